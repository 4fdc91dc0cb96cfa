mkdir -p gpurun_out/r5d
O=gpurun_out/r5d
timeout 600 python bench.py > $O/bench_full.json 2> $O/bench_full.err
echo "bench rc=$?"; cut -c1-200 $O/bench_full.json
for tool in memcheck racecheck; do
  timeout 170 compute-sanitizer --tool $tool --print-limit 30 python tools/sanitize_target.py > $O/sanitizer_$tool.log 2>&1
  echo "$tool rc=$?"; tail -2 $O/sanitizer_$tool.log | cut -c1-200
done
