"""ncu CSV (per-launch dram bytes / duration / throughput) of tools/hbm_bench.py -> markdown table, last launch per kernel signature.
Usage: ncu_hbm_table.py in.csv out.md"""
import collections, csv, sys
src, dst = sys.argv[1], sys.argv[2]
lines = [l for l in open(src) if l.startswith('"')]
rows = list(csv.DictReader(lines))
per = collections.OrderedDict()
for r in rows:
    key = (r["ID"], r["Kernel Name"], r["Grid Size"])
    per.setdefault(key, {})[r["Metric Name"]] = (float(r["Metric Value"].replace(",", "")), r["Metric Unit"])
def mb(v):
    val, unit = v
    return val * {"byte": 1e-6, "Kbyte": 1e-3, "Mbyte": 1.0, "Gbyte": 1e3}.get(unit, 1e-6)
def us(v):
    val, unit = v
    return val * {"ns": 1e-3, "us": 1.0, "ms": 1e3, "nsecond": 1e-3, "usecond": 1.0, "msecond": 1e3}.get(unit, 1e-3)
seen = collections.OrderedDict()
for (_, name, grid), m in per.items():
    if "at::" in name or "dram__bytes_read.sum" not in m:
        continue
    rd, wr, t = mb(m["dram__bytes_read.sum"]), mb(m["dram__bytes_write.sum"]), us(m["gpu__time_duration.sum"])
    if rd + wr < 1.0:
        continue
    short = name.replace("void ", "").replace("b3d::", "")
    short = short[:short.index("(")] if "(" in short else short
    seen.setdefault((short, grid), []).append((t, rd, wr))
with open(dst, "w") as f:
    f.write("# ncu DRAM counters of the bandwidth-bound kernels (`ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,"
            "gpu__time_duration.sum --clock-control none python tools/hbm_bench.py 1`; median launch per kernel and grid; "
            "cold cache, serialised)\n\n| kernel | grid | launches | us | dram read MB | dram write MB | DRAM GB/s | % of 6,547 |\n"
            "|---|---|---:|---:|---:|---:|---:|---:|\n")
    for (short, grid), v in seen.items():
        v.sort()
        t, rd, wr = v[len(v) // 2]
        gbs = (rd + wr) / t * 1e3
        f.write(f"| `{short}` | {grid} | {len(v)} | {t:.1f} | {rd:.1f} | {wr:.1f} | {gbs:.0f} | {100 * gbs / 6546.6:.1f} |\n")
print(open(dst).read()[:3000])
