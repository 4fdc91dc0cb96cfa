"""Aggregate an `ncu --metrics gpu__time_duration.sum --csv` launch list into a per-kernel table for the LAST
training step in the log (delimited by adam_kernel launches).  Usage: summarize_launches.py in.csv out.md [title]"""
import collections
import csv
import re
import sys

src, dst = sys.argv[1], sys.argv[2]
title = sys.argv[3] if len(sys.argv) > 3 else src
with open(src) as f:
    lines = [l for l in f if not l.startswith("==")]
rows = list(csv.DictReader(lines))
names = [r["Kernel Name"] for r in rows]
vals = [float(r["Metric Value"].replace(",", "")) for r in rows]
idx = [i for i, n in enumerate(names) if "adam_kernel" in n]
s, e = (idx[-2] + 2, idx[-1] + 2) if len(idx) >= 2 else (0, len(rows))
agg = collections.defaultdict(lambda: [0, 0.0])
for n, v in zip(names[s:e], vals[s:e]):
    k = re.sub(r"[<(].*", "", n).replace("void ", "")
    if "conv_tc_kernel" in n:
        m = re.search(r"TcCfg<(\d+), *\d+, *\d+, *(\d+), *(\d+), *(\w+)(?:, *(\d+))?>", n)
        k = f"b3d::conv_tc_kernel KS={m.group(2)}" + (" (stride-2 family)" if m.group(2) == "2" else "") + \
            (" kd-folded" if m.group(5) == "1" else "")
    elif "wgrad_tc" in n:
        m = re.search(r"<(\d+), *(\d+)>", n)
        k = f"b3d::conv3_wgrad_tc_kernel KS={m.group(1)} TG={m.group(2)}"
    elif k.startswith("at::"):
        k = "at:: (torch elementwise: autograd grad accumulation, zeros, stack)"
    agg[k][0] += 1
    agg[k][1] += v
tot = sum(v for _, v in agg.values())
with open(dst, "w") as f:
    f.write(f"# {title}\n\nOne training step (default model, 128^3, batch 1), eager launches under "
            f"`ncu --metrics gpu__time_duration.sum --clock-control none` (cold-cache, serialised: compare SHARES).\n\n"
            f"launches: {e - s}, sum of kernel durations: {tot * 1e-6:.2f} ms\n\n| kernel | launches | ms | share |\n|---|---:|---:|---:|\n")
    for k, (c, v) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        f.write(f"| `{k}` | {c} | {v * 1e-6:.3f} | {100 * v / tot:.1f}% |\n")
print(open(dst).read()[:1500])
