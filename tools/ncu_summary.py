"""Condense an `ncu -i X.ncu-rep --page raw --csv` dump of ONE kernel launch into the small JSON bench.py's `roofline`
object reads (profiles/roofline_kernel.json): DRAM traffic per launch and both tensor-pipe readings.

    ncu -i gpurun_out/prof.ncu-rep --page raw --csv > profiles/rNN_ncu_<kernel>.csv
    python tools/ncu_summary.py profiles/rNN_ncu_<kernel>.csv profiles/roofline_kernel.json "<workload note>"
"""
import csv
import json
import sys

src, dst = sys.argv[1], sys.argv[2]
note = sys.argv[3] if len(sys.argv) > 3 else ""
rows = list(csv.reader(open(src)))
hdr, units, vals = rows[0], rows[1], rows[2]


def get(name, scale_units=True):
    for i, h in enumerate(hdr):
        if h == name:
            try:
                v = float(vals[i].replace(",", ""))
            except ValueError:          # "no data": the counter was not collected in this capture
                return None
            u = units[i]
            if scale_units:
                v *= {"Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "byte": 1.0}.get(u, 1.0)
            return v
    return None


kernel = vals[hdr.index("Kernel Name")]
rd, wr = get("dram__bytes_read.sum"), get("dram__bytes_write.sum")
out = {
    "kernel": kernel,
    "source": f"{src} (ncu --set full --clock-control none, one launch{'; ' + note if note else ''})",
    "dram_bytes_read": rd, "dram_bytes_write": wr, "dram_bytes": (rd or 0) + (wr or 0),
    "duration_us_under_ncu": get("gpu__time_duration.sum", False),
    "sm_cycles_elapsed": get("sm__cycles_elapsed.avg", False),
    # the two readings of the tensor pipe: ncu's normalised figure, and the raw counter, which is a SUM
    # over the 4 sub-pipes of an SM (divide by 4 * sm_cycles_elapsed to get the same fraction)
    "pipe_tensor_cycles_active_pct_of_peak_sustained_elapsed":
        get("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed", False),
    "hmma_cycles_active_realtime_sum_over_subpipes":
        get("TPC.TriageCompute.sm__pipe_tensor_subpipe_hmma_cycles_active_realtime.avg", False),
    "l1tex_throughput_pct": get("l1tex__throughput.avg.pct_of_peak_sustained_elapsed", False),
    "lts_throughput_pct": get("lts__throughput.avg.pct_of_peak_sustained_elapsed", False),
    "dram_throughput_pct": get("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", False),
    "registers_per_thread": get("launch__registers_per_thread", False),
}
json.dump(out, open(dst, "w"), indent=1)
print(json.dumps(out, indent=1))
