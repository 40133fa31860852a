#!/usr/bin/env python
"""Summarise an .ncu-rep (one `ncu --set full` capture) into the handful of metrics DESIGN.md / bench.py quote.

    python profiles/summarize_ncu.py gpurun_out/prof.ncu-rep > profiles/rNN_<kernel>.txt
"""
import csv
import subprocess
import sys

KEYS = [
    "gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
    "launch__occupancy_limit_shared_mem", "sm__cycles_elapsed.avg.per_second",
    "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed", "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
    "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
    "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_tensor.avg.pct_of_peak_sustained_active",
    "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed",
    "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
]
STALLS = "smsp__average_warps_issue_stalled_"


def main(path):
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units = rows[0], rows[1]
    for r in rows[2:]:
        d = dict(zip(hdr, r))
        print("kernel:", d.get("Kernel Name"), "| id", d.get("ID"))
        for k in KEYS:
            for h, u in zip(hdr, units):
                if h == k:
                    print("  %-80s %s %s" % (k, d[h], u))
        st = sorted(((float(d[h]), h[len(STALLS):]) for h in hdr if h.startswith(STALLS) and h.endswith("_per_issue_active.ratio")),
                    reverse=True)
        print("  stalls per issue:", ", ".join("%s %.2f" % (n.replace("_per_issue_active.ratio", ""), v) for v, n in st[:6]))
        tr = None
        try:
            f = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
            ur = dict(zip(hdr, units))
            tr = float(d["dram__bytes_read.sum"]) * f[ur["dram__bytes_read.sum"]] + \
                float(d["dram__bytes_write.sum"]) * f[ur["dram__bytes_write.sum"]]
        except Exception:
            pass
        print("  dram traffic per launch (bytes):", tr)


if __name__ == "__main__":
    main(sys.argv[1])
