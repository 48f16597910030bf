"""Turns the two ncu artifacts of a round into the text summaries kept under profiles/.
  python scripts/summarize_ncu.py launches gpurun_out/r1e_launches.csv "<command that was profiled>" > profiles/r1e_launch_list_summary.txt
  python scripts/summarize_ncu.py full gpurun_out/r1e_top.ncu-rep > profiles/r1e_top_kernels_summary.txt"""
import csv
import subprocess
import sys
from collections import OrderedDict

RAW_KEYS = [
    "gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread", "launch__shared_mem_per_block_dynamic", "launch__shared_mem_per_block_static",
    "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_bytes.sum", "l1tex__t_bytes.sum",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum", "smsp__thread_inst_executed.sum",
    "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__average_warp_latency_per_inst_issued.ratio",
    "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio", "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio", "smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio", "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio", "smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio",
]


def launches(path, command):
    rows = [r for r in csv.reader(l for l in open(path) if l.startswith('"'))]
    hdr = rows[0]
    ik, iv = hdr.index("Kernel Name"), hdr.index("Metric Value")
    agg = OrderedDict()
    for r in rows[1:]:
        name = r[ik]
        d = agg.setdefault(name, [0, 0.0])
        d[0] += 1
        d[1] += float(r[iv].replace(",", "")) / 1e3
    total = sum(v[1] for v in agg.values())
    print("# ncu --metrics gpu__time_duration.sum --clock-control none  %s" % command)
    print("# %d launches captured; per-launch times under ncu are cold-cache and serialised: compare SHARES with bench.py's own per-kernel CUDA-event split" % sum(v[0] for v in agg.values()))
    for name, (n, us) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print("%-72s launches %4d  total %10.1f us  avg %9.1f us  share %5.1f%%" % (name[:72], n, us, us / n, 100.0 * us / total))


def full(path):
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr = rows[0]
    for r in rows[2:]:
        name = r[hdr.index("Kernel Name")]
        print("== %s" % name)
        for k in RAW_KEYS:
            if k in hdr:
                print("   %-90s %s %s" % (k, r[hdr.index(k)], rows[1][hdr.index(k)]))
        if "dram__bytes_read.sum" in hdr:
            def num(k):
                v, u = float(r[hdr.index(k)].replace(",", "")), rows[1][hdr.index(k)]
                return v * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(u, 1)
            print("   traffic (dram read + write) per launch: %.1f MB" % ((num("dram__bytes_read.sum") + num("dram__bytes_write.sum")) / 1e6))


if __name__ == "__main__":
    if sys.argv[1] == "launches":
        launches(sys.argv[2], sys.argv[3] if len(sys.argv) > 3 else "")
    else:
        full(sys.argv[2])
