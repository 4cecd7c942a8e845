"""Turns an `ncu --set full` report into the text + JSON summaries kept under profiles/.
usage: python scripts/summarise_ncu.py gpurun_out/prof.ncu-rep profiles/r1_hist_ncu.txt [profiles/hist_root_traffic.json]"""
import csv
import json
import subprocess
import sys

rep, out_txt = sys.argv[1], sys.argv[2]
out_json = sys.argv[3] if len(sys.argv) > 3 else None
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
hdr, units = rows[0], rows[1]
want = ["Kernel Name", "launch__grid_size", "launch__block_size", "launch__registers_per_thread", "gpu__time_duration.sum",
        "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active", "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "l1tex__lsu_writeback_active_mem_lg.avg.pct_of_peak_sustained_elapsed",
        "l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed", "smsp__pcsamp_warps_issue_stalled_mio_throttle",
        "smsp__pcsamp_warps_issue_stalled_lg_throttle", "smsp__pcsamp_warps_issue_stalled_math_pipe_throttle",
        "smsp__pcsamp_warps_issue_stalled_not_selected", "smsp__pcsamp_warps_issue_stalled_dispatch_stall",
        "smsp__pcsamp_warps_issue_stalled_wait", "smsp__pcsamp_warps_issue_stalled_selected",
        "smsp__pcsamp_warps_issue_stalled_long_scoreboard", "smsp__pcsamp_warps_issue_stalled_short_scoreboard",
        "smsp__pcsamp_warps_issue_stalled_barrier", "smsp__pcsamp_warps_issue_stalled_branch_resolving"]
idx = [hdr.index(w) for w in want if w in hdr]
lines = [f"source report: {rep} (ncu --set full --clock-control none --import-source on)"]
first_root = None
for r in rows[2:]:
    lines.append("-" * 100)
    d = {}
    for i in idx:
        lines.append(f"{hdr[i]:70s} {r[i]} {units[i]}")
        d[hdr[i]] = r[i]
    if first_root is None and "k_hist_root" in d["Kernel Name"]:
        first_root = d
open(out_txt, "w").write("\n".join(lines) + "\n")
if out_json and first_root:
    def to_bytes(v, u):
        return float(v) * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}[u]
    ur, uw = units[hdr.index("dram__bytes_read.sum")], units[hdr.index("dram__bytes_write.sum")]
    tr = to_bytes(first_root["dram__bytes_read.sum"], ur) + to_bytes(first_root["dram__bytes_write.sum"], uw)
    json.dump({"kernel": first_root["Kernel Name"], "dram_bytes_per_launch": tr, "duration_us_under_ncu": float(first_root["gpu__time_duration.sum"]),
               "source": rep}, open(out_json, "w"), indent=1)
print("\n".join(lines[:30]))
