"""Extracts a per-kernel summary (JSON) from an ncu report: python tools/ncu_summary.py <rep> <out.json> [kernel-regex]"""
import csv, json, re, subprocess, sys
rep, out = sys.argv[1], sys.argv[2]
pat = re.compile(sys.argv[3]) if len(sys.argv) > 3 else None
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
hdr, units = rows[0], rows[1]
keep = ["Kernel Name", "Block Size", "Grid Size", "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
        "l1tex__data_pipe_lsu_wavefronts.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum", "launch__registers_per_thread",
        "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem", "launch__shared_mem_per_block_dynamic",
        "smsp__pcsamp_warps_issue_stalled_long_scoreboard", "smsp__pcsamp_warps_issue_stalled_short_scoreboard", "smsp__pcsamp_warps_issue_stalled_barrier",
        "smsp__pcsamp_warps_issue_stalled_mio_throttle", "smsp__pcsamp_warps_issue_stalled_math_pipe_throttle", "smsp__pcsamp_warps_issue_stalled_not_selected",
        "smsp__pcsamp_warps_issue_stalled_selected", "smsp__pcsamp_warps_issue_stalled_wait", "smsp__pcsamp_sample_count"]
idx = {h: i for i, h in enumerate(hdr)}
res = []
for r in rows[2:]:
    name = r[idx["Kernel Name"]]
    if pat and not pat.search(name):
        continue
    res.append({k: {"value": r[idx[k]], "unit": units[idx[k]]} for k in keep if k in idx})
json.dump({"report": rep, "how": "ncu --set full --clock-control none --import-source on (one launch each, cold cache, serialised)", "kernels": res}, open(out, "w"), indent=1)
print(f"{len(res)} kernels -> {out}")
