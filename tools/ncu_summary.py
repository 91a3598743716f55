"""Condenses an `ncu --set full` report into the JSON summary committed under profiles/ (and read by bench.py for roofline.traffic).

    ncu --set full --clock-control none --import-source on -k regex:"mdp_step_kernel|taxel_kernel|ppo_loss_kernel" -c 9 \
        -o gpurun_out/r1c_full python tools/prof_traffic.py
    ncu -i gpurun_out/r1c_full.ncu-rep --page raw --csv > gpurun_out/r1c_full_raw.csv
    python tools/ncu_summary.py gpurun_out/r1c_full_raw.csv > profiles/r1c_k1_k2_ncu_full_summary.json
"""
import csv
import json
import sys

KEEP = (
    "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "lts__t_bytes.sum", "smsp__inst_executed.sum",
    "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "launch__registers_per_thread", "launch__grid_size", "launch__block_size", "launch__occupancy_limit_shared_mem",
    "launch__occupancy_limit_registers",
    "smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_imc_miss_per_issue_active.ratio",
    "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_tensor.sum",
    "lts__t_sector_hit_rate.pct", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
)

lines = [ln for ln in open(sys.argv[1], newline="") if not ln.startswith("==")]
rows = list(csv.reader(lines))
header, units = rows[0], rows[1]
out = []
for r in rows[2:]:
    if len(r) != len(header):
        continue
    rec = {"kernel": r[header.index("Kernel Name")][:60]}
    for key in KEEP:
        if key in header:
            i = header.index(key)
            rec[key] = f"{r[i]} {units[i]}".strip()
    out.append(rec)
json.dump(out, sys.stdout, indent=1)
