"""Condenses an `ncu --metrics gpu__time_duration.sum --csv` launch list into a per-kernel share table.

    python tools/summarize_launches.py gpurun_out/launches_r1.csv [--last-iteration] > profiles/launches_r1_summary.md

--last-iteration keeps only the launches after the second-to-last `counter_add_kernel` boundary (= the replayed, timed PPO
iteration: one counter_add per rollout).  ncu serialises launches and runs them cold-cache: compare SHARES, not absolutes."""
import csv
import re
import sys
from collections import defaultdict

path = sys.argv[1]
rows = []
with open(path, newline="") as f:
    lines = [ln for ln in f if not ln.startswith("==")]
reader = csv.reader(lines)
hdr = next(reader)
ki, vi, ui = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
for r in reader:
    if len(r) <= vi:
        continue
    v = float(r[vi].replace(",", ""))
    unit = r[ui]
    us = v / 1e3 if unit in ("ns", "nsecond") else (v if unit.startswith("u") else v * 1e3)
    rows.append((r[ki], us))
if "--last-iteration" in sys.argv:
    marks = [i for i, (k, _) in enumerate(rows) if "counter_add" in k]
    if len(marks) >= 2:
        rows = rows[marks[-2] + 1:]
        # the rollout graph ends with counter_add; the iteration = [rollout ... counter_add] + [GAE + update]: rotate
def short(name):
    name = re.sub(r"<unnamed>::", "", name)
    name = re.sub(r"\(.*", "", name)
    name = re.sub(r"void ", "", name)
    return name[:70]
agg = defaultdict(lambda: [0, 0.0])
for k, us in rows:
    a = agg[short(k)]
    a[0] += 1
    a[1] += us
total = sum(a[1] for a in agg.values())
mine = ("mdp_step", "taxel", "gae_scan", "adv_normalize", "ppo_loss", "clip_adam", "grad_sqnorm", "act_sample", "store_scalars", "gather_rows",
        "process_actions", "counter_add", "bias_act_bwd", "copy4", "copy1", "any_nonzero", "adaptive_lr", "masked_mse", "pad_traj", "tactile_delay",
        "command_step", "vel_curriculum", "lt_wgrad::", "lt_mlp3::", "act_heads", "ppo_heads", "gae_fused", "peer_sum", "contact_time", "student_", "policy_",
        # K12 (gemm_fused.cu): our translation unit built from the CUTLASS sm100 collective builders; cuBLAS's own kernels are
        # named cutlass3x_* / cutlass::Kernel2<cutlass_80_*> and are NOT ours
        "cutlass::device_kernel<")
print(f"launches: {len(rows)}   summed kernel time: {total / 1e3:.3f} ms (ncu: serialised, cold cache)\n")
print("| kernel | launches | total us | share | avg us | ours |")
print("|---|---:|---:|---:|---:|:-:|")
for k, (n, us) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    ours = "x" if any(m in k for m in mine) else ""
    print(f"| `{k}` | {n} | {us:.1f} | {100 * us / total:.1f}% | {us / n:.2f} | {ours} |")
own = sum(us for k, (n, us) in agg.items() if any(m in k for m in mine))
print(f"\nkernels of this library (hand-written + K12 fused tcgen05 GEMMs): {100 * own / total:.1f}% of the summed kernel time; "
      f"cuBLAS / ATen (rollout head layers, fills, randperm, copies): {100 - 100 * own / total:.1f}%")
