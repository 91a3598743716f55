"""In-kernel timeline of the fused MDP kernel (profiling build only).

    LT_MDP_PROF=1 python -m locotouch_b200.csrc.build --force && python tools/mdp_timeline.py [teacher|locomotion] [envs]

Every warp stamps clock64 at the stage boundaries; this prints, per warp (== task slot), the mean over blocks of each stamp
relative to the earliest stamp of its block, in microseconds at the sampled SM clock.
"""
import ctypes
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from locotouch_b200 import _C  # noqa: E402
from locotouch_b200.mdp import task_spec as TS  # noqa: E402
from locotouch_b200.mdp.fused import FusedMdp  # noqa: E402
from locotouch_b200.sim import synth  # noqa: E402

task = sys.argv[1] if len(sys.argv) > 1 else "teacher"
n = int(sys.argv[2]) if len(sys.argv) > 2 else 4096
phases = sys.argv[3] if len(sys.argv) > 3 else "both"
lib = _C.lib()
fn = getattr(lib, "lt_debug_mdp_prof", None)
if fn is None:
    raise SystemExit("library was not built with LT_MDP_PROF=1")
fn.restype = ctypes.c_int
fn.argtypes = [ctypes.c_void_p]
spec = TS.SPECS[task]()
copies = 8
mdps = []
for i in range(copies):
    env = synth.make_env(n, seed=i, with_object=spec.with_object).to("cuda")
    m = FusedMdp(env, spec, seed=i)
    m.step(True, True)
    mdps.append(m)
blocks = (n + 31) // 32
WARPS = int(os.environ.get("LT_MDP_THREADS", "1024")) // 32
prof = torch.zeros(blocks, WARPS, 32, dtype=torch.int64, device="cuda")
torch.cuda.synchronize()


def run(m):
    if phases == "both":
        m.step(True, True)
    elif phases == "rew":
        m.compute_rewards()
    else:
        m.compute_observations()


for i in range(copies):  # warm
    run(mdps[i])
torch.cuda.synchronize()
acc = torch.zeros(WARPS, 32, dtype=torch.float64)
cnt = 0
assert fn(prof.data_ptr()) == 0
for rep in range(3):
    for i in range(copies):
        prof.zero_()
        run(mdps[i])
        torch.cuda.synchronize()
        p = prof.cpu().double()
        t0 = p[:, :, 0].min(dim=1).values  # block start
        rel = p - t0[:, None, None]
        rel[p == 0] = float("nan")
        acc += torch.nan_to_num(rel.nanmean(dim=0), nan=-1.0)
        cnt += 1
assert fn(0) == 0
mhz = 1965.0
mean = acc / cnt / mhz
names = {16: "bulk_iss", 17: "gathers", 18: "nonfloat", 21: "pre_done", 19: "2a_loop", 20: "2a_sum", 0: "entry", 1: "issued", 2: "landed", 3: "sync1", 8: "round0", 10: "round1", 13: "pass0_end", 12: "pass1", 4: "tasks_done", 5: "hist_landed", 6: "sync2", 7: "2a_done", 15: "end"}
order = [0, 16, 17, 18, 1, 2, 3, 21, 8, 10, 4, 5, 6, 19, 20, 7, 15]
print(f"task={task} envs={n} phases={phases}; microseconds since block start (mean over {blocks} blocks, {cnt} launches)")
print("warp " + " ".join(f"{names[s]:>11s}" for s in order))
for w in range(WARPS):
    print(f"{w:4d} " + " ".join(f"{mean[w, s].item():11.2f}" for s in order))
