"""Launches each hand-written kernel a few times at BASELINE size (for `ncu`): python tools/prof_kernels.py [mdp|taxel|ppo|all]"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from locotouch_b200 import ops  # noqa: E402
from locotouch_b200.mdp import task_spec as TS  # noqa: E402
from locotouch_b200.mdp.fused import FusedMdp  # noqa: E402
from locotouch_b200.sim import synth  # noqa: E402

what = sys.argv[1] if len(sys.argv) > 1 else "all"
n = int(sys.argv[2]) if len(sys.argv) > 2 else 4096
reps = 6
if what in ("mdp", "all"):
    for task in ("locomotion", "teacher"):
        spec = TS.SPECS[task]()
        env = synth.make_env(n, seed=1, with_object=spec.with_object).to("cuda")
        m = FusedMdp(env, spec, seed=1)
        for _ in range(reps):
            m.step(True, True)
        for _ in range(3):
            m.compute_rewards()
            m.compute_observations()
if what in ("taxel", "all"):
    g = torch.Generator().manual_seed(0)
    q = torch.randn(n, 238, 4, generator=g)
    q = (q / q.norm(dim=-1, keepdim=True)).cuda()
    f = (torch.randn(n, 221, 3, generator=g) * 0.1).cuda()
    thr = (0.05 + (torch.rand(n, 221, generator=g) - 0.5) * 0.02).cuda()
    for i in range(reps):
        ops.taxel_synth(q, f, thr, quat_body_offset=17, seed=1, offset=i)
if what in ("ppo", "all"):
    T = 24
    r, v = torch.randn(T, n, device="cuda"), torch.randn(T, n, device="cuda")
    d = (torch.rand(T, n, device="cuda") < 0.02).byte()
    for _ in range(reps):
        ops.gae(r, v, d, torch.randn(n, device="cuda"), 0.99, 0.95)
    b, A = n * T // 4, 12
    rn = lambda *s: torch.randn(*s, device="cuda")  # noqa: E731
    args = dict(mu=rn(b, A), sigma=0.5 + torch.rand(A, device="cuda"), value=rn(b), actions=rn(b, A), old_logp=rn(b), old_mu=rn(b, A),
                old_sigma=0.5 + torch.rand(b, A, device="cuda"), advantages=rn(b), returns=rn(b), old_values=rn(b))
    lr = torch.tensor([1e-3], device="cuda")
    for _ in range(reps):
        ops.ppo_loss(**args, entropy_coef=0.01, desired_kl=0.01, lr=lr)
    p, gr, m1, m2 = (torch.randn(607644, device="cuda") for _ in range(4))
    step = torch.zeros(1, device="cuda")
    for _ in range(reps):
        ops.clip_adam(p, gr, m1.abs(), m2.abs(), lr, step)
torch.cuda.synchronize()
print("ok")
