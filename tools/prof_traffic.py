"""One launch of each streaming kernel at BASELINE size (for one `ncu --set full` capture: DRAM bytes per launch).

    ncu --set full --clock-control none --import-source on \
        -k regex:"mdp_step_kernel|taxel_kernel|ppo_loss_kernel|ppo_heads_kernel|wgrad_splitk_kernel|device_kernel|mlp3_forward_kernel" -c 30 -o out \
        python tools/prof_traffic.py        # then tools/ncu_summary.py
"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from locotouch_b200 import ops  # noqa: E402
from locotouch_b200.mdp import task_spec as TS  # noqa: E402
from locotouch_b200.mdp.fused import FusedMdp  # noqa: E402
from locotouch_b200.sim import synth  # noqa: E402

n = 4096
spec = TS.SPECS["teacher"]()
env = synth.make_env(n, seed=1, with_object=True).to("cuda")
m = FusedMdp(env, spec, seed=1)
for _ in range(3):  # launches 0-2: the fused teacher step (the first fills the history)
    m.step(True, True)
g = torch.Generator().manual_seed(0)
q = torch.randn(n, 238, 4, generator=g)
q = (q / q.norm(dim=-1, keepdim=True)).cuda()
f = (torch.randn(n, 221, 3, generator=g) * 0.1).cuda()
thr = (0.05 + (torch.rand(n, 221, generator=g) - 0.5) * 0.02).cuda()
for i in range(3):
    ops.taxel_synth(q, f, thr, quat_body_offset=17, seed=1, offset=i)
b, A = n * 24 // 4, 12  # launches 6-8: the fused PPO loss on one mini-batch
rn = lambda *s: torch.randn(*s, device="cuda")  # noqa: E731
args = dict(mu=rn(b, A), sigma=0.5 + torch.rand(A, device="cuda"), value=rn(b), actions=rn(b, A), old_logp=rn(b), old_mu=rn(b, A),
            old_sigma=0.5 + torch.rand(b, A, device="cuda"), advantages=rn(b), returns=rn(b), old_values=rn(b))
lr = torch.tensor([1e-3], device="cuda")
for _ in range(3):
    ops.ppo_loss(**args, entropy_coef=0.01, desired_kl=0.01, lr=lr)
# launches 9-11: K16 (heads + loss + head dgrad) on one mini-batch
H = 128
bufs = ops.PpoLossBuffers(b, A, "cuda")
hs = [rn(b, H) for _ in range(2)]
gh = [torch.empty(b, H, device="cuda") for _ in range(2)]
for _ in range(3):
    ops.ppo_heads_loss(hs[0], hs[1], rn(A, H) * 0.1, rn(A), rn(1, H) * 0.1, rn(1), args["sigma"], args["actions"], args["old_logp"], args["old_mu"],
                       args["old_sigma"], args["advantages"], args["returns"], args["old_values"], gh[0], gh[1], entropy_coef=0.01, desired_kl=0.01, lr=lr,
                       buffers=bufs)
# launches 12-17: K15 weight + bias gradient of the first two layers; 18-20: K12 forward of the first layer
torch.backends.cuda.matmul.allow_tf32 = True
for (no, k) in ((512, 348), (256, 512)):
    g_, x_, dw, db = rn(b, no), rn(b, k), torch.zeros(no, k, device="cuda"), torch.zeros(no, device="cuda")
    for _ in range(3):
        ops.wgrad(g_, x_, dw, db, zero_first=False)
x_, w_, b_ = rn(b, 348), rn(512, 348) / 18.0, rn(512)
o_ = torch.empty(b, 512, device="cuda")
for _ in range(3):
    ops.linear_bias_act(x_, w_, b_, out=o_, elu=True)
# K19 (three hidden layers of both MLPs in one persistent kernel): 3 launches at the mini-batch size (h1 / h2 stored), 3 at the rollout size
for rows_, keep in ((b, True), (n, False)):
    nets = []
    for _ in range(2):
        ps = []
        for (no, k) in ((512, 348), (256, 512), (128, 256)):
            ps += [rn(no, k) / k ** 0.5, rn(no)]
        nets.append((rn(rows_, 348), tuple(ps), (torch.empty(rows_, 512, device="cuda") if keep else None, torch.empty(rows_, 256, device="cuda") if keep else None,
                                                  torch.empty(rows_, 128, device="cuda"))))
    for _ in range(3):
        ops.mlp3_forward(nets)
# K15 pair launch (layer 1 of actor and critic)
g2, x2 = [rn(b, 512) for _ in range(2)], [rn(b, 348) for _ in range(2)]
dw2, db2 = [torch.zeros(512, 348, device="cuda") for _ in range(2)], [torch.zeros(512, device="cuda") for _ in range(2)]
for _ in range(3):
    ops.wgrad_pair(g2[0], x2[0], dw2[0], db2[0], g2[1], x2[1], dw2[1], db2[1])
torch.cuda.synchronize()
print("ok")
