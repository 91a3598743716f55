"""K15 pair launch repeated: worst deviation of dW / db from the float64 products.   python tools/wgrad_pair_stress.py"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from locotouch_b200 import ops  # noqa: E402

B = 24576
for (n, k) in ((512, 348), (256, 512), (128, 256), (512, 352), (512, 320)):
    gen = torch.Generator().manual_seed(n + k)
    gs = [(torch.randn(B, n, generator=gen) / B ** 0.5).cuda() for _ in range(2)]
    xs = [torch.randn(B, k, generator=gen).cuda() for _ in range(2)]
    refw = [g.double().t() @ x.double() for g, x in zip(gs, xs)]
    refb = [g.double().sum(0) for g in gs]
    worst_w = worst_b = worst_b1 = 0.0
    for rep in range(30):
        outs = [torch.zeros(n, k, device="cuda") for _ in range(2)]
        dbs = [torch.zeros(n, device="cuda") for _ in range(2)]
        ops.wgrad_pair(gs[0], xs[0], outs[0], dbs[0], gs[1], xs[1], outs[1], dbs[1])
        torch.cuda.synchronize()
        for o, d, rw, rb in zip(outs, dbs, refw, refb):
            worst_w = max(worst_w, (o.double() - rw).abs().max().item())
            worst_b = max(worst_b, (d.double() - rb).abs().max().item())
        o1, d1 = torch.zeros(n, k, device="cuda"), torch.zeros(n, device="cuda")
        ops.wgrad(gs[0], xs[0], o1, d1, zero_first=False)
        torch.cuda.synchronize()
        worst_b1 = max(worst_b1, (d1.double() - refb[0]).abs().max().item())
    print(f"{n}x{k}: pair worst |dW err| {worst_w:.2e}  |db err| {worst_b:.2e}   single launch |db err| {worst_b1:.2e}")
