"""K19 (fused three-layer MLP forward, both networks) vs the six per-layer K12 launches it replaces: python tools/mlp3_bench.py"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from locotouch_b200 import ops  # noqa: E402
from tools.linear_bench import timed  # noqa: E402

torch.backends.cuda.matmul.allow_tf32 = True

for B, keep in ((4096, False), (24576, True)):
    nets, sets = [], 6  # several input sets so that x comes from HBM like in the update
    for s in range(sets):
        pair = []
        for _ in range(2):
            x = torch.randn(B, 348, device="cuda")
            ps = []
            for (n, k) in ((512, 348), (256, 512), (128, 256)):
                ps += [torch.randn(n, k, device="cuda") / k ** 0.5, torch.randn(n, device="cuda")]
            hs = (torch.empty(B, 512, device="cuda") if keep else None, torch.empty(B, 256, device="cuda") if keep else None, torch.empty(B, 128, device="cuda"))
            pair.append((x, tuple(ps), hs))
        nets.append(pair)
    bufs = [[torch.empty(B, n, device="cuda") for n in (512, 256, 128)] for _ in range(2)]
    i = [0]

    def fused():
        i[0] = (i[0] + 1) % sets
        assert ops.mlp3_forward(nets[i[0]]) is not None

    def per_layer():
        i[0] = (i[0] + 1) % sets
        for (x, ps, _), hb in zip(nets[i[0]], bufs):
            h = x
            for w, b, o in zip(ps[0::2], ps[1::2], hb):
                h = ops.linear_bias_act(h, w, b, out=o, elu=True)

    t0, t1 = timed(per_layer, reps=24), timed(fused, reps=24)
    flop = 2 * 2 * B * (348 * 512 + 512 * 256 + 256 * 128)
    print(f"B = {B:6d} (h1/h2 stored: {keep}): 6 x K12 {t0:7.1f} us   K19 fused {t1:7.1f} us   ({flop / t1 / 1e6:6.1f} TFLOP/s TF32)")
