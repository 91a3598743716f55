"""Where does K15 (lt_wgrad_splitk) differ from the fp64 product?  Prints, per shape, the maximum error and a map of the
32 x 32 output blocks whose error exceeds the TF32 bound.

    python tools/wgrad_debug.py
"""
from __future__ import annotations

import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from locotouch_b200 import ops  # noqa: E402


def main():
    dev = torch.device("cuda:0")
    shapes = [(32, 128, 32), (64, 128, 64), (256, 128, 128), (4096, 128, 256), (4096, 256, 256), (4096, 512, 348), (24576, 512, 348), (24576, 256, 512),
              (24576, 128, 256), (1000, 132, 100), (37, 20, 64), (4096, 64, 36), (8192, 512, 512), (333, 20, 12), (2048, 640, 600), (5000, 256, 272), (24576, 12, 128), (24576, 1, 128)]
    for B, n, k in shapes:
        gen = torch.Generator().manual_seed(B + n + k)
        g = (torch.randn(B, n, generator=gen) / B ** 0.5).to(dev)
        x = torch.randn(B, k, generator=gen).to(dev)
        out = torch.full((n, k), 7.0, device=dev)
        db = torch.full((n,), 5.0, device=dev)
        res = ops.wgrad(g, x, out, db)
        torch.cuda.synchronize()
        if res is None:
            print(f"B={B} n={n} k={k}: unsupported")
            continue
        ref = g.double().t() @ x.double()
        err = (out.double() - ref).abs()
        errb = (db.double() - g.double().sum(0)).abs().max().item()
        print(f"B={B} n={n} k={k}: max err {err.max().item():.3e}  bias err {errb:.3e}  (ref rms {ref.pow(2).mean().sqrt().item():.3f})")
        if err.max().item() > 4e-3:
            bad = err > 4e-3
            for r0 in range(0, n, 32):
                print("   rows %4d: " % r0 + "".join("X" if bad[r0:r0 + 32, c0:c0 + 32].any().item() else "." for c0 in range(0, k, 32)))
            # is it a multiple of the reference (double add / missing slices)?
            ratio = (out.double() / ref)[bad]
            print("   out/ref over the bad entries: median %.3f  min %.3f  max %.3f" % (ratio.median().item(), ratio.min().item(), ratio.max().item()))


if __name__ == "__main__":
    main()
