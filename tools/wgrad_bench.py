"""K15 (lt_wgrad_splitk) against the round-1 path (8-way torch.bmm split + torch.sum) and one cuBLAS mm, per layer shape of the PPO
mini-batch.  Same harness as tools/kbench.py: launches captured in a CUDA graph, input sets cycled through more than L2.

    python tools/wgrad_bench.py [--batch 24576]
"""
from __future__ import annotations

import argparse
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tools"))

from kbench import time_graph  # noqa: E402
from locotouch_b200 import ops  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=24576)
    ap.add_argument("--no-zero", action="store_true", help="K15 without its memset nodes (accumulate mode)")
    args = ap.parse_args()
    B = args.batch
    torch.backends.cuda.matmul.allow_tf32 = True
    print(f"{'shape (n x k)':<16}{'K15 us':>10}{'TFLOP/s':>10}{'GB/s (min bytes)':>18}{'bmm8+sum us':>14}{'mm us':>10}")
    for n, k in [(512, 348), (256, 512), (128, 256), (12, 128), (1, 128), (512, 272)]:
        copies = 4
        sets = [(torch.randn(B, n, device="cuda"), torch.randn(B, k, device="cuda"), torch.zeros(n, k, device="cuda"),
                 torch.empty(8, n, k, device="cuda")) for _ in range(copies)]
        db = torch.zeros(n, device="cuda")

        def k15(i):
            g, x, o, _ = sets[i]
            ops.wgrad(g, x, o, db, zero_first=not args.no_zero)

        def bmm8(i):
            g, x, o, part = sets[i]
            if n >= 64:
                torch.bmm(g.view(8, B // 8, -1).transpose(1, 2), x.view(8, B // 8, -1), out=part)
                torch.sum(part, dim=0, out=o)
            else:
                torch.mm(g.t(), x, out=o)

        def mm(i):
            g, x, o, _ = sets[i]
            torch.mm(g.t(), x, out=o)

        t = time_graph(k15, copies, 40)
        tb = time_graph(bmm8, copies, 40)
        tm = time_graph(mm, copies, 40)
        flops = 2.0 * B * n * k
        nbytes = 4.0 * (B * (n + k) + n * k)
        print(f"{n:>4} x {k:<9}{t:>10.2f}{flops / t / 1e6:>10.1f}{nbytes / t / 1e3:>18.1f}{tb:>14.2f}{tm:>10.2f}")


if __name__ == "__main__":
    main()
