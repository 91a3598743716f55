"""K12 vs cuBLAS addmm + elu_ on the actor-critic layer shapes (TF32): python tools/linear_bench.py"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from locotouch_b200 import ops  # noqa: E402
from locotouch_b200.streams import graph_capture  # noqa: E402

torch.backends.cuda.matmul.allow_tf32 = True


def timed(fn, reps=40):
    """Mean microseconds per call, the calls replayed from one CUDA graph (no host launch overhead)."""
    s = torch.cuda.Stream()
    with torch.cuda.stream(s):
        for _ in range(3):
            fn()
        s.synchronize()
        g = torch.cuda.CUDAGraph()
        with graph_capture(g, stream=s):
            for _ in range(reps):
                fn()
        g.replay()
        s.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        best = 1e30
        for _ in range(3):
            e0.record(s)
            g.replay()
            e1.record(s)
            e1.synchronize()
            best = min(best, e0.elapsed_time(e1) / reps * 1e3)
    return best


for (M, K, N, elu) in ((24576, 348, 512, True), (24576, 512, 256, True), (24576, 256, 128, True), (24576, 128, 12, False),
                       (4096, 348, 512, True), (4096, 512, 256, True), (4096, 256, 128, True)):
    x, w, b = torch.randn(M, K, device="cuda"), torch.randn(N, K, device="cuda"), torch.randn(N, device="cuda")
    out = torch.empty(M, N, device="cuda")

    def cublas():
        torch.addmm(b, x, w.t(), out=out)
        if elu:
            torch.nn.functional.elu_(out)

    t0 = timed(cublas)
    t1 = timed(lambda: ops.linear_bias_act(x, w, b, out=out, elu=elu))
    print(f"{M:6d} x {K:3d} x {N:3d} elu={int(elu)}: cuBLAS addmm{'+elu_' if elu else ''} {t0:7.1f} us   fused tcgen05 {t1:7.1f} us   ({2 * M * K * N / t1 / 1e6:6.1f} TFLOP/s)")

# dgrad with the ELU backward fused (K12 backward)
for (M, Nout, Kin) in ((24576, 256, 512), (24576, 128, 256)):
    g, w, h = torch.randn(M, Nout, device="cuda"), torch.randn(Nout, Kin, device="cuda"), torch.randn(M, Kin, device="cuda")
    out = torch.empty(M, Kin, device="cuda")
    t1 = timed(lambda: ops.dgrad_act_bwd(g, w, h, out=out))
    print(f"dgrad {M:6d} x {Nout:3d} -> {Kin:3d}: fused tcgen05 {t1:7.1f} us   ({2 * M * Kin * Nout / t1 / 1e6:6.1f} TFLOP/s, {4 * (M * Nout + 2 * M * Kin) / t1 / 1e3:6.0f} GB/s)")
