"""Wait-cycle attribution of K19's MMA thread: LT_MLP3_DBG=1 python tools/mlp3_dbg.py [B]"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from locotouch_b200 import ops  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
keep = (B > 8192) if os.environ.get("KEEP") is None else os.environ["KEEP"] == "1"
nets = []
for _ in range(2):
    x = torch.randn(B, 348, device="cuda")
    ps = []
    for (n, k) in ((512, 348), (256, 512), (128, 256)):
        ps += [torch.randn(n, k, device="cuda") / k ** 0.5, torch.randn(n, device="cuda")]
    hs = (torch.empty(B, 512, device="cuda") if keep else None, torch.empty(B, 256, device="cuda") if keep else None, torch.empty(B, 128, device="cuda"))
    nets.append((x, tuple(ps), hs))
for _ in range(3):
    ops.mlp3_forward(nets)
torch.cuda.synchronize()
