"""Repeatability of the TF32 training pass: the same forward + backward N times, deviation of every gradient tensor from the float64
autograd reference.    [LT_WGRAD_PAIR=0] [LT_MLP3=0] python tools/backward_stress.py [reps]"""
import copy
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from locotouch_b200.loco_rl.modules.actor_critic import ActorCritic  # noqa: E402

reps = int(sys.argv[1]) if len(sys.argv) > 1 else 20
B, A, OBS, hidden = 24576, 12, 348, [512, 256, 128]
torch.manual_seed(B)
ac = ActorCritic(OBS, OBS, A, actor_hidden_dims=hidden, critic_hidden_dims=hidden, activation="elu").cuda()
ref = copy.deepcopy(ac).double()
ac.flatten_parameters()
obs, cobs = torch.randn(B, OBS, device="cuda"), torch.randn(B, OBS, device="cuda")
g_mu, g_v = torch.randn(B, A, device="cuda") / B, torch.randn(B, 1, device="cuda") / B
mu64, v64 = ref.actor(obs.double()), ref.critic(cobs.double())
torch.autograd.backward([mu64, v64], [g_mu.double(), g_v.double()])
want = {n: p.grad for n, p in ref.named_parameters()}
torch.backends.cuda.matmul.allow_tf32 = True
worst = {}
for r in range(reps):
    ac.train_forward(obs, cobs)
    ac.train_backward(g_mu, g_v)
    torch.cuda.synchronize()
    for n, p in ac.named_parameters():
        if n in ("std", "log_std"):
            continue
        rel = ((p.grad.double() - want[n]).abs().max() / want[n].abs().max()).item()
        worst[n] = max(worst.get(n, 0.0), rel)
print(f"LT_WGRAD_PAIR={os.environ.get('LT_WGRAD_PAIR', '1')} LT_MLP3={os.environ.get('LT_MLP3', '1')} reps={reps}")
for n, v in worst.items():
    print(f"  {n:20s} worst relative-to-max error {v:.4f}")
