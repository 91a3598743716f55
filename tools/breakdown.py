"""Where does one PPO iteration go?  CUDA-event timing of the rollout graph, the update graph and the pieces of one
mini-batch step (eager, 20 reps each).  python tools/breakdown.py"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from locotouch_b200 import ops  # noqa: E402
from locotouch_b200.engine import HotPathEngine  # noqa: E402


def timed(fn, reps=20, warm=3):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    e1.synchronize()
    return e0.elapsed_time(e1) / reps


eng = HotPathEngine(num_envs=4096, task="teacher", tactile=True, device="cuda:0")
eng.capture()
g_roll, g_upd = eng._graphs['roll'][0], eng._graphs['update']
print(f"rollout graph  {timed(g_roll.replay):8.3f} ms")
print(f"update graph   {timed(g_upd.replay):8.3f} ms")
print(f"randperm       {timed(eng.draw_permutation):8.3f} ms")
alg, st, ac, opt = eng.alg, eng.alg.storage, eng.alg.actor_critic, eng.alg.optimizer
eng.rollout()
eng.draw_permutation()
bufs, _ = st.gather_permuted(eng.perm)
print(f"gather (once)  {timed(lambda: st.gather_permuted(eng.perm)):8.3f} ms")
B = 24576
obs, cobs = bufs[0][:B], bufs[1][:B]
actions, values, returns, logp, adv, mu_old, sig_old = (b[:B] for b in bufs[2:9])
lb = ops.PpoLossBuffers(B, 12, "cuda:0")
state = {}


def fwd():
    state["mu"] = ac.actor(obs)
    state["v"] = ac.critic(cobs)


def loss():
    ops.ppo_loss(state["mu"].detach(), ac.std.detach(), state["v"].detach().view(-1), actions, logp.view(-1), mu_old, sig_old, adv.view(-1),
                 returns.view(-1), values.view(-1), entropy_coef=0.01, desired_kl=0.01, lr=opt.lr_t, buffers=lb)


def fwd_bwd():
    fwd()
    opt.zero_grad()
    torch.autograd.backward([state["mu"], state["v"]], [lb.grad_mu, lb.grad_value.view_as(state["v"])])


with torch.no_grad():
    print(f"mlp fwd (no grad) {timed(fwd):8.3f} ms")
print(f"mlp fwd (grad)    {timed(fwd):8.3f} ms")
fwd()
print(f"ppo_loss          {timed(loss):8.3f} ms")
print(f"mlp fwd+bwd       {timed(fwd_bwd):8.3f} ms")
print(f"clip+adam         {timed(lambda: opt.step(max_grad_norm=1.0)):8.3f} ms")
# one rollout env step pieces
with torch.no_grad():
    o, c = st._obs_buf[0], st._priv_buf[0]
    def act0():
        st.step = 0
        alg.act(o, c)
    print(f"act (2 MLPs + sample) {timed(act0):8.3f} ms")
    a = st.actions[0]
    st.step = 0
    print(f"env_step (K0+K1+K2)   {timed(lambda: eng.env_step(0, a)):8.3f} ms")
# GEMM efficiency: the four layer shapes, TF32
for (m, k, n) in ((B, 348, 512), (B, 512, 256), (B, 256, 128), (B, 128, 12), (4096, 348, 512)):
    x, w = torch.randn(m, k, device="cuda"), torch.randn(n, k, device="cuda")
    t = timed(lambda: torch.nn.functional.linear(x, w), reps=50)
    print(f"linear {m}x{k}x{n}: {t * 1e3:7.1f} us  {2 * m * k * n / t / 1e9:7.1f} TFLOP/s")
# backward-shaped GEMMs: dgrad [B,n]x[n,k] and wgrad [n,B]x[B,k]
for (m, k, n) in ((B, 348, 512), (B, 512, 256), (B, 256, 128)):
    gy, w, x = torch.randn(m, n, device="cuda"), torch.randn(n, k, device="cuda"), torch.randn(m, k, device="cuda")
    t = timed(lambda: gy @ w, reps=50)
    print(f"dgrad {m}x{n}x{k}: {t * 1e3:7.1f} us  {2 * m * k * n / t / 1e9:7.1f} TFLOP/s")
    t = timed(lambda: gy.t() @ x, reps=50)
    print(f"wgrad {n}x{m}x{k}: {t * 1e3:7.1f} us  {2 * m * k * n / t / 1e9:7.1f} TFLOP/s")
    t = timed(lambda: gy.sum(0), reps=50)
    print(f"bias grad {m}x{n}: {t * 1e3:7.1f} us")
x = torch.randn(B, 512, device="cuda")
print(f"elu {B}x512: {timed(lambda: torch.nn.functional.elu(x), reps=50) * 1e3:7.1f} us")
# wgrad alternatives: explicit split-K through bmm
for (m, k, n) in ((B, 348, 512), (B, 512, 256), (B, 256, 128)):
    gy, x = torch.randn(m, n, device="cuda"), torch.randn(m, k, device="cuda")
    for S in (4, 8, 16):
        gs, xs = gy.view(S, m // S, n), x.view(S, m // S, k)
        t = timed(lambda: torch.bmm(gs.transpose(1, 2), xs).sum(0), reps=50)
        print(f"wgrad split-{S:2d} {n}x{m}x{k}: {t * 1e3:7.1f} us  {2 * m * k * n / t / 1e9:7.1f} TFLOP/s")
    t = timed(lambda: torch.mm(x.t(), gy), reps=50)
    print(f"wgrad (x^T g) {k}x{m}x{n}: {t * 1e3:7.1f} us")
