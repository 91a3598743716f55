"""Where the rollout graph's time goes: 24 x policy step (K19 + K3b) and 24 x env step (K0 + K1 || K2 + store) captured as separate
CUDA graphs next to the full rollout graph.    python tools/rollout_split.py"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from locotouch_b200.engine import HotPathEngine  # noqa: E402
from locotouch_b200.streams import graph_capture  # noqa: E402


def timed(g, reps=10):
    for _ in range(3):
        g.replay()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        g.replay()
    e1.record()
    e1.synchronize()
    return e0.elapsed_time(e1) / reps


eng = HotPathEngine(num_envs=4096, task="teacher", tactile=True, device=torch.device("cuda:0"), seed=0, num_state_sets=6)
eng.capture()
alg, st = eng.alg, eng.alg.storage
print(f"rollout graph (24 steps + GAE): {timed(eng._graphs['roll'][0]):.3f} ms   update graph: {timed(eng._graphs['update']):.3f} ms")
ac = alg.actor_critic
g_pol = torch.cuda.CUDAGraph()
with graph_capture(g_pol):
    for t in range(eng.T):
        st.step = t
        alg.act(st._obs_buf[t], st._priv_buf[t])
st.step = 0
print(f"24 x policy step (hidden layers K19, heads + sample K3b): {timed(g_pol):.3f} ms = {timed(g_pol) / 24 * 1e3:.1f} us per step")
actions = st.actions[0]
g_env = torch.cuda.CUDAGraph()
with graph_capture(g_env):
    for t in range(eng.T):
        eng.env_step(t, actions, 0)
print(f"24 x env step (K0 + K1 + K3 store in one launch || K2, delay-line bookkeeping; LT_FUSE_K3=0: K3 separate, below): {timed(g_env):.3f} ms = {timed(g_env) / 24 * 1e3:.1f} us per step")
g_store = torch.cuda.CUDAGraph()
rew, dn, inf = eng.mdp.reward_buf, eng.mdp.dones, {"time_outs": eng.mdp.time_outs}
with graph_capture(g_store):
    for t in range(eng.T):
        st.step = t
        alg.transition.observations, alg.transition.critic_observations = st._obs_buf[t], st._priv_buf[t]
        alg.transition.actions, alg.transition.values = st.actions[t], st.values[t]
        alg.transition.actions_log_prob, alg.transition.action_mean, alg.transition.action_sigma = st.actions_log_prob[t], st.mu[t], st.sigma[t]
        alg.process_env_step(rew, dn, inf)
st.step = 0
print(f"24 x process_env_step as a separate K3 launch (not part of the rollout graph when K3 is fused into K1): {timed(g_store):.3f} ms = {timed(g_store) / 24 * 1e3:.1f} us per step")
