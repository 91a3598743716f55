"""TEST INFRASTRUCTURE ONLY -- CPU restatement (plain Python loops) of the reference's DAgger ReplayBuffer.  Never imported
by the product path (locotouch_b200/).

Follows locotouch/distill/replay_buffer.py:20-80 (collect_data / _record_new_traj), :82-112 (to_recurrent_generator /
_prepare_padded_sequence) and :118-140 (evaluate).  Pinned against the unmodified reference class by
tests/golden/dagger_c4.npz (tests/golden/make_golden.py::golden_dagger)."""
from __future__ import annotations

import numpy as np
import torch

from .tactile import TactileDelayOracle


class ReplayBufferOracle:
    def __init__(self, env, proprioception_dim: int, tactile_dim: int, min_delay=1, max_delay=2):
        self.env, self.P = env, proprioception_dim
        self.N = env.num_envs
        self.delay = TactileDelayOracle(self.N, tactile_dim, min_delay, max_delay)
        self.trajs = []  # (proprioceptions [len, P], teacher obs [len, E], tactile [len, X]) per recorded trajectory
        self.steps_count = 0
        self.reward_sums = torch.zeros(self.N)

    def collect_data(self, act, num_steps: int, with_student: bool):
        if with_student:
            self.env.reset()  # replay_buffer.py:22-23
        self.delay.reset(torch.arange(self.N))
        rewards, lengths = [], []
        prop, teach, tact = [], [], []  # one [N, D] entry per step
        steps, start_count = 0, self.steps_count
        start = [0] * self.N
        obs = self.env.get_observations()
        pos, tactile = obs["policy"], obs["tactile"]
        while self.steps_count - start_count < num_steps:
            prop.append(pos[:, :self.P])
            teach.append(pos[:, self.P:])
            act(pos, tactile)
            self.delay.record(tactile)
            tact.append(self.delay.get().clone())
            nxt, reward, dones, _ = self.env.step(None)
            pos, tactile = nxt["policy"], nxt["tactile"]
            self.reward_sums += reward
            steps += 1
            done_ids = [n for n in range(self.N) if bool(dones[n])]
            if done_ids:
                self.delay.reset(torch.tensor(done_ids))
                for n in done_ids:
                    rewards.append(float(self.reward_sums[n]))
                    lengths.append(steps - start[n])
                    self.reward_sums[n] = 0
                for n in done_ids:
                    if self.steps_count - start_count < num_steps:  # the budget is tested env by env (:68), then break (:73)
                        s0 = start[n]
                        self.steps_count += steps - s0
                        self.trajs.append(tuple(torch.stack([x[n] for x in buf[s0:steps]]) for buf in (prop, teach, tact)))
                        start[n] = steps
                    else:
                        break
        return rewards, lengths

    def padded(self, traj_indices):
        """_prepare_padded_sequence (:90-112)."""
        lens = [self.trajs[i][0].shape[0] for i in traj_indices]
        L, B = max(lens), len(traj_indices)
        outs = [torch.zeros(L, B, self.trajs[0][k].shape[1]) for k in range(3)]
        masks = torch.zeros(L, B, dtype=torch.bool)
        for j, i in enumerate(traj_indices):
            for k in range(3):
                outs[k][:lens[j], j] = self.trajs[i][k]
            masks[:lens[j], j] = True
        return dict(proprioceptions=outs[0], teacher_encoder_obses=outs[1], tactile_signals=outs[2], masks=masks)

    def batches(self, batch_size: int):
        idx = np.random.permutation(np.arange(len(self.trajs)))  # :84-85
        return [self.padded(idx[s:min(s + batch_size, len(idx))]) for s in range(0, len(idx), batch_size)]

    def evaluate(self, num_trajs: int):
        """:118-140 on an RSL-style env (tuple API)."""
        rewards, lengths = [], []
        count = [0] * self.N
        self.env.get_observations()
        while len(rewards) < num_trajs:
            _, reward, dones, _ = self.env.step(None)
            self.reward_sums += reward
            count = [c + 1 for c in count]
            for n in range(self.N):
                if bool(dones[n]):
                    rewards.append(float(self.reward_sums[n]))
                    lengths.append(float(count[n]))
                    self.reward_sums[n] = 0
                    count[n] = 0
        return rewards, lengths
