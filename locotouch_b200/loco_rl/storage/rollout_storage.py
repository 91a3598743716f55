"""RolloutStorage with the reference's interface (reference loco_rl/loco_rl/storage/rollout_storage.py:13-318).

Same constructor, public ``[T, N, ...]`` tensor attributes, ``add_transitions`` / ``clear`` / ``compute_returns`` /
``get_statistics`` / ``mini_batch_generator`` and the same ``OverflowError``; below the interface

* ``add_transitions`` is the fused store kernel (K3) and skips every copy whose source already *is* the slot
  (``PPO.act`` of this package writes actions / log-prob / mu / sigma / values straight into ``slot(step)``);
* ``compute_returns`` is the GAE scan + normalisation kernel pair (K4);
* ``mini_batch_generator`` draws the single permutation of rollout_storage.py:189, gathers the permuted rollout ONCE
  with the fused row-gather kernel (K5) and yields contiguous slices of it for every epoch -- the reference re-gathers
  nine tensors for each of the epochs x mini-batches steps.
"""
from __future__ import annotations

import torch

from ... import ops


class RolloutStorage:
    class Transition:
        def __init__(self):
            self.observations = None
            self.critic_observations = None
            self.actions = None
            self.rewards = None
            self.dones = None
            self.values = None
            self.actions_log_prob = None
            self.action_mean = None
            self.action_sigma = None
            self.hidden_states = None
            self.rnd_state = None

        def clear(self):
            self.__init__()

    def __init__(self, num_envs, num_transitions_per_env, obs_shape, privileged_obs_shape, actions_shape, rnd_state_shape=None, device="cpu"):
        self.device = device
        self.num_transitions_per_env = num_transitions_per_env
        self.num_envs = num_envs
        self.obs_shape = obs_shape
        self.privileged_obs_shape = privileged_obs_shape
        self.rnd_state_shape = rnd_state_shape
        self.actions_shape = actions_shape
        T, N = num_transitions_per_env, num_envs
        # one extra observation slot: slot T receives the observation that follows the last transition (zero-copy
        # rollouts write the env's observation pass directly into slot t+1); the public attributes are the [:T] views
        self._obs_buf = torch.zeros(T + 1, N, *obs_shape, device=self.device)
        self.observations = self._obs_buf[:T]
        if privileged_obs_shape is not None:
            self._priv_buf = torch.zeros(T + 1, N, *privileged_obs_shape, device=self.device)
            self.privileged_observations = self._priv_buf[:T]
        else:
            self._priv_buf = None
            self.privileged_observations = None
        self.rewards = torch.zeros(T, N, 1, device=self.device)
        self.actions = torch.zeros(T, N, *actions_shape, device=self.device)
        self.dones = torch.zeros(T, N, 1, device=self.device).byte()
        self.actions_log_prob = torch.zeros(T, N, 1, device=self.device)
        self.values = torch.zeros(T, N, 1, device=self.device)
        self.returns = torch.zeros(T, N, 1, device=self.device)
        self.advantages = torch.zeros(T, N, 1, device=self.device)
        self.mu = torch.zeros(T, N, *actions_shape, device=self.device)
        self.sigma = torch.zeros(T, N, *actions_shape, device=self.device)
        if rnd_state_shape is not None:
            self.rnd_state = torch.zeros(T, N, *rnd_state_shape, device=self.device)
        self.saved_hidden_states_a = None
        self.saved_hidden_states_c = None
        self.step = 0
        self._perm_bufs = None

    # ------------------------------------------------------------------------------------------------------ zero copy
    def slot(self, step: int | None = None) -> dict:
        """Row views of transition ``step`` that producers may write into directly."""
        s = self.step if step is None else step
        if s >= self.num_transitions_per_env:
            raise OverflowError("Rollout buffer overflow! You should call clear() before adding new transitions.")
        return dict(actions=self.actions[s], logp=self.actions_log_prob[s].view(-1), mu=self.mu[s], sigma=self.sigma[s],
                    values=self.values[s], observations=self._obs_buf[s],
                    critic_observations=self._priv_buf[s] if self._priv_buf is not None else None,
                    next_observations=self._obs_buf[s + 1], next_critic_observations=self._priv_buf[s + 1] if self._priv_buf is not None else None)

    @staticmethod
    def _copy_if_needed(dst: torch.Tensor, src: torch.Tensor):
        if src is None:
            return
        if src.data_ptr() != dst.data_ptr():
            dst.copy_(src.view(dst.shape) if src.numel() == dst.numel() else src)

    def add_transitions(self, transition: "RolloutStorage.Transition", time_outs=None, gamma: float = 0.0):
        """``time_outs`` / ``gamma``: when given, the time-out bootstrap of ppo.py:162-165 is fused into the store."""
        if self.step >= self.num_transitions_per_env:
            raise OverflowError("Rollout buffer overflow! You should call clear() before adding new transitions.")
        s = self.step
        t = transition
        obs, cobs = t.observations, t.critic_observations
        # everything already in its row (a caller that let the MDP launch write rewards + dones into the slot -- K3 inside K1 -- and built the
        # observations in place): nothing to launch
        in_place = (time_outs is None and t.rewards.data_ptr() == self.rewards[s].data_ptr() and t.dones.data_ptr() == self.dones[s].data_ptr()
                    and t.dones.dtype == torch.uint8 and obs.data_ptr() == self._obs_buf[s].data_ptr()
                    and (self._priv_buf is None or cobs is None or cobs.data_ptr() == self._priv_buf[s].data_ptr()))
        if not in_place:
            ops.store_step(
                t.rewards.view(-1), t.dones.view(-1), time_outs, t.values.view(-1) if t.values is not None else None, gamma,
                self.rewards[s].view(-1), self.dones[s].view(-1),
                obs, self._obs_buf[s], cobs if self._priv_buf is not None else None, self._priv_buf[s] if self._priv_buf is not None else None)
        self._copy_if_needed(self.actions[s], t.actions)
        self._copy_if_needed(self.values[s], t.values)
        self._copy_if_needed(self.actions_log_prob[s], t.actions_log_prob)
        self._copy_if_needed(self.mu[s], t.action_mean)
        self._copy_if_needed(self.sigma[s], t.action_sigma)
        if self.rnd_state_shape is not None:
            self.rnd_state[s].copy_(t.rnd_state)
        self._save_hidden_states(t.hidden_states)
        self.step += 1

    def _save_hidden_states(self, hidden_states):
        """reference rollout_storage.py:112-146: per-step copies of the actor / critic RNN states ([layers, N, hidden]; an LSTM
        passes a tuple); the critic entry may be None."""
        if hidden_states is None or hidden_states == (None, None):
            return
        hid_a = hidden_states[0] if isinstance(hidden_states[0], tuple) else (hidden_states[0],)
        hid_c = None
        if hidden_states[1] is not None:
            hid_c = hidden_states[1] if isinstance(hidden_states[1], tuple) else (hidden_states[1],)
        T = self.observations.shape[0]
        if self.saved_hidden_states_a is None:
            self.saved_hidden_states_a = [torch.zeros(T, *h.shape, device=self.device) for h in hid_a]
            if hid_c is not None:
                self.saved_hidden_states_c = [torch.zeros(T, *h.shape, device=self.device) for h in hid_c]
        for i, h in enumerate(hid_a):
            self.saved_hidden_states_a[i][self.step].copy_(h)
            if hid_c is not None:
                self.saved_hidden_states_c[i][self.step].copy_(hid_c[i])

    def clear(self):
        self.step = 0

    # ----------------------------------------------------------------------------------------------------------- K4
    def compute_returns(self, last_values, gamma, lam, normalize_advantage: bool = True):
        ops.gae(self.rewards, self.values, self.dones, last_values.detach().contiguous().view(-1), gamma, lam, normalize_advantage,
                self.returns, self.advantages)

    def get_statistics(self):
        done = self.dones
        done[-1] = 1  # reference mutates the buffer in place (rollout_storage.py:177-178); preserved
        flat_dones = done.permute(1, 0, 2).reshape(-1, 1)
        done_indices = torch.cat((flat_dones.new_tensor([-1], dtype=torch.int64), flat_dones.nonzero(as_tuple=False)[:, 0]))
        trajectory_lengths = done_indices[1:] - done_indices[:-1]
        return trajectory_lengths.float().mean(), self.rewards.mean()

    # ----------------------------------------------------------------------------------------------------------- K5
    def _flat_sources(self):
        obs = self.observations.flatten(0, 1)
        cobs = self.privileged_observations.flatten(0, 1) if self.privileged_observations is not None else None
        srcs = [obs] + ([cobs] if cobs is not None else []) + [
            self.actions.flatten(0, 1), self.values.flatten(0, 1), self.returns.flatten(0, 1), self.actions_log_prob.flatten(0, 1),
            self.advantages.flatten(0, 1), self.mu.flatten(0, 1), self.sigma.flatten(0, 1)]
        if self.rnd_state_shape is not None:
            srcs.append(self.rnd_state.flatten(0, 1))
        return srcs, cobs is not None

    def gather_permuted(self, indices: torch.Tensor):
        """One fused gather of every per-sample tensor in permuted order; buffers are reused across updates."""
        srcs, has_priv = self._flat_sources()
        count = indices.numel()
        if self._perm_bufs is None or self._perm_bufs[0].shape[0] != count:
            self._perm_bufs = [torch.empty((count,) + tuple(s.shape[1:]), device=s.device) for s in srcs]
        ops.gather_rows(srcs, indices, self._perm_bufs)
        return self._perm_bufs, has_priv

    def mini_batch_generator(self, num_mini_batches, num_epochs=8, indices=None):
        batch_size = self.num_envs * self.num_transitions_per_env
        mini_batch_size = batch_size // num_mini_batches
        if indices is None:
            indices = torch.randperm(num_mini_batches * mini_batch_size, requires_grad=False, device=self.device)
        bufs, has_priv = self.gather_permuted(indices)
        it = iter(bufs)
        obs = next(it)
        cobs = next(it) if has_priv else obs
        actions, values, returns, logp, adv, mu, sigma = (next(it) for _ in range(7))
        rnd = next(it) if self.rnd_state_shape is not None else None
        for _ in range(num_epochs):
            for i in range(num_mini_batches):
                sl = slice(i * mini_batch_size, (i + 1) * mini_batch_size)
                yield obs[sl], cobs[sl], actions[sl], values[sl], adv[sl], returns[sl], logp[sl], mu[sl], sigma[sl], (None, None), None, (
                    rnd[sl] if rnd is not None else None)

    def recurrent_mini_batch_generator(self, num_mini_batches, num_epochs=8):
        """reference rollout_storage.py:246-318.  Mini-batches are env ranges; observations travel as zero-padded trajectories
        (split at the dones) with their masks and the RNN states at each trajectory's first step.  The trajectory index is
        built once on the device (K10) and its per-env offsets are read back once, instead of the per-mini-batch
        ``torch.sum(last_was_done[:, start:stop])`` host reads and boolean-mask gathers of the reference."""
        index = ops.TrajectoryIndex(self.dones)
        padded_obs, masks = index.split_and_pad(self.observations)
        if self.privileged_observations is not None:
            padded_cobs, _ = index.split_and_pad(self.privileged_observations, want_masks=False)
        else:
            padded_cobs = padded_obs
        padded_rnd = index.split_and_pad(self.rnd_state, want_masks=False)[0] if self.rnd_state_shape is not None else None
        start_steps, envs = index.start.long(), index.env.long()

        def first_step_states(saved):  # [T, layers, N, hidden] -> [layers, M, hidden] at every trajectory's first step
            if saved is None:
                return None
            return [h[start_steps, :, envs].transpose(1, 0).contiguous() for h in saved]

        hid_a_all = first_step_states(self.saved_hidden_states_a)
        hid_c_all = first_step_states(self.saved_hidden_states_c)
        mini_batch_size = self.num_envs // num_mini_batches
        for _ in range(num_epochs):
            for i in range(num_mini_batches):
                start, stop = i * mini_batch_size, (i + 1) * mini_batch_size
                tr = slice(index.base[start], index.base[stop])
                hid_a = [h[:, tr].contiguous() for h in hid_a_all] if hid_a_all is not None else None
                hid_c = [h[:, tr].contiguous() for h in hid_c_all] if hid_c_all is not None else None
                if hid_a is not None and len(hid_a) == 1:  # remove the tuple for GRU
                    hid_a = hid_a[0]
                if hid_c is not None and len(hid_c) == 1:
                    hid_c = hid_c[0]
                yield (padded_obs[:, tr], padded_cobs[:, tr], self.actions[:, start:stop], self.values[:, start:stop],
                       self.advantages[:, start:stop], self.returns[:, start:stop], self.actions_log_prob[:, start:stop],
                       self.mu[:, start:stop], self.sigma[:, start:stop], (hid_a, hid_c), masks[:, tr],
                       padded_rnd[:, tr] if padded_rnd is not None else None)
