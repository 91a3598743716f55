"""Synthetic stand-in for the ``Isaac-RandCylinderTransportStudent`` env as the distillation loop sees it
(reference locotouch/distill/replay_buffer.py:33-56: ``get_observations() -> {"policy", "tactile"}``,
``step(action) -> (obs dict, reward, dones, extras)``, ``reset()``, ``num_envs``, ``device``).

PhysX is replaced by pre-generated state sets (BASELINE.json north_star); everything the managers compute around it runs on
the hot-path kernels: action processing (K0), the fused teacher MDP step (K1: rewards, terminations, the 348-wide policy
observation = 270 proprioceptive values + the 78-wide object-state history) and the binary taxel synthesis (K2, undelayed:
the delay line belongs to the caller's TactileRecorder)."""
from __future__ import annotations

import torch

from .. import ops
from ..mdp import task_spec as TS
from ..mdp.fused import FusedMdp
from . import synth
from .scene import ActionTermState

ACTION_CLIP, ACTION_RAW_SCALE = 100.0, 0.25  # reference locotouch/config/base/locomotion_base_env_cfg.py:127-135


class SyntheticTransportEnv:
    def __init__(self, num_envs: int, device="cuda:0", seed: int = 0, num_state_sets: int = 4, max_episode_length: int | None = None):
        self.num_envs, self.device = num_envs, torch.device(device)
        self.spec = TS.teacher_spec()
        if max_episode_length is not None:
            self.spec.max_episode_length = max_episode_length
        base = synth.make_env(num_envs, seed=seed, with_object=True, with_tactile=True, max_episode_length=self.spec.max_episode_length)
        self.action_term = ActionTermState(num_envs, synth.NUM_JOINTS, self.device)
        self.sets = []
        for k in range(num_state_sets):
            if k:
                synth.advance(base, keep_cmd_prob=0.9)
            denv = base.to(self.device)
            denv.action_manager._terms["joint_pos"] = self.action_term
            self.sets.append(denv)
        self.default_joint_pos = self.sets[0].scene["robot"].data.default_joint_pos.clone()
        self.mdp = FusedMdp(self.sets[0], self.spec, seed=seed)
        g = torch.Generator().manual_seed(seed + 17)
        self.taxel_thr = (0.05 + (torch.rand(num_envs, 221, generator=g) * 0.02 - 0.01)).to(self.device)
        self.tactile = torch.zeros(num_envs, 442, device=self.device)
        self.episode_length = torch.zeros(num_envs, device=self.device, dtype=torch.int64)
        self.t = 0
        self._observe()

    def _bind(self):
        env = self.sets[self.t % len(self.sets)]
        env.episode_length_buf.copy_(self.episode_length)  # episode progress is the env's, not the state set's
        self.mdp.env = env
        self.mdp._bound_ptrs = None
        return env

    def _taxels(self, env):
        ops.taxel_synth(env.scene["robot"].data.body_quat_w, env.scene.sensors["tactile_contact_sensor"].data.net_forces_w, self.taxel_thr,
                        quat_body_offset=synth.NUM_ROBOT_BODIES, p_drop=0.005, p_add=0.005, seed=self.mdp.seed + 1, offset=self.t,
                        signal=self.tactile, want_packed=False)

    def _observe(self):
        env = self._bind()
        self.mdp.compute_observations()
        self._taxels(env)

    def get_observations(self):
        return {"policy": self.mdp.policy_obs, "critic": self.mdp.critic_obs, "tactile": self.tactile}

    def reset(self):
        self.episode_length.zero_()
        self.mdp.reset()
        self._observe()

    def step(self, action: torch.Tensor):
        a = self.action_term
        ops.process_actions(action.contiguous(), a.raw_actions, a.prev_raw_actions, a.prev_prev_raw_actions, a.processed_actions,
                            clip=ACTION_CLIP, raw_scale=ACTION_RAW_SCALE, scale=1.0, offset=self.default_joint_pos)
        self.t += 1
        self.episode_length += 1
        env = self._bind()
        self.mdp.step(True, True)  # terminations -> rewards -> reset of done envs -> observations, one launch
        self._taxels(env)
        dones = self.mdp.dones.bool()
        self.episode_length.masked_fill_(dones, 0)
        obs = self.get_observations()
        return obs, self.mdp.reward_buf, dones, {"time_outs": self.mdp.time_outs, "observations": obs}
