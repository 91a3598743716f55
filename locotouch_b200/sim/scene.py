"""Synthetic stand-in for the slice of the IsaacLab scene API that the LocoTouch hot path reads.

PhysX stepping stays IsaacLab's; in the bench and in the tests *synthetic state tensors* stand in for it
(BASELINE.json north_star).  This module provides the container objects those tensors live in, shaped exactly
like the attributes the reference terms dereference (SURVEY.md section 8b "What the terms read"):

    env.scene["robot"].data.root_lin_vel_b            reference locotouch/mdp/rewards.py:19
    env.scene.sensors["robot_contact_senosr"].data    reference locotouch/mdp/rewards.py:38-39
    env.command_manager.get_command("base_velocity")  reference locotouch/mdp/rewards.py:19
    env.action_manager.get_term("joint_pos")          reference locotouch/mdp/rewards.py:455
    env.reward_manager.get_term_cfg(name).params      reference locotouch/mdp/rewards.py:380
    env.termination_manager.terminated                [IL] isaaclab.envs.mdp.is_alive

The same objects drive (a) the CUDA path of this package, (b) the CPU oracle, and (c) -- in the build container
only -- the unmodified reference modules loaded by ``oracle/ref_loader.py``.  Nothing here computes anything.
"""
from __future__ import annotations

import re
from types import SimpleNamespace
from typing import Sequence

import torch

# Go1/LocoTouch rigid bodies seen by the robot contact sensor, breadth-first as IsaacLab enumerates them
# (prim path "{ENV_REGEX_NS}/Robot/(?!sensor.*).*", reference locomotion_base_env_cfg.py:35-39 -> 17 bodies).
_LEGS = ("a_FR", "b_FL", "c_RR", "d_RL")
ROBOT_BODY_NAMES = (
    ["trunk"]
    + [f"{leg}_hip" for leg in _LEGS]
    + [f"{leg}_thigh" for leg in _LEGS]
    + [f"{leg}_calf" for leg in _LEGS]
    + [f"{leg}_foot" for leg in _LEGS]
)
ROBOT_JOINT_NAMES = (
    [f"{leg}_hip_joint" for leg in _LEGS]
    + [f"{leg}_thigh_joint" for leg in _LEGS]
    + [f"{leg}_calf_joint" for leg in _LEGS]
)
# 17 x 13 taxel links "sensor_RR_CC" (reference utils/urdf_processor/generate_locotouch_urdf.py:59-75)
TAXEL_ROWS, TAXEL_COLS = 17, 13
TAXEL_BODY_NAMES = [f"sensor_{r + 1:02d}_{c + 1:02d}" for r in range(TAXEL_ROWS) for c in range(TAXEL_COLS)]


def match_names(keys: str | Sequence[str], names: Sequence[str]) -> tuple[list[int], list[str]]:
    """Regex full-match of ``keys`` against ``names``; ids are returned in ``names`` order.

    [IL] mirrors isaaclab.utils.string.resolve_matching_names(preserve_order=False), which is what
    ``ContactSensor.find_bodies`` / ``SceneEntityCfg.resolve`` use.
    """
    if isinstance(keys, str):
        keys = [keys]
    ids, out = [], []
    for i, n in enumerate(names):
        if any(re.fullmatch(k, n) for k in keys):
            ids.append(i)
            out.append(n)
    if not ids:
        raise ValueError(f"no name in {list(names)[:6]}... matches {list(keys)}")
    return ids, out


class SceneEntityCfg:
    """[IL] isaaclab.managers.SceneEntityCfg: names an asset/sensor and (after resolve) the body / joint ids."""

    def __init__(self, name: str, body_names=None, joint_names=None, body_ids=slice(None), joint_ids=slice(None), **_):
        self.name = name
        self.body_names = body_names
        self.joint_names = joint_names
        self.body_ids = body_ids
        self.joint_ids = joint_ids

    def resolve(self, scene: "Scene") -> "SceneEntityCfg":
        entity = scene.sensors[self.name] if self.name in scene.sensors else scene[self.name]
        if self.body_names is not None:
            ids, _ = match_names(self.body_names, entity.body_names)
            self.body_ids = slice(None) if len(ids) == len(entity.body_names) else ids
        if self.joint_names is not None and hasattr(entity, "joint_names"):
            ids, _ = match_names(self.joint_names, entity.joint_names)
            self.joint_ids = slice(None) if len(ids) == len(entity.joint_names) else ids
        return self

    def __repr__(self):
        return f"SceneEntityCfg({self.name!r}, body_ids={self.body_ids}, joint_ids={self.joint_ids})"


class Entity:
    """An articulation, rigid object or contact sensor: ``.data`` namespace of tensors + name lookup."""

    def __init__(self, body_names: Sequence[str], joint_names: Sequence[str] = (), device="cpu"):
        self.body_names = list(body_names)
        self.joint_names = list(joint_names)
        self.device = device
        self.data = SimpleNamespace()

    @property
    def num_bodies(self) -> int:
        return len(self.body_names)

    def find_bodies(self, name_keys, preserve_order: bool = False):
        return match_names(name_keys, self.body_names)

    def find_joints(self, name_keys, preserve_order: bool = False):
        return match_names(name_keys, self.joint_names)


class Scene:
    def __init__(self, num_envs: int):
        self.num_envs = num_envs
        self._entities: dict[str, Entity] = {}
        self.sensors: dict[str, Entity] = {}

    def __getitem__(self, key: str) -> Entity:
        if key in self._entities:
            return self._entities[key]
        return self.sensors[key]

    def __setitem__(self, key: str, value: Entity):
        self._entities[key] = value

    def keys(self):
        return list(self._entities) + list(self.sensors)


class CommandManager:
    def __init__(self):
        self._commands: dict[str, torch.Tensor] = {}
        self._terms: dict[str, object] = {}

    def set_command(self, name: str, value: torch.Tensor):
        self._commands[name] = value

    def get_command(self, name: str) -> torch.Tensor:
        return self._commands[name]

    def get_term(self, name: str):
        return self._terms[name]


class ActionTermState:
    """State block of JointPositionActionPrevPrev that the hot path reads (reference mdp/actions.py:55-69)."""

    def __init__(self, num_envs: int, action_dim: int, device="cpu"):
        self.raw_actions = torch.zeros(num_envs, action_dim, device=device)
        self.prev_raw_actions = torch.zeros(num_envs, action_dim, device=device)
        self.prev_prev_raw_actions = torch.zeros(num_envs, action_dim, device=device)
        self.processed_actions = torch.zeros(num_envs, action_dim, device=device)


class ActionManager:
    def __init__(self):
        self._terms: dict[str, object] = {}

    def get_term(self, name: str):
        return self._terms[name]

    @property
    def action(self):
        return self._terms["joint_pos"].raw_actions


class RewardManagerView:
    """Just enough of [IL] RewardManager for ``get_term_cfg(name).params`` (reference rewards.py:380-381)."""

    def __init__(self):
        self._term_cfgs: dict[str, SimpleNamespace] = {}
        self._episode_sums: dict[str, torch.Tensor] = {}

    def set_term_cfg(self, name: str, cfg):
        self._term_cfgs[name] = cfg

    def get_term_cfg(self, name: str):
        return self._term_cfgs[name]

    @property
    def active_terms(self):
        return list(self._term_cfgs)


class TerminationManagerView:
    def __init__(self, num_envs: int, device="cpu"):
        self.terminated = torch.zeros(num_envs, dtype=torch.bool, device=device)
        self.time_outs = torch.zeros(num_envs, dtype=torch.bool, device=device)

    @property
    def dones(self):
        return self.terminated | self.time_outs


class SynthEnv:
    """Duck-typed ManagerBasedRLEnv holding synthetic state (no simulator)."""

    def __init__(self, num_envs: int, device="cpu", step_dt: float = 0.02, max_episode_length: int = 1000):
        self.num_envs = num_envs
        self.device = torch.device(device) if not isinstance(device, torch.device) else device
        self.step_dt = step_dt  # sim.dt 0.005 x decimation 4 (reference locomotion_base_env_cfg.py:345-348)
        self.max_episode_length = max_episode_length  # ceil(20 s / 0.02 s) (reference :346)
        self.max_episode_length_s = max_episode_length * step_dt
        self.common_step_counter = 0
        self.episode_length_buf = torch.zeros(num_envs, dtype=torch.long, device=self.device)
        self.scene = Scene(num_envs)
        self.command_manager = CommandManager()
        self.action_manager = ActionManager()
        self.reward_manager = RewardManagerView()
        self.termination_manager = TerminationManagerView(num_envs, self.device)
        self.extras: dict = {}

    # ------------------------------------------------------------------ helpers
    def named_tensors(self) -> dict[str, torch.Tensor]:
        """Flat ``{"robot.root_pos_w": tensor, ...}`` view of every state tensor (used to move / save / compare)."""
        out: dict[str, torch.Tensor] = {}
        for key in self.scene.keys():
            ent = self.scene[key]
            for attr, val in vars(ent.data).items():
                if torch.is_tensor(val):
                    out[f"{key}.{attr}"] = val
        for name, val in self.command_manager._commands.items():
            out[f"command.{name}"] = val
        for name, term in self.action_manager._terms.items():
            for attr in ("raw_actions", "prev_raw_actions", "prev_prev_raw_actions", "processed_actions"):
                out[f"action.{name}.{attr}"] = getattr(term, attr)
        out["episode_length_buf"] = self.episode_length_buf
        out["terminated"] = self.termination_manager.terminated
        out["time_outs"] = self.termination_manager.time_outs
        return out

    def load_named_tensors(self, tensors: dict[str, torch.Tensor]):
        for key, val in tensors.items():
            val = val.to(self.device)
            if key == "episode_length_buf":
                self.episode_length_buf = val
            elif key == "terminated":
                self.termination_manager.terminated = val
            elif key == "time_outs":
                self.termination_manager.time_outs = val
            elif key.startswith("command."):
                self.command_manager._commands[key.split(".", 1)[1]] = val
            elif key.startswith("action."):
                _, name, attr = key.split(".")
                setattr(self.action_manager._terms[name], attr, val)
            else:
                ent, attr = key.split(".", 1)
                setattr(self.scene[ent].data, attr, val)

    def to(self, device) -> "SynthEnv":
        """Deep-copy onto ``device`` (same bits; CPU oracle and GPU path see identical inputs)."""
        import copy

        other = copy.copy(self)
        other.device = torch.device(device)
        other.scene = Scene(self.num_envs)
        for key in self.scene._entities:
            src = self.scene._entities[key]
            dst = Entity(src.body_names, src.joint_names, other.device)
            other.scene[key] = dst
        for key, src in self.scene.sensors.items():
            other.scene.sensors[key] = Entity(src.body_names, src.joint_names, other.device)
        for key in self.scene.keys():
            for attr, val in vars(self.scene[key].data).items():
                setattr(other.scene[key].data, attr, val.to(other.device).clone() if torch.is_tensor(val) else val)
        other.command_manager = CommandManager()
        for name, val in self.command_manager._commands.items():
            other.command_manager.set_command(name, val.to(other.device).clone())
        other.action_manager = ActionManager()
        for name, term in self.action_manager._terms.items():
            new = ActionTermState(self.num_envs, term.raw_actions.shape[1], other.device)
            for attr in ("raw_actions", "prev_raw_actions", "prev_prev_raw_actions", "processed_actions"):
                setattr(new, attr, getattr(term, attr).to(other.device).clone())
            other.action_manager._terms[name] = new
        other.reward_manager = RewardManagerView()
        other.reward_manager._term_cfgs = dict(self.reward_manager._term_cfgs)
        other.termination_manager = TerminationManagerView(self.num_envs, other.device)
        other.termination_manager.terminated = self.termination_manager.terminated.to(other.device).clone()
        other.termination_manager.time_outs = self.termination_manager.time_outs.to(other.device).clone()
        other.episode_length_buf = self.episode_length_buf.to(other.device).clone()
        other.extras = {}
        return other
