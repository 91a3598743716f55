"""ORACLE (test infrastructure, never imported by the product path).

Loads the UNMODIFIED reference modules from ``/root/reference`` under a small stub of the ``isaaclab.*`` names they
import.  Only usable in the build container -- ``/root/reference`` does not exist on the GPU box,
so nothing in ``-m gpu`` tests, ``smoke()`` or ``bench.py`` may call this.  Its two users are
``tests/golden/make_golden.py`` (generates the committed fixtures) and the ``not gpu`` tests that pin the oracle
restatement against the live reference (skipped when the reference is absent).

No reference source is copied: modules are executed from where they lie.
"""
from __future__ import annotations

import copy
import importlib
import importlib.util
import os
import sys
import types

REFERENCE_ROOT = os.environ.get("LOCOTOUCH_REFERENCE", "/root/reference")


def reference_available() -> bool:
    return os.path.isdir(os.path.join(REFERENCE_ROOT, "locotouch", "mdp"))


def _configclass(cls):
    """Minimal stand-in for [IL] isaaclab.utils.configclass: per-instance deep copies + __post_init__ chaining."""
    if not any("__post_init__" in vars(k) for k in cls.__mro__):
        cls.__post_init__ = lambda self: None

    def __init__(self, **kwargs):
        for klass in reversed(type(self).__mro__):
            for key, val in vars(klass).items():
                if key.startswith("__") or callable(val) or isinstance(val, (property, staticmethod, classmethod)):
                    continue
                setattr(self, key, copy.deepcopy(val))
        for key, val in kwargs.items():
            setattr(self, key, val)
        self.__post_init__()

    cls.__init__ = __init__
    return cls


def _module(name: str, **attrs) -> types.ModuleType:
    mod = sys.modules.get(name)
    if mod is None:
        mod = types.ModuleType(name)
        mod.__path__ = []  # behave as a package
        sys.modules[name] = mod
    for key, val in attrs.items():
        setattr(mod, key, val)
    return mod


_INSTALLED = False


def install_isaaclab_stub():
    """Register stub modules for every ``isaaclab.*`` name imported by the hot-path files of the reference."""
    global _INSTALLED
    if _INSTALLED:
        return
    from locotouch_b200.sim.scene import SceneEntityCfg
    from oracle import il_math

    class ManagerTermBase:  # [IL] isaaclab.managers.ManagerTermBase
        def __init__(self, cfg, env):
            self.cfg = cfg
            self._env = env

        @property
        def num_envs(self):
            return self._env.num_envs

        @property
        def device(self):
            return self._env.device

        def reset(self, env_ids=None):
            pass

    class _Cfg:
        def __init__(self, **kw):
            self.__dict__.update(kw)

    class JointPositionAction:  # only subclassed, never instantiated here
        def __init__(self, cfg, env):
            raise RuntimeError("stub")

    dummy = type("Dummy", (), {})
    _module("isaaclab")
    _module("isaaclab.assets", Articulation=dummy, RigidObject=dummy)
    _module(
        "isaaclab.managers",
        SceneEntityCfg=SceneEntityCfg,
        ManagerTermBase=ManagerTermBase,
        RewardTermCfg=_Cfg,
        ObservationTermCfg=_Cfg,
    )
    _module("isaaclab.managers.action_manager", ActionTerm=dummy)
    _module("isaaclab.sensors", ContactSensor=dummy)
    _module("isaaclab.utils", configclass=_configclass)
    _module(
        "isaaclab.utils.math",
        quat_apply=il_math.quat_apply,
        quat_apply_inverse=il_math.quat_apply_inverse,
        quat_mul=il_math.quat_mul,
        quat_inv=il_math.quat_inv,
        quat_from_euler_xyz=il_math.quat_from_euler_xyz,
        euler_xyz_from_quat=il_math.euler_xyz_from_quat,
    )
    _module("isaaclab.envs", ManagerBasedRLEnv=dummy, ManagerBasedEnv=dummy)
    _module("isaaclab.envs.mdp")
    _module("isaaclab.envs.mdp.actions", JointPositionAction=JointPositionAction, JointPositionActionCfg=_Cfg)
    from oracle import il_commands

    _module("isaaclab.managers", CurriculumTermCfg=_Cfg, CommandTerm=il_commands.CommandTerm, CommandTermCfg=il_commands.CommandTermCfg)
    _module("isaaclab.envs.mdp.commands", UniformVelocityCommand=il_commands.UniformVelocityCommand,
            UniformVelocityCommandCfg=il_commands.UniformVelocityCommandCfg)
    _INSTALLED = True


def _load_file(mod_name: str, rel_path: str):
    if mod_name in sys.modules and getattr(sys.modules[mod_name], "__file__", None):
        return sys.modules[mod_name]
    path = os.path.join(REFERENCE_ROOT, rel_path)
    spec = importlib.util.spec_from_file_location(mod_name, path)
    mod = importlib.util.module_from_spec(spec)
    sys.modules[mod_name] = mod
    spec.loader.exec_module(mod)
    return mod


def load_reference_mdp():
    """Returns ``(rewards, observations, terminations, actions)`` modules of reference ``locotouch/mdp``."""
    install_isaaclab_stub()
    _module("locotouch")
    _module("locotouch.mdp")
    actions = _load_file("locotouch.mdp.actions", "locotouch/mdp/actions.py")
    rewards = _load_file("locotouch.mdp.rewards", "locotouch/mdp/rewards.py")
    observations = _load_file("locotouch.mdp.observations", "locotouch/mdp/observations.py")
    terminations = _load_file("locotouch.mdp.terminations", "locotouch/mdp/terminations.py")
    return rewards, observations, terminations, actions


def load_reference_commands():
    """Returns ``(commands, curriculums)`` modules of reference ``locotouch/mdp`` (command terms over the [IL] base classes
    restated in ``oracle/il_commands.py``; the velocity curriculum term)."""
    install_isaaclab_stub()
    _module("locotouch")
    _module("locotouch.mdp")
    commands = _load_file("locotouch.mdp.commands", "locotouch/mdp/commands.py")
    curriculums = _load_file("locotouch.mdp.curriculums", "locotouch/mdp/curriculums.py")
    return commands, curriculums


def load_reference_loco_rl():
    """Returns the reference ``loco_rl`` package (PPO, RolloutStorage, ActorCritic, models)."""
    install_isaaclab_stub()
    path = os.path.join(REFERENCE_ROOT, "loco_rl")
    if path not in sys.path:
        sys.path.insert(0, path)
    # make sure we do not pick up the drop-in package of the same leaf name
    mod = importlib.import_module("loco_rl")
    assert os.path.realpath(mod.__file__).startswith(os.path.realpath(path)), mod.__file__
    return mod


def load_reference_distill():
    """Returns ``(student, tactile_recorder, distillation_cfg)`` modules of reference ``locotouch/distill``."""
    load_reference_loco_rl()
    _module("locotouch")
    _module("locotouch.config")
    _module("locotouch.config.locotouch")
    _module("locotouch.config.locotouch.agents")
    _module("locotouch.distill")
    cfg = _load_file(
        "locotouch.config.locotouch.agents.distillation_cfg", "locotouch/config/locotouch/agents/distillation_cfg.py"
    )
    recorder = _load_file("locotouch.distill.tactile_recorder", "locotouch/distill/tactile_recorder.py")
    _load_file("locotouch.distill.replay_buffer", "locotouch/distill/replay_buffer.py")
    student = _load_file("locotouch.distill.student", "locotouch/distill/student.py")
    return student, recorder, cfg
