"""CPU checks of the drop-in boundary: the C-ABI library builds, loads, and exports every symbol the header declares."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "locotouch_b200.h")


def declared_functions():
    text = open(HEADER).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(lt_[a-z0-9_]+)\s*\(", text)))


def test_header_declares_the_survey_entry_points():
    names = declared_functions()
    for want in ("lt_mdp_step", "lt_mdp_reset", "lt_taxel_synth", "lt_tactile_delay", "lt_store_step", "lt_gae", "lt_gather_rows",
                 "lt_ppo_loss", "lt_clip_adam", "lt_masked_mse", "lt_act_sample", "lt_pad_trajectories"):
        assert want in names


def test_library_exports_every_declared_symbol(lt_lib):
    from locotouch_b200 import _C

    raw = ctypes.CDLL(_C.LIB_PATH)
    for name in declared_functions():
        assert hasattr(raw, name), f"{name} is declared in include/locotouch_b200.h but not exported"
        assert name in _C.EXPORTED_SYMBOLS, f"{name} has no ctypes signature in _C.py"


def test_struct_layouts_match(lt_lib):
    from locotouch_b200 import _C

    for which, struct in enumerate((_C.LtGatherArgs, _C.LtPpoLossArgs, _C.LtTaxelArgs, _C.LtMdpArgs, _C.LtGaitState, _C.LtGaitParams, _C.LtTaxelForceArgs,
                                    _C.LtCommandRanges, _C.LtCommandArgs, _C.LtVelCurriculumArgs, _C.LtPpoHeadsArgs, _C.LtStudentCnnArgs, _C.LtMlp3Net)):
        assert lt_lib.lt_struct_size(which) == ctypes.sizeof(struct), struct.__name__
    assert lt_lib.lt_struct_size(99) == -1


def test_error_strings_and_workspace_queries(lt_lib):
    assert lt_lib.lt_abi_version() == 1
    assert lt_lib.lt_error_string(0) == b"ok"
    assert b"workspace" in lt_lib.lt_error_string(3)
    assert lt_lib.lt_gae_workspace_bytes(24, 4096) >= 16 + 2 * 32 * 8
    assert lt_lib.lt_ppo_loss_workspace_bytes(24576, 12) >= 384 * 15 * 4
    assert lt_lib.lt_clip_adam_workspace_bytes(607641) >= 1024 * 8
    assert lt_lib.lt_masked_mse_workspace_bytes(20000) > 0


def test_invalid_arguments_are_rejected_without_a_gpu(lt_lib):
    # argument validation happens before any CUDA call, so it can be exercised on the CPU box
    assert lt_lib.lt_gae(None, None, None, None, None, None, 24, 64, 0.99, 0.95, 1, None, 0, None) == 3  # workspace
    assert lt_lib.lt_adv_normalize(None, 10, None, None) == 1
    assert lt_lib.lt_ppo_loss(None, None) == 1
    assert lt_lib.lt_taxel_synth(None, None) == 1
    assert lt_lib.lt_mdp_step(None, None) == 1
    assert lt_lib.lt_clip_adam(None, None, None, None, 10, None, None, 1.0, 0.9, 0.999, 1e-8, 0.0, 1.0, None, None, 0, None) == 1


def test_product_path_refuses_cpu_tensors(lt_lib):
    import torch

    from locotouch_b200 import _C, ops

    with pytest.raises(_C.LocoTouchLibraryError):
        ops.gae(torch.zeros(4, 8), torch.zeros(4, 8), torch.zeros(4, 8, dtype=torch.uint8), torch.zeros(8), 0.99, 0.95)


def test_product_never_imports_the_oracle():
    """The oracle is test infrastructure: nothing under locotouch_b200/ may import it."""
    pkg = os.path.join(ROOT, "locotouch_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith(".py"):
                src = open(os.path.join(dirpath, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", src, flags=re.M), os.path.join(dirpath, f)


def test_mdp_lookup_tables_are_built_on_the_host(lt_lib):
    """lt_mdp_build_tables: observation column map ([term][history][dim] rows), per-value term info, reward kind -> slot and
    the per-kind parameter rows -- host arithmetic only, checked against a direct restatement."""
    import ctypes as C
    import struct

    from locotouch_b200 import _C
    from locotouch_b200.mdp import task_spec as TS

    RK_COUNT = TS.RK_OBJ_DANGER + 1  # LT_RK_COUNT
    a = _C.LtMdpArgs()
    a.J = 12
    dims, H = [3, 12, 3], 4
    a.num_obs_terms, a.history_length = len(dims), H
    for i, (kind, d) in enumerate(zip((0, 3, 2), dims)):  # LT_OK_COMMAND, LT_OK_JOINT_POS_REL, LT_OK_PROJECTED_GRAVITY
        a.obs_terms[i].kind, a.obs_terms[i].dim, a.obs_terms[i].scale = kind, d, 1.0
    kinds = [(TS.RK_ALIVE, 1.0, ()), (TS.RK_TRACK_LIN_VEL_XY, 2.0, (0.25,)), (TS.RK_FOOT_SLIP, 0.0, (0.5,)), (TS.RK_BASE_HEIGHT, -1.0, (0.42,))]
    a.num_reward_terms = len(kinds)
    for i, (k, w, p) in enumerate(kinds):
        a.reward_terms[i].kind, a.reward_terms[i].weight = k, w
        for j, v in enumerate(p):
            a.reward_terms[i].p[j] = v
    n = lt_lib.lt_mdp_tables_len(C.byref(a))
    dps, D = sum(dims), sum(dims) * H
    assert n == (((D + 3) // 4 * 4) + dps + RK_COUNT * 7 + 3) // 4 * 4
    out = (C.c_int32 * n)()
    assert lt_lib.lt_mdp_build_tables(C.byref(a), out, n) == 0
    assert lt_lib.lt_mdp_build_tables(C.byref(a), out, n - 1) == 1  # too small
    out = list(out)
    col = jb = 0
    for t, d in enumerate(dims):
        for h in range(H):
            for i in range(d):
                k = col + h * d + i
                m = out[k] & 0xFFFFFFFF
                assert m & 0xFFFF == jb + i
                assert m >> 16 == (0xFFFF if h == H - 1 else k + d)
        for i in range(d):
            assert out[(D + 3) // 4 * 4 + jb + i] == t | (i << 8)
        col += d * H
        jb += d
    slot0 = (D + 3) // 4 * 4 + dps
    slots = out[slot0:slot0 + RK_COUNT]
    assert slots[TS.RK_ALIVE] == 0 and slots[TS.RK_TRACK_LIN_VEL_XY] == 1 and slots[TS.RK_BASE_HEIGHT] == 3
    assert slots[TS.RK_FOOT_SLIP] == -1 and slots[TS.RK_GAIT] == -1  # zero weight / absent
    par = out[slot0 + RK_COUNT:]
    as_float = lambda v: struct.unpack("f", struct.pack("i", v))[0]  # noqa: E731
    assert as_float(par[TS.RK_TRACK_LIN_VEL_XY * 6]) == 0.25
    assert abs(as_float(par[TS.RK_BASE_HEIGHT * 6]) - 0.42) < 1e-7
    assert as_float(par[TS.RK_FOOT_SLIP * 6]) == 0.0


def test_actor_critic_rejects_action_widths_the_kernels_do_not_take():
    """ADVICE (round 1): lt_act_sample / lt_ppo_loss need num_actions % 4 == 0 and <= 64 -- fail at construction with a clear message,
    not at the first act() with LT_ERR_INVALID_ARG."""
    import pytest

    from locotouch_b200.loco_rl import ActorCritic

    for bad in (6, 13, 68):
        with pytest.raises(ValueError, match="num_actions"):
            ActorCritic(8, 8, bad, [8], [8])
    ActorCritic(8, 8, 12, [8], [8])
