"""DAgger ReplayBuffer: device-side collection bookkeeping (K11), trajectory packing and batch padding
against the oracle and the vectors of the unmodified reference class, all driven by the same deterministic tape env."""
import os

import numpy as np
import pytest
import torch

from oracle.dagger import ReplayBufferOracle
from tests import helpers as H
from tests.golden.make_golden import DAGGER_SMALL, _DummyStudent

GOLDEN = os.path.join(os.path.dirname(__file__), "golden", "dagger_c4.npz")


def _check_against_golden(g, r1, l1, r2, l2, steps, trajs, lengths, flats, batches, tol=dict(rtol=1e-6, atol=1e-7)):
    assert np.allclose(np.array(r1), g["rewards1"], **tol) and list(l1) == list(g["lengths1"])
    assert np.allclose(np.array(r2), g["rewards2"], **tol) and list(l2) == list(g["lengths2"])
    assert steps == (int(g["steps1"]), int(g["steps2"])) and trajs == (int(g["trajs1"]), int(g["trajs2"]))
    assert list(lengths) == list(g["traj_lengths"])
    for name, f in zip(("flat_prop", "flat_teacher", "flat_tactile"), flats):
        assert np.array_equal(np.asarray(f), g[name]), name
    assert len(batches) == int(g["num_batches"])
    for i in (0, len(batches) - 1):
        b = batches[i]
        for key, name in (("proprioceptions", "prop"), ("teacher_encoder_obses", "teacher"), ("tactile_signals", "tactile"), ("masks", "masks")):
            assert np.array_equal(b[key].cpu().numpy(), g[f"b{i}_{name}"]), (i, key)


def test_oracle_matches_reference_golden():
    g, c = np.load(GOLDEN), DAGGER_SMALL
    env = H.TapeEnv(c["N"], c["steps"], obs_dim=c["P"] + 8, tactile_dim=c["tactile"])
    rb = ReplayBufferOracle(env, c["P"], c["tactile"])
    r1, l1 = rb.collect_data(lambda *a: None, c["first"], with_student=False)
    s1, t1 = rb.steps_count, len(rb.trajs)
    assert env.t == int(g["t1"])
    r2, l2 = rb.collect_data(lambda *a: None, c["second"], with_student=True)
    assert env.t == int(g["t2"]) and env.resets == 1
    np.random.seed(c["np_seed"])
    batches = rb.batches(c["batch"])
    flats = [torch.cat([t[k] for t in rb.trajs]).numpy() for k in range(3)]
    _check_against_golden(g, r1, l1, r2, l2, (s1, rb.steps_count), (t1, len(rb.trajs)), [t[0].shape[0] for t in rb.trajs], flats, batches)
    env2 = H.TapeEnv(c["N"], c["steps"], seed=5, obs_dim=c["P"] + 8, tactile_dim=c["tactile"], rsl_style=True)
    er, el = ReplayBufferOracle(env2, c["P"], c["tactile"]).evaluate(40)
    assert np.allclose(np.array(er), g["eval_rewards"], rtol=1e-6, atol=1e-7) and el == list(g["eval_lengths"])


@pytest.mark.gpu
def test_replay_buffer_matches_reference(cuda, lt_lib):
    from locotouch_b200.distill import ReplayBuffer, TactileRecorder

    g, c = np.load(GOLDEN), DAGGER_SMALL
    env = H.TapeEnv(c["N"], c["steps"], device=cuda, obs_dim=c["P"] + 8, tactile_dim=c["tactile"])
    rb = ReplayBuffer(env, TactileRecorder(cuda, c["N"], c["tactile"], 1, 2), c["P"])
    teacher = lambda x: torch.zeros(c["N"], 12, device=cuda)  # noqa: E731
    r1, l1 = rb.collect_data(teacher, None, c["first"])
    s1, t1 = rb.num_steps, rb.num_trajs
    assert env.t == int(g["t1"])
    student = _DummyStudent(c["N"], cuda)
    r2, l2 = rb.collect_data(teacher, student, c["second"])
    assert env.t == int(g["t2"]) and env.resets == 1 and student.resets > 0
    np.random.seed(c["np_seed"])
    batches = list(rb.to_recurrent_generator(c["batch"]))
    assert batches[0]["masks"].dtype == torch.bool
    _check_against_golden(g, r1, l1, r2, l2, (s1, rb.num_steps), (t1, rb.num_trajs), rb._lengths, [f.cpu().numpy() for f in rb._flat], batches)
    # Student.train_on_data derives its batch size from these two properties (reference student.py:113)
    assert rb.num_steps == sum(rb._lengths) and rb.num_trajs == len(rb._lengths)
    rb.clear_buffer()
    assert rb.num_trajs == 0 and rb.num_steps == 0 and float(rb._reward_sums.abs().sum()) == 0.0
    env2 = H.TapeEnv(c["N"], c["steps"], device=cuda, seed=5, obs_dim=c["P"] + 8, tactile_dim=c["tactile"], rsl_style=True)
    rb2 = ReplayBuffer(env2, TactileRecorder(cuda, c["N"], c["tactile"], 1, 2), c["P"])
    er, el = rb2.evaluate(student, 40)
    assert np.allclose(np.array(er), g["eval_rewards"], rtol=1e-6, atol=1e-7) and el == list(g["eval_lengths"])


@pytest.mark.gpu
@pytest.mark.parametrize("N,steps,budget", [(1, 60, 20), (405, 120, 3000), (2050, 40, 5000)])
def test_replay_buffer_matches_oracle_on_other_sizes(cuda, lt_lib, N, steps, budget):
    from locotouch_b200.distill import ReplayBuffer, TactileRecorder

    kw = dict(steps=steps, seed=N, p_done=0.08, obs_dim=20, tactile_dim=12)
    env_o, env_g = H.TapeEnv(N, **kw), H.TapeEnv(N, device=cuda, **kw)
    if N == 1:  # the single env must finish sometimes (TapeEnv pins env 0 to "done every step")
        env_o.dones[::2, 0] = False
        env_g.dones[::2, 0] = False
    oracle = ReplayBufferOracle(env_o, 14, 12)
    ro, lo = oracle.collect_data(lambda *a: None, budget, with_student=False)
    rb = ReplayBuffer(env_g, TactileRecorder(cuda, N, 12, 1, 2), 14)
    rg, lg = rb.collect_data(lambda x: torch.zeros(N, 12, device=cuda), None, budget)
    assert lg == lo and np.allclose(rg, ro, rtol=1e-5, atol=1e-6)
    assert rb.num_steps == oracle.steps_count and rb.num_trajs == len(oracle.trajs) and env_g.t == env_o.t
    for k in range(3):
        H.assert_equal(rb._flat[k].cpu(), torch.cat([t[k] for t in oracle.trajs]), f"packed store {k}")
    # size-independent property: padding then masking returns exactly the stored rows
    idx = np.arange(rb.num_trajs)
    b = rb._prepare_padded_sequence(idx)
    assert int(b["masks"].sum()) == rb.num_steps
    H.assert_equal(b["proprioceptions"].transpose(0, 1)[b["masks"].transpose(0, 1)], rb._flat[0], "masked rows == packed store")


@pytest.mark.gpu
def test_distillation_iteration_on_the_synthetic_transport_env(cuda, lt_lib):
    """collect_data -> to_recurrent_generator -> Student.train_on_batch on the hot-path env stand-in: shapes, bookkeeping
    invariants and a finite, decreasing behaviour-cloning loss."""
    from locotouch_b200.distill import DistillationRandCylinderCNNRNNMonCfg, ReplayBuffer, Student, TactileRecorder
    from locotouch_b200.sim.transport_env import SyntheticTransportEnv

    N = 64
    env = SyntheticTransportEnv(N, cuda, seed=1, max_episode_length=40)
    obs = env.get_observations()
    assert obs["policy"].shape == (N, 348) and obs["tactile"].shape == (N, 442)
    assert set(obs["tactile"].unique().tolist()) <= {0.0, 1.0}
    torch.manual_seed(0)
    teacher = torch.nn.Linear(348, 12).to(cuda)
    cfg = DistillationRandCylinderCNNRNNMonCfg(device=str(cuda))
    student = Student(cfg, 270, 442, 12, teacher_policy_inference=teacher)
    rb = ReplayBuffer(env, TactileRecorder(cuda, N, 442, 1, 2), 270)
    rewards, lengths = rb.collect_data(teacher, None, 1500)
    assert rb.num_steps >= 1500 and rb.num_trajs > 0 and len(rewards) == len(lengths) >= rb.num_trajs
    assert max(lengths) <= 40 and rb.num_steps == sum(rb._lengths)
    student.train()
    losses = []
    for _ in range(6):
        for b in rb.to_recurrent_generator(16):
            L, B = b["masks"].shape
            assert b["proprioceptions"].shape == (L, B, 270) and b["teacher_encoder_obses"].shape == (L, B, 78)
            assert b["tactile_signals"].shape == (L, B, 442)
            losses.append(float(student.train_on_batch(b).reshape(-1)[0]))
    assert all(np.isfinite(losses)) and np.mean(losses[-3:]) < np.mean(losses[:3])
    # a second, student-driven collection appends to the buffer (DAgger)
    before = rb.num_trajs
    student.eval()
    rb.collect_data(teacher, student, 500)
    assert rb.num_trajs > before
