"""Recurrent mini-batch path: trajectory split / pad / unpad (K10) and
RolloutStorage.recurrent_mini_batch_generator against the oracle and the vectors of the unmodified reference."""
import os

import numpy as np
import pytest
import torch

from oracle import recurrent as R
from tests import helpers as H
from tests.golden.make_golden import RECURRENT_SMALL, recurrent_inputs

GOLDEN = os.path.join(os.path.dirname(__file__), "golden", "recurrent_c1.npz")


# --------------------------------------------------------------------------------------------------------- CPU: oracle
def test_oracle_matches_reference_golden():
    g = np.load(GOLDEN)
    c, r = RECURRENT_SMALL, recurrent_inputs()
    dones = r["dones"][:, :, 0].numpy()
    padded, masks = R.split_and_pad_trajectories(r["obs"].numpy(), dones)
    assert padded.shape == g["padded"].shape
    assert np.array_equal(padded, g["padded"]) and np.array_equal(masks, g["masks"])
    assert np.array_equal(R.unpad_trajectories(padded, masks), r["obs"].numpy())
    cpad, _ = R.split_and_pad_trajectories(r["critic_obs"].numpy(), dones)
    hid_a = R.first_step_hidden(r["hid_a"].numpy(), dones)
    hid_c = R.first_step_hidden(r["hid_c"].numpy(), dones)
    for i, (start, stop, first, last) in enumerate(R.recurrent_mini_batches(dones, c["N"], c["num_mini_batches"])):
        assert np.array_equal(padded[:, first:last], g[f"mb{i}_obs"])
        assert np.array_equal(cpad[:, first:last], g[f"mb{i}_cobs"])
        assert np.array_equal(masks[:, first:last], g[f"mb{i}_masks"])
        assert np.array_equal(hid_a[:, first:last], g[f"mb{i}_hid_a"]) and np.array_equal(hid_c[:, first:last], g[f"mb{i}_hid_c"])
        assert np.array_equal(r["actions"][:, start:stop].numpy(), g[f"mb{i}_actions"])


def test_oracle_edge_cases():
    # every step done -> T*N trajectories of length 1; never done -> N trajectories of length T
    x = np.arange(3 * 2 * 2, dtype=np.float32).reshape(3, 2, 2)
    p, m = R.split_and_pad_trajectories(x, np.ones((3, 2), dtype=bool))
    assert p.shape == (3, 6, 2) and m[0].all() and not m[1:].any()
    p, m = R.split_and_pad_trajectories(x, np.zeros((3, 2), dtype=bool))
    assert p.shape == (3, 2, 2) and m.all() and np.array_equal(p, x)


# --------------------------------------------------------------------------------------------------------- GPU: kernels
@pytest.mark.gpu
def test_split_pad_unpad_match_reference_golden(cuda, lt_lib):
    from locotouch_b200.loco_rl.utils import split_and_pad_trajectories, unpad_trajectories

    g = np.load(GOLDEN)
    r = recurrent_inputs()
    obs, dones = r["obs"].to(cuda), r["dones"].to(cuda)
    padded, masks = split_and_pad_trajectories(obs, dones)
    assert masks.dtype == torch.bool
    H.assert_equal(padded.cpu(), torch.from_numpy(g["padded"]), "padded trajectories")
    H.assert_equal(masks.cpu(), torch.from_numpy(g["masks"]), "trajectory masks")
    H.assert_equal(unpad_trajectories(padded, masks), obs, "unpad(split_and_pad(x)) == x")
    # dones as stored by RolloutStorage (uint8) and as int64 give the same cut
    p8, _ = split_and_pad_trajectories(obs, dones.byte())
    H.assert_equal(p8, padded, "uint8 dones")


@pytest.mark.gpu
@pytest.mark.parametrize("T,N,D", [(1, 1, 1), (24, 1, 7), (5, 33, 4), (24, 4097, 30), (48, 257, 270)])
def test_split_pad_matches_oracle_on_ragged_sizes(cuda, lt_lib, T, N, D):
    from locotouch_b200.loco_rl.utils import split_and_pad_trajectories, unpad_trajectories

    g = torch.Generator().manual_seed(T * 1000 + N)
    x = torch.randn(T, N, D, generator=g)
    dones = torch.rand(T, N, 1, generator=g) < 0.1
    padded, masks = split_and_pad_trajectories(x.to(cuda), dones.to(cuda))
    ref_p, ref_m = R.split_and_pad_trajectories(x.numpy(), dones[:, :, 0].numpy())
    H.assert_equal(padded.cpu(), torch.from_numpy(ref_p), "padded")
    H.assert_equal(masks.cpu(), torch.from_numpy(ref_m), "masks")
    H.assert_equal(unpad_trajectories(padded, masks).cpu(), x, "round trip")
    # size-independent properties: every step appears exactly once; padding is zero
    assert int(masks.sum()) == T * N
    assert float(padded[~masks].abs().sum()) == 0.0


@pytest.mark.gpu
def test_recurrent_mini_batch_generator_matches_reference(cuda, lt_lib):
    from locotouch_b200.loco_rl import RolloutStorage

    g = np.load(GOLDEN)
    c, r = RECURRENT_SMALL, recurrent_inputs()
    st = RolloutStorage(c["N"], c["T"], [c["D"]], [c["D"] + 2], [c["A"]], device=cuda)
    zeros = torch.zeros(c["N"], device=cuda)
    for t in range(c["T"]):
        tr = RolloutStorage.Transition()
        tr.observations, tr.critic_observations, tr.actions = r["obs"][t].to(cuda), r["critic_obs"][t].to(cuda), r["actions"][t].to(cuda)
        tr.rewards, tr.dones, tr.values = zeros, r["dones"][t, :, 0].to(cuda), zeros.view(-1, 1)
        tr.actions_log_prob, tr.action_mean, tr.action_sigma = zeros, tr.actions, tr.actions.abs()
        tr.hidden_states = (r["hid_a"][t].to(cuda), r["hid_c"][t].to(cuda))
        st.add_transitions(tr)
    batches = list(st.recurrent_mini_batch_generator(c["num_mini_batches"], num_epochs=2))
    assert len(batches) == 2 * c["num_mini_batches"]
    for i, b in enumerate(batches):
        obs_b, cobs_b, act_b, val_b, adv_b, ret_b, logp_b, mu_b, sig_b, (hid_a, hid_c), masks_b, rnd_b = b
        k = i % c["num_mini_batches"]
        H.assert_equal(obs_b.cpu(), torch.from_numpy(g[f"mb{k}_obs"]), f"mini-batch {k} observations")
        H.assert_equal(cobs_b.cpu(), torch.from_numpy(g[f"mb{k}_cobs"]), f"mini-batch {k} critic observations")
        H.assert_equal(masks_b.cpu(), torch.from_numpy(g[f"mb{k}_masks"]), f"mini-batch {k} masks")
        H.assert_equal(hid_a.cpu(), torch.from_numpy(g[f"mb{k}_hid_a"]), f"mini-batch {k} actor hidden states")
        H.assert_equal(hid_c.cpu(), torch.from_numpy(g[f"mb{k}_hid_c"]), f"mini-batch {k} critic hidden states")
        H.assert_equal(act_b.cpu(), torch.from_numpy(g[f"mb{k}_actions"]), f"mini-batch {k} actions")
        assert rnd_b is None and val_b.shape[1] == c["N"] // c["num_mini_batches"]
