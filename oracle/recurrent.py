"""TEST INFRASTRUCTURE ONLY -- CPU restatement (plain loops) of the reference's trajectory split / pad / unpad and of the
recurrent mini-batch selection.  Never imported by the product path (locotouch_b200/).

Follows loco_rl/loco_rl/utils/utils.py:37-83 (split_and_pad_trajectories, unpad_trajectories) and
loco_rl/loco_rl/storage/rollout_storage.py:246-318 (recurrent_mini_batch_generator).  Pinned against the unmodified
reference functions by tests/golden/recurrent_c1.npz (tests/golden/make_golden.py::golden_recurrent)."""
from __future__ import annotations

import numpy as np


def trajectory_index(dones: np.ndarray):
    """dones [T, N] (any truthy dtype) -> list of (env, first step, length) in the reference's order: env by env, in time order
    (utils.py:58 transposes to [N, T] before flattening); the last step always closes a trajectory (utils.py:55-56)."""
    T, N = dones.shape
    out = []
    for n in range(N):
        start = 0
        for t in range(T):
            if t == T - 1 or dones[t, n]:
                out.append((n, start, t + 1 - start))
                start = t + 1
    return out


def split_and_pad_trajectories(x: np.ndarray, dones: np.ndarray):
    """x [T, N, D] -> padded [T, M, D] (zeros after each trajectory's end), masks [T, M] (utils.py:37-73)."""
    T, N, D = x.shape
    index = trajectory_index(dones.reshape(T, N))
    M = len(index)
    padded = np.zeros((T, M, D), dtype=x.dtype)
    masks = np.zeros((T, M), dtype=bool)
    for j, (n, start, length) in enumerate(index):
        padded[:length, j] = x[start:start + length, n]
        masks[:length, j] = True
    return padded, masks


def unpad_trajectories(padded: np.ndarray, masks: np.ndarray):
    """utils.py:76-83: valid rows, trajectory by trajectory, re-tile the env-major flattened rollout."""
    T, M, D = padded.shape
    rows = [padded[:int(masks[:, j].sum()), j] for j in range(M)]
    flat = np.concatenate(rows, axis=0)
    return flat.reshape(-1, T, D).transpose(1, 0, 2)


def recurrent_mini_batches(dones: np.ndarray, num_envs: int, num_mini_batches: int):
    """Per mini-batch (an env range, rollout_storage.py:262-271): (env start, env stop, first trajectory, last trajectory)."""
    T, N = dones.shape
    index = trajectory_index(dones)
    per_env = np.zeros(N + 1, dtype=np.int64)
    for n, _, _ in index:
        per_env[n + 1] += 1
    base = np.cumsum(per_env)
    size = num_envs // num_mini_batches
    return [(i * size, (i + 1) * size, int(base[i * size]), int(base[(i + 1) * size])) for i in range(num_mini_batches)]


def first_step_hidden(saved: np.ndarray, dones: np.ndarray):
    """saved [T, L, N, H] -> [L, M, H]: the RNN state at every trajectory's first step (rollout_storage.py:291-299)."""
    index = trajectory_index(dones)
    return np.stack([saved[start, :, n] for n, start, _ in index], axis=1)
