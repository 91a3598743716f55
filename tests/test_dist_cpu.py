"""world_size-2 gloo tests (CPU) of the env-sharding logic of SURVEY.md 8e: averaged shard gradients == gradient on the
concatenated mini-batch, all-reduced advantage statistics == statistics of the concatenated rollout, and identical
learning-rate decisions on every rank."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from oracle import ppo as OP


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _batch(B, seed):
    g = torch.Generator().manual_seed(seed)
    A = 12
    old_mu = torch.randn(B, A, generator=g)
    old_sigma = (0.5 + torch.rand(A, generator=g)).expand(B, A).contiguous()
    actions = old_mu + old_sigma * torch.randn(B, A, generator=g)
    old_logp = (-((actions - old_mu) ** 2) / (2 * old_sigma**2) - old_sigma.log() - 0.9189385332046727).sum(-1, keepdim=True)
    return dict(mu=old_mu + 0.2 * torch.randn(B, A, generator=g), sigma=old_sigma[0] * 1.05, value=torch.randn(B, 1, generator=g), actions=actions,
                old_logp=old_logp, old_mu=old_mu, old_sigma=old_sigma, adv=torch.randn(B, 1, generator=g), returns=torch.randn(B, 1, generator=g),
                old_values=torch.randn(B, 1, generator=g))


def _grads(b):
    mu, sigma, value = b["mu"].clone().requires_grad_(True), b["sigma"].clone().requires_grad_(True), b["value"].clone().requires_grad_(True)
    res = OP.ppo_loss(mu, sigma, value, b["actions"], b["old_logp"], b["old_mu"], b["old_sigma"], b["adv"], b["returns"], b["old_values"])
    res["loss"].backward()
    return res, sigma.grad, mu.grad, value.grad


def _worker(rank, world, port, tmp):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from locotouch_b200 import dist as D

        assert D.world_info() == (rank, world)
        start, count = D.shard_envs(4096 * world)
        assert (start, count) == (rank * 4096, 4096)
        with pytest.raises(ValueError):
            D.shard_envs(4097)
        # (1) gradient of the shared parameter (sigma): mean over ranks of shard gradients == gradient on the union batch
        B = 512
        full = _batch(B * world, seed=1)
        shard = {k: (v[rank * B:(rank + 1) * B] if v.dim() == 2 and v.shape[0] == B * world else v) for k, v in full.items()}
        res, g_sigma, _, _ = _grads(shard)
        flat = g_sigma.clone()
        scale = D.average_gradients_(flat)
        _, g_full, _, _ = _grads(full)
        torch.testing.assert_close(flat * scale, g_full, rtol=1e-5, atol=1e-7)
        D.assert_same_on_all_ranks(flat, "all-reduced gradient")
        # (2) KL statistic -> identical learning-rate decision
        kl = res["kl_mean"].detach().clone().view(1)
        D.global_kl_mean_(kl)
        res_full, *_ = _grads(full)
        torch.testing.assert_close(kl[0], res_full["kl_mean"], rtol=1e-5, atol=1e-8)
        lr = torch.tensor([OP.adaptive_lr(1e-3, float(kl[0]), 0.01)])
        D.assert_same_on_all_ranks(lr, "learning rate")
        # (3) advantage statistics
        g = torch.Generator().manual_seed(7)
        adv_full = torch.randn(24, 64 * world, generator=g) * 3 + 0.5
        adv = adv_full[:, rank * 64:(rank + 1) * 64].double()
        stats = torch.tensor([adv.sum(), (adv * adv).sum(), adv.numel(), 0.0], dtype=torch.float64)
        D.reduce_adv_stats_(stats)
        mean, std = D.mean_and_unbiased_std(stats)
        assert abs(mean - float(adv_full.double().mean())) < 1e-9
        assert abs(std - float(adv_full.double().std())) < 1e-9
        with open(os.path.join(tmp, f"ok{rank}"), "w") as f:
            f.write("ok")
    finally:
        dist.destroy_process_group()


def test_env_sharding_world_size_2(tmp_path):
    world = 2
    port = _free_port()
    mp.spawn(_worker, args=(world, port, str(tmp_path)), nprocs=world, join=True)
    assert all((tmp_path / f"ok{r}").exists() for r in range(world))


def test_peer_gradient_exchange_validates_its_arguments():
    """The C entry point of K14 rejects bad arguments before any CUDA call."""
    import ctypes as C

    from locotouch_b200 import _C

    lib = _C.lib()
    ptrs = (C.c_void_p * 2)(None, None)
    assert lib.lt_peer_sum_clip_adam(None, ptrs, 2, None, 4, None, None, 16, None, None, 1.0, 0.9, 0.999, 1e-8, 0.0, 0.5, 0.01, 0.5, None, None, 0, None) == 1


def test_random_streams_of_one_rank_and_of_neighbouring_ranks_never_share_a_philox_key():
    """engine.stream_seed: exploration noise, observation noise and taxel dropout of every (seed, rank) draw from distinct keys."""
    from locotouch_b200.engine import stream_seed

    keys = {}
    for seed in range(0, 40):
        for rank in range(0, 16):
            for consumer in range(3):
                k = stream_seed(seed, rank, consumer)
                assert 0 <= k < 2**63
                assert k not in keys, (seed, rank, consumer, keys[k])
                keys[k] = (seed, rank, consumer)


def test_state_upload_layout_carries_only_what_the_path_reads():
    """engine.pack_host: one contiguous host buffer; the per-step upload prefix holds every per-step tensor the path reads whole, the
    feet rows of body_pos_w / body_lin_vel_w and the taxel-body rows of body_quat_w compact; articulation constants and unread
    tensors sit behind the prefix (uploaded once)."""
    import torch

    from locotouch_b200.engine import STATIC, UNREAD, pack_host, row_subsets
    from locotouch_b200.sim import synth

    env = synth.make_env(16, seed=3, with_object=True, with_tactile=True)
    layout, compact, upload_bytes, host = pack_host(env)
    tensors = env.named_tensors()
    subsets = row_subsets(env)
    assert list(subsets["robot.body_quat_w"]) == list(range(17, 238)) and len(subsets["robot.body_pos_w"]) == 4
    for name, off, nbytes, dtype, shape in layout:
        got = host[off:off + nbytes].view(dtype).view(shape)
        assert torch.equal(got, tensors[name]), name
        behind = name in UNREAD or name in STATIC or name in subsets
        assert (off >= upload_bytes) == behind, f"{name}: wrong side of the upload prefix"
    for name, off, nbytes, dtype, shape in compact:
        assert off + nbytes <= upload_bytes
        assert torch.equal(host[off:off + nbytes].view(dtype).view(shape), tensors[name][:, subsets[name]]), name
    per_env = upload_bytes / 16
    assert 7000 < per_env < 8200, per_env  # 7 476 B of payload + 256-byte slice alignment at 16 envs
    # no tactile sensor: body_quat_w is not read at all, nothing of it travels
    env2 = synth.make_env(16, seed=3)
    _, compact2, ub2, _ = pack_host(env2)
    assert [c[0] for c in compact2] == ["robot.body_pos_w", "robot.body_lin_vel_w"]
