"""K6 parity: fused PPO loss forward + backward vs autograd over the reference expressions (oracle/ppo.py)."""
import pytest
import torch

from oracle import ppo as OP
from tests import helpers as H

pytestmark = pytest.mark.gpu


def make_batch(B, A=12, seed=0, shift=0.3):
    g = torch.Generator().manual_seed(seed)
    old_mu = torch.randn(B, A, generator=g)
    old_sigma = (0.5 + torch.rand(A, generator=g)).expand(B, A).contiguous()
    actions = old_mu + old_sigma * torch.randn(B, A, generator=g)
    old_logp = (-((actions - old_mu) ** 2) / (2 * old_sigma**2) - old_sigma.log() - 0.9189385332046727).sum(-1)
    mu = old_mu + shift * torch.randn(B, A, generator=g)  # policy moved: some ratios leave the clip range
    sigma = old_sigma[0] * (1 + 0.1 * torch.randn(A, generator=g))
    value = torch.randn(B, generator=g)
    old_values = value + 0.3 * torch.randn(B, generator=g)  # some |V - V_old| > clip
    returns = torch.randn(B, generator=g)
    adv = torch.randn(B, generator=g)
    adv[::17] = 0.0
    return dict(mu=mu, sigma=sigma, value=value, actions=actions, old_logp=old_logp, old_mu=old_mu, old_sigma=old_sigma,
                advantages=adv, returns=returns, old_values=old_values)


def oracle_loss_and_grads(b, **cfg):
    mu = b["mu"].clone().requires_grad_(True)
    sigma = b["sigma"].clone().requires_grad_(True)
    value = b["value"].clone().unsqueeze(1).requires_grad_(True)
    res = OP.ppo_loss(mu, sigma, value, b["actions"], b["old_logp"].unsqueeze(1), b["old_mu"], b["old_sigma"], b["advantages"].unsqueeze(1),
                      b["returns"].unsqueeze(1), b["old_values"].unsqueeze(1), **cfg)
    res["loss"].backward()
    return res, mu.grad, sigma.grad, value.grad.squeeze(1)


@pytest.mark.parametrize("B,shift,clipped_value", [(768, 0.3, True), (24576, 0.1, True), (1000, 0.5, False), (3, 0.2, True)])
def test_ppo_loss_matches_autograd(cuda, lt_lib, B, shift, clipped_value):
    from locotouch_b200 import ops

    b = make_batch(B, seed=B, shift=shift)
    cfg = dict(clip_param=0.2, value_loss_coef=1.0, entropy_coef=0.01, use_clipped_value_loss=clipped_value)
    res, g_mu, g_sigma, g_value = oracle_loss_and_grads(b, **cfg)
    d = {k: v.to(cuda).contiguous() for k, v in b.items()}
    lr = torch.tensor([1e-3], device=cuda)
    out = ops.ppo_loss(**d, **cfg, desired_kl=0.01, lr=lr)
    torch.cuda.synchronize()
    o = out.out.cpu()
    H.assert_close(o[0], res["loss"].detach(), "loss")
    H.assert_close(o[1], res["surrogate"].detach(), "surrogate loss")
    H.assert_close(o[2], res["value_loss"].detach(), "value loss")
    H.assert_close(o[3], res["entropy_mean"].detach(), "entropy")
    H.assert_close(o[4], res["kl_mean"], "kl")
    # gradients: absolute floor scaled to 1/B (each per-sample gradient carries the 1/B of the batch mean)
    H.assert_close(out.grad_mu, g_mu, "dL/dmu", rtol=1e-5, atol=1e-6 / B * 10)
    H.assert_close(out.grad_value, g_value, "dL/dvalue", rtol=1e-5, atol=1e-6 / B * 10)
    H.assert_close(out.grad_sigma, g_sigma, "dL/dsigma", rtol=2e-5, atol=1e-6)
    # adaptive learning rate decision (ppo.py:275-281)
    H.assert_close(lr.cpu(), torch.tensor([OP.adaptive_lr(1e-3, float(res["kl_mean"]), 0.01)]), "learning rate", rtol=1e-6, atol=0)
    # deterministic: a second launch reproduces the first bit for bit
    out2 = ops.ppo_loss(**d, **cfg)
    H.assert_equal(out2.grad_sigma, out.grad_sigma, "run-to-run determinism of the reductions")
    H.assert_equal(out2.out[:5], out.out[:5], "run-to-run determinism of the losses")


def test_adaptive_lr_kernel_and_accumulators(cuda, lt_lib):
    from locotouch_b200 import ops

    for kl, want in ((0.05, 1e-3 / 1.5), (0.001, 1e-3 * 1.5), (0.01, 1e-3), (0.0, 1e-3)):
        lr = torch.tensor([1e-3], device=cuda)
        ops.adaptive_lr(torch.tensor([kl * 4], device=cuda), 0.25, 0.01, lr)
        H.assert_close(lr.cpu(), torch.tensor([want]), f"kl={kl}", rtol=1e-6, atol=0)
    b = {k: v.to(cuda).contiguous() for k, v in make_batch(512, seed=1).items()}
    acc = torch.zeros(4, device=cuda)
    for _ in range(3):
        out = ops.ppo_loss(**b, entropy_coef=0.01, loss_accum=acc)
    H.assert_close(acc.cpu(), torch.stack([out.out[2] * 3, out.out[1] * 3, out.out[3] * 3, torch.tensor(3.0, device=cuda)]).cpu(), "loss accumulators", rtol=1e-6)
