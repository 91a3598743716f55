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


@pytest.mark.parametrize("B,Hd,A,clipped_value", [(24576, 128, 12, True), (1000, 128, 12, False), (333, 256, 8, True), (5, 128, 16, True), (4099, 128, 4, True)])
def test_fused_heads_loss_matches_autograd_through_the_heads(cuda, lt_lib, B, Hd, A, clipped_value):
    """K16: head layers of both MLPs + PPO loss + head dgrad with the ELU backward of the last hidden layer in one kernel, against
    autograd over the reference expressions (oracle/ppo.py = reference ppo.py:252-302) applied to elu(z) @ W^T + b in float64 /
    float32: losses, dL/dmu, dL/dV, dL/dsigma, the gradient w.r.t. the pre-activations z, and the learning-rate decision."""
    from locotouch_b200 import ops

    g = torch.Generator().manual_seed(B + Hd + A)
    b = make_batch(B, A=A, seed=B, shift=0.3)
    z_a, z_c = torch.randn(B, Hd, generator=g), torch.randn(B, Hd, generator=g)
    w_a, b_a = 0.3 * torch.randn(A, Hd, generator=g) / Hd ** 0.5, 0.1 * torch.randn(A, generator=g)
    w_c, b_c = torch.randn(1, Hd, generator=g) / Hd ** 0.5, 0.1 * torch.randn(1, generator=g)
    # the policy mean of make_batch() is realised through the head: shift the bias-free part onto old_mu so that ratios stay O(1)
    za = z_a.clone().requires_grad_(True)
    zc = z_c.clone().requires_grad_(True)
    wa, ba, wc, bc = (t.clone().requires_grad_(True) for t in (w_a, b_a, w_c, b_c))
    sigma = b["sigma"].clone().requires_grad_(True)
    h_a, h_c = torch.nn.functional.elu(za), torch.nn.functional.elu(zc)
    mu = b["old_mu"] + h_a @ wa.t() + ba
    value = h_c @ wc.t() + bc
    cfg = dict(clip_param=0.2, value_loss_coef=1.0, entropy_coef=0.01, use_clipped_value_loss=clipped_value)
    res = OP.ppo_loss(mu, sigma, value, b["actions"], b["old_logp"].unsqueeze(1), b["old_mu"], b["old_sigma"], b["advantages"].unsqueeze(1),
                      b["returns"].unsqueeze(1), b["old_values"].unsqueeze(1), **cfg)
    mu.retain_grad()
    value.retain_grad()
    res["loss"].backward()

    d = {k: v.to(cuda).contiguous() for k, v in b.items() if k not in ("mu", "value")}
    # mu = old_mu + head(h): fold old_mu into the kernel's view by handing it the head with zero "old_mu offset" is not possible, so
    # the comparison shifts the actions / old_mu instead: (a - mu) and (old_mu - mu) only see differences
    shift = b["old_mu"].to(cuda)
    d["actions"] = (d["actions"] - shift).contiguous()
    d["old_mu"] = torch.zeros_like(shift)
    ha, hc = h_a.detach().to(cuda).contiguous(), h_c.detach().to(cuda).contiguous()
    g_ha, g_hc = torch.empty_like(ha), torch.empty_like(hc)
    bufs = ops.PpoLossBuffers(B, A, cuda)
    lr = torch.tensor([1e-3], device=cuda)
    mu_out, v_out = torch.empty(B, A, device=cuda), torch.empty(B, device=cuda)
    ops.ppo_heads_loss(ha, hc, w_a.to(cuda), b_a.to(cuda), w_c.to(cuda), b_c.to(cuda), d["sigma"], d["actions"], d["old_logp"], d["old_mu"],
                       d["old_sigma"], d["advantages"], d["returns"], d["old_values"], g_ha, g_hc, **cfg, desired_kl=0.01, lr=lr, buffers=bufs,
                       mu_out=mu_out, value_out=v_out)
    torch.cuda.synchronize()
    H.assert_close(mu_out.cpu() + b["old_mu"], mu.detach(), "mu through the fused head", rtol=1e-5, atol=1e-5)
    H.assert_close(v_out.cpu(), value.detach().squeeze(1), "value through the fused head", rtol=1e-5, atol=1e-5)
    o = bufs.out.cpu()
    H.assert_close(o[0], res["loss"].detach(), "loss", rtol=2e-5)
    H.assert_close(o[1], res["surrogate"].detach(), "surrogate loss", rtol=2e-5, atol=2e-6)
    H.assert_close(o[2], res["value_loss"].detach(), "value loss", rtol=2e-5)
    H.assert_close(o[4], res["kl_mean"], "kl", rtol=2e-5)
    H.assert_close(bufs.grad_mu, mu.grad, "dL/dmu", rtol=1e-4, atol=2e-5 / B)
    H.assert_close(bufs.grad_value, value.grad.squeeze(1), "dL/dvalue", rtol=1e-4, atol=2e-5 / B)
    H.assert_close(bufs.grad_sigma, sigma.grad, "dL/dsigma", rtol=1e-4, atol=1e-6)
    H.assert_close(g_ha, za.grad, "dL/dz_actor (head dgrad x elu')", rtol=1e-4, atol=2e-5 / B)
    H.assert_close(g_hc, zc.grad, "dL/dz_critic (head dgrad x elu')", rtol=1e-4, atol=2e-5 / B)
    H.assert_close(lr.cpu(), torch.tensor([OP.adaptive_lr(1e-3, float(res["kl_mean"]), 0.01)]), "learning rate", rtol=1e-6, atol=0)
    # deterministic statistics
    ops.ppo_heads_loss(ha, hc, w_a.to(cuda), b_a.to(cuda), w_c.to(cuda), b_c.to(cuda), d["sigma"], d["actions"], d["old_logp"], d["old_mu"],
                       d["old_sigma"], d["advantages"], d["returns"], d["old_values"], g_ha, g_hc, **cfg, buffers=bufs)
    H.assert_equal(bufs.out[:5].cpu(), o[:5], "run-to-run determinism of the losses")
