"""K3 / K5 / K7 parity: act epilogue, transition store, mini-batch gather, clip + Adam."""
import pytest
import torch

from oracle import ppo as OP
from tests import helpers as H

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("n", [46000, 1001, 607641, 5])
def test_clip_adam_matches_torch_semantics(cuda, lt_lib, n):
    from locotouch_b200 import ops

    g = torch.Generator().manual_seed(n)
    n_pad = (n + 3) // 4 * 4
    p = torch.randn(n, generator=g)
    p_o = [p.clone()]
    m_o, v_o = [torch.zeros(n)], [torch.zeros(n)]
    pd, md, vd = (torch.zeros(n_pad, device=cuda) for _ in range(3))
    pd[:n] = p.to(cuda)
    lr = torch.tensor([1e-3], device=cuda)
    step = torch.zeros(1, device=cuda)
    norm = torch.zeros(1, device=cuda)
    step_o = 0
    for it in range(4):
        scale = [3.0, 0.01, 1.0, 30.0][it]  # clipped, unclipped, ...
        grad = torch.randn(n, generator=g) * scale / (n**0.5)
        total, step_o = OP.clip_and_adam(p_o, [grad.clone()], m_o, v_o, step_o, 1e-3, 1.0)
        gd = torch.zeros(n_pad, device=cuda)
        gd[:n] = grad.to(cuda)
        ops.clip_adam(pd[:n], gd[:n], md[:n], vd[:n], lr, step, max_grad_norm=1.0, grad_norm_out=norm)
        torch.cuda.synchronize()
        H.assert_close(norm.cpu()[0], total, f"grad norm it={it}", rtol=1e-5)
        H.assert_close(pd[:n], p_o[0], f"params it={it}", rtol=1e-5, atol=1e-6)
        H.assert_close(md[:n], m_o[0], f"exp_avg it={it}", rtol=1e-5, atol=1e-7)
        H.assert_close(vd[:n], v_o[0], f"exp_avg_sq it={it}", rtol=1e-5, atol=1e-9)
    assert float(step.cpu()) == 4.0


def test_adamw_decay_and_grad_scale(cuda, lt_lib):
    from locotouch_b200 import ops

    n = 4096
    g = torch.Generator().manual_seed(0)
    p, grad = torch.randn(n, generator=g), torch.randn(n, generator=g)
    ref = torch.nn.Parameter(p.clone())
    opt = torch.optim.AdamW([ref], lr=5e-4)  # student.py:82 defaults: weight_decay=1e-2
    ref.grad = grad.clone() * 0.5
    opt.step()
    pd, gd = p.to(cuda), grad.to(cuda)
    md, vd = torch.zeros(n, device=cuda), torch.zeros(n, device=cuda)
    ops.clip_adam(pd, gd, md, vd, torch.tensor([5e-4], device=cuda), torch.zeros(1, device=cuda), max_grad_norm=None,
                  weight_decay=1e-2, grad_scale=0.5)
    H.assert_close(pd, ref.detach(), "AdamW step", rtol=1e-5, atol=1e-6)


def test_act_sample_matches_oracle(cuda, lt_lib):
    from locotouch_b200 import ops

    g = torch.Generator().manual_seed(3)
    for n in (1, 405, 4097):
        mu, sigma, eps = torch.randn(n, 12, generator=g), 0.3 + torch.rand(12, generator=g), torch.randn(n, 12, generator=g)
        a_o, lp_o = OP.act_sample(mu, sigma, eps)
        mu_out, sig_out = torch.empty(n, 12, device=cuda), torch.empty(n, 12, device=cuda)
        a, lp = ops.act_sample(mu.to(cuda), sigma.to(cuda), eps.to(cuda), mu_out=mu_out, sigma_out=sig_out)
        H.assert_equal(a, a_o, "actions (bit-exact)")
        H.assert_close(lp, lp_o, "log prob", rtol=1e-5, atol=1e-5)
        H.assert_equal(mu_out, mu, "mu copy")
        H.assert_equal(sig_out, sigma.expand(n, 12), "sigma rows")
    # Philox path: moments of the standardised sample
    mu = torch.zeros(16384, 12, device=cuda)
    a, _ = ops.act_sample(mu, torch.ones(12, device=cuda), None, seed=5, offset=9)
    assert abs(float(a.mean())) < 0.01 and abs(float(a.std()) - 1.0) < 0.01
    a2, _ = ops.act_sample(mu, torch.ones(12, device=cuda), None, seed=5, offset=9)
    H.assert_equal(a2, a, "counter RNG is reproducible")
    a3, _ = ops.act_sample(mu, torch.ones(12, device=cuda), None, seed=5, offset=10)
    assert not torch.equal(a3, a)


def test_store_step_and_gather(cuda, lt_lib):
    from locotouch_b200 import ops

    g = torch.Generator().manual_seed(4)
    n = 1000
    r, v = torch.randn(n, generator=g), torch.randn(n, 1, generator=g)
    dones = (torch.rand(n, generator=g) < 0.1).long()
    to = (torch.rand(n, generator=g) < 0.05)
    obs, cobs = torch.randn(n, 270, generator=g), torch.randn(n, 348, generator=g)
    r_out, d_out = torch.empty(n, 1, device=cuda), torch.empty(n, 1, device=cuda, dtype=torch.uint8)
    o_out, c_out = torch.empty(n, 270, device=cuda), torch.empty(n, 348, device=cuda)
    ops.store_step(r.to(cuda), dones.to(cuda), to.to(cuda), v.to(cuda), 0.99, r_out, d_out, obs.to(cuda), o_out, cobs.to(cuda), c_out)
    H.assert_equal(r_out[:, 0], OP.bootstrap_rewards(r, v, to, 0.99), "bootstrapped rewards")
    H.assert_equal(d_out[:, 0], dones.byte(), "dones")
    H.assert_equal(o_out, obs, "obs rows")
    H.assert_equal(c_out, cobs, "critic obs rows")
    # gather: 9 tensors of the PPO mini-batch in one launch (rollout_storage.py:221-231)
    rows = 24 * 64
    srcs = [torch.randn(rows, d, generator=g) for d in (270, 270, 12, 1, 1, 1, 1, 12, 12)]
    idx = torch.randperm(rows, generator=g)[:500]
    outs = ops.gather_rows([s.to(cuda) for s in srcs], idx.to(cuda))
    for s, o in zip(srcs, outs):
        H.assert_equal(o, s[idx], "gathered rows")


@pytest.mark.parametrize("B,n", [(24576, 512), (1000, 256), (333, 128), (4097, 12), (500, 1), (7, 6)])
def test_bias_act_bwd_matches_autograd(cuda, lt_lib, B, n):
    from locotouch_b200 import ops

    g = torch.Generator().manual_seed(B + n)
    z = torch.randn(B, n, generator=g, requires_grad=True)
    h = torch.nn.functional.elu(z)
    go = torch.randn(B, n, generator=g)
    (h * go).sum().backward()
    bias = torch.empty(n, device=cuda)
    gd = go.to(cuda).clone()
    out = ops.bias_act_bwd(gd, h.detach().to(cuda), bias)
    assert out.data_ptr() == gd.data_ptr()
    H.assert_close(out, z.grad, "grad wrt pre-activation", rtol=1e-5, atol=1e-7)
    H.assert_close(bias, z.grad.sum(0), "bias gradient", rtol=2e-5, atol=1e-5 * (B ** 0.5))
    bias2 = torch.empty(n, device=cuda)
    ops.bias_act_bwd(go.to(cuda), None, bias2)  # activation-free last layer: column sums only
    H.assert_close(bias2, go.sum(0), "column sums", rtol=2e-5, atol=1e-5 * (B ** 0.5))
    bias3 = torch.empty(n, device=cuda)
    ops.bias_act_bwd(go.to(cuda), h.detach().to(cuda), bias3)
    H.assert_equal(bias3, bias, "deterministic reduction")
