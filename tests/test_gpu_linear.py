"""K12: fused linear layer (tcgen05 TF32 GEMM + bias + ELU epilogue) against torch on the layer shapes of the actor-critic."""
import pytest
import torch

from tests import helpers as H

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("M,K,N,elu", [(24576, 348, 512, True), (24576, 512, 256, True), (4096, 256, 128, True), (24576, 128, 12, False),
                                       (1000, 348, 512, True), (37, 64, 8, True), (129, 36, 132, False)])
def test_fused_linear_matches_torch(cuda, lt_lib, M, K, N, elu):
    from locotouch_b200 import ops

    g = torch.Generator().manual_seed(M + K + N)
    x = torch.randn(M, K, generator=g).to(cuda)
    w = (torch.randn(N, K, generator=g) / K ** 0.5).to(cuda)
    b = torch.randn(N, generator=g).to(cuda)
    out = ops.linear_bias_act(x, w, b, elu=elu)
    assert out is not None, "shape should be supported"
    ref64 = torch.nn.functional.linear(x.double(), w.double(), b.double())
    if elu:
        ref64 = torch.nn.functional.elu(ref64)
    # TF32 operands (10-bit mantissa): |error| <~ 2^-11 * sqrt(K) * |x||w| per output, fp32 accumulation
    err = (out.double() - ref64).abs().max().item()
    assert err < 4e-3, f"max abs error {err}"
    prev = torch.backends.cuda.matmul.allow_tf32
    torch.backends.cuda.matmul.allow_tf32 = True
    try:
        ref_tf32 = torch.nn.functional.linear(x, w, b)
        if elu:
            ref_tf32 = torch.nn.functional.elu(ref_tf32)
    finally:
        torch.backends.cuda.matmul.allow_tf32 = prev
    H.assert_close(out, ref_tf32, "fused linear vs cuBLAS TF32 + ELU", rtol=2e-3, atol=2e-3)


def test_unsupported_shapes_are_reported_not_faked(cuda, lt_lib):
    from locotouch_b200 import ops

    x, b = torch.randn(64, 270, device=cuda), torch.randn(16, device=cuda)
    assert ops.linear_bias_act(x, torch.randn(16, 270, device=cuda), b) is None  # K = 270 is not a multiple of 4
    assert ops.linear_bias_act(torch.randn(64, 128, device=cuda), torch.randn(1, 128, device=cuda), torch.randn(1, device=cuda)) is None  # N = 1


@pytest.mark.parametrize("M,Nout,Kin", [(24576, 512, 348), (24576, 256, 512), (4096, 128, 256), (1000, 64, 128), (37, 8, 64)])
def test_fused_dgrad_elu_backward_matches_torch(cuda, lt_lib, M, Nout, Kin):
    from locotouch_b200 import ops

    g = torch.Generator().manual_seed(M + Nout + Kin)
    go = torch.randn(M, Nout, generator=g).to(cuda)
    w = (torch.randn(Nout, Kin, generator=g) / Nout ** 0.5).to(cuda)
    h = torch.nn.functional.elu(torch.randn(M, Kin, generator=g)).to(cuda)
    out = ops.dgrad_act_bwd(go, w, h)
    assert out is not None
    ref = (go.double() @ w.double()) * torch.where(h > 0, torch.ones_like(h), h + 1.0).double()
    err = (out.double() - ref).abs().max().item()
    assert err < 6e-3, f"max abs error {err}"


def test_inference_forward_uses_fused_layers_and_matches_the_modules(cuda, lt_lib):
    """ActorCritic.act_inference / evaluate in TF32 mode (rollout path): K12 per hidden layer, same numbers as the torch modules
    at TF32 tolerance; with autograd recording or TF32 off the torch modules run (bit-identical to calling them)."""
    from locotouch_b200 import _C
    from locotouch_b200.loco_rl import ActorCritic

    torch.manual_seed(0)
    ac = ActorCritic(348, 348, 12, [512, 256, 128], [512, 256, 128], "elu", 1.0).to(cuda)
    obs = torch.randn(4096, 348, device=cuda)
    prev = torch.backends.cuda.matmul.allow_tf32
    try:
        torch.backends.cuda.matmul.allow_tf32 = True
        with torch.no_grad():
            n0 = _C.launch_count
            mu, v = ac.act_inference(obs), ac.evaluate(obs)
            assert _C.launch_count - n0 == 6, "three fused hidden layers per network"
            H.assert_close(mu, ac.actor(obs), "actor: fused inference vs modules (TF32)", rtol=5e-3, atol=5e-3)
            H.assert_close(v, ac.critic(obs), "critic: fused inference vs modules (TF32)", rtol=5e-3, atol=5e-3)
        n0 = _C.launch_count
        assert ac.act_inference(obs).requires_grad and _C.launch_count == n0  # autograd recording: torch modules
        torch.backends.cuda.matmul.allow_tf32 = False
        with torch.no_grad():
            H.assert_equal(ac.act_inference(obs), ac.actor(obs), "fp32 mode: the modules themselves")
            assert _C.launch_count == n0
    finally:
        torch.backends.cuda.matmul.allow_tf32 = prev


@pytest.mark.parametrize("B,n,k", [(24576, 512, 348), (24576, 256, 512), (24576, 128, 256), (24576, 12, 128), (24576, 1, 128),
                                   (1000, 132, 100), (37, 8, 64), (4096, 64, 36), (8192, 512, 512), (333, 20, 12), (2048, 640, 600), (5000, 256, 272), (777, 3, 50)])
def test_wgrad_splitk_matches_torch(cuda, lt_lib, B, n, k):
    """K15: dW = g^T x with in-kernel split-K (tcgen05 kind::tf32 on row-major = MN-major operands, TMEM accumulator, vector
    reductions into dW) against the fp64 product; TF32 operand rounding bounds the error by ~2^-11 sqrt(B) |g||x|."""
    from locotouch_b200 import ops

    gen = torch.Generator().manual_seed(B + n + k)
    g = (torch.randn(B, n, generator=gen) / B ** 0.5).to(cuda)
    x = torch.randn(B, k, generator=gen).to(cuda)
    out = torch.full((n, k), 7.0, device=cuda)  # zero_first must clear it
    db = torch.full((n,), -3.0, device=cuda)
    res = ops.wgrad(g, x, out, db)
    assert res is not None, "shape should be supported"
    ref = g.double().t() @ x.double()
    ref_b = g.double().sum(0)
    err = (out.double() - ref).abs().max().item()
    assert err < 4e-3, f"max abs error {err}"
    # bias gradient: column sums of g (TF32-rounded operands on the tensor-core path: |g| ~ B^-1/2, B terms, 2^-11 relative each)
    err_b = (db.double() - ref_b).abs().max().item()
    assert err_b < 2e-3, f"bias gradient: max abs error {err_b}"
    # accumulate mode: a second call without clearing doubles the result
    ops.wgrad(g, x, out, db, zero_first=False)
    err2 = (out.double() - 2 * ref).abs().max().item()
    assert err2 < 8e-3, f"accumulate: max abs error {err2}"
    assert (db.double() - 2 * ref_b).abs().max().item() < 4e-3
    # structure check that a transposed / shifted operand would fail: compare against cuBLAS TF32 as well
    prev = torch.backends.cuda.matmul.allow_tf32
    torch.backends.cuda.matmul.allow_tf32 = True
    try:
        ref_tf32 = g.t() @ x
    finally:
        torch.backends.cuda.matmul.allow_tf32 = prev
    ops.wgrad(g, x, out)  # without the bias output
    H.assert_close(out, ref_tf32, "wgrad vs cuBLAS TF32", rtol=2e-3, atol=2e-3)


def test_wgrad_unsupported_shapes_are_reported(cuda, lt_lib):
    from locotouch_b200 import ops

    assert ops.wgrad(torch.randn(64, 32, device=cuda), torch.randn(64, 270, device=cuda), torch.zeros(32, 270, device=cuda)) is None  # k % 4
    assert ops.wgrad(torch.randn(64, 34, device=cuda), torch.randn(64, 128, device=cuda), torch.zeros(34, 128, device=cuda)) is None  # n % 4


@pytest.mark.parametrize("B,hidden,OBS", [(24576, [512, 256, 128], 348), (1000, [128, 64], 348), (24576, [512, 256, 128], 270), (777, [128, 64], 270)])
def test_tf32_training_pass_gradients_match_autograd(cuda, lt_lib, B, hidden, OBS):
    """The production (TF32) composition bench.py times -- K12 forward, K12 dgrad + ELU backward, K15 weight + bias gradients
    accumulated into the flat buffer -- against autograd through the same torch modules in float64 (reference
    loco_rl/loco_rl/modules/actor_critic.py:33-56 under loss.backward(), algorithms/ppo.py:350).  TF32 operand rounding bounds
    the error: every gradient tensor must agree to 1 % of its own largest entry."""
    import copy

    from locotouch_b200 import _C
    from locotouch_b200.loco_rl.modules.actor_critic import ActorCritic

    torch.manual_seed(B)
    A = 12  # OBS = 270: the locomotion task's observation width (not a multiple of 4: zero-padded copies feed K12 / K15)
    ac = ActorCritic(OBS, OBS, A, actor_hidden_dims=hidden, critic_hidden_dims=hidden, activation="elu").to(cuda)
    ref = copy.deepcopy(ac).double()
    ac.flatten_parameters()
    obs, cobs = torch.randn(B, OBS, device=cuda), torch.randn(B, OBS, device=cuda)
    g_mu, g_v = torch.randn(B, A, device=cuda) / B, torch.randn(B, 1, device=cuda) / B
    prev = torch.backends.cuda.matmul.allow_tf32
    torch.backends.cuda.matmul.allow_tf32 = True
    try:
        n0 = _C.launch_count
        mu, v = ac.train_forward(obs, cobs)
        ac.train_backward(g_mu, g_v)
        torch.cuda.synchronize()
        own = _C.launch_count - n0
    finally:
        torch.backends.cuda.matmul.allow_tf32 = prev
    # library launches only: per network 2 fused dgrad layers and one K15 per Linear (heads included), plus the forward: ONE K19
    # launch for both networks at the [512, 256, 128] stack, otherwise one K12 per hidden layer and network
    fwd = 1 if list(hidden) == [512, 256, 128] else 2 * len(hidden)
    if OBS % 4 == 0:  # paired backward: one K15 launch per hidden layer for BOTH networks, two head wgrads, two K9 below the heads, dgrads
        assert own >= fwd + len(hidden) + 2 + 2 + 2 * (len(hidden) - 1), own
    else:
        assert own >= fwd + 2 * ((len(hidden) - 1) + len(hidden) + 1), own
    mu64, v64 = ref.actor(obs.double()), ref.critic(cobs.double())
    H.assert_close(mu, mu64.float(), "mu (TF32 forward)", rtol=2e-2, atol=2e-2)
    torch.autograd.backward([mu64, v64], [g_mu.double(), g_v.double()])
    got = dict(ac.named_parameters())
    for name, p in ref.named_parameters():
        if name in ("std", "log_std"):
            continue
        want = p.grad
        have = got[name].grad.double()
        scale = want.abs().max().item()
        err = (have - want).abs().max().item()
        # bias gradients are sums of mixed-sign terms that went through up to three TF32 GEMMs (cancellation): 3 % of the largest entry
        tol = 3e-2 if name.endswith("bias") else 1e-2
        assert err <= tol * scale + 1e-9, f"{name}: max abs error {err:.3e} against a largest entry of {scale:.3e}"


@pytest.mark.parametrize("N,Hd,A,critic", [(4096, 128, 12, True), (405, 128, 12, True), (1, 128, 4, False), (4097, 256, 16, True)])
def test_rollout_heads_kernel_matches_torch(cuda, lt_lib, N, Hd, A, critic):
    """K3b: actor / critic head layers + Normal sample + log-prob in one launch against the torch expressions of reference
    actor_critic.py:105-131 (explicit eps), and the Philox path against lt_act_sample's stream on the same means."""
    from locotouch_b200 import ops

    g = torch.Generator().manual_seed(N + A)
    rn = lambda *s: torch.randn(*s, generator=g).to(cuda)  # noqa: E731
    ha, hc = rn(N, Hd), rn(N, Hd)
    wa, ba, wc, bc = rn(A, Hd) / Hd ** 0.5, rn(A), rn(1, Hd) / Hd ** 0.5, rn(1)
    sigma = (0.5 + torch.rand(A, generator=g)).to(cuda)
    eps = rn(N, A)
    actions, logp, mu, values = ops.act_heads(ha, hc if critic else None, wa, ba, wc, bc, sigma, eps=eps)
    mu_ref = (ha.double() @ wa.double().t() + ba.double())
    H.assert_close(mu, mu_ref.float(), "mu", rtol=1e-5, atol=1e-5)
    a_ref = mu_ref + sigma.double() * eps.double()
    H.assert_close(actions, a_ref.float(), "actions", rtol=1e-5, atol=1e-5)
    lp_ref = torch.distributions.Normal(mu_ref, sigma.double().expand_as(mu_ref)).log_prob(a_ref).sum(-1)
    H.assert_close(logp, lp_ref.float(), "log prob", rtol=1e-5, atol=1e-4)
    if critic:
        H.assert_close(values.view(-1), (hc.double() @ wc.double().t() + bc.double()).view(-1).float(), "values", rtol=1e-5, atol=1e-5)
    # Philox: the same normal stream as lt_act_sample (key (seed, offset), counter (env, chunk))
    a2, lp2, mu2, _ = ops.act_heads(ha, None, wa, ba, None, None, sigma, seed=11, offset=5)
    a3, lp3 = ops.act_sample(mu2.contiguous(), sigma, seed=11, offset=5)
    H.assert_close(a2, a3, "Philox actions == lt_act_sample on the same means", rtol=0, atol=1e-6)
    H.assert_close(lp2, lp3, "log prob == lt_act_sample", rtol=1e-5, atol=1e-4)


@pytest.mark.parametrize("B,n,k", [(24576, 512, 348), (24576, 256, 512), (24576, 128, 256), (1000, 132, 100), (37, 64, 64)])
def test_wgrad_pair_matches_two_single_launches(cuda, lt_lib, B, n, k):
    """K15 with two problems of one shape (actor and critic layer) in ONE launch against the fp64 products; bias both set / both None."""
    from locotouch_b200 import ops

    gen = torch.Generator().manual_seed(B + n + k + 1)
    gs = [(torch.randn(B, n, generator=gen) / B ** 0.5).to(cuda) for _ in range(2)]
    xs = [torch.randn(B, k, generator=gen).to(cuda) for _ in range(2)]
    outs = [torch.zeros(n, k, device=cuda) for _ in range(2)]
    dbs = [torch.zeros(n, device=cuda) for _ in range(2)]
    assert ops.wgrad_pair(gs[0], xs[0], outs[0], dbs[0], gs[1], xs[1], outs[1], dbs[1]) is not None
    for g, x, out, db in zip(gs, xs, outs, dbs):
        assert (out.double() - g.double().t() @ x.double()).abs().max().item() < 4e-3
        assert (db.double() - g.double().sum(0)).abs().max().item() < 2e-3
    outs2 = [torch.zeros(n, k, device=cuda) for _ in range(2)]
    assert ops.wgrad_pair(gs[0], xs[0], outs2[0], None, gs[1], xs[1], outs2[1], None) is not None
    for a, b in zip(outs, outs2):
        H.assert_close(a, b, "pair launch with / without bias outputs", rtol=1e-3, atol=1e-3)
    assert ops.wgrad_pair(gs[0], xs[0], outs[0], dbs[0], gs[1], xs[1], outs[1], None) is None  # bias outputs: both or none


@pytest.mark.parametrize("B", [24576, 1000])
def test_tf32_production_backward_from_hidden_matches_autograd(cuda, lt_lib, B):
    """The backward exactly as PPO.minibatch_grads drives it: K19 forward without heads, the gradients below the heads handed over in
    ``hidden_grad_buffers()`` (what K16 writes), then ``train_backward(from_hidden=True)`` = paired K15 launches + K12 dgrads; against
    float64 autograd of the reference modules (actor_critic.py:33-56 under loss.backward())."""
    import copy

    from locotouch_b200.loco_rl.modules.actor_critic import ActorCritic

    torch.manual_seed(B + 3)
    A, OBS, hidden = 12, 348, [512, 256, 128]
    ac = ActorCritic(OBS, OBS, A, actor_hidden_dims=hidden, critic_hidden_dims=hidden, activation="elu").to(cuda)
    ref = copy.deepcopy(ac).double()
    ac.flatten_parameters()
    obs, cobs = torch.randn(B, OBS, device=cuda), torch.randn(B, OBS, device=cuda)
    g_mu, g_v = torch.randn(B, A, device=cuda) / B, torch.randn(B, 1, device=cuda) / B
    prev = torch.backends.cuda.matmul.allow_tf32
    torch.backends.cuda.matmul.allow_tf32 = True
    try:
        h_a, h_c = ac.train_forward(obs, cobs, heads=False)
        assert tuple(h_a.shape) == (B, 128) and tuple(h_c.shape) == (B, 128)
        gh_a, gh_c = ac.hidden_grad_buffers()
        with torch.no_grad():  # dL/d(pre-activation of the last hidden layer) = (g W_head) * elu'(h), elu' through the stored activation
            gh_a.copy_(((g_mu.double() @ ac.actor[-1].weight.double()) * torch.where(h_a > 0, 1.0, h_a.double() + 1.0)).float())
            gh_c.copy_(((g_v.double() @ ac.critic[-1].weight.double()) * torch.where(h_c > 0, 1.0, h_c.double() + 1.0)).float())
        ac.train_backward(g_mu, g_v, from_hidden=True)
        torch.cuda.synchronize()
    finally:
        torch.backends.cuda.matmul.allow_tf32 = prev
    mu64, v64 = ref.actor(obs.double()), ref.critic(cobs.double())
    torch.autograd.backward([mu64, v64], [g_mu.double(), g_v.double()])
    got = dict(ac.named_parameters())
    for name, p in ref.named_parameters():
        if name in ("std", "log_std"):
            continue
        want, have = p.grad, got[name].grad.double()
        scale = want.abs().max().item()
        err = (have - want).abs().max().item()
        tol = 3e-2 if name.endswith("bias") else 1e-2
        assert err <= tol * scale + 1e-9, f"{name}: max abs error {err:.3e} against a largest entry of {scale:.3e}"


@pytest.mark.parametrize("n,k,pair", [(512, 352, False), (512, 320, False), (128, 256, True), (512, 348, True)])
def test_wgrad_bias_fold_is_repeatable_with_warm_l2(cuda, lt_lib, n, k, pair):
    """Regression: K15's bias-folding warps released a pipeline stage while their shared-memory loads of it were still in flight (generic
    LD.E is not ordered with the mbarrier unit), so with fast TMA (128-byte-aligned row pitch, warm L2, the longer slices of the pair
    launch) a few rows of the NEXT batch slice leaked into the column sums: errors of 5-10 % that no single cold launch shows."""
    from locotouch_b200 import ops

    B = 24576
    gen = torch.Generator().manual_seed(n * k)
    g = (torch.randn(B, n, generator=gen) / B ** 0.5).to(cuda)
    x = torch.randn(B, k, generator=gen).to(cuda)
    ref = g.double().sum(0)
    for _ in range(12):
        out, db = torch.zeros(n, k, device=cuda), torch.zeros(n, device=cuda)
        if pair:
            out2, db2 = torch.zeros(n, k, device=cuda), torch.zeros(n, device=cuda)
            assert ops.wgrad_pair(g, x, out, db, g, x, out2, db2) is not None
            assert (db2.double() - ref).abs().max().item() < 2e-3
        else:
            assert ops.wgrad(g, x, out, db, zero_first=False) is not None
        assert (db.double() - ref).abs().max().item() < 2e-3
