"""K19: the three hidden layers of actor and critic in one persistent tcgen05 kernel, against torch (float64 and cuBLAS TF32) and
against the per-layer K12 path it replaces."""
import pytest
import torch

from tests import helpers as H

pytestmark = pytest.mark.gpu


def _net(g, k0, dev, k_pad=None):
    k_pad = k_pad or k0
    ps = []
    for (n, k) in ((512, k0), (256, 512), (128, 256)):
        w = torch.randn(n, k, generator=g) / k ** 0.5
        if k == k0 and k_pad != k0:
            w = torch.cat([w, torch.zeros(n, k_pad - k0)], dim=1)
        ps += [w.to(dev).contiguous(), (0.5 * torch.randn(n, generator=g)).to(dev)]
    return ps


def _ref64(x, ps):
    h, outs = x.double(), []
    for w, b in zip(ps[0::2], ps[1::2]):
        h = torch.nn.functional.elu(torch.nn.functional.linear(h, w.double(), b.double()))
        outs.append(h)
    return outs


@pytest.mark.parametrize("B,k0,n_nets,keep", [(4096, 348, 2, False), (1000, 348, 2, True), (99, 348, 1, True), (128, 64, 1, False), (24576, 348, 2, True),
                                               (300, 272, 2, True), (4097, 352, 1, True), (129, 4, 2, False)])
def test_mlp3_matches_torch(cuda, lt_lib, B, k0, n_nets, keep):
    from locotouch_b200 import ops

    g = torch.Generator().manual_seed(B + k0)
    nets = []
    for _ in range(n_nets):
        x = torch.randn(B, k0, generator=g).to(cuda)
        ps = _net(g, k0, cuda)
        hs = (torch.full((B, 512), float("nan"), device=cuda) if keep else None, torch.full((B, 256), float("nan"), device=cuda) if keep else None,
              torch.full((B, 128), float("nan"), device=cuda))
        nets.append((x, tuple(ps), hs))
    res = ops.mlp3_forward(nets)
    assert res is not None, "shape should be supported"
    torch.cuda.synchronize()
    for (x, ps, hs) in nets:
        refs = _ref64(x, ps)
        for got, ref, tol in zip(hs, refs, (4e-3, 8e-3, 1.2e-2)):  # TF32 operands, the error of a layer feeds the next one
            if got is None:
                continue
            assert torch.isfinite(got).all()
            err = (got.double() - ref).abs().max().item()
            assert err < tol, f"max abs error {err} at width {got.shape[1]}"
        # the per-layer kernels this replaces (same TF32 products, other accumulation order / exp)
        h = x
        for w, b in zip(ps[0::2], ps[1::2]):
            h = ops.linear_bias_act(h, w, b, elu=True)
        H.assert_close(hs[2], h, "fused MLP vs per-layer K12", rtol=4e-3, atol=4e-3)


def test_mlp3_rows_beyond_batch_are_not_written(cuda, lt_lib):
    from locotouch_b200 import ops

    B = 200  # second slab is partial: rows 200..255 of the slab must not be stored
    g = torch.Generator().manual_seed(5)
    x = torch.randn(B, 348, generator=g).to(cuda)
    ps = _net(g, 348, cuda)
    big = [torch.full((B + 64, n), 7.0, device=cuda) for n in (512, 256, 128)]
    hs = tuple(t[:B] for t in big)
    assert ops.mlp3_forward([(x, tuple(ps), hs)]) is not None
    torch.cuda.synchronize()
    for t in big:
        assert (t[B:] == 7.0).all()
        assert (t[:B] != 7.0).any()


def test_mlp3_unsupported_shapes_are_reported(cuda, lt_lib):
    from locotouch_b200 import ops

    g = torch.Generator().manual_seed(1)
    x = torch.randn(64, 270, generator=g).to(cuda)
    ps = _net(g, 270, cuda)
    assert ops.mlp3_forward([(x, tuple(ps), (None, None, torch.empty(64, 128, device=cuda)))]) is None  # k0 = 270: not a multiple of 4
    x = torch.randn(64, 356, generator=g).to(cuda)
    ps = _net(g, 356, cuda)
    assert ops.mlp3_forward([(x, tuple(ps), (None, None, torch.empty(64, 128, device=cuda)))]) is None  # wider than the resident slab
