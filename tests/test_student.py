"""Student drop-in (CNN2dHead -> GRU+MLP -> MLP) vs the reference Student (golden student_c4.npz): identical state_dict
keys, forward actions, masked BC loss and parameters after one AdamW step."""
import pytest
import torch

from tests import helpers as H
from tests.golden.make_golden import STUDENT_SMALL, shrink_student_cfg, student_batch


def _build(device):
    from locotouch_b200.distill import DistillationRandCylinderCNNRNNMonCfg, Student

    gold = H.load_golden("student_c4.npz")
    cfg = shrink_student_cfg(DistillationRandCylinderCNNRNNMonCfg())
    cfg.device = str(device)
    batch, tw, tb = student_batch()
    tw, tb = tw.to(device), tb.to(device)
    student = Student(cfg, 270, 442, 12, teacher_policy_inference=lambda x: torch.nn.functional.linear(x, tw, tb))
    assert list(student.state_dict().keys()) == [str(n) for n in gold["names"]], "state_dict keys must match reference checkpoints"
    flat = torch.as_tensor(gold["init"])
    off = 0
    sd = {}
    for k, v in student.state_dict().items():
        sd[k] = flat[off:off + v.numel()].view(v.shape).clone()
        off += v.numel()
    student.load_state_dict(sd)
    return student, {k: v.to(device) for k, v in batch.items()}, gold


def test_full_size_student_architecture():
    from locotouch_b200.distill import DistillationRandCylinderCNNRNNMonCfg, Student

    s = Student(DistillationRandCylinderCNNRNNMonCfg(device="cpu"), 270, 442, 12)
    assert sum(p.numel() for p in s.parameters()) == 1422420
    keys = list(s.state_dict())
    assert keys[0] == "pre_encoder.conv.conv.0.weight" and "student_encoder.memory.rnn.weight_ih_l0" in keys
    assert "student_encoder.mlp.model.0.weight" in keys and "student_backbone.model.6.bias" in keys
    assert s.pre_encoder.conv.conv_out_size(17, 13) == 192


def test_student_forward_matches_reference_on_cpu():
    student, batch, gold = _build(torch.device("cpu"))
    with torch.no_grad():
        actions = student(batch["proprioceptions"], batch["tactile_signals"])
    H.assert_close(actions, gold["actions"], "student actions", rtol=1e-5, atol=1e-6)


@pytest.mark.gpu
def test_student_training_step_matches_reference(cuda, lt_lib):
    torch.backends.cuda.matmul.allow_tf32 = False
    torch.backends.cudnn.allow_tf32 = False
    student, batch, gold = _build(cuda)
    student.train()
    with torch.no_grad():
        actions = student(batch["proprioceptions"], batch["tactile_signals"])
    H.assert_close(actions, gold["actions"], "student actions (cuDNN conv / GRU)", rtol=1e-4, atol=1e-5)
    loss = student.train_on_batch(batch)
    H.assert_close(loss, gold["loss"], "masked behaviour-cloning loss", rtol=1e-5, atol=1e-6)
    after = torch.cat([v.flatten() for v in student.state_dict().values()])
    H.assert_close(after, gold["after"], "parameters after one AdamW step", rtol=1e-4, atol=2e-5)
    # the gradient really moved every trainable tensor
    assert float((after.cpu() - torch.as_tensor(gold["init"])).abs().max()) > 1e-4


def _full_cnn(device):
    """The drop-in CNN2dHead at the LocoTouch geometry, loaded with the parameters of the reference instance (golden)."""
    from locotouch_b200.distill import DistillationRandCylinderCNNRNNMonCfg
    from locotouch_b200.loco_rl.models import generate_model

    gold = H.load_golden("student_cnn_c4.npz")
    cfg = DistillationRandCylinderCNNRNNMonCfg(device="cpu")
    net = generate_model(442, cfg.pre_encoder.embedding_dim, cfg.pre_encoder)
    assert list(net.state_dict().keys()) == [str(n) for n in gold["names"]]
    flat, off, sd = torch.as_tensor(gold["params"]), 0, {}
    for k, v in net.state_dict().items():
        sd[k] = flat[off:off + v.numel()].view(v.shape).clone()
        off += v.numel()
    net.load_state_dict(sd)
    return net.to(device), gold


def test_full_geometry_cnn_modules_match_reference_on_cpu():
    from tests.golden.make_golden import student_cnn_inputs

    net, gold = _full_cnn(torch.device("cpu"))
    bits, dense = student_cnn_inputs()
    assert abs(float(bits.double().sum() + dense.double().sum()) - float(gold["checksum"])) < 1e-6
    with torch.no_grad():
        H.assert_close(net(bits.reshape(-1, 2, 17, 13)), gold["out_bits"], "CNN2dHead (binary frames)", rtol=1e-5, atol=1e-6)
        H.assert_close(net(dense.reshape(-1, 2, 17, 13)), gold["out_dense"], "CNN2dHead (dense frames)", rtol=1e-5, atol=1e-6)


@pytest.mark.gpu
def test_fused_student_cnn_matches_reference(cuda, lt_lib):
    """K17 (conv1 + ReLU + pool + conv2 + ReLU + conv3 + ReLU + head in one kernel) against the reference CNN2dHead's own outputs
    (golden, fp32 CPU): fp32 image input, the ballot-packed bitmap input, and other batch sizes against the torch modules."""
    from locotouch_b200 import _C
    from tests.golden.make_golden import student_cnn_inputs

    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    net, gold = _full_cnn(cuda)
    bits, dense = student_cnn_inputs()
    n0 = _C.launch_count
    with torch.no_grad():
        out_bits = net(bits.to(cuda).reshape(-1, 2, 17, 13))
        out_dense = net(dense.to(cuda).reshape(-1, 2, 17, 13))
    assert _C.launch_count == n0 + 2, "the no-grad forward must be the fused kernel (one launch per call)"
    H.assert_close(out_bits, gold["out_bits"], "K17, binary frames (fp32 image input)", rtol=1e-5, atol=2e-6)
    H.assert_close(out_dense, gold["out_dense"], "K17, dense frames", rtol=1e-5, atol=2e-6)
    # packed input: bit t % 32 of word t // 32 = taxel t (what K2 writes)
    b = bits[:, :221].to(torch.int64)
    words = torch.zeros(bits.shape[0], 7, dtype=torch.int64)
    for t in range(221):
        words[:, t // 32] |= b[:, t] << (t % 32)
    packed = words.to(torch.int32).to(cuda) if int(words.max()) < 2 ** 31 else (words - (words >= 2 ** 31) * 2 ** 32).to(torch.int32).to(cuda)
    H.assert_equal(net.forward_packed(packed), out_bits, "packed-bitmap input == image input")
    # other sizes (ragged last block, more frames than warps) against the torch modules themselves
    for M in (1, 405, 4097):
        g = torch.Generator().manual_seed(M)
        x = (torch.rand(M, 442, generator=g) * (torch.rand(M, 442, generator=g) < 0.3)).to(cuda)
        with torch.no_grad():
            fused = net(x.reshape(-1, 2, 17, 13))
        with torch.enable_grad():
            ref = net(x.reshape(-1, 2, 17, 13)).detach()   # autograd on: torch modules (cuDNN / cuBLAS)
        H.assert_close(fused, ref, f"K17 vs torch modules, M={M}", rtol=2e-5, atol=5e-6)
