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
