"""Student-batch helpers (K8): padded batch assembly and the masked behaviour-cloning loss."""
from __future__ import annotations

import torch

from .. import ops


def pad_trajectories(flat: torch.Tensor, offsets: torch.Tensor, lengths: torch.Tensor, max_length: int | None = None):
    """ReplayBuffer._prepare_padded_sequence (reference replay_buffer.py:90-112) for trajectories stored back to back in
    ``flat`` [total_steps, D]: returns ([L_max, B, D] zero padded, masks [L_max, B] bool)."""
    if max_length is None:
        max_length = int(lengths.max().item())
    return ops.pad_trajectories(flat, offsets, lengths, max_length)


class _MaskedMSE(torch.autograd.Function):
    @staticmethod
    def forward(ctx, student, teacher, masks):
        out, grad = ops.masked_mse(student.contiguous(), teacher.contiguous(), masks.contiguous())
        ctx.save_for_backward(grad)
        ctx.mark_non_differentiable(out)
        ctx.out = out
        return out[0].clone()

    @staticmethod
    def backward(ctx, g):
        (grad,) = ctx.saved_tensors
        return grad * g, None, None


def masked_mse_loss(student_actions, teacher_actions, masks):
    """``((s - t)**2).mean(-1)`` masked-averaged over valid (t, b) (reference student.py:131,142) with the gradient produced by
    the same kernel pair; returns the scalar loss (differentiable w.r.t. ``student_actions``)."""
    return _MaskedMSE.apply(student_actions, teacher_actions.detach(), masks)
