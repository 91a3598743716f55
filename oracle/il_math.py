"""ORACLE (test infrastructure, never imported by the product path).

CPU restatement of the six ``isaaclab.utils.math`` helpers the LocoTouch terms call.

Third-party arithmetic: these live in **IsaacLab 2.2.1** (pinned only by reference README.md:4; not vendored under
/root/reference and not installable here), so this file restates IsaacLab's published algorithm.
PARITY UNPINNED for this file: no reference-side golden vector exists for these helpers; they are pinned only by the
closed-form identities in tests/test_oracle_math.py.  Call sites in the reference: locotouch/mdp/rewards.py:7,
375-378,490,500,510-511,521,531-532,542,555-559,581,592 and locotouch/mdp/observations.py:8,55-58,79-80,83,156-158.

Quaternions are (w, x, y, z).
"""
from __future__ import annotations

import math

import torch


def quat_apply(quat: torch.Tensor, vec: torch.Tensor) -> torch.Tensor:
    """v' = v + w t + xyz x t with t = 2 (xyz x v)  -- rotate ``vec`` by ``quat``."""
    shape = vec.shape
    quat = quat.reshape(-1, 4)
    vec = vec.reshape(-1, 3)
    xyz = quat[:, 1:]
    t = xyz.cross(vec, dim=-1) * 2
    return (vec + quat[:, 0:1] * t + xyz.cross(t, dim=-1)).view(shape)


def quat_apply_inverse(quat: torch.Tensor, vec: torch.Tensor) -> torch.Tensor:
    """v' = v - w t + xyz x t with t = 2 (xyz x v)  -- rotate ``vec`` by the inverse of ``quat``."""
    shape = vec.shape
    quat = quat.reshape(-1, 4)
    vec = vec.reshape(-1, 3)
    xyz = quat[:, 1:]
    t = xyz.cross(vec, dim=-1) * 2
    return (vec - quat[:, 0:1] * t + xyz.cross(t, dim=-1)).view(shape)


def quat_mul(q1: torch.Tensor, q2: torch.Tensor) -> torch.Tensor:
    """Hamilton product in IsaacLab's 8-multiply factored form."""
    shape = q1.shape
    q1 = q1.reshape(-1, 4)
    q2 = q2.reshape(-1, 4)
    w1, x1, y1, z1 = q1[:, 0], q1[:, 1], q1[:, 2], q1[:, 3]
    w2, x2, y2, z2 = q2[:, 0], q2[:, 1], q2[:, 2], q2[:, 3]
    ww = (z1 + x1) * (x2 + y2)
    yy = (w1 - y1) * (w2 + z2)
    zz = (w1 + y1) * (w2 - z2)
    xx = ww + yy + zz
    qq = 0.5 * (xx + (z1 - x1) * (x2 - y2))
    w = qq - ww + (z1 - y1) * (y2 - z2)
    x = qq - xx + (x1 + w1) * (x2 + w2)
    y = qq - yy + (w1 - x1) * (y2 + z2)
    z = qq - zz + (z1 + y1) * (w2 - x2)
    return torch.stack([w, x, y, z], dim=-1).view(shape)


def quat_conjugate(q: torch.Tensor) -> torch.Tensor:
    shape = q.shape
    q = q.reshape(-1, 4)
    return torch.cat((q[:, 0:1], -q[:, 1:]), dim=-1).view(shape)


def quat_inv(q: torch.Tensor, eps: float = 1e-9) -> torch.Tensor:
    """conj(q) / max(|q|^2, eps)  (IsaacLab >= 2.1; older releases normalised the conjugate instead)."""
    return quat_conjugate(q) / q.pow(2).sum(dim=-1, keepdim=True).clamp(min=eps)


def quat_from_euler_xyz(roll: torch.Tensor, pitch: torch.Tensor, yaw: torch.Tensor) -> torch.Tensor:
    cy = torch.cos(yaw * 0.5)
    sy = torch.sin(yaw * 0.5)
    cr = torch.cos(roll * 0.5)
    sr = torch.sin(roll * 0.5)
    cp = torch.cos(pitch * 0.5)
    sp = torch.sin(pitch * 0.5)
    qw = cy * cr * cp + sy * sr * sp
    qx = cy * sr * cp - sy * cr * sp
    qy = cy * cr * sp + sy * sr * cp
    qz = sy * cr * cp - cy * sr * sp
    return torch.stack([qw, qx, qy, qz], dim=-1)


def _copysign(mag: float, other: torch.Tensor) -> torch.Tensor:
    return torch.abs(torch.full_like(other, mag)) * torch.where(other < 0, -1.0, 1.0).to(other.dtype)


def euler_xyz_from_quat(quat: torch.Tensor, wrap_to_2pi: bool = False):
    """roll/pitch/yaw in (-pi, pi] (IsaacLab 2.2); ``wrap_to_2pi`` reproduces older releases' [0, 2 pi)."""
    q_w, q_x, q_y, q_z = quat[..., 0], quat[..., 1], quat[..., 2], quat[..., 3]
    sin_roll = 2.0 * (q_w * q_x + q_y * q_z)
    cos_roll = 1 - 2 * (q_x * q_x + q_y * q_y)
    roll = torch.atan2(sin_roll, cos_roll)
    sin_pitch = 2.0 * (q_w * q_y - q_z * q_x)
    pitch = torch.where(torch.abs(sin_pitch) >= 1, _copysign(math.pi / 2.0, sin_pitch), torch.asin(sin_pitch))
    sin_yaw = 2.0 * (q_w * q_z + q_x * q_y)
    cos_yaw = 1 - 2 * (q_y * q_y + q_z * q_z)
    yaw = torch.atan2(sin_yaw, cos_yaw)
    if wrap_to_2pi:
        return roll % (2 * math.pi), pitch % (2 * math.pi), yaw % (2 * math.pi)
    return roll, pitch, yaw
