"""Functional wrappers over the C ABI: tensors in, tensors out, one ABI call each.

These are the thinnest possible host side -- argument checking, workspace ownership and pointer extraction.  The
classes that mirror the reference interfaces (``locotouch_b200.loco_rl``, ``locotouch_b200.mdp``,
``locotouch_b200.distill``) are built on top of them.  Every function requires CUDA tensors; there is no CPU path.
"""
from __future__ import annotations

import ctypes as C
import os

import torch

from . import _C
from ._C import check, count_launches, current_stream, lib, ptr

_workspaces: dict[tuple, torch.Tensor] = {}


def _workspace(key: str, nbytes: int, device) -> torch.Tensor:
    """Zero-initialised, cached device workspace (the kernels keep their counters self-cleaning)."""
    if torch.device(device).type != "cuda":
        raise _C.LocoTouchLibraryError(f"workspace requested on {device}: locotouch_b200 has no CPU path")
    k = (key, str(device), torch.cuda.current_stream(device).cuda_stream if torch.cuda.is_available() else 0)
    ws = _workspaces.get(k)
    if ws is None or ws.numel() < nbytes:
        ws = torch.zeros(max(nbytes, 256), dtype=torch.uint8, device=device)
        _workspaces[k] = ws
    return ws


# ----------------------------------------------------------------------------------------------------------- K4 GAE
def gae(rewards, values, dones, last_values, gamma: float, lam: float, normalize_advantage: bool = True,
        returns=None, advantages=None):
    """RolloutStorage.compute_returns (reference rollout_storage.py:152-174).  Tensors are [T,N] or [T,N,1]."""
    T, N = rewards.shape[0], rewards.shape[1]
    if returns is None:
        returns = torch.empty_like(rewards)
    if advantages is None:
        advantages = torch.empty_like(rewards)
    if dones.dtype != torch.uint8:
        raise _C.LocoTouchLibraryError("dones must be uint8 (RolloutStorage stores them with .byte())")
    nbytes = lib().lt_gae_workspace_bytes(T, N)
    ws = _workspace("gae", nbytes, rewards.device)
    check(lib().lt_gae(ptr(rewards, torch.float32, "rewards"), ptr(values, torch.float32, "values"), ptr(dones, torch.uint8, "dones"),
                       ptr(last_values, torch.float32, "last_values"), ptr(returns, torch.float32), ptr(advantages, torch.float32),
                       T, N, gamma, lam, int(normalize_advantage), ptr(ws), nbytes, current_stream()), "lt_gae")
    fused = normalize_advantage and T == 24 and (N + 31) // 32 <= 8 * torch.cuda.get_device_properties(rewards.device).multi_processor_count
    count_launches(2 if (normalize_advantage and not fused) else 1)
    return returns, advantages


def gae_scan(rewards, values, dones, last_values, gamma, lam, returns, advantages, stats):
    """First half of ``gae`` (no normalisation); ``stats`` (4 x float64) receives sum, sum of squares, count."""
    T, N = rewards.shape[0], rewards.shape[1]
    nbytes = lib().lt_gae_workspace_bytes(T, N)
    ws = _workspace("gae", nbytes, rewards.device)
    check(lib().lt_gae_scan(ptr(rewards, torch.float32), ptr(values, torch.float32), ptr(dones, torch.uint8), ptr(last_values, torch.float32),
                            ptr(returns, torch.float32), ptr(advantages, torch.float32), T, N, gamma, lam, ptr(stats, torch.float64),
                            ptr(ws), nbytes, current_stream()), "lt_gae_scan")
    count_launches(1)


def adv_normalize(advantages, stats):
    check(lib().lt_adv_normalize(ptr(advantages, torch.float32), advantages.numel(), ptr(stats, torch.float64), current_stream()), "lt_adv_normalize")
    count_launches(1)


def process_actions(actions, raw, prev_raw, prev_prev_raw=None, processed=None, prev_processed=None, prev_prev_processed=None, *,
                    clip: float = 0.0, raw_scale: float = 1.0, scale: float = 1.0, offset=None):
    """JointPositionActionPrevPrev.process_actions (reference mdp/actions.py:30-44) on [N,J] state tensors, in place."""
    check(lib().lt_process_actions(ptr(actions, torch.float32, "actions"), clip, raw_scale, scale, ptr(offset, torch.float32, "offset"),
                                   ptr(raw, torch.float32, "raw"), ptr(prev_raw, torch.float32), ptr(prev_prev_raw, torch.float32),
                                   ptr(processed, torch.float32), ptr(prev_processed, torch.float32), ptr(prev_prev_processed, torch.float32),
                                   actions.numel(), current_stream()), "lt_process_actions")
    count_launches(1)


def counter_add(counter, inc: int = 1):
    """``counter += inc`` on the current stream (device-resident step counter for CUDA-graph replays)."""
    check(lib().lt_counter_add(ptr(counter, torch.int64, "counter"), inc, current_stream()), "lt_counter_add")
    count_launches(1)


# ------------------------------------------------------------------------------------------------- K3 act / store / K5
def act_sample(mu, sigma, eps=None, actions=None, logp=None, mu_out=None, sigma_out=None, seed: int = 0, offset: int = 0, offset_base=None):
    """Normal(mu, sigma).sample() + log_prob().sum(-1) (reference actor_critic.py:105-123)."""
    N, A = mu.shape
    if actions is None:
        actions = torch.empty_like(mu)
    if logp is None:
        logp = torch.empty(N, device=mu.device, dtype=torch.float32)
    check(lib().lt_act_sample(ptr(mu, torch.float32, "mu"), ptr(sigma, torch.float32, "sigma"), ptr(eps, torch.float32, "eps"),
                              ptr(actions, torch.float32), ptr(logp, torch.float32), ptr(mu_out, torch.float32), ptr(sigma_out, torch.float32),
                              N, A, seed, offset, ptr(offset_base, torch.int64, "offset_base"), current_stream()), "lt_act_sample")
    count_launches(1)
    return actions, logp


def store_step(rewards, dones, time_outs, values, gamma, rewards_out, dones_out, obs=None, obs_out=None, critic_obs=None,
               critic_obs_out=None):
    """PPO.process_env_step bootstrap + RolloutStorage.add_transitions (reference ppo.py:143-170, rollout_storage.py:80-107)."""
    N = rewards.shape[0]
    d64 = ptr(dones, torch.int64, "dones") if dones.dtype == torch.int64 else None
    d8 = ptr(dones.view(torch.uint8) if dones.dtype == torch.bool else dones, torch.uint8, "dones") if dones.dtype != torch.int64 else None
    to = None
    if time_outs is not None:
        to = ptr(time_outs.view(torch.uint8) if time_outs.dtype == torch.bool else time_outs, torch.uint8, "time_outs")
    check(lib().lt_store_step(ptr(rewards, torch.float32, "rewards"), d64, d8, to, ptr(values, torch.float32, "values"), gamma,
                              ptr(rewards_out, torch.float32), ptr(dones_out, torch.uint8),
                              ptr(obs, torch.float32), ptr(obs_out, torch.float32), obs.shape[-1] if obs is not None else 0,
                              ptr(critic_obs, torch.float32), ptr(critic_obs_out, torch.float32),
                              critic_obs.shape[-1] if critic_obs is not None else 0, N, current_stream()), "lt_store_step")
    n = 1
    if obs is not None and obs_out is not None and obs.data_ptr() != obs_out.data_ptr():
        n += 1
    if critic_obs is not None and critic_obs_out is not None and critic_obs.data_ptr() != critic_obs_out.data_ptr():
        n += 1
    count_launches(n)


def gather_rows(sources, indices, outs=None):
    """``[src[indices] for src in sources]`` in one launch (reference rollout_storage.py:221-231)."""
    count = indices.numel()
    if outs is None:
        outs = [torch.empty((count,) + tuple(s.shape[1:]), device=s.device, dtype=s.dtype) for s in sources]
    args = _C.LtGatherArgs()
    args.num_tensors = len(sources)
    if len(sources) > _C.LT_GATHER_MAX:
        raise _C.LocoTouchLibraryError("too many tensors for one gather")
    for i, (s, o) in enumerate(zip(sources, outs)):
        args.src[i] = ptr(s, torch.float32, "gather source")
        args.dst[i] = ptr(o, torch.float32, "gather destination")
        args.row_len[i] = s[0].numel()
    check(lib().lt_gather_rows(C.byref(args), ptr(indices, torch.int64, "indices"), count, current_stream()), "lt_gather_rows")
    count_launches(1)
    return outs


# --------------------------------------------------------------------------------------------------------- K6 PPO loss
class PpoLossBuffers:
    """Pre-allocated outputs + workspace for ``ppo_loss`` (no allocation inside the update loop / graph capture)."""

    def __init__(self, B: int, A: int, device):
        self.B, self.A = B, A
        self.grad_mu = torch.empty(B, A, device=device)
        self.grad_value = torch.empty(B, device=device)
        self.grad_sigma = torch.empty(A, device=device)
        self.out = torch.zeros(8, device=device)
        self.nbytes = max(lib().lt_ppo_loss_workspace_bytes(B, A), lib().lt_ppo_heads_workspace_bytes(B, A))
        self.ws = torch.zeros(self.nbytes, dtype=torch.uint8, device=device)


def ppo_loss(mu, sigma, value, actions, old_logp, old_mu, old_sigma, advantages, returns, old_values, *, clip_param=0.2,
             value_loss_coef=1.0, entropy_coef=0.0, use_clipped_value_loss=True, desired_kl=None, lr=None, loss_accum=None,
             grad_scale=1.0, buffers: PpoLossBuffers | None = None):
    """Fused PPO loss forward + backward (reference ppo.py:252-302).  Returns the ``PpoLossBuffers`` holding
    ``grad_mu`` [B,A], ``grad_value`` [B], ``grad_sigma`` [A] and ``out`` = (loss, surrogate, value_loss, entropy, kl, lr)."""
    B, A = mu.shape
    if buffers is None:
        buffers = PpoLossBuffers(B, A, mu.device)
    a = _C.LtPpoLossArgs()
    a.B, a.A = B, A
    a.mu = ptr(mu, torch.float32, "mu")
    a.sigma = ptr(sigma, torch.float32, "sigma")
    a.value = ptr(value, torch.float32, "value")
    a.actions = ptr(actions, torch.float32, "actions")
    a.old_logp = ptr(old_logp, torch.float32, "old_logp")
    a.old_mu = ptr(old_mu, torch.float32, "old_mu")
    a.old_sigma = ptr(old_sigma, torch.float32, "old_sigma")
    a.advantages = ptr(advantages, torch.float32, "advantages")
    a.returns = ptr(returns, torch.float32, "returns")
    a.old_values = ptr(old_values, torch.float32, "old_values")
    a.clip_param, a.value_loss_coef, a.entropy_coef = clip_param, value_loss_coef, entropy_coef
    a.use_clipped_value_loss = int(use_clipped_value_loss)
    a.desired_kl = float(desired_kl) if (desired_kl is not None and lr is not None) else 0.0
    a.grad_scale = grad_scale
    a.grad_mu, a.grad_value, a.grad_sigma = ptr(buffers.grad_mu), ptr(buffers.grad_value), ptr(buffers.grad_sigma)
    a.out = ptr(buffers.out)
    a.lr_inout = ptr(lr, torch.float32, "lr")
    a.loss_accum = ptr(loss_accum, torch.float32, "loss_accum")
    a.workspace, a.workspace_bytes = ptr(buffers.ws), buffers.nbytes
    check(lib().lt_ppo_loss(C.byref(a), current_stream()), "lt_ppo_loss")
    count_launches(1)
    return buffers


def _fill_loss_args(a, B, A, sigma, actions, old_logp, old_mu, old_sigma, advantages, returns, old_values, clip_param, value_loss_coef, entropy_coef,
                    use_clipped_value_loss, desired_kl, lr, loss_accum, grad_scale, buffers):
    a.B, a.A = B, A
    a.sigma = ptr(sigma, torch.float32, "sigma")
    a.actions = ptr(actions, torch.float32, "actions")
    a.old_logp = ptr(old_logp, torch.float32, "old_logp")
    a.old_mu = ptr(old_mu, torch.float32, "old_mu")
    a.old_sigma = ptr(old_sigma, torch.float32, "old_sigma")
    a.advantages = ptr(advantages, torch.float32, "advantages")
    a.returns = ptr(returns, torch.float32, "returns")
    a.old_values = ptr(old_values, torch.float32, "old_values")
    a.clip_param, a.value_loss_coef, a.entropy_coef = clip_param, value_loss_coef, entropy_coef
    a.use_clipped_value_loss = int(use_clipped_value_loss)
    a.desired_kl = float(desired_kl) if (desired_kl is not None and lr is not None) else 0.0
    a.grad_scale = grad_scale
    a.grad_mu, a.grad_value, a.grad_sigma = ptr(buffers.grad_mu), ptr(buffers.grad_value), ptr(buffers.grad_sigma)
    a.out = ptr(buffers.out)
    a.lr_inout = ptr(lr, torch.float32, "lr")
    a.loss_accum = ptr(loss_accum, torch.float32, "loss_accum")
    a.workspace, a.workspace_bytes = ptr(buffers.ws), buffers.nbytes


def ppo_heads_supported(A: int, H_actor: int, H_critic: int) -> bool:
    """Shapes K16 takes: A % 4 == 0, A <= 16, both last hidden layers of one width H with H % 128 == 0 and H <= 256."""
    return A % 4 == 0 and A <= 16 and H_actor == H_critic and H_actor % 128 == 0 and H_actor <= 256


def ppo_heads_loss(h_actor, h_critic, w_actor, b_actor, w_critic, b_critic, sigma, actions, old_logp, old_mu, old_sigma, advantages, returns,
                   old_values, g_h_actor, g_h_critic, *, clip_param=0.2, value_loss_coef=1.0, entropy_coef=0.0, use_clipped_value_loss=True,
                   desired_kl=None, lr=None, loss_accum=None, grad_scale=1.0, buffers: PpoLossBuffers, mu_out=None, value_out=None):
    """K16: head layers of both MLPs + PPO loss (reference ppo.py:252-302) + head dgrad with the ELU backward of the last hidden
    layer, one pass.  Writes ``g_h_actor`` / ``g_h_critic`` [B,H], ``buffers.grad_mu`` / ``grad_value`` (inputs of the head weight
    gradients), ``buffers.grad_sigma`` and ``buffers.out``; ``mu_out`` [B,A] / ``value_out`` [B] optionally receive the head outputs."""
    B, H = h_actor.shape
    A = w_actor.shape[0]
    if tuple(h_critic.shape) != (B, H) or tuple(w_actor.shape) != (A, H) or w_critic.numel() != H or tuple(g_h_actor.shape) != (B, H) or tuple(g_h_critic.shape) != (B, H):
        raise _C.LocoTouchLibraryError("ppo_heads_loss: shape mismatch")
    h = _C.LtPpoHeadsArgs()
    _fill_loss_args(h.loss, B, A, sigma, actions, old_logp, old_mu, old_sigma, advantages, returns, old_values, clip_param, value_loss_coef, entropy_coef,
                    use_clipped_value_loss, desired_kl, lr, loss_accum, grad_scale, buffers)
    h.loss.mu = ptr(mu_out, torch.float32, "mu_out")
    h.loss.value = ptr(value_out, torch.float32, "value_out")
    h.H = H
    h.h_actor, h.h_critic = ptr(h_actor, torch.float32, "h_actor"), ptr(h_critic, torch.float32, "h_critic")
    h.w_actor, h.b_actor = ptr(w_actor, torch.float32, "w_actor"), ptr(b_actor, torch.float32, "b_actor")
    h.w_critic, h.b_critic = ptr(w_critic, torch.float32, "w_critic"), ptr(b_critic, torch.float32, "b_critic")
    h.g_h_actor, h.g_h_critic = ptr(g_h_actor, torch.float32, "g_h_actor"), ptr(g_h_critic, torch.float32, "g_h_critic")
    check(lib().lt_ppo_heads_loss(C.byref(h), current_stream()), "lt_ppo_heads_loss")
    count_launches(1)
    return buffers


def adaptive_lr(kl_sum, kl_scale: float, desired_kl: float, lr):
    check(lib().lt_adaptive_lr(ptr(kl_sum, torch.float32), kl_scale, desired_kl, ptr(lr, torch.float32), current_stream()), "lt_adaptive_lr")
    count_launches(1)


# ------------------------------------------------------------------------------------------------------- K7 clip + Adam
def _adam_launches(n: int, device) -> int:
    """1 when the parameter count fits the co-resident grid of the one-launch kernel (optim.cu: 4 blocks x 256 threads x 4 float4
    per SM, at most 1024 blocks), else the norm pass + update pass."""
    blocks = -(-max(n // 4, 1) // (256 * 4))
    return 1 if blocks <= min(1024, 4 * torch.cuda.get_device_properties(device).multi_processor_count) else 2


def clip_adam(params, grads, exp_avg, exp_avg_sq, lr, step, *, max_grad_norm=1.0, betas=(0.9, 0.999), eps=1e-8, weight_decay=0.0,
              grad_scale=1.0, grad_norm_out=None):
    """clip_grad_norm_ + Adam/AdamW step over flat fp32 buffers (reference ppo.py:350-353).  ``lr``/``step`` are device scalars."""
    n = params.numel()
    nbytes = lib().lt_clip_adam_workspace_bytes(n)
    ws = _workspace("adam", nbytes, params.device)
    check(lib().lt_clip_adam(ptr(params, torch.float32, "params"), ptr(grads, torch.float32, "grads"), ptr(exp_avg, torch.float32),
                             ptr(exp_avg_sq, torch.float32), n, ptr(lr, torch.float32, "lr"), ptr(step, torch.float32, "step"),
                             float(max_grad_norm if max_grad_norm is not None else 0.0), betas[0], betas[1], eps, weight_decay,
                             grad_scale, ptr(grad_norm_out, torch.float32), ptr(ws), nbytes, current_stream()), "lt_clip_adam")
    count_launches(_adam_launches(n, params.device))


def peer_sum_clip_adam(params, peer_ptrs, grad_sum, tail, exp_avg, exp_avg_sq, lr, step, *, max_grad_norm=1.0, betas=(0.9, 0.999), eps=1e-8,
                       weight_decay=0.0, grad_scale=1.0, desired_kl=None, kl_scale=1.0, grad_norm_out=None, gather: bool = False):
    """K14: grad_sum = sum over ranks of the flat gradient buffers at ``peer_ptrs`` (raw device-visible addresses, rank order), then
    (``desired_kl``: adaptive learning-rate decision from the summed KL statistic in the tail) clip_grad_norm_ + Adam on
    grad_sum * grad_scale.  The caller has put a cross-GPU barrier between the ranks' gradient writes and
    this call.  ``gather``: second half of the two-shot exchange (slice q of the sum is read from rank q, which reduced it with
    ``peer_reduce_scatter``; a second cross-GPU barrier separates the two)."""
    n = params.numel()
    nbytes = lib().lt_clip_adam_workspace_bytes(n)
    ws = _workspace("adam", nbytes, params.device)
    arr = (C.c_void_p * len(peer_ptrs))(*[int(x) for x in peer_ptrs])
    fn = lib().lt_peer_gather_clip_adam if gather else lib().lt_peer_sum_clip_adam
    check(fn(ptr(params, torch.float32, "params"), arr, len(peer_ptrs), ptr(grad_sum, torch.float32, "grad_sum"), int(tail),
             ptr(exp_avg, torch.float32), ptr(exp_avg_sq, torch.float32), n, ptr(lr, torch.float32, "lr"),
             ptr(step, torch.float32, "step"), float(max_grad_norm if max_grad_norm is not None else 0.0), betas[0], betas[1],
             eps, weight_decay, grad_scale, float(desired_kl) if desired_kl else 0.0, float(kl_scale),
             ptr(grad_norm_out, torch.float32), ptr(ws), nbytes, current_stream()),
          "lt_peer_gather_clip_adam" if gather else "lt_peer_sum_clip_adam")
    count_launches(_adam_launches(n, params.device))


def peer_reduce_scatter(peer_ptrs, rank: int, n: int):
    """K14 two-shot, first half: this rank sums ITS slice of all ranks' gradient buffers (peer loads, rank order) in place."""
    arr = (C.c_void_p * len(peer_ptrs))(*[int(x) for x in peer_ptrs])
    check(lib().lt_peer_reduce_scatter(arr, len(peer_ptrs), int(rank), int(n), current_stream()), "lt_peer_reduce_scatter")
    count_launches(1)


# ------------------------------------------------------------------------------------------------- K9 MLP backward helper
def bias_act_bwd(grad_out, act_out, bias_grad, grad_pre=None, alpha: float = 1.0):
    """grad_pre = grad_out * elu'(act_out) (in place when ``grad_pre`` is None) and bias_grad = grad_pre.sum(0), one pass."""
    B, n = grad_out.shape
    if grad_pre is None:
        grad_pre = grad_out
    nbytes = lib().lt_bias_act_bwd_workspace_bytes(B, n)
    ws = _workspace("biasbwd", nbytes, grad_out.device)
    check(lib().lt_bias_act_bwd(ptr(grad_out, torch.float32, "grad_out"), ptr(act_out, torch.float32, "act_out"), ptr(grad_pre, torch.float32),
                                ptr(bias_grad, torch.float32, "bias_grad"), B, n, alpha, ptr(ws), nbytes, current_stream()), "lt_bias_act_bwd")
    count_launches(1)
    return grad_pre


# ----------------------------------------------------------------------------------------------------------- K2 taxels
def taxel_synth(body_quat_w, net_forces_w, thresholds, *, quat_body_offset=0, u_drop=None, u_add=None, p_drop=0.005, p_add=0.005,
                seed=0, offset=0, offset_base=None, signal=None, packed=None, normal_forces=None, original_contact=None, delay_ring=None,
                delay_first=None, delay_steps=None, delayed_signal=None, want_signal=True, want_packed=True, delay_reset=None):
    """Binary taxel bitmap from contact forces (reference observations.py:154-199, 281-308)."""
    N, T = net_forces_w.shape[0], net_forces_w.shape[1]
    dev = net_forces_w.device
    if signal is None and want_signal:
        signal = torch.empty(N, 2 * T, device=dev)
    words = (T + 31) // 32
    if packed is None and want_packed:
        packed = torch.empty(N, words, device=dev, dtype=torch.int32)
    a = _C.LtTaxelArgs()
    a.N, a.T = N, T
    a.body_quat_w = ptr(body_quat_w, torch.float32, "body_quat_w")
    a.quat_num_bodies = body_quat_w.shape[1]
    a.quat_body_offset = quat_body_offset
    a.net_forces_w = ptr(net_forces_w, torch.float32, "net_forces_w")
    a.thresholds = ptr(thresholds, torch.float32, "thresholds")
    a.u_drop, a.u_add = ptr(u_drop, torch.float32), ptr(u_add, torch.float32)
    a.p_drop, a.p_add = p_drop, p_add
    a.seed, a.offset = seed, offset
    a.offset_base = ptr(offset_base, torch.int64, "offset_base")
    a.signal = ptr(signal, torch.float32)
    a.packed = ptr(packed, torch.int32)
    a.normal_forces = ptr(normal_forces, torch.float32)
    a.original_contact = ptr(original_contact.view(torch.uint8) if original_contact is not None and original_contact.dtype == torch.bool else original_contact, torch.uint8)
    if delay_ring is not None:
        a.delay_ring = ptr(delay_ring, torch.int32)
        a.delay_first = ptr(delay_first.view(torch.uint8) if delay_first.dtype == torch.bool else delay_first, torch.uint8)
        a.delay_steps = ptr(delay_steps, torch.int64)
        a.max_delay = delay_ring.shape[1]
        a.delayed_signal = ptr(delayed_signal, torch.float32)
        if delay_reset is not None:  # [N] bytes: envs reset since the previous frame (read only)
            a.delay_reset = ptr(delay_reset.view(torch.uint8) if delay_reset.dtype == torch.bool else delay_reset, torch.uint8)
    check(lib().lt_taxel_synth(C.byref(a), current_stream()), "lt_taxel_synth")
    count_launches(1)
    return signal, packed


def tactile_delay(ring, first, delay_steps, signal, out=None):
    """TactileRecorder.record_new_tactile_signals + get_tactile_signals (reference tactile_recorder.py:25-34)."""
    N, max_delay, D = ring.shape
    if out is None:
        out = torch.empty(N, D, device=ring.device)
    check(lib().lt_tactile_delay(ptr(ring, torch.float32, "ring"), ptr(first.view(torch.uint8) if first.dtype == torch.bool else first, torch.uint8),
                                 ptr(delay_steps, torch.int64, "delay_steps"), ptr(signal, torch.float32, "signal"), ptr(out, torch.float32),
                                 N, max_delay, D, current_stream()), "lt_tactile_delay")
    count_launches(1)
    return out


# ------------------------------------------------------------------------------------------------------ K8 student batch
def pad_trajectories(flat, offsets, lengths, L_max: int, out=None, masks=None):
    """ReplayBuffer._prepare_padded_sequence (reference replay_buffer.py:90-112): [sum_len, D] -> [L_max, B, D] + mask."""
    B, D = offsets.numel(), flat.shape[1]
    if out is None:
        out = torch.empty(L_max, B, D, device=flat.device)
    if masks is None:
        masks = torch.empty(L_max, B, device=flat.device, dtype=torch.bool)
    check(lib().lt_pad_trajectories(ptr(flat, torch.float32, "flat"), ptr(offsets, torch.int64), ptr(lengths, torch.int64), B, L_max, D,
                                    ptr(out, torch.float32), ptr(masks.view(torch.uint8), torch.uint8), current_stream()), "lt_pad_trajectories")
    count_launches(1)
    return out, masks


def masked_mse(student, teacher, masks, grad=None, out=None, want_grad=True):
    """Masked behaviour-cloning loss + gradient (reference student.py:131,142-151).  out = (loss, mae, count, 0)."""
    A = student.shape[-1]
    rows = student.numel() // A
    if grad is None and want_grad:
        grad = torch.empty_like(student)
    if out is None:
        out = torch.zeros(4, device=student.device)
    nbytes = lib().lt_masked_mse_workspace_bytes(rows)
    ws = _workspace("mse", nbytes, student.device)
    m = masks.view(torch.uint8) if masks.dtype == torch.bool else masks
    check(lib().lt_masked_mse(ptr(student, torch.float32, "student"), ptr(teacher, torch.float32, "teacher"), ptr(m, torch.uint8, "masks"),
                              rows, A, ptr(grad, torch.float32), ptr(out, torch.float32), ptr(ws), nbytes, current_stream()), "lt_masked_mse")
    count_launches(2 if grad is not None else 1)
    return out, grad


# ------------------------------------------------------------------------------------------- K10 trajectory split / pad
class TrajectoryIndex:
    """(env, first step, length) of every trajectory of a rollout cut at its dones (reference utils.py:55-61), on the device.
    ``base`` is also read back once (N + 1 host ints): a caller slicing trajectories per env range needs the counts."""

    def __init__(self, dones: torch.Tensor):
        d = dones.view(dones.shape[0], dones.shape[1])
        if d.dtype != torch.uint8:
            d = d.to(torch.uint8)
        d = d.contiguous()
        T, N = d.shape
        counts = 1 + (d[:-1] != 0).sum(dim=0, dtype=torch.int64)  # the last step always ends a trajectory
        ends = torch.cumsum(counts, dim=0)
        self.base_dev = (ends - counts).contiguous()
        self.base = [0] + ends.tolist()  # the one device -> host read (the reference does trajectory_lengths.tolist())
        self.T, self.N, self.M = T, N, self.base[-1]
        dev = d.device
        self.env = torch.empty(self.M, dtype=torch.int32, device=dev)
        self.start = torch.empty(self.M, dtype=torch.int32, device=dev)
        self.length = torch.empty(self.M, dtype=torch.int32, device=dev)
        check(lib().lt_trajectory_index(ptr(d, torch.uint8, "dones"), ptr(self.base_dev, torch.int64), ptr(self.env, torch.int32),
                                        ptr(self.start, torch.int32), ptr(self.length, torch.int32), T, N, current_stream()), "lt_trajectory_index")
        count_launches(1)

    def split_and_pad(self, x: torch.Tensor, want_masks: bool = True):
        """x [T, N, D] -> (padded [T, M, D], masks [T, M] bool)."""
        T, N = self.T, self.N
        xc = x.reshape(T, N, -1).contiguous()
        D = xc.shape[-1]
        out = torch.empty(T, self.M, D, device=x.device, dtype=torch.float32)
        masks = torch.empty(T, self.M, device=x.device, dtype=torch.uint8) if want_masks else None
        check(lib().lt_split_pad_trajectories(ptr(xc, torch.float32, "x"), ptr(self.env), ptr(self.start), ptr(self.length), ptr(out),
                                              ptr(masks), T, N, D, self.M, current_stream()), "lt_split_pad_trajectories")
        count_launches(1)
        return out, (masks.view(torch.bool) if masks is not None else None)

    def unpad(self, padded: torch.Tensor):
        """padded [T, M, D] -> [T, N, D]."""
        T, N = self.T, self.N
        pc = padded.contiguous()
        D = pc.shape[-1]
        out = torch.empty(T, N, D, device=padded.device, dtype=torch.float32)
        check(lib().lt_unpad_trajectories(ptr(pc, torch.float32, "padded"), ptr(self.env), ptr(self.start), ptr(self.length), ptr(out), T, N, D,
                                          self.M, current_stream()), "lt_unpad_trajectories")
        count_launches(1)
        return out

    @classmethod
    def from_masks(cls, masks: torch.Tensor):
        """Rebuilds the index from trajectory masks [T, M] alone (what ``unpad_trajectories(trajectories, masks)`` receives):
        lengths are column sums; trajectories tile the env-major flattened rollout, so start offsets are their exclusive scan."""
        self = cls.__new__(cls)
        T, M = masks.shape
        length = masks.sum(dim=0, dtype=torch.int64)
        flat_start = torch.cumsum(length, dim=0) - length
        self.T, self.M = T, M
        self.N = int(length.sum().item()) // T
        self.env = (flat_start // T).to(torch.int32).contiguous()
        self.start = (flat_start % T).to(torch.int32).contiguous()
        self.length = length.to(torch.int32).contiguous()
        self.base = None
        return self


# ------------------------------------------------------------------------------------------- K11 DAgger replay buffer
def dagger_step(dones, reward, reward_sums, start_idx, step_now: int, limit: int, state, traj_env, traj_start, traj_len, ep_reward, ep_length,
                always_restart: bool = False):
    """Per-step done handling of ReplayBuffer.collect_data (reference replay_buffer.py:57-73) on the device."""
    d = dones.view(torch.uint8) if dones.dtype == torch.bool else dones
    N = reward_sums.numel()
    check(lib().lt_dagger_step(ptr(d, torch.uint8, "dones"), ptr(reward, torch.float32, "reward"), ptr(reward_sums, torch.float32), ptr(start_idx, torch.int32),
                               N, int(step_now), int(limit), int(always_restart), ptr(state, torch.int64), ptr(traj_env, torch.int32),
                               ptr(traj_start, torch.int32), ptr(traj_len, torch.int32), ptr(ep_reward, torch.float32), ptr(ep_length, torch.int32),
                               current_stream()), "lt_dagger_step")
    count_launches(1)


def pack_trajectories(x, traj_env, traj_start, traj_offset, total_rows: int, flat):
    """Copies the recorded trajectories out of the step-major collection buffer x [S, N, D] into flat [total_rows, D]."""
    S, N, D = x.shape
    M = traj_env.numel()
    check(lib().lt_pack_trajectories(ptr(x, torch.float32, "x"), ptr(traj_env, torch.int32), ptr(traj_start, torch.int32), ptr(traj_offset, torch.int64),
                                     M, int(total_rows), N, D, ptr(flat, torch.float32, "flat"), current_stream()), "lt_pack_trajectories")
    count_launches(1)
    return flat


def taxel_forces(body_quat_w, net_forces_w, thresholds, out, channels, *, quat_body_offset=0, u=None, p_drop=0.0, p_add=0.0,
                 add_force_noise=False, force_n_prop_min=0.0, force_n_prop_max=0.0, maximal_force=1.0, total_levels=5, add_level_noise=False,
                 level_n_min=0.0, level_n_max=0.0, seed=0, offset=0, offset_base=None, normal_forces=None):
    """Force-valued tactile encodings (reference observations.py:166-237) into ``out`` [N, C, T]: ``channels`` lists what each of
    the C channels receives ("contact", "normalized", "minmax", "discretized").  ``u``: optional dict of explicit [N, T] uniforms
    (drop, drop_force, add, add_force, noise, small, level)."""
    N, T = net_forces_w.shape[0], net_forces_w.shape[1]
    a = _C.LtTaxelForceArgs()
    a.N, a.T = N, T
    a.body_quat_w = ptr(body_quat_w, torch.float32, "body_quat_w")
    a.quat_num_bodies, a.quat_body_offset = body_quat_w.shape[1], quat_body_offset
    a.net_forces_w = ptr(net_forces_w, torch.float32, "net_forces_w")
    a.thresholds = ptr(thresholds, torch.float32, "thresholds")
    if u is not None:
        for k, name in enumerate(("drop", "drop_force", "add", "add_force", "noise", "small", "level")):
            a.u[k] = ptr(u[name], torch.float32, "u_" + name) if u.get(name) is not None else None
    a.seed, a.offset = seed, offset
    a.offset_base = ptr(offset_base, torch.int64, "offset_base")
    a.p_drop, a.p_add = p_drop, p_add
    a.add_force_noise = int(add_force_noise)
    a.force_noise_min, a.force_noise_range = force_n_prop_min, force_n_prop_max - force_n_prop_min  # double arithmetic, rounded once
    a.maximal_force = maximal_force
    a.level_bin = 1.0 / total_levels
    a.add_level_noise = int(add_level_noise)
    a.level_noise_min, a.level_noise_range = level_n_min, level_n_max - level_n_min
    C_ = len(channels)
    if out.shape != (N, C_, T) or not out.is_contiguous():
        raise _C.LocoTouchLibraryError(f"out must be a contiguous [N, {C_}, T] tensor")
    a.out_stride = C_ * T
    base = ptr(out, torch.float32, "out")
    for c, name in enumerate(channels):
        if name not in ("contact", "normalized", "minmax", "discretized"):
            raise _C.LocoTouchLibraryError(f"unknown channel {name}")
        setattr(a, name, base + 4 * c * T)
    if normal_forces is not None:
        if C_ * T != T and normal_forces.stride(0) != C_ * T:
            raise _C.LocoTouchLibraryError("normal_forces must have the row stride of `out` (pass a [N, T] slice of a [N, C, T] tensor)")
        a.normal_forces = ptr(normal_forces, torch.float32, "normal_forces") if normal_forces.is_contiguous() else normal_forces.data_ptr()
    check(lib().lt_taxel_forces(C.byref(a), current_stream()), "lt_taxel_forces")
    count_launches(1)
    return out


# ------------------------------------------------------------------------------------------------- K12 fused linear layer
def linear_bias_act(x, weight, bias, out=None, elu: bool = True):
    """out = elu(x @ weight.T + bias) (or without the activation) in ONE tcgen05 TF32 GEMM with a fused epilogue.  Returns None
    when the shape / alignment is not supported (K, N multiples of 4, 16-byte aligned pointers) so that the caller can keep
    its cuBLAS path."""
    M, K = x.shape
    N = weight.shape[0]
    if (K & 3) or (N & 3) or not (x.is_contiguous() and weight.is_contiguous() and bias.is_contiguous()):
        return None
    if out is None:
        out = torch.empty(M, N, device=x.device, dtype=torch.float32)
    if (x.data_ptr() | weight.data_ptr() | bias.data_ptr() | out.data_ptr()) & 15:
        return None
    nbytes = lib().lt_linear_bias_act_workspace_bytes(M, N, K)
    ws = _workspace("linear", max(nbytes, 256), x.device)
    rc = lib().lt_linear_bias_act(ptr(x, torch.float32, "x"), ptr(weight, torch.float32, "weight"), ptr(bias, torch.float32, "bias"),
                                  ptr(out, torch.float32, "out"), M, N, K, int(elu), ptr(ws), ws.numel(), current_stream())
    if rc == _C.LT_ERR_UNSUPPORTED:
        return None
    check(rc, "lt_linear_bias_act")
    count_launches(1)
    return out


def dgrad_act_bwd(grad_out, weight, act_in, out=None):
    """out = (grad_out @ weight) * elu'(act_in): the dgrad GEMM with the ELU backward of the layer below in its epilogue (K12).
    Returns None when the shape / alignment is not supported."""
    M, Nout = grad_out.shape
    Kin = weight.shape[1]
    if (Kin & 3) or (Nout & 3) or not (grad_out.is_contiguous() and weight.is_contiguous() and act_in.is_contiguous()):
        return None
    if out is None:
        out = torch.empty(M, Kin, device=grad_out.device, dtype=torch.float32)
    if (grad_out.data_ptr() | weight.data_ptr() | act_in.data_ptr() | out.data_ptr()) & 15:
        return None
    nbytes = lib().lt_linear_bias_act_workspace_bytes(M, Kin, Nout)
    ws = _workspace("linear", max(nbytes, 256), grad_out.device)
    rc = lib().lt_dgrad_act_bwd(ptr(grad_out, torch.float32, "grad_out"), ptr(weight, torch.float32, "weight"), ptr(act_in, torch.float32, "act_in"),
                                ptr(out, torch.float32, "out"), M, Nout, Kin, ptr(ws), ws.numel(), current_stream())
    if rc == _C.LT_ERR_UNSUPPORTED:
        return None
    check(rc, "lt_dgrad_act_bwd")
    count_launches(1)
    return out


# ------------------------------------------------------------------------------------------- K19 fused three-layer MLP forward
MLP3_HIDDEN = (512, 256, 128)


def mlp3_supported(linears, k_in: int) -> bool:
    """The stacks K19 takes: three hidden layers [512, 256, 128] over an input width that is a multiple of 4 and <= 352."""
    return (len(linears) >= 3 and tuple(lin.out_features for lin in linears[:3]) == MLP3_HIDDEN and linears[0].in_features <= k_in
            and (k_in & 3) == 0 and k_in <= 352 and os.environ.get("LT_MLP3", "1") != "0")


def mlp3_forward(nets):
    """K19: the three hidden Linear + ELU layers of one or two MLPs in ONE persistent tcgen05 kernel.  ``nets`` = list of
    ``(x, (w1, b1, w2, b2, w3, b3), (h1, h2, h3))`` with x [B, k0], w1 [512, k0] (k0 % 4 == 0: zero-padded copies for other
    widths), h3 [B, 128] and h1 [B, 512] / h2 [B, 256] or None (rollout: only h3 is needed).  Returns the h3 tensors, or None
    when the shapes / alignment are not taken (callers keep the per-layer K12 path)."""
    if not 1 <= len(nets) <= 2:
        return None
    arr = (_C.LtMlp3Net * len(nets))()
    B, k0 = nets[0][0].shape
    for a, (x, params, hs) in zip(arr, nets):
        w1, b1, w2, b2, w3, b3 = params
        h1, h2, h3 = hs
        if (tuple(x.shape) != (B, k0) or tuple(w1.shape) != (512, k0) or tuple(w2.shape) != (256, 512) or tuple(w3.shape) != (128, 256)
                or (k0 & 3) or k0 > 352 or tuple(h3.shape) != (B, 128)):
            return None
        ts = [x, w1, b1, w2, b2, w3, b3, h3] + [h for h in (h1, h2) if h is not None]
        if any((not t.is_contiguous()) or (t.data_ptr() & 15) or t.dtype != torch.float32 for t in ts) or any(h is not None and (h.data_ptr() & 31) for h in hs):
            return None
        if (h1 is not None and tuple(h1.shape) != (B, 512)) or (h2 is not None and tuple(h2.shape) != (B, 256)):
            return None
        a.x, a.k0 = ptr(x, torch.float32, "x"), k0
        a.w1, a.b1, a.w2, a.b2, a.w3, a.b3 = (ptr(t, torch.float32, "param") for t in params)
        a.h1 = ptr(h1, torch.float32, "h1") if h1 is not None else None
        a.h2 = ptr(h2, torch.float32, "h2") if h2 is not None else None
        a.h3 = ptr(h3, torch.float32, "h3")
    rc = lib().lt_mlp3_forward(arr, len(nets), B, current_stream())
    if rc == _C.LT_ERR_UNSUPPORTED:
        return None
    check(rc, "lt_mlp3_forward")
    count_launches(1)
    return [hs[2] for (_, _, hs) in nets]


# ------------------------------------------------------------------------------------------- K15 split-K weight gradient
def wgrad(grad_out, act_in, out, bias_out=None, zero_first: bool = True):
    """out[n, k] (+)= grad_out[B, n]^T @ act_in[B, k] and, with ``bias_out`` [n], bias_out (+)= grad_out.sum(0), in ONE tcgen05 TF32
    kernel with in-kernel split-K (fp32 vector reductions straight into ``out``, e.g. a view of the flat gradient buffer; the bias
    gradient is summed from the tiles that pass through shared memory).  ``zero_first=False``: the caller cleared the outputs (or
    wants accumulation).  Returns None when the shape / alignment is not supported so that the caller can keep its cuBLAS path."""
    B, n = grad_out.shape
    k = act_in.shape[1]
    if act_in.shape[0] != B or tuple(out.shape) != (n, k) or (bias_out is not None and bias_out.numel() != n):
        raise _C.LocoTouchLibraryError("wgrad: shape mismatch")
    if not (grad_out.is_contiguous() and act_in.is_contiguous() and out.is_contiguous()) or (bias_out is not None and not bias_out.is_contiguous()):
        return None
    if n > 16 and ((n & 3) or (k & 3) or ((grad_out.data_ptr() | act_in.data_ptr() | out.data_ptr()) & 15)):
        return None
    rc = lib().lt_wgrad_splitk(ptr(grad_out, torch.float32, "grad_out"), ptr(act_in, torch.float32, "act_in"), ptr(out, torch.float32, "out"),
                               ptr(bias_out, torch.float32, "bias_out") if bias_out is not None else None, B, n, k, int(zero_first), current_stream())
    if rc == _C.LT_ERR_UNSUPPORTED:
        return None
    check(rc, "lt_wgrad_splitk")
    count_launches(1)
    return out


def wgrad_pair(grad_out0, act_in0, out0, bias_out0, grad_out1, act_in1, out1, bias_out1):
    """K15 for two layers of the same shape (actor and critic) in ONE launch; outputs are accumulated into (cleared by the caller).
    Returns None when the shapes / alignment are not taken (the caller launches them one by one)."""
    B, n = grad_out0.shape
    k = act_in0.shape[1]
    if (tuple(grad_out1.shape) != (B, n) or tuple(act_in1.shape) != (B, k) or tuple(out0.shape) != (n, k) or tuple(out1.shape) != (n, k) or n <= 16 or (n & 3) or (k & 3)
            or os.environ.get("LT_WGRAD_PAIR", "1") == "0"):
        return None
    ts = [grad_out0, act_in0, out0, grad_out1, act_in1, out1]
    if any((not t.is_contiguous()) or (t.data_ptr() & 15) for t in ts) or (bias_out0 is None) != (bias_out1 is None):
        return None
    rc = lib().lt_wgrad_splitk_pair(ptr(grad_out0, torch.float32), ptr(act_in0, torch.float32), ptr(out0, torch.float32),
                                    ptr(bias_out0, torch.float32) if bias_out0 is not None else None, ptr(grad_out1, torch.float32), ptr(act_in1, torch.float32),
                                    ptr(out1, torch.float32), ptr(bias_out1, torch.float32) if bias_out1 is not None else None, B, n, k, current_stream())
    if rc == _C.LT_ERR_UNSUPPORTED:
        return None
    check(rc, "lt_wgrad_splitk_pair")
    count_launches(1)
    return out0, out1


# ------------------------------------------------------------------------------------------- K17 student tactile pre-encoder
def student_cnn_supported(image_shape, channels, kernel_sizes, pool_strides, paddings, embedding_dim, nonlinearity="relu", normlayer=None) -> bool:
    """The geometry K17 takes (the LocoTouch student pre-encoder, reference loco_rl/models/model_cfg.py:17-25)."""
    return (tuple(image_shape) == (2, 17, 13) and tuple(channels) == (24, 24, 24) and tuple(kernel_sizes) == (4, 3, 2)
            and tuple(pool_strides) == (2, 1, 1) and all(int(q) == 0 for q in paddings) and 0 < embedding_dim <= 64
            and nonlinearity in ("relu", "crelu") and normlayer is None)


def student_cnn_forward(weights, *, image=None, packed=None, out=None):
    """CNN2dHead.forward (reference loco_rl/models/cnn_2d.py:75-131) of the LocoTouch student cfg in one kernel.  ``weights`` =
    (w1, b1, w2, b2, w3, b3, wh, bh) as the modules hold them; ``image`` [M, 442] fp32 or ``packed`` [M, words] int32 bitmaps."""
    if (image is None) == (packed is None):
        raise _C.LocoTouchLibraryError("student_cnn_forward: give exactly one of image / packed")
    src = image if image is not None else packed
    M = src.shape[0]
    w1, b1, w2, b2, w3, b3, wh, bh = weights
    E = wh.shape[0]
    if out is None:
        out = torch.empty(M, E, device=src.device)
    a = _C.LtStudentCnnArgs()
    a.M, a.in_channels, a.height, a.width = M, 2, 17, 13
    a.channels[:] = [w1.shape[0], w2.shape[0], w3.shape[0]]
    a.kernel_sizes[:] = [w1.shape[-1], w2.shape[-1], w3.shape[-1]]
    a.pool[:] = [2, 1, 1]
    a.embedding_dim = E
    a.image = ptr(image, torch.float32, "image")
    a.packed = ptr(packed, torch.int32, "packed")
    a.packed_words = packed.shape[1] if packed is not None else 0
    for name, t in zip(("w1", "b1", "w2", "b2", "w3", "b3", "wh", "bh"), weights):
        setattr(a, name, ptr(t.detach(), torch.float32, name))
    a.out = ptr(out, torch.float32, "out")
    check(lib().lt_student_cnn_forward(C.byref(a), current_stream()), "lt_student_cnn_forward")
    count_launches(1)
    return out


# ------------------------------------------------------------------------------------------------ K18 ContactSensor bookkeeping
def contact_sensor_update(forces, *, net_forces_w=None, history=None, current_air_time=None, last_air_time=None, current_contact_time=None,
                          last_contact_time=None, dt: float = 0.0, dt_per_env=None, force_threshold: float = 1.0, reset_mask=None):
    """[IL] ContactSensor._update_buffers_impl (+ reset for the envs in ``reset_mask``) in one launch: history ring shift + insert and
    the air / contact-time state machine of every body.  All state tensors are updated in place."""
    N, Bd = forces.shape[0], forces.shape[1]
    H = history.shape[1] if history is not None else 0
    if history is not None and tuple(history.shape) != (N, H, Bd, 3):
        raise _C.LocoTouchLibraryError("contact_sensor_update: history must be [N, H, bodies, 3]")
    for t in (current_air_time, last_air_time, current_contact_time, last_contact_time):
        if t is not None and tuple(t.shape) != (N, Bd):
            raise _C.LocoTouchLibraryError("contact_sensor_update: timers must be [N, bodies]")
    check(lib().lt_contact_sensor_update(ptr(forces, torch.float32, "forces"), ptr(net_forces_w, torch.float32, "net_forces_w"),
                                         ptr(history, torch.float32, "history"), H, N, Bd, ptr(current_air_time, torch.float32),
                                         ptr(last_air_time, torch.float32), ptr(current_contact_time, torch.float32),
                                         ptr(last_contact_time, torch.float32), ptr(dt_per_env, torch.float32, "dt_per_env"), float(dt),
                                         float(force_threshold), ptr(reset_mask, torch.uint8, "reset_mask"), current_stream()),
          "lt_contact_sensor_update")
    count_launches(1)


# ------------------------------------------------------------------------------------------------ K3b rollout heads + sampling
def act_heads(h_actor, h_critic, w_actor, b_actor, w_critic, b_critic, sigma, *, eps=None, actions=None, logp=None, mu=None, sigma_rows=None,
              values=None, seed: int = 0, offset: int = 0, offset_base=None):
    """Head layers of actor (and critic) on the last hidden activations + Normal sample + log-prob in one launch (K3b).  Outputs may
    be RolloutStorage slot rows.  Returns (actions, logp, mu, values)."""
    N, H = h_actor.shape
    A = w_actor.shape[0]
    dev = h_actor.device
    actions = actions if actions is not None else torch.empty(N, A, device=dev)
    logp = logp if logp is not None else torch.empty(N, device=dev)
    mu = mu if mu is not None else torch.empty(N, A, device=dev)
    if h_critic is not None and values is None:
        values = torch.empty(N, device=dev)
    check(lib().lt_act_heads(ptr(h_actor, torch.float32, "h_actor"), ptr(h_critic, torch.float32, "h_critic"), ptr(w_actor, torch.float32, "w_actor"),
                             ptr(b_actor, torch.float32), ptr(w_critic, torch.float32) if h_critic is not None else None,
                             ptr(b_critic, torch.float32) if h_critic is not None else None, ptr(sigma, torch.float32, "sigma"),
                             ptr(eps, torch.float32, "eps"), ptr(actions, torch.float32, "actions"), ptr(logp, torch.float32, "logp"),
                             ptr(mu, torch.float32, "mu"), ptr(sigma_rows, torch.float32, "sigma_rows"),
                             ptr(values.view(-1), torch.float32, "values") if values is not None else None, N, A, H, int(seed), int(offset),
                             ptr(offset_base, torch.int64, "offset_base"), current_stream()), "lt_act_heads")
    count_launches(1)
    return actions, logp, mu, values
