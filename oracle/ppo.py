"""ORACLE (test infrastructure, never imported by the product path).

CPU (torch fp32) restatement of the loco_rl learning math on the hot path:

* ``gae_returns``        reference loco_rl/loco_rl/storage/rollout_storage.py:152-174
* ``store_transition``   reference loco_rl/loco_rl/algorithms/ppo.py:143-170 + rollout_storage.py:80-107
* ``act_sample``         reference loco_rl/loco_rl/modules/actor_critic.py:105-123 (Normal sample / log_prob)
* ``ppo_loss``           reference loco_rl/loco_rl/algorithms/ppo.py:264-302 (KL, adaptive LR, surrogate, value, entropy)
* ``clip_and_adam``      reference ppo.py:350-353 = torch.nn.utils.clip_grad_norm_ + torch.optim.Adam.step
* ``mlp_forward``        reference actor_critic.py:33-56 (Linear/ELU stack)

Pinned against the live reference classes by tests/test_oracle_vs_reference.py and by tests/golden/ppo_*.npz.
"""
from __future__ import annotations

import math

import torch


def gae_returns(rewards, values, dones, last_values, gamma, lam, normalize_advantage=True):
    """rewards/values [T,N,1] f32, dones [T,N,1] u8, last_values [N,1] -> (returns, advantages) [T,N,1]."""
    T = rewards.shape[0]
    returns = torch.zeros_like(rewards)
    adv = 0
    for t in reversed(range(T)):
        nxt = last_values if t == T - 1 else values[t + 1]
        not_term = 1.0 - dones[t].float()
        delta = rewards[t] + not_term * gamma * nxt - values[t]
        adv = delta + not_term * gamma * lam * adv
        returns[t] = adv + values[t]
    advantages = returns - values
    if normalize_advantage:
        advantages = (advantages - advantages.mean()) / (advantages.std() + 1e-8)  # std is unbiased (n-1)
    return returns, advantages


def bootstrap_rewards(rewards, values, time_outs, gamma):
    """ppo.py:162-165: r += gamma * V * time_out."""
    return rewards + gamma * torch.squeeze(values * time_outs.unsqueeze(1), 1)


def act_sample(mean, std, eps):
    """a = mu + sigma*eps (torch.normal semantics), log-prob summed over actions (actor_critic.py:118-123)."""
    std_b = std.expand_as(mean)
    actions = mean + std_b * eps
    var = std_b**2
    logp = (-((actions - mean) ** 2) / (2 * var) - std_b.log() - math.log(math.sqrt(2 * math.pi))).sum(dim=-1)
    return actions, logp


def mlp_forward(x, weights, biases):
    """Linear -> ELU -> ... -> Linear (no activation after the last layer)."""
    h = x
    for i, (w, b) in enumerate(zip(weights, biases)):
        h = torch.nn.functional.linear(h, w, b)
        if i < len(weights) - 1:
            h = torch.nn.functional.elu(h)
    return h


def ppo_loss(mu, sigma, value, actions, old_logp, old_mu, old_sigma, adv, returns, old_values,
             clip_param=0.2, value_loss_coef=1.0, entropy_coef=0.01, use_clipped_value_loss=True):
    """Differentiable PPO loss exactly as ppo.py:252-302 evaluates it.

    mu [B,A] (requires grad), sigma [A] (requires grad, expanded), value [B,1] (requires grad); the rest are data.
    Returns dict(loss, surrogate, value_loss, entropy_mean, kl_mean).
    """
    sig = sigma.expand_as(mu)
    var = sig**2
    logp = (-((actions - mu) ** 2) / (2 * var) - sig.log() - math.log(math.sqrt(2 * math.pi))).sum(dim=-1)
    entropy = (0.5 + 0.5 * math.log(2 * math.pi) + torch.log(sig)).sum(dim=-1)
    with torch.no_grad():
        kl = torch.sum(
            torch.log(sig / old_sigma + 1.0e-5) + (torch.square(old_sigma) + torch.square(old_mu - mu)) / (2.0 * torch.square(sig)) - 0.5,
            axis=-1,
        )
        kl_mean = torch.mean(kl)
    ratio = torch.exp(logp - torch.squeeze(old_logp))
    a = torch.squeeze(adv)
    surrogate = -a * ratio
    surrogate_clipped = -a * torch.clamp(ratio, 1.0 - clip_param, 1.0 + clip_param)
    surrogate_loss = torch.max(surrogate, surrogate_clipped).mean()
    if use_clipped_value_loss:
        value_clipped = old_values + (value - old_values).clamp(-clip_param, clip_param)
        vl = (value - returns).pow(2)
        vlc = (value_clipped - returns).pow(2)
        value_loss = torch.max(vl, vlc).mean()
    else:
        value_loss = (returns - value).pow(2).mean()
    loss = surrogate_loss + value_loss_coef * value_loss - entropy_coef * entropy.mean()
    return dict(loss=loss, surrogate=surrogate_loss, value_loss=value_loss, entropy_mean=entropy.mean(), kl_mean=kl_mean)


def adaptive_lr(lr, kl_mean, desired_kl=0.01):
    """ppo.py:275-278"""
    if kl_mean > desired_kl * 2.0:
        return max(1e-5, lr / 1.5)
    if kl_mean < desired_kl / 2.0 and kl_mean > 0.0:
        return min(1e-2, lr * 1.5)
    return lr


def clip_and_adam(params, grads, exp_avg, exp_avg_sq, step, lr, max_grad_norm=1.0, beta1=0.9, beta2=0.999, eps=1e-8):
    """clip_grad_norm_(max_norm) then one torch.optim.Adam step (defaults: no weight decay, no amsgrad), on lists of
    tensors, in place.  Returns the total gradient norm before clipping."""
    norms = torch.stack([torch.linalg.vector_norm(g, 2.0) for g in grads])
    total = torch.linalg.vector_norm(norms, 2.0)
    coef = torch.clamp(max_grad_norm / (total + 1e-6), max=1.0)
    step = step + 1
    bc1 = 1 - beta1**step
    bc2 = 1 - beta2**step
    for p, g, m, v in zip(params, grads, exp_avg, exp_avg_sq):
        g.mul_(coef)
        m.lerp_(g, 1 - beta1)
        v.mul_(beta2).addcmul_(g, g, value=1 - beta2)
        denom = (v.sqrt() / math.sqrt(bc2)).add_(eps)
        p.addcdiv_(m, denom, value=-(lr / bc1))
    return total, step


def actor_critic_shapes(obs_dim, critic_obs_dim, num_actions, actor_hidden, critic_hidden):
    """Parameter shapes in ``ActorCritic.parameters()`` order (actor_critic.py:33-64): std, actor.*, critic.*."""
    shapes = [("std", (num_actions,))]
    dims = [obs_dim] + list(actor_hidden) + [num_actions]
    for i in range(len(dims) - 1):
        shapes += [(f"actor.{2 * i}.weight", (dims[i + 1], dims[i])), (f"actor.{2 * i}.bias", (dims[i + 1],))]
    dims = [critic_obs_dim] + list(critic_hidden) + [1]
    for i in range(len(dims) - 1):
        shapes += [(f"critic.{2 * i}.weight", (dims[i + 1], dims[i])), (f"critic.{2 * i}.bias", (dims[i + 1],))]
    return shapes


def unflatten(flat, shapes):
    out, off = {}, 0
    for name, shape in shapes:
        n = 1
        for s in shape:
            n *= s
        out[name] = flat[off : off + n].view(shape)
        off += n
    assert off == flat.numel()
    return out


def _split(params):
    aw = [v for k, v in params.items() if k.startswith("actor") and k.endswith("weight")]
    ab = [v for k, v in params.items() if k.startswith("actor") and k.endswith("bias")]
    cw = [v for k, v in params.items() if k.startswith("critic") and k.endswith("weight")]
    cb = [v for k, v in params.items() if k.startswith("critic") and k.endswith("bias")]
    return aw, ab, cw, cb


def ppo_update(flat_params, shapes, storage, perm, *, num_learning_epochs, num_mini_batches, clip_param, value_loss_coef,
               entropy_coef, learning_rate, max_grad_norm, desired_kl, use_clipped_value_loss=True):
    """PPO.update() (ppo.py:179-385) for the plain ActorCritic / adaptive schedule, on a flat parameter vector.

    ``storage``: dict of flattened [T*N, ...] tensors obs, critic_obs, actions, values, returns, logp, advantages, mu,
    sigma.  ``perm`` is the permutation mini_batch_generator draws once (rollout_storage.py:189).
    Returns (mean_value_loss, mean_surrogate_loss, mean_entropy, lr_sequence, flat_params_after).
    """
    flat = flat_params.clone().requires_grad_(True)
    m = torch.zeros_like(flat_params)
    v = torch.zeros_like(flat_params)
    step = 0
    lr = learning_rate
    bs = perm.numel() // num_mini_batches
    sums = [0.0, 0.0, 0.0]
    lrs = []
    for _ in range(num_learning_epochs):
        for i in range(num_mini_batches):
            idx = perm[i * bs : (i + 1) * bs]
            params = unflatten(flat, shapes)
            aw, ab, cw, cb = _split(params)
            mu = mlp_forward(storage["obs"][idx], aw, ab)
            value = mlp_forward(storage["critic_obs"][idx], cw, cb)
            res = ppo_loss(mu, params["std"], value, storage["actions"][idx], storage["logp"][idx], storage["mu"][idx],
                           storage["sigma"][idx], storage["advantages"][idx], storage["returns"][idx], storage["values"][idx],
                           clip_param=clip_param, value_loss_coef=value_loss_coef, entropy_coef=entropy_coef,
                           use_clipped_value_loss=use_clipped_value_loss)
            if desired_kl is not None:
                lr = adaptive_lr(lr, float(res["kl_mean"]), desired_kl)
            lrs.append(lr)
            (grad,) = torch.autograd.grad(res["loss"], flat)
            with torch.no_grad():
                # clip_grad_norm_ takes the norm of the per-tensor norms; Adam is element-wise, so one flat tensor per
                # parameter tensor reproduces the reference exactly
                g_list = [grad[o : o + n] for o, n in _offsets(shapes)]
                p_list = [flat[o : o + n] for o, n in _offsets(shapes)]
                m_list = [m[o : o + n] for o, n in _offsets(shapes)]
                v_list = [v[o : o + n] for o, n in _offsets(shapes)]
                _, step = clip_and_adam(p_list, [g.clone() for g in g_list], m_list, v_list, step, lr, max_grad_norm)
            sums[0] += float(res["value_loss"].detach())
            sums[1] += float(res["surrogate"].detach())
            sums[2] += float(res["entropy_mean"].detach())
    k = num_learning_epochs * num_mini_batches
    return sums[0] / k, sums[1] / k, sums[2] / k, lrs, flat.detach()


def _offsets(shapes):
    off, out = 0, []
    for _, shape in shapes:
        n = 1
        for s in shape:
            n *= s
        out.append((off, n))
        off += n
    return out
