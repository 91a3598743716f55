"""PPO with the reference's interface (reference loco_rl/loco_rl/algorithms/ppo.py:19-385).

Same constructor keywords, attributes (``actor_critic``, ``optimizer``, ``storage``, ``transition``, ``learning_rate``,
``rnd``, ``symmetry``) and methods (``init_storage``, ``test_mode`` / ``train_mode``, ``act``, ``process_env_step``,
``compute_returns``, ``update`` returning the 5-tuple of ppo.py:385), so ``OnPolicyRunner`` drives it unchanged.

What changed below the interface:
* ``act`` writes actions / log-prob / mu / sigma / values directly into the RolloutStorage slot (K3);
* ``process_env_step`` fuses the time-out bootstrap into the store kernel (K3);
* ``compute_returns`` is the GAE kernel pair (K4);
* ``update``: one fused gather of the permuted rollout (K5); per mini-batch the hidden layers of both MLPs as tcgen05 GEMMs with
  bias + ELU in their epilogue (K12; fp32 parity mode: cuBLAS), ONE kernel for the head layers + the loss (log-prob, entropy, KL,
  adaptive learning rate on the device, clipped surrogate + value loss and their analytic gradients) + the head dgrad (K16), an
  explicit backward (K12 dgrad with the ELU backward fused, K15 weight + bias gradients into the flat gradient buffer), ONE fused
  clip + Adam over the flat parameter buffer (K7).  No ``.item()`` inside the loop: the learning rate lives on the device and the
  three logged means are accumulated on the device and read back once per update.  Recurrent policies take ``update_recurrent``
  (K10 trajectories, cuDNN GRU under autograd, K6, K7).  With ``torch.distributed`` initialised, gradients are exchanged by peer loads
  inside the optimizer kernel (K14; NCCL all-reduce of one flat buffer as fallback) with the KL statistic in the buffer's tail
  (SURVEY.md 8e).
RND and symmetry augmentation (unused by every LocoTouch runner cfg) are not part of the hot path and raise.
"""
from __future__ import annotations

import torch
import torch.distributed as dist

from ... import dist as D
from ... import ops
from ..modules import ActorCritic
from ..storage import RolloutStorage


_C_MAX_PEERS = 16  # LT_MAX_PEERS of the C ABI


class _NoBarrier:
    """Barrier of same-process replicas driven in lock step on one stream: stream order already is the barrier."""

    def barrier(self, channel: int = 0):
        pass


class _FusedAdam:
    """``optimizer``-shaped facade over the fused clip + Adam kernel: keeps ``param_groups[0]['lr']`` and
    ``state_dict`` / ``load_state_dict`` in torch.optim.Adam's format so runner checkpoints stay interchangeable."""

    def __init__(self, actor_critic: ActorCritic, lr: float):
        self.ac = actor_critic
        self.flat, self.grads = actor_critic.flatten_parameters()
        dev = self.flat.device
        self.exp_avg = torch.zeros_like(self.flat)
        self.exp_avg_sq = torch.zeros_like(self.flat)
        self.step_t = torch.zeros(1, device=dev)
        self.lr_t = torch.full((1,), lr, device=dev)
        self.grad_norm = torch.zeros(1, device=dev)
        self.param_groups = [dict(lr=lr, betas=(0.9, 0.999), eps=1e-8, weight_decay=0, amsgrad=False, params=list(range(len(list(actor_critic.parameters())))))]
        self._host_lr = lr
        self._peer_buf = None

    def _check_group(self):
        g = self.param_groups[0]
        if g.get("weight_decay", 0):  # torch.optim.Adam's weight decay is the L2 form (grad += wd * p); the kernel's is AdamW's
            raise NotImplementedError("PPO's optimizer is torch.optim.Adam: weight_decay != 0 (L2 form) is not implemented by the fused kernel")
        return g

    def use_gradient_buffer(self, buf: torch.Tensor, copy: bool = True):
        self.grads = self.ac.rebind_gradients(buf, copy=copy)
        self._peer_buf = buf

    def refresh(self):
        """Re-attaches the optimizer to the actor-critic's flat buffers when they were re-created (``module.to(...)`` to another
        device / dtype after construction re-homes the parameters): the Adam moments follow the parameters, and a gradient buffer
        installed with ``use_gradient_buffer`` (K14 symmetric memory) is bound again.  No-op in the common case."""
        flat, grads = self.ac.flatten_parameters()
        if flat is self.flat:
            return False
        if flat.numel() != self.flat.numel():
            raise RuntimeError("the actor-critic's parameter count changed after the optimizer was built")
        dev = flat.device
        self.exp_avg, self.exp_avg_sq = self.exp_avg.to(dev), self.exp_avg_sq.to(dev)
        self.step_t, self.lr_t, self.grad_norm = self.step_t.to(dev), self.lr_t.to(dev), self.grad_norm.to(dev)
        self.flat, self.grads = flat, grads
        buf = getattr(self, "_peer_buf", None)
        if buf is not None and buf.device == dev:
            self.grads = self.ac.rebind_gradients(buf)
        return True

    def step_peer_sum(self, peer_ptrs, grad_sum, tail: int, max_grad_norm=None, grad_scale: float = 1.0, desired_kl=None, kl_scale=1.0, gather=False):
        """K14: sum of the ranks' gradient buffers (peer loads, rank order) + learning-rate decision + clip + Adam."""
        g = self._check_group()
        ops.peer_sum_clip_adam(self.flat, peer_ptrs, grad_sum, tail, self.exp_avg, self.exp_avg_sq, self.lr_t, self.step_t,
                               max_grad_norm=max_grad_norm, betas=g["betas"], eps=g["eps"], weight_decay=g["weight_decay"],
                               grad_scale=grad_scale, desired_kl=desired_kl, kl_scale=kl_scale, grad_norm_out=self.grad_norm, gather=gather)

    def sync_lr_to_device(self):
        lr = self.param_groups[0]["lr"]
        if lr != self._host_lr:  # someone (runner, user) wrote param_groups[...]["lr"]
            self.lr_t.fill_(lr)
            self._host_lr = lr

    def zero_grad(self, set_to_none: bool = False):
        self.grads.zero_()

    def step(self, max_grad_norm=None, grad_scale: float = 1.0):
        g = self._check_group()
        ops.clip_adam(self.flat, self.grads, self.exp_avg, self.exp_avg_sq, self.lr_t, self.step_t, max_grad_norm=max_grad_norm,
                      betas=g["betas"], eps=g["eps"], weight_decay=g["weight_decay"], grad_scale=grad_scale, grad_norm_out=self.grad_norm)

    def state_dict(self):
        state = {}
        step = self.step_t.detach().cpu().clone().squeeze(0)
        for i, (name, p) in enumerate(self.ac.named_parameters()):
            off, n = self.ac._slices[name]
            state[i] = dict(step=step.clone(), exp_avg=self.exp_avg[off:off + n].view(p.shape).clone(), exp_avg_sq=self.exp_avg_sq[off:off + n].view(p.shape).clone())
        pg = dict(self.param_groups[0])
        pg["lr"] = float(self.lr_t.item())
        return dict(state=state, param_groups=[pg])

    def load_state_dict(self, sd):
        for i, (name, p) in enumerate(self.ac.named_parameters()):
            if i in sd["state"]:
                off, n = self.ac._slices[name]
                self.exp_avg[off:off + n].copy_(sd["state"][i]["exp_avg"].flatten())
                self.exp_avg_sq[off:off + n].copy_(sd["state"][i]["exp_avg_sq"].flatten())
                self.step_t.fill_(float(sd["state"][i]["step"]))
        lr = sd["param_groups"][0]["lr"]
        self.param_groups[0]["lr"] = lr
        self.lr_t.fill_(lr)
        self._host_lr = lr


class PPO:
    """Proximal Policy Optimization algorithm (https://arxiv.org/abs/1707.06347)."""

    actor_critic: ActorCritic

    def __init__(self, actor_critic, num_learning_epochs=1, num_mini_batches=1, clip_param=0.2, gamma=0.998, lam=0.95,
                 value_loss_coef=1.0, entropy_coef=0.0, learning_rate=1e-3, max_grad_norm=1.0, use_clipped_value_loss=True,
                 schedule="fixed", desired_kl=0.01, device="cpu", normalize_advantage_per_mini_batch=False,
                 rnd_cfg: dict | None = None, symmetry_cfg: dict | None = None):
        self.device = device
        if torch.device(device).type != "cuda":
            from ... import _C
            raise _C.LocoTouchLibraryError(f"locotouch_b200 PPO runs on CUDA only (device={device!r}); there is no CPU fallback")
        self.desired_kl = desired_kl
        self.schedule = schedule
        self.learning_rate = learning_rate
        self.normalize_advantage_per_mini_batch = normalize_advantage_per_mini_batch
        if rnd_cfg is not None:
            raise NotImplementedError("RND is not used by any LocoTouch cfg and is outside the hot path (SURVEY.md 2, #24)")
        self.rnd = None
        self.rnd_optimizer = None
        if symmetry_cfg is not None:
            use_symmetry = symmetry_cfg["use_data_augmentation"] or symmetry_cfg["use_mirror_loss"]
            if symmetry_cfg["use_data_augmentation"] and not callable(symmetry_cfg["data_augmentation_func"]):
                raise ValueError("Data augmentation enabled but the function is not callable:" f" {symmetry_cfg['data_augmentation_func']}")
            if use_symmetry:
                raise NotImplementedError("symmetry augmentation is not used by any LocoTouch cfg and is outside the hot path")
        self.symmetry = None
        self.actor_critic = actor_critic
        self.actor_critic.to(self.device)
        A = getattr(actor_critic, "num_actions", None)
        if A is not None and (A % 4 != 0 or A > 64):  # lt_act_sample / lt_ppo_loss move [A] rows as float4 chunks, up to 16 of them
            raise ValueError(f"locotouch_b200 PPO needs num_actions to be a multiple of 4 and <= 64 (got {A})")
        self.optimizer = _FusedAdam(self.actor_critic, learning_rate)
        # K16 (heads + loss + head dgrad in one kernel).  The head arithmetic is fp32 FMA in either GEMM mode, so it is used in the
        # fp32 parity mode too; set False to keep the separate head GEMMs + K6.
        self.fused_heads = True
        self.trace = None  # list -> reduce_and_step appends (local KL, lr, grad norm) per mini-batch (eager diagnostics)
        self._world_override = None  # (rank, world) of same-process replicas attached with attach_local_peers()
        self._peer = None
        self.storage: RolloutStorage = None  # type: ignore
        self.transition = RolloutStorage.Transition()
        self.clip_param = clip_param
        self.num_learning_epochs = num_learning_epochs
        self.num_mini_batches = num_mini_batches
        self.value_loss_coef = value_loss_coef
        self.entropy_coef = entropy_coef
        self.gamma = gamma
        self.lam = lam
        self.max_grad_norm = max_grad_norm
        self.use_clipped_value_loss = use_clipped_value_loss
        self._loss_bufs = None
        self._loss_accum = torch.zeros(4, device=self.device)
        self._adv_stats = torch.zeros(4, device=self.device, dtype=torch.float64)
        self.global_advantage_normalization = True  # multi-GPU: normalise over ALL ranks' envs (== the single-process reference)

    # ----------------------------------------------------------------------------------------------------------- setup
    def init_storage(self, num_envs, num_transitions_per_env, actor_obs_shape, critic_obs_shape, action_shape):
        self.storage = RolloutStorage(num_envs, num_transitions_per_env, actor_obs_shape, critic_obs_shape, action_shape, None, self.device)

    def test_mode(self):
        self.actor_critic.eval()

    def train_mode(self):
        self.actor_critic.train()

    # --------------------------------------------------------------------------------------------------------- rollout
    def act(self, obs, critic_obs):
        slot = self.storage.slot()
        t = self.transition
        if self.actor_critic.is_recurrent:  # the states BEFORE this step (reference ppo.py:130-131); cloned: the GRU output is re-used
            t.hidden_states = tuple(None if h is None else (tuple(x.clone() for x in h) if isinstance(h, tuple) else h.clone())
                                    for h in self.actor_critic.get_hidden_states())
        with torch.no_grad():
            t.actions = self.actor_critic.act_evaluate_fused(obs, critic_obs, slot) if self.fused_heads else None  # K12 x 6 + K3b
            if t.actions is None:
                side = self.actor_critic.side_streams(obs.device)[0]
                with side.forked():  # the critic runs next to the actor (two independent chains of small GEMMs)
                    values = self.actor_critic.evaluate(critic_obs)
                    slot["values"].copy_(values)  # the critic GEMM's output row is the only copy left on this path
                t.actions = self.actor_critic.act(obs, out=slot)
                side.join()
        t.values = slot["values"]
        t.actions_log_prob = slot["logp"]
        t.action_mean = slot["mu"]
        t.action_sigma = slot["sigma"]
        t.observations = obs
        t.critic_observations = critic_obs
        return t.actions

    def process_env_step(self, rewards, dones, infos):
        t = self.transition
        t.rewards = rewards
        t.dones = dones
        time_outs = infos["time_outs"] if "time_outs" in infos else None
        self.storage.add_transitions(t, time_outs=time_outs, gamma=self.gamma)
        t.clear()
        self.actor_critic.reset(dones)

    def _world_info(self):
        """(rank, world) of the env-sharded job this learner is part of: the process group's, or the one declared by
        ``attach_local_peers`` for replicas that live in ONE process."""
        return self._world_override if self._world_override is not None else D.world_info()

    def compute_returns(self, last_critic_obs):
        if self.compute_returns_scan(last_critic_obs):
            self.reduce_adv_stats()
            self.compute_returns_normalize()

    def reduce_adv_stats(self):
        """(sum, sum of squares, count) of the advantages over all ranks.  Under K14 the three doubles sit in symmetric memory: a
        cross-GPU barrier, then every rank adds the W blocks in rank order by peer loads -- plain launches, capturable in the
        rollout graph; otherwise one NCCL all-reduce."""
        pr = getattr(self, "_peer", None)
        if pr is not None and pr.get("adv_stats") is not None:
            pr["handle"].barrier(channel=2)
            total = pr["adv_stats"][0].clone()
            for t in pr["adv_stats"][1:]:
                total += t
            self._adv_total = total
        else:
            D.reduce_adv_stats_(self._adv_stats)
            self._adv_total = self._adv_stats

    def compute_returns_scan(self, last_critic_obs) -> bool:
        """GAE scan.  Returns True when the advantage statistics still have to be summed over the shards (env-sharded job with
        global normalisation): the caller reduces ``self._adv_stats`` = (sum, sum of squares, count) and then calls
        ``compute_returns_normalize``; otherwise everything, normalisation included, is done (K4, one launch)."""
        with torch.no_grad():
            last_values = self.actor_critic.evaluate(last_critic_obs).detach()
        st = self.storage
        normalize = not self.normalize_advantage_per_mini_batch
        _, world = self._world_info()
        if world == 1 or not normalize or not self.global_advantage_normalization:
            st.compute_returns(last_values, self.gamma, self.lam, normalize_advantage=normalize)
            return False
        # env-sharded ranks: sum / sum of squares / count are all-reduced between the scan and the normalisation
        ops.gae_scan(st.rewards, st.values, st.dones, last_values.contiguous().view(-1), self.gamma, self.lam, st.returns, st.advantages, self._adv_stats)
        return True

    def compute_returns_normalize(self):
        ops.adv_normalize(self.storage.advantages, getattr(self, "_adv_total", self._adv_stats))

    # ---------------------------------------------------------------------------------------------------------- update
    def update(self, indices=None):
        """reference ppo.py:179-385.  ``indices``: optional explicit permutation (the reference draws it with randperm)."""
        self.optimizer.sync_lr_to_device()
        if self.actor_critic.is_recurrent:
            self.update_recurrent()
        else:
            self.update_body(indices)
        return self.update_epilogue()

    def update_body(self, indices=None):
        """The device work of ``update()``: no host synchronisation, capturable in a CUDA graph when no process group is
        active (``indices`` must then be a persistent tensor that is refilled before each replay)."""
        self.update_begin(indices)
        for _epoch in range(self.num_learning_epochs):
            for i in range(self.num_mini_batches):
                self.minibatch_grads(i)
                self.reduce_and_step()
        self.peer_update_fence()
        self.storage.clear()

    # The three stages below are public so that a caller can capture them separately (CUDA graphs must not contain the
    # NCCL collectives of ``reduce_and_step`` when several processes train together).
    def update_begin(self, indices=None):
        """Draws / takes the permutation (rollout_storage.py:189) and gathers the permuted rollout once (K5)."""
        ac, st = self.actor_critic, self.storage
        if ac.is_recurrent:
            raise ValueError("recurrent policies update through update_recurrent()")
        if self.normalize_advantage_per_mini_batch:
            raise NotImplementedError("per-mini-batch advantage normalisation is not used by the LocoTouch cfgs")
        if ac.noise_std_type != "scalar":
            raise NotImplementedError("noise_std_type='log' is not used by the LocoTouch cfgs")
        self.optimizer.refresh()  # no-op unless the module was moved after construction
        self._loss_accum.zero_()
        batch_size = st.num_envs * st.num_transitions_per_env
        self._mb_size = batch_size // self.num_mini_batches
        if indices is None:
            indices = torch.randperm(self.num_mini_batches * self._mb_size, requires_grad=False, device=self.device)
        bufs, has_priv = st.gather_permuted(indices)
        it = iter(bufs)
        obs = next(it)
        cobs = next(it) if has_priv else obs
        if (obs.shape[1] & 3) and torch.backends.cuda.matmul.allow_tf32 and ac.supports_explicit_backward:
            # observation widths that are not a multiple of 4 (locomotion: 270) reach the tcgen05 kernels through zero-padded rows
            # (TMA needs a 16-byte pitch): pad the gathered rollout ONCE per update instead of every mini-batch slice
            obs_p = self._padded_obs(obs, 0)
            cobs = self._padded_obs(cobs, 1) if has_priv else obs_p
            obs = obs_p
        self._mb = (obs, cobs) + tuple(next(it) for _ in range(7))

    def select_gradient_buffer(self):
        """K14, double-buffered exchange: mini-batch k writes gradient buffer k % 2, so a rank may start its next backward while
        slower ranks still read its previous buffer; the barrier in front of mini-batch k + 1's peer reads orders buffer k % 2's
        next rewrite (mini-batch k + 2) after everybody's reads of it -- ONE cross-GPU barrier per mini-batch.  Call before
        anything writes gradients for the next ``reduce_and_step`` (``minibatch_grads`` does)."""
        if self.peer_gradients:
            pr = self._peer
            self.optimizer.use_gradient_buffer(pr["bufs"][pr["k"] % 2], copy=False)

    def _padded_obs(self, x, slot: int):
        kp = (x.shape[1] + 3) // 4 * 4
        bufs = self.__dict__.setdefault("_obs_pad", {})
        key = (slot, x.shape[0], kp)
        if key not in bufs:
            bufs[key] = torch.zeros(x.shape[0], kp, device=x.device)
        bufs[key][:, :x.shape[1]].copy_(x)
        return bufs[key]

    def minibatch_grads(self, i: int):
        """Forward of both MLPs on mini-batch slice ``i``, fused loss (K6), backward into the flat gradient buffer."""
        ac, opt = self.actor_critic, self.optimizer
        self.select_gradient_buffer()
        sl = slice(i * self._mb_size, (i + 1) * self._mb_size)
        obs, cobs, actions, values, returns, logp, adv, mu_old, sigma_old = (t[sl] for t in self._mb)
        B, A = actions.shape
        if self._loss_bufs is None or self._loss_bufs.B != B:
            self._loss_bufs = ops.PpoLossBuffers(B, A, self.device)
        bufs = self._loss_bufs
        adaptive = self.desired_kl is not None and self.schedule == "adaptive"
        _, world = self._world_info()
        local_lr = adaptive and world == 1  # one process: the learning-rate decision is taken inside the loss kernel
        explicit = ac.supports_explicit_backward
        off, n = ac._slices["std"]
        loss_kw = dict(clip_param=self.clip_param, value_loss_coef=self.value_loss_coef, entropy_coef=self.entropy_coef,
                       use_clipped_value_loss=self.use_clipped_value_loss, desired_kl=self.desired_kl if local_lr else None,
                       lr=opt.lr_t if local_lr else None, loss_accum=self._loss_accum, buffers=bufs)
        if explicit and self.fused_heads and ac.supports_fused_heads:
            # K16: head layers + loss + head dgrad in one kernel between the fused hidden layers (K12) and their backward
            h_a, h_c = ac.train_forward(obs, cobs, heads=False)
            g_ha, g_hc = ac.hidden_grad_buffers()
            head_a, head_c = ac.actor[-1], ac.critic[-1]
            ops.ppo_heads_loss(h_a, h_c, head_a.weight, head_a.bias, head_c.weight, head_c.bias, ac.std.detach(), actions, logp.view(-1), mu_old,
                               sigma_old, adv.view(-1), returns.view(-1), values.view(-1), g_ha, g_hc, **loss_kw)
            ac.train_backward(bufs.grad_mu, bufs.grad_value.view(-1, 1), from_hidden=True)
            opt.grads[off:off + n].copy_(bufs.grad_sigma)
        else:
            if explicit:
                mu, value = ac.train_forward(obs, cobs)
            else:  # non-ELU activations: autograd through the torch modules
                mu = ac.actor(obs)
                value = ac.critic(cobs)
                opt.zero_grad()
            ops.ppo_loss(mu.detach(), ac.std.detach(), value.detach().view(-1), actions, logp.view(-1), mu_old, sigma_old, adv.view(-1),
                         returns.view(-1), values.view(-1), **loss_kw)
            if explicit:
                ac.train_backward(bufs.grad_mu, bufs.grad_value.view(-1, 1))
                opt.grads[off:off + n].copy_(bufs.grad_sigma)  # every gradient slot is overwritten: no zero_grad pass
            else:
                torch.autograd.backward([mu, value], [bufs.grad_mu, bufs.grad_value.view_as(value)])
                opt.grads[off:off + n].add_(bufs.grad_sigma)
        if adaptive and world > 1:  # the local KL mean travels in the tail of the gradient all-reduce
            ac.flat_grads_ext[-4:-3].copy_(bufs.out[4:5])

    def update_recurrent(self):
        """The recurrent branch of reference ppo.py:195-196,251-302,350-353: mini-batches are env ranges whose observations travel as
        zero-padded trajectories with the RNN states of every trajectory's first step (K10 builds them once per update); the policy
        is re-evaluated in batch mode through cuDNN's GRU under autograd, the loss and its gradients w.r.t. (mu, V, sigma) are K6,
        autograd carries them back through the MLPs / GRU into the flat gradient buffer, clip + Adam is K7."""
        ac, opt, st = self.actor_critic, self.optimizer, self.storage
        if ac.noise_std_type != "scalar":
            raise NotImplementedError("noise_std_type='log' is not used by the LocoTouch cfgs")
        if self.normalize_advantage_per_mini_batch:
            raise NotImplementedError("per-mini-batch advantage normalisation is not used by the LocoTouch cfgs")
        opt.refresh()
        self._loss_accum.zero_()
        adaptive = self.desired_kl is not None and self.schedule == "adaptive"
        _, world = self._world_info()
        local_lr = adaptive and world == 1
        off, n = ac._slices["std"]
        for (obs, cobs, actions, values, adv, returns, logp, mu_old, sigma_old, hid, masks, _rnd) in st.recurrent_mini_batch_generator(
                self.num_mini_batches, self.num_learning_epochs):
            self.select_gradient_buffer()
            with torch.enable_grad():
                ac.act(obs, masks=masks, hidden_states=hid[0])
                mu = ac.action_mean                                                  # [T, n, A]
                value = ac.evaluate(cobs, masks=masks, hidden_states=hid[1])         # [T, n, 1]
            A = mu.shape[-1]
            B = mu.numel() // A
            if self._loss_bufs is None or self._loss_bufs.B != B:
                self._loss_bufs = ops.PpoLossBuffers(B, A, self.device)
            bufs = self._loss_bufs
            opt.zero_grad()
            flat = lambda x, *shape: x.detach().reshape(*shape).contiguous()  # noqa: E731
            ops.ppo_loss(flat(mu, B, A), ac.std.detach(), flat(value, B), flat(actions, B, A), flat(logp, B), flat(mu_old, B, A), flat(sigma_old, B, A),
                         flat(adv, B), flat(returns, B), flat(values, B), clip_param=self.clip_param, value_loss_coef=self.value_loss_coef,
                         entropy_coef=self.entropy_coef, use_clipped_value_loss=self.use_clipped_value_loss,
                         desired_kl=self.desired_kl if local_lr else None, lr=opt.lr_t if local_lr else None, loss_accum=self._loss_accum, buffers=bufs)
            torch.autograd.backward([mu, value], [bufs.grad_mu.view_as(mu), bufs.grad_value.view_as(value)])
            opt.grads[off:off + n].add_(bufs.grad_sigma)
            if adaptive and world > 1:
                ac.flat_grads_ext[-4:-3].copy_(bufs.out[4:5])
            self.reduce_and_step()
        self.peer_update_fence()
        st.clear()

    # ---------------------------------------------------------------------------------- K14: peer-memory gradient exchange
    def enable_peer_gradients(self, group=None) -> bool:
        """Puts the flat gradient buffer into NVLink-mapped symmetric memory (``torch.distributed._symmetric_memory``) so that
        every rank can read every other rank's gradients directly: the per-mini-batch NCCL all-reduce becomes a cross-GPU barrier
        + ONE kernel that sums the W buffers in rank order while it computes the clip norm (K14).  Returns False (and leaves the
        NCCL path in place) when there is one process, when symmetric memory cannot be set up, or when LT_PEER_GRADS=0.
        Call before capturing CUDA graphs."""
        import os

        self._peer = None
        _, world = self._world_info()
        mode = os.environ.get("LT_PEER_GRADS", "auto")
        if world == 1 or mode == "0" or world > _C_MAX_PEERS:
            return False
        try:
            import torch.distributed._symmetric_memory as symm

            ac = self.actor_critic
            ac.flatten_parameters()
            group = group if group is not None else dist.group.WORLD
            L = ac.flat_grads_ext.numel()
            dev = ac.flat_grads_ext.device
            # one symmetric allocation: two gradient buffers (double-buffered exchange) + 16 doubles for the advantage statistics
            big = symm.empty(2 * L + 32, dtype=torch.float32, device=dev)
            hdl = symm.rendezvous(big, group)
            big.zero_()
            bufs = [big[:L], big[L:2 * L]]
            base = [int(x) for x in hdl.buffer_ptrs]
            self.optimizer.use_gradient_buffer(bufs[0])
            world = len(base)
            stats = [hdl.get_buffer(r, (4,), torch.float64, storage_offset=(2 * L * 4) // 8) for r in range(world)] if (2 * L * 4) % 8 == 0 else None
            mode = os.environ.get("LT_PEER_TWO_SHOT", "auto")
            self._peer = dict(handle=hdl, ptrs=[base, [b + 4 * L for b in base]], bufs=bufs, sum=torch.zeros(L, device=dev), buf=big, k=0,
                              adv_stats=stats, rank=dist.get_rank(group),
                              two_shot=(mode == "1") or (mode == "auto" and world > 4 and ops._adam_launches(L - 4, dev) == 1))
            if stats is not None:
                self._adv_stats = stats[self._peer["rank"]]  # gae_scan writes the local statistics where the peers can read them
        except Exception as exc:  # noqa: BLE001 -- any failure leaves the NCCL path untouched
            self._peer = None
            self._peer_error = repr(exc)
            return False
        return True

    @staticmethod
    def attach_local_peers(learners: "list[PPO]") -> None:
        """Makes ``learners`` -- replicas of one policy that live in THIS process, one per env shard, on one device or on
        peer-accessible devices -- an env-sharded job of world size ``len(learners)``: every learner's flat gradient buffer is an
        ordinary device allocation whose address all the others pass to ``lt_peer_sum_clip_adam`` (K14 takes raw pointers; symmetric
        memory is only how SEPARATE processes obtain them).  The caller runs the learners in lock step on one stream -- all
        ``minibatch_grads(i)`` before any ``step_after_reduce()`` -- which is what the cross-GPU barriers enforce between processes.
        Advantage statistics: ``compute_returns_scan`` on every learner, ``sum_local_adv_stats(learners)``, then
        ``compute_returns_normalize``."""
        W = len(learners)
        if not 1 <= W <= _C_MAX_PEERS:
            raise ValueError(f"1..{_C_MAX_PEERS} learners")
        bufs = []
        for alg in learners:
            ac = alg.actor_critic
            ac.flatten_parameters()
            pair = [torch.zeros(ac.flat_grads_ext.numel(), dtype=torch.float32, device=ac.flat_grads_ext.device) for _ in range(2)]
            alg.optimizer.use_gradient_buffer(pair[0])
            bufs.append(pair)
        for r, alg in enumerate(learners):
            alg._world_override = (r, W)
            alg._peer = dict(handle=_NoBarrier(), ptrs=[[b[q].data_ptr() for b in bufs] for q in range(2)], bufs=bufs[r], sum=torch.zeros_like(bufs[r][0]),
                             buf=None, k=0, adv_stats=None, rank=r) if W > 1 else None

    @staticmethod
    def sum_local_adv_stats(learners: "list[PPO]") -> None:
        """The 3-double exchange of ``compute_returns`` for same-process replicas (rank order, like the all-reduce of equal shards)."""
        total = learners[0]._adv_stats.clone()
        for alg in learners[1:]:
            total += alg._adv_stats.to(total.device)
        for alg in learners:
            alg._adv_stats.copy_(total)

    @property
    def peer_gradients(self) -> bool:
        return getattr(self, "_peer", None) is not None

    def reduce_and_step(self):
        """[NCCL: ONE all-reduce of the flat gradients with the KL statistic in its tail] then the learning-rate decision and
        the fused clip + Adam (K7)."""
        self.allreduce_grads()
        self.step_after_reduce()
        self.after_step_barrier()
        if self.trace is not None:  # diagnostics only (host reads): local KL of this mini-batch, learning rate after the decision
            self.trace.append((float(self._loss_bufs.out[4]), float(self.optimizer.lr_t), float(self.optimizer.grad_norm)))

    def allreduce_grads(self):
        """The only collective of a mini-batch step (kept outside CUDA-graph capture)."""
        _, world = self._world_info()
        if self.peer_gradients:
            self._peer["handle"].barrier(channel=0)  # every rank's backward has written its gradient buffer
        elif world > 1:
            D.average_gradients_(self.actor_critic.flat_grads_ext)

    def step_after_reduce(self):
        """Learning-rate decision from the reduced KL statistic + clip + Adam; collective-free, capturable."""
        opt = self.optimizer
        _, world = self._world_info()
        adaptive = self.desired_kl is not None and self.schedule == "adaptive"
        if self.peer_gradients:
            pr = self._peer
            # the summed KL statistic sits behind the summed gradients: the first kernel takes the learning-rate decision
            # (every rank the same one), the second applies it
            ptrs = pr["ptrs"][pr["k"] % 2]
            two_shot = pr.get("two_shot", False)
            if two_shot:
                # larger worlds: every rank first reduces ITS slice of the W buffers in place (W - 1 remote slices), a second
                # barrier, then the optimizer kernel gathers the W reduced slices: 2 (W - 1) / W buffers per rank instead of W - 1
                ops.peer_reduce_scatter(ptrs, pr["rank"], opt.flat.numel())
                pr["handle"].barrier(channel=3)
            opt.step_peer_sum(ptrs, pr["sum"], 4, max_grad_norm=self.max_grad_norm, grad_scale=1.0 / world,
                              desired_kl=self.desired_kl if adaptive else None, kl_scale=1.0 / world, gather=two_shot)
            pr["k"] += 1
            return
        if world > 1 and adaptive:  # every rank takes the same decision
            ops.adaptive_lr(self.actor_critic.flat_grads_ext[-4:-3], 1.0 / world, self.desired_kl, opt.lr_t)
        opt.step(max_grad_norm=self.max_grad_norm, grad_scale=1.0 / world)

    def after_step_barrier(self):
        """K14: nothing to wait for -- the gradient buffers alternate per mini-batch (see ``minibatch_grads``), so the barrier in
        front of the NEXT mini-batch's peer reads already orders the rewrite of this one's buffer after everybody's reads."""

    def peer_update_fence(self):
        """End of an update under K14: one barrier, so that the first two mini-batches of the next update (which reuse both
        buffers) cannot overtake a rank that is still reading in this update's last step; it also re-arms the buffer parity."""
        if self.peer_gradients:
            self._peer["handle"].barrier(channel=1)
            self._peer["k"] = 0

    def update_epilogue(self):
        """The only device->host read of an update: the three logged means (reference ppo.py:361-363 reads them with
        ``.item()`` after every mini-batch) and the learning rate."""
        acc = self._loss_accum.tolist()
        n = max(acc[3], 1.0)
        opt = self.optimizer
        self.learning_rate = float(opt.lr_t.item())
        opt.param_groups[0]["lr"] = self.learning_rate
        opt._host_lr = self.learning_rate
        return acc[0] / n, acc[1] / n, acc[2] / n, None, None
