"""PPO with the reference's interface (reference loco_rl/loco_rl/algorithms/ppo.py:19-385).

Same constructor keywords, attributes (``actor_critic``, ``optimizer``, ``storage``, ``transition``, ``learning_rate``,
``rnd``, ``symmetry``) and methods (``init_storage``, ``test_mode`` / ``train_mode``, ``act``, ``process_env_step``,
``compute_returns``, ``update`` returning the 5-tuple of ppo.py:385), so ``OnPolicyRunner`` drives it unchanged.

What changed below the interface:
* ``act`` writes actions / log-prob / mu / sigma / values directly into the RolloutStorage slot (K3);
* ``process_env_step`` fuses the time-out bootstrap into the store kernel (K3);
* ``compute_returns`` is the GAE kernel pair (K4);
* ``update``: one fused gather of the permuted rollout (K5); per mini-batch the cuBLAS MLP forward, ONE fused loss
  kernel (K6: log-prob, entropy, KL, adaptive learning rate on the device, clipped surrogate + value loss and their
  analytic gradients), autograd only through the two MLPs, ONE fused clip + Adam over the flat parameter buffer (K7).
  No ``.item()`` inside the loop: the learning rate lives on the device and the three logged means are accumulated on
  the device and read back once per update.  With ``torch.distributed`` initialised, gradients are all-reduced as one
  flat NCCL buffer and the KL statistic is all-reduced before the learning-rate decision.
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

    def use_gradient_buffer(self, buf: torch.Tensor):
        self.grads = self.ac.rebind_gradients(buf)

    def step_peer_sum(self, peer_ptrs, grad_sum, tail: int, max_grad_norm=None, grad_scale: float = 1.0, desired_kl=None, kl_scale=1.0):
        """K14: sum of the ranks' gradient buffers (peer loads, rank order) + learning-rate decision + clip + Adam."""
        g = self.param_groups[0]
        ops.peer_sum_clip_adam(self.flat, peer_ptrs, grad_sum, tail, self.exp_avg, self.exp_avg_sq, self.lr_t, self.step_t,
                               max_grad_norm=max_grad_norm, betas=g["betas"], eps=g["eps"], weight_decay=g["weight_decay"],
                               grad_scale=grad_scale, desired_kl=desired_kl, kl_scale=kl_scale, grad_norm_out=self.grad_norm)

    def sync_lr_to_device(self):
        lr = self.param_groups[0]["lr"]
        if lr != self._host_lr:  # someone (runner, user) wrote param_groups[...]["lr"]
            self.lr_t.fill_(lr)
            self._host_lr = lr

    def zero_grad(self, set_to_none: bool = False):
        self.grads.zero_()

    def step(self, max_grad_norm=None, grad_scale: float = 1.0):
        g = self.param_groups[0]
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
        self.optimizer = _FusedAdam(self.actor_critic, learning_rate)
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
        with torch.no_grad():
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

    def compute_returns(self, last_critic_obs):
        with torch.no_grad():
            last_values = self.actor_critic.evaluate(last_critic_obs).detach()
        st = self.storage
        normalize = not self.normalize_advantage_per_mini_batch
        world = dist.get_world_size() if dist.is_available() and dist.is_initialized() else 1
        if world == 1 or not normalize or not self.global_advantage_normalization:
            st.compute_returns(last_values, self.gamma, self.lam, normalize_advantage=normalize)
            return
        # env-sharded ranks: sum / sum of squares / count are all-reduced between the scan and the normalisation
        ops.gae_scan(st.rewards, st.values, st.dones, last_values.contiguous().view(-1), self.gamma, self.lam, st.returns, st.advantages, self._adv_stats)
        D.reduce_adv_stats_(self._adv_stats)
        ops.adv_normalize(st.advantages, self._adv_stats)

    # ---------------------------------------------------------------------------------------------------------- update
    def update(self, indices=None):
        """reference ppo.py:179-385.  ``indices``: optional explicit permutation (the reference draws it with randperm)."""
        self.optimizer.sync_lr_to_device()
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
        self.storage.clear()

    # The three stages below are public so that a caller can capture them separately (CUDA graphs must not contain the
    # NCCL collectives of ``reduce_and_step`` when several processes train together).
    def update_begin(self, indices=None):
        """Draws / takes the permutation (rollout_storage.py:189) and gathers the permuted rollout once (K5)."""
        ac, st = self.actor_critic, self.storage
        if ac.is_recurrent:
            raise NotImplementedError("recurrent policies are outside the LocoTouch hot path")
        if self.normalize_advantage_per_mini_batch:
            raise NotImplementedError("per-mini-batch advantage normalisation is not used by the LocoTouch cfgs")
        if ac.noise_std_type != "scalar":
            raise NotImplementedError("noise_std_type='log' is not used by the LocoTouch cfgs")
        ac.flatten_parameters()
        self._loss_accum.zero_()
        batch_size = st.num_envs * st.num_transitions_per_env
        self._mb_size = batch_size // self.num_mini_batches
        if indices is None:
            indices = torch.randperm(self.num_mini_batches * self._mb_size, requires_grad=False, device=self.device)
        bufs, has_priv = st.gather_permuted(indices)
        it = iter(bufs)
        obs = next(it)
        cobs = next(it) if has_priv else obs
        self._mb = (obs, cobs) + tuple(next(it) for _ in range(7))

    def minibatch_grads(self, i: int):
        """Forward of both MLPs on mini-batch slice ``i``, fused loss (K6), backward into the flat gradient buffer."""
        ac, opt = self.actor_critic, self.optimizer
        sl = slice(i * self._mb_size, (i + 1) * self._mb_size)
        obs, cobs, actions, values, returns, logp, adv, mu_old, sigma_old = (t[sl] for t in self._mb)
        B, A = actions.shape
        if self._loss_bufs is None or self._loss_bufs.B != B:
            self._loss_bufs = ops.PpoLossBuffers(B, A, self.device)
        bufs = self._loss_bufs
        adaptive = self.desired_kl is not None and self.schedule == "adaptive"
        _, world = D.world_info()
        local_lr = adaptive and world == 1  # one process: the learning-rate decision is taken inside the loss kernel
        explicit = ac.supports_explicit_backward
        if explicit:
            mu, value = ac.train_forward(obs, cobs)
        else:  # non-ELU activations: autograd through the torch modules
            mu = ac.actor(obs)
            value = ac.critic(cobs)
            opt.zero_grad()
        ops.ppo_loss(mu.detach(), ac.std.detach(), value.detach().view(-1), actions, logp.view(-1), mu_old, sigma_old, adv.view(-1),
                     returns.view(-1), values.view(-1), clip_param=self.clip_param, value_loss_coef=self.value_loss_coef,
                     entropy_coef=self.entropy_coef, use_clipped_value_loss=self.use_clipped_value_loss,
                     desired_kl=self.desired_kl if local_lr else None, lr=opt.lr_t if local_lr else None, loss_accum=self._loss_accum, buffers=bufs)
        off, n = ac._slices["std"]
        if explicit:
            ac.train_backward(bufs.grad_mu, bufs.grad_value.view(-1, 1))
            opt.grads[off:off + n].copy_(bufs.grad_sigma)  # every gradient slot is overwritten: no zero_grad pass
        else:
            torch.autograd.backward([mu, value], [bufs.grad_mu, bufs.grad_value.view_as(value)])
            opt.grads[off:off + n].add_(bufs.grad_sigma)
        if adaptive and world > 1:  # the local KL mean travels in the tail of the gradient all-reduce
            ac.flat_grads_ext[-4:-3].copy_(bufs.out[4:5])

    # ---------------------------------------------------------------------------------- K14: peer-memory gradient exchange
    def enable_peer_gradients(self, group=None) -> bool:
        """Puts the flat gradient buffer into NVLink-mapped symmetric memory (``torch.distributed._symmetric_memory``) so that
        every rank can read every other rank's gradients directly: the per-mini-batch NCCL all-reduce becomes a cross-GPU barrier
        + ONE kernel that sums the W buffers in rank order while it computes the clip norm (K14).  Returns False (and leaves the
        NCCL path in place) when there is one process, when symmetric memory cannot be set up, or when LT_PEER_GRADS=0.
        Call before capturing CUDA graphs."""
        import os

        self._peer = None
        _, world = D.world_info()
        mode = os.environ.get("LT_PEER_GRADS", "auto")
        if world == 1 or mode == "0" or world > _C_MAX_PEERS:
            return False
        try:
            import torch.distributed._symmetric_memory as symm

            ac = self.actor_critic
            ac.flatten_parameters()
            group = group if group is not None else dist.group.WORLD
            buf = symm.empty(ac.flat_grads_ext.numel(), dtype=torch.float32, device=ac.flat_grads_ext.device)
            hdl = symm.rendezvous(buf, group)
            self.optimizer.use_gradient_buffer(buf)
            self._peer = dict(handle=hdl, ptrs=[int(x) for x in hdl.buffer_ptrs], sum=torch.zeros_like(buf), buf=buf)
        except Exception as exc:  # noqa: BLE001 -- any failure leaves the NCCL path untouched
            self._peer = None
            self._peer_error = repr(exc)
            return False
        return True

    @property
    def peer_gradients(self) -> bool:
        return getattr(self, "_peer", None) is not None

    def reduce_and_step(self):
        """[NCCL: ONE all-reduce of the flat gradients with the KL statistic in its tail] then the learning-rate decision and
        the fused clip + Adam (K7)."""
        self.allreduce_grads()
        self.step_after_reduce()
        self.after_step_barrier()

    def allreduce_grads(self):
        """The only collective of a mini-batch step (kept outside CUDA-graph capture)."""
        _, world = D.world_info()
        if self.peer_gradients:
            self._peer["handle"].barrier(channel=0)  # every rank's backward has written its gradient buffer
        elif world > 1:
            D.average_gradients_(self.actor_critic.flat_grads_ext)

    def step_after_reduce(self):
        """Learning-rate decision from the reduced KL statistic + clip + Adam; collective-free, capturable."""
        opt = self.optimizer
        _, world = D.world_info()
        adaptive = self.desired_kl is not None and self.schedule == "adaptive"
        if self.peer_gradients:
            pr = self._peer
            # the summed KL statistic sits behind the summed gradients: the first kernel takes the learning-rate decision
            # (every rank the same one), the second applies it
            opt.step_peer_sum(pr["ptrs"], pr["sum"], 4, max_grad_norm=self.max_grad_norm, grad_scale=1.0 / world,
                              desired_kl=self.desired_kl if adaptive else None, kl_scale=1.0 / world)
            return
        if world > 1 and adaptive:  # every rank takes the same decision
            ops.adaptive_lr(self.actor_critic.flat_grads_ext[-4:-3], 1.0 / world, self.desired_kl, opt.lr_t)
        opt.step(max_grad_norm=self.max_grad_norm, grad_scale=1.0 / world)

    def after_step_barrier(self):
        """K14: nobody may overwrite its gradient buffer (next backward) before every rank has read it."""
        if self.peer_gradients:
            self._peer["handle"].barrier(channel=1)

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
