"""ActorCritic with the reference's interface (reference loco_rl/loco_rl/modules/actor_critic.py:8-144).

Same constructor signature, ``state_dict`` keys (``std`` | ``log_std``, ``actor.{0,2,4,6}.*``, ``critic.*``) and methods,
so reference checkpoints load unchanged.  Differences are below the interface:

* all parameters are views into ONE flat fp32 buffer (``flat_params``) with a matching flat gradient buffer, so that the
  gradient all-reduce is a single NCCL call and clip + Adam is a single fused kernel (K7);
* ``act`` / ``get_actions_log_prob`` run the fused act epilogue (K3) instead of ``torch.distributions.Normal``: no
  distribution object, no expanded-std tensor, outputs can be written straight into a RolloutStorage slot.
In TF32 mode (allowed exactly like reference locotouch/scripts/train.py:66-69) the MLP GEMMs are this library's tcgen05 kernels
(K12 forward / dgrad with fused epilogues, K15 weight + bias gradients); fp32 parity mode keeps cuBLAS.
"""
from __future__ import annotations

import math
import os

import torch
import torch.nn as nn
import torch.nn.functional as F

from ... import _C, ops
from ...streams import SideStream
from ..utils import resolve_nn_activation


class _GaussianView:
    """Minimal stand-in for the ``self.distribution`` attribute some callers poke at (mean / stddev / entropy)."""

    def __init__(self, mean, std_rows):
        self.mean, self.stddev = mean, std_rows
        self.loc, self.scale = mean, std_rows

    def entropy(self):
        return 0.5 + 0.5 * math.log(2 * math.pi) + torch.log(self.stddev)

    def log_prob(self, value):
        var = self.stddev**2
        return -((value - self.mean) ** 2) / (2 * var) - self.stddev.log() - math.log(math.sqrt(2 * math.pi))

    def sample(self):
        return torch.normal(self.mean, self.stddev)


class ActorCritic(nn.Module):
    is_recurrent = False

    def __init__(self, num_actor_obs, num_critic_obs, num_actions, actor_hidden_dims=[256, 256, 256], critic_hidden_dims=[256, 256, 256],
                 activation="elu", init_noise_std=1.0, noise_std_type: str = "scalar", **kwargs):
        if kwargs:
            print("ActorCritic.__init__ got unexpected arguments, which will be ignored: " + str([key for key in kwargs.keys()]))
        super().__init__()
        act = resolve_nn_activation(activation)

        def mlp(inp, hidden, out):
            layers, prev = [], inp
            for h in hidden:
                layers += [nn.Linear(prev, h), act]
                prev = h
            layers.append(nn.Linear(prev, out))
            return nn.Sequential(*layers)

        self.actor = mlp(num_actor_obs, actor_hidden_dims, num_actions)
        self.critic = mlp(num_critic_obs, critic_hidden_dims, 1)
        self.init_noise_std = init_noise_std
        self.noise_std_type = noise_std_type
        if noise_std_type == "scalar":
            self.std = nn.Parameter(init_noise_std * torch.ones(num_actions))
        elif noise_std_type == "log":
            self.log_std = nn.Parameter(torch.log(init_noise_std * torch.ones(num_actions)))
        else:
            raise ValueError(f"Unknown standard deviation type: {self.noise_std_type}. Should be 'scalar' or 'log'")
        self.num_actions = num_actions
        if num_actions % 4 or num_actions > 64:  # lt_act_sample / lt_ppo_loss / K16 move one float4 of every [A] row per lane
            raise ValueError(f"locotouch_b200.ActorCritic: num_actions = {num_actions} is not supported by the sampling / PPO-loss kernels "
                             "(a multiple of 4, at most 64; the LocoTouch tasks use 12)")
        self.distribution = None
        self.flat_params = None
        self.flat_grads = None
        self._logp = None
        self._sampled = None
        self.rng = "philox"  # "torch": eps = torch.randn (same generator stream as torch.normal); "philox": in-kernel
        self.seed = 0
        self._draws = 0
        self._offset_base = None  # optional device int64 counter added to the Philox offset (CUDA-graph replays)
        self._graph_slot = 0

    # ------------------------------------------------------------------------------------------ flat parameter storage
    def flatten_parameters(self):
        """Re-homes every parameter (and its .grad) as a view of one 16-byte aligned flat buffer.  Idempotent."""
        params = list(self.parameters())
        dev = params[0].device
        if self.flat_params is not None and self.flat_params.device == dev and all(p.data.data_ptr() == self._offsets[i][0] for i, p in enumerate(params)):
            return self.flat_params, self.flat_grads
        total = sum((p.numel() + 3) // 4 * 4 for p in params)  # every tensor starts 16-byte aligned
        flat = torch.zeros(total, device=dev, dtype=torch.float32)
        # four spare floats after the gradients: statistics that must be summed over ranks (the KL mean) ride along with the
        # gradient all-reduce instead of paying for a collective of their own
        self.flat_grads_ext = torch.zeros(total + 4, device=dev, dtype=torch.float32)
        grads = self.flat_grads_ext[:total]
        off, self._offsets, self._slices = 0, [], {}
        names = [k for k, _ in self.named_parameters()]
        for name, p in zip(names, params):
            n = p.numel()
            flat[off:off + n].copy_(p.data.flatten())
            p.data = flat[off:off + n].view(p.shape)
            p.grad = grads[off:off + n].view(p.shape)
            self._offsets.append((p.data.data_ptr(), off, n))
            self._slices[name] = (off, n)
            off += (n + 3) // 4 * 4
        self.flat_params, self.flat_grads = flat, grads
        return flat, grads

    def rebind_gradients(self, buf: torch.Tensor, copy: bool = True):
        """Moves the flat gradient storage (and every ``p.grad`` view) into ``buf`` [total + 4] -- e.g. a buffer in NVLink-mapped
        symmetric memory that the other ranks read directly (K14).  ``copy=False``: the old contents are not carried over (the
        double-buffered exchange alternates between two such buffers, one per mini-batch; every backward rewrites all of it)."""
        self.flatten_parameters()
        if buf.numel() != self.flat_grads_ext.numel() or buf.dtype != torch.float32 or buf.device != self.flat_grads_ext.device:
            raise ValueError("gradient buffer must be float32 [total + 4] on the parameters' device")
        if buf.data_ptr() == self.flat_grads_ext.data_ptr():
            return self.flat_grads
        if copy:
            buf.copy_(self.flat_grads_ext)
        self.flat_grads_ext = buf
        total = self.flat_params.numel()
        grads = buf[:total]
        for name, p in self.named_parameters():
            off, n = self._slices[name]
            p.grad = grads[off:off + n].view(p.shape)
        self.flat_grads = grads
        return grads

    def _apply(self, fn, *a, **k):
        """``.to()`` / ``.cuda()`` / ``.float()``: the flat buffers are dropped only when the parameters really moved (another
        device or dtype re-creates every tensor); a no-op conversion keeps them -- the optimizer and a peer-mapped gradient buffer
        hold references to them (``_FusedAdam.refresh`` re-attaches after a real move)."""
        out = super()._apply(fn, *a, **k)
        if self.flat_params is not None:
            params = list(self.parameters())
            still_views = (len(params) == len(self._offsets) and params[0].device == self.flat_params.device
                           and all(p.dtype == torch.float32 and p.data.data_ptr() == self._offsets[i][0] for i, p in enumerate(params)))
            if not still_views:
                self.flat_params = None
            else:  # torch may have replaced p.grad by a converted copy: point it at the flat gradient buffer again
                for (name, p) in self.named_parameters():
                    off, n = self._slices[name]
                    if p.grad is None or p.grad.data_ptr() != self.flat_grads[off:off + n].data_ptr():
                        p.grad = self.flat_grads[off:off + n].view(p.shape)
        return out

    # ------------------------------------------------------------------------------ explicit training forward / backward
    def _elu_stack(self, net) -> bool:
        mods = list(net)
        return all(isinstance(m, (nn.Linear, nn.ELU)) for m in mods) and all(m.alpha == 1.0 for m in mods if isinstance(m, nn.ELU))

    @property
    def supports_explicit_backward(self) -> bool:
        return self.flat_params is not None and self._elu_stack(self.actor) and self._elu_stack(self.critic)

    def side_streams(self, device):
        """(critic chain, actor wgrad, critic wgrad) side streams, created on first use."""
        if getattr(self, "_side", None) is None or self._side[0].device != torch.device(device):
            self._side = tuple(SideStream(device) for _ in range(3))
        return self._side

    def _train_buffers(self, batch: int, device):
        """Persistent activation / gradient buffers of the explicit training pass (keyed by batch size): nothing is allocated
        inside the two concurrently running chains, so stream-ordered reuse by the caching allocator cannot alias them."""
        key = (batch, str(device))
        if getattr(self, "_train_buf_key", None) != key:
            bufs = []
            for net in (self.actor, self.critic):
                linears = [m for m in net if isinstance(m, nn.Linear)]
                hs = [torch.empty(batch, lin.out_features, device=device) for lin in linears]
                gs = [None] + [torch.empty(batch, lin.in_features, device=device) for lin in linears[1:]]
                for lin in linears:  # split-K partial products of the weight gradients
                    lin._wgrad_part = (torch.empty(self._WGRAD_SPLIT, lin.out_features, lin.in_features, device=device)
                                       if self._wgrad_split_ok(batch, lin.out_features) else None)
                bufs.append((linears, hs, gs))
            self._train_bufs, self._train_buf_key = bufs, key
        return self._train_bufs

    @property
    def supports_fused_heads(self) -> bool:
        """K16 (heads + PPO loss + head dgrad in one kernel) takes this module: explicit backward, scalar std, supported widths."""
        if not self.supports_explicit_backward or self.noise_std_type != "scalar":
            return False
        la = [m for m in self.actor if isinstance(m, nn.Linear)]
        lc = [m for m in self.critic if isinstance(m, nn.Linear)]
        return len(la) >= 2 and len(lc) >= 2 and ops.ppo_heads_supported(self.num_actions, la[-1].in_features, lc[-1].in_features)

    def _pad4(self, t, tag: str):
        """[rows, k] -> a persistent [rows, ceil4(k)] copy with zero tail columns (None when k is already a multiple of 4).  TMA needs a
        16-byte row pitch: the locomotion task's 270-wide observations / first-layer weights reach K12 / K15 through this copy (one
        strided-copy launch) instead of falling back to cuBLAS + ELU + split-K + K9."""
        k = t.shape[1]
        kp = (k + 3) // 4 * 4
        if kp == k:
            return None
        if os.environ.get("LT_PAD_K", "1") == "0":  # measurement knob: keep the cuBLAS path of round 1 for unaligned K
            return t
        if tag is None:  # one-off copy owned by the caller
            out = torch.zeros(t.shape[0], kp, device=t.device)
            out[:, :k].copy_(t)
            return out
        cache = self.__dict__.setdefault("_pad_bufs", {})
        key = (tag, t.shape[0], kp, str(t.device))
        buf = cache.get(key)
        if buf is None:
            buf = cache[key] = torch.zeros(t.shape[0], kp, device=t.device)
        buf[:, :k].copy_(t)
        return buf

    @torch.no_grad()
    def train_forward(self, observations, critic_observations, heads: bool = True):
        """Forward of both MLPs keeping the post-activation tensors (K12: tcgen05 GEMM with bias + ELU in its epilogue; fp32 mode:
        cuBLAS GEMM with fused bias, ELU in place); no autograd graph is built -- ``train_backward`` produces the parameter
        gradients explicitly.  ``heads=False`` stops at the last hidden layers and returns their activations (h_actor, h_critic):
        the head layers then belong to K16 (``ops.ppo_heads_loss``)."""
        bufs = self._train_buffers(observations.shape[0], observations.device)
        self._saved = []
        outs = []

        fused = torch.backends.cuda.matmul.allow_tf32  # K12 computes in TF32, like cuBLAS does under the reference's train.py:66-69

        def chain(linears, hs, x, tag, acts=None):
            acts = [x] if acts is None else acts
            h, first = acts[-1], len(acts) - 1  # K19 may already have produced the first three hidden layers
            for i, lin in enumerate(linears):
                if i < first:
                    continue
                hidden = i < len(linears) - 1
                if not hidden and not heads:
                    break
                out = None
                if fused and lin.out_features >= 64:  # one tcgen05 GEMM with bias + ELU in the epilogue (K12)
                    w = lin.weight
                    if i == 0 and (lin.in_features & 3):  # e.g. K = 270: zero-padded copies with a 16-byte row pitch
                        if h.shape[1] == lin.in_features:  # (PPO pads the gathered rollout once per update: nothing to copy then)
                            h = acts[0] = self._pad4(h, "x" + tag)
                        w = self._pad4(w, "w" + tag)
                    out = ops.linear_bias_act(h, w, lin.bias, out=hs[i], elu=hidden)
                if out is None:  # fp32 mode, narrow output layers, unaligned K: cuBLAS GEMM with fused bias, ELU in place
                    out = torch.addmm(lin.bias, h, lin.weight.t(), out=hs[i])
                    if hidden:
                        F.elu_(out)
                h = out
                acts.append(h)
            return acts

        # the critic chain runs on a side stream next to the actor chain (eagerly and as a parallel branch of a captured graph):
        # the GEMMs of one network overlap the memory-bound ELU passes of the other
        side = self.side_streams(observations.device)[0]
        pre = self._mlp3_hidden([(bufs[0][0], observations, "a", bufs[0][1]), (bufs[1][0], critic_observations, "c", bufs[1][1])], keep=True) if fused else None
        for k, ((linears, hs, gs), x) in enumerate(zip(bufs, (observations, critic_observations))):
            if k == 0:
                acts = chain(linears, hs, x, "a", pre[0] if pre else None)
            elif pre and (not heads or len(linears) == 4):  # nothing (or one small head GEMM) left: not worth a forked branch
                acts = chain(linears, hs, x, "c", pre[1])
            else:
                with side.forked():
                    acts = chain(linears, hs, x, "c", pre[1] if pre else None)
                side.join()
            self._saved.append((linears, acts, gs))
            outs.append(acts[-1])
        return outs[0], outs[1]

    def _mlp3_hidden(self, nets, keep: bool):
        """K19: the three hidden layers [512, 256, 128] of actor AND critic in one persistent tcgen05 kernel (x slab resident in shared
        memory, activations handed from layer to layer inside TMEM).  ``nets`` = [(linears, x, tag, hs)] with ``hs`` the [B, 512] /
        [B, 256] / [B, 128] buffers; ``keep`` stores h1 / h2 as well (the training pass needs them).  Returns per net the list
        [x (padded to a multiple of 4 columns if need be), h1, h2, h3], or None when a stack is not taken (callers keep K12 per layer)."""
        args, outs = [], []
        for linears, x, tag, hs in nets:
            if len(linears) < 4 or x.dim() != 2 or not x.is_cuda or x.dtype != torch.float32:  # three hidden layers + a head
                return None
            lin1, w1 = linears[0], linears[0].weight
            if lin1.in_features & 3:  # e.g. K = 270: zero-padded copies with a 16-byte row pitch (like the per-layer path)
                if x.shape[1] == lin1.in_features:
                    x = self._pad4(x, "x" + tag)
                w1 = self._pad4(w1, "w" + tag)
            if not ops.mlp3_supported(linears, x.shape[1]) or w1.shape[1] != x.shape[1]:
                return None
            x = x.contiguous()
            args.append((x, (w1, lin1.bias, linears[1].weight, linears[1].bias, linears[2].weight, linears[2].bias),
                         (hs[0] if keep else None, hs[1] if keep else None, hs[2])))
            outs.append([x, hs[0], hs[1], hs[2]])
        if len({a[0].shape for a in args}) != 1:  # the two networks share one launch only when their inputs have the same shape
            return None
        return outs if ops.mlp3_forward(args) is not None else None

    def hidden_grad_buffers(self):
        """(g_h_actor, g_h_critic): where the gradient w.r.t. the pre-activation of the last hidden layers goes (K16 output)."""
        return tuple(gs[-1] for (_, _, gs) in self._saved)

    @torch.no_grad()
    def train_backward(self, grad_mu, grad_value, from_hidden: bool = False):
        """Writes dLoss/dW and dLoss/db of every layer straight into the flat gradient buffer.  TF32 mode: per layer one K15 launch
        (weight + bias gradient) on a trailing stream and one K12 dgrad GEMM with the ELU backward in its epilogue; fp32 mode: K9 +
        cuBLAS.  ``from_hidden``: the gradients of the last hidden layers already sit in ``hidden_grad_buffers()`` (K16 wrote
        them), so the head layers only need their weight gradients."""

        fused = torch.backends.cuda.matmul.allow_tf32
        if fused:
            # K15 ADDS the batch slices into the gradient views (vector reductions): clear the flat buffer once, ahead of both
            # chains, instead of one memset node per layer
            self.flat_grads.zero_()

        def chain(linears, acts, gs, g, wstream):
            # dgrad feeds the next layer; the weight gradients only have to be complete before the optimizer step, so they
            # trail on their own stream.  acts[i] is the input of layer i = the post-ELU output of layer i - 1.
            # TF32 mode: K15 produces dW and db of a layer from ONE read of g (the bias gradient is summed from the tiles in
            # shared memory), so no separate reduction pass (K9) runs; fp32 parity mode keeps K9 + cuBLAS.
            last = len(linears) - 1
            have_bias = False  # bias gradient of the layer that produced g already written (by the K9 pass of the unfused path)
            for i in range(last, -1, -1):
                lin = linears[i]
                if not self._wgrad_launch(g, acts[i], lin, wstream, with_bias=not have_bias) and not have_bias:
                    ops.bias_act_bwd(g, None, lin.bias.grad)            # K9, reduction only (bias gradient = column sums)
                have_bias = False
                if i == last and from_hidden:                           # K16 already produced the gradient below the head
                    g = gs[i]
                    continue
                if i > 0:
                    out = None
                    if fused and lin.out_features >= 64:  # K12: dgrad GEMM with the ELU backward of the layer below in its epilogue
                        out = ops.dgrad_act_bwd(g, lin.weight, acts[i], out=gs[i])
                    if out is None:
                        out = torch.mm(g, lin.weight, out=gs[i])
                        ops.bias_act_bwd(out, acts[i], linears[i - 1].bias.grad)  # K9: ELU backward in place + bias gradient
                        have_bias = True
                    g = out

        if fused and self._train_backward_paired(grad_mu, grad_value, from_hidden):
            self._saved = None
            return

        side, w_actor, w_critic = self.side_streams(grad_mu.device)
        for k, ((linears, acts, gs), g) in enumerate(zip(self._saved, (grad_mu, grad_value))):
            if k == 0:
                chain(linears, acts, gs, g, w_actor)
            else:
                with side.forked():
                    chain(linears, acts, gs, g, w_critic)
                    w_critic.join()
        w_actor.join()
        side.join()
        self._saved = None

    def _train_backward_paired(self, grad_mu, grad_value, from_hidden: bool) -> bool:
        """TF32 backward of two networks with identical hidden stacks: layer i of actor and critic share ONE K15 launch
        (``ops.wgrad_pair``: the CTAs are divided between the two problems, the fixed cost of a launch is paid once) on the trailing
        stream; the dgrad GEMMs (K12, ELU backward in the epilogue) run actor on the caller's stream, critic on the side stream.
        Returns False (nothing launched) when the stacks differ or an input is a zero-padded copy -- the per-network path then runs."""
        (la, acts_a, gs_a), (lc, acts_c, gs_c) = self._saved
        if len(la) != len(lc) or len(la) < 2 or os.environ.get("LT_WGRAD_PAIR", "1") == "0" or _C.gemm_backend() == "stub":
            return False
        for a, c in zip(la[:-1], lc[:-1]):
            if (a.in_features, a.out_features) != (c.in_features, c.out_features) or a.out_features < 64:
                return False
        if acts_a[0].shape[1] != la[0].in_features or acts_c[0].shape[1] != lc[0].in_features or (la[0].in_features & 3):
            return False
        side, wstream, _ = self.side_streams(grad_mu.device)
        ga, gc = grad_mu, grad_value
        last = len(la) - 1
        have_bias = False  # bias gradients of layer i already written (by the K9 pass below the narrow head layers)
        for i in range(last, -1, -1):
            with wstream.forked():  # ordered after everything enqueued so far on this stream (both g of this layer)
                done = None
                if i < last:
                    done = ops.wgrad_pair(ga, acts_a[i], la[i].weight.grad, None if have_bias else la[i].bias.grad, gc, acts_c[i], lc[i].weight.grad,
                                          None if have_bias else lc[i].bias.grad)
                if done is None:  # the narrow head layers (CUDA-core kernels) and anything the pair launch does not take
                    for g, x, lin in ((ga, acts_a[i], la[i]), (gc, acts_c[i], lc[i])):
                        if ops.wgrad(g, x, lin.weight.grad, None if have_bias else lin.bias.grad, zero_first=False) is None:
                            self._wgrad(g, x, lin.weight.grad, lin._wgrad_part)
                            if not have_bias:
                                ops.bias_act_bwd(g, None, lin.bias.grad)
            have_bias = False
            if i == last and from_hidden:  # K16 already produced the gradients below the heads
                ga, gc = gs_a[i], gs_c[i]
                continue
            if i == last:  # narrow head layers (n = 12 / 1): cuBLAS dgrad, then K9 = ELU backward in place + bias gradient of the layer below
                with side.forked():
                    gc = torch.mm(gc, lc[i].weight, out=gs_c[i])
                    ops.bias_act_bwd(gc, acts_c[i], lc[i - 1].bias.grad)
                ga = torch.mm(ga, la[i].weight, out=gs_a[i])
                ops.bias_act_bwd(ga, acts_a[i], la[i - 1].bias.grad)
                side.join()
                have_bias = True
                continue
            if i > 0:
                with side.forked():
                    out_c = ops.dgrad_act_bwd(gc, lc[i].weight, acts_c[i], out=gs_c[i])
                out_a = ops.dgrad_act_bwd(ga, la[i].weight, acts_a[i], out=gs_a[i])
                side.join()
                if out_a is None or out_c is None:  # shapes were checked above: only a misaligned buffer gets here
                    raise _C.LocoTouchLibraryError("K12 dgrad rejected a layer of the paired backward (misaligned activation buffer?)")
                ga, gc = out_a, out_c
        wstream.join()
        return True

    _WGRAD_SPLIT = 8

    def _wgrad_split_ok(self, batch: int, out_features: int) -> bool:
        S = self._WGRAD_SPLIT
        return batch % S == 0 and batch // S >= 512 and out_features >= 64

    def _wgrad_launch(self, g, x, lin, wstream, with_bias: bool = True) -> bool:
        """Weight gradient of ``lin`` on the trailing stream.  Returns True when the same kernel also produced the bias gradient
        (K15, TF32 mode, ``with_bias``)."""
        with wstream.forked():
            if torch.backends.cuda.matmul.allow_tf32:
                if x.shape[1] != lin.in_features:  # zero-padded input (K % 4 != 0): K15 into a padded scratch, valid columns copied out
                    cache = self.__dict__.setdefault("_pad_bufs", {})
                    key = ("dw", lin.out_features, x.shape[1], str(x.device), id(lin))
                    dwp = cache.get(key)
                    if dwp is None:
                        dwp = cache[key] = torch.empty(lin.out_features, x.shape[1], device=x.device)
                    if ops.wgrad(g, x, dwp, lin.bias.grad if with_bias else None, zero_first=True) is not None:
                        lin.weight.grad.copy_(dwp[:, :lin.in_features])
                        return with_bias
                    x = x[:, :lin.in_features].contiguous()
                elif ops.wgrad(g, x, lin.weight.grad, lin.bias.grad if with_bias else None, zero_first=False) is not None:
                    return with_bias  # K15: one tcgen05 kernel, in-kernel split-K, dW and db
            self._wgrad(g, x, lin.weight.grad, lin._wgrad_part)
        return False

    def _wgrad(self, g, x, out, part=None):
        """out[n,k] = g[B,n]^T x[B,k] through cuBLAS (fp32 parity mode / shapes K15 does not take): for the tall-skinny shapes of a
        PPO mini-batch an explicit 8-way split through one batched GEMM + a sum (cuBLAS' own split-K choice leaves most SMs idle)."""
        B = g.shape[0]
        S = self._WGRAD_SPLIT
        if self._wgrad_split_ok(B, g.shape[1]):
            part = torch.bmm(g.view(S, B // S, -1).transpose(1, 2), x.view(S, B // S, -1), out=part)
            torch.sum(part, dim=0, out=out)
        else:
            torch.mm(g.t(), x, out=out)

    def _infer(self, net, x, tag: str):
        """Inference forward of one MLP (rollout ``act`` / ``evaluate``): K12 per hidden layer -- one tcgen05 GEMM with bias + ELU in
        its epilogue instead of a cuBLAS GEMM and an ELU launch -- into persistent hidden buffers (graph-capture safe; actor and
        critic, which run on parallel streams, own separate buffers).  Falls back to the torch modules whenever autograd is
        recording, TF32 is off (fp32 parity mode), the stack is not Linear/ELU or a layer shape is not supported."""
        if (torch.is_grad_enabled() or not torch.backends.cuda.matmul.allow_tf32 or not x.is_cuda or x.dtype != torch.float32
                or x.dim() != 2 or not self._elu_stack(net)):
            return net(x)
        linears = [m for m in net if isinstance(m, nn.Linear)]
        key = (tag, x.shape[0], str(x.device))
        cache = self.__dict__.setdefault("_infer_bufs", {})
        if key not in cache:
            cache[key] = [torch.empty(x.shape[0], lin.out_features, device=x.device) for lin in linears[:-1]]
        hs = cache[key]
        h = x.contiguous()
        for i, lin in enumerate(linears):
            hidden = i < len(linears) - 1
            out = None
            if hidden and lin.out_features >= 64:
                w = lin.weight
                if i == 0 and (lin.in_features & 3):
                    h, w = self._pad4(h, "ix" + tag), self._pad4(w, "iw" + tag)
                out = ops.linear_bias_act(h, w, lin.bias, out=hs[i], elu=True)
            if out is None:
                out = torch.addmm(lin.bias, h, lin.weight.t(), out=hs[i] if hidden else None)
                if hidden:
                    F.elu_(out)
            h = out
        return h

    def _infer_buffers(self, linears, tag: str, x):
        key = (tag, x.shape[0], str(x.device))
        cache = self.__dict__.setdefault("_infer_bufs", {})
        if key not in cache:
            cache[key] = [torch.empty(x.shape[0], lin.out_features, device=x.device) for lin in linears[:-1]]
        return cache[key]

    def _infer_hidden(self, net, x, tag: str):
        """The hidden layers of one MLP through K12 (bias + ELU fused) into persistent buffers; returns the last hidden activation,
        or None when a layer is not taken by K12 (the caller then uses ``_infer``)."""
        linears = [m for m in net if isinstance(m, nn.Linear)]
        key = (tag, x.shape[0], str(x.device))
        cache = self.__dict__.setdefault("_infer_bufs", {})
        if key not in cache:
            cache[key] = [torch.empty(x.shape[0], lin.out_features, device=x.device) for lin in linears[:-1]]
        h = x.contiguous()
        for i, (lin, buf) in enumerate(zip(linears[:-1], cache[key])):
            w = lin.weight
            if i == 0 and (lin.in_features & 3):
                h, w = self._pad4(h, "ix" + tag), self._pad4(w, "iw" + tag)
            h = ops.linear_bias_act(h, w, lin.bias, out=buf, elu=True) if lin.out_features >= 64 else None
            if h is None:
                return None
        return h

    @torch.no_grad()
    def act_evaluate_fused(self, observations, critic_observations, out: dict):
        """Rollout step of BOTH networks with the heads, the action sample and the log-prob in one kernel (K3b): hidden layers
        through K12 (critic on the side stream), then ``lt_act_heads`` writes actions / log-prob / mu / sigma / values straight into
        the RolloutStorage slot ``out``.  Returns the actions, or None when this module / mode is not taken (fp32 parity mode,
        non-ELU stacks, unsupported widths, recurrent subclasses): the caller then runs ``evaluate`` + ``act``."""
        if (type(self).act is not ActorCritic.act or type(self).evaluate is not ActorCritic.evaluate or not torch.backends.cuda.matmul.allow_tf32
                or not observations.is_cuda or observations.dim() != 2 or not self.supports_fused_heads or out is None
                or any(out.get(k) is None for k in ("actions", "logp", "mu", "sigma", "values"))):
            return None
        h_a = h_c = None
        la = [m for m in self.actor if isinstance(m, nn.Linear)]
        lc = [m for m in self.critic if isinstance(m, nn.Linear)]
        if len(la) == 4 and len(lc) == 4 and observations.shape == critic_observations.shape:  # K19: both networks, one launch
            bufs = self._infer_buffers(la, "actor", observations), self._infer_buffers(lc, "critic", critic_observations)
            pre = self._mlp3_hidden([(la, observations, "iactor", bufs[0]), (lc, critic_observations, "icritic", bufs[1])], keep=False)
            if pre is not None:
                h_a, h_c = pre[0][-1], pre[1][-1]
        if h_a is None:
            side = self.side_streams(observations.device)[0]
            with side.forked():
                h_c = self._infer_hidden(self.critic, critic_observations, "critic")
            h_a = self._infer_hidden(self.actor, observations, "actor")
            side.join()
        if h_a is None or h_c is None:
            return None
        if callable(self.rng):
            eps = self.rng(out["mu"])
        else:
            eps = torch.randn_like(out["mu"]) if self.rng == "torch" else None
        head_a, head_c = self.actor[-1], self.critic[-1]
        actions, logp, mu, values = ops.act_heads(h_a, h_c, head_a.weight, head_a.bias, head_c.weight, head_c.bias, self._std_vector().detach().contiguous(),
                                                  eps=eps, actions=out["actions"], logp=out["logp"], mu=out["mu"], sigma_rows=out["sigma"], values=out["values"],
                                                  seed=self.seed, offset=self._graph_slot if self._offset_base is not None else self._draws,
                                                  offset_base=self._offset_base)
        self._draws += 1
        self.distribution = _GaussianView(mu, out["sigma"])
        self._logp, self._sampled = logp, actions
        return actions

    # ---------------------------------------------------------------------------------------------- reference interface
    @staticmethod
    def init_weights(sequential, scales):
        [torch.nn.init.orthogonal_(module.weight, gain=scales[idx]) for idx, module in enumerate(mod for mod in sequential if isinstance(mod, nn.Linear))]

    def reset(self, dones=None):
        pass

    def forward(self):
        raise NotImplementedError

    def _std_vector(self):
        if self.noise_std_type == "scalar":
            return self.std
        return torch.exp(self.log_std)

    @property
    def action_mean(self):
        return self.distribution.mean

    @property
    def action_std(self):
        return self.distribution.stddev

    @property
    def entropy(self):
        return self.distribution.entropy().sum(dim=-1)

    def update_distribution(self, observations):
        mean = self.actor(observations)
        self.distribution = _GaussianView(mean, self._std_vector().expand_as(mean))
        self._logp = self._sampled = None

    def act(self, observations, out=None, **kwargs):
        """Samples actions.  ``out`` = dict(actions, logp, mu, sigma) of pre-allocated rows (e.g. a RolloutStorage slot)."""
        mean = self._infer(self.actor, observations, "actor")
        std = self._std_vector()
        if mean.requires_grad:  # training-time call (PPO.update in the reference): distribution only
            self.distribution = _GaussianView(mean, std.expand_as(mean))
            self._logp = self._sampled = None
            return self.distribution.sample()
        out = out or {}
        if callable(self.rng):
            eps = self.rng(mean)  # explicit standard-normal draws (parity tests)
        else:
            eps = torch.randn_like(mean) if self.rng == "torch" else None
        sigma_rows = out.get("sigma")
        if sigma_rows is None:
            sigma_rows = torch.empty_like(mean)
        mean = mean.contiguous()
        actions, logp = ops.act_sample(mean, std.detach().contiguous(), eps, out.get("actions"), out.get("logp"), out.get("mu"), sigma_rows,
                                       seed=self.seed, offset=self._graph_slot if self._offset_base is not None else self._draws,
                                       offset_base=self._offset_base)
        self._draws += 1
        self.distribution = _GaussianView(out.get("mu", mean) if out.get("mu") is not None else mean, sigma_rows)
        self._logp, self._sampled = logp, actions
        return actions

    def get_actions_log_prob(self, actions):
        if self._sampled is not None and actions.data_ptr() == self._sampled.data_ptr():
            return self._logp  # computed by the same kernel that drew the sample
        return self.distribution.log_prob(actions).sum(dim=-1)

    def act_inference(self, observations):
        return self._infer(self.actor, observations, "actor")

    def evaluate(self, critic_observations, **kwargs):
        return self._infer(self.critic, critic_observations, "critic")

    def reset_init_std(self):
        if self.noise_std_type == "scalar":
            self.std.data.fill_(self.init_noise_std)
        elif self.noise_std_type == "log":
            self.log_std.data.fill_(math.log(self.init_noise_std))
        else:
            raise ValueError(f"Unknown standard deviation type: {self.noise_std_type}. Should be 'scalar' or 'log'")

    def get_actor_critic_obs_from_obs_dict(self, obs_dict):
        actor_obs = obs_dict["policy"]
        critic_obs = obs_dict.get("critic", actor_obs)
        return actor_obs, critic_obs
