"""Student with the reference's interface (reference locotouch/distill/student.py:12-187): CNN2dHead pre-encoder -> GRU+MLP
encoder -> MLP backbone, behaviour cloning against the teacher with a masked MSE.

Same constructor, ``forward`` / ``encoder_forward`` / ``backbone_forward`` / ``train_on_data`` / ``save_model`` /
``load_checkpoint`` / ``extract_input_and_forward`` / ``reset`` / ``get_hidden_states`` and ``state_dict`` keys
(``pre_encoder.conv.conv.*``, ``student_encoder.memory.rnn.*``, ...).  Below the interface: the loss and its gradient are the
K8 kernel pair (``masked_mse_loss``), the optimiser step is the fused AdamW kernel (K7) over one flat parameter buffer, and
the two ``.item()`` syncs per batch of student.py:143-151 are replaced by device-side accumulation read once per epoch.
"""
from __future__ import annotations

import os
import warnings

import torch
import torch.nn as nn

from .. import ops
from ..loco_rl.models import generate_model
from .batch import masked_mse_loss


class Student(nn.Module):
    def __init__(self, cfg, proprioception_dim: int, tactile_signal_dim: int, action_dim: int, teacher_policy_inference=None,
                 teacher_encoder_inference=None, teacher_backbone_weights=None, logger=None):
        super().__init__()
        self.cfg, self.logger = cfg, logger
        self.device, self.log_dir = cfg.device, cfg.log_dir
        self.proprioception_dim, self.tactile_signal_dim, self.action_dim = proprioception_dim, tactile_signal_dim, action_dim
        self.tactile_signal_img_shape = cfg.pre_encoder.img_shape
        self.tactile_embedding_dim = cfg.tactile_encoder.embedding_dim
        pre_type = cfg.pre_encoder.model_type
        self.use_pre_encoder = True if "CNN" in pre_type else (cfg.pre_encoder.hidden_dims is not None)
        if self.use_pre_encoder:
            self.pre_encoder = generate_model(tactile_signal_dim, cfg.pre_encoder.embedding_dim, cfg.pre_encoder).to(self.device)
        enc_in = tactile_signal_dim if not self.use_pre_encoder else cfg.pre_encoder.embedding_dim
        self.student_encoder = generate_model(enc_in, self.tactile_embedding_dim, cfg.tactile_encoder).to(self.device)
        self.student_backbone = generate_model(proprioception_dim + self.tactile_embedding_dim, action_dim, cfg.student_policy).to(self.device)
        self.MonolithicDistillation = cfg.distillation_type == "Monolithic"
        self.RMA_distillation = not self.MonolithicDistillation
        if self.RMA_distillation:
            raise NotImplementedError("RMA distillation is not used by the LocoTouch student cfg (SURVEY.md 8a18: Monolithic)")
        self.teacher_policy_inference = teacher_policy_inference
        self.teacher_encoder_inference = teacher_encoder_inference
        self.teacher_backbone_weights = teacher_backbone_weights
        if teacher_backbone_weights is not None:
            self.student_backbone.model.load_state_dict(teacher_backbone_weights)
            for p in self.student_backbone.parameters():
                p.requires_grad = False
        self.max_iterations = cfg.num_iterations
        self.initial_epoches, self.incremental_epoches, self.final_epoches = cfg.initial_epoches, cfg.incremental_epoches, cfg.final_epoches
        self.batch_steps = cfg.batch_steps
        self._distill_lr = cfg.distill_lr
        self.clip_actions, self.clip_range, self.action_scale_within_env = cfg.clip_actions, cfg.clip_range, cfg.action_scale_within_env
        self._flat = None

    # ------------------------------------------------------------------------------------------------ fused AdamW state
    def _flatten(self):
        params = [p for p in self.parameters() if p.requires_grad]
        # cuDNN takes RNN weights in place only when they sit at the START of their storage in its own layout (ATen's
        # try_get_weight_buf builds the candidate buffer from storage offset 0); otherwise every call copies them into a fresh
        # buffer first.  So the GRU's tensors (weight_ih, weight_hh, bias_ih, bias_hh: cuDNN's order for one layer) lead the flat buffer.
        rnn_first = [p for m in self.modules() if isinstance(m, nn.RNNBase) for p in m._flat_weights if p is not None and p.requires_grad]
        seen = {id(p) for p in rnn_first}
        params = rnn_first + [p for p in params if id(p) not in seen]
        dev = params[0].device
        if self._flat is not None and self._flat["params"].device == dev and all(p.data_ptr() == q for p, q in zip(params, self._flat["ptrs"])):
            return self._flat
        total = sum((p.numel() + 3) // 4 * 4 for p in params)
        flat, grads = torch.zeros(total, device=dev), torch.zeros(total, device=dev)
        off = 0
        for p in params:
            n = p.numel()
            flat[off:off + n].copy_(p.data.flatten())
            p.data = flat[off:off + n].view(p.shape)
            p.grad = grads[off:off + n].view(p.shape)
            off += (n + 3) // 4 * 4
        # NOTE: nn.GRU.flatten_parameters() must NOT be called afterwards: it would re-home the recurrent weights into a
        # cuDNN-owned buffer and detach them from the flat buffer the fused AdamW kernel updates (it is not needed either:
        # the tensors already lie the way cuDNN wants them).
        self._flat = dict(params=flat, grads=grads, m=torch.zeros_like(flat), v=torch.zeros_like(flat), step=torch.zeros(1, device=dev),
                          lr=torch.full((1,), self._distill_lr, device=dev), ptrs=[p.data_ptr() for p in params])
        return self._flat

    def optimizer_step(self):
        """torch.optim.AdamW(lr=distill_lr) semantics (betas 0.9/0.999, eps 1e-8, weight_decay 1e-2), no gradient clipping."""
        f = self._flatten()
        ops.clip_adam(f["params"], f["grads"], f["m"], f["v"], f["lr"], f["step"], max_grad_norm=None, weight_decay=1e-2)

    # --------------------------------------------------------------------------------------------------------- forward
    def encoder_forward(self, tactile_signal, hidden_states=None):
        if self.use_pre_encoder:
            shape = tactile_signal.shape
            if len(shape) <= 3:
                tactile_signal = tactile_signal.reshape(*shape[:-1], *self.tactile_signal_img_shape)
                shape = tactile_signal.shape
            flat = tactile_signal.reshape(-1, *shape[-3:])
            tactile_signal = self.pre_encoder(flat).reshape(*shape[:-3], -1)
        return self.student_encoder(tactile_signal, hidden_states)

    def backbone_forward(self, proprioception, tactile_embedding):
        return self.student_backbone(torch.cat((proprioception, tactile_embedding), dim=-1))

    def forward(self, proprioception, tactile_signal, hidden_states=None):
        return self.backbone_forward(proprioception, self.encoder_forward(tactile_signal, hidden_states))

    # -------------------------------------------------------------------------------------------------------- training
    def train_on_batch(self, batch):
        """One behaviour-cloning step on a padded batch dict (reference student.py:121-151).  Returns the device tensor
        ``(loss, mae, count, 0)`` of the K8 kernel; nothing is synchronised."""
        f = self._flatten()
        f["grads"].zero_()
        prop, teach_obs, tac, masks = batch["proprioceptions"], batch["teacher_encoder_obses"], batch["tactile_signals"], batch["masks"]
        student_actions = self.forward(prop, tac)
        with torch.no_grad():
            teacher_actions = self.teacher_policy_inference(torch.cat((prop, teach_obs), dim=-1))
        loss = masked_mse_loss(student_actions, teacher_actions, masks)
        loss.backward()
        self.optimizer_step()
        return loss.detach()

    def train_on_data(self, replay_buffer, num_iter: int):
        self.train()
        epochs = self.initial_epoches + self.incremental_epoches * num_iter
        epochs += self.final_epoches if num_iter == self.max_iterations - 1 else 0
        batch_trajs = int(self.batch_steps / (replay_buffer.num_steps / replay_buffer.num_trajs)) + 1
        mean = 0.0
        for _ in range(epochs):
            losses = [self.train_on_batch(b) for b in replay_buffer.to_recurrent_generator(batch_size=batch_trajs)]
            mean = float(torch.stack(losses).mean().item())  # one device->host read per epoch
            if self.logger is not None and getattr(self.cfg, "logger", None) == "wandb":
                self.logger.log({"train/Action MSE": mean})
        print(f"[Distillation iteration {num_iter}] Action MSE: {mean}")
        self.save_model(num_iter)

    def save_model(self, iteration):
        torch.save(self.state_dict(), os.path.join(self.log_dir, f"model_{iteration}.pt"))

    def load_checkpoint(self, model_path):
        self.load_state_dict(torch.load(model_path, map_location=self.device))

    def extract_input_and_forward(self, obs):
        if "tactile_packed" in obs and self.use_pre_encoder and hasattr(self.pre_encoder, "forward_packed") and not torch.is_grad_enabled():
            # K2's ballot-packed bitmap straight into K17: the 442-float image is never read
            emb = self.pre_encoder.forward_packed(obs["tactile_packed"])
            return self.backbone_forward(obs["policy"][:, :self.proprioception_dim], self.student_encoder(emb, None))
        return self.forward(obs["policy"][:, :self.proprioception_dim], obs["tactile"])

    def reset(self, dones=None):
        if self.use_pre_encoder:
            self.pre_encoder.reset(dones)
        self.student_encoder.reset(dones)
        self.student_backbone.reset(dones)

    def get_hidden_states(self):
        if hasattr(self.student_encoder, "get_hidden_states"):
            return self.student_encoder.get_hidden_states()
