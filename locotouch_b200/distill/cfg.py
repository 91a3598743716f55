"""Distillation configuration (reference locotouch/config/locotouch/agents/distillation_cfg.py:5-90) as plain classes."""
from __future__ import annotations

from ..loco_rl.models import ModelCfg


class PreEncoderCfg(ModelCfg):
    def __init__(self, **kw):
        super().__init__(**{"model_type": "MLP", "hidden_dims": None, "embedding_dim": None, **kw})


class TactileEncoderCfg(ModelCfg):
    def __init__(self, **kw):
        super().__init__(**{"model_type": "MLP", "hidden_dims": [256, 128, 64], "embedding_dim": 64, **kw})


class StudentPolicyCfg(ModelCfg):
    pass


class DistillationCfg:
    def __init__(self, **kw):
        self.distillation_type = "Monolithic"
        self.pre_encoder = PreEncoderCfg()
        self.tactile_encoder = TactileEncoderCfg()
        self.student_policy = StudentPolicyCfg()
        self.device = "cuda:0"
        self.log_root_path, self.experiment_name = "logs/distillation", "object"
        self.log_dir, self.log_dir_distill, self.checkpoint_distill = "specify_log_dir", "specify_log_dir_distill", "specify_checkpoint_distill"
        self.logger, self.wandb_project = "wandb", "Transport_Distillation"
        self.num_iterations, self.bc_data_steps, self.dagger_data_steps = 8, 400000, 200000
        self.initial_epoches, self.incremental_epoches, self.final_epoches = 2000, 500, 0
        self.batch_steps, self.distill_lr, self.evaluation_trajs_num = 20000, 5.0e-4, 2000
        self.clip_actions, self.clip_range, self.action_scale_within_env = False, 100.0, 0.25
        self.min_delay, self.max_delay = 1, 2
        for k, v in kw.items():
            setattr(self, k, v)


class DistillationRandCylinderCNNRNNMonCfg(DistillationCfg):
    """RandCylinderTransportStudent_SingleBinaryTac_CNNRNN_Mon (reference distillation_cfg.py:78-85)."""

    def __init__(self, **kw):
        super().__init__(**kw)
        self.pre_encoder.model_type = "CNN2dHead"
        self.pre_encoder.embedding_dim = 64
        self.tactile_encoder.model_type = "RNN"
        self.tactile_encoder.rnn_hidden_size = 512
        self.experiment_name = "rand_cylinder"
