"""Recurrent policies with the reference's interfaces:

* ``Memory``                          -- reference loco_rl/loco_rl/modules/actor_critic_recurrent.py:68-95
* ``ActorCriticRecurrent``            -- reference actor_critic_recurrent.py:13-65
* ``ActorCriticRNNEncoder``           -- reference actor_critic_rnn_encoder.py:11-153
* ``ActorCriticPreEncoderRNNEncoder`` -- reference actor_critic_pre_encoder_rnn_encoder.py:9-181
  (cfg blocks locotouch/config/locotouch/agents/rsl_rl_ppo_cfg.py:147,181-226)

Same constructors, attribute names and ``state_dict`` keys (``memory_a.rnn.*``, ``actor_encoder.model.*``,
``actor_pre_encoder.model.*``, ...).  Below the interface they sit on this package's ``ActorCritic``: rollout ``act`` samples through
the fused act epilogue (K3) straight into the RolloutStorage slot, the update's batch mode un-pads the GRU outputs and the flat
observation parts with the trajectory kernels (K10), the loss / optimiser step are K6 / K7.  The GRU itself is cuDNN through
``nn.GRU``, as in the reference."""
from __future__ import annotations

import warnings

import torch
import torch.nn as nn

from ..models import MLP
from ..utils import unpad_trajectories
from .actor_critic import ActorCritic


# The GRU tensors are views into the policy's flat parameter buffer (one Adam launch, one gradient exchange); with two GRUs at most
# one could lead that buffer, so cuDNN packs these (small: 3 x hidden x (in + hidden) floats) weights per call and says so every time.
warnings.filterwarnings("ignore", message="RNN module weights are not part of single contiguous chunk")


class Memory(nn.Module):
    def __init__(self, input_size, type="lstm", num_layers=1, hidden_size=256):
        super().__init__()
        rnn_cls = nn.GRU if type.lower() == "gru" else nn.LSTM
        self.rnn = rnn_cls(input_size=input_size, hidden_size=hidden_size, num_layers=num_layers)
        self.hidden_states = None

    def forward(self, input, masks=None, hidden_states=None):
        if masks is not None:  # batch mode (policy update): the saved hidden states of every trajectory's first step
            if hidden_states is None:
                raise ValueError("Hidden states not passed to memory module during policy update")
            out, _ = self.rnn(input, hidden_states)
            out = unpad_trajectories(out, masks)
        else:  # collection: carry the hidden states of the last step
            out, self.hidden_states = self.rnn(input.unsqueeze(0), self.hidden_states)
        return out

    def reset(self, dones=None):
        if self.hidden_states is None:
            return
        states = self.hidden_states if isinstance(self.hidden_states, tuple) else (self.hidden_states,)
        if dones is None:
            return  # reference: `hidden_state[..., None == 1, :] = 0.0` indexes with False, i.e. clears nothing; preserved
        for state in states:  # [layers, N, hidden]: clear the envs that finished (reference: `hidden_state[..., dones == 1, :] = 0.0`)
            state.masked_fill_((dones == 1).view(1, -1, 1), 0.0)


def _obs_split(obs_dim, flatten_end_idx, encoder_start_idx):
    enc = abs(encoder_start_idx) if encoder_start_idx < 0 else (obs_dim - encoder_start_idx)
    flat = flatten_end_idx if flatten_end_idx > 0 else (obs_dim - abs(flatten_end_idx))
    return flat, enc


class _EncoderPolicy(ActorCritic):
    """Shared mechanics of the two encoder policies: [flat part | encoder part] observations, a (pre-encoder ->) GRU -> encoder MLP
    branch whose embedding is concatenated to the flat part in front of the ActorCritic backbone."""

    is_recurrent = True

    def reset(self, dones=None):
        super().reset(dones)
        self.memory_a.reset(dones)
        if self.critic_with_encoder:
            self.memory_c.reset(dones)

    def _embed(self, which: str, encoder_obs, masks=None, hidden_states=None):
        pre = getattr(self, f"{which}_pre_encoder", None)
        x = pre(encoder_obs) if pre is not None else encoder_obs
        mem = self.memory_a if which == "actor" else self.memory_c
        x = mem(x, masks, hidden_states).squeeze(0)
        return getattr(self, f"{which}_encoder")(x)

    def act(self, obs, masks=None, hidden_states=None, out=None):
        flatten_obs = obs[..., :self.actor_flatten_obs_dim]
        if masks is not None:
            flatten_obs = unpad_trajectories(flatten_obs, masks)
        embedding = self._embed("actor", obs[..., -self.actor_encoder_obs_dim:], masks, hidden_states)
        return super().act(torch.cat([flatten_obs, embedding], dim=-1), out=out)

    def act_encoder_inference(self, encoder_obs):
        return self._embed("actor", encoder_obs)

    def act_backbone_inference(self, flatten_obs, embedding):
        return super().act_inference(torch.cat([flatten_obs, embedding], dim=-1))

    def evaluate(self, obs, masks=None, hidden_states=None):
        if self.critic_with_encoder:
            flatten_obs = obs[..., :self.critic_flatten_obs_dim]
            if masks is not None:
                flatten_obs = unpad_trajectories(flatten_obs, masks)
            embedding = self._embed("critic", obs[..., -self.critic_encoder_obs_dim:], masks, hidden_states)
            input_c = torch.cat([flatten_obs, embedding], dim=-1)
        else:
            input_c = unpad_trajectories(obs, masks) if (masks is not None and self._unpad_plain_critic) else obs
        return super().evaluate(input_c)

    def get_hidden_states(self):
        return self.memory_a.hidden_states, (self.memory_c.hidden_states if self.critic_with_encoder else None)


class ActorCriticRNNEncoder(_EncoderPolicy):
    _unpad_plain_critic = False  # reference actor_critic_rnn_encoder.py:146-147 hands the padded observations to the critic as they are

    def __init__(self, actor_obs_dim, critic_obs_dim, num_actions, actor_flatten_obs_end_idx, actor_encoder_obs_start_idx, actor_encoder_hidden_dims,
                 actor_encoder_embedding_dim, actor_hidden_dims, critic_flatten_obs_end_idx, critic_encoder_obs_start_idx, critic_encoder_hidden_dims,
                 critic_encoder_embedding_dim, critic_hidden_dims, encoder_rnn_type="gru", encoder_rnn_hidden_size=256, encoder_rnn_num_layers=1,
                 encoder_activation="elu", encoder_final_activation=None, activation="elu", init_noise_std=1.0, **kwargs):
        if kwargs:
            print("ActorCriticEncoder.__init__ got unexpected arguments, which will be ignored: " + str(kwargs.keys()))
        flat_a, enc_a = _obs_split(actor_obs_dim, actor_flatten_obs_end_idx, actor_encoder_obs_start_idx)
        with_c = critic_encoder_hidden_dims is not None
        if with_c:
            assert critic_flatten_obs_end_idx is not None, "Critic flatten obs end index is required"
            assert critic_encoder_obs_start_idx is not None, "Critic encoder obs start index is required"
            assert critic_encoder_embedding_dim is not None, "Critic encoder embedding dim is required"
            flat_c, enc_c = _obs_split(critic_obs_dim, critic_flatten_obs_end_idx, critic_encoder_obs_start_idx)
        super().__init__(num_actor_obs=flat_a + actor_encoder_embedding_dim,
                         num_critic_obs=(flat_c + critic_encoder_embedding_dim) if with_c else critic_obs_dim, num_actions=num_actions,
                         actor_hidden_dims=actor_hidden_dims, critic_hidden_dims=critic_hidden_dims, activation=activation, init_noise_std=init_noise_std)
        self.actor_encoder_obs_dim, self.actor_flatten_obs_dim, self.critic_with_encoder = enc_a, flat_a, with_c
        self.memory_a = Memory(input_size=enc_a, type=encoder_rnn_type, num_layers=encoder_rnn_num_layers, hidden_size=encoder_rnn_hidden_size)
        self.actor_encoder = MLP(encoder_rnn_hidden_size, actor_encoder_hidden_dims, actor_encoder_embedding_dim, activation=encoder_activation,
                                 final_layer_activation=encoder_final_activation)
        if with_c:
            self.critic_encoder_obs_dim, self.critic_flatten_obs_dim = enc_c, flat_c
            self.memory_c = Memory(input_size=enc_c, type=encoder_rnn_type, num_layers=encoder_rnn_num_layers, hidden_size=encoder_rnn_hidden_size)
            self.critic_encoder = MLP(encoder_rnn_hidden_size, critic_encoder_hidden_dims, critic_encoder_embedding_dim, activation=encoder_activation,
                                      final_layer_activation=encoder_final_activation)

    def act_inference(self, obs):
        embedding = self._embed("actor", obs[..., -self.actor_encoder_obs_dim:])
        return ActorCritic.act_inference(self, torch.cat([obs[..., :self.actor_flatten_obs_dim], embedding], dim=-1))


class ActorCriticPreEncoderRNNEncoder(_EncoderPolicy):
    _unpad_plain_critic = True  # reference actor_critic_pre_encoder_rnn_encoder.py:172-174

    def __init__(self, actor_obs_dim, critic_obs_dim, num_actions, actor_flatten_obs_end_idx, actor_encoder_obs_start_idx,
                 actor_pre_encoder_hidden_dims, actor_pre_encoder_embedding_dim, actor_encoder_hidden_dims, actor_encoder_embedding_dim,
                 actor_hidden_dims, critic_flatten_obs_end_idx, critic_encoder_obs_start_idx, critic_pre_encoder_hidden_dims,
                 critic_pre_encoder_embedding_dim, critic_encoder_hidden_dims, critic_encoder_embedding_dim, critic_hidden_dims,
                 encoder_rnn_type="gru", encoder_rnn_hidden_size=256, encoder_rnn_num_layers=1, pre_encoder_activation="elu",
                 pre_encoder_final_activation=None, encoder_activation="elu", encoder_final_activation=None, activation="elu",
                 init_noise_std=1.0, **kwargs):
        if kwargs:
            print("ActorCriticEncoder.__init__ got unexpected arguments, which will be ignored: " + str(kwargs.keys()))
        flat_a, enc_a = _obs_split(actor_obs_dim, actor_flatten_obs_end_idx, actor_encoder_obs_start_idx)
        with_c = not (critic_pre_encoder_hidden_dims is None or critic_encoder_hidden_dims is None)
        if with_c:
            assert critic_flatten_obs_end_idx is not None, "Critic flatten obs end index is required"
            assert critic_encoder_obs_start_idx is not None, "Critic encoder obs start index is required"
            assert critic_encoder_embedding_dim is not None, "Critic encoder embedding dim is required"
            flat_c, enc_c = _obs_split(critic_obs_dim, critic_flatten_obs_end_idx, critic_encoder_obs_start_idx)
        super().__init__(num_actor_obs=flat_a + actor_encoder_embedding_dim,
                         num_critic_obs=(flat_c + critic_encoder_embedding_dim) if with_c else critic_obs_dim, num_actions=num_actions,
                         actor_hidden_dims=actor_hidden_dims, critic_hidden_dims=critic_hidden_dims, activation=activation, init_noise_std=init_noise_std)
        self.actor_encoder_obs_dim, self.actor_flatten_obs_dim, self.critic_with_encoder = enc_a, flat_a, with_c
        self.actor_pre_encoder = MLP(enc_a, actor_pre_encoder_hidden_dims, actor_pre_encoder_embedding_dim, activation=pre_encoder_activation,
                                     final_layer_activation=pre_encoder_final_activation)
        self.memory_a = Memory(input_size=actor_pre_encoder_embedding_dim, type=encoder_rnn_type, num_layers=encoder_rnn_num_layers,
                               hidden_size=encoder_rnn_hidden_size)
        self.actor_encoder = MLP(encoder_rnn_hidden_size, actor_encoder_hidden_dims, actor_encoder_embedding_dim, activation=encoder_activation,
                                 final_layer_activation=encoder_final_activation)
        if with_c:
            self.critic_encoder_obs_dim, self.critic_flatten_obs_dim = enc_c, flat_c
            self.critic_pre_encoder = MLP(enc_c, critic_pre_encoder_hidden_dims, critic_pre_encoder_embedding_dim, activation=pre_encoder_activation,
                                          final_layer_activation=pre_encoder_final_activation)
            self.memory_c = Memory(input_size=critic_pre_encoder_embedding_dim, type=encoder_rnn_type, num_layers=encoder_rnn_num_layers,
                                   hidden_size=encoder_rnn_hidden_size)
            self.critic_encoder = MLP(encoder_rnn_hidden_size, critic_encoder_hidden_dims, critic_encoder_embedding_dim, activation=encoder_activation,
                                      final_layer_activation=encoder_final_activation)

    def act_inference(self, obs):
        # reference actor_critic_pre_encoder_rnn_encoder.py:146-153 samples here too (`super().act`); preserved
        embedding = self._embed("actor", obs[..., -self.actor_encoder_obs_dim:])
        return ActorCritic.act(self, torch.cat([obs[..., :self.actor_flatten_obs_dim], embedding], dim=-1))


class ActorCriticRecurrent(ActorCritic):
    is_recurrent = True

    def __init__(self, num_actor_obs, num_critic_obs, num_actions, actor_hidden_dims=[256, 256, 256], critic_hidden_dims=[256, 256, 256],
                 activation="elu", rnn_type="lstm", rnn_hidden_size=256, rnn_num_layers=1, init_noise_std=1.0, **kwargs):
        if kwargs:
            print("ActorCriticRecurrent.__init__ got unexpected arguments, which will be ignored: " + str(kwargs.keys()))
        super().__init__(num_actor_obs=rnn_hidden_size, num_critic_obs=rnn_hidden_size, num_actions=num_actions, actor_hidden_dims=actor_hidden_dims,
                         critic_hidden_dims=critic_hidden_dims, activation=activation, init_noise_std=init_noise_std)
        self.memory_a = Memory(num_actor_obs, type=rnn_type, num_layers=rnn_num_layers, hidden_size=rnn_hidden_size)
        self.memory_c = Memory(num_critic_obs, type=rnn_type, num_layers=rnn_num_layers, hidden_size=rnn_hidden_size)

    def reset(self, dones=None):
        self.memory_a.reset(dones)
        self.memory_c.reset(dones)

    def act(self, observations, masks=None, hidden_states=None, out=None):
        return super().act(self.memory_a(observations, masks, hidden_states).squeeze(0), out=out)

    def act_inference(self, observations):
        return super().act_inference(self.memory_a(observations).squeeze(0))

    def evaluate(self, critic_observations, masks=None, hidden_states=None):
        return super().evaluate(self.memory_c(critic_observations, masks, hidden_states).squeeze(0))

    def get_hidden_states(self):
        return self.memory_a.hidden_states, self.memory_c.hidden_states
