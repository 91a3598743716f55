"""Building blocks of the student network with the reference's names, constructor arguments and ``state_dict`` keys
(reference loco_rl/loco_rl/models/{mlp,cnn_2d,rnn,memory_module,model_cfg,model_generation,activation}.py).
Training runs the layers through torch (cuBLAS / cuDNN, as in the reference) -- autograd needs their intermediates; the
inference path of the LocoTouch student's tactile pre-encoder (``CNN2dHead`` under ``torch.no_grad()``, i.e. the student acting in
the DAgger collection / evaluation) is ONE hand-written kernel (K17, ``lt_student_cnn_forward``), optionally fed with the
ballot-packed taxel bitmap K2 emits.  The other hand-written parts of the student batch are the padded-batch assembly and the
masked loss (K8) and the fused AdamW step (K7)."""
from __future__ import annotations

import copy

import torch
import torch.nn as nn

from ... import ops


def get_activation(act_name):
    """reference models/activation.py:3-19 ("crelu" -> ReLU at this call site)."""
    table = {"elu": nn.ELU, "selu": nn.SELU, "relu": nn.ReLU, "crelu": nn.ReLU, "lrelu": nn.LeakyReLU, "tanh": nn.Tanh, "sigmoid": nn.Sigmoid}
    if act_name not in table:
        raise ValueError(f"Invalid activation function: {act_name}")
    return table[act_name]()


class MLP(nn.Module):
    """reference models/mlp.py:4-25 (keys ``model.{0,2,...}``)."""

    def __init__(self, input_dim, hidden_dims, output_dim, activation="elu", final_layer_activation=None):
        super().__init__()
        layers, prev = [], input_dim
        for h in hidden_dims or []:
            layers += [nn.Linear(prev, h), get_activation(activation)]
            prev = h
        layers.append(nn.Linear(prev, output_dim))
        if final_layer_activation is not None:
            layers.append(get_activation(final_layer_activation))
        self.model = nn.Sequential(*layers)

    def forward(self, x):
        return self.model(x)

    def reset(self, dones=None):
        pass


def conv2d_output_shape(h, w, kernel_size=1, stride=1, padding=0, dilation=1):
    kh, kw = kernel_size if isinstance(kernel_size, tuple) else (kernel_size,) * 2
    sh, sw = stride if isinstance(stride, tuple) else (stride,) * 2
    ph, pw = padding if isinstance(padding, tuple) else (padding,) * 2
    h = (h + 2 * ph - dilation * (kh - 1) - 1) // sh + 1
    w = (w + 2 * pw - dilation * (kw - 1) - 1) // sw + 1
    return h, w


class CNN2d(nn.Module):
    """reference models/cnn_2d.py:16-72 (keys ``conv.{i}``; with ``use_maxpool`` the convs run at stride 1 and a MaxPool2d
    with the configured stride follows every conv whose stride is > 1)."""

    def __init__(self, in_channels=2, channels=(2, 4, 8), kernel_sizes=(5, 4, 3), strides=(2, 1, 1), paddings=None, nonlinearity="relu",
                 use_maxpool=True, normlayer=None):
        super().__init__()
        paddings = [0] * len(channels) if paddings is None else paddings
        act = get_activation(nonlinearity)
        normlayer = getattr(nn, normlayer) if isinstance(normlayer, str) else normlayer
        assert len(channels) == len(kernel_sizes) == len(strides) == len(paddings)
        ins = [in_channels] + list(channels)[:-1]
        pool_strides = strides if use_maxpool else [1] * len(strides)
        conv_strides = [1] * len(strides) if use_maxpool else strides
        seq = []
        for ic, oc, k, s, p, ps in zip(ins, channels, kernel_sizes, conv_strides, paddings, pool_strides):
            seq.append(nn.Conv2d(in_channels=ic, out_channels=oc, kernel_size=k, stride=s, padding=p))
            if normlayer is not None:
                seq.append(normlayer(oc))
            seq.append(act)
            if ps > 1:
                seq.append(nn.MaxPool2d(ps))
        self.conv = nn.Sequential(*seq)

    def forward(self, x):
        return self.conv(x)

    def conv_out_size(self, h, w, c=None):
        for child in self.conv.children():
            if isinstance(child, (nn.Conv2d, nn.MaxPool2d)):
                h, w = conv2d_output_shape(h, w, child.kernel_size, child.stride, child.padding)
            if isinstance(child, nn.Conv2d):
                c = child.out_channels
        return h * w * c

    def reset(self, dones=None):
        pass


class CNN2dHead(nn.Module):
    """reference models/cnn_2d.py:75-131 (keys ``conv.conv.{i}``, ``head.model.{i}``)."""

    def __init__(self, image_shape, channels=(2, 4, 8), kernel_sizes=(5, 4, 3), strides=(2, 1, 1), paddings=None, hidden_sizes=None,
                 output_size=None, nonlinearity="relu", use_maxpool=False, normlayer=None):
        super().__init__()
        c, h, w = image_shape
        self._image_shape = tuple(image_shape)
        self._fused_ok = None if (nonlinearity in ("relu", "crelu") and normlayer is None and use_maxpool and tuple(strides) == (2, 1, 1)) else False
        self.conv = CNN2d(c, channels, kernel_sizes, strides, paddings, nonlinearity, use_maxpool, normlayer)
        conv_out = self.conv.conv_out_size(h, w)
        if hidden_sizes or output_size:
            self.head = MLP(conv_out, hidden_sizes, output_size, activation=nonlinearity)
            self._output_size = output_size if output_size is not None else (hidden_sizes if isinstance(hidden_sizes, int) else hidden_sizes[-1])
        else:
            self.head = lambda x: x
            self._output_size = conv_out

    def _fused_weights(self):
        """(w1, b1, w2, b2, w3, b3, wh, bh) when K17 takes this instance's geometry, else None."""
        ok = getattr(self, "_fused_ok", None)
        if ok is None:
            convs = [m for m in self.conv.conv if isinstance(m, nn.Conv2d)]
            head = getattr(self.head, "model", None)
            ok = (len(convs) == 3 and head is not None and len(head) == 1 and isinstance(head[0], nn.Linear)
                  and all(isinstance(m, (nn.Conv2d, nn.ReLU, nn.MaxPool2d)) for m in self.conv.conv)
                  and [type(m).__name__ for m in self.conv.conv] == ["Conv2d", "ReLU", "MaxPool2d", "Conv2d", "ReLU", "Conv2d", "ReLU"]
                  and self.conv.conv[2].kernel_size in (2, (2, 2)) and self.conv.conv[2].stride in (2, (2, 2))
                  and all(c.stride == (1, 1) and c.dilation == (1, 1) and c.groups == 1 for c in convs)
                  and ops.student_cnn_supported(self._image_shape, [c.out_channels for c in convs], [c.kernel_size[0] for c in convs],
                                                (2, 1, 1), [c.padding[0] for c in convs], head[0].out_features))
            self._fused_ok = ok
        if not ok:
            return None
        convs = [m for m in self.conv.conv if isinstance(m, nn.Conv2d)]
        lin = self.head.model[0]
        return (convs[0].weight, convs[0].bias, convs[1].weight, convs[1].bias, convs[2].weight, convs[2].bias, lin.weight, lin.bias)

    def forward(self, x):
        if not torch.is_grad_enabled() and x.is_cuda and x.dtype == torch.float32:
            w = self._fused_weights()
            if w is not None and w[0].is_cuda:
                return ops.student_cnn_forward(w, image=x.reshape(x.shape[0], -1).contiguous())  # K17: one kernel
        return self.head(self.conv(x).view(x.shape[0], -1))

    def forward_packed(self, packed):
        """Inference from the ballot-packed taxel bitmap [M, words] (int32; bit t % 32 of word t // 32 = taxel t, both image channels
        equal) that K2 (``lt_taxel_synth``) emits next to -- or instead of -- the 442-float observation."""
        w = self._fused_weights()
        if w is None:
            raise NotImplementedError("forward_packed needs the LocoTouch student pre-encoder geometry (K17)")
        return ops.student_cnn_forward(w, packed=packed.contiguous())

    @property
    def output_size(self):
        return self._output_size

    def reset(self, dones=None):
        pass


class Memory(nn.Module):
    """reference models/memory_module.py:3-30 (key ``rnn.*``)."""

    def __init__(self, memory_type, input_dim, hidden_size, num_layers):
        super().__init__()
        cls = nn.GRU if memory_type.lower() == "gru" else nn.LSTM
        self.rnn = cls(input_size=input_dim, hidden_size=hidden_size, num_layers=num_layers)
        self.hidden_states = None

    def forward(self, input, hidden_states=None):
        if len(input.shape) == 3:  # batch mode during training: [L, B, D]
            out, _ = self.rnn(input, hidden_states)
        else:  # collection: carry the hidden state of the last step
            out, self.hidden_states = self.rnn(input.unsqueeze(0), self.hidden_states)
            out = out.squeeze(0)
        return out

    def reset(self, dones=None):
        if self.hidden_states is not None:
            if dones is None:
                self.hidden_states = None
            else:
                for state in (self.hidden_states if isinstance(self.hidden_states, tuple) else (self.hidden_states,)):
                    state[..., dones, :] = 0.0

    def get_hidden_states(self):
        return self.hidden_states


class RNN(nn.Module):
    """reference models/rnn.py:6-22 (keys ``memory.rnn.*``, ``mlp.model.*``)."""

    def __init__(self, input_dim, hidden_dims, output_dim, activation="elu", rnn_memory_type="gru", rnn_hidden_size=256, rnn_num_layers=1):
        super().__init__()
        self.memory = Memory(rnn_memory_type, input_dim, rnn_hidden_size, rnn_num_layers)
        self.mlp = MLP(rnn_hidden_size, hidden_dims, output_dim, activation)

    def forward(self, x, hidden_states=None):
        return self.mlp(self.memory(x, hidden_states=hidden_states))

    def reset(self, dones=None):
        self.memory.reset(dones=dones)

    def get_hidden_states(self):
        return self.memory.get_hidden_states()


class ModelCfg:
    """reference models/model_cfg.py:5-25 without the IsaacLab ``configclass`` dependency."""

    model_type = "MLP"
    hidden_dims = [512, 256, 128]
    activation = "elu"
    final_layer_activation = None
    rnn_type = "gru"
    rnn_hidden_size = 256
    rnn_num_layers = 1
    img_shape = (2, 17, 13)
    cnn_channels = (24, 24, 24)
    cnn_kernel_size = (4, 3, 2)
    cnn_stride = (2, 1, 1)
    cnn_nonlinearity = "relu"
    cnn_padding = None
    cnn_use_maxpool = True
    cnn_normlayer = None

    def __init__(self, **kw):
        for k in dir(type(self)):
            if not k.startswith("_") and not callable(getattr(type(self), k)):
                setattr(self, k, copy.deepcopy(getattr(type(self), k)))
        for k, v in kw.items():
            setattr(self, k, v)


def generate_model(input_dim: int, output_dim: int, cfg):
    """reference models/model_generation.py:3-22"""
    t = cfg.model_type
    if t == "MLP":
        return MLP(input_dim, cfg.hidden_dims, output_dim, cfg.activation, cfg.final_layer_activation)
    if t == "RNN":
        return RNN(input_dim, cfg.hidden_dims, output_dim, cfg.activation, cfg.rnn_type, cfg.rnn_hidden_size, cfg.rnn_num_layers)
    if t == "CNN2d":
        return CNN2d(input_dim, cfg.cnn_channels, cfg.cnn_kernel_size, cfg.cnn_stride, cfg.cnn_padding, cfg.cnn_nonlinearity, cfg.cnn_use_maxpool, cfg.cnn_normlayer)
    if t == "CNN2dHead":
        return CNN2dHead(cfg.img_shape, cfg.cnn_channels, cfg.cnn_kernel_size, cfg.cnn_stride, cfg.cnn_padding, cfg.hidden_dims, output_dim,
                         cfg.cnn_nonlinearity, cfg.cnn_use_maxpool, cfg.cnn_normlayer)
    raise NotImplementedError(f"Model type {t} not implemented")
