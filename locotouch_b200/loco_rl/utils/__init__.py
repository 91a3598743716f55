import torch


def resolve_nn_activation(act_name: str) -> torch.nn.Module:
    """reference loco_rl/loco_rl/utils/utils.py:15-34 ("crelu" -> CELU at this call site)."""
    table = {
        "elu": torch.nn.ELU, "selu": torch.nn.SELU, "relu": torch.nn.ReLU, "crelu": torch.nn.CELU, "lrelu": torch.nn.LeakyReLU,
        "tanh": torch.nn.Tanh, "sigmoid": torch.nn.Sigmoid, "identity": torch.nn.Identity,
    }
    if act_name not in table:
        raise ValueError(f"Invalid activation function '{act_name}'.")
    return table[act_name]()


def split_and_pad_trajectories(tensor, dones):
    """Drop-in for reference loco_rl/loco_rl/utils/utils.py:37-73: cuts ``tensor`` [T, N, ...] at the done flags (the last step
    always ends a trajectory), lists the trajectories env by env in time order and pads them with zeros to T steps.
    Returns (padded [T, M, D], masks [T, M] bool).  Two launches (index + output-driven copy) instead of clone / nonzero /
    tolist / split into M views / pad_sequence."""
    from ... import ops

    return ops.TrajectoryIndex(dones).split_and_pad(tensor)


class _UnpadFn(torch.autograd.Function):
    """Differentiable face of the un-pad kernel: the gradient of a gather of the valid rows is the zero-padded scatter of the
    incoming gradient -- the split-and-pad kernel over the same trajectory index."""

    @staticmethod
    def forward(ctx, trajectories, masks):
        from ... import ops

        ctx.index = ops.TrajectoryIndex.from_masks(masks)
        return ctx.index.unpad(trajectories)

    @staticmethod
    def backward(ctx, grad):
        return ctx.index.split_and_pad(grad, want_masks=False)[0], None


def unpad_trajectories(trajectories, masks):
    """Inverse of split_and_pad_trajectories (reference utils.py:76-83): [T, M, D] + masks [T, M] -> [T, N, D].  Differentiable
    (the recurrent PPO update back-propagates through it into the GRU outputs)."""
    from ... import ops

    if torch.is_grad_enabled() and trajectories.requires_grad:
        return _UnpadFn.apply(trajectories, masks)
    return ops.TrajectoryIndex.from_masks(masks).unpad(trajectories)
