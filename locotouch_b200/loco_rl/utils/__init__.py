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
