"""MLP container with the reference's parameter naming (metabo/policies/mlp.py:25-61:
`fc.{i}.weight/bias`), so the shipped iclr2020 `weights_N` state_dicts load unchanged."""
from torch import nn


class MLP(nn.Module):
    def __init__(self, d_in, d_out, arch_spec, f_act=None):
        super().__init__()
        self.arch_spec = list(arch_spec)
        self.f_act = f_act
        self.is_linear = (self.arch_spec == [])
        if not self.is_linear:
            assert f_act is not None
        dims = [d_in] + self.arch_spec + [d_out]
        self.fc = nn.ModuleList([nn.Linear(dims[i], dims[i + 1]) for i in range(len(dims) - 1)])

    def forward(self, X):
        """autograd path (torch ops on the tensor's device) -- used by the PPO update only"""
        Y = X
        for layer in self.fc[:-1]:
            Y = self.f_act(layer(Y))
        return self.fc[-1](Y)
