"""A tiny data-parallel MLP whose gradient sync runs through gloo_b200.

This is the "model running end to end" of a collectives library: forward/backward in
torch, gradient allreduce through parallel.DataParallel (fused NVLink kernels for CUDA
tensors, TCP ring for CPU tensors), SGD step. Used by tests and examples.
"""
from __future__ import annotations

import torch
from torch import nn

from ..parallel import DataParallel


class DDPMLP(nn.Module):
    def __init__(self, d_in=64, d_hidden=256, d_out=8):
        super().__init__()
        self.net = nn.Sequential(nn.Linear(d_in, d_hidden), nn.GELU(), nn.Linear(d_hidden, d_out))

    def forward(self, x):
        return self.net(x)


def train_step(model: nn.Module, dp: DataParallel, x: torch.Tensor, y: torch.Tensor, lr: float = 1e-2) -> float:
    model.zero_grad(set_to_none=True)
    loss = nn.functional.mse_loss(model(x), y)
    loss.backward()
    dp.allreduce_gradients(list(model.parameters()))
    with torch.no_grad():
        for p in model.parameters():
            p.add_(p.grad, alpha=-lr)
    return float(loss.detach())
