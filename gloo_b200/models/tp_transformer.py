"""A tensor-parallel transformer block on the library's collectives (Megatron layout).

Attention: the QKV projection is column-parallel (every rank owns H/P heads), attention runs
locally over those heads, the output projection is row-parallel -> ONE allreduce. MLP: column-
parallel up-projection, GELU, row-parallel down-projection -> ONE allreduce. Backward mirrors it
(one allreduce per half). CUDA tensors run the fused NVLink kernels, CPU tensors the TCP transport.
``DenseBlock`` is the single-process reference with the same parameters (tests load one into the other).
"""
from __future__ import annotations

import torch
from torch import nn

from ..parallel import TensorParallel


class DenseBlock(nn.Module):
    def __init__(self, d_model=64, n_heads=4, d_ff=256):
        super().__init__()
        self.n_heads = n_heads
        self.ln1, self.ln2 = nn.LayerNorm(d_model), nn.LayerNorm(d_model)
        self.qkv, self.proj = nn.Linear(d_model, 3 * d_model), nn.Linear(d_model, d_model)
        self.up, self.down = nn.Linear(d_model, d_ff), nn.Linear(d_ff, d_model)

    def forward(self, x):                                           # x: [B, S, d_model]
        B, S, E = x.shape
        h = self.ln1(x)
        q, k, v = self.qkv(h).view(B, S, 3, self.n_heads, E // self.n_heads).unbind(2)
        a = nn.functional.scaled_dot_product_attention(q.transpose(1, 2), k.transpose(1, 2), v.transpose(1, 2), is_causal=True)
        x = x + self.proj(a.transpose(1, 2).reshape(B, S, E))
        return x + self.down(nn.functional.gelu(self.up(self.ln2(x))))


class TPBlock(nn.Module):
    def __init__(self, tp: TensorParallel, d_model=64, n_heads=4, d_ff=256, device=None, dtype=None):
        super().__init__()
        assert n_heads % tp.size == 0 and d_ff % tp.size == 0
        self.tp, self.n_heads, self.d_model = tp, n_heads, d_model
        self.local_heads = n_heads // tp.size
        self.ln1, self.ln2 = nn.LayerNorm(d_model, device=device, dtype=dtype), nn.LayerNorm(d_model, device=device, dtype=dtype)
        # q, k and v each column-parallel over heads (one fused parameter per projection kind)
        self.q = tp.column_linear(d_model, d_model, device=device, dtype=dtype)
        self.k = tp.column_linear(d_model, d_model, device=device, dtype=dtype)
        self.v = tp.column_linear(d_model, d_model, device=device, dtype=dtype)
        self.proj = tp.row_linear(d_model, d_model, device=device, dtype=dtype)
        self.up = tp.column_linear(d_model, d_ff, device=device, dtype=dtype)
        self.down = tp.row_linear(d_ff, d_model, device=device, dtype=dtype)

    def load_dense(self, dense: DenseBlock):
        """Take this rank's shards of a DenseBlock's parameters."""
        E, Hh = self.d_model, self.n_heads
        w = dense.qkv.weight.detach().view(3, Hh, E // Hh, E)       # qkv rows are laid out (kind, head, dim)
        b = dense.qkv.bias.detach().view(3, Hh, E // Hh)
        for i, lin in enumerate((self.q, self.k, self.v)):
            lin.load_full(w[i].reshape(E, E), b[i].reshape(E))
        self.proj.load_full(dense.proj.weight.detach(), dense.proj.bias.detach())
        self.up.load_full(dense.up.weight.detach(), dense.up.bias.detach())
        self.down.load_full(dense.down.weight.detach(), dense.down.bias.detach())
        self.ln1.load_state_dict(dense.ln1.state_dict())
        self.ln2.load_state_dict(dense.ln2.state_dict())

    def forward(self, x):
        B, S, E = x.shape
        hd = E // self.n_heads
        h = self.ln1(x)
        q, k, v = (lin(h).view(B, S, self.local_heads, hd).transpose(1, 2) for lin in (self.q, self.k, self.v))
        a = nn.functional.scaled_dot_product_attention(q, k, v, is_causal=True)
        x = x + self.proj(a.transpose(1, 2).reshape(B, S, self.local_heads * hd))
        return x + self.down(nn.functional.gelu(self.up(self.ln2(x))))
