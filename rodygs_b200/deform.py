"""Time embedding + motion-basis MLP, state-dict compatible with the reference's
`MLPBasisNetwork` (/root/reference/src/model/rodygs_dynamic.py:243-327) and
`TimestepEmbedder` (:192-220).

This part of the path is 70 K MACs per query time and stays in PyTorch (SURVEY.md
§8 a1); what changes is the launch count: the 16 per-basis heads, which the
reference runs as a Python loop of 16 x 2 `nn.Linear` calls (:314-317), are
evaluated as two batched matmuls over stacked weights, and B(t) for the query time
and for all T training times come out of ONE forward pass ([T+1, 53] batch).
Its outputs feed the fused kernel (rodygs_b200.dynamic.render_dynamic).
"""
from __future__ import annotations

import math

import torch
import torch.nn as nn
import torch.nn.functional as F


class TimestepEmbedder(nn.Module):
    """emb(t) = [t, sin(f_0 pi t), cos(f_0 pi t), sin(f_1 pi t), ...] (rodygs_dynamic.py:202-220)."""

    def __init__(self, emb_multires: int, input_dims: int = 1, log_sampling: bool = False):
        super().__init__()
        self.num_freqs = emb_multires
        if log_sampling:
            freqs = 2.0 ** torch.linspace(0.0, emb_multires - 1, emb_multires)
        else:
            freqs = torch.linspace(1.0, 2.0 ** (emb_multires - 1), emb_multires)
        self.register_buffer("freqs_pi", freqs * math.pi, persistent=False)

    def forward(self, t: torch.Tensor) -> torch.Tensor:
        t = torch.as_tensor(t, dtype=torch.float32, device=self.freqs_pi.device)
        ang = t.unsqueeze(-1) * self.freqs_pi                       # [..., F]
        sc = torch.stack((torch.sin(ang), torch.cos(ang)), dim=-1)  # [..., F, 2] -> interleaved sin, cos
        return torch.cat((t.unsqueeze(-1), sc.flatten(-2)), dim=-1)


class _Head(nn.Module):
    """Parameter container with the reference's names: basis_xyz.{k}.basis.{0,2}."""

    def __init__(self, width: int, out_dim: int, act: nn.Module):
        super().__init__()
        self.basis = nn.Sequential(nn.Linear(width, width // 2), act, nn.Linear(width // 2, out_dim))
        for m in self.basis.modules():
            if isinstance(m, nn.Linear):
                nn.init.normal_(m.weight, mean=0, std=1e-2)
                nn.init.constant_(m.bias, 0)


class MotionBasisNetwork(nn.Module):
    trans_dim = 3
    rot_dim = 4

    def __init__(self, netwidth: int = 128, num_basis: int = 16, t_emb_multires: int = 26,
                 t_log_sampling: bool = False, activation: str = "gelu"):
        super().__init__()
        self.num_basis = num_basis
        self.t_embed_dim = t_emb_multires * 2 + 1
        self.t_embedder = TimestepEmbedder(t_emb_multires, 1, t_log_sampling)
        self.activation = nn.GELU() if activation.lower() != "relu" else nn.ReLU(inplace=False)
        self.timenet = nn.Sequential(
            nn.Linear(self.t_embed_dim, netwidth), self.activation,
            nn.Linear(netwidth, netwidth), self.activation,
            nn.Linear(netwidth, netwidth // 2), self.activation)
        for m in self.timenet.modules():
            if isinstance(m, nn.Linear):
                nn.init.normal_(m.weight, mean=0, std=1e-2)
                nn.init.constant_(m.bias, 0)
        self.basis_xyz = nn.ModuleList([_Head(netwidth // 2, 7, self.activation) for _ in range(num_basis)])

    def bases(self, t_embs: torch.Tensor) -> torch.Tensor:
        """t_embs [..., 53] -> B [..., num_basis, 7]; equals `batch_inference` (:296-306)."""
        h = self.timenet(t_embs)
        w0 = torch.stack([hd.basis[0].weight for hd in self.basis_xyz])   # [K, w/2, w]
        b0 = torch.stack([hd.basis[0].bias for hd in self.basis_xyz])     # [K, w/2]
        w2 = torch.stack([hd.basis[2].weight for hd in self.basis_xyz])   # [K, 7, w/2]
        b2 = torch.stack([hd.basis[2].bias for hd in self.basis_xyz])     # [K, 7]
        u = self.activation(torch.einsum("...i,koi->...ko", h, w0) + b0)
        return torch.einsum("...ki,koi->...ko", u, w2) + b2

    def batch_embedding(self, timesteps: torch.Tensor) -> torch.Tensor:
        return self.t_embedder(timesteps)

    def batch_inference(self, t_embs: torch.Tensor) -> torch.Tensor:
        return self.bases(t_embs)

    def query_and_table(self, t: torch.Tensor, train_time_embs: torch.Tensor):
        """(B(t) [K,7], table [T,K,7]) from a single batched pass."""
        emb = torch.cat((self.t_embedder(t).reshape(1, -1), train_time_embs), dim=0)
        out = self.bases(emb)
        return out[0], out[1:]

    def forward(self, coeff: torch.Tensor, timestep: torch.Tensor):
        """Reference signature (:308-327): returns (translation, rotation) = coeff @ B(t)."""
        basis = self.bases(self.t_embedder(timestep))
        tot = torch.squeeze(torch.squeeze(coeff) @ basis)
        return tot[..., : self.trans_dim], tot[..., self.trans_dim:]
