"""Host-side helpers of the VMM path: responsibility initialisation and the view prior.

Mirrors the call surface of the reference's tools/NEM_utilities.py (bin_uniform :29-36,
bin_rand :19-27, compute_norm :9-17, compute_prior :66-72, clip_gradient :60-64) with the
same names and argument meaning.  These run once per batch on the host (tiny K x M
tables); the per-object work happens in the CUDA library.
"""
from __future__ import annotations

import math
import random

import numpy as np
import torch


def compute_norm(center: int, m: int) -> torch.Tensor:
    """Gaussian(0, 0.15) pdf sampled on linspace(-.5, .5, m), rolled so its mode sits on
    view `center`, normalised to sum 1; shape (m, 1)  (tools/NEM_utilities.py:9-17)."""
    grid = np.linspace(-0.5, 0.5, m, endpoint=True)
    grid = np.roll(grid, int(m / 2) - center)
    y = grid / 0.15
    pdf = np.exp(-y * y / 2.0) / (math.sqrt(2.0 * math.pi) * 0.15)
    return torch.Tensor(pdf / np.sum(pdf)).unsqueeze(1)


def view_tables(k: int, m: int, init_type: str = "bin_uniform"):
    """(K, M) one-hot responsibility table and (K, M) view prior shared by every object
    of the batch (the reference builds them per object and they are identical across B)."""
    gamma = torch.zeros(k, m)
    prior = torch.zeros(k, m)
    if init_type == "bin_uniform":
        stride = int(m / k)
        views = [i * stride for i in range(k)]
    elif init_type == "bin_rand":
        order = list(range(m))
        random.shuffle(order)          # one draw for the whole batch, as the reference does
        views = order[:k]
    else:
        raise KeyError(init_type)
    for i, v in enumerate(views):
        gamma[i, v] = 1.0
        prior[i] = compute_norm(v, m).squeeze(1)
    return gamma, prior


def bin_uniform(B: int, k: int, m: int):
    """(gamma, gamma_prior), each (B, k, m, 1)  (tools/NEM_utilities.py:29-36)."""
    g, p = view_tables(k, m, "bin_uniform")
    return g.view(1, k, m, 1).repeat(B, 1, 1, 1), p.view(1, k, m, 1).repeat(B, 1, 1, 1)


def bin_rand(B: int, k: int, m: int):
    """(tools/NEM_utilities.py:19-27)."""
    g, p = view_tables(k, m, "bin_rand")
    return g.view(1, k, m, 1).repeat(B, 1, 1, 1), p.view(1, k, m, 1).repeat(B, 1, 1, 1)


def compute_prior(type: str, device=None):
    """Bernoulli prior placeholder: zeros (1,1,1,1) for 'fc' features (tools/NEM_utilities.py:66-72)."""
    shape = (1, 1, 1, 1) if type == "fc" else (1, 1, 1, 1, 1, 1)
    return torch.zeros(*shape, device=device if device is not None else "cuda")


def clip_gradient(optimizer, grad_clip):
    """Clamp every existing gradient to [-c, c]  (tools/NEM_utilities.py:60-64)."""
    for group in optimizer.param_groups:
        for param in group["params"]:
            if param.grad is not None:
                param.grad.data.clamp_(-grad_clip, grad_clip)
