"""Seeded synthetic inputs and weights: there is no network for datasets or
pretrained ANI parameters, so benchmarks and smoke tests use these deterministic generators.
(The test oracle carries its own copy; tests/test_host_logic.py checks they agree.)"""
from __future__ import annotations

import math
import typing as tp

import numpy as np
import torch

DIMS_2X = {"H": (256, 192, 160), "C": (224, 192, 160), "N": (192, 160, 128), "O": (192, 160, 128),
           "S": (160, 128, 96), "F": (160, 128, 96), "Cl": (160, 128, 96)}
DIMS_1X = {"H": (160, 128, 96), "C": (144, 112, 96), "N": (128, 112, 96), "O": (128, 112, 96)}


def make_weights(symbols: tp.Sequence[str], dims: tp.Dict[str, tp.Tuple[int, ...]], in_dim: int, members: int,
                 seed: int = 1234, dtype: torch.dtype = torch.float32):
    """weights[member][symbol] = [(W [out,in], b [out]) x 4], uniform in +-1/sqrt(fan_in) from numpy's
    PCG64 stream (stable across versions and machines)."""
    rng = np.random.default_rng(seed)
    out = []
    for _ in range(members):
        per_sym = {}
        for s in symbols:
            layer_dims = (in_dim,) + tuple(dims[s]) + (1,)
            layers = []
            for fan_in, fan_out in zip(layer_dims[:-1], layer_dims[1:]):
                bound = 1.0 / math.sqrt(fan_in)
                w = rng.uniform(-bound, bound, size=(fan_out, fan_in))
                b = rng.uniform(-bound, bound, size=(fan_out,))
                layers.append((torch.tensor(w, dtype=dtype), torch.tensor(b, dtype=dtype)))
            per_sym[s] = layers
        out.append(per_sym)
    return out


def water_box(n_molecules: int, seed: int = 0, density: float = 0.1, dtype: torch.dtype = torch.float32):
    """Periodic water box: O on a jittered cubic lattice, two H at 0.9572 A in random directions.
    Returns (Z (1,N), element index (1,N), coords (1,N,3), cell (3,3), pbc (3,))."""
    g = torch.Generator().manual_seed(seed)
    n_atoms = 3 * n_molecules
    L = (n_atoms / density) ** (1.0 / 3.0)
    k = int(math.ceil(n_molecules ** (1.0 / 3.0)))
    grid = torch.stack(torch.meshgrid(*[torch.arange(k)] * 3, indexing="ij"), -1).view(-1, 3)
    grid = grid[:n_molecules].to(torch.float64)
    o = (grid + 0.5) * (L / k) + (torch.rand(n_molecules, 3, generator=g, dtype=torch.float64) - 0.5) * 0.3
    dirs = torch.randn(n_molecules, 2, 3, generator=g, dtype=torch.float64)
    dirs = dirs / dirs.norm(dim=-1, keepdim=True)
    h = o.unsqueeze(1) + 0.9572 * dirs
    coords = torch.cat([o.unsqueeze(1), h], 1).view(1, n_atoms, 3).to(dtype)
    z = torch.tensor([8, 1, 1]).repeat(n_molecules).view(1, -1)
    idx = torch.tensor([3, 0, 0]).repeat(n_molecules).view(1, -1)
    return z, idx, coords, torch.eye(3, dtype=dtype) * L, torch.tensor([True, True, True])


def conformer_batch(n_conf: int = 256, a_min: int = 9, a_max: int = 26, seed: int = 1234,
                    dtype: torch.dtype = torch.float32):
    """GDB-11-like padded batch: random-walk chains of H/C/N/O, -1 padding, no PBC."""
    rng = np.random.default_rng(seed)
    species = -np.ones((n_conf, a_max), dtype=np.int64)
    coords = np.zeros((n_conf, a_max, 3))
    for c in range(n_conf):
        n = int(rng.integers(a_min, a_max + 1))
        pts = [np.zeros(3)]
        while len(pts) < n:
            anchor = pts[int(rng.integers(0, len(pts)))]
            d = rng.normal(size=3)
            p = anchor + d / np.linalg.norm(d) * rng.uniform(1.1, 1.5)
            if min(np.linalg.norm(p - q) for q in pts) >= 0.8:
                pts.append(p)
        coords[c, :n] = np.array(pts)
        species[c, :n] = rng.choice(4, size=n, p=[0.5, 0.3, 0.1, 0.1])
    return torch.tensor(species), torch.tensor(coords, dtype=dtype)
