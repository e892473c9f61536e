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


def protein_in_water(n_atoms: int = 50001, seed: int = 0, density: float = 0.1, pad: float = 3.0,
                     protein_npz: tp.Optional[str] = None, dtype: torch.dtype = torch.float32):
    """BASELINE config 5: the 1C17 protein (H C N O S; tests/golden/1c17_protein.npz, written by
    oracle/make_protein_fixture.py from the reference's dataset/pdb/1C17.pdb) in a cubic periodic box of
    ``L = (n_atoms / density)^(1/3)`` (79.4 A for 50 001 atoms), the remaining volume filled with lattice water.

    1C17 is elongated (78 x 96 x 86 A): the atoms that fit into the box with ``pad`` A to every face are kept
    (the protein is cut, which does not matter for a throughput workload; composition and local density are
    those of the protein).  Water oxygens sit on a jittered cubic lattice wherever no solute atom is within
    2.6 A; hydrogens as in ``water_box``.  Returns (Z (1,N), element index (1,N), coords (1,N,3), cell, pbc)
    with N <= n_atoms (the water count is rounded down to whole molecules)."""
    import os
    if protein_npz is None:
        protein_npz = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden",
                                   "1c17_protein.npz")
    rec = np.load(protein_npz)
    p_idx, p_xyz = rec["species"].astype(np.int64), rec["coords"].astype(np.float64)
    L = (n_atoms / density) ** (1.0 / 3.0)
    p_xyz = p_xyz - 0.5 * (p_xyz.max(0) + p_xyz.min(0)) + 0.5 * L
    inside = ((p_xyz >= pad) & (p_xyz <= L - pad)).all(1)
    p_idx, p_xyz = p_idx[inside], p_xyz[inside]
    n_water = (n_atoms - len(p_idx)) // 3
    rng = np.random.default_rng(seed)
    # occupancy grid of the solute (2.6 A exclusion around every atom, periodic)
    h = 0.65
    ng = int(math.ceil(L / h))
    occ = np.zeros((ng, ng, ng), dtype=bool)
    r = int(math.ceil(2.6 / (L / ng)))
    offs = np.array([(a, b, c) for a in range(-r, r + 1) for b in range(-r, r + 1) for c in range(-r, r + 1)
                     if (a * a + b * b + c * c) * (L / ng) ** 2 <= 2.6 ** 2])
    cells = np.floor(p_xyz / (L / ng)).astype(np.int64)
    for o in offs:
        q = (cells + o) % ng
        occ[q[:, 0], q[:, 1], q[:, 2]] = True
    # candidate oxygen sites: cubic lattice, densest spacing that still offers enough free sites
    k = int(math.ceil((n_water * 1.02 / max(1e-9, 1.0 - occ.mean())) ** (1.0 / 3.0)))
    while True:
        g = np.stack(np.meshgrid(*[np.arange(k)] * 3, indexing="ij"), -1).reshape(-1, 3)
        sites = (g + 0.5) * (L / k)
        c = np.floor(sites / (L / ng)).astype(np.int64) % ng
        free = sites[~occ[c[:, 0], c[:, 1], c[:, 2]]]
        if len(free) >= n_water:
            break
        k += 1
    pick = np.sort(rng.choice(len(free), size=n_water, replace=False))
    o = free[pick] + (rng.random((n_water, 3)) - 0.5) * 0.3
    d = rng.normal(size=(n_water, 2, 3))
    d /= np.linalg.norm(d, axis=-1, keepdims=True)
    hyd = o[:, None, :] + 0.9572 * d
    w_xyz = np.concatenate([o[:, None, :], hyd], 1).reshape(-1, 3)
    w_idx = np.tile(np.array([3, 0, 0]), n_water)
    idx = np.concatenate([p_idx, w_idx])
    xyz = np.concatenate([p_xyz, w_xyz])
    z_of = np.array([1, 6, 7, 8, 16, 9, 17])
    return (torch.tensor(z_of[idx]).view(1, -1), torch.tensor(idx).view(1, -1),
            torch.tensor(xyz, dtype=dtype).view(1, -1, 3), torch.eye(3, dtype=dtype) * L,
            torch.tensor([True, True, True]))
