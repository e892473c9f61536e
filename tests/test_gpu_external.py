"""neighbour lists that come from OUTSIDE -- the entry points MD engines use.

* ``ANI.compute_from_external_neighbors`` (arch.py:171-206): a half pair list built with a skin + lattice shifts,
  screened by ``narrow_down`` (neighbors.py:64-113);
* ``AEVComputer.compute_from_full_neighborlist`` / ``ANI.compute_from_full_neighborlist``: the LAMMPS-style full
  list (ilist / jlist / numneigh) over local + GHOST atoms (``cuaev::run_with_full_nbrlist``,
  aev/_computer.py:409-438, csrc/aev.cu:1048-1126,1868-1956).

Both must reproduce the periodic calculation they stand for: energies, AEVs and -- after folding the ghost
gradients back onto their parent atoms, which is what the MD engine does -- forces, checked against the float64
oracle of the periodic system."""
import numpy as np
import pytest
import torch

import oracle.ani_oracle as orc
from helpers import AEV_ATOL, AEV_RTOL, F_ATOL, assert_close, oracle_model

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def _model(members=8):
    from torchani_b200 import models, synthetic
    w = synthetic.make_weights(models.SYMBOLS_2X, synthetic.DIMS_2X, 1008, members, seed=1234)
    return models.from_weight_lists("2x", w, device=DEV, periodic_table_index=True)


def _box():
    from torchani_b200 import synthetic
    z, idx, coords, cell, pbc = synthetic.water_box(100, seed=4)          # 300 atoms, L = 14.42 A
    ref = orc.compute(oracle_model("2x", torch.float64), idx, coords.double(), cell.double(), pbc)
    return z, idx, coords, cell, pbc, ref


def test_compute_from_external_neighbors_with_a_skin_list():
    from torchani_b200.neighbors import CellList
    model = _model()
    z, idx, coords, cell, pbc, ref = _box()
    zd, cd, celld, pbcd = z.to(DEV), coords.to(DEV), cell.to(DEV), pbc.to(DEV)
    wide = CellList()(6.5, model.species_converter(zd), cd, celld, pbcd)                # cutoff 5.1 + skin 1.4
    flat = cd.view(-1, 3)
    shifts = wide.diff_vectors - (flat[wide.indices[0]] - flat[wide.indices[1]])
    assert wide.indices.shape[1] > 1.5 * int(ref["num_pairs"])
    c = cd.clone().requires_grad_(True)
    e = model.compute_from_external_neighbors(zd, c, wide.indices, shifts)
    (g,) = torch.autograd.grad(e.sum(), c)
    assert abs(float(e[0]) - float(ref["energy"][0])) < 5e-3                              # float32 total
    assert_close("forces", -g.cpu().numpy(), ref["forces"].numpy(), 0.0, F_ATOL)
    at = model.compute_from_external_neighbors(zd, cd, wide.indices, shifts, atomic=True)
    assert at.shape == (1, 300)
    # a dummy atom in the pair list is dropped (neighbors.py:72-82)
    z2 = zd.clone()
    z2[0, 5] = -1
    e2 = model.compute_from_external_neighbors(z2, cd, wide.indices, shifts)
    assert bool(torch.isfinite(e2).all()) and abs(float(e2[0]) - float(e[0])) > 1e-3


def _with_ghosts(coords, idx, L, halo):
    """local atoms + every periodic image that lies within `halo` of the box: (coords_all, idx_all, parent)."""
    x = coords[0].double().numpy() % L
    sp = idx[0].numpy()
    pos, spc, par = [x], [sp], [np.arange(len(x))]
    for a in (-1, 0, 1):
        for b in (-1, 0, 1):
            for c in (-1, 0, 1):
                if (a, b, c) == (0, 0, 0):
                    continue
                y = x + np.array([a, b, c]) * L
                keep = ((y > -halo) & (y < L + halo)).all(1)
                pos.append(y[keep]); spc.append(sp[keep]); par.append(np.arange(len(x))[keep])
    return np.concatenate(pos), np.concatenate(spc), np.concatenate(par)


def test_full_neighbor_list_with_ghost_atoms():
    model = _model()
    z, idx, coords, cell, pbc, ref = _box()
    L, n_loc, rlist = float(cell[0, 0]), 300, 6.0                                          # list cutoff = 5.1 + skin
    pos, spc, parent = _with_ghosts(coords, idx, L, rlist)
    n_all = len(pos)
    assert n_all > 3 * n_loc
    d = np.linalg.norm(pos[:n_loc, None, :] - pos[None, :, :], axis=-1)
    within = d <= rlist
    within[np.arange(n_loc), np.arange(n_loc)] = False
    ilist = np.arange(n_loc)
    numneigh = within.sum(1)
    jlist = np.concatenate([np.nonzero(within[i])[0] for i in range(n_loc)])
    to_z = torch.tensor([1, 6, 7, 8, 16, 9, 17])
    z_all = to_z[torch.tensor(spc)].view(1, -1).to(DEV)
    c_all = torch.tensor(pos, dtype=torch.float32).view(1, -1, 3).to(DEV).requires_grad_(True)
    il, jl, nn = (torch.tensor(a, dtype=torch.int32, device=DEV) for a in (ilist, jlist, numneigh))
    # AEVs of the local atoms == periodic AEVs; ghost rows are zero
    aev = model.aev_computer.compute_from_full_neighborlist(model.species_converter(z_all), c_all, il, jl, nn)
    assert aev.shape == (1, n_all, 1008) and float(aev[0, n_loc:].abs().max()) == 0.0
    assert_close("aev", aev[0, :n_loc].detach().cpu().numpy(), ref["aev"][0].numpy(), AEV_RTOL, AEV_ATOL)
    # energy of the local atoms and forces after folding the ghost gradients onto their parents
    e = model.compute_from_full_neighborlist(z_all, c_all, il, jl, nn)
    (g,) = torch.autograd.grad(e.sum(), c_all)
    folded = torch.zeros(n_loc, 3, dtype=torch.float64)
    folded.index_add_(0, torch.tensor(parent), g[0].double().cpu())
    assert abs(float(e[0]) - float(ref["energy"][0])) < 5e-3
    assert_close("forces", -folded.numpy(), ref["forces"][0].numpy(), 0.0, F_ATOL)
    # only a subset of local atoms requested: the other rows stay zero, energies are those atoms' only
    half = model.compute_from_full_neighborlist(z_all, c_all.detach(), il[:150], jl[: int(numneigh[:150].sum())],
                                                nn[:150], atomic=True)
    full = model.compute_from_full_neighborlist(z_all, c_all.detach(), il, jl, nn, atomic=True)
    assert float(half[0, 150:].abs().max()) == 0.0
    assert float((half[0, :150] - full[0, :150]).abs().max()) < 1e-6
    with pytest.raises(ValueError):
        model.aev_computer.compute_from_full_neighborlist(torch.zeros(2, 3, dtype=torch.long, device=DEV),
                                                          torch.zeros(2, 3, 3, device=DEV), il, jl, nn)
