"""The CPU oracle against the committed golden vectors (generated from the REAL reference by
oracle/make_golden.py).  Runs everywhere, no GPU, no /root/reference."""
import os

import numpy as np
import pytest
import torch

import oracle.ani_oracle as orc
from helpers import GOLD, GOLDEN_CASES, golden_inputs, load_golden, oracle_model


@pytest.mark.parametrize("name", GOLDEN_CASES)
def test_oracle_matches_reference_golden(name):
    rec = load_golden(name)
    species, coords, cell, pbc = golden_inputs(rec, torch.float64)
    model = oracle_model(rec["kind"], torch.float64, rec["neighborlist"])
    out = orc.compute(model, species, coords, cell, pbc)
    assert int(out["num_pairs"]) == int(rec["num_pairs"])
    if "aev" in rec:
        assert np.abs(out["aev"].numpy() - rec["aev"]).max() < 1e-12
    assert np.abs(out["aev"].numpy() @ rec["aev_proj_matrix"] - rec["aev_proj"]).max() < 1e-10
    assert np.abs(out["aev"].numpy().sum(-1) - rec["aev_rowsum"]).max() < 1e-10
    assert np.abs(out["member_atomic"].numpy() - rec["member_atomic"]).max() < 1e-12
    assert np.abs(out["energy_nn"].numpy() - rec["energy_nn"]).max() < 1e-10
    assert np.abs(out["forces"].numpy() - rec["forces"]).max() < 1e-12


def test_loop_oracle_matches_vector_oracle():
    """Independent loop-level restatement (per-term formulas) vs the vectorised oracle."""
    rec = load_golden("water30_pbc_ani2x")
    species, coords, cell, pbc = golden_inputs(rec)
    spec = orc.aev_spec_2x()
    loop = orc.aev_loops(spec, species[0].tolist(), coords[0].numpy(), cell.numpy())
    assert np.abs(loop - rec["aev"][0]).max() < 1e-12
    rec = load_golden("kat2x5_ani2x")
    species, coords, _, _ = golden_inputs(rec)
    for c in range(2):
        loop = orc.aev_loops(spec, species[c].tolist(), coords[c].numpy())
        assert np.abs(loop - rec["aev"][c]).max() < 1e-12


def test_cell_list_equals_all_pairs():
    """tests/test_neighbors.py:173-290 of the reference: both lists give the same pair set."""
    _, idx, coords, cell, pbc = orc.water_box(60, seed=3, dtype=torch.float64)
    a = orc.cell_list(5.1, idx, coords, cell, pbc)
    b = orc.all_pairs(5.1, idx, coords, cell, pbc)
    assert a.indices.shape[1] == b.indices.shape[1]
    assert torch.allclose(a.distances.sort().values, b.distances.sort().values, atol=1e-12)


def test_oracle_fp32_close_to_fp64():
    rec = load_golden("water30_pbc_ani2x")
    species, coords, cell, pbc = golden_inputs(rec, torch.float32)
    out = orc.compute(oracle_model("2x", torch.float32, "cell_list"), species, coords, cell, pbc)
    assert np.abs(out["forces"].numpy() - rec["forces"]).max() < 1e-5
    assert np.abs(out["aev"].numpy() - rec["aev"]).max() < 1e-4


@pytest.mark.parametrize("name", ["water30_pbc_ani2x", "benzene_pbc_ani2x", "tightcell_pbc_ani2x"])
def test_oracle_stress_matches_reference_fdotr(name):
    """tests/golden/stress_pbc_ani2x.npz holds the reference's own "f dot r" stress (ase.py:164-168, computed by
    oracle/make_golden.py --stress with the real reference modules); the oracle's strain-derivative stress must
    reproduce it."""
    fix = np.load(os.path.join(GOLD, "stress_pbc_ani2x.npz"))
    rec = load_golden(name)
    sp, co, cell, pbc = golden_inputs(rec, torch.float64)
    co = co - torch.floor(co @ torch.linalg.inv(cell)) @ cell
    out = orc.compute(oracle_model("2x", torch.float64, "cell_list"), sp, co, cell, pbc, stress=True)
    assert float(np.abs(out["stress"].numpy() - fix[name]).max()) < 1e-12
