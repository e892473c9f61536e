"""The ASE calculator shim (torchani_b200.ase.Calculator, counterpart of torchani/ase.py:32-173).

The image has no `ase`; when the real package is absent a TEST STUB with the minimal Atoms / Calculator protocol
(tests/stubs/ase) is put on sys.path so that the shim's calculate() -- host positions in, energy / forces / stress
in ASE units out, through the persistent HostCalculator -- is exercised.  Energies and forces are checked against
the float64 oracle, the stress against the fixture the REAL reference produced (tests/golden/stress_pbc_ani2x.npz,
its "f dot r" stress of ase.py:164-168)."""
import os
import sys

import numpy as np
import pytest
import torch

import oracle.ani_oracle as orc
from helpers import GOLD, ROOT, golden_inputs, load_golden, oracle_model

pytestmark = pytest.mark.gpu
DEV = "cuda:0"

try:
    import ase  # noqa: F401
except ImportError:
    sys.path.insert(0, os.path.join(ROOT, "tests", "stubs"))
    import ase  # noqa: F401

HA = 27.211386024367243
Z_OF = {0: 1, 1: 6, 2: 7, 3: 8, 4: 16, 5: 9, 6: 17}


def _model():
    from torchani_b200 import models, synthetic
    w = synthetic.make_weights(models.SYMBOLS_2X, synthetic.DIMS_2X, 1008, 8, seed=1234)
    return models.from_weight_lists("2x", w, device=DEV, periodic_table_index=True)


@pytest.mark.parametrize("name", ["water30_pbc_ani2x", "benzene_pbc_ani2x"])
def test_ase_calculator_energy_forces_stress(name):
    from ase import Atoms
    rec = load_golden(name)
    species, coords, cell, pbc = golden_inputs(rec, torch.float64)
    numbers = [Z_OF[int(s)] for s in species[0]]
    atoms = Atoms(numbers, coords[0].numpy(), cell=cell.numpy(), pbc=True)
    model = _model()
    om = oracle_model("2x", torch.float64)
    ref = orc.compute(om, species, coords, cell, pbc)
    stress_ref = np.load(os.path.join(GOLD, "stress_pbc_ani2x.npz"))[name]
    for kind in ("fdotr", "scaling"):
        atoms.calc = model.ase(stress_kind=kind)
        e = atoms.get_potential_energy()
        f = atoms.get_forces()
        s = atoms.get_stress()
        assert abs(e / HA - float(ref["energy"][0])) < 1e-5 * len(numbers) + 1e-4
        assert np.abs(f / HA - ref["forces"][0].numpy()).max() < 1e-4
        assert s.shape == (3, 3)
        assert np.abs(s / HA - stress_ref).max() < 2e-6 + 1e-3 * np.abs(stress_ref).max()
    # results are fresh arrays (the pinned staging buffer is reused by the next call)
    f1 = atoms.get_forces()
    keep = f1.copy()
    moved = atoms.copy()
    moved.set_positions(atoms.get_positions() + np.random.default_rng(0).normal(size=(len(numbers), 3)) * 0.05)
    moved.calc = atoms.calc
    f2 = moved.get_forces()
    assert f1 is not f2 and np.abs(f1 - keep).max() == 0.0 and np.abs(f2 - f1).max() > 1e-3
    assert np.abs(f1 - f).max() < 1e-5            # (float atomics: two evaluations agree to rounding)
    # repeated calls reuse one persistent HostCalculator (CUDA graph after a few steps)
    host = atoms.calc._host
    for _ in range(6):
        atoms.get_potential_energy()
    assert atoms.calc._host is host and len(host._graphs) == 1


def test_ase_calculator_nonperiodic_and_errors():
    from ase import Atoms
    rec = load_golden("small264_nopbc_ani2x")
    species, coords, _, _ = golden_inputs(rec, torch.float64)
    numbers = [Z_OF[int(s)] for s in species[0]]
    atoms = Atoms(numbers, coords[0].numpy())
    model = _model()
    atoms.calc = model.ase()
    om = oracle_model("2x", torch.float64)
    ref = orc.compute(om, species, coords, None, None)
    assert abs(atoms.get_potential_energy() / HA - float(ref["energy"][0])) < 1e-5 * len(numbers) + 1e-4
    assert np.abs(atoms.get_forces() / HA - ref["forces"][0].numpy()).max() < 1e-4
    with pytest.raises(ValueError):
        atoms.get_stress()                       # no cell, no stress
    with pytest.raises(ValueError):
        model.ase(stress_kind="virial")
    from torchani_b200 import models, synthetic
    w = synthetic.make_weights(models.SYMBOLS_2X, synthetic.DIMS_2X, 1008, 1, seed=1)
    with pytest.raises(ValueError):              # ase.py:69-70
        models.from_weight_lists("2x", w, device=DEV, periodic_table_index=False).ase()
