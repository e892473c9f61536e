"""TEST STUB -- NOT the ASE package.  The image has no `ase`; this is the minimal protocol that
``torchani.ase.Calculator`` / ``torchani_b200.ase.Calculator`` rely on (``ase.Atoms`` accessors,
``ase.units.Hartree``, ``ase.calculators.calculator.Calculator`` / ``all_changes``), so that the calculator shim can
be exercised.  tests/test_gpu_ase.py puts this directory on sys.path ONLY when the real package is absent."""
import copy

import numpy as np

IS_TEST_STUB = True


class _Cell:
    def __init__(self, array):
        self.array = np.array(array, dtype=float).reshape(3, 3)

    def __array__(self, dtype=None, copy=None):
        return self.array if dtype is None else self.array.astype(dtype)


class Atoms:
    def __init__(self, numbers, positions, cell=None, pbc=False):
        self.numbers = np.array(numbers, dtype=int)
        self.positions = np.array(positions, dtype=float).reshape(-1, 3)
        self.cell = _Cell(np.zeros((3, 3)) if cell is None else cell)
        self.pbc = np.array([pbc] * 3 if isinstance(pbc, bool) else pbc, dtype=bool)
        self.calc = None

    def copy(self):
        a = Atoms(self.numbers.copy(), self.positions.copy(), self.cell.array.copy(), self.pbc.copy())
        return a

    def get_atomic_numbers(self):
        return self.numbers.copy()

    def get_positions(self):
        return self.positions.copy()

    def set_positions(self, p):
        self.positions = np.array(p, dtype=float).reshape(-1, 3)

    def get_cell(self, complete=False):
        return _Cell(self.cell.array.copy())

    def set_cell(self, cell, scale_atoms=False):
        new = np.array(cell, dtype=float).reshape(3, 3)
        if scale_atoms:
            self.positions = self.positions @ np.linalg.inv(self.cell.array) @ new
        self.cell = _Cell(new)

    def get_pbc(self):
        return self.pbc.copy()

    def get_volume(self):
        return abs(float(np.linalg.det(self.cell.array)))

    def get_potential_energy(self, force_consistent=False):
        return self.calc.get_property("free_energy" if force_consistent else "energy", self)

    def get_forces(self):
        return self.calc.get_property("forces", self)

    def get_stress(self):
        return self.calc.get_property("stress", self)

    def __len__(self):
        return len(self.numbers)
