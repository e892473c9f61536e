"""TEST STUB (see ase/__init__.py): the calculator protocol of ase.calculators.calculator."""
import numpy as np

all_changes = ["positions", "numbers", "cell", "pbc", "initial_charges", "initial_magmoms"]


class Calculator:
    implemented_properties = []

    def __init__(self, **kwargs):
        self.atoms = None
        self.results = {}

    def calculate(self, atoms=None, properties=["energy"], system_changes=all_changes):
        if atoms is not None:
            self.atoms = atoms.copy()

    def get_property(self, name, atoms):
        self.results = {}
        self.calculate(atoms, [name], all_changes)
        return self.results[name]

    def calculate_numerical_stress(self, atoms, d=1e-6, voigt=True):
        """Central differences of the energy under symmetric strains (ase/calculators/calculator.py)."""
        stress = np.zeros((3, 3))
        cell = atoms.get_cell().array.copy()
        vol = atoms.get_volume()
        work = atoms.copy()
        for i in range(3):
            for j in range(i, 3):
                e = []
                for sign in (1.0, -1.0):
                    x = np.eye(3)
                    if i == j:
                        x[i, i] += sign * d
                    else:
                        x[i, j] += sign * d / 2
                        x[j, i] += sign * d / 2
                    work.set_cell(cell @ x, scale_atoms=True)
                    self.results = {}
                    self.calculate(work, ["energy"], all_changes)
                    e.append(self.results["free_energy"])
                    work.set_cell(cell, scale_atoms=True)
                stress[i, j] = stress[j, i] = (e[0] - e[1]) / (2 * d * vol)
        return stress.flat[[0, 4, 8, 5, 2, 1]] if voigt else stress
