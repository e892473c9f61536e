r"""ASE ``Calculator`` for the B200 path -- the counterpart of ``torchani.ase.Calculator`` (ase.py:32-173).

    calc = model.ase()                 # torchani_b200.models.ANI.ase(overwrite, stress_kind)
    atoms.calc = calc                  # energy / free_energy / forces / stress, ASE units (eV, eV/A, eV/A^3)

The reference converts numpy -> tensors -> GPU on every call, runs the model under autograd and reads the
energy and the forces back with two blocking copies.  Here ``calculate`` hands the positions to a persistent
``calculator.HostCalculator`` (pinned staging buffers, the whole step -- H2D, kernels, D2H -- one captured CUDA
graph, one synchronisation); a new one is built only when the atomic numbers or the periodicity change.

``stress_kind`` (ase.py:52-55,142-156): ``"fdotr"`` and ``"scaling"`` both come from the virial the force kernel
accumulates (sum over pairs of dE/dDelta (x) Delta, divided by the volume -- for a potential that depends on the
pair vectors only, the strain derivative of ``"scaling"`` IS that virial, DESIGN.md / tests/test_gpu_api.py);
``"numerical"`` uses ASE's own finite-difference ``calculate_numerical_stress``.

``ase`` is imported lazily and only here, exactly as in the reference: without it this module raises the
reference's ImportError and ``ANI.ase()`` is unavailable -- nothing else in the package depends on it.
"""
from __future__ import annotations

import typing as tp
import warnings

import numpy as np

try:
    import ase.units
    from ase.calculators.calculator import Calculator as AseCalculator, all_changes
except ImportError:
    raise ImportError(
        "Error when trying to import 'torchani_b200.ase':"
        " The ASE package could not be found. 'torchani_b200.ase' and the '*.ase()' methods"
        " of models won't be available. Please install ase if you want to use them."
    ) from None

from .calculator import HostCalculator

StressKind = str   # "scaling" | "fdotr" | "numerical"  (annotations.py of the reference)


class Calculator(AseCalculator):
    """TorchANI-B200 calculator for ASE (same constructor and results as ase.py:32-173)."""

    implemented_properties = ["energy", "free_energy", "forces", "stress"]

    def __init__(self, model, overwrite: bool = False, stress_kind: StressKind = "scaling", skin: float = 0.0):
        super().__init__()
        self.model = model
        buf = next(model.buffers())
        self.device = buf.device
        if not model.periodic_table_index:
            raise ValueError("ASE models must have periodic_table_index=True")
        if stress_kind not in ("scaling", "fdotr", "numerical"):
            raise ValueError(f"Unsupported stress kind {stress_kind}")
        self.overwrite = overwrite
        self.stress_kind = stress_kind
        self.skin = skin
        self._host: tp.Optional[HostCalculator] = None
        self._host_key: tp.Any = None

    def _host_calculator(self, numbers: np.ndarray, periodic: bool, cell) -> HostCalculator:
        key = (numbers.tobytes(), periodic)
        if self._host is None or self._host_key != key:
            self._host = HostCalculator(self.model, numbers, cell if periodic else None, pbc=periodic, skin=self.skin)
            self._host_key = key
        return self._host

    def calculate(self, atoms=None, properties=["energy"], system_changes=all_changes):
        super().calculate(atoms, properties, system_changes)
        assert self.atoms is not None
        needs_stress = "stress" in properties
        numbers = np.asarray(self.atoms.get_atomic_numbers(), dtype=np.int64)
        positions = np.asarray(self.atoms.get_positions(), dtype=np.float64)
        cell = np.asarray(self.atoms.get_cell(complete=True).array if hasattr(self.atoms.get_cell(complete=True), "array")
                          else self.atoms.get_cell(complete=True), dtype=np.float64)
        pbc = np.asarray(self.atoms.get_pbc(), dtype=bool)
        periodic = bool(pbc.any())
        if periodic and not pbc.all():
            # PBC in some directions only (slabs, wires): an equivalent fully periodic cell whose non-periodic
            # lattice vectors are long enough that no pair within the cutoff crosses them (neighbors.py:214-275
            # enumerates image shifts along the periodic vectors only -- same pair set)
            import torch
            from .neighbors import effective_periodic_cell
            cell = effective_periodic_cell(torch.from_numpy(positions), torch.from_numpy(cell), torch.from_numpy(pbc),
                                           float(self.model.cutoff)).numpy()
        if periodic and self.overwrite:
            warnings.warn("'overwrite' set, info about crossing PBC *will be lost*")
            frac = positions @ np.linalg.inv(cell)
            positions = (frac - np.floor(frac)) @ cell            # utils.py:237-255 (map_to_central)
            self.atoms.set_positions(positions)
        host = self._host_calculator(numbers, periodic, cell)
        if periodic and (host.h_cell.numpy().reshape(3, 3) != cell.astype(np.float32)).any():
            host.set_cell(cell)
        ha = ase.units.Hartree
        if needs_stress and self.stress_kind in ("scaling", "fdotr"):
            if not periodic:
                raise ValueError("the stress needs a periodic cell")
            e, f, stress = host.calculate_with_stress(positions)
            self.results["stress"] = stress * ha
        else:
            e, f = host.calculate(positions)
        self.results["energy"] = e * ha
        self.results["free_energy"] = e * ha
        self.results["forces"] = f.astype(np.float64) * ha
        if needs_stress and self.stress_kind == "numerical":
            self.results["stress"] = self.calculate_numerical_stress(self.atoms)
