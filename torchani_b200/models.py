"""ANI model assembly with the interface of ``torchani.arch.ANI`` / ``torchani.models``.

``ANI.forward((species, coords), cell, pbc, charge, atomic, ensemble_values)`` mirrors
arch.py:302-349 and returns ``SpeciesEnergies``; energies carry an autograd edge to ``coords``,
so ``torch.autograd.grad(energy, coords)`` (ase.py:159-162, grad.py:57-62) works unchanged.  For
the plain energy(+force) call the whole step runs through the fused ``Engine`` (one launch
sequence, forces produced in the forward pass and stashed for backward); ``atomic=True`` /
``ensemble_values=True`` use the same engine outputs.

Pretrained ANI parameters are not shipped (the reference downloads them, arch.py:1185-1220):
``ANI2x()`` / ``ANI1x()`` build the published architectures with seeded random weights, and
``load_state_dict`` accepts the reference's state-dict keys.  ``from_torchani(model)`` converts
an existing ``torchani.arch.ANI`` instance (weights, AEV constants, self energies).
"""
from __future__ import annotations

import typing as tp

import torch
from torch import Tensor

from . import _lib
from .aev import AEVComputer
from .engine import Engine, StepResult
from .neighbors import (CellList, Neighbors, NeighborlistArg, _validate_inputs, discard_outside_cutoff,
                        effective_periodic_cell, narrow_down, supercell_for_thin_cell)
from .nn import (ANINetworks, ATOMIC_NUMBER, AtomicContainer, AtomicNetwork, Ensemble, SpeciesConverter,
                 SpeciesEnergies)

SYMBOLS_1X = ("H", "C", "N", "O")
SYMBOLS_2X = ("H", "C", "N", "O", "S", "F", "Cl")
# constants.py:88-96 (wb97x-631gd ground-state atomic energies, Hartree)
GSAES_WB97X_631GD = {"H": -0.4993212, "C": -37.8338334, "N": -54.5732825, "O": -75.0424519,
                     "S": -398.0814169, "F": -99.6949007, "Cl": -460.1167008}


class SpeciesForces(tp.NamedTuple):        # tuples.py of the reference
    species: Tensor
    energies: Tensor
    forces: Tensor


class SpeciesEnergiesQBC(tp.NamedTuple):
    species: Tensor
    energies: Tensor
    qbcs: Tensor


class AtomicStdev(tp.NamedTuple):
    species: Tensor
    energies: Tensor
    stdev_atomic_energies: Tensor


class ForceMagnitudes(tp.NamedTuple):
    species: Tensor
    magnitudes: Tensor


class ForceStdev(tp.NamedTuple):
    species: Tensor
    magnitudes: Tensor
    relative_stdev: Tensor
    relative_range: Tensor


class SelfEnergy(torch.nn.Module):
    """Constant per-element energies (sae.py:16-64).  The fused engine adds them in float64."""

    def __init__(self, symbols: tp.Sequence[str], self_energies: tp.Sequence[float]):
        super().__init__()
        if len(symbols) != len(self_energies):
            raise ValueError("one self energy per symbol is required")
        self.symbols = tuple(symbols)
        self.register_buffer("self_energies", torch.tensor(list(self_energies), dtype=torch.float64))
        self._enabled = True

    @classmethod
    def with_gsaes(cls, symbols: tp.Sequence[str], functional: str = "wb97x", basis_set: str = "631gd"):
        if (functional.lower(), basis_set.lower()) != ("wb97x", "631gd"):
            raise ValueError("only the wb97x-631gd GSAEs are bundled")
        return cls(symbols, [GSAES_WB97X_631GD[s] for s in symbols])

    def forward(self, elem_idxs: Tensor, atomic: bool = False) -> Tensor:
        e = self.self_energies.to(elem_idxs.device)[elem_idxs.clamp(min=0)].masked_fill(elem_idxs == -1, 0.0)
        return e if atomic else e.sum(dim=-1)


class _FusedEnergy(torch.autograd.Function):
    """coords -> (energies (C,), atomic NN energies (C, A), member atomic (M, C, A)); the
    gradient of the ensemble-mean energy is produced by the forward launch sequence."""

    @staticmethod
    def forward(ctx, coords: Tensor, species: Tensor, cell: tp.Optional[Tensor], pbc: bool, engine: Engine,
                want_grad: bool):
        res: StepResult = engine.step(species, coords.detach(), cell, pbc, want_grad=want_grad, check=True)
        ctx.want_grad = want_grad
        if want_grad:
            ctx.save_for_backward(res.grad.clone())
        ctx.mark_non_differentiable(res.member_atomic)
        return res.energies.clone(), res.atomic_energies.clone(), res.member_atomic.clone()

    @staticmethod
    def backward(ctx, g_energy, g_atomic, g_member):
        if not ctx.want_grad:
            raise RuntimeError("energies were computed without forces (coords.requires_grad was False)")
        (grad,) = ctx.saved_tensors
        out = None
        if g_energy is not None:
            out = grad * g_energy.to(grad.dtype).view(-1, 1, 1)
        if g_atomic is not None and bool((g_atomic != 0).any()):
            raise NotImplementedError("per-atom upstream gradients are not supported by the fused engine; "
                                      "use AEVComputer + Ensemble modules directly")
        return out, None, None, None, None, None


class ANI(torch.nn.Module):
    r"""ANI-style neural network interatomic potential (interface of arch.py:90-381)."""

    def __init__(self, symbols: tp.Sequence[str], aev_computer: AEVComputer, neural_networks: AtomicContainer,
                 energy_shifter: SelfEnergy, periodic_table_index: bool = True):
        super().__init__()
        self.register_buffer("atomic_numbers", torch.tensor([ATOMIC_NUMBER[e] for e in symbols], dtype=torch.long))
        assert len(energy_shifter.self_energies) == len(symbols)
        assert aev_computer.num_species == len(symbols)
        assert neural_networks.num_species == len(symbols)
        self.symbols = tuple(symbols)
        self.aev_computer = aev_computer
        self.neural_networks = neural_networks
        self.neighborlist = aev_computer.neighborlist
        self.energy_shifter = energy_shifter
        self.species_converter = SpeciesConverter(symbols)
        self.cutoff = aev_computer.radial.cutoff
        self.periodic_table_index = periodic_table_index
        self._engine: tp.Optional[Engine] = None
        self._engine_key: tp.Any = None

    # -- reference conveniences (arch.py:132-146,253-275) ----------------------------------
    def set_active_members(self, idxs: tp.List[int]) -> None:
        self.neural_networks.set_active_members(idxs)

    def set_strategy(self, strategy: str) -> None:
        self.aev_computer.set_strategy(strategy)

    def __len__(self) -> int:
        return self.neural_networks.get_active_members_num()

    def __getitem__(self, idx: int) -> "ANI":
        import copy
        nets = self.neural_networks
        if not isinstance(nets, Ensemble):
            raise ValueError("Only ensembles can be indexed")
        return ANI(self.symbols, copy.deepcopy(self.aev_computer), copy.deepcopy(nets[idx]),
                   copy.deepcopy(self.energy_shifter), self.periodic_table_index)

    def to_infer_model(self, use_mnp: bool = False) -> "ANI":
        return self

    def ase(self, overwrite: bool = False, stress_kind: str = "scaling", jit: bool = False, skin: float = 0.0):
        """An ASE ``Calculator`` that uses this model (arch.py:219-241).  ``jit`` is accepted for interface
        compatibility and ignored (the hot path is a CUDA graph already)."""
        from .ase import Calculator
        return Calculator(self, overwrite=overwrite, stress_kind=stress_kind, skin=skin)

    # -- fused engine ----------------------------------------------------------------------
    def engine(self, device: torch.device) -> Engine:
        # (the layer-1 operands are packed in the AEV kernels' internal column order: 32-aligned angular block)
        c = self.aev_computer.constants
        self.neural_networks._radial_len = c.num_species * len(c.shf_r)
        nets = self.neural_networks.packed(device)
        shifter = self.energy_shifter
        # the engine bakes the self energies in: key it on their values (edits, `_enabled` toggles)
        key = (str(device), id(nets), bool(shifter._enabled), shifter.self_energies._version,
               shifter.self_energies.data_ptr())
        if self._engine is None or self._engine_key != key:
            sae = self.energy_shifter.self_energies.tolist() if self.energy_shifter._enabled else None
            self._engine = Engine(self.aev_computer.constants, nets, sae, nbr_cap=self.aev_computer.nbr_cap)
            self._engine_key = key
        return self._engine

    def _guarded(self, fn: tp.Callable[[], tp.Any]) -> tp.Any:
        """Run an engine call; if a value left the range of the 2 x fp16 operand pieces
        (ANI_STATUS_OPERAND_RANGE), switch the networks to the 3 x bfloat16 build of the library -- same kernels,
        6 instead of 4 bytes per operand element, no range limit -- and redo the call."""
        try:
            return fn()
        except _lib.OperandRangeError:
            nets = self.neural_networks
            if nets._variant == "bf16x3" or not _lib.available("bf16x3"):
                raise
            import warnings
            warnings.warn("torchani_b200: operand range of the 2 x fp16 GEMM format exceeded; switching this model "
                          "to the 3 x bfloat16 build of the library")
            nets.use_variant("bf16x3")
            return fn()

    @staticmethod
    def _check_inputs(elem_idxs: Tensor, coords: Tensor, charge: int = 0) -> None:
        assert elem_idxs.dim() == 2
        assert coords.shape == (elem_idxs.shape[0], elem_idxs.shape[1], 3)
        assert charge == 0, "Model only supports neutral molecules"

    def forward(self, species_coordinates: tp.Tuple[Tensor, Tensor], cell: tp.Optional[Tensor] = None,
                pbc: tp.Optional[Tensor] = None, charge: int = 0, atomic: bool = False,
                ensemble_values: bool = False) -> SpeciesEnergies:
        species, coords = species_coordinates
        self._check_inputs(species, coords, charge)
        elem_idxs = self.species_converter(species, nop=not self.periodic_table_index)
        _validate_inputs(self.cutoff, elem_idxs, coords, cell, pbc)
        if pbc is not None and elem_idxs.shape[0] > 1:
            # periodic batch (C conformers in one shared cell): one engine step per conformer
            parts = [self((species[c:c + 1], coords[c:c + 1]), cell, pbc, charge, atomic, ensemble_values).energies
                     for c in range(elem_idxs.shape[0])]
            return SpeciesEnergies(elem_idxs, torch.cat(parts, dim=1 if ensemble_values else 0))
        # (with a cell-list neighbour list the reference raises "Cell is too small" instead, neighbors.py:402-403: so
        # does the engine; all_pairs / adaptive models take the supercell route)
        sup = None if isinstance(self.neighborlist, CellList) else \
            supercell_for_thin_cell(species, coords, cell, pbc, self.cutoff)
        if sup is not None:
            # periodic cell thinner than the cutoff (neighbors.py:245-275): evaluate a supercell of R translation-
            # equivalent copies; atomic energies of the original atoms are the first A, the energy is 1/R of the whole
            sp_rep, co_rep, cell_rep, R = sup
            a = elem_idxs.shape[1]
            if atomic:
                out = self((sp_rep, co_rep), cell_rep, pbc, charge, True, ensemble_values).energies[..., :a]
            else:
                out = self((sp_rep, co_rep), cell_rep, pbc, charge, False, ensemble_values).energies / R
            return SpeciesEnergies(elem_idxs, out)
        cell = effective_periodic_cell(coords, cell, pbc, self.cutoff)   # PBC in some directions only
        e, e_atomic, e_member = self._guarded(lambda: _FusedEnergy.apply(
            coords, elem_idxs, cell, pbc is not None, self.engine(coords.device), bool(coords.requires_grad)))
        energies: Tensor
        if ensemble_values:
            active = self.neural_networks.active_members_idxs
            energies = e_member[active].to(coords.dtype)
            if self.energy_shifter._enabled:
                energies = energies + self.energy_shifter(elem_idxs, atomic=True).to(coords.dtype)
            if not atomic:
                energies = energies.sum(-1)
        elif atomic:
            energies = e_atomic.to(coords.dtype)
            if self.energy_shifter._enabled:
                energies = energies + self.energy_shifter(elem_idxs, atomic=True).to(coords.dtype)
        else:
            energies = e.to(coords.dtype)
        return SpeciesEnergies(elem_idxs, energies)

    # -- entry point with a caller-supplied neighbour list (arch.py:354-381, potentials/nnp.py:20-32) ----
    def compute_from_neighbors(self, elem_idxs: Tensor, coords: Tensor, neighbors: Neighbors, charge: int = 0,
                               atomic: bool = False, ensemble_values: bool = False) -> Tensor:
        """Energies from element indices, coordinates and the result of a neighbour-list calculation
        (pairs beyond the radial cutoff are discarded first, as ``discard_outside_cutoff`` does for
        every potential of the reference).  Shape: (C,), (C, A), (M, C) or (M, C, A)."""
        self._check_inputs(elem_idxs, coords, charge)
        neighbors = discard_outside_cutoff(neighbors, self.cutoff)
        aevs = self.aev_computer.compute_from_neighbors(elem_idxs, coords, neighbors)
        energies = self.neural_networks(elem_idxs, aevs, atomic, ensemble_values)
        if self.energy_shifter._enabled:
            energies = energies + self.energy_shifter(elem_idxs, atomic=atomic).to(energies.dtype)
        return energies

    def compute_from_external_neighbors(self, species: Tensor, coords: Tensor, neighbor_idxs: Tensor,
                                        shifts: tp.Optional[Tensor], charge: int = 0, atomic: bool = False,
                                        ensemble_values: bool = False) -> Tensor:
        r"""Entry point for an EXTERNAL neighbour list (arch.py:171-206), e.g. the Verlet list of an MD engine:
        ``neighbor_idxs (2, P)`` candidate pairs (possibly built with a skin) and their lattice ``shifts (P, 3)``
        (None without PBC); pairs beyond the cutoff and pairs with dummy atoms are screened out first
        (``narrow_down``).  IMPORTANT, as in the reference: coords must be mapped to the central cell."""
        self._check_inputs(species, coords, charge)
        elem_idxs = self.species_converter(species, nop=not self.periodic_table_index)
        neighbors = narrow_down(self.cutoff, elem_idxs, coords, neighbor_idxs, shifts)
        return self.compute_from_neighbors(elem_idxs, coords, neighbors, charge, atomic, ensemble_values)

    def compute_from_full_neighborlist(self, species: Tensor, coords: Tensor, ilist_unique: Tensor, jlist: Tensor,
                                       numneigh: Tensor, atomic: bool = False, ensemble_values: bool = False) -> Tensor:
        r"""Energies of the LOCAL atoms (``ilist_unique``) of a local + ghost atom set with the full neighbour list
        of an MD engine (LAMMPS ``ilist / jlist / numneigh``; csrc/cuaev.cpp:225-246 + the network pass of
        potentials/nnp.py:20-32).  Ghost atoms contribute as neighbours only: their energies are masked out, and
        ``torch.autograd.grad`` of the result gives dE/dx for local and ghost coordinates alike."""
        self._check_inputs(species, coords, 0)
        elem_idxs = self.species_converter(species, nop=not self.periodic_table_index)
        aevs = self.aev_computer.compute_from_full_neighborlist(elem_idxs, coords, ilist_unique, jlist, numneigh)
        local = torch.zeros_like(elem_idxs, dtype=torch.bool).view(-1)
        local[ilist_unique.long()] = True
        masked = torch.where(local.view_as(elem_idxs), elem_idxs, torch.full_like(elem_idxs, -1))
        energies = self.neural_networks(masked, aevs, atomic, ensemble_values)
        if self.energy_shifter._enabled:
            energies = energies + self.energy_shifter(masked, atomic=atomic).to(energies.dtype)
        return energies

    # -- ensemble statistics (arch.py:385-576; query-by-committee active learning) ---------------------
    def atomic_energies(self, species_coordinates: tp.Tuple[Tensor, Tensor], cell: tp.Optional[Tensor] = None,
                        pbc: tp.Optional[Tensor] = None, charge: int = 0,
                        ensemble_values: bool = False) -> SpeciesEnergies:
        return self(species_coordinates, cell, pbc, charge, True, ensemble_values)

    def members_forces(self, species_coordinates: tp.Tuple[Tensor, Tensor], cell: tp.Optional[Tensor] = None,
                       pbc: tp.Optional[Tensor] = None, charge: int = 0) -> SpeciesForces:
        """Energies (M, C) and forces (M, C, A, 3) of every active ensemble member (arch.py:403-436).
        The reference differentiates each member's energy by autograd; here ONE pass of the engine
        (``Engine.step_members``): the layer-1 backward GEMM writes one dE_m/dAEV slab per member instead of their
        sum, and the force kernel runs once per slab."""
        species, coords = species_coordinates
        self._check_inputs(species, coords, charge)
        elem_idxs = self.species_converter(species, nop=not self.periodic_table_index)
        _validate_inputs(self.cutoff, elem_idxs, coords, cell, pbc)
        if pbc is not None and elem_idxs.shape[0] > 1:
            parts = [self.members_forces((species[c:c + 1], coords[c:c + 1]), cell, pbc, charge)
                     for c in range(elem_idxs.shape[0])]
            return SpeciesForces(elem_idxs, torch.cat([p.energies for p in parts], 1), torch.cat([p.forces for p in parts], 1))
        cell = effective_periodic_cell(coords, cell, pbc, self.cutoff)

        def run():
            eng = self.engine(coords.device)
            out = eng.step_members(elem_idxs, coords.detach(), cell, pbc is not None)
            eng.check_status()
            return out
        e_m, g_m = self._guarded(run)
        return SpeciesForces(elem_idxs, e_m.to(coords.dtype), -g_m)

    def energies_qbcs(self, species_coordinates: tp.Tuple[Tensor, Tensor], cell: tp.Optional[Tensor] = None,
                      pbc: tp.Optional[Tensor] = None, unbiased: bool = True, charge: int = 0) -> SpeciesEnergiesQBC:
        """Mean energies and query-by-committee factors std_m(E_m) / sqrt(num_atoms) (arch.py:439-485)."""
        elem_idxs, energies = self(species_coordinates, cell, pbc, charge, False, True)
        if energies.shape[0] == 1:
            qbc = torch.zeros_like(energies).squeeze(0)
        else:
            qbc = energies.std(0, unbiased=unbiased)
        qbc = qbc / (elem_idxs >= 0).sum(dim=1, dtype=energies.dtype).sqrt()
        return SpeciesEnergiesQBC(elem_idxs, energies.mean(dim=0), qbc)

    def atomic_stdev(self, species_coordinates: tp.Tuple[Tensor, Tensor], cell: tp.Optional[Tensor] = None,
                     pbc: tp.Optional[Tensor] = None, charge: int = 0, ensemble_values: bool = False,
                     unbiased: bool = True) -> AtomicStdev:
        elem_idxs, energies = self(species_coordinates, cell, pbc, charge, True, True)
        stdev = torch.zeros_like(energies).squeeze(0) if energies.shape[0] == 1 else energies.std(0, unbiased=unbiased)
        if not ensemble_values:
            energies = energies.mean(0)
        return AtomicStdev(elem_idxs, energies, stdev)

    def force_magnitudes(self, species_coordinates: tp.Tuple[Tensor, Tensor], cell: tp.Optional[Tensor] = None,
                         pbc: tp.Optional[Tensor] = None, ensemble_values: bool = False) -> ForceMagnitudes:
        species, _, members_forces = self.members_forces(species_coordinates, cell, pbc)
        magnitudes = members_forces.norm(dim=-1)
        if not ensemble_values:
            magnitudes = magnitudes.mean(0)
        return ForceMagnitudes(species, magnitudes)

    def force_qbc(self, species_coordinates: tp.Tuple[Tensor, Tensor], cell: tp.Optional[Tensor] = None,
                  pbc: tp.Optional[Tensor] = None, ensemble_values: bool = False, unbiased: bool = True) -> ForceStdev:
        """Mean force magnitudes, relative std and relative range across the ensemble (arch.py:543-576)."""
        species, mags = self.force_magnitudes(species_coordinates, cell, pbc, True)
        eps = 1e-8
        mean_mags = mags.mean(0)
        if mags.shape[0] == 1:
            relative_std = torch.zeros_like(mags).squeeze(0)
            relative_range = torch.ones_like(mags).squeeze(0)
        else:
            relative_std = (mags.std(0, unbiased=unbiased) + eps) / (mean_mags + eps)
            relative_range = ((mags.max(dim=0).values - mags.min(dim=0).values) + eps) / (mean_mags + eps)
        if not ensemble_values:
            mags = mean_mags
        return ForceStdev(species, mags, relative_std, relative_range)

    def energies_f64(self, species_coordinates: tp.Tuple[Tensor, Tensor], cell: tp.Optional[Tensor] = None,
                     pbc: tp.Optional[Tensor] = None) -> Tensor:
        """Total energies in float64 as accumulated on the device (the float32 tensor returned by
        ``forward`` loses ~3e-3 Ha at |E| ~ 2.5e4 Ha)."""
        species, coords = species_coordinates
        elem_idxs = self.species_converter(species, nop=not self.periodic_table_index)
        cell = effective_periodic_cell(coords, cell, pbc, self.cutoff)
        return self._guarded(lambda: self.engine(coords.device).step(
            elem_idxs, coords.detach(), cell, pbc is not None, want_grad=False, check=True).energies.clone())

    def energies_and_forces(self, species: Tensor, coords: Tensor, cell: tp.Optional[Tensor] = None,
                            pbc: tp.Optional[Tensor] = None) -> tp.Tuple[Tensor, Tensor]:
        """grad.py:263-290 without the autograd round trip: (energies f64 (C,), forces f32 (C,A,3))."""
        elem_idxs = self.species_converter(species, nop=not self.periodic_table_index)
        cell = effective_periodic_cell(coords, cell, pbc, self.cutoff)
        res = self._guarded(lambda: self.engine(coords.device).step(
            elem_idxs, coords.detach(), cell, pbc is not None, want_grad=True, check=True))
        return res.energies.clone(), -res.grad

    def energies_forces_stress(self, species: Tensor, coords: Tensor, cell: Tensor,
                               pbc: tp.Optional[Tensor] = None) -> tp.Tuple[Tensor, Tensor, Tensor]:
        """(energies f64 (1,), forces f32 (1, A, 3), stress f64 (3, 3) in Hartree/A^3) of one periodic
        system.  The stress is the "f dot r" virial of ase.py:164-168 -- sum over the pairs of
        (dE/dDelta)_a Delta_b, accumulated by the force kernel -- divided by the cell volume; it equals
        the strain derivative dE/d(scaling)/V of ase.py:170-173 and, unlike that one, does not need the
        atoms to be wrapped into the cell."""
        elem_idxs = self.species_converter(species, nop=not self.periodic_table_index)
        res = self._guarded(lambda: self.engine(coords.device).step(
            elem_idxs, coords.detach(), cell, True, want_grad=True, want_virial=True, check=True))
        volume = torch.det(cell.detach().double()).abs()
        return res.energies.clone(), -res.grad, res.virial / volume

    # -- state dicts of the reference (arch.py:278-290) ------------------------------------
    def _load_from_state_dict(self, state_dict, prefix, *args, **kwargs) -> None:
        for old in list(state_dict.keys()):
            new = old.replace("potentials.nnp.aev_computer.", "aev_computer.")
            new = new.replace("potentials.nnp.neural_networks.", "neural_networks.")
            new = new.replace("radial_terms", "radial").replace("angular_terms", "angular")
            new = new.replace(".EtaR", ".eta").replace(".ShfR", ".shifts").replace(".EtaA", ".eta")
            new = new.replace(".Zeta", ".zeta").replace(".ShfA", ".shifts").replace(".ShfZ", ".sections")
            if new != old:
                state_dict[new] = state_dict.pop(old).reshape(-1) if any(
                    new.endswith(k) for k in (".eta", ".zeta", ".shifts", ".sections")) else state_dict.pop(old)
        super()._load_from_state_dict(state_dict, prefix, *args, **kwargs)


def _assemble(symbols, aev_ctor, net_ctor, ensemble_size: int, neighborlist: NeighborlistArg,
              periodic_table_index: bool, seed: tp.Optional[int], device) -> ANI:
    if seed is not None:
        torch.manual_seed(seed)
    aevc = aev_ctor(num_species=len(symbols), neighborlist=neighborlist)
    members = [net_ctor(symbols, aevc.out_dim) for _ in range(ensemble_size)]
    nets: AtomicContainer = Ensemble(members) if ensemble_size > 1 else members[0]
    model = ANI(symbols, aevc, nets, SelfEnergy.with_gsaes(symbols), periodic_table_index)
    model.requires_grad_(False)
    return model.to(device) if device is not None else model


def ANI2x(model_index: tp.Optional[int] = None, neighborlist: NeighborlistArg = "cell_list",
          periodic_table_index: bool = True, device=None, seed: tp.Optional[int] = 0) -> ANI:
    """ANI-2x architecture (models.py:165-198): 7 elements, 1008-dim AEV, 8-member ensemble.
    Weights are seeded random (no download in this environment)."""
    model = _assemble(SYMBOLS_2X, AEVComputer.like_2x, ANINetworks.like_2x, 8, neighborlist,
                      periodic_table_index, seed, device)
    return model if model_index is None else model[model_index]


def ANI1x(model_index: tp.Optional[int] = None, neighborlist: NeighborlistArg = "cell_list",
          periodic_table_index: bool = True, device=None, seed: tp.Optional[int] = 0) -> ANI:
    """ANI-1x architecture (models.py:105-131): 4 elements, 384-dim AEV, 8-member ensemble."""
    model = _assemble(SYMBOLS_1X, AEVComputer.like_1x, ANINetworks.like_1x, 8, neighborlist,
                      periodic_table_index, seed, device)
    return model if model_index is None else model[model_index]


def from_weight_lists(kind: str, weights, device=None, neighborlist: NeighborlistArg = "cell_list",
                      periodic_table_index: bool = False) -> ANI:
    """Build an ANI-1x / ANI-2x shaped model from ``weights[member][symbol] = [(W, b) x 4]``."""
    symbols = SYMBOLS_2X if kind == "2x" else SYMBOLS_1X
    aev_ctor = AEVComputer.like_2x if kind == "2x" else AEVComputer.like_1x
    net_ctor = ANINetworks.like_2x if kind == "2x" else ANINetworks.like_1x
    model = _assemble(symbols, aev_ctor, net_ctor, len(weights), neighborlist, periodic_table_index, None, None)
    members = model.neural_networks.member_networks()
    with torch.no_grad():
        for net, w_m in zip(members, weights):
            for s in symbols:
                for (w_dst, b_dst), (w, b) in zip(net.atomics[s].linear_pairs(), w_m[s]):
                    w_dst.copy_(w)
                    b_dst.copy_(b)
    return model.to(device) if device is not None else model


def aev_computer_from_torchani(aevr, neighborlist: NeighborlistArg = "cell_list") -> AEVComputer:
    """``torchani.aev.AEVComputer`` (standard ANIRadial / ANIAngular terms, cosine or smooth cutoff:
    what its own ``_check_cuaev_avail`` accepts, aev/_computer.py:151-168) -> the B200 module with the
    same constants."""
    cutoff_fn = getattr(aevr, "_cuaev_cutoff_fn", "") or getattr(aevr.radial.cutoff_fn, "_cuaev_name", "") or ""
    if cutoff_fn not in ("cosine", "smooth"):
        raise ValueError("only the cosine and smooth(order=2, eps=1e-10) cutoffs are supported")
    return AEVComputer.from_constants(
        aevr.radial.cutoff, aevr.angular.cutoff, float(aevr.radial.eta), aevr.radial.shifts.tolist(),
        float(aevr.angular.eta), float(aevr.angular.zeta), aevr.angular.shifts.tolist(),
        aevr.angular.sections.tolist(), aevr.num_species, cutoff_fn=cutoff_fn, neighborlist=neighborlist)


def networks_from_torchani(netr) -> AtomicContainer:
    """``torchani.nn.Ensemble`` / ``ANINetworks`` / ``BmmEnsemble`` (nn/_containers.py:319-660,
    nn/_infer.py:61-216) -> the B200 container with the same weights and active members."""
    first = next(iter(netr.atomics.values())) if hasattr(netr, "atomics") else None
    # (BmmEnsemble does not carry atomic_numbers: the keys of `atomics` are the symbols everywhere)
    symbols = tuple(netr.atomics.keys()) if first is not None else tuple(netr.members[0].atomics.keys())
    per_member: tp.List[tp.Dict[str, tp.List[tp.Tuple[Tensor, tp.Optional[Tensor]]]]] = []
    if hasattr(netr, "members"):
        for rm in netr.members:
            per_member.append({s: [(l.weight, l.bias) for l in list(rm.atomics[s].layers) + [rm.atomics[s].final_layer]]
                               for s in symbols})
    elif first is not None and hasattr(first, "_num_batched_networks"):
        # BmmAtomicNetwork: weights stacked as (members, in, out), biases (members, 1, out)
        for m in range(int(first._num_batched_networks)):
            per_member.append({s: [(l.weight[m].t(), l.bias[m, 0] if l._beta else None)
                                   for l in list(netr.atomics[s].layers) + [netr.atomics[s].final_layer]]
                               for s in symbols})
    else:
        per_member.append({s: [(l.weight, l.bias) for l in list(netr.atomics[s].layers) + [netr.atomics[s].final_layer]]
                           for s in symbols})
    members = []
    for layers in per_member:
        mods = {}
        for s in symbols:
            dims = [layers[s][0][0].shape[1]] + [w.shape[0] for w, _ in layers[s]]
            an = AtomicNetwork(dims)
            with torch.no_grad():
                for (w_dst, b_dst), (w, b) in zip(an.linear_pairs(), layers[s]):
                    if b is None:
                        raise ValueError("bias-free networks are not supported")
                    w_dst.copy_(w)
                    b_dst.copy_(b)
            mods[s] = an
        members.append(ANINetworks(mods))
    nets: AtomicContainer = Ensemble(members) if len(members) > 1 else members[0]
    if hasattr(netr, "members"):
        nets.set_active_members(list(netr.active_members_idxs))
    nets.requires_grad_(False)
    return nets


def from_torchani(ref_model, device=None) -> ANI:
    """Convert a ``torchani.arch.ANI`` instance (reference, e.g. ``torchani.models.ANI2x()``) into
    the B200 model: same symbols, AEV constants, network weights and self energies."""
    aevr = ref_model.potentials["nnp"].aev_computer
    netr = ref_model.potentials["nnp"].neural_networks
    symbols = tuple(ref_model.symbols)
    aevc = aev_computer_from_torchani(aevr)
    nets = networks_from_torchani(netr)
    sae = SelfEnergy(symbols, ref_model.energy_shifter.self_energies.double().tolist())
    sae._enabled = bool(ref_model.energy_shifter._enabled)
    model = ANI(symbols, aevc, nets, sae, ref_model.periodic_table_index)
    model.requires_grad_(False)
    if device is None:
        device = ref_model.energy_shifter.self_energies.device
    return model.to(device)


def accelerate_torchani_(ref_model):
    """Swap the B200 modules INTO an existing ``torchani.arch.ANI`` in place -- its three plug-in points
    (arch.py:116-127, 208-217, 264-275): ``model.neighborlist``, ``model.potentials["nnp"].aev_computer`` and
    ``model.potentials["nnp"].neural_networks`` -- and return it.  Everything else of the reference model
    (species converter, energy shifter, extra pair potentials, its ``forward`` / ``compute_from_neighbors``
    / ``ase()``) keeps running unchanged on top of them; CUDA float32 only."""
    nnp = ref_model.potentials["nnp"]
    dev = ref_model.energy_shifter.self_energies.device
    aevc = aev_computer_from_torchani(nnp.aev_computer).to(dev)
    nets = networks_from_torchani(nnp.neural_networks).to(dev)
    nnp.aev_computer = aevc
    nnp.neural_networks = nets
    ref_model.neighborlist = aevc.neighborlist
    return ref_model
