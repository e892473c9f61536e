"""Host-buffer calculator: the B200 counterpart of ``torchani.ase.Calculator.calculate``
(ase.py:75-173) without the ``ase`` dependency.

An MD driver owns positions on the HOST (numpy / pinned torch tensors).  Per step the
reference converts numpy -> tensors, copies to the GPU, runs the model, calls ``energy.item()``
and ``forces.cpu().numpy()`` (two blocking D2H syncs).  Here the positions go straight into the
engine's persistent input buffer (one async H2D), the captured CUDA graph of the step is
replayed, energy and forces come back through pinned buffers (async D2H) and ONE stream
synchronisation ends the step.  After a few eager steps the copies themselves are captured too
(the pinned host buffers are persistent): a step is then ONE graph launch + ONE synchronisation.
"""
from __future__ import annotations

import typing as tp

import numpy as np
import torch
from torch import Tensor

from . import _lib
from .models import ANI

HARTREE_TO_EV = 27.211386024367243  # units.py (CODATA 2014, as used by the reference's ASE interface)


class HostCalculator:
    """Energy + forces for a fixed set of atoms from host coordinates.

    Args:
        model: a ``torchani_b200.models.ANI`` on a CUDA device
        atomic_numbers: (A,) or (1, A) integer array (atomic numbers, or element indices if the
            model was built with ``periodic_table_index=False``)
        cell: (3, 3) lattice vectors as rows, or None for a non-periodic system
        pbc: periodic in all three directions (partial periodicity is not supported)
        skin: > 0 switches on Verlet-skin reuse of the bucket grid (the counterpart of
            ``torchani.neighbors.VerletCellList``, neighbors.py:759-884): the grid is built for
            cutoff + skin and reused -- only the positions inside it are refreshed -- until an atom has
            moved more than 70 % of skin/2 from where it was binned; a step in which an atom left its
            skin/2 sphere is recomputed after a rebuild, so the results are always those of a fresh
            grid.  Pays off when consecutive calls move the atoms by a small fraction of ``skin``.
    """

    def __init__(self, model: ANI, atomic_numbers, cell=None, pbc: bool = False, shard: tp.Tuple[int, int] = (0, 1),
                 skin: float = 0.0, sharded: tp.Any = None):
        dev = next(model.buffers()).device
        if dev.type != "cuda":
            raise ValueError("HostCalculator needs a model on a CUDA device")
        self.model, self.device, self.pbc, self.shard = model, dev, bool(pbc), shard
        z = torch.as_tensor(np.asarray(atomic_numbers)).reshape(1, -1).to(dev)
        self.elem_idxs = model.species_converter(z, nop=not model.periodic_table_index)   # validated once
        self.n = z.shape[1]
        self.engine = model.engine(dev)
        self.skin = float(skin)
        if self.skin < 0:
            raise ValueError("skin must be >= 0")
        if self.skin > 0:
            self.engine.skin = self.skin   # buckets for cutoff + skin (shared by every user of this engine)
        # multi-GPU (one process per GPU): `sharded` = parallel.ShardedEngine over this model's engine.  This rank
        # evaluates its slice of the atoms; the partial forces / energies are summed over the ranks ON THE DEVICE
        # (peer-memory reduction inside the captured graph) before the single D2H copy, so every rank returns the
        # full result
        self.sharded = sharded
        if sharded is not None:
            if sharded.engine is not self.engine:
                raise ValueError("`sharded` must wrap this model's engine")
            if self.skin > 0:
                raise ValueError("Verlet-skin reuse is not combined with multi-GPU sharding")
            self.shard = shard = (sharded.rank, sharded.world)
            self.ws = sharded.attach(1, self.n)
        else:
            self.ws = self.engine.workspace(1, self.n)
        self.ws.species_i32.copy_(self.elem_idxs.reshape(-1))
        # pinned mirrors of the workspace's two transfer blocks (engine.Workspace.in_block / out_block): ONE H2D copy of
        # [cell | coords] and ONE D2H copy of [energy | status | gradient] per step
        self.h_in = torch.zeros(self.ws.in_block.numel(), dtype=torch.float32).pin_memory()
        self.h_out = torch.zeros(self.ws.out_block.numel(), dtype=torch.uint8).pin_memory()
        self.h_coords = self.h_in[16:].view(self.n, 3)
        self.h_cell = self.h_in[:9]
        off = self.ws._out_grad_off
        self.h_grad = self.h_out[off:].view(torch.float32).view(self.n, 3)
        self.h_energy = self.h_out[:8].view(torch.float64)
        self.h_status = self.h_out[8:12].view(torch.int32)
        self.h_moved = torch.zeros(1, dtype=torch.int32).pin_memory()
        self._have_grid = False       # a grid built with the skin exists and may be reused
        self.rebuilds = 0             # steps that built the grid
        self.redone = 0               # reuse steps that had to be recomputed (an atom left its skin/2 sphere)
        if pbc and cell is None:
            raise ValueError("If pbc is not None, cell should be present")
        if cell is not None:
            self.set_cell(cell)
        self.h2d_bytes = self.h_in.numel() * 4
        self.d2h_bytes = self.h_out.numel()
        self.graph_after = 3          # eager host-driven steps before copies + kernels are captured as one graph
        self._calls = 0
        self._graphs: tp.Dict[bool, torch.cuda.CUDAGraph] = {}   # reuse flag -> captured step
        self._mode_calls = {False: 0, True: 0}
        self._graph_version: tp.Any = None

    def set_cell(self, cell) -> None:
        self.h_cell.copy_(torch.as_tensor(np.asarray(cell, dtype=np.float32)).reshape(-1))
        self._have_grid = False   # a grid belongs to one cell

    def calculate_with_stress(self, positions, cell=None) -> tp.Tuple[float, np.ndarray, np.ndarray]:
        """(energy [Ha], forces (A, 3) [Ha/A], stress (3, 3) [Ha/A^3]) of a periodic system: the "f dot r"
        virial of ase.py:164-168 accumulated by the force kernel, divided by the cell volume.  Runs the
        step eagerly with a freshly built grid (the virial slots are not part of the captured graph)."""
        if not self.pbc:
            raise ValueError("the stress needs a periodic cell")
        if cell is not None:
            self.set_cell(cell)
        if isinstance(positions, Tensor):
            self.h_coords.copy_(positions.reshape(self.n, 3))
        else:
            self.h_coords.numpy()[...] = np.asarray(positions, dtype=np.float32).reshape(self.n, 3)
        with torch.cuda.device(self.device):
            self._copies_in(False)
            res = self.engine.run(self.ws, True, want_grad=True, shard=self.shard, want_virial=True)
            self._copies_out(False)
            virial = res.virial.cpu()
        self._have_grid = False
        if int(self.h_status[0]):
            self.engine.check_status(self.ws)
        volume = abs(float(np.linalg.det(self.h_cell.numpy().reshape(3, 3).astype(np.float64))))
        return float(self.h_energy[0]), np.negative(self.h_grad.numpy()), virial.numpy() / volume

    def calculate(self, positions, cell=None) -> tp.Tuple[float, np.ndarray]:
        """positions: (A, 3) float array on the host (Angstrom).  Returns (energy [Hartree],
        forces (A, 3) [Hartree/Angstrom]) -- multiply by HARTREE_TO_EV for ASE units."""
        if cell is not None:
            self.set_cell(cell)
        if isinstance(positions, Tensor):
            self.h_coords.copy_(positions.reshape(self.n, 3))
        else:
            self.h_coords.numpy()[...] = np.asarray(positions, dtype=np.float32).reshape(self.n, 3)
        eng = self.engine
        self._calls += 1
        if self._graphs and self._graph_version != eng.nets.active_key:
            self._graphs = {}   # the active ensemble members changed: the captured scales are stale
        reuse = self.skin > 0 and self._have_grid
        self._run(reuse)
        if reuse and (int(self.h_moved[0]) & 1):
            # an atom left its skin/2 sphere: the refreshed grid may have missed pairs -> rebuild, redo
            self.redone += 1
            self._run(False)
        elif reuse and int(self.h_moved[0]):
            self._have_grid = False   # close to the limit: rebuild before the next step
        if int(self.h_status[0]):   # came back with the results: raise what the reference raises
            try:
                eng.check_status(self.ws)
            except _lib.OperandRangeError:
                # the 2 x fp16 operand format ran out of range: rebuild on the 3 x bfloat16 build and redo the step
                nets = self.model.neural_networks
                if nets._variant == "bf16x3" or not _lib.available("bf16x3") or self.sharded is not None:
                    raise
                import warnings
                warnings.warn("torchani_b200: operand range of the 2 x fp16 GEMM format exceeded; switching this "
                              "model to the 3 x bfloat16 build of the library")
                nets.use_variant("bf16x3")
                self.engine = self.model.engine(self.device)
                self.engine.skin = self.skin
                self.ws = self.engine.workspace(1, self.n)
                self.ws.species_i32.copy_(self.elem_idxs.reshape(-1))
                self._graphs, self._have_grid = {}, False
                self._mode_calls = {False: 0, True: 0}
                return self.calculate(positions)
        # a fresh array: h_grad is the persistent pinned D2H buffer and is overwritten by the next call
        return float(self.h_energy[0]), np.negative(self.h_grad.numpy())

    def _run(self, reuse: bool) -> None:
        """One step (H2D, kernels, D2H, synchronise); after a few eager uses of a mode its whole sequence
        is captured into one CUDA graph."""
        with torch.cuda.device(self.device):   # the C-ABI launches on the current device's stream
            self._run_on_device(reuse)

    def _run_on_device(self, reuse: bool) -> None:
        ws, eng = self.ws, self.engine
        self._mode_calls[reuse] += 1
        if not reuse:
            self.rebuilds += 1
        graph = self._graphs.get(reuse)
        nccl_fallback = self.sharded is not None and self.sharded.world > 1 and self.sharded.mode == "nccl"
        if nccl_fallback:
            # peer memory could not be mapped: eager step, library all-reduce of the partials on the device
            import torch.distributed as dist
            self._copies_in(reuse)
            eng.run(ws, self.pbc, want_grad=True, shard=self.shard, reuse=reuse)
            dist.all_reduce(ws.grad, group=self.sharded.group)
            dist.all_reduce(ws.energies, group=self.sharded.group)
            self._copies_out(reuse)
            torch.cuda.current_stream(self.device).synchronize()
            return
        if graph is None and eng.cuda_graph and not eng.profile and self._mode_calls[reuse] > self.graph_after:
            rank, world = self.shard
            lo, hi = (self.n * rank) // world, (self.n * (rank + 1)) // world
            eng.note_composition(ws)
            graph = torch.cuda.CUDAGraph()
            with torch.cuda.graph(graph):
                self._copies_in(reuse)
                eng._launch(ws, self.pbc, True, lo, hi, reuse)
                self._copies_out(reuse)
            self._graphs[reuse], self._graph_version = graph, eng.nets.active_key
        if graph is not None:
            graph.replay()
        else:
            self._copies_in(reuse)
            eng.run(ws, self.pbc, want_grad=True, shard=self.shard, reuse=reuse)
            self._copies_out(reuse)
        torch.cuda.current_stream(self.device).synchronize()
        if not reuse:
            self._have_grid = self.skin > 0

    def _copies_in(self, reuse: bool = False) -> None:
        if reuse:
            self.ws.moved.zero_()
        self.ws.in_block.copy_(self.h_in, non_blocking=True)

    def _copies_out(self, reuse: bool = False) -> None:
        self.h_out.copy_(self.ws.out_block, non_blocking=True)
        if reuse:
            self.h_moved.copy_(self.ws.moved, non_blocking=True)

    def check_status(self) -> None:
        self.engine.check_status(self.ws)
