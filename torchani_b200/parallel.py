"""Multi-GPU sharding of one energy+force step (one process per GPU, torch.distributed for the plumbing).

The reference has no distributed runtime at all; the path shards naturally
because AEV rows and atomic energies are independent per central atom:

* every rank holds the full (species, coords, cell) -- 120 KB at 10 k atoms -- and builds the
  same bucket grid (deterministic, so all ranks agree on the bucket-sorted order),
* rank r owns the bucket-sorted atoms ``[n*r/W, n*(r+1)/W)``: it runs neighbour search + AEV +
  MLP + AEV-backward for those central atoms only and scatters dE_owned/dx_j into a
  full-length gradient buffer (contributions land on non-owned neighbours too),
* ONE reduction of ``3N`` float32 + ``C`` float64 values completes the step.  Nothing else is exchanged.

The reduction is this library's own kernel (``csrc/comm.cu``, ``ani_b200_comm_allreduce``): the partial
buffers live in CUDA-IPC memory mapped by every peer over NVLink, and one launch at the tail of the step's
stream -- inside the step's CUDA graph -- does barrier / read-all-peers-and-sum / barrier.  ``torch.distributed``
only carries the 64-byte IPC handles at set-up.  ``allreduce_partials`` (a plain ``dist.all_reduce``) remains as
the host-logic reference of that reduction and as the fallback when peer memory cannot be mapped.
"""
from __future__ import annotations

import ctypes as C
import typing as tp

import torch
import torch.distributed as dist
from torch import Tensor

from . import _lib
from ._lib import check
from .engine import Engine, StepResult


def shard_bounds(n: int, rank: int, world: int) -> tp.Tuple[int, int]:
    """Slice of bucket-sorted positions owned by ``rank`` (same formula as Engine.step)."""
    return (n * rank) // world, (n * (rank + 1)) // world


def pack_partials(grad: Tensor, energies: Tensor) -> Tensor:
    """[grad (3N) | energies (C)] as one float64 message."""
    return torch.cat([grad.reshape(-1).to(torch.float64), energies.reshape(-1).to(torch.float64)])


def unpack_totals(buf: Tensor, n_conf: int, n_per_conf: int) -> tp.Tuple[Tensor, Tensor]:
    n3 = n_conf * n_per_conf * 3
    return buf[:n3].view(n_conf, n_per_conf, 3), buf[n3:n3 + n_conf]


def allreduce_partials(grad: Tensor, energies: Tensor, group=None) -> tp.Tuple[Tensor, Tensor]:
    """Sum the per-rank partial gradients / energies with a single library collective (fallback path and
    the statement the gloo test checks)."""
    buf = pack_partials(grad, energies)
    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(buf, op=dist.ReduceOp.SUM, group=group)
    return unpack_totals(buf, grad.shape[0], grad.shape[1])


class PeerReducer:
    """The partial-sum buffers of one problem shape in peer-mapped memory + the one-launch reduction.

    ``partial_f32`` / ``partial_f64`` are raw device pointers (IPC memory owned by the C library) that the
    step's kernels accumulate into; ``launch(out_f32, out_f64)`` enqueues the reduction on the current stream."""

    def __init__(self, n_f32: int, n_f64: int, device: torch.device, group=None):
        self.device = torch.device(device)
        self.rank, self.world = dist.get_rank(group), dist.get_world_size(group)
        self.lib = _lib.lib()
        handle = C.c_void_p()
        with torch.cuda.device(self.device):
            check(self.lib.ani_b200_comm_create(self.rank, self.world, n_f32, n_f64, C.byref(handle)), "comm_create")
            self.handle = handle
            mine = (C.c_char * 64)()
            check(self.lib.ani_b200_comm_handle(self.handle, mine), "comm_handle")
            gathered: tp.List[tp.Any] = [None] * self.world
            dist.all_gather_object(gathered, bytes(mine), group=group)
            blob = b"".join(gathered)
            check(self.lib.ani_b200_comm_connect(self.handle, blob), "comm_connect")
            p32, p64, perr = C.c_void_p(), C.c_void_p(), C.c_void_p()
            check(self.lib.ani_b200_comm_buffers(self.handle, C.byref(p32), C.byref(p64), C.byref(perr)), "comm_buffers")
        self.partial_f32, self.partial_f64, self.error_ptr = p32.value, p64.value, perr.value
        dist.barrier(group)   # every rank has mapped every peer before the first launch

    def launch(self, out_f32: Tensor, out_f64: Tensor) -> None:
        st = torch.cuda.current_stream(self.device).cuda_stream
        check(self.lib.ani_b200_comm_allreduce(self.handle, out_f32.data_ptr(), out_f64.data_ptr(), st),
              "comm_allreduce")

    def close(self) -> None:
        if getattr(self, "handle", None):
            torch.cuda.synchronize(self.device)
            self.lib.ani_b200_comm_destroy(self.handle)
            self.handle = None


class ShardedEngine:
    """Runs ``Engine.step`` on this rank's slice; the partial results are summed over the ranks on the
    device, inside the step's launch sequence / CUDA graph (``reduce="peer"``), or -- ``reduce="nccl"`` -- by a
    ``dist.all_reduce`` after it."""

    def __init__(self, engine: Engine, group=None, reduce: str = "auto"):
        self.engine = engine
        self.group = group
        if dist.is_available() and dist.is_initialized():
            self.rank, self.world = dist.get_rank(group), dist.get_world_size(group)
        else:
            self.rank, self.world = 0, 1
        if reduce not in ("auto", "peer", "nccl"):
            raise ValueError("reduce must be 'auto', 'peer' or 'nccl'")
        self.reduce = reduce
        self._reducers: tp.Dict[tp.Tuple[int, int], PeerReducer] = {}
        self.mode = "none" if self.world == 1 else ""

    def attach(self, n_conf: int, n_per_conf: int):
        """Workspace of this shape with its gradient / energy accumulators placed in peer-mapped memory."""
        ws = self.engine.workspace(n_conf, n_per_conf)
        if self.world == 1 or ws.reducer is not None or self.mode == "nccl":
            return ws
        want_peer = self.reduce in ("auto", "peer")
        ok = False
        if want_peer:
            try:
                red = PeerReducer(3 * ws.n, n_conf, self.engine.device, self.group)
                ok = True
            except Exception:   # peer memory cannot be mapped here (no P2P / IPC): every rank falls back together
                if self.reduce == "peer":
                    raise
        flag = torch.tensor([1 if ok else 0], device=self.engine.device)
        dist.all_reduce(flag, op=dist.ReduceOp.MIN, group=self.group)
        if int(flag.item()) == 1:
            ws.reducer = red
            ws.grad_ptr, ws.energies_ptr = red.partial_f32, red.partial_f64
            self.mode = "peer"
        else:
            self.mode = "nccl"
        return ws

    def step(self, species: Tensor, coords: Tensor, cell: tp.Optional[Tensor] = None, pbc: bool = False
             ) -> tp.Tuple[Tensor, Tensor]:
        """-> (energies f64 (C,), dE/dcoords (C, A, 3)), identical on every rank."""
        self.attach(species.shape[0], species.shape[1])
        res: StepResult = self.engine.step(species, coords, cell, pbc, want_grad=True,
                                           shard=(self.rank, self.world))
        if self.world == 1 or self.mode == "peer":
            return res.energies, res.grad
        energies, grad = allreduce_partials(res.grad, res.energies, self.group)[::-1]
        return energies, grad
