#!/usr/bin/env python
"""Benchmark of the ANI-2x energy+force hot path (BASELINE.json metric).

    python bench.py --gpus N --steps K --warmup W            # B200 arm (this repo)
    python bench.py --impl reference --steps K --warmup W    # reference arm: CPU port of the
                                                             # reference algorithm (oracle/)

One "step" = energy AND forces of the whole system (neighbour search -> AEV -> 8-member MLP
ensemble -> forces).  Workload = BASELINE.json's metric configuration: ANI-2x x8, periodic
10k-atom water box (9999 atoms, L = 46.38 A), seeded synthetic coordinates and weights.
Metric: atom-steps/s (= N_atoms / t_step; ns/day at 1 fs = 0.0864 / t_step is reported too).
With N > 1 GPUs (torchrun, one rank per GPU) the SAME box is sharded over the ranks (strong
scaling) and one NCCL all-reduce of 3N+1 float64 values closes every step.

Timing: W >= 3 warm-up steps, then K steps each bracketed by CUDA events on the launch stream;
between timed steps a 256 MiB buffer is overwritten (L2 flush, untimed); barrier +
synchronize on both sides of the region; per-step times are summed, max over ranks.
Prints ONE JSON line on rank 0.
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "ani2x_energy_force_atom_steps_per_s"
UNIT = "atom-steps/s"
# algorithmic bytes per atom of the AEV kernels (fp32, ANI-2x)
AEV_FWD_BYTES_PER_ATOM = 4276.0
AEV_BWD_BYTES_PER_ATOM = 4288.0
# dram__bytes_read.sum + dram__bytes_write.sum from the ncu --set full capture of this build
# (profiles/r01_ncu_full_v10_selected.csv), per launch (AEV kernels) / per six-GEMM sequence at the
# 9999-atom box.  ncu flushes the caches before every kernel, so these are COLD-cache figures: inside
# a step the activations a GEMM reads were just written by the previous launch and sit in the L2.
MLP_DRAM_BYTES_PER_STEP = 574.1e6      # reads 494.6 MB + writes 79.5 MB over the six launches
AEV_FWD_DRAM_BYTES = 0.93e6            # reads; its 6.4 MB of live AEV blocks stay in the L2 (write-back)
AEV_BWD_DRAM_BYTES = 14.5e6            # reads (live blocks of dE/dAEV)


def workload(n_molecules: int):
    from torchani_b200.synthetic import water_box
    return water_box(n_molecules, seed=0)


def load_oracle():
    """The CPU port of the reference algorithm: test infrastructure, imported ONLY by the
    cpu_baseline / --impl reference legs as the thing that is timed beside the GPU path."""
    import oracle.ani_oracle as orc
    return orc


def peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        p = json.load(open(path))
        return {"hbm_gbs": p["hbm_gbs"], "bf16_tflops": p["bf16_tflops"], "source": "measured (MEASURED_PEAKS.json)"}
    return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "source": "fallback (B200_PROFILING.md)"}


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled every 200 ms during the timed region."""

    FIELDS = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
              "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--id={self.index}", f"--query-gpu={self.FIELDS}", "--format=csv,noheader,nounits",
                 "-lms", "100"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
            t0 = time.time()
            while not self.rows and time.time() - t0 < 5.0:   # nvidia-smi takes a moment to start sampling
                time.sleep(0.05)
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([v.strip() for v in line.split(",")])

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        sm = [float(r[0]) for r in self.rows if r and r[0].replace(".", "").isdigit()]
        mx = [float(r[1]) for r in self.rows if len(r) > 1 and r[1].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [n for k, n in enumerate(names) if any(len(r) > 2 + k and r[2 + k] == "Active" for r in self.rows)]
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": reasons, "samples": len(sm)}


def cpu_reference_step(orc, model, idx, coords, cell, pbc):
    out = orc.compute(model, idx, coords, cell, pbc, forces=True)
    return out["energy"], out["forces"]


def run_cpu_baseline(idx, coords, cell, pbc, steps: int, warmup: int, budget_s: float = 60.0):
    """The reference algorithm's CPU port (oracle/ani_oracle.py, float32, all host threads).
    Bounded: stops early (after at least one timed evaluation) once `budget_s` seconds are spent,
    warm-up included, so a slow or busy host cannot stretch the run."""
    torch.set_num_threads(os.cpu_count() or 1)  # torchrun pins OMP_NUM_THREADS=1: undo that for the CPU arm
    orc = load_oracle()
    model = orc.ani2x_model(seed=1234, members=8, neighborlist="cell_list")
    t_begin = time.perf_counter()
    for _ in range(warmup):
        cpu_reference_step(orc, model, idx, coords, cell, pbc)
        if time.perf_counter() - t_begin > budget_s / 2:
            break
    times = []
    for _ in range(steps):
        t0 = time.perf_counter()
        cpu_reference_step(orc, model, idx, coords, cell, pbc)
        times.append(time.perf_counter() - t0)
        if time.perf_counter() - t_begin > budget_s:
            break
    return times


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--molecules", type=int, default=3333, help="water molecules (3333 -> the 10k-atom box)")
    ap.add_argument("--cpu-steps", type=int, default=2, help="steps of the cpu_baseline sample (0 = skip)")
    ap.add_argument("--skin", type=float, default=0.0,
                    help="experiment: Verlet skin (A) of the end-to-end arm; the atoms then move ballistically "
                         "(300 K Maxwell-Boltzmann velocities, 1 fs per step) so that grid reuse is exercised")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    warmup = max(args.warmup, 3) if args.impl == "b200" else args.warmup
    n_atoms = 3 * args.molecules
    config = {"workload": f"ANI-2x x8 ensemble, periodic water box, {n_atoms} atoms (BASELINE configs[3]), "
                          "energy+forces per step",
              "atoms": n_atoms, "ensemble": 8, "aev_dim": 1008, "cutoffs_A": [5.1, 3.5],
              "parallelism": f"atom-range sharding x{world} + 1 all-reduce" if world > 1 else "single GPU",
              "l2": "256 MiB flush between timed steps (untimed)"}

    # ------------------------------------------------------------------ reference arm (CPU)
    if args.impl == "reference":
        if rank != 0:
            return
        z, idx, coords, cell, pbc = workload(args.molecules)
        cores = torch.get_num_threads()
        times = run_cpu_baseline(idx, coords, cell, pbc, args.steps, args.warmup, budget_s=150.0)
        t = sum(times) / len(times)
        value = n_atoms / t
        sample = f"{len(times)} full energy+force evaluations of the {n_atoms}-atom box (float32, torch CPU ops)"
        line = {"impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
                "steps": len(times), "warmup": args.warmup, "ms_per_step": t * 1e3, "higher_is_better": True,
                "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
                "ns_per_day": 0.0864 / t, "config": config,
                "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
                "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
                "gpu_launches": 0}
        print(json.dumps(line))
        return

    # ------------------------------------------------------------------ B200 arm
    import torch.distributed as dist
    if os.environ.get("NCCL_DEBUG", "VERSION").upper() == "VERSION":
        os.environ["NCCL_DEBUG"] = "WARN"  # keep stdout to the one JSON line (NCCL prints its version there)
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    from torchani_b200 import models
    from torchani_b200.parallel import ShardedEngine

    from torchani_b200.synthetic import DIMS_2X, make_weights
    z, idx, coords, cell, pbc = workload(args.molecules)
    weights = make_weights(models.SYMBOLS_2X, DIMS_2X, 1008, 8, seed=1234)
    model = models.from_weight_lists("2x", weights, device=dev, periodic_table_index=True)
    eng = model.engine(dev)
    sharded = ShardedEngine(eng)
    sp_d, co_d, ce_d = idx.to(dev), coords.to(dev), cell.to(dev)
    z_d = z.to(dev)
    flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=dev)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    def one_step():
        return sharded.step(sp_d, co_d, ce_d, True)

    sampler = ClockSampler(local_rank)
    sampler.start()  # sampled from the warm-up to the end of the end-to-end region (GPU under load throughout)
    for _ in range(warmup):
        one_step()
    eng.check_status()
    barrier()
    evs = []
    for _ in range(args.steps):
        flush.fill_(1)
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        one_step()
        b.record()
        evs.append((a, b))
    barrier()
    total_ms = sum(a.elapsed_time(b) for a, b in evs)
    t_all = torch.tensor([total_ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t_all, op=dist.ReduceOp.MAX)
    ms_per_step = float(t_all.item()) / args.steps
    value = n_atoms / (ms_per_step * 1e-3)
    eng.check_status()

    # ---- end to end through the public API with HOST buffers: calculator.HostCalculator.calculate
    #      (host positions in, host energy + forces out; H2D, graph replay, D2H and the sync inside)
    from torchani_b200.calculator import HostCalculator
    calc = HostCalculator(model, z[0].numpy(), cell.numpy(), pbc=True, shard=(rank, world), skin=args.skin)
    h_pos = coords[0].numpy().copy()
    h_vel = None
    if args.skin > 0:   # A/fs: sqrt(kT/m), kT = 0.02585 eV, 1 amu A^2/fs^2 = 103.6427 eV
        import numpy as np
        mass = np.where(z[0].numpy() == 1, 1.008, 15.999)[:, None]
        h_vel = (np.random.default_rng(0).normal(size=h_pos.shape) * np.sqrt(0.02585 / (103.6427 * mass))).astype("float32")

    def e2e_step():
        if h_vel is not None:
            h_pos[...] += h_vel
        e, f = calc.calculate(h_pos)
        if world > 1:  # partial results of this rank's atom slice -> one all-reduce (as in ShardedEngine)
            buf = torch.cat([torch.from_numpy(f).reshape(-1).double(), torch.tensor([e], dtype=torch.float64)]).to(dev)
            dist.all_reduce(buf)
            buf.cpu()
        return e, f

    for _ in range(3):
        e2e_step()
    barrier()
    e2e_evs = []
    for _ in range(args.steps):
        flush.fill_(1)
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        e2e_step()
        b.record()
        e2e_evs.append((a, b))
    barrier()
    clocks = sampler.stop()
    e2e_ms = sum(a.elapsed_time(b) for a, b in e2e_evs)
    t_e2e = torch.tensor([e2e_ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t_e2e, op=dist.ReduceOp.MAX)
    e2e_value = n_atoms / (float(t_e2e.item()) / args.steps * 1e-3)
    h2d = calc.h2d_bytes
    d2h = calc.d2h_bytes

    # ---- per-stage device times for the roofline (separate short run, events per C-ABI call)
    eng.profile = True
    for _ in range(5):
        flush.fill_(1)
        one_step()
    stage = eng.stage_times_ms()
    eng.profile = False
    pk = peaks()
    owned = n_atoms / world
    flops = sum(eng.nets.flops_per_atom(3 if k % 3 == 0 else 0) for k in range(3)) / 3.0 * owned  # O,H,H
    mlp_s = stage.get("mlp_forward_backward", float("nan")) * 1e-3
    fwd_s = stage.get("aev_forward", float("nan")) * 1e-3
    bwd_s = stage.get("aev_backward", float("nan")) * 1e-3
    from torchani_b200._lib import operand_format
    fmt = operand_format()
    split = ("2 x fp16 split of every (power-of-two scaled) fp32 operand, 3 products" if fmt.parts == 2
             else "3 x bf16 split of every fp32 operand, 6 products")
    nprod = 3 if fmt.parts == 2 else 6
    roofline = {"kernel": "tc::k_gemm_tc<EPI> x6 (ani_b200_mlp_forward_backward): ensemble MLP fwd + bwd-to-input, "
                          f"tcgen05 kind::f16 on a {split} (fp32-accurate)",
                "bound": "tensor", "achieved": flops / mlp_s / 1e12, "peak": pk["bf16_tflops"], "unit": "TFLOP/s",
                "frac": flops / mlp_s / 1e12 / pk["bf16_tflops"], "traffic": MLP_DRAM_BYTES_PER_STEP if world == 1 and args.molecules == 3333 else None,
                "algorithmic_flops_per_launch_sequence": flops, "peak_source": pk["source"],
                "note": "achieved = dense algorithmic FLOPs of SURVEY 8(d) (98.2 GFLOP/step at 10k atoms) / "
                        "device time of the six GEMM launches; peak = measured dense bf16 rate (= the fp16 rate). The "
                        f"kernel issues {nprod} 16-bit MMAs per product but skips the AEV column blocks of absent "
                        f"element pairs in layer 1, so the executed tensor work is {nprod} x 36% of the dense count "
                        "for water. "
                        "traffic = dram bytes of the six launches (ncu, profiles/), null if not captured for this build"}
    roofline_aev = {
        "forward": {"kernel": "k_aev_forward_cta<8,4>", "bound": "hbm", "achieved": AEV_FWD_BYTES_PER_ATOM * owned / fwd_s / 1e9,
                    "peak": pk["hbm_gbs"], "unit": "GB/s",
                    "frac": AEV_FWD_BYTES_PER_ATOM * owned / fwd_s / 1e9 / pk["hbm_gbs"],
                    "traffic": AEV_FWD_DRAM_BYTES if world == 1 and args.molecules == 3333 else None},
        "backward": {"kernel": "k_aev_backward<8,4>", "bound": "hbm", "achieved": AEV_BWD_BYTES_PER_ATOM * owned / bwd_s / 1e9,
                     "peak": pk["hbm_gbs"], "unit": "GB/s",
                     "frac": AEV_BWD_BYTES_PER_ATOM * owned / bwd_s / 1e9 / pk["hbm_gbs"],
                     "traffic": AEV_BWD_DRAM_BYTES if world == 1 and args.molecules == 3333 else None},
        "peak_source": pk["source"]}

    # ---- CPU baseline beside it (rank 0, N = 1 only): bounded sample of the same workload
    cpu = None
    if rank == 0 and world == 1 and args.cpu_steps > 0:
        cores = torch.get_num_threads()
        times = run_cpu_baseline(idx, coords, cell, pbc, args.cpu_steps, 1, budget_s=45.0)
        t = sum(times) / len(times)
        cpu = {"value": n_atoms / t, "unit": UNIT, "cores": cores, "kind": "port",
               "sample": f"{len(times)} full energy+force evaluations of the same {n_atoms}-atom box with the "
                         f"CPU port of the reference algorithm (oracle/ani_oracle.py, float32, {cores} threads)"}

    if rank == 0:
        line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
                "warmup": warmup, "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "strong",
                "vs_baseline": None, "dtype": "f32", "data": "synthetic", "config": config,
                "ns_per_day": 0.0864 / (ms_per_step * 1e-3), "clocks": clocks,
                "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                        "ms_per_step": float(t_e2e.item()) / args.steps,
                        "api": "torchani_b200.calculator.HostCalculator.calculate (host positions in, host "
                               "energy+forces out; counterpart of torchani.ase.Calculator.calculate)",
                        **({"verlet_skin_A": args.skin, "grid_rebuilds": calc.rebuilds, "steps_redone": calc.redone,
                            "calls": calc._calls} if args.skin > 0 else {})},
                "gpu_launches": eng.launches_per_step * args.steps,
                "operand_format": {"parts": fmt.parts, "bytes_per_element": 2 * fmt.parts},
                "stage_ms": stage, "roofline": roofline, "roofline_aev": roofline_aev, "cpu_baseline": cpu}
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
