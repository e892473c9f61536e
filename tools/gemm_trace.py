#!/usr/bin/env python
"""Timeline of the tensor-core GEMM roles (timing experiment; needs a B200).

Runs one eager energy+force step of the 10k-atom water box with ani_b200_debug_gemm_trace
switched on and prints, for the first CTA of each of the six GEMM launches, per tile:
  producer : first copy issued, all copies issued
  MMA      : accumulator free, first operands landed, last MMA issued
  epilogue : accumulator ready, tile stored
in microseconds relative to the CTA's first stamp (clock64 / SM clock).
"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from torchani_b200 import _lib, models  # noqa: E402
from torchani_b200.synthetic import DIMS_2X, make_weights, water_box  # noqa: E402


def main():
    dev = torch.device("cuda", 0)
    mhz = float(os.environ.get("SM_MHZ", "1965"))
    z, idx, coords, cell, pbc = water_box(int(os.environ.get("MOLECULES", "3333")), seed=0)
    weights = make_weights(models.SYMBOLS_2X, DIMS_2X, 1008, 8, seed=1234)
    model = models.from_weight_lists("2x", weights, device=dev, periodic_table_index=True)
    eng = model.engine(dev)
    eng.graph_after = 10 ** 9  # eager launches only
    sp_d, co_d, ce_d = idx.to(dev), coords.to(dev), cell.to(dev)
    for _ in range(3):
        eng.step(sp_d, co_d, ce_d, True, True)
    torch.cuda.synchronize()
    launches = 6
    buf = torch.zeros(launches * 4 * 8 * 3 * 4, dtype=torch.int64, device=dev)
    _lib.check(_lib.lib().ani_b200_debug_gemm_trace(buf.data_ptr(), launches), "trace")
    eng.step(sp_d, co_d, ce_d, True, True)
    torch.cuda.synchronize()
    _lib.lib().ani_b200_debug_gemm_trace(None, 0)
    names = ["fwd1", "fwd2", "fwd3+head", "bwd3", "bwd2", "bwd1"]
    if eng.mlp_fused:
        # one data-flow launch (csrc/gemm_fused.cuh): [cta 4][unit 32][role: producer, MMA, epilogue, info][4]
        t = buf[:4 * 32 * 16].view(4, 32, 4, 4).cpu()
        for cta in (0, 3):
            x = t[cta]
            nz = x[:, :3][x[:, :3] > 0]
            if nz.numel() == 0:
                continue
            t0 = int(nz.min())
            us = lambda v: (int(v) - t0) / mhz if int(v) > 0 else float("nan")  # noqa: E731
            print(f"== data-flow launch, cta {cta}: span {(int(nz.max()) - t0) / mhz:.1f} us")
            for u in range(32):
                if int(x[u, :3].max()) == 0:
                    break
                p, m, e, info = x[u, 0], x[u, 1], x[u, 2], x[u, 3]
                print(f"  unit {u:2d} {names[int(info[0])]:9s} rt {int(info[1]):3d} m {int(info[2])} bn {int(info[3]):3d}: "
                      f"prod start {us(p[0]):6.1f} dep-ok {us(p[1]):6.1f} first {us(p[2]):6.1f} done {us(p[3]):6.1f} | "
                      f"mma begin {us(m[0]):6.1f} acc-free {us(m[1]):6.1f} data {us(m[2]):6.1f} issued {us(m[3]):6.1f} | "
                      f"epi begin {us(e[0]):6.1f} start {us(e[1]):6.1f} done {us(e[2]):6.1f} flushed {us(e[3]):6.1f}")
        return
    t = buf.view(launches, 4, 8, 3, 4).cpu()
    for la in range(launches):
        for cta in (0, 3):
            x = t[la, cta]
            nz = x[x > 0]
            if nz.numel() == 0:
                continue
            t0 = int(nz.min())
            us = lambda v: (int(v) - t0) / mhz if int(v) > 0 else float("nan")  # noqa: E731
            print(f"== {names[la]} cta {cta}: span {(int(nz.max()) - t0) / mhz:.1f} us")
            for tile in range(8):
                if int(x[tile].max()) == 0:
                    break
                p, m, e = x[tile, 0], x[tile, 1], x[tile, 2]
                print(f"  tile {tile}: prod start {us(p[0]):6.1f} first {us(p[1]):6.1f} done {us(p[2]):6.1f} | "
                      f"mma begin {us(m[0]):6.1f} acc-free {us(m[1]):6.1f} data {us(m[2]):6.1f} issued {us(m[3]):6.1f} | "
                      f"epi begin {us(e[0]):6.1f} acc-ready {us(e[1]):6.1f} done {us(e[2]):6.1f}")


if __name__ == "__main__":
    main()
