"""Phase timeline of the cluster preparation kernel (k_prep_cluster): %globaltimer stamps at its phase boundaries.

    ANI_B200_PREP_TRACE=1 python tools/prep_trace.py [molecules ...]
"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
os.environ["ANI_B200_PREP_TRACE"] = "1"

from torchani_b200 import synthetic  # noqa: E402
from torchani_b200.engine import Engine, PackedNetworks, constants_2x  # noqa: E402

NAMES = ["grid", "zero", "barrier0", "assign", "barrier1", "scan", "scatter+ranges+fills", "barrier2", "order", "barrier3",
         "layout", "rows"]


def main():
    dev = torch.device("cuda:0")
    consts = constants_2x()
    symbols = ("H", "C", "N", "O", "S", "F", "Cl")
    w = synthetic.make_weights(symbols, synthetic.DIMS_2X, 1008, 8, seed=1)
    nets = PackedNetworks([[wm[s] for s in symbols] for wm in w], consts.out_dim, dev)
    for n_mol in [int(a) for a in sys.argv[1:]] or [333, 3333]:
        _, idx, coords, cell, _ = synthetic.water_box(n_mol, seed=3)
        eng = Engine(consts, nets, None, cuda_graph=False)
        sp, co, ce = idx.to(dev), coords.to(dev), cell.to(dev)
        for _ in range(5):
            eng.step(sp, co, ce, True)
        torch.cuda.synchronize()
        ws = eng.workspace(*idx.shape)
        n = ws.n
        n_chunks = max(1, (n + 255) // 256)
        off = 3 * n + 2 + (ws.max_bins + 1) + 2 + (n_chunks + 1) * 8 + 8       # ints before the trace words
        base = ws.scratch.data_ptr() + 4 * off
        pad = (-base) % 8
        t = ws.scratch[off + pad // 4: off + pad // 4 + 24].cpu().view(torch.int64).tolist()
        d = [t[k + 1] - t[k] for k in range(11)]
        print(f"{3 * n_mol} atoms: total {t[11] - t[0]} ns  " + "  ".join(f"{NAMES[k]} {d[k]}" for k in range(11)))


if __name__ == "__main__":
    main()
