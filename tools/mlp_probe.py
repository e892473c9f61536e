"""Timing probe of the MLP stage under the GEMM debug switches (ANI_B200_GEMM_DEBUG: 2 no copies, 4 no MMA, 8 no
epilogue, 32 epilogue math only (no staging / stores), 128 no stored-activation loads).  Results are garbage when a
switch is set; only the stage time is read.  One process per setting (the switch is read once)."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from torchani_b200 import synthetic  # noqa: E402
from torchani_b200.engine import Engine, PackedNetworks, constants_2x  # noqa: E402


def main():
    dev = torch.device("cuda:0")
    n_mol = int(os.environ.get("MOLECULES", "3333"))
    consts = constants_2x()
    symbols = ("H", "C", "N", "O", "S", "F", "Cl")
    w = synthetic.make_weights(symbols, synthetic.DIMS_2X, 1008, 8, seed=1)
    nets = PackedNetworks([[wm[s] for s in symbols] for wm in w], consts.out_dim, dev)
    _, idx, coords, cell, _ = synthetic.water_box(n_mol, seed=3)
    eng = Engine(consts, nets, None, cuda_graph=False)
    eng.profile = True
    sp, co, ce = idx.to(dev), coords.to(dev), cell.to(dev)
    flush = torch.empty(64 << 20, dtype=torch.float32, device=dev)
    for it in range(25):
        if it == 5:
            eng.stage_events.clear()
        flush.fill_(1.0)
        eng.step(sp, co, ce, True)
    t = eng.stage_times_ms()
    print(f"debug={os.environ.get('ANI_B200_GEMM_DEBUG', '0'):>4s} fused={os.environ.get('ANI_B200_MLP_FUSED', 'auto')} "
          f"mlp {t['mlp_forward_backward'] * 1e3:7.1f} us")


if __name__ == "__main__":
    main()
