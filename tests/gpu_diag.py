"""Stage-by-stage GPU diagnostic (run by hand under gpurun; the pytest suite is the gate).

Compares every stage of the fused engine (bucket grid, half neighbour list, AEV forward,
MLP forward/backward, AEV backward) against the CPU oracle on the committed golden cases and
prints max errors, so that a single GPU round trip localises a bug.
"""
from __future__ import annotations

import os
import sys
import time
import ctypes as C

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import oracle.ani_oracle as orc  # noqa: E402
from torchani_b200 import _lib  # noqa: E402
from torchani_b200.engine import Engine, PackedNetworks, constants_1x, constants_2x  # noqa: E402
from torchani_b200._lib import check, ptr  # noqa: E402

GOLD = os.path.join(ROOT, "tests", "golden")
dev = torch.device("cuda:0")


def weights_as_lists(w, symbols):
    return [[w_m[s] for s in symbols] for w_m in w]


def make_engine(kind):
    if kind == "2x":
        model = orc.ani2x_model()
        consts = constants_2x()
    else:
        model = orc.ani1x_model()
        consts = constants_1x()
    nets = PackedNetworks(weights_as_lists(model.weights, model.symbols), consts.out_dim, dev)
    sae = [model.sae[s] for s in model.symbols]
    return model, Engine(consts, nets, sae)


def relerr(a, b, floor):
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    return float(np.max(np.abs(a - b) / (np.abs(b) + floor))) if a.size else 0.0


def run_case(name, engines):
    z = np.load(os.path.join(GOLD, name + ".npz"))
    kind = str(z["kind"])
    model, eng = engines[kind]
    model = model._replace(neighborlist=str(z["neighborlist"]))
    species = torch.tensor(z["species"])
    coords64 = torch.tensor(z["coords"])
    cell64 = torch.tensor(z["cell"]) if z["cell"].size else None
    pbc = torch.tensor(z["pbc"]) if z["pbc"].size else None
    Cn, A = species.shape
    print(f"\n=== {name}: C={Cn} A={A} pbc={pbc is not None} kind={kind}")
    # oracle in fp64 with the fp32-rounded inputs the GPU sees
    coords32 = coords64.float()
    cell32 = None if cell64 is None else cell64.float()
    ref = orc.compute(model, species, coords32.double(), None if cell32 is None else cell32.double(), pbc)
    # oracle gradient wrt the AEV (for the MLP stage)
    aev_ref = ref["aev"].clone().requires_grad_(True)
    e_m = orc.ensemble_atomic_energies(model.symbols, [
        {s: [(w.double(), b.double()) for w, b in lay] for s, lay in wm.items()} for wm in model.weights
    ], species, aev_ref)
    g_aev_ref = torch.autograd.grad(e_m.mean(0).sum(), aev_ref)[0]

    sp_d = species.to(dev)
    co_d = coords32.to(dev)
    ce_d = None if cell32 is None else cell32.to(dev)
    res = eng.step(sp_d, co_d, ce_d, pbc is not None, want_grad=True)
    torch.cuda.synchronize()
    ws = eng.workspace(Cn, A)
    status = int(ws.status.item())
    g = eng.grid_info(ws)
    n_real = int((species >= 0).sum())
    print(f"status={status} grid dims={list(g.dims)} nbins={g.nbins} n_real={g.n_real} (expect {n_real}) "
          f"layout={ws.layout_info.tolist()}")
    so = ws.sorted_orig.cpu().numpy()
    assert sorted(so.tolist()) == list(range(Cn * A)), "sorted_orig is not a permutation"
    o2s = ws.orig_to_sorted.cpu().numpy()
    assert (so[o2s] == np.arange(Cn * A)).all()
    cnt = ws.nbr_cnt.cpu().numpy()
    tot_nbr = int(cnt[:n_real].sum())
    print(f"directed neighbours {tot_nbr} (= 2 x half pairs {tot_nbr // 2}; golden num_pairs {int(z['num_pairs'])}),"
          f" max per atom {cnt[:n_real].max() if n_real else 0}")

    # ---- half neighbour list API kernel
    L = _lib.lib()
    st = torch.cuda.current_stream().cuda_stream
    n = Cn * A
    pair_start = torch.zeros(2 * n + 2, dtype=torch.int32, device=dev)
    check(L.ani_b200_half_neighbor_count(ptr(ws.grid), ptr(ws.bin_start), ptr(ws.spos), ptr(ws.sbin),
                                         ptr(ws.sorted_orig), n, eng.consts.rcr, ptr(pair_start), st))
    P = int(pair_start[n].item())
    idx0 = torch.zeros(max(P, 1), dtype=torch.int64, device=dev)
    idx1 = torch.zeros_like(idx0)
    dist = torch.zeros(max(P, 1), dtype=torch.float32, device=dev)
    diff = torch.zeros(max(P, 1), 3, dtype=torch.float32, device=dev)
    check(L.ani_b200_half_neighbor_fill(ptr(ws.grid), ptr(ws.bin_start), ptr(ws.spos), ptr(ws.sbin),
                                        ptr(ws.sorted_orig), n, eng.consts.rcr, ptr(pair_start), P, ptr(idx0),
                                        ptr(idx1), ptr(dist), ptr(diff), ptr(ws.status), st))
    torch.cuda.synchronize()
    print(f"half list pairs {P} (golden {int(z['num_pairs'])})")
    if "pairs" in z and P == int(z["num_pairs"]):
        mine = np.stack([idx0[:P].cpu().numpy(), idx1[:P].cpu().numpy()])
        d_m = dist[:P].cpu().numpy()
        lo_ = np.minimum(mine[0], mine[1]); hi_ = np.maximum(mine[0], mine[1])
        order = np.lexsort((d_m, hi_, lo_))
        gp = z["pairs"]; gd = z["distances"]
        glo = np.minimum(gp[0], gp[1]); ghi = np.maximum(gp[0], gp[1])
        gorder = np.lexsort((gd, ghi, glo))
        same = (lo_[order] == glo[gorder]).all() and (hi_[order] == ghi[gorder]).all()
        print(f"  pair sets equal: {same}; max |dist - golden| = {np.abs(d_m[order] - gd[gorder]).max():.3e}")

    # ---- AEV forward (x was overwritten by dE/dAEV in the step -> recompute the forward only)
    from torchani_b200.engine import untile_a_operand
    xt = untile_a_operand(ws.x.reshape(-1), ws.rows_cap, eng.nets.ldx)   # AEVs of the step (tiled -> plain)
    row_of = ws.row_of.cpu().numpy()
    x = xt.cpu().numpy()
    D = eng.consts.out_dim
    aev_mine = np.zeros((n, D))
    for i in range(n_real):
        aev_mine[so[i]] = x[row_of[i], :D]
    aev_r = ref["aev"].reshape(n, D).numpy()
    RL = eng.consts.radial_len
    print(f"AEV fwd: max-abs radial {np.abs(aev_mine[:, :RL] - aev_r[:, :RL]).max():.3e} "
          f"angular {np.abs(aev_mine[:, RL:] - aev_r[:, RL:]).max():.3e}; "
          f"max-rel(floor 1e-3) {relerr(aev_mine, aev_r, 1e-3):.3e}; |aev|max {np.abs(aev_r).max():.3f}")
    pad_cols = x[:, D:]
    print(f"  pad columns max {np.abs(pad_cols).max() if pad_cols.size else 0:.1e}")
    # ---- MLP
    em_mine = res.member_atomic.cpu().numpy()
    em_ref = ref["member_atomic"].numpy()
    print(f"MLP: member atomic energies max-abs {np.abs(em_mine - em_ref).max():.3e} "
          f"(|e|max {np.abs(em_ref).max():.3f}); ensemble atomic max-abs "
          f"{np.abs(res.atomic_energies.cpu().numpy() - ref['atomic_nn'].numpy()).max():.3e}")
    e_tot = res.energies.cpu().numpy()
    print(f"  energies (NN+SAE) mine {e_tot[:3]} ref {ref['energy'].numpy()[:3]} "
          f"max-abs {np.abs(e_tot - ref['energy'].numpy()).max():.3e}")
    # dE/dAEV of the step (plain rows)
    gx = ws.dx.cpu().numpy()
    g_mine = np.zeros((n, D))
    for i in range(n_real):
        g_mine[so[i]] = gx[row_of[i], :D]
    g_r = g_aev_ref.reshape(n, D).numpy()
    blocks = ws.aev_blocks.cpu().numpy()
    live = np.zeros(eng.nets.ldx, dtype=bool)
    for b in blocks[1:1 + blocks[0]]:
        live[b * 32:(b + 1) * 32] = True
    live = live[:D]
    print(f"MLP bwd: dE/dAEV max-abs over the {blocks[0]} live column blocks "
          f"{np.abs(g_mine - g_r)[:, live].max():.3e} (|g|max {np.abs(g_r).max():.3e}); live blocks {blocks[1:1 + blocks[0]].tolist()}")
    # ---- forces
    f_mine = -res.grad.cpu().numpy()
    f_ref = ref["forces"].numpy()
    print(f"forces: max-abs {np.abs(f_mine - f_ref).max():.3e} (|F|max {np.abs(f_ref).max():.3e}); "
          f"golden(fp64 inputs) max-abs {np.abs(f_mine - z['forces']).max():.3e}")
    print(f"  net force (should be ~0): {np.abs(f_mine.reshape(Cn, A, 3).sum(1)).max():.2e}")


def timing(engines):
    model, eng = engines["2x"]
    for nmol in (333, 3333):
        _, idx, coords, cell, pbc = orc.water_box(nmol, seed=0)
        sp_d, co_d, ce_d = idx.to(dev), coords.to(dev), cell.to(dev)
        for _ in range(3):
            eng.step(sp_d, co_d, ce_d, True)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        reps = 10
        for _ in range(reps):
            eng.step(sp_d, co_d, ce_d, True)
        torch.cuda.synchronize()
        dt = (time.perf_counter() - t0) / reps
        eng.check_status()
        print(f"water {3 * nmol} atoms: {dt * 1e3:.3f} ms/step, {0.0864 / dt:.2f} ns/day, "
              f"{3 * nmol / dt:.3e} atom-steps/s")


if __name__ == "__main__":
    print(torch.cuda.get_device_name(0))
    engines = {"2x": make_engine("2x"), "1x": make_engine("1x")}
    names = ["kat2x5_ani2x", "ch4_ani1x", "water30_pbc_ani2x", "benzene_pbc_ani2x", "tightcell_pbc_ani2x",
             "randbatch_ani2x", "small264_nopbc_ani2x", "water999_pbc_ani2x", "6w8h_triclinic_ani2x"]
    if len(sys.argv) > 1:
        names = sys.argv[1:]
    for nm in names:
        try:
            run_case(nm, engines)
        except Exception as e:  # keep going: one round trip should report everything
            import traceback
            traceback.print_exc()
            print(f"!!! case {nm} raised {e!r}")
            try:
                torch.cuda.synchronize()
            except Exception as e2:
                print("CUDA context is broken:", e2)
                break
    try:
        timing(engines)
    except Exception:
        import traceback
        traceback.print_exc()
