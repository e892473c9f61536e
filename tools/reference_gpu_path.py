#!/usr/bin/env python
"""The UNMODIFIED reference's GPU path timed end to end on this box, beside this repo's path.

What the north star is defined against (BASELINE.json: ">= 10x the reference cuAEV+PyTorch GPU path on ANI-2x
energy+force for a 10k-atom periodic water box at 1xB200"; SURVEY.md 8(d) "GPU baseline beside it"):

    torchani.arch.ANI (ANI-2x shaped, 8 members, the seeded synthetic weights of this repo)
      strategy="cuaev"   (its CUDA AEV extension, aev/_computer.py:383-407, csrc/aev.cu)
      neighborlist="cell_list" (neighbors.py:366-507, ATen ops) -- and "fast_cell_list" (csrc/cell_list.cpp) as a
                         second row
      .to_infer_model()  (BmmEnsemble, nn/_infer.py:61-216) -- and use_mnp=True / the plain Ensemble loop as rows
    driven through torchani.grad.energies_and_forces (grad.py:263-290), float32, TF32 off and on,
    100 warm-up + 50 timed steps bracketed by CUDA events (the method of tools/tool_utils.py:198-275).

The reference comes from oracle/_ref/torchani (staged by oracle/build_ref.sh; baseline infrastructure, never part
of the product).  Both arms see the same inputs; the script also checks that they agree (energy, forces) before
it reports any time.  Writes one JSON record (gpurun_out/reference_gpu_path.json when that directory exists).

    python tools/reference_gpu_path.py [--molecules 3333] [--warmup 100] [--steps 50]
"""
from __future__ import annotations

import argparse
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def timeit(fn, warmup: int, steps: int) -> float:
    for _ in range(warmup):
        fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(steps):
        fn()
    b.record()
    b.synchronize()
    return a.elapsed_time(b) / steps


def main() -> None:
    ap = argparse.ArgumentParser()
    ap.add_argument("--molecules", type=int, default=3333)
    ap.add_argument("--warmup", type=int, default=100)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--out", default=None)
    args = ap.parse_args()
    import oracle.ref_torchani as rt
    from torchani_b200 import models, synthetic
    from torchani_b200.calculator import HostCalculator

    dev = torch.device("cuda", 0)
    torch.cuda.set_device(dev)
    ext = rt.extensions()
    z, idx, coords, cell, pbc = synthetic.water_box(args.molecules, seed=0)
    n_atoms = z.shape[1]
    weights = synthetic.make_weights(models.SYMBOLS_2X, synthetic.DIMS_2X, 1008, 8, seed=1234)
    z_d, cell_d, pbc_d = z.to(dev), cell.to(dev), pbc.to(dev)
    rec = {"atoms": n_atoms, "warmup": args.warmup, "steps": args.steps, "gpu": torch.cuda.get_device_name(dev),
           "torch": torch.__version__, "reference_extensions": ext, "rows": {}}

    # ---- this repo
    ours = models.from_weight_lists("2x", weights, device=dev, periodic_table_index=True)
    eng = ours.engine(dev)
    sp_d, co_d = idx.to(dev), coords.to(dev)

    def ours_device():
        eng.step(sp_d, co_d, cell_d, True)

    def ours_module():
        c = co_d.clone().requires_grad_(True)
        e = ours((z_d, c), cell_d, pbc_d).energies
        torch.autograd.grad(e.sum(), c)

    e_ours, f_ours = ours.energies_and_forces(z_d, co_d, cell_d, pbc_d)
    eng.check_status()
    e_ours, f_ours = e_ours.clone(), f_ours.clone()
    rec["rows"]["ours_fused_engine_device_inputs"] = timeit(ours_device, 20, args.steps)
    rec["rows"]["ours_ANI_forward_plus_autograd_grad"] = timeit(ours_module, 20, args.steps)
    calc = HostCalculator(ours, z[0].numpy(), cell.numpy(), pbc=True)
    h_pos = coords[0].numpy().copy()
    rec["rows"]["ours_host_calculator_e2e"] = timeit(lambda: calc.calculate(h_pos), 20, args.steps)

    # ---- the reference, GPU path
    def ref_row(name, tf32, **kw):
        torch.backends.cuda.matmul.allow_tf32 = tf32
        torch.backends.cudnn.allow_tf32 = tf32
        try:
            model = rt.build_model(weights, "2x", dev, **kw)
            c = coords.to(dev)

            def step():
                return rt.energies_and_forces(model, z_d, c, cell_d, pbc_d)

            e, f = step()
            de = abs(float(e[0]) - float(e_ours[0]))
            df = float((f - f_ours).abs().max())
            t = timeit(step, args.warmup, args.steps)
            rec["rows"][name] = {"ms_per_step": t, "abs_dE_vs_ours_Ha": de, "max_abs_dF_vs_ours_Ha_per_A": df,
                                 "tf32": tf32, **{k: str(v) for k, v in kw.items()}}
        except Exception as exc:   # a row that cannot run here is reported, not hidden
            rec["rows"][name] = {"error": f"{type(exc).__name__}: {exc}"}
        finally:
            torch.backends.cuda.matmul.allow_tf32 = False
            torch.backends.cudnn.allow_tf32 = False

    head = dict(strategy="cuaev", neighborlist="cell_list", infer=True)
    ref_row("reference_cuaev_celllist_bmm_fp32", False, **head)                      # the north-star baseline
    ref_row("reference_cuaev_celllist_bmm_tf32", True, **head)
    ref_row("reference_cuaev_fastcelllist_bmm_fp32", False, strategy="cuaev", neighborlist="fast_cell_list", infer=True)
    ref_row("reference_cuaev_celllist_mnp_fp32", False, strategy="cuaev", neighborlist="cell_list", infer=True,
            use_mnp=True)
    ref_row("reference_cuaev_celllist_ensemble_loop_fp32", False, strategy="cuaev", neighborlist="cell_list")
    ref_row("reference_pyaev_celllist_bmm_fp32", False, strategy="pyaev", neighborlist="cell_list", infer=True)
    base = rec["rows"].get("reference_cuaev_celllist_bmm_fp32", {})
    if "ms_per_step" in base:
        rec["ratio_reference_fp32_over_ours_device"] = base["ms_per_step"] / rec["rows"]["ours_fused_engine_device_inputs"]
        rec["ratio_reference_fp32_over_ours_e2e"] = base["ms_per_step"] / rec["rows"]["ours_host_calculator_e2e"]
        t32 = rec["rows"].get("reference_cuaev_celllist_bmm_tf32", {}).get("ms_per_step")
        if t32:
            rec["ratio_reference_tf32_over_ours_e2e"] = t32 / rec["rows"]["ours_host_calculator_e2e"]
    print(json.dumps(rec, indent=1))
    out = args.out or (os.path.join(ROOT, "gpurun_out", "reference_gpu_path.json")
                       if os.path.isdir(os.path.join(ROOT, "gpurun_out")) else None)
    if out:
        with open(out, "w") as fh:
            json.dump(rec, fh, indent=1)


if __name__ == "__main__":
    main()
