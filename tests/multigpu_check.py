"""Multi-GPU check of the sharded step with the peer-memory reduction (run under torchrun, one rank per GPU):

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 \
        tests/multigpu_check.py

Every rank evaluates its slice of a periodic water box; the in-graph reduction (csrc/comm.cu) must give, on EVERY
rank, bitwise identical totals that agree with the single-GPU evaluation of the whole system, eagerly and through
CUDA-graph replays, and the host calculator must return the full forces on every rank.  Prints one line per rank;
exit code 0 = all checks passed.  (tests/test_gpu_multi.py runs it when the box has >= 2 GPUs.)"""
import os
import sys

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    local = int(os.environ.get("LOCAL_RANK", rank))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    from torchani_b200 import models, synthetic
    from torchani_b200.calculator import HostCalculator
    from torchani_b200.parallel import ShardedEngine
    w = synthetic.make_weights(models.SYMBOLS_2X, synthetic.DIMS_2X, 1008, 8, seed=1234)
    model = models.from_weight_lists("2x", w, device=dev, periodic_table_index=True)
    eng = model.engine(dev)
    mode = os.environ.get("ANI_B200_REDUCE", "auto")
    for n_mol in (333, 3333):
        z, idx, coords, cell, pbc = synthetic.water_box(n_mol, seed=0)
        sp, co, ce = idx.to(dev), coords.to(dev), cell.to(dev)
        whole = eng.step(sp, co, ce, True)
        e_ref, g_ref = whole.energies.clone(), whole.grad.clone()
        sh = ShardedEngine(eng, reduce=mode)
        for it in range(8):      # eager uses, then the captured graph (Engine.graph_after = 3)
            e, g = sh.step(sp, co, ce, True)
            torch.cuda.synchronize()
            assert abs(float(e[0]) - float(e_ref[0])) < 1e-5, (it, float(e[0]), float(e_ref[0]))
            assert float((g - g_ref).abs().max()) < 2e-5, (it, float((g - g_ref).abs().max()))
            # identical bits on every rank
            gs = [torch.empty_like(g) for _ in range(world)]
            dist.all_gather(gs, g.contiguous())
            assert all(torch.equal(gs[0], x) for x in gs), "ranks disagree on the reduced forces"
        eng.check_status()
        calc = HostCalculator(model, z[0].numpy(), cell.numpy(), pbc=True, sharded=sh)
        for it in range(7):
            e_h, f_h = calc.calculate(coords[0].numpy())
            assert abs(e_h - float(e_ref[0])) < 1e-5
            assert float((torch.from_numpy(f_h) + g_ref[0].cpu()).abs().max()) < 2e-5
        print(f"rank {rank}/{world}: {3 * n_mol} atoms ok, reduction mode = {sh.mode}, E = {e_h:.6f}", flush=True)
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
