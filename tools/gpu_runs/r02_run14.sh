#!/bin/bash
# round 2, GPU run 14: suite after thin-cell supercells, one-pass member forces, pair-potential drop-in
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -rfEs --tb=short > gpurun_out/r02_run14_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02_run14_pytest.log
grep -v "^  File\|^Extension" gpurun_out/r02_run14_pytest.log | tail -40
timeout 300 python bench.py --steps 20 --warmup 5 --cpu-steps 0 > gpurun_out/r02_run14_10k.json 2> gpurun_out/r02_run14_10k.err
python -c "
import json; d=json.loads(open('gpurun_out/r02_run14_10k.json').read().strip().splitlines()[-1]); print('10k', d['ms_per_step'], 'e2e', d['e2e']['ms_per_step'], d['stage_ms'])"
python - <<'PY'
# member forces: one pass vs the time of 8 separate steps
import torch, time
from torchani_b200 import models, synthetic
dev = torch.device("cuda:0")
w = synthetic.make_weights(models.SYMBOLS_2X, synthetic.DIMS_2X, 1008, 8, seed=1234)
m = models.from_weight_lists("2x", w, device=dev, periodic_table_index=True)
z, idx, coords, cell, pbc = synthetic.water_box(3333, seed=0)
args = ((z.to(dev), coords.to(dev)), cell.to(dev), pbc.to(dev))
for _ in range(3): out = m.members_forces(*args)
torch.cuda.synchronize(); t0 = time.perf_counter()
for _ in range(10): out = m.members_forces(*args)
torch.cuda.synchronize(); print("members_forces one pass: %.3f ms" % ((time.perf_counter() - t0) * 100), out.energies.shape, out.forces.shape)
PY
