for d in 0 8 31; do
ANI_B200_GEMM_DEBUG=$d python - <<PY
import sys, json, torch
sys.path.insert(0,'.')
from torchani_b200 import models, synthetic
dev=torch.device('cuda:0')
z, idx, coords, cell, pbc = synthetic.water_box(3333, seed=0)
w = synthetic.make_weights(models.SYMBOLS_2X, synthetic.DIMS_2X, 1008, 8, seed=1234)
m = models.from_weight_lists("2x", w, device=dev, periodic_table_index=True)
eng = m.engine(dev); eng.cuda_graph=False
sp, co, ce = idx.to(dev), coords.to(dev), cell.to(dev)
for _ in range(3): eng.step(sp, co, ce, True)
eng.profile=True
for _ in range(10): eng.step(sp, co, ce, True)
t=eng.stage_times_ms()
print("debug=$d mlp_ms=%.3f" % t['mlp_forward_backward'])
PY
done
