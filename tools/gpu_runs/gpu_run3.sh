set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
( time timeout 600 python -m pytest tests -m gpu -x -q ) > gpurun_out/r3_pytest.log 2>&1
tail -5 gpurun_out/r3_pytest.log
if ! grep -q " passed" gpurun_out/r3_pytest.log || grep -q "failed" gpurun_out/r3_pytest.log; then
  ( ANI_B200_AEV_LEGACY=1 timeout 600 python -m pytest tests -m gpu -x -q ) > gpurun_out/r3_pytest_legacy.log 2>&1
  tail -5 gpurun_out/r3_pytest_legacy.log
fi
timeout 300 python tests/gpu_diag.py > gpurun_out/r3_diag.log 2>&1
grep -E "^===|AEV fwd|forces:|status" gpurun_out/r3_diag.log | head -60
timeout 300 python bench.py --steps 50 --warmup 10 > gpurun_out/r3_bench.json 2> gpurun_out/r3_bench.err
cat gpurun_out/r3_bench.json
ANI_B200_AEV_LEGACY=1 timeout 300 python bench.py --steps 50 --warmup 10 --cpu-steps 0 > gpurun_out/r3_bench_legacy.json 2> gpurun_out/r3_bench_legacy.err
python - <<'PY'
import json
for f in ("gpurun_out/r3_bench.json","gpurun_out/r3_bench_legacy.json"):
    try:
        d=json.load(open(f)); print(f, d["ms_per_step"], d["e2e"]["ms_per_step"], d["stage_ms"])
    except Exception as e: print(f, "ERR", e)
PY
# every launch of this library in two steps (eager steps 2-3 of the warm-up), device time per launch
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:'^k_|k_gemm_tc|k_aev|k_prep|k_layout|k_reduce|k_zero' -s 16 -c 34 --csv \
  --log-file gpurun_out/r3_launches.csv python bench.py --steps 2 --warmup 3 --cpu-steps 0 > gpurun_out/r3_launches.out 2>&1
tail -40 gpurun_out/r3_launches.csv | cut -c1-200
