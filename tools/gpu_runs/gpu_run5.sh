set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
( time timeout 600 python -m pytest tests -m gpu -x -q ) > gpurun_out/r5_pytest.log 2>&1
tail -5 gpurun_out/r5_pytest.log
if grep -q "failed\|error" gpurun_out/r5_pytest.log; then
  ( ANI_B200_PDL=0 timeout 600 python -m pytest tests -m gpu -x -q ) > gpurun_out/r5_pytest_nopdl.log 2>&1
  tail -5 gpurun_out/r5_pytest_nopdl.log
fi
timeout 300 python bench.py --steps 50 --warmup 10 > gpurun_out/r5_bench.json 2> gpurun_out/r5_bench.err
ANI_B200_PDL=0 timeout 300 python bench.py --steps 50 --warmup 10 --cpu-steps 0 > gpurun_out/r5_bench_nopdl.json 2> gpurun_out/r5_bench_nopdl.err
timeout 300 python bench.py --steps 50 --warmup 10 --cpu-steps 0 --molecules 333 > gpurun_out/r5_bench_1k.json 2> gpurun_out/r5_bench_1k.err
timeout 300 python bench.py --steps 20 --warmup 5 --cpu-steps 0 --molecules 16667 > gpurun_out/r5_bench_50k.json 2> gpurun_out/r5_bench_50k.err
python - <<'PY'
import json
for f in ("r5_bench","r5_bench_nopdl","r5_bench_1k","r5_bench_50k"):
    try:
        d=json.load(open(f"gpurun_out/{f}.json")); print(f, d["config"]["atoms"], round(d["ms_per_step"],4), round(d["e2e"]["ms_per_step"],4), {k:round(v,4) for k,v in d["stage_ms"].items()})
    except Exception as e: print(f, "ERR", e)
PY
tail -3 gpurun_out/r5_bench.err
