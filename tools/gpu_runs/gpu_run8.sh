set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
( time timeout 600 python -m pytest tests -m gpu -x -q ) > gpurun_out/r8_pytest.log 2>&1
tail -5 gpurun_out/r8_pytest.log
timeout 300 python bench.py --steps 50 --warmup 10 --cpu-steps 0 > gpurun_out/r8_bench.json 2> gpurun_out/r8_bench.err
timeout 300 python bench.py --steps 50 --warmup 10 --cpu-steps 0 --skin 0.4 > gpurun_out/r8_bench_skin04.json 2> gpurun_out/r8_bench_skin04.err
timeout 300 python bench.py --steps 50 --warmup 10 --cpu-steps 0 --skin 1.0 > gpurun_out/r8_bench_skin10.json 2> gpurun_out/r8_bench_skin10.err
python - <<'PY'
import json
for f in ("r8_bench","r8_bench_skin04","r8_bench_skin10"):
    try:
        d=json.load(open(f"gpurun_out/{f}.json")); print(f, round(d["ms_per_step"],4), d["e2e"])
    except Exception as e: print(f, "ERR", e)
PY
tail -3 gpurun_out/r8_bench_skin04.err
