set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
( time timeout 600 python -m pytest tests -m gpu -q ) > gpurun_out/r12_pytest.log 2>&1
tail -30 gpurun_out/r12_pytest.log
timeout 200 python bench.py --steps 50 --warmup 10 --cpu-steps 0 > gpurun_out/r12_bench.json 2> gpurun_out/r12_bench.err
python - <<'PY'
import json
d=json.load(open("gpurun_out/r12_bench.json")); print(round(d["ms_per_step"],4), round(d["e2e"]["ms_per_step"],4), {k:round(v,4) for k,v in d["stage_ms"].items()})
PY
