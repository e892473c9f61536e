set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
( time timeout 600 python -m pytest tests -m gpu -x -q ) > gpurun_out/r6_pytest.log 2>&1
tail -5 gpurun_out/r6_pytest.log
if grep -q "failed\|error" gpurun_out/r6_pytest.log; then
  ( ANI_B200_SIDE_STREAM=0 timeout 600 python -m pytest tests -m gpu -x -q ) > gpurun_out/r6_pytest_noside.log 2>&1
  tail -5 gpurun_out/r6_pytest_noside.log
fi
timeout 300 python bench.py --steps 50 --warmup 10 > gpurun_out/r6_bench.json 2> gpurun_out/r6_bench.err
ANI_B200_SIDE_STREAM=0 timeout 300 python bench.py --steps 50 --warmup 10 --cpu-steps 0 > gpurun_out/r6_bench_noside.json 2> gpurun_out/r6_bench_noside.err
for v in bwd5 bwd6 fwd7; do
  ANI_B200_LIB=$GRAFT_REPO_ROOT/torchani_b200/libani_b200_$v.so timeout 300 python bench.py --steps 50 --warmup 10 --cpu-steps 0 > gpurun_out/r6_bench_$v.json 2> gpurun_out/r6_bench_$v.err
done
python - <<'PY'
import json
for f in ("r6_bench","r6_bench_noside","r6_bench_bwd5","r6_bench_bwd6","r6_bench_fwd7"):
    try:
        d=json.load(open(f"gpurun_out/{f}.json")); print(f, d["config"]["atoms"], round(d["ms_per_step"],4), round(d["e2e"]["ms_per_step"],4), {k:round(v,4) for k,v in d["stage_ms"].items()})
    except Exception as e: print(f, "ERR", e)
PY
tail -3 gpurun_out/r6_bench.err
