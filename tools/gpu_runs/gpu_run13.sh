set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
for v in "" fwd8w fwd6w; do
  if [ -z "$v" ]; then lib=$GRAFT_REPO_ROOT/torchani_b200/libani_b200.so; else lib=$GRAFT_REPO_ROOT/torchani_b200/libani_b200_$v.so; fi
  ANI_B200_LIB=$lib timeout 200 python bench.py --steps 50 --warmup 10 --cpu-steps 0 > gpurun_out/r13_bench_${v:-default}.json 2> gpurun_out/r13_bench_${v:-default}.err
done
ANI_B200_LIB=$GRAFT_REPO_ROOT/torchani_b200/libani_b200_fwd8w.so timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -x -q 2>&1 | tail -3
python - <<'PY'
import json
for f in ("default","fwd8w","fwd6w"):
    try:
        d=json.load(open(f"gpurun_out/r13_bench_{f}.json")); print(f, round(d["ms_per_step"],4), round(d["e2e"]["ms_per_step"],4), {k:round(v,4) for k,v in d["stage_ms"].items()})
    except Exception as e: print(f, "ERR", e)
PY
