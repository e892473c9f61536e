set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 200 python bench.py --steps 50 --warmup 10 --cpu-steps 0 > gpurun_out/r11_bench_epi1.json 2> gpurun_out/r11_bench_epi1.err
ANI_B200_GEMM_DEBUG=64 timeout 200 python bench.py --steps 50 --warmup 10 --cpu-steps 0 > gpurun_out/r11_bench_epi2.json 2> gpurun_out/r11_bench_epi2.err
ANI_B200_GEMM_PAIR=1 timeout 200 python bench.py --steps 50 --warmup 10 --cpu-steps 0 > gpurun_out/r11_bench_pair_epi1.json 2> gpurun_out/r11_bench_pair_epi1.err
python - <<'PY'
import json
for f in ("r11_bench_epi1","r11_bench_epi2","r11_bench_pair_epi1"):
    try:
        d=json.load(open(f"gpurun_out/{f}.json")); print(f, round(d["ms_per_step"],4), {k:round(v,4) for k,v in d["stage_ms"].items()})
    except Exception as e: print(f, "ERR", e)
PY
( timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -x -q ) 2>&1 | tail -3
