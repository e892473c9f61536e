set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
ANI_B200_GEMM_PAIR=1 timeout 200 python tools/gemm_trace.py > gpurun_out/r10_gemm_trace_pair.log 2>&1
timeout 200 python tools/gemm_trace.py > gpurun_out/r10_gemm_trace_single.log 2>&1
ANI_B200_GEMM_PAIR=1 ANI_B200_PDL=0 timeout 200 python bench.py --steps 50 --warmup 10 --cpu-steps 0 > gpurun_out/r10_bench_pair_nopdl.json 2> gpurun_out/r10_bench_pair_nopdl.err
python - <<'PY'
import json
d=json.load(open("gpurun_out/r10_bench_pair_nopdl.json")); print(round(d["ms_per_step"],4), {k:round(v,4) for k,v in d["stage_ms"].items()})
PY
