set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv
( time timeout 600 python -m pytest tests -m gpu -x -q ) > gpurun_out/r1_pytest_fp16.log 2>&1
tail -5 gpurun_out/r1_pytest_fp16.log
timeout 300 python bench.py --steps 50 --warmup 10 > gpurun_out/r1_bench_fp16.json 2> gpurun_out/r1_bench_fp16.err
cat gpurun_out/r1_bench_fp16.json
ANI_B200_LIB=$GRAFT_REPO_ROOT/torchani_b200/libani_b200_bf16x3.so timeout 300 python bench.py --steps 50 --warmup 10 --cpu-steps 0 > gpurun_out/r1_bench_bf16x3.json 2> gpurun_out/r1_bench_bf16x3.err
cat gpurun_out/r1_bench_bf16x3.json
timeout 300 python tests/gpu_diag.py > gpurun_out/r1_diag_fp16.log 2>&1
tail -30 gpurun_out/r1_diag_fp16.log
