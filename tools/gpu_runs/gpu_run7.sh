set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
# launch list: two eager steps of this library's kernels (device time per launch; compare shares)
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:'^k_|k_gemm_tc|k_aev|k_prep|k_layout|k_reduce|k_zero' -s 15 -c 30 --csv \
  --log-file gpurun_out/r7_launches.csv python bench.py --steps 2 --warmup 3 --cpu-steps 0 > gpurun_out/r7_launches.out 2>&1
# full sections for the 8 main kernels of one step
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'k_gemm_tc|k_aev_' -s 16 -c 8 \
  -f -o gpurun_out/r7_full python bench.py --steps 1 --warmup 3 --cpu-steps 0 > gpurun_out/r7_full.out 2>&1
timeout 300 python tools/gemm_trace.py > gpurun_out/r7_gemm_trace.log 2>&1
timeout 300 python bench.py --steps 50 --warmup 10 > gpurun_out/r7_bench.json 2> gpurun_out/r7_bench.err
timeout 300 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r7_bench_reference.json 2> gpurun_out/r7_bench_reference.err
cat gpurun_out/r7_bench_reference.json | cut -c1-400
ls -la gpurun_out | head -40
