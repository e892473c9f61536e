set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
# (1) every launch of two eager steps with its device time (shares, not absolutes)
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 48 -c 48 --csv \
  --log-file gpurun_out/r2_launches.csv python bench.py --steps 2 --warmup 3 --cpu-steps 0 > gpurun_out/r2_launches.out 2>&1
# (2) full sections + source counters for the AEV kernels and the six GEMMs of one step
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'k_gemm_tc|k_aev_' -s 16 -c 8 \
  -f -o gpurun_out/r2_full python bench.py --steps 1 --warmup 3 --cpu-steps 0 > gpurun_out/r2_full.out 2>&1
ls -la gpurun_out
# (3) GEMM role timeline
timeout 300 python tools/gemm_trace.py > gpurun_out/r2_gemm_trace.log 2>&1
tail -5 gpurun_out/r2_gemm_trace.log
