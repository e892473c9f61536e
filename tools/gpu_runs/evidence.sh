# One gpurun call that captures the evidence of a build (launch list, full ncu sections of the eight main
# kernels, GEMM role timeline, bench line):  gpurun --timeout 1800 -- 'bash tools/gpu_runs/evidence.sh TAG'
# Read here with: ncu -i gpurun_out/TAG_full.ncu-rep --page raw --csv > x.csv; python tools/ncu_summary.py x.csv
set -x
TAG=${1:-ev}
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:'^k_|k_gemm_tc|k_aev|k_prep|k_layout|k_reduce|k_zero|k_verlet' -s 15 -c 30 --csv \
  --log-file gpurun_out/${TAG}_launches.csv python bench.py --steps 2 --warmup 3 --cpu-steps 0 > gpurun_out/${TAG}_launches.out 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'k_gemm_tc|k_aev_' -s 16 -c 8 \
  -f -o gpurun_out/${TAG}_full python bench.py --steps 1 --warmup 3 --cpu-steps 0 > gpurun_out/${TAG}_full.out 2>&1
timeout 300 python tools/gemm_trace.py > gpurun_out/${TAG}_gemm_trace.log 2>&1
timeout 300 python bench.py --steps 50 --warmup 10 --cpu-steps 0 > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err
