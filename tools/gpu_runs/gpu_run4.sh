set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
( time timeout 600 python -m pytest tests -m gpu -x -q ) > gpurun_out/r4_pytest.log 2>&1
tail -5 gpurun_out/r4_pytest.log
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'k_aev_' -s 4 -c 2 \
  -f -o gpurun_out/r4_aev python bench.py --steps 1 --warmup 3 --cpu-steps 0 > gpurun_out/r4_aev.out 2>&1
ls -la gpurun_out
