set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
( timeout 150 python -m pytest tests/test_gpu_reference_cuaev.py -m gpu -q -s -x ) > gpurun_out/r16_ref_cuaev.log 2>&1
tail -25 gpurun_out/r16_ref_cuaev.log
