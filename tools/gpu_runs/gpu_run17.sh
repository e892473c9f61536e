set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
( ANI_B200_REF_PATH_TEST=1 timeout 100 python -m pytest tests/test_gpu_reference_cuaev.py -m gpu -q -s -x -k pieces ) 2>&1 | grep -v lazyInitCUDA > gpurun_out/r17_ref_path.log
tail -12 gpurun_out/r17_ref_path.log
