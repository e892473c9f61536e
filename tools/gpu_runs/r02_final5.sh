#!/bin/bash
# last check of the round: the two preparation kernels bit for bit, incl. a dilute system with more buckets than atoms
mkdir -p gpurun_out
timeout 120 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "cluster_preparation or water_1k" > gpurun_out/r02_final5_pytest.log 2>&1; echo "pytest rc=$?"; tail -2 gpurun_out/r02_final5_pytest.log
