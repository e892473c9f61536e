#!/bin/bash
# where the time of the six chained GEMM launches goes: debug switches of k_gemm_tc
mkdir -p gpurun_out
for d in 0 8 6 38 134 166 2 4 10 12; do
  ANI_B200_MLP_FUSED=0 ANI_B200_GEMM_DEBUG=$d timeout 120 python tools/mlp_probe.py 2>&1 | tail -1
done | tee gpurun_out/r20_mlp_probe.log
