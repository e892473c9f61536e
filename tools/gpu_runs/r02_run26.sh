#!/bin/bash
# source-level stall sampling of the data-flow MLP launch (where do the epilogue warps wait?)
mkdir -p gpurun_out
ANI_B200_MLP_CHAIN=0 ANI_B200_MLP_FUSED=1 timeout 600 ncu --set full --import-source on --clock-control none -k regex:k_mlp_fused -s 3 -c 1 -o gpurun_out/r26_fused -f python tools/mlp_probe.py > gpurun_out/r26_ncu.log 2>&1; echo "ncu rc=$?"; tail -2 gpurun_out/r26_ncu.log
ls -la gpurun_out/r26_fused.ncu-rep
