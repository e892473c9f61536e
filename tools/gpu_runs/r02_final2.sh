#!/bin/bash
# after the last library change (fallback of the 16-CTA cluster launch): parity of the preparation kernels, the default
# bench line, and the ncu captures of THIS build (traffic.json is keyed by the library hash)
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q > gpurun_out/r02_final2_pytest.log 2>&1; echo "pytest rc=$?"; tail -2 gpurun_out/r02_final2_pytest.log
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02_final_launches.csv python bench.py --steps 2 --warmup 3 --cpu-steps 0 > gpurun_out/r02_final_launches.log 2>&1; echo "launches rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'k_aev_forward_cta|k_aev_backward|k_mlp_fused|k_gemm_tc|k_prep_fused|k_prep_cluster' -c 12 -o gpurun_out/r02_final_ncu_full -f python bench.py --steps 1 --warmup 3 --cpu-steps 0 > gpurun_out/r02_final_ncu.log 2>&1; echo "ncu full rc=$?"
ncu -i gpurun_out/r02_final_ncu_full.ncu-rep --page raw --csv > gpurun_out/r02_final_ncu_full_raw.csv 2>/dev/null
python tools/make_traffic.py gpurun_out/r02_final_ncu_full_raw.csv water10k; cp profiles/traffic.json gpurun_out/traffic.json
timeout 400 python bench.py --steps 20 --warmup 5 > gpurun_out/r02_final_bench_water10k.json 2> gpurun_out/r02_final_bench_water10k.err; echo "bench rc=$?"
python -c "
import json; d=json.loads(open('gpurun_out/r02_final_bench_water10k.json').read().strip().splitlines()[-1]); print(d['ms_per_step'], d['e2e']['ms_per_step'], d['roofline']['traffic'], d['library_sha256_16'])"
