#!/bin/bash
# host loop with ONE H2D and ONE D2H copy per step (python-only change): API / calculator tests, bench line
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_api.py tests/test_gpu_ase.py tests/test_gpu_external.py tests/test_gpu_dropin.py -m gpu -x -q > gpurun_out/r02_final4_pytest.log 2>&1; echo "pytest rc=$?"; tail -2 gpurun_out/r02_final4_pytest.log
timeout 400 python bench.py --steps 30 --warmup 5 --cpu-steps 0 > gpurun_out/r02_final4_bench_water10k.json 2> gpurun_out/r02_final4_bench_water10k.err; echo "bench rc=$?"
python -c "
import json; d=json.loads(open('gpurun_out/r02_final4_bench_water10k.json').read().strip().splitlines()[-1]); print(d['ms_per_step'], d['e2e'])"
