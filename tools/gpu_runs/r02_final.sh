#!/bin/bash
# round 2, final evidence run (1 GPU): smoke, full GPU suite, every bench config with the CPU reference beside it,
# the reference arm, the reference's GPU path, launch list + ncu --set full capture of the main kernels
mkdir -p gpurun_out
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r02_final_smoke.log 2>&1; echo "smoke rc=$?"; tail -2 gpurun_out/r02_final_smoke.log
timeout 1500 python -m pytest tests -m gpu -q -rfEs --tb=short > gpurun_out/r02_final_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02_final_pytest.log
grep -v "^  File\|^Extension" gpurun_out/r02_final_pytest.log | tail -8
for cfg in water10k water1k gdb256 protein50k; do
  timeout 400 python bench.py --config $cfg --steps 20 --warmup 5 > gpurun_out/r02_final_bench_$cfg.json 2> gpurun_out/r02_final_bench_$cfg.err; echo "bench $cfg rc=$?"
  python - <<PY
import json
try:
    d = json.loads(open("gpurun_out/r02_final_bench_$cfg.json").read().strip().splitlines()[-1])
    print("$cfg", round(d["ms_per_step"],4), "e2e", round(d["e2e"]["ms_per_step"],4), {k: round(v,4) for k,v in d["stage_ms"].items()}, "cpu ms", d["cpu_baseline"] and round(d["cpu_baseline"]["ms_per_step"],1), "frac", round(d["roofline"]["frac"],3), round(d["roofline_aev"]["forward"]["frac"],3), round(d["roofline_aev"]["backward"]["frac"],3))
except Exception as e:
    print("$cfg failed", e); print(open("gpurun_out/r02_final_bench_$cfg.err").read()[-1500:])
PY
done
timeout 400 python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/r02_final_bench_reference.json 2> gpurun_out/r02_final_bench_reference.err; echo "reference arm rc=$?"
timeout 600 python tools/reference_gpu_path.py --out gpurun_out/r02_final_reference_gpu_path.json > gpurun_out/r02_final_refpath.log 2>&1; echo "refpath rc=$?"; grep ratio gpurun_out/r02_final_refpath.log
# launch list of two steps (eager warm-up launches + graph replays are all visible to ncu)
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02_final_launches.csv python bench.py --steps 2 --warmup 3 --cpu-steps 0 > gpurun_out/r02_final_launches.log 2>&1; echo "launches rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'k_aev_forward_cta|k_aev_backward|k_mlp_fused|k_gemm_tc|k_prep_fused|k_prep_cluster' -c 12 -o gpurun_out/r02_final_ncu_full -f python bench.py --steps 1 --warmup 3 --cpu-steps 0 > gpurun_out/r02_final_ncu.log 2>&1; echo "ncu full rc=$?"
ncu -i gpurun_out/r02_final_ncu_full.ncu-rep --page raw --csv > gpurun_out/r02_final_ncu_full_raw.csv 2>/dev/null
python tools/make_traffic.py gpurun_out/r02_final_ncu_full_raw.csv water10k; cp profiles/traffic.json gpurun_out/traffic.json
