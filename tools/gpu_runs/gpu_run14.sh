set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 50 --warmup 10 > gpurun_out/r14_bench_2gpu.json 2> gpurun_out/r14_bench_2gpu.err
tail -2 gpurun_out/r14_bench_2gpu.err
cat gpurun_out/r14_bench_2gpu.json | cut -c1-900
