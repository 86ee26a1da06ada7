#!/bin/bash
# Box visit on 2 GPUs: the GPU test suite (the sharded test runs here), the sharded parity script with its
# log kept, and the bench at N=1 and N=2.     gpurun --gpus 2 --timeout 1500 -- 'bash tools/visit_2gpu.sh TAG'
TAG=${1:-r2}
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -rs > gpurun_out/pytest_gpu_${TAG}_2gpu.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/pytest_gpu_${TAG}_2gpu.log
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tests/dist_gpu_check.py > gpurun_out/dist_gpu_check_${TAG}_2gpu.log 2>&1; echo "dist check rc=$?"; grep -v Warn gpurun_out/dist_gpu_check_${TAG}_2gpu.log | tail -8
timeout 400 python bench.py --steps 10 --warmup 3 > gpurun_out/bench_${TAG}_1gpu.json 2> gpurun_out/bench_${TAG}_1gpu.err; echo "bench1 rc=$?"; cat gpurun_out/bench_${TAG}_1gpu.json; tail -3 gpurun_out/bench_${TAG}_1gpu.err
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/bench_${TAG}_2gpu.json 2> gpurun_out/bench_${TAG}_2gpu.err; echo "bench2 rc=$?"; cat gpurun_out/bench_${TAG}_2gpu.json; tail -3 gpurun_out/bench_${TAG}_2gpu.err
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_${TAG}_ref.json 2> gpurun_out/bench_${TAG}_ref.err; echo "ref rc=$?"; cat gpurun_out/bench_${TAG}_ref.json
nproc; free -g | head -2
