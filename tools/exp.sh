#!/bin/bash
mkdir -p gpurun_out
N=$1
timeout 300 python -m pytest tests -m gpu -x -q -k "multi or dist or shard" > gpurun_out/pytest_gpu2.log 2>&1; echo "pytest rc=$?"; tail -2 gpurun_out/pytest_gpu2.log
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 5 --warmup 3 > gpurun_out/bench_r01_final_${N}gpu.json 2> gpurun_out/bench_r01_final_${N}gpu.err; echo "bench$N rc=$?"; cat gpurun_out/bench_r01_final_${N}gpu.json | cut -c1-200; grep -o '"comm_ms_per_step": [0-9.]*' gpurun_out/bench_r01_final_${N}gpu.json; grep -o '"qubit_swaps": [0-9]*' gpurun_out/bench_r01_final_${N}gpu.json
if [ "$N" = "8" ]; then
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29512 tools/run_circuit_dist.py qft 36 > gpurun_out/qft36_8gpu.json 2> gpurun_out/qft36_8gpu.err; echo "qft36 rc=$?"; cat gpurun_out/qft36_8gpu.json; tail -2 gpurun_out/qft36_8gpu.err
fi
