#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -2 gpurun_out/pytest_gpu.log
timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/bench_r01_v7.json 2> gpurun_out/bench_r01_v7.err; echo "bench rc=$?"; cat gpurun_out/bench_r01_v7.json; tail -2 gpurun_out/bench_r01_v7.err
timeout 100 python tools/run_circuit.py qft 30
