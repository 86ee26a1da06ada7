#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -2 gpurun_out/pytest_gpu.log
timeout 100 python tools/run_circuit.py qft 30
timeout 100 python tools/run_circuit.py random 30 20
QB2_NO_CHAIN=1 timeout 100 python tools/run_circuit.py random 30 20
