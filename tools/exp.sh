#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -2 gpurun_out/pytest_gpu.log
timeout 100 python tools/run_circuit.py qft 30
timeout 100 python tools/run_circuit.py random 30 20
timeout 120 python tools/trace_tile.py qft 30 > gpurun_out/trace_qft30.txt 2>&1; echo rc=$?; grep -A1 "=== pass" gpurun_out/trace_qft30.txt | cut -c1-900 | sed 's/: total [0-9]* (to next start \([0-9]*\)).*start+0/ next=\1 :/'
