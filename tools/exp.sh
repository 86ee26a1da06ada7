#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -2 gpurun_out/pytest_gpu.log
for cfg in "13 5" "12 5"; do set -- $cfg
  echo "=== T=$1 r=$2"
  QB2_TILE_BITS=$1 QB2_REG_BITS=$2 timeout 100 python tools/run_circuit.py qft 30
  QB2_TILE_BITS=$1 QB2_REG_BITS=$2 timeout 100 python tools/run_circuit.py qftonly 30
  QB2_TILE_BITS=$1 QB2_REG_BITS=$2 timeout 100 python tools/run_circuit.py qft 24
done
timeout 300 python tools/run_circuit.py qft 33
timeout 300 python tools/run_circuit.py random 33 20
