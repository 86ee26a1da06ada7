#!/bin/bash
mkdir -p gpurun_out
for v in 0 1; do
  if [ $v = 1 ]; then export QB2_NO_TILE_PREFETCH=1; fi
  echo "=== no_prefetch=$v"
  timeout 100 python tools/run_circuit.py qft 30
  timeout 100 python tools/run_circuit.py qftonly 30
  timeout 100 python tools/run_circuit.py random 30 20
done
