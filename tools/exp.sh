#!/bin/bash
# One GPU box visit: parity tests, the headline bench, two circuit timings and the tile trace.
#   gpurun --timeout 900 -- 'bash tools/exp.sh'
# Every command has its own timeout: a hung kernel must not eat the box's budget.
mkdir -p gpurun_out
timeout 400 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -2 gpurun_out/pytest_gpu.log
timeout 600 python bench.py > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench rc=$?"; cat gpurun_out/bench.json
timeout 100 python tools/run_circuit.py qft 30
timeout 100 python tools/run_circuit.py random 30 20
if [ -f qrisp_b200/lib/libqb2_trace.so ]; then
  timeout 120 python tools/trace_tile.py qft 30 > gpurun_out/trace_qft30.txt 2>&1; echo "trace rc=$?"
  grep -A1 "=== pass" gpurun_out/trace_qft30.txt | cut -c1-600
fi
