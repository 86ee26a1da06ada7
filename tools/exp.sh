#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest_gpu.log
for cfg in "13 5 0" "12 5 0"; do set -- $cfg
  echo "=== T=$1 r=$2 L2PF=$3"
  export QB2_TILE_BITS=$1 QB2_REG_BITS=$2 QB2_L2PF=$3
  timeout 100 python tools/run_circuit.py qft 30
  timeout 100 python tools/run_circuit.py random 30 20
  timeout 100 python tools/pass_micro.py 28 copy h8 h_cp_ladder4 cx4_reg_ctrl h8_two_phases h12_three_phases 2>&1 | tail -6
done
echo "=== no slots"; QB2_NO_SLOTS=1 QB2_TILE_BITS=13 QB2_REG_BITS=5 QB2_L2PF=0 timeout 100 python tools/run_circuit.py qft 30
