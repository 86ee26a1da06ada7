#!/bin/bash
# box visit: TMA tile-traffic probe (see tma_probe.cu)
cd "$(dirname "$0")"
out=../../gpurun_out/tma_probe.log
: > $out
run() { echo "== $*" >> $out; timeout 120 ./tma_probe "$@" >> $out 2>&1; echo "rc=$?" >> $out; }
run 30 0 1 1 0
run 30 0 1 0 0
run 30 0 0 1 0
run 30 0 1 1 0 12 13 14 15 16 17 18 19 20
run 30 0 1 1 0 21 22 23 24 25 26 27 28 29
run 30 0 1 1 0 5 6 7 12 13 14 20 21 22
run 30 16 1 1 0
run 30 32 1 1 0
run 30 48 1 1 0
run 30 16 1 1 1
run 30 0 1 1 1
run 30 0 0 1 1
run 33 0 1 1 0
python -c "import ctypes; print('nvrtc', ctypes.CDLL('libnvrtc.so.12'))" >> $out 2>&1
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv >> $out 2>&1
