#!/bin/bash
# The non-headline workloads of BASELINE.json on one B200 (best of three runs each, warmed up).
mkdir -p gpurun_out
out=gpurun_out/other_workloads.jsonl; : > $out
for args in "qft 30" "qftonly 30" "random 30 20" "qft 24" "qft 33" "random 33 20"; do
  timeout 300 python tools/run_circuit.py $args | tee -a $out
done
