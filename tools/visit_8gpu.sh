#!/bin/bash
# Box visit on 8 GPUs: the bench at N = 4 and 8 (parity cases inside), then BASELINE.json configs[4]: GHZ+QFT on
# 36 qubits (closed-form amplitude check + chi-square of samples), a random circuit on 22 qubits against the
# oracle (qubit swaps exercised), random circuits on 33 and 36 qubits (norm).
#   gpurun --gpus 8 --timeout 1500 -- 'bash tools/visit_8gpu.sh TAG'
TAG=${1:-r2}
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
for N in 4 8; do
  timeout 300 $TR --nproc-per-node $N --master-port 2951$N bench.py --gpus $N --steps 10 --warmup 3 > gpurun_out/bench_${TAG}_${N}gpu.json 2> gpurun_out/bench_${TAG}_${N}gpu.err; echo "bench$N rc=$?"; cut -c1-700 gpurun_out/bench_${TAG}_${N}gpu.json
done
: > gpurun_out/dist_${TAG}_8gpu.jsonl
export QB2_SPECIALIZE=1
for ARGS in "qft 36" "random 22 20" "random 33 20" "random 36 20" "qft 33"; do
  timeout 600 $TR --nproc-per-node 8 --master-port 29530 tools/run_circuit_dist.py $ARGS 2> gpurun_out/dist_${TAG}_8gpu.err | grep '^{' | tee -a gpurun_out/dist_${TAG}_8gpu.jsonl; echo "$ARGS rc=${PIPESTATUS[0]}"
done
tail -5 gpurun_out/dist_${TAG}_8gpu.err
