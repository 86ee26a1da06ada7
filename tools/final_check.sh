#!/bin/bash
# The driver's end-of-round sequence on one GPU, with the ncu records that go under profiles/:
#   gpurun --timeout 1500 -- 'bash tools/final_check.sh TAG'
TAG=${1:-r2}
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -rs > gpurun_out/pytest_gpu_${TAG}.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/pytest_gpu_${TAG}.log
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke_${TAG}.log 2>&1; echo "smoke rc=$?"; tail -1 gpurun_out/smoke_${TAG}.log
timeout 400 python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/bench_${TAG}_ref.json 2> gpurun_out/bench_${TAG}_ref.err; echo "ref rc=$?"; cut -c1-300 gpurun_out/bench_${TAG}_ref.json
timeout 400 python bench.py > gpurun_out/bench_${TAG}_1gpu.json 2> gpurun_out/bench_${TAG}_1gpu.err; echo "bench rc=$?"; cat gpurun_out/bench_${TAG}_1gpu.json
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_${TAG}.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-parity > gpurun_out/ncu_launches_${TAG}.log 2>&1; echo "launch list rc=$?"
timeout 300 ncu --set full --clock-control none --import-source on -k regex:qb2_pass_jit --launch-skip 2 --launch-count 2 -o gpurun_out/${TAG}_pass_jit -f env QB2_SPECIALIZE=1 python tools/r2_prof.py ghzqft 30 2 > gpurun_out/ncu_${TAG}_jit.log 2>&1; echo "ncu jit rc=$?"
timeout 300 ncu --set full --clock-control none --import-source on -k regex:block5_tc --launch-skip 9 --launch-count 1 -o gpurun_out/${TAG}_block_tc -f python tools/block_bench.py 30 > gpurun_out/ncu_${TAG}_tc.log 2>&1; echo "ncu tc rc=$?"
