#!/bin/bash
mkdir -p gpurun_out
timeout 600 python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/launches_v5.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_launches.log 2>&1
echo "ncu launches rc=$?"
ncu --set full --clock-control none --import-source on -k regex:pass_kernel -s 5 -c 5 -o gpurun_out/prof_v7 python bench.py --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_v7.log 2>&1
echo "ncu full rc=$?"
ncu -i gpurun_out/prof_v7.ncu-rep --page raw --csv > gpurun_out/prof_v7_raw.csv 2>/dev/null
ncu -i gpurun_out/prof_v7.ncu-rep --page source --csv > gpurun_out/prof_v7_source.csv 2>/dev/null
ls -la gpurun_out; rm -f gpurun_out/prof_v7.ncu-rep
