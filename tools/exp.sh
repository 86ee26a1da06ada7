run() { timeout 100 python tools/pass_micro.py 28 copy h4 h8 cx4_reg_ctrl h8_two_phases h_cp_ladder4 2>&1 | tail -6; timeout 200 python bench.py --steps 3 --warmup 2 --no-cpu-baseline 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('BENCH ms_per_step', d['ms_per_step'])"; }
echo "=== lean/full split"; run
