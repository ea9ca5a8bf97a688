#!/bin/bash
# Round 2, session v: ncu launch list of the bench command, full captures of the step's two kernels and the streamed kernel,
# caller-pinned arrays (test + bench leg).
set -u
O=gpurun_out
mkdir -p $O
( timeout -s KILL 300 python -m pytest tests/test_gpu_robustness.py -x -q -m gpu ) > $O/r2v_pytest_robust.log 2>&1
tail -4 $O/r2v_pytest_robust.log
timeout -s KILL 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/r2v_ncu_launches_bench.csv python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu --no-fftn --no-other > $O/r2v_bench_under_ncu.log 2>&1
tail -c 300 $O/r2v_bench_under_ncu.log; wc -l $O/r2v_ncu_launches_bench.csv
timeout -s KILL 300 ncu --set full --clock-control none --import-source on -k regex:fused2 -s 1 -c 1 -o $O/r2v_cols_fused2 python tools/prof_target.py cols 3 > $O/r2v_ncu_cols.log 2>&1
tail -2 $O/r2v_ncu_cols.log
timeout -s KILL 300 ncu --set full --clock-control none --import-source on -k regex:dual -s 1 -c 1 -o $O/r2v_rows_dual python tools/prof_target.py rows 3 > $O/r2v_ncu_rows.log 2>&1
tail -2 $O/r2v_ncu_rows.log
timeout -s KILL 300 ncu --set full --clock-control none --import-source on -k regex:stream -s 1 -c 1 -o $O/r2v_axis1_stream python tools/prof_target.py cfg3 2 > $O/r2v_ncu_stream.log 2>&1
tail -2 $O/r2v_ncu_stream.log
( time timeout 900 python bench.py --steps 20 --warmup 5 --no-fftn --no-other ) > $O/r2v_bench.json 2> $O/r2v_bench.err
tail -c 1500 $O/r2v_bench.json; tail -3 $O/r2v_bench.err
