#!/bin/bash
# Round 2, session i: the whole GPU test-suite (timed), smoke(), the default bench line, the reference arm.
set -u
O=gpurun_out
mkdir -p $O
( time timeout 2400 python -m pytest tests/ -x -q -m gpu --durations=15 ) > $O/r2i_pytest_gpu.log 2>&1
tail -30 $O/r2i_pytest_gpu.log
( time timeout 300 python -c "import __graft_entry__ as g; g.smoke()" ) > $O/r2i_smoke.log 2>&1
tail -4 $O/r2i_smoke.log
( time timeout 900 python bench.py --gpus 1 --steps 20 --warmup 5 ) > $O/r2i_bench.json 2> $O/r2i_bench.err
tail -c 6000 $O/r2i_bench.json; tail -5 $O/r2i_bench.err
( time timeout 600 python bench.py --impl reference --gpus 1 --steps 20 --warmup 5 ) > $O/r2i_bench_ref.json 2> $O/r2i_bench_ref.err
tail -c 1500 $O/r2i_bench_ref.json; tail -4 $O/r2i_bench_ref.err
