#!/bin/bash
# Round 2, session 3q: whole GPU test-suite after the DCT/DST test fix, smoke(), default bench line.
set -u
O=gpurun_out
mkdir -p $O
( time timeout 2400 python -m pytest tests/ -x -q -m gpu --durations=15 ) > $O/r3q_pytest_gpu.log 2>&1
tail -30 $O/r3q_pytest_gpu.log
( time timeout 300 python -c "import __graft_entry__ as g; g.smoke()" ) > $O/r3q_smoke.log 2>&1
tail -4 $O/r3q_smoke.log
( time timeout 900 python bench.py ) > $O/r3q_bench.json 2> $O/r3q_bench.err
tail -c 3000 $O/r3q_bench.json; tail -5 $O/r3q_bench.err
( time timeout 600 python bench.py --impl reference --gpus 1 --steps 10 --warmup 3 ) > $O/r3q_bench_ref.json 2> $O/r3q_bench_ref.err
tail -c 600 $O/r3q_bench_ref.json
