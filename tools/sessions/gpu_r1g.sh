#!/bin/bash
# Last GPU session of the round: smoke() and the device-resident timings of every BASELINE config with the final library.
set -u
O=gpurun_out
mkdir -p $O
echo "== smoke"
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3 | tee $O/r1g_smoke.log
echo "== microbench all configs"
timeout 260 python tools/microbench.py cfg1 cfg2 cfg3 cfg4 cfg5 2>&1 | tee $O/r1g_microbench_all_configs_final.log
echo "== done"
