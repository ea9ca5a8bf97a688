#!/bin/bash
# Round 2, session 3f: ncu captures of the kernels that bound cfg3 axis 0 and cfg5.
set -u
O=gpurun_out
mkdir -p $O
timeout -s KILL 300 ncu --set full --clock-control none --import-source on -k regex:pair -s 1 -c 1 -o $O/r3f_axis0_pair python tools/prof_target.py cfg3_axis0 2 > $O/r3f_ncu_axis0.log 2>&1
tail -2 $O/r3f_ncu_axis0.log
timeout -s KILL 300 ncu --set full --clock-control none --import-source on -k regex:pow2_kernel -s 2 -c 2 -o $O/r3f_cfg5_dct python tools/prof_target.py cfg5 2 > $O/r3f_ncu_cfg5.log 2>&1
tail -2 $O/r3f_ncu_cfg5.log
