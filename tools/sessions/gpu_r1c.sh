#!/bin/bash
# GPU session: parity of the fused four-step kernel (+ everything else), knob sweep, bench line, ncu.
set -u
O=gpurun_out
mkdir -p $O
echo "== fused four-step parity (with the dependency check on)"
RFB200_FUSE4_CHECK=1 timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -q -k "fused_fourstep or long_real or device_arrays" 2>&1 | tail -15 | tee $O/r1c_pytest_fused.log
echo "== pytest -m gpu (all)"
timeout 420 python -m pytest tests -m gpu -q 2>&1 | tail -15 | tee $O/r1c_pytest_gpu.log
echo "== sweep"
run() { echo "-- $*"; env "$@" timeout 120 python tools/microbench.py cfg2 2>&1 | grep -v "^rocketfft\|cuFFT"; }
( run RFB200_FUSE4=0
  run RFB200_FUSE4_COLS=64 RFB200_FUSE4_RING=4 RFB200_FUSE4_LAG=2
  run RFB200_FUSE4_COLS=128 RFB200_FUSE4_RING=4 RFB200_FUSE4_LAG=2
  run RFB200_FUSE4_COLS=256 RFB200_FUSE4_RING=4 RFB200_FUSE4_LAG=2
  run RFB200_FUSE4_COLS=64 RFB200_FUSE4_RING=6 RFB200_FUSE4_LAG=3
  run RFB200_FUSE4_COLS=128 RFB200_FUSE4_RING=6 RFB200_FUSE4_LAG=3
  run RFB200_FUSE4_COLS=96 RFB200_FUSE4_RING=5 RFB200_FUSE4_LAG=2
  run RFB200_FUSE4_COLS=128 RFB200_FUSE4_RING=3 RFB200_FUSE4_LAG=1 ) 2>&1 | tee $O/r1c_sweep_fuse4.log
echo "== bench default"
timeout 300 python bench.py > $O/r1c_bench_1gpu.json 2> $O/r1c_bench_1gpu.err
RFB200_FUSE4=0 timeout 200 python bench.py --steps 30 --no-e2e --no-cpu > $O/r1c_bench_fuse0.json 2>/dev/null
python - <<'P'
import json
for f in ("r1c_bench_1gpu","r1c_bench_fuse0"):
    try:
        d=json.load(open(f"gpurun_out/{f}.json")); print(f, round(d["value"]), d["ms_per_step"], d["roofline"]["kernel"], round(d["roofline"]["frac"],3), [(round(s["ms"],3), s["launches"]) for s in d["stages"]], d["clocks"], d.get("e2e",{}) and round(d["e2e"]["value"]))
    except Exception as e: print(f, "unreadable", e)
P
echo "== ncu launch list"
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/r1c_ncu_launches_bench.csv python bench.py --steps 2 --warmup 1 --no-e2e --no-cpu > $O/r1c_ncu_bench.log 2>&1
tail -8 $O/r1c_ncu_launches_bench.csv | cut -c1-200
echo "== ncu full: fused column kernel"
timeout 400 ncu --set full --clock-control none --import-source on -k regex:fourstep_fused -c 1 -f -o $O/r1c_cols_fused python tools/prof_target.py cols 2 > $O/r1c_ncu_cols.log 2>&1
python tools/ncu_summarize.py $O/r1c_cols_fused.ncu-rep > $O/r1c_ncu_cols_fused.txt 2>&1
head -40 $O/r1c_ncu_cols_fused.txt
echo "== done"
