#!/bin/bash
# One GPU session: parity of the two-transform real kernel, A/B against the single-transform kernel,
# bench line, ncu launch list and full captures.  Everything lands in gpurun_out/.
set -u
O=gpurun_out
mkdir -p $O
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > $O/r1b_smi.txt 2>&1
echo "== pytest -m gpu" 
timeout 420 python -m pytest tests -m gpu -q 2>&1 | tail -15 | tee $O/r1b_pytest_gpu.log
for v in 0 1 2; do
  echo "== microbench cfg2 RFB200_DUAL=$v"
  RFB200_DUAL=$v timeout 120 python tools/microbench.py cfg2 2>&1 | tee $O/r1b_microbench_cfg2_dual$v.log
done
echo "== bench default"
timeout 300 python bench.py > $O/r1b_bench_1gpu.json 2> $O/r1b_bench_1gpu.err; tail -c 1500 $O/r1b_bench_1gpu.json
echo "== bench DUAL=2 / DUAL=0 (short)"
RFB200_DUAL=2 timeout 200 python bench.py --steps 30 --no-e2e --no-cpu > $O/r1b_bench_dual2.json 2>/dev/null
RFB200_DUAL=0 timeout 200 python bench.py --steps 30 --no-e2e --no-cpu > $O/r1b_bench_dual0.json 2>/dev/null
python - <<'P'
import json
for f in ("r1b_bench_1gpu","r1b_bench_dual2","r1b_bench_dual0"):
    try:
        d=json.load(open(f"gpurun_out/{f}.json")); print(f, round(d["value"]), d["ms_per_step"], d["roofline"]["kernel"], round(d["roofline"]["frac"],3), [round(s["ms"],3) for s in d["stages"]], d["clocks"])
    except Exception as e: print(f, "unreadable", e)
P
echo "== ncu launch list"
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/r1b_ncu_launches_bench.csv python bench.py --steps 2 --warmup 1 --no-e2e --no-cpu > $O/r1b_ncu_bench.log 2>&1
for v in 1 2; do
  echo "== ncu full rows DUAL=$v"
  RFB200_DUAL=$v timeout 300 ncu --set full --clock-control none --import-source on -k regex:dual -c 1 -f -o $O/r1b_rows_dual$v python tools/prof_target.py rows 2 > $O/r1b_ncu_rows$v.log 2>&1
  python tools/ncu_summarize.py $O/r1b_rows_dual$v.ncu-rep > $O/r1b_ncu_rows_dual$v.txt 2>&1
  head -30 $O/r1b_ncu_rows_dual$v.txt
done
echo "== done"
