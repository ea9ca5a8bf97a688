#!/bin/bash
# GPU session: the cp.async-staged persistent kernel for strided 128/256-point lines (RFB200_LFMODE=2): parity, A/B, ncu.
set -u
O=gpurun_out
mkdir -p $O
echo "== parity with RFB200_LFMODE=2"
RFB200_LFMODE=2 timeout 400 python -m pytest tests/test_gpu_parity.py -m gpu -q 2>&1 | tail -15 | tee $O/r1e_pytest_lfmode2.log
for cfg in "RFB200_LFMODE=1" "RFB200_LFMODE=2" "RFB200_LFMODE=2 RFB200_ASYNC_CTAS=2" "RFB200_LFMODE=2 RFB200_ASYNC_CTAS=6" "RFB200_LFMODE=0"; do
  echo "== microbench cfg2 $cfg"
  env $cfg timeout 200 python tools/microbench.py cfg2 2>&1 | grep -v "cuFFT\|^rocketfft" | tee -a $O/r1e_microbench_cfg2_lfmode.log
done
echo "== bench LFMODE=2 (short)"
RFB200_LFMODE=2 timeout 200 python bench.py --steps 30 --no-e2e --no-cpu > $O/r1e_bench_lfmode2.json 2>/dev/null
python - <<'P'
import json
for f in ("r1e_bench_lfmode2",):
    try:
        d=json.load(open(f"gpurun_out/{f}.json")); print(f, round(d["value"]), d["ms_per_step"], round(d["roofline"]["frac"],3), [(round(s["ms"],3), s["launches"]) for s in d["stages"]], d["clocks"])
    except Exception as e: print(f, "unreadable", e)
P
echo "== ncu full: async column kernels"
RFB200_LFMODE=2 timeout 400 ncu --set full --clock-control none --import-source on -k regex:async -c 2 -f -o $O/r1e_cols_async python tools/prof_target.py cols 1 > $O/r1e_ncu_cols.log 2>&1
python tools/ncu_summarize.py $O/r1e_cols_async.ncu-rep > $O/r1e_ncu_cols_async.txt 2>&1
grep -E "^## kernel|gpu__time_duration|dram__bytes|smsp__inst_executed|issue_active|l1tex__throughput|gpu__dram_throughput|registers_per_thread|warps_active|^    \{" $O/r1e_ncu_cols_async.txt | head -40
echo "== done"
