#!/bin/bash
# GPU session: parity with the two-lines-per-thread strided kernel, A/B, configs 2/3, bench, ncu.
set -u
O=gpurun_out
mkdir -p $O
echo "== pytest -m gpu (all)"
timeout 500 python -m pytest tests -m gpu -q 2>&1 | tail -15 | tee $O/r1d_pytest_gpu.log
for v in 0 1; do
  echo "== microbench cfg2 cfg3 RFB200_PAIR=$v"
  RFB200_PAIR=$v timeout 200 python tools/microbench.py cfg2 2>&1 | grep -v cuFFT | tee $O/r1d_microbench_cfg2_pair$v.log
  RFB200_PAIR=$v timeout 200 python tools/microbench.py cfg3 2>&1 | tee $O/r1d_microbench_cfg3_pair$v.log
done
echo "== fused four-step on top (opt-in)"
RFB200_FUSE4=1 RFB200_FUSE4_COLS=128 timeout 120 python tools/microbench.py cfg2 2>&1 | grep -v cuFFT
echo "== bench default"
timeout 300 python bench.py > $O/r1d_bench_1gpu.json 2> $O/r1d_bench_1gpu.err
python - <<'P'
import json
for f in ("r1d_bench_1gpu",):
    try:
        d=json.load(open(f"gpurun_out/{f}.json")); print(f, round(d["value"]), d["ms_per_step"], d["roofline"], [(round(s["ms"],3), s["launches"]) for s in d["stages"]], d["clocks"], d.get("e2e",{}) and round(d["e2e"]["value"]), d["cpu_baseline"])
    except Exception as e: print(f, "unreadable", e)
P
echo "== ncu launch list"
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/r1d_ncu_launches_bench.csv python bench.py --steps 2 --warmup 1 --no-e2e --no-cpu > $O/r1d_ncu_bench.log 2>&1
tail -4 $O/r1d_ncu_launches_bench.csv | cut -c1-220
echo "== ncu full: rfft2 kernels"
timeout 400 ncu --set full --clock-control none --import-source on -c 3 -f -o $O/r1d_rfft2 python tools/prof_target.py rfft2 1 > $O/r1d_ncu_rfft2.log 2>&1
python tools/ncu_summarize.py $O/r1d_rfft2.ncu-rep > $O/r1d_ncu_rfft2_kernels.txt 2>&1
grep -E "^## kernel|gpu__time_duration|dram__bytes|smsp__inst_executed|issue_active|l1tex__throughput|gpu__dram_throughput|registers_per_thread" $O/r1d_ncu_rfft2_kernels.txt
echo "== done"
