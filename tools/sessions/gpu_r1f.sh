#!/bin/bash
# Final GPU session of the round: full parity suite with the default library, prefetch-distance sweep for the row
# kernel, bench (own arm + reference arm), ncu launch list.
set -u
O=gpurun_out
mkdir -p $O
echo "== pytest -m gpu (all, defaults)"
timeout 500 python -m pytest tests -m gpu -q 2>&1 | tail -15 | tee $O/r1f_pytest_gpu_final.log
echo "== L2 prefetch distance sweep, row kernel"
for pf in 0 74 148 296 444 592; do
  echo "-- RFB200_PF=$pf"; RFB200_PF=$pf timeout 100 python tools/microbench.py cfg2 2>&1 | grep "rows\|whole"
done 2>&1 | tee $O/r1f_pf_sweep_rows.log
echo "== bench default"
timeout 300 python bench.py > $O/r1f_bench_1gpu.json 2> $O/r1f_bench_1gpu.err
timeout 300 python bench.py --impl reference > $O/r1f_bench_reference_arm.json 2>/dev/null
python - <<'P'
import json
d=json.load(open("gpurun_out/r1f_bench_1gpu.json")); print(round(d["value"]), d["ms_per_step"], d["roofline"], [(round(s["ms"],3), s["launches"]) for s in d["stages"]], d["clocks"], d["e2e"], d["cpu_baseline"], d["gpu_launches"])
d=json.load(open("gpurun_out/r1f_bench_reference_arm.json")); print("reference arm", round(d["value"],1), d["cpu_baseline"])
P
echo "== ncu launch list"
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/r1f_ncu_launches_bench.csv python bench.py --steps 2 --warmup 1 --no-e2e --no-cpu > $O/r1f_ncu_bench.log 2>&1
python - <<'P'
import csv,collections
rows=[r for r in csv.reader(open("gpurun_out/r1f_ncu_launches_bench.csv")) if len(r)>10 and r[0].isdigit()]
agg=collections.defaultdict(lambda:[0,0.0])
for r in rows:
    k=r[4][:70]; agg[k][0]+=1; agg[k][1]+=float(r[-1].replace(",",""))
tot=sum(v[1] for k,v in agg.items() if "rfb::" in k)
for k,v in sorted(agg.items(), key=lambda kv:-kv[1][1])[:8]: print(f"{v[0]:4d} launches {v[1]/1e3:10.1f} us  {100*v[1]/tot if 'rfb::' in k else 0:5.1f}%  {k}")
P
echo "== done"
