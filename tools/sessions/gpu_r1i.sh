#!/bin/bash
# Closing GPU session of round 1: the whole parity suite against the final library, then the bench line with 8 images per GPU.
set -u
O=gpurun_out
mkdir -p $O
echo "== pytest -m gpu (all, final library)"
timeout 400 python -m pytest tests -m gpu -q 2>&1 | tail -8 | tee $O/r1i_pytest_gpu_final.log
echo "== bench --images 8"
timeout 200 python bench.py --images 8 > $O/r1i_bench_1gpu_8img.json 2> $O/r1i_bench.err
free -g | head -2
python - <<'P'
import json
d=json.load(open("gpurun_out/r1i_bench_1gpu_8img.json")); print(round(d["value"]), d["ms_per_step"], round(d["roofline"]["frac"],3), d["roofline"]["kernel"][:40], [(round(s["ms"],3), s["launches"]) for s in d["stages"]], d["clocks"], d["e2e"], d["cpu_baseline"]["value"])
P
echo "== done"
