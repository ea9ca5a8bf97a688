#!/bin/bash
# Round 2, session 3b (2 GPUs): bench.py at N=2 with per-rank NUMA binding; host topology.
set -u
O=gpurun_out
mkdir -p $O
( nvidia-smi topo -m; echo; lscpu | grep -i -E "numa|socket|model name|^CPU\(s\)"; echo; cat /sys/fs/cgroup/cpuset.cpus.effective 2>/dev/null; python -c "import os; print('affinity', sorted(os.sched_getaffinity(0)))"; numactl -H 2>/dev/null | head -20 ) > $O/r3b_topology.txt 2>&1
cat $O/r3b_topology.txt | head -60
( time timeout -s KILL 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 10 --warmup 3 --no-fftn ) > $O/r3b_bench_2gpu.json 2> $O/r3b_bench_2gpu.err
python - <<'PY'
import json
j=[json.loads(l) for l in open('gpurun_out/r3b_bench_2gpu.json') if l.startswith('{')][0]
e=j['e2e']; print({k:e[k] for k in e if k not in ('call',)})
PY
tail -3 $O/r3b_bench_2gpu.err
