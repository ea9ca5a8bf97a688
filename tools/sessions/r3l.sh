#!/bin/bash
# Round 2, session 3l (2 GPUs): padded plan buffers of the slab transforms: parity (all engines, complex + real), bench.
set -u
O=gpurun_out
mkdir -p $O
( time timeout -s KILL 600 python -m pytest tests/test_gpu_distributed.py -x -q -m gpu ) > $O/r3l_pytest_2gpu.log 2>&1
tail -12 $O/r3l_pytest_2gpu.log
( time timeout -s KILL 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 10 --warmup 3 --no-e2e ) > $O/r3l_bench_2gpu.json 2> $O/r3l_bench_2gpu.err
python - <<'PY'
import json
j=[json.loads(l) for l in open('gpurun_out/r3l_bench_2gpu.json') if l.startswith('{')][0]
f=j['fftn']; print({k:f[k] for k in ('ms','alltoall_ms','bus_GBps','parity_rel_l2')}); print(json.dumps(f.get('rfftn'), indent=1))
PY
tail -3 $O/r3l_bench_2gpu.err
