#!/bin/bash
# Round 2, session 3i (2 GPUs): scatter output through the two-lines-per-thread kernel: slab checks and the bench's fftn.
set -u
O=gpurun_out
mkdir -p $O
( time timeout -s KILL 600 python -m pytest tests/test_gpu_distributed.py tests/test_gpu_parity.py -x -q -m gpu -k "slab" ) > $O/r3i_pytest_2gpu.log 2>&1
tail -5 $O/r3i_pytest_2gpu.log
( time timeout -s KILL 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 10 --warmup 3 --no-e2e ) > $O/r3i_bench_2gpu.json 2> $O/r3i_bench_2gpu.err
python - <<'PY'
import json
j=[json.loads(l) for l in open('gpurun_out/r3i_bench_2gpu.json') if l.startswith('{')][0]
print(json.dumps(j['fftn'], indent=1))
PY
tail -3 $O/r3i_bench_2gpu.err
