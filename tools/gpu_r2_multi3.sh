#!/bin/bash
N=${1:-8}
mkdir -p gpurun_out
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29521 bench.py --gpus $N --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/multi3_2048_n$N.json 2> gpurun_out/multi3_2048_n$N.err; echo "rc=$?"
python - <<PY
import json
d=json.loads(open('gpurun_out/multi3_2048_n$N.json').read().strip().splitlines()[-1]); print(d.get('n_gpus'), round(d['value'],1), 'e2e', round(d['e2e']['value'],1), d['clocks'])
PY
