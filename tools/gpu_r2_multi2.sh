#!/bin/bash
N=${1:-2}
mkdir -p gpurun_out
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29519 bench.py --gpus $N --steps 5 --warmup 3 > gpurun_out/multi2_2048_n$N.json 2> gpurun_out/multi2_2048_n$N.err; echo "rc=$?"
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29519 bench.py --gpus $N --steps 2 --warmup 1 --impl reference > gpurun_out/multi2_ref_n$N.json 2> gpurun_out/multi2_ref_n$N.err; echo "ref rc=$?"
python - <<PY
import json
for f in ['gpurun_out/multi2_2048_n$N.json','gpurun_out/multi2_ref_n$N.json']:
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1]); print(f, d.get('n_gpus'), round(d['value'],2), 'e2e', d['e2e']['value'], d.get('impl'))
    except Exception as e: print(f,'ERR',e)
PY
