#!/bin/bash
# round 2: the other BASELINE configs on N GPUs of one box (N = $1): configs[1] 2048 kp, configs[3] 4096 kp (256 pairs per
# step over 8 GPUs), configs[4] 8192 kp (fp32 path and the bf16 variant)
N=${1:-8}
mkdir -p gpurun_out
run() {  # name, args...
  name=$1; shift
  timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 \
    bench.py --gpus $N --no-cpu-baseline "$@" > gpurun_out/multi_${name}_n$N.json 2> gpurun_out/multi_${name}_n$N.err
  echo "$name rc=$?"
}
run 2048 --steps 5 --warmup 3
run 4096 --kpts 4096 --weights damped --pairs-per-step 32 --steps 3 --warmup 3
run 8192 --kpts 8192 --weights damped --pairs-per-step 8 --steps 3 --warmup 3
run 8192_bf16 --kpts 8192 --weights damped --pairs-per-step 8 --steps 3 --warmup 3 --gemm-mode bf16
python - <<PY
import json,glob
for f in sorted(glob.glob('gpurun_out/multi_*_n$N.json')):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        print(f, 'n_gpus', d['n_gpus'], 'value', round(d['value'],1), 'e2e', round(d['e2e']['value'],1), d['config'].get('parallelism'))
    except Exception as e: print(f, 'ERR', e)
PY
