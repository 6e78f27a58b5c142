#!/bin/bash
mkdir -p gpurun_out
N=${1:-4}
nvidia-smi -L > gpurun_out/r02c11_gpus.log
SECONDS=0
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29521 bench.py --gpus $N \
   > gpurun_out/r02c11_bench_n$N.json 2> gpurun_out/r02c11_bench_n$N.err
echo "bench n$N rc=$? wall=${SECONDS}s" >> gpurun_out/r02c11_gpus.log
cat gpurun_out/r02c11_gpus.log; tail -3 gpurun_out/r02c11_bench_n$N.err | cut -c1-300; python - <<PY
import json
d=json.loads(open("gpurun_out/r02c11_bench_n$N.json").read().strip().splitlines()[-1])
print({k:d[k] for k in ['value','ms_per_step','n_gpus']}, d['e2e']['value'])
print(json.dumps(d.get("multi_gpu"), indent=1)[:6000])
PY
