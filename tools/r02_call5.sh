#!/bin/bash
mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/r02c5_gpus.log
SECONDS=0
timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 \
   > gpurun_out/r02c5_bench_n2.json 2> gpurun_out/r02c5_bench_n2.err
echo "bench n2 rc=$? wall=${SECONDS}s" >> gpurun_out/r02c5_gpus.log
SECONDS=0
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --impl reference --steps 3 --warmup 1 \
   > gpurun_out/r02c5_ref_n2.json 2> gpurun_out/r02c5_ref_n2.err
echo "ref n2 rc=$? wall=${SECONDS}s" >> gpurun_out/r02c5_gpus.log
cat gpurun_out/r02c5_gpus.log; tail -5 gpurun_out/r02c5_bench_n2.err; head -c 600 gpurun_out/r02c5_bench_n2.json; echo; python - <<'PY'
import json
d=json.loads(open("gpurun_out/r02c5_bench_n2.json").read().strip().splitlines()[-1])
print(json.dumps(d.get("multi_gpu"), indent=1)[:6000])
PY
head -c 400 gpurun_out/r02c5_ref_n2.json
