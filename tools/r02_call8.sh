#!/bin/bash
mkdir -p gpurun_out
( timeout 900 python -m pytest tests/test_gpu_t5.py tests/test_gpu_vae.py tests/test_gpu_dit.py tests/test_gpu_wan.py -x -q 2>&1 | tail -6 ) > gpurun_out/r02c8_tests.log 2>&1
for b in 0 1; do VGPA_GEMM_BN128=$b timeout 300 python tools/dev_t5.py 2>&1 | tail -3 | sed "s/^/BN128=$b /"; done > gpurun_out/r02c8_t5.log 2>&1
cat gpurun_out/r02c8_tests.log gpurun_out/r02c8_t5.log
timeout 300 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum --clock-control none -s 40 -c 40 --csv --log-file gpurun_out/r02c8_t5_launches.csv python tools/dev_t5.py > /dev/null 2>&1
python - <<'PY'
import csv,re
rows=[r for r in csv.reader(open("gpurun_out/r02c8_t5_launches.csv")) if len(r)>10]
h=rows[0]; ki,mi,vi,gi=h.index("Kernel Name"),h.index("Metric Name"),h.index("Metric Value"),h.index("Grid Size")
cur={}
for r in rows[1:]:
    key=(r[0],r[ki][:60],r[gi]); cur.setdefault(key,{})[r[mi]]=r[vi]
for k,v in cur.items(): print(k[1],k[2],v)
PY
