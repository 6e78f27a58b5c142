#!/bin/bash
mkdir -p gpurun_out
timeout 300 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum --clock-control none -k regex:"gemm_bf16|t5_|gated|rmsnorm|linear_small" -s 18 -c 27 --csv --log-file gpurun_out/r02c9_t5_launches.csv python tools/dev_t5.py > /dev/null 2>&1
python - <<'PY'
import csv,re
rows=[r for r in csv.reader(open("gpurun_out/r02c9_t5_launches.csv")) if len(r)>10]
h=rows[0]; ki,mi,vi,gi=h.index("Kernel Name"),h.index("Metric Name"),h.index("Metric Value"),h.index("Grid Size")
cur={}
for r in rows[1:]:
    key=(r[0],r[ki][:70],r[gi]); cur.setdefault(key,{})[r[mi]]=r[vi]
for k,v in cur.items(): print(k[1],k[2],v)
PY
