#!/bin/bash
mkdir -p gpurun_out
( timeout 900 python -m pytest tests/test_gpu_wan.py tests/test_gpu_t5.py "tests/test_gpu_parity_full.py::test_wan_ti2v_5b_forward_full_size" -x -q 2>&1 | tail -4 ) > gpurun_out/r02c24_tests.log 2>&1
timeout 900 python bench.py > gpurun_out/r02c24_bench.json 2> gpurun_out/r02c24_bench.err
cat gpurun_out/r02c24_tests.log; python - <<'PY'
import json
d=json.load(open("gpurun_out/r02c24_bench.json"))
print(d["value"], d["ms_per_step"], d["roofline"]["frac"], d["kernel_families"]["ln_modulate_kernel"], d["wan_step"]["ms_per_step"], d["encoders"]["t5_xxl_ms_per_prompt"], d["vae_decode"]["ms_per_clip"], d["dpo_train_step"]["ms_per_step"], d["secondary"]["value"])
PY
tail -2 gpurun_out/r02c24_bench.err
