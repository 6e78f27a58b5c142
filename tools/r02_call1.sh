#!/bin/bash
# Round-2 GPU validation call: GPU tests, smoke, bench (N = 1), ncu launch list of a short bench, ncu --set full of the
# top kernels. Everything lands in gpurun_out/.
mkdir -p gpurun_out
( timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -15 ) > gpurun_out/r02_gputests.log 2>&1
( timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3 ) > gpurun_out/r02_smoke.log 2>&1
timeout 900 python bench.py > gpurun_out/r02_bench_a.json 2> gpurun_out/r02_bench_a.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file gpurun_out/r02_launches.csv \
    python bench.py --layers 4 --steps 1 --warmup 1 --no-cpu-baseline --no-gpu-eager > gpurun_out/r02_launches_bench.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:attn_fwd_d64_bounded -s 2 -c 1 -f -o gpurun_out/r02_attn \
    python tools/dev_attn.py bench > gpurun_out/r02_attn_ncu.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"gemm_bf16_kernel|mvcs|ln_modulate" -s 4 -c 6 -f -o gpurun_out/r02_kernels \
    python tools/dev_profile_kernels.py > gpurun_out/r02_kernels_ncu.log 2>&1
python tools/dev_attn.py bench > gpurun_out/r02_attn_bench.log 2>&1
tail -5 gpurun_out/r02_gputests.log; cat gpurun_out/r02_smoke.log; head -c 1500 gpurun_out/r02_bench_a.json; tail -3 gpurun_out/r02_bench_a.err
cat gpurun_out/r02_attn_bench.log
