timeout 700 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
timeout 200 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -1
timeout 600 python bench.py > gpurun_out/bench_r01_v5.json 2> gpurun_out/bench_r01_v5.err
python - <<PY
import json
d=json.load(open("gpurun_out/bench_r01_v5.json"))
print("value", d["value"], "ms", d["ms_per_step"], "e2e", d["e2e"]["value"], "frac", d["roofline"]["frac"])
print("train", d["dpo_train_step"])
PY
tail -2 gpurun_out/bench_r01_v5.err
