#!/bin/bash
# Quick GPU iteration: conv/net parity + bench summary (no CPU baseline, no ncu).
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_conv.py tests/test_gpu_net.py tests/test_gpu_elementwise.py -m gpu -x -q 2>&1 | tail -8
timeout 600 python bench.py --steps ${STEPS:-30} --warmup 5 --no-cpu-baseline ${BENCH_ARGS} > gpurun_out/bench_quick.json 2> gpurun_out/bench_quick.err || tail -5 gpurun_out/bench_quick.err
python - <<'PY'
import json
d=json.loads(open("gpurun_out/bench_quick.json").read())
print("value %.0f img/s  e2e %.0f  ms/step %.3f  clocks %s" % (d["value"], d["e2e"]["value"], d["ms_per_step"], d["clocks"]))
for k,v in d["roofline"]["stages"].items(): print("  %-12s %.3f ms  frac %.3f" % (k, v["ms"], v["frac"]))
PY
