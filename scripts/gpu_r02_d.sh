#!/bin/bash
# Round 2, fourth GPU visit: parity suite (bit-exact linear_dynamic, fused small-batch head), flip stats, sweep, element-wise.
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -p no:cacheprovider > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit=$? :: $(tail -n 1 gpurun_out/pytest_gpu.log)"
grep -E "^(FAILED|ERROR)" gpurun_out/pytest_gpu.log | head -40
timeout 600 python scripts/dynamic_flip_stats.py --images 4096 > gpurun_out/dynamic_flip_stats.json 2> gpurun_out/dynamic_flip_stats.err; echo "flip stats exit=$?"; python -c "
import json; d=json.load(open('gpurun_out/dynamic_flip_stats.json')); print(d['full_gpu_vs_cpu']); print(d['gpu_linears_on_cpu_features_vs_cpu'])"
timeout 600 python scripts/batch_sweep.py --batches 1,2,4,8,16,32,33,64,128,256,512,1024,2048 > gpurun_out/batch_sweep.json 2> gpurun_out/batch_sweep.log; echo "sweep exit=$?"; cat gpurun_out/batch_sweep.log
timeout 300 python scripts/prof_elementwise.py > gpurun_out/elementwise.json 2> gpurun_out/elementwise.log; echo "elementwise exit=$?"; tail -3 gpurun_out/elementwise.log
timeout 300 python bench.py --variant dynamic --steps 20 --warmup 3 > gpurun_out/bench_dynamic.json 2> gpurun_out/bench_dynamic.err; echo "bench dynamic exit=$?"; python -c "
import json; d=json.load(open('gpurun_out/bench_dynamic.json')); print(d['value'], d['parity'], d['roofline'])"
