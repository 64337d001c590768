#!/bin/bash
# Round 2, second GPU visit: full parity suite again (fma activation quantisation, small-batch kernels), sweep, bench.
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -p no:cacheprovider > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit=$? :: $(tail -n 1 gpurun_out/pytest_gpu.log)"
grep -E "^(FAILED|ERROR)" gpurun_out/pytest_gpu.log | head -40
timeout 600 python scripts/batch_sweep.py > gpurun_out/batch_sweep.json 2> gpurun_out/batch_sweep.log; echo "sweep exit=$?"; cat gpurun_out/batch_sweep.log
B200Q_LIB=$PWD/convnet_quantization_b200/libb200q_dev.so B200Q_NO_SMALL=1 timeout 300 python scripts/batch_sweep.py --batches 1,8,32,64,128,256 > gpurun_out/batch_sweep_nosmall.json 2> gpurun_out/batch_sweep_nosmall.log; echo "sweep nosmall exit=$?"; cat gpurun_out/batch_sweep_nosmall.log
timeout 900 python bench.py --steps 50 --warmup 5 > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench exit=$?"; python -c "
import json; d=json.load(open('gpurun_out/bench.json')); print({k:d[k] for k in ('value','ms_per_step','parity','sustained')}); print(d['e2e']['value'], d['e2e_u8']['value'], d['config'].get('host_binding_rank0'))"
timeout 300 python bench.py --variant dynamic --steps 20 --warmup 3 > gpurun_out/bench_dynamic.json 2> gpurun_out/bench_dynamic.err; echo "bench dynamic exit=$?"; cat gpurun_out/bench_dynamic.json
timeout 120 python scripts/h2d_probe.py > gpurun_out/h2d_n1.json 2> gpurun_out/h2d_n1.err; echo "h2d exit=$?"; cat gpurun_out/h2d_n1.json
