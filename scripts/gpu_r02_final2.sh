#!/bin/bash
# Round 2, last visit: full GPU suite, smoke, default bench, batch sweep, chain breakdown, the reference driver's view.
mkdir -p gpurun_out
DEV=$PWD/convnet_quantization_b200/libb200q_dev.so
timeout 1500 python -m pytest tests -m gpu -q -p no:cacheprovider > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit=$? :: $(tail -n 1 gpurun_out/pytest_gpu.log)"
grep -E "^(FAILED|ERROR)" gpurun_out/pytest_gpu.log | head -40
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke exit=$? :: $(tail -n 1 gpurun_out/smoke.log)"
timeout 900 python bench.py --steps 50 --warmup 5 > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench exit=$?"; python -c "
import json; d=json.load(open('gpurun_out/bench.json')); print({k:d[k] for k in ('value','ms_per_step','parity','clocks')}); print(d['sustained']['value'], d['e2e']['value'], d['e2e_u8']['value'], d['cpu_baseline']['value'])"
timeout 600 python scripts/batch_sweep.py > gpurun_out/batch_sweep.json 2> gpurun_out/batch_sweep.log; echo "sweep exit=$?"; cat gpurun_out/batch_sweep.log
B200Q_LIB=$DEV timeout 600 python scripts/graph_breakdown.py > gpurun_out/graph_breakdown.json 2> gpurun_out/graph_breakdown.log; echo "breakdown exit=$?"; cat gpurun_out/graph_breakdown.log
timeout 600 python scripts/reference_driver_view.py > gpurun_out/reference_driver_view.json 2> gpurun_out/reference_driver_view.err; echo "driver view exit=$?"; cat gpurun_out/reference_driver_view.json | tr -d '\n' | cut -c1-1500; echo
timeout 300 python scripts/prof_net.py --batch 1 --forwards 40 > gpurun_out/prof_b1_plain.log 2>&1 &&
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_b1.csv python scripts/prof_net.py --batch 1 --forwards 4 > gpurun_out/ncu_b1.log 2>&1
echo "ncu b1 exit=$?"
