#!/bin/bash
# Round 2, final single-GPU visit on the final code: parity suite, smoke, bench, sweep, variants, driver view, ncu passes.
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -p no:cacheprovider > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit=$? :: $(tail -n 1 gpurun_out/pytest_gpu.log)"
grep -E "^(FAILED|ERROR)" gpurun_out/pytest_gpu.log | head -40
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke exit=$? :: $(tail -n 1 gpurun_out/smoke.log)"
timeout 900 python bench.py --steps 50 --warmup 5 > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench exit=$?"; python -c "
import json; d=json.load(open('gpurun_out/bench.json')); print({k:d[k] for k in ('value','ms_per_step','parity','clocks')}); print(d['sustained']['value'], d['e2e']['value'], d['e2e_u8']['value'], d['cpu_baseline']['value'])"
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_reference.json 2> gpurun_out/bench_reference.err; echo "bench reference exit=$?"; cut -c1-300 gpurun_out/bench_reference.json
timeout 600 python scripts/batch_sweep.py > gpurun_out/batch_sweep.json 2> gpurun_out/batch_sweep.log; echo "sweep exit=$?"; cat gpurun_out/batch_sweep.log
timeout 600 python scripts/reference_driver_view.py > gpurun_out/reference_driver_view.json 2> gpurun_out/reference_driver_view.err; echo "driver view exit=$?"; cat gpurun_out/reference_driver_view.json | tr -d '\n' | cut -c1-1500; echo
for V in dynamic fp32 custom custom_sandwich; do
  timeout 300 python bench.py --variant $V --steps 20 --warmup 3 > gpurun_out/bench_$V.json 2> gpurun_out/bench_$V.err; echo "bench $V exit=$?"; python -c "
import json; d=json.load(open('gpurun_out/bench_$V.json')); print(d['value'], d['e2e']['value'], d['parity']['max_abs_err_over_logit_range'], d['roofline'] and d['roofline']['frac'], d['cpu_baseline']['value'])"
done
timeout 300 python scripts/prof_elementwise.py > gpurun_out/elementwise.json 2> gpurun_out/elementwise.log; echo "elementwise exit=$?"; cat gpurun_out/elementwise.log
timeout 600 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/elementwise_ncu.csv python scripts/prof_elementwise.py --once > gpurun_out/elementwise_ncu.log 2>&1; echo "ncu elementwise exit=$?"
SHORT="python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-parity --sustain-s 0"
timeout 600 $SHORT > gpurun_out/bench_short.json 2> gpurun_out/bench_short.err &&
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches.csv $SHORT > gpurun_out/ncu_launches.log 2>&1
echo "ncu launches exit=$?"
timeout 300 python scripts/prof_net.py > gpurun_out/prof_net_plain.log 2>&1 &&
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:'conv1_tc|conv_halo|conv_pair|igemm_tc|linear_simt|linear_head' -s 16 -c 8 -f -o gpurun_out/prof_net python scripts/prof_net.py > gpurun_out/prof_net_ncu.log 2>&1
echo "ncu net exit=$? :: $(tail -n 1 gpurun_out/prof_net_ncu.log)"
timeout 300 python scripts/prof_net.py --batch 1 --forwards 40 > gpurun_out/prof_b1_plain.log 2>&1 &&
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_b1.csv python scripts/prof_net.py --batch 1 --forwards 4 > gpurun_out/ncu_b1.log 2>&1
echo "ncu b1 exit=$?"
