#!/bin/bash
# Round 2, first GPU visit: parity tests (all, no -x), smoke, int8 peak probe, bench, small-batch sweep, variants,
# element-wise profile (+ ncu DRAM bytes).
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,driver_version,clocks.max.sm,memory.total --format=csv > gpurun_out/gpu.txt 2>&1
nproc > gpurun_out/host.txt; lscpu | grep -E "Model name|^CPU\(s\)|NUMA" >> gpurun_out/host.txt
timeout 1500 python -m pytest tests -m gpu -q -p no:cacheprovider > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit=$? :: $(tail -n 1 gpurun_out/pytest_gpu.log)"
grep -E "^(FAILED|ERROR)" gpurun_out/pytest_gpu.log | head -40
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke exit=$? :: $(tail -n 1 gpurun_out/smoke.log)"
timeout 120 tools/bin/int8_peak > gpurun_out/int8_peak.json 2> gpurun_out/int8_peak.err; echo "int8_peak exit=$?"; cat gpurun_out/int8_peak.json
timeout 900 python bench.py --steps 50 --warmup 5 > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench exit=$?"; cat gpurun_out/bench.json; tail -3 gpurun_out/bench.err
timeout 600 python scripts/batch_sweep.py > gpurun_out/batch_sweep.json 2> gpurun_out/batch_sweep.log; echo "sweep exit=$?"; cat gpurun_out/batch_sweep.log
B200Q_LIB=$PWD/convnet_quantization_b200/libb200q_dev.so B200Q_NO_HALO=1 timeout 300 python scripts/batch_sweep.py --batches 1,8,32,128,512 > gpurun_out/batch_sweep_nohalo.json 2> gpurun_out/batch_sweep_nohalo.log; echo "sweep nohalo exit=$?"; cat gpurun_out/batch_sweep_nohalo.log
for V in dynamic fp32 custom custom_sandwich; do
  timeout 300 python bench.py --variant $V --steps 20 --warmup 3 > gpurun_out/bench_$V.json 2> gpurun_out/bench_$V.err; echo "bench $V exit=$?"; cat gpurun_out/bench_$V.json; tail -2 gpurun_out/bench_$V.err
done
timeout 300 python scripts/prof_elementwise.py > gpurun_out/elementwise.json 2> gpurun_out/elementwise.log; echo "elementwise exit=$?"; cat gpurun_out/elementwise.log
timeout 600 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/elementwise_ncu.csv python scripts/prof_elementwise.py --once > gpurun_out/elementwise_ncu.log 2>&1; echo "ncu elementwise exit=$?"
