#!/bin/bash
# One GPU-box visit: parity tests, smoke, bench, then the ncu launch list of a short bench (same command run plain first).
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,driver_version,clocks.max.sm,memory.total --format=csv > gpurun_out/gpu.txt 2>&1
nproc > gpurun_out/host.txt; lscpu | grep -E "Model name|^CPU\(s\)" >> gpurun_out/host.txt
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit=$? :: $(tail -n 1 gpurun_out/pytest_gpu.log)"
timeout 600 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke exit=$? :: $(tail -n 1 gpurun_out/smoke.log)"
timeout 900 python bench.py --steps ${STEPS:-50} --warmup 5 > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench exit=$?"; cat gpurun_out/bench.json
SHORT="python bench.py --steps 2 --warmup 3 --no-cpu-baseline"
timeout 600 $SHORT > gpurun_out/bench_short.json 2> gpurun_out/bench_short.err &&
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches.csv $SHORT > gpurun_out/ncu_launches.log 2>&1
echo "ncu launches exit=$?"
# ncu --set full of the 8 kernels of one forward at the bench batch (after the same command ran plain)
timeout 300 python scripts/prof_net.py > gpurun_out/prof_net_plain.log 2>&1 &&
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:'conv1_tc|conv_halo|conv_pair|igemm_tc|linear_simt|linear_head' -s 16 -c 8 -f -o gpurun_out/prof_net python scripts/prof_net.py > gpurun_out/prof_net_ncu.log 2>&1
echo "ncu net exit=$? :: $(tail -n 1 gpurun_out/prof_net_ncu.log)"
