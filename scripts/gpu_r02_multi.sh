#!/bin/bash
# Round 2 multi-GPU visit: bash scripts/gpu_r02_multi.sh N   (run under `gpurun --gpus N`)
# Concurrent pinned H2D bandwidth per GPU (with / without NUMA binding) and the bench at N ranks.
N=$1
mkdir -p gpurun_out
nvidia-smi topo -m > gpurun_out/topo_n$N.txt 2>&1
lscpu | grep -E "Model name|^CPU\(s\)|NUMA|Socket" > gpurun_out/host_n$N.txt; nproc >> gpurun_out/host_n$N.txt; cat /sys/fs/cgroup/cpuset.cpus.effective >> gpurun_out/host_n$N.txt 2>/dev/null
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
timeout 300 $TR --master-port 29511 scripts/h2d_probe.py --out=gpurun_out/h2d_n$N.json > gpurun_out/h2d_n$N.log 2> gpurun_out/h2d_n$N.err; echo "h2d exit=$?"; python -c "
import json; d=json.load(open('gpurun_out/h2d_n$N.json')); print(d['aggregate_gbs'], d['min_per_gpu_gbs']); print([r['numa'] for r in d['ranks']][:2])"
timeout 300 $TR --master-port 29512 scripts/h2d_probe.py --no-bind --out=gpurun_out/h2d_n${N}_nobind.json > gpurun_out/h2d_n${N}_nobind.log 2> gpurun_out/h2d_n${N}_nobind.err; echo "h2d nobind exit=$?"; python -c "
import json; d=json.load(open('gpurun_out/h2d_n${N}_nobind.json')); print(d['aggregate_gbs'], d['min_per_gpu_gbs'])"
timeout 600 $TR --master-port 29513 bench.py --gpus $N --steps 20 --warmup 5 --sustain-s 0 > gpurun_out/bench_n$N.json 2> gpurun_out/bench_n$N.err; echo "bench exit=$?"; python -c "
import json; d=json.load(open('gpurun_out/bench_n$N.json')); print('value', d['value'], 'e2e', d['e2e']['value'], 'e2e_u8', d['e2e_u8']['value'], d['parity'], d['eval'])"; tail -2 gpurun_out/bench_n$N.err
if [ "$N" = "2" ]; then
  timeout 600 python -m pytest tests/test_gpu_net.py tests/test_gpu_models.py tests/test_reference_drivers.py -m gpu -q -p no:cacheprovider > gpurun_out/pytest_gpu_n2.log 2>&1; echo "pytest exit=$? :: $(tail -n 1 gpurun_out/pytest_gpu_n2.log)"
  grep -E "^(FAILED|ERROR)" gpurun_out/pytest_gpu_n2.log | head
  timeout 300 python scripts/batch_sweep.py --batches 1,2,4,8,16,32,64,65,128 > gpurun_out/batch_sweep_small.json 2> gpurun_out/batch_sweep_small.log; echo "sweep exit=$?"; cat gpurun_out/batch_sweep_small.log
fi
