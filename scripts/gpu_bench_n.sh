#!/bin/bash
# The bench line at N ranks, nothing else: bash scripts/gpu_bench_n.sh N   (under `gpurun --gpus N`)
N=$1
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
timeout 600 $TR --master-port 29513 bench.py --gpus $N --steps 20 --warmup 5 --sustain-s 0 > gpurun_out/bench_n$N.json 2> gpurun_out/bench_n$N.err; echo "bench exit=$?"; python -c "
import json; d=json.load(open('gpurun_out/bench_n$N.json')); print('value', d['value'], 'e2e', d['e2e']['value'], 'e2e_u8', d['e2e_u8']['value'], d['parity']['bit_exact'], d['eval'])"; tail -2 gpurun_out/bench_n$N.err
