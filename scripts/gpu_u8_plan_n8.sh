#!/bin/bash
# uint8 host route at N ranks under different chunk ramps (first, cap): bash scripts/gpu_u8_plan_n8.sh N
N=${1:-8}
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
P=29600
for rep in 1 2; do
for CFG in "1024 8192" "1024 4096" "2048 4096" "2048 2048"; do set -- $CFG
P=$((P+1))
B200Q_U8_FIRST_CHUNK=$1 B200Q_U8_MAX_CHUNK=$2 timeout 300 $TR --master-port $P bench.py --gpus $N --steps 10 --warmup 3 --no-cpu-baseline --no-parity --sustain-s 0 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('first $1 cap $2: value %.0f e2e %.0f e2e_u8 %.0f'%(d['value'],d['e2e']['value'],d['e2e_u8']['value']))" | tee -a gpurun_out/u8_plan_n$N.txt
done; done
