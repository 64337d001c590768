#!/bin/bash
# e2e (host input -> host logits) under different chunkings, same box, interleaved twice.
for rep in 1 2; do
for CFG in "2048 256" "2048 100000" "4096 256"; do set -- $CFG
B200Q_HOST_CHUNK=$1 B200Q_TAIL_MIN=$2 python bench.py --steps 20 --warmup 3 --no-cpu-baseline 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('chunk $1 tail_min $2: value %.0f e2e %.0f (%.3f ms/step e2e)'%(d['value'],d['e2e']['value'],16384e3/d['e2e']['value']))"; done; done
