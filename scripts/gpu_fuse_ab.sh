#!/bin/bash
# Fused conv1+conv2 kernel vs the two separate kernels, same build, same box, interleaved.
for rep in 1 2; do
  for V in 0 1; do
    B200Q_FUSE12=$V timeout 300 python bench.py --steps 20 --warmup 3 --stages-only 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('fuse12=$V ms/step %.3f ' % d['ms_per_step'], ' '.join('%s %.3f' % (k[:10], v['ms']) for k,v in d['roofline']['stages'].items()))"
  done
done
