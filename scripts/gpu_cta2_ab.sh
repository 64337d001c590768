#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_conv.py -m gpu -q -p no:cacheprovider -x -k "conv2" > gpurun_out/pytest_cta2_conv.log 2>&1; echo "pytest conv2 exit=$? :: $(tail -n 1 gpurun_out/pytest_cta2_conv.log)"
DEV=$PWD/convnet_quantization_b200/libb200q_dev.so
for i in 1 2 3; do
  for CFG in "B200Q_NO_CTA2=1" "B200Q_H2_SLOTS=4" "B200Q_H2_SLOTS=8"; do
    env B200Q_LIB=$DEV $CFG timeout 300 python bench.py --steps 50 --warmup 5 --stages-only 2>gpurun_out/cta2_ab.err | python -c "
import json,sys; d=json.loads(sys.stdin.read()); st=d['roofline']['stages']; print('$CFG run $i: ms/step %.4f  conv2_pool %.4f ms' % (d['ms_per_step'], st['conv2_pool']['ms']))"
  done
done
