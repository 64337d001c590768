#!/bin/bash
# conv2 on CTA pairs (tcgen05.mma.cta_group::2): parity first, then interleaved A-B against the single-CTA kernel.
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_conv.py -m gpu -q -p no:cacheprovider -x -k "conv2" > gpurun_out/pytest_cta2_conv.log 2>&1; echo "pytest conv2 exit=$? :: $(tail -n 1 gpurun_out/pytest_cta2_conv.log)"
grep -E "^(FAILED|ERROR)|Error|error|trap|timeout" gpurun_out/pytest_cta2_conv.log | head -20
timeout 900 python -m pytest tests/test_gpu_net.py -m gpu -q -p no:cacheprovider -x > gpurun_out/pytest_cta2_net.log 2>&1; echo "pytest net exit=$? :: $(tail -n 1 gpurun_out/pytest_cta2_net.log)"
grep -E "^(FAILED|ERROR)" gpurun_out/pytest_cta2_net.log | head
DEV=$PWD/convnet_quantization_b200/libb200q_dev.so
for i in 1 2; do
  for V in 1 0; do
    B200Q_LIB=$DEV B200Q_NO_CTA2=$V timeout 300 python bench.py --steps 50 --warmup 5 --stages-only 2>gpurun_out/cta2_ab.err | python -c "
import json,sys; d=json.loads(sys.stdin.read()); st=d['roofline']['stages']; print('no_cta2=$V run $i: ms/step %.4f  conv2_pool %.4f ms  conv1 %.4f' % (d['ms_per_step'], st['conv2_pool']['ms'], st['quant_conv1']['ms']))"
  done
done
tail -3 gpurun_out/cta2_ab.err
