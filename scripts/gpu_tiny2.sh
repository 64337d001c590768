#!/bin/bash
# Cut-off of the 16-CTA-per-image CUDA-core conv1 shape (development library): sweep batches at several cut-offs.
mkdir -p gpurun_out
DEV=$PWD/convnet_quantization_b200/libb200q_dev.so
for M in 2 8 32 64 2 32; do
  echo "== B200Q_CONV1_TINY_MAX_B=$M"
  B200Q_LIB=$DEV B200Q_CONV1_TINY_MAX_B=$M timeout 300 python scripts/batch_sweep.py --batches 3,4,8,16,32,64 > gpurun_out/batch_sweep_c1cut$M.json 2> gpurun_out/batch_sweep_c1cut$M.log; echo "exit=$?"; grep -v Warn gpurun_out/batch_sweep_c1cut$M.log
done
