#!/bin/bash
# Cut-off of the CUDA-core conv2-conv6 kernel with the conv1 shape at its product cut-off (development library).
mkdir -p gpurun_out
DEV=$PWD/convnet_quantization_b200/libb200q_dev.so
for M in 1 2 3 1 2 3; do
  echo "== B200Q_TINY_MAX_B=$M"
  B200Q_LIB=$DEV B200Q_TINY_MAX_B=$M timeout 300 python scripts/batch_sweep.py --batches 1,2,3,4 > gpurun_out/batch_sweep_cutb$M.json 2> gpurun_out/batch_sweep_cutb$M.log; echo "exit=$?"; grep -v Warn gpurun_out/batch_sweep_cutb$M.log
done
