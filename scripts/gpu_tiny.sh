#!/bin/bash
# CUDA-core tiny-batch kernels (conv_small.cu, conv1_kernel<.,1>): parity, then the small-batch sweep with them off / on
# (development library), then the per-kernel breakdown of the captured chain.
mkdir -p gpurun_out
DEV=$PWD/convnet_quantization_b200/libb200q_dev.so
timeout 600 python -m pytest tests/test_gpu_conv.py tests/test_gpu_net.py tests/test_gpu_guard_bands.py -m gpu -q -x -p no:cacheprovider > gpurun_out/pytest_tiny.log 2>&1; echo "pytest exit=$? :: $(tail -n 1 gpurun_out/pytest_tiny.log)"
grep -E "^(FAILED|ERROR)|Error|mismatch" gpurun_out/pytest_tiny.log | head -20
for M in 0 2 0 2; do
  echo "== B200Q_TINY_MAX_B=$M"
  B200Q_LIB=$DEV B200Q_TINY_MAX_B=$M timeout 300 python scripts/batch_sweep.py --batches 1,2,3 > gpurun_out/batch_sweep_tiny$M.json 2> gpurun_out/batch_sweep_tiny$M.log; echo "exit=$?"; grep -v Warn gpurun_out/batch_sweep_tiny$M.log
done
B200Q_LIB=$DEV timeout 600 python scripts/graph_breakdown.py > gpurun_out/graph_breakdown.json 2> gpurun_out/graph_breakdown.log; echo "breakdown exit=$?"; cat gpurun_out/graph_breakdown.log
