#!/bin/bash
# Kernel-level GPU parity run: every family in its own process so that a trap in one cannot mask the others.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,driver_version,clocks.max.sm,memory.total --format=csv > gpurun_out/gpu.txt 2>&1
run() { # name, timeout, args...
  local name=$1; shift; local to=$1; shift
  timeout $to python -m pytest -q -m gpu "$@" > gpurun_out/$name.log 2>&1
  echo "$name exit=$? :: $(tail -n 1 gpurun_out/$name.log)"
}
run elem 600 tests/test_gpu_elementwise.py
run simt 900 tests/test_gpu_conv.py -k "conv1 or simt or fc2 or dynamic"
for L in conv2 conv3 conv4 conv5 conv6; do
  run tc_$L 300 tests/test_gpu_conv.py -k "test_conv_tc and $L"
done
run tc_fc1 300 tests/test_gpu_conv.py -k "linear_tc"
for L in conv2 conv3 conv4; do
  B200Q_TC_STREAM_WEIGHTS=1 run tcs_$L 300 tests/test_gpu_conv.py -k "test_conv_tc and $L"
done
