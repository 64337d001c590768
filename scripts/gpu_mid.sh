#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_net.py tests/test_gpu_conv.py -m gpu -q -p no:cacheprovider > gpurun_out/pytest_mid.log 2>&1; echo "pytest exit=$? :: $(tail -n 1 gpurun_out/pytest_mid.log)"
grep -E "^(FAILED|ERROR)" gpurun_out/pytest_mid.log | head
timeout 600 python scripts/batch_sweep.py --batches 64,96,128,192,256,384,443,444,512,1024 > gpurun_out/batch_sweep_mid.json 2> gpurun_out/batch_sweep_mid.log; echo "sweep exit=$?"; cat gpurun_out/batch_sweep_mid.log
