#!/bin/bash
# A-B: programmatic dependent launch in the EAGER forward (development library switch), interleaved runs on one box.
mkdir -p gpurun_out
DEV=$PWD/convnet_quantization_b200/libb200q_dev.so
for i in 1 2 3; do
  for V in 0 1; do
    B200Q_LIB=$DEV B200Q_EAGER_PDL=$V timeout 300 python bench.py --steps 50 --warmup 5 --no-cpu-baseline --sustain-s 1 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('eager_pdl=$V run $i: ms/step %.4f  value %.0f  sustained %.0f  parity %s' % (d['ms_per_step'], d['value'], d['sustained']['value'], d['parity']['bit_exact']))"
  done
done
