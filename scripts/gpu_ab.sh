#!/bin/bash
# A-B timing on one box: baseline build (convnet_quantization_b200/build/libb200q_base.so) vs the current one,
# interleaved.  Extra environment (e.g. B200Q_HALO_EW=8) applies to both.
mkdir -p gpurun_out
for rep in 1 2; do
  for V in base new; do
    if [ $V = base ]; then export B200Q_LIB=$PWD/convnet_quantization_b200/build/libb200q_base.so; else unset B200Q_LIB; fi
    timeout 300 python bench.py --steps 20 --warmup 3 --stages-only > gpurun_out/ab_$V.json 2> gpurun_out/ab_$V.err || tail -3 gpurun_out/ab_$V.err
    python - <<PY
import json
d=json.loads(open("gpurun_out/ab_$V.json").read())
print("%-4s ms/step %.3f " % ("$V", d["ms_per_step"]), " ".join("%s %.3f" % (k[:8], v["ms"]) for k,v in d["roofline"]["stages"].items()))
PY
  done
done
