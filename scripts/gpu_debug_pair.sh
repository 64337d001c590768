#!/bin/bash
# Timing experiments: conv_pair kernels with one role disabled (results are wrong in those modes).
mkdir -p gpurun_out
for D in 0 2 4 8 32 36 44 46; do
  B200Q_PAIR_DEBUG=$D timeout 600 python bench.py --steps 10 --warmup 3 --stages-only > gpurun_out/bench_pdbg$D.json 2> gpurun_out/bench_pdbg$D.err || tail -3 gpurun_out/bench_pdbg$D.err
  python - <<PY
import json
d=json.loads(open("gpurun_out/bench_pdbg$D.json").read())
st=d["roofline"]["stages"]
print("pair debug=$D", " ".join("%s %.3f" % (k, v["ms"]) for k,v in st.items() if k in ("conv5","conv6_pool")))
PY
done
