#!/bin/bash
# Timing experiments: halo kernels with one role disabled (results are wrong in those modes; only stage times matter).
mkdir -p gpurun_out
for D in 0 2 4 8 10 12; do
  B200Q_HALO_DEBUG=$D timeout 600 python bench.py --steps 10 --warmup 3 --stages-only > gpurun_out/bench_dbg$D.json 2> gpurun_out/bench_dbg$D.err || tail -3 gpurun_out/bench_dbg$D.err
  python - <<PY
import json
d=json.loads(open("gpurun_out/bench_dbg$D.json").read())
st=d["roofline"]["stages"]
print("debug=$D", " ".join("%s %.3f" % (k, v["ms"]) for k,v in st.items()))
PY
done
