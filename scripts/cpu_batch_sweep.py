"""SURVEY 8(d) "CPU baseline timed beside it": the torch/fbgemm static-PTQ oracle on the box's host cores at batch
1 / 32 / 64 / 256 (all threads; the core count and CPU model are recorded), next to the B200 numbers of
profiles/r02_batch_sweep.json.  Test infrastructure: this is one of the places that may execute oracle/.

    python scripts/cpu_batch_sweep.py > gpurun_out/cpu_batch_sweep.json
"""
import json, os, platform, sys, time, warnings
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
warnings.filterwarnings("ignore")
import torch
import bench
from convnet_quantization_b200 import synth

q = bench.build_cpu_oracle()
cores = os.cpu_count() or 1
torch.set_num_threads(cores)
torch.backends.quantized.engine = "fbgemm"
model = ""
try:
    with open("/proc/cpuinfo") as f:
        model = next((l.split(":", 1)[1].strip() for l in f if l.startswith("model name")), "")
except OSError:
    pass
rows = []
with torch.no_grad():
    for b in (1, 32, 64, 256):
        x = synth.images_f32(b, seed=11)
        for _ in range(5):
            q(x)
        best = None
        for _ in range(3):  # best of three ~2 s loops
            n, t0 = 0, time.perf_counter()
            while time.perf_counter() - t0 < 2.0:
                q(x)
                n += 1
            dt = time.perf_counter() - t0
            ips = n * b / dt
            best = ips if best is None or ips > best else best
        rows.append({"batch": b, "images_per_s": best, "ms_per_batch": 1e3 * b / best})
        print(f"batch {b:4d}: {best:10.0f} img/s  ({1e3 * b / best:.3f} ms per batch)", file=sys.stderr)
print(json.dumps({"what": "torch/fbgemm static-PTQ CPU oracle (oracle/torch_oracle.py), all host threads, best of 3 x 2 s",
                  "torch": torch.__version__, "threads": cores, "cpu": model or platform.processor(), "rows": rows}))
