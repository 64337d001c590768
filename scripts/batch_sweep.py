"""Batch sweep of the static-PTQ forward (BASELINE config 2): device-resident images/s and latency per batch size for
eight plain launches, one CUDA-graph replay, and one CUDA-graph replay with programmatic dependent launch (the
whole-network executor), plus the per-kernel split of the plain forward.
    python scripts/batch_sweep.py [--batches 1,2,...] > profiles/r02_batch_sweep.json
Every row is checked bit for bit against the plain forward."""
import argparse, json, os, sys, warnings
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
warnings.filterwarnings("ignore")
import torch
from convnet_quantization_b200 import synth
from convnet_quantization_b200.engine import _Graph
from convnet_quantization_b200.models.static_ptq_model import StaticPTQModel

ap = argparse.ArgumentParser()
ap.add_argument("--batches", default="1,2,4,8,16,32,64,128,256,512,1024,2048,8192,16384,65536")
a = ap.parse_args()
dev = torch.device("cuda", 0)
model = StaticPTQModel(device=dev)
model.fp32_model.load_state_dict(synth.make_state_dict(0))
engine = model.quantize().engine
rows = []
for b in (int(v) for v in a.batches.split(",")):
    g = torch.Generator(device=dev).manual_seed(b)
    x = synth.normalize(torch.randint(0, 256, (b, 3, 32, 32), dtype=torch.uint8, device=dev, generator=g)).contiguous()
    out = torch.empty((b, 10), dtype=torch.float32, device=dev)
    iters = max(5, min(500, int(2e6 / max(b, 1000))))

    def timed(fn):
        for _ in range(5):
            fn()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        best = None
        for _ in range(3):
            e0.record()
            for _ in range(iters):
                fn()
            e1.record()
            torch.cuda.synchronize()
            t = e0.elapsed_time(e1) / iters
            best = t if best is None else min(best, t)
        return best

    ms_plain = timed(lambda: engine.forward(x, out=out, graph=False))
    ref = out.clone()
    row = {"batch": b, "ms_launches": ms_plain, "images_per_s_launches": b / ms_plain * 1e3}
    stream = torch.cuda.current_stream().cuda_stream
    for key, pdl in (("graph", False), ("graph_pdl", True)):
        gr = _Graph(engine, x, out, pdl)
        out.zero_()
        ms = timed(lambda: gr.launch(stream))
        assert torch.equal(out, ref), f"{key} replay differs from the plain forward at batch {b}"
        row[f"ms_{key}"] = ms
        row[f"images_per_s_{key}"] = b / ms * 1e3
        del gr
    acc = {}
    for i in range(7):
        _, st = engine.forward_profiled(x)
        if i >= 2:
            for k, v in st.items():
                acc[k] = acc.get(k, 0.0) + v / 5
    row["stage_us_plain"] = {k: round(v * 1e3, 2) for k, v in acc.items()}
    rows.append(row)
    print(f"batch {b:6d}: plain {ms_plain * 1e3:9.1f} us  graph {row['ms_graph'] * 1e3:9.1f} us  graph+pdl "
          f"{row['ms_graph_pdl'] * 1e3:9.1f} us  ({b / row['ms_graph_pdl'] * 1e3:12.0f} img/s)  stages "
          + " ".join(f"{v:.0f}" for v in row["stage_us_plain"].values()), file=sys.stderr)
json.dump({"what": "static-PTQ forward, device-resident fp32 input, one B200; best of 3 timed loops (CUDA events)",
           "lib": os.environ.get("B200Q_LIB", "libb200q.so"), "no_halo": os.environ.get("B200Q_NO_HALO", ""),
           "rows": rows}, sys.stdout, indent=1)
