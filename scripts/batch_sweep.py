"""Batch sweep of the static-PTQ forward (BASELINE config 2): device-resident images/s and latency per batch size,
plain launches vs one CUDA-graph replay.  python scripts/batch_sweep.py > profiles/r01_batch_sweep.json"""
import json, os, sys, warnings
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
warnings.filterwarnings("ignore")
import torch
from convnet_quantization_b200 import synth
from convnet_quantization_b200.models.static_ptq_model import StaticPTQModel

dev = torch.device("cuda", 0)
model = StaticPTQModel(device=dev)
model.fp32_model.load_state_dict(synth.make_state_dict(0))
engine = model.quantize().engine
rows = []
for b in (1, 8, 32, 128, 512, 2048, 8192, 16384, 65536):
    g = torch.Generator(device=dev).manual_seed(b)
    x = synth.normalize(torch.randint(0, 256, (b, 3, 32, 32), dtype=torch.uint8, device=dev, generator=g)).contiguous()
    out = torch.empty((b, 10), dtype=torch.float32, device=dev)
    iters = max(5, min(200, int(2e6 / max(b, 1000))))
    def timed(fn):
        for _ in range(3):
            fn()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(iters):
            fn()
        e1.record()
        torch.cuda.synchronize()
        return e0.elapsed_time(e1) / iters
    ms_plain = timed(lambda: engine.forward(x, out=out))
    ref = out.clone()
    graph = torch.cuda.CUDAGraph()
    s = torch.cuda.Stream(dev)
    s.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(s):
        engine.forward(x, out=out)  # workspace for this stream
        torch.cuda.synchronize()
        with torch.cuda.graph(graph, stream=s):
            engine.forward(x, out=out)
    torch.cuda.current_stream().wait_stream(s)
    out.zero_()
    ms_graph = timed(graph.replay)
    assert torch.equal(out, ref), "graph replay differs"
    rows.append({"batch": b, "ms_launches": ms_plain, "ms_graph": ms_graph, "images_per_s_launches": b / ms_plain * 1e3,
                 "images_per_s_graph": b / ms_graph * 1e3})
    print(f"batch {b:6d}: {ms_plain:8.4f} ms plain ({b / ms_plain * 1e3:12.0f} img/s)   {ms_graph:8.4f} ms graph "
          f"({b / ms_graph * 1e3:12.0f} img/s)", file=sys.stderr)
json.dump({"what": "static-PTQ forward, device-resident fp32 input, one B200", "rows": rows}, sys.stdout, indent=1)
