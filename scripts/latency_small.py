"""Wall-clock latency of model(x) at small batch (the reference's InferenceBenchmark protocol: batch 32, per-call
synchronised timing), device-resident input: plain engine call vs a captured CUDA graph with static buffers."""
import os, sys, time, warnings
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
warnings.filterwarnings("ignore")
import torch
from convnet_quantization_b200 import synth
from convnet_quantization_b200.models.static_ptq_model import StaticPTQModel
dev = torch.device("cuda", 0)
m = StaticPTQModel(device=dev); m.fp32_model.load_state_dict(synth.make_state_dict(0)); q = m.quantize()
for b in (1, 32, 256, 1024):
    x = synth.images_f32(b, seed=b).to(dev)
    def wall(fn, n=300):
        for _ in range(20): fn()
        torch.cuda.synchronize(); t0 = time.perf_counter()
        for _ in range(n):
            y = fn(); torch.cuda.synchronize()
        return (time.perf_counter() - t0) / n * 1e6
    us_model = wall(lambda: q(x))
    us_engine = wall(lambda: q.engine.forward(x))
    print(f"batch {b:5d}: model(x) {us_model:7.1f} us   engine.forward {us_engine:7.1f} us (wall, synchronised per call)")
