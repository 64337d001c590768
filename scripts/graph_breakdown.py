"""Where the small-batch forward spends its time: CUDA graphs of the first k kernels of the forward (development library,
flag bits 8..11 of b200q_graph_create), with programmatic dependent launch; T(k) - T(k-1) is what kernel k adds to the
dependent chain.  B200Q_LIB must point at libb200q_dev.so.
    B200Q_LIB=$PWD/convnet_quantization_b200/libb200q_dev.so python scripts/graph_breakdown.py > profiles/r02_graph_breakdown.json"""
import ctypes as C, json, os, sys, warnings
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
warnings.filterwarnings("ignore")
import torch
from convnet_quantization_b200 import _lib, synth
from convnet_quantization_b200.models.static_ptq_model import StaticPTQModel

dev = torch.device("cuda", 0)
model = StaticPTQModel(device=dev)
model.fp32_model.load_state_dict(synth.make_state_dict(0))
engine = model.quantize().engine
lib = engine.lib
names = ["quant_conv1", "conv2_pool", "conv3", "conv4_pool", "conv5", "conv6_pool", "head (fc1+fc2+dequant)"]
side = torch.cuda.Stream(dev)
rows = []
for b in tuple(int(v) for v in os.environ.get("BREAKDOWN_BATCHES", "1,8,32").split(",")):
    x = synth.images_f32(b, seed=b).cuda()
    out = torch.empty((b, 10), device=dev)
    ws = torch.empty(int(lib.b200q_static_workspace_bytes(b)), dtype=torch.uint8, device=dev)
    t = []
    for pdl in (1, 0):
        t.append([])
        for k in range(1, 8):
            g = C.c_void_p()
            torch.cuda.synchronize()
            _lib.check(lib.b200q_graph_create(engine.packed.ptr(), x.data_ptr(), out.data_ptr(), b, ws.data_ptr(), ws.numel(),
                                              pdl | (k << 8), side.cuda_stream, C.byref(g)), "graph_create")
            torch.cuda.synchronize()
            s = torch.cuda.current_stream().cuda_stream
            best = None
            for _ in range(3):
                for _ in range(20):
                    lib.b200q_graph_launch(g, s)
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                for _ in range(500):
                    lib.b200q_graph_launch(g, s)
                e1.record()
                torch.cuda.synchronize()
                us = e0.elapsed_time(e1) / 500 * 1e3
                best = us if best is None else min(best, us)
            lib.b200q_graph_destroy(g)
            t[-1].append(best)
    row = {"batch": b, "kernels": names, "cumulative_us_pdl": t[0], "cumulative_us_no_pdl": t[1],
           "marginal_us_pdl": [t[0][0]] + [t[0][i] - t[0][i - 1] for i in range(1, 7)],
           "marginal_us_no_pdl": [t[1][0]] + [t[1][i] - t[1][i - 1] for i in range(1, 7)]}
    rows.append(row)
    print(f"batch {b}: pdl   " + " ".join(f"{v:5.1f}" for v in row["marginal_us_pdl"]) + f"  = {t[0][-1]:.1f}", file=sys.stderr)
    print(f"batch {b}: nopdl " + " ".join(f"{v:5.1f}" for v in row["marginal_us_no_pdl"]) + f"  = {t[1][-1]:.1f}", file=sys.stderr)
json.dump({"what": "CUDA graph of the first k kernels of the static-PTQ forward, replayed back to back (best of 3 x 500); marginal = "
                   "T(k) - T(k-1); the first entry includes the graph launch itself", "rows": rows}, sys.stdout, indent=1)
