"""Stand-alone memory-bound kernels at the bench batch (16 384 images): CUDA-event time -> achieved algorithmic GB/s
against the measured HBM peak (north_star subsystem 2).  With --once every kernel is launched exactly once after a
warm-up, for `ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum`.
    python scripts/prof_elementwise.py > profiles/r02_elementwise.json"""
import argparse, json, os, sys, warnings
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
warnings.filterwarnings("ignore")
import torch
from convnet_quantization_b200 import ops

ap = argparse.ArgumentParser()
ap.add_argument("--batch", type=int, default=16384)
ap.add_argument("--once", action="store_true")
a = ap.parse_args()
B = a.batch
dev = torch.device("cuda", 0)
peak = 6541.5
try:
    peak = float(json.load(open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "MEASURED_PEAKS.json")))["hbm_gbs"])
except Exception:
    pass
g = torch.Generator(device=dev).manual_seed(0)
x_img = torch.randn(B, 3, 32, 32, device=dev, generator=g)
a64 = torch.randint(0, 256, (B, 32, 32, 64), dtype=torch.uint8, device=dev, generator=g)
a128 = torch.randint(0, 256, (B, 16, 16, 128), dtype=torch.uint8, device=dev, generator=g)
a256 = torch.randint(0, 256, (B, 8, 8, 256), dtype=torch.uint8, device=dev, generator=g)
feat = torch.randn(B, 4096, device=dev, generator=g).abs()
hid = torch.randn(B, 512, device=dev, generator=g).abs()
q10 = torch.randint(0, 256, (B, 16), dtype=torch.uint8, device=dev, generator=g)
lut = torch.arange(255, -1, -1, dtype=torch.uint8)
w1 = ops.DynamicLinearWeights(torch.randint(-127, 128, (512, 4096), dtype=torch.int8), 0.01, torch.zeros(512), dev)
w2 = ops.DynamicLinearWeights(torch.randint(-127, 128, (10, 512), dtype=torch.int8), 0.01, torch.zeros(10), dev)
lo, hi = float(feat.min()), float(feat.max())
MB = 1 << 20
CASES = [  # name, callable, algorithmic bytes (read + written)
    ("quantize_per_tensor fp32 NCHW -> u8 NHWC4 [B,3,32,32]", lambda: ops.quantize_per_tensor(x_img, 0.04, 60, 4), B * 3072 * 4 + B * 4096),
    ("quantize_flat fp32 [B,4096]", lambda: ops.quantize_flat(feat, 0.04, 0), B * 4096 * 5),
    ("dequantize u8 [B,32,32,64] -> fp32", lambda: ops.dequantize(a64, 0.04, 60), B * 65536 * 5),
    ("relu_q u8 [B,32,32,64]", lambda: ops.relu_q(a64, 70), B * 65536 * 2),
    ("lut_u8 u8 [B,32,32,64]", lambda: ops.lut_u8(a64, lut), B * 65536 * 2),
    ("max_pool2d_q u8 [B,32,32,64]", lambda: ops.max_pool2d_q(a64), B * 65536 * 5 // 4),
    ("max_pool2d_q u8 [B,16,16,128]", lambda: ops.max_pool2d_q(a128), B * 32768 * 5 // 4),
    ("max_pool2d_q u8 [B,8,8,256]", lambda: ops.max_pool2d_q(a256), B * 16384 * 5 // 4),
    ("minmax (+ dynamic qparams) fp32 [B,4096]", lambda: ops.minmax(feat), B * 4096 * 4),
    ("aminmax fp32 [B,4096]", lambda: ops.aminmax(feat), B * 4096 * 4),
    ("histc 2048 bins fp32 [B,4096]", lambda: ops.histc(feat, 2048, lo, hi), B * 4096 * 4),
    ("linear_dynamic fc1 fp32 [B,4096] -> [B,512] (minmax + tcgen05 GEMM, x read twice)", lambda: ops.linear_dynamic(feat, w1, True),
     B * 4096 * 4 * 2 + B * 512 * 4 + 4096 * 512),
    ("linear_dynamic fc2 fp32 [B,512] -> [B,10] (minmax + tcgen05 GEMM, x read twice)", lambda: ops.linear_dynamic(hid, w2, False),
     B * 512 * 4 * 2 + B * 10 * 4 + 512 * 10),
]
rows = []
for name, fn, nbytes in CASES:
    for _ in range(1 if a.once else 3):
        fn()
    torch.cuda.synchronize()
    if a.once:
        fn()
        torch.cuda.synchronize()
        continue
    iters = 20
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / iters
    rows.append({"op": name, "ms": ms, "algorithmic_bytes": nbytes, "achieved_gbs": nbytes / ms / 1e6, "frac_of_hbm_peak": nbytes / ms / 1e6 / peak})
    print(f"{name:90s} {ms * 1e3:9.1f} us  {nbytes / ms / 1e6:8.0f} GB/s  {nbytes / ms / 1e6 / peak:5.2f}", file=sys.stderr)
if not a.once:
    json.dump({"what": f"stand-alone memory-bound kernels, batch {B}, CUDA events over 20 launches (includes the output allocation of "
                       "the op wrapper); peak = MEASURED_PEAKS.json hbm_gbs", "hbm_peak_gbs": peak, "rows": rows}, sys.stdout, indent=1)
