"""Where the time of one host-input call goes: per-chunk CUDA-event timeline of the pipelined `model(x_cpu)` /
`model.forward_uint8(pixels_cpu)` call, and an A/B of pipeline shapes (alternating streams vs a dedicated copy stream).

    python scripts/e2e_timeline.py [--batch 16384] > gpurun_out/e2e_timeline.json
"""
import argparse, json, os, statistics, sys, time, warnings
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
warnings.filterwarnings("ignore")
import torch
from convnet_quantization_b200 import synth
from convnet_quantization_b200.models.static_ptq_model import StaticPTQModel

ap = argparse.ArgumentParser()
ap.add_argument("--batch", type=int, default=16384)
ap.add_argument("--reps", type=int, default=15)
a = ap.parse_args()
dev = torch.device("cuda", 0)
model = StaticPTQModel(device=dev)
model.fp32_model.load_state_dict(synth.make_state_dict(0))
q = model.quantize()
eng = q.engine
B = a.batch
g = torch.Generator().manual_seed(0)
pix = torch.randint(0, 256, (B, 32, 32, 3), dtype=torch.uint8, generator=g).pin_memory()
x = synth.normalize(pix.permute(0, 3, 1, 2).contiguous().to(dev)).contiguous().cpu().pin_memory()


def wall(fn, reps=a.reps):
    for _ in range(3):
        fn()
    ts = []
    for _ in range(reps):
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        fn()
        torch.cuda.synchronize()
        ts.append((time.perf_counter() - t0) * 1e3)
    return {"median_ms": statistics.median(ts), "min_ms": min(ts), "images_per_s_median": B / statistics.median(ts) * 1e3}


def plain_copy_ms(t):
    d = torch.empty_like(t, device=dev)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    for _ in range(3):
        d.copy_(t, non_blocking=True)
    torch.cuda.synchronize()
    e0.record()
    for _ in range(5):
        d.copy_(t, non_blocking=True)
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / 5


def timeline(u8: bool):
    """The module's own loop (models/_gpu_modules.py::_forward_host), with events after every copy / forward / readback."""
    src = pix if u8 else x
    q.forward_uint8(src) if u8 else q(src)  # buffers, graphs
    pipe = q._pipeline()
    chunk = q.U8_MAX_CHUNK if u8 else q.HOST_CHUNK
    plan = list(q._chunks_ramp(B, q.U8_FIRST_CHUNK, chunk) if u8 else q._chunks(B, chunk))
    out = pipe["out"][:B]
    ev = lambda: torch.cuda.Event(enable_timing=True)
    rows = []
    torch.cuda.synchronize()
    cur = torch.cuda.current_stream(dev)
    t0 = ev(); t0.record(cur)
    for s in pipe["streams"]:
        s.wait_stream(cur)
    c0 = time.perf_counter()
    for i, (lo, n) in enumerate(plan):
        k = i & 1
        st = pipe["streams"][k]
        with torch.cuda.stream(st):
            xin, yout = (pipe["xu8"] if u8 else pipe["x"])[k][:n], (pipe["yu8"] if u8 else pipe["y"])[k][:n]
            e_a = ev(); e_a.record(st)
            xin.copy_(src[lo:lo + n], non_blocking=True)
            e_b = ev(); e_b.record(st)
            eng.forward_u8(xin, out=yout) if u8 else eng.forward(xin, out=yout)
            e_c = ev(); e_c.record(st)
            out[lo:lo + n].copy_(yout, non_blocking=True)
            e_d = ev(); e_d.record(st)
        rows.append((n, e_a, e_b, e_c, e_d, (time.perf_counter() - c0) * 1e3))
    for s in pipe["streams"]:
        s.synchronize()
    return [{"images": n, "h2d_start_ms": t0.elapsed_time(e_a), "h2d_end_ms": t0.elapsed_time(e_b),
             "kernels_end_ms": t0.elapsed_time(e_c), "d2h_end_ms": t0.elapsed_time(e_d), "cpu_enqueued_ms": c}
            for n, e_a, e_b, e_c, e_d, c in rows]


def halving_plan(b, chunk, smallest):
    """Full chunks, then the last `chunk` images in halves (1/2, 1/4, ... down to `smallest`): the exposed tail is the
    kernels of a `smallest`-image chunk."""
    plan, lo = [], 0
    while b - lo > chunk:
        plan.append((lo, chunk)); lo += chunk
    rest = b - lo
    while rest >= 2 * smallest:
        n = rest // 2
        plan.append((lo, n)); lo += n; rest -= n
    plan.append((lo, rest))
    return plan


class AltStreamsPipe:
    """The module's shape (two alternating streams) with another chunk plan."""

    def __init__(self, chunk, smallest):
        self.chunk, self.smallest = chunk, smallest
        self.xb = [torch.empty((chunk, 3, 32, 32), dtype=torch.float32, device=dev) for _ in range(2)]
        self.yb = [torch.empty((chunk, 10), dtype=torch.float32, device=dev) for _ in range(2)]
        self.streams = [torch.cuda.Stream(dev), torch.cuda.Stream(dev)]
        self.out = torch.empty((B, 10), dtype=torch.float32).pin_memory()

    def __call__(self, src):
        b = src.shape[0]
        cur = torch.cuda.current_stream(dev)
        for s in self.streams:
            s.wait_stream(cur)
        for i, (lo, n) in enumerate(halving_plan(b, self.chunk, self.smallest)):
            k = i & 1
            with torch.cuda.stream(self.streams[k]):
                xin, yout = self.xb[k][:n], self.yb[k][:n]
                xin.copy_(src[lo:lo + n], non_blocking=True)
                eng.forward(xin, out=yout)
                self.out[lo:lo + n].copy_(yout, non_blocking=True)
        for s in self.streams:
            s.synchronize()
        return self.out[:b].clone()


class CopyStreamPipe:
    """Alternative shape: every host->device copy on ONE dedicated stream (back to back, no cross-stream hand-offs),
    kernels + readback on a second stream, `ring` staging buffers guarded by events."""

    def __init__(self, u8: bool, chunk: int, ring: int = 3, first: int | None = None, halving: int = 0):
        self.u8, self.chunk, self.ring, self.first, self.halving = u8, chunk, ring, first, halving
        shape = (chunk, 32, 32, 3) if u8 else (chunk, 3, 32, 32)
        self.xb = [torch.empty(shape, dtype=torch.uint8 if u8 else torch.float32, device=dev) for _ in range(ring)]
        self.yb = [torch.empty((chunk, 10), dtype=torch.float32, device=dev) for _ in range(ring)]
        self.copy_s, self.comp_s = torch.cuda.Stream(dev), torch.cuda.Stream(dev)
        self.filled = [torch.cuda.Event() for _ in range(ring)]
        self.freed = [torch.cuda.Event() for _ in range(ring)]
        self.out = torch.empty((B, 10), dtype=torch.float32).pin_memory()

    def plan(self, b):
        if self.first:
            return list(q._chunks_ramp(b, self.first, self.chunk))
        if self.halving:
            return halving_plan(b, self.chunk, self.halving)
        return list(q._chunks(b, self.chunk))

    def __call__(self, src):
        b = src.shape[0]
        cur = torch.cuda.current_stream(dev)
        self.copy_s.wait_stream(cur)
        self.comp_s.wait_stream(cur)
        for i, (lo, n) in enumerate(self.plan(b)):
            k = i % self.ring
            xin, yout = self.xb[k][:n], self.yb[k][:n]
            with torch.cuda.stream(self.copy_s):
                if i >= self.ring:
                    self.copy_s.wait_event(self.freed[k])
                xin.copy_(src[lo:lo + n], non_blocking=True)
                self.filled[k].record(self.copy_s)
            with torch.cuda.stream(self.comp_s):
                self.comp_s.wait_event(self.filled[k])
                eng.forward_u8(xin, out=yout) if self.u8 else eng.forward(xin, out=yout)
                self.freed[k].record(self.comp_s)
                self.out[lo:lo + n].copy_(yout, non_blocking=True)
        self.comp_s.synchronize()
        return self.out[:b].clone()


res = {"batch": B, "h2d_alone_ms": {"fp32": plain_copy_ms(x), "uint8": plain_copy_ms(pix)}}
ref32, ref8 = q(x), q.forward_uint8(pix)
res["module_fp32"] = wall(lambda: q(x))
res["module_uint8"] = wall(lambda: q.forward_uint8(pix))
res["timeline_fp32"] = timeline(False)
res["timeline_uint8"] = timeline(True)
alts = {}
for name, u8, chunk, ring, first in [("fp32_copystream_2048_r3", False, 2048, 3, None), ("fp32_copystream_4096_r3", False, 4096, 3, None),
                                     ("fp32_copystream_1024_r4", False, 1024, 4, None),
                                     ("uint8_copystream_ramp1024_8192_r3", True, 8192, 3, 1024),
                                     ("uint8_copystream_ramp512_4096_r3", True, 4096, 3, 512),
                                     ("uint8_copystream_ramp2048_8192_r3", True, 8192, 3, 2048)]:
    p = CopyStreamPipe(u8, chunk, ring, first)
    src = pix if u8 else x
    assert torch.equal(p(src), ref8 if u8 else ref32), name
    alts[name] = wall(lambda: p(src))
    del p
for name, mk in [("fp32_alt_2048_halve256", lambda: AltStreamsPipe(2048, 256)), ("fp32_alt_2048_halve128", lambda: AltStreamsPipe(2048, 128)),
                 ("fp32_alt_2048_halve512", lambda: AltStreamsPipe(2048, 512)),
                 ("fp32_copystream_2048_r4_halve256", lambda: CopyStreamPipe(False, 2048, 4, None, 256)),
                 ("fp32_copystream_2048_r4_halve128", lambda: CopyStreamPipe(False, 2048, 4, None, 128)),
                 ("fp32_copystream_1024_r6_halve128", lambda: CopyStreamPipe(False, 1024, 6, None, 128)),
                 ("fp32_copystream_1024_r6_halve256", lambda: CopyStreamPipe(False, 1024, 6, None, 256))]:
    p = mk()
    assert torch.equal(p(x), ref32), name
    alts[name] = wall(lambda: p(x))
    del p
res["alternatives"] = alts
# the module's uint8 route under other ramps
first0, cap0 = q.U8_FIRST_CHUNK, q.U8_MAX_CHUNK
for first, cap in ((512, 4096), (1024, 2048), (1024, 4096), (1024, 8192), (2048, 2048), (2048, 4096), (4096, 4096), (8192, 8192)):
    q.U8_FIRST_CHUNK, q.U8_MAX_CHUNK, q._pipe = first, cap, None
    assert torch.equal(q.forward_uint8(pix), ref8)
    res[f"module_uint8_first{first}_cap{cap}"] = wall(lambda: q.forward_uint8(pix))
q.U8_FIRST_CHUNK, q.U8_MAX_CHUNK, q._pipe = first0, cap0, None
print(json.dumps(res))
for k, v in res.items():
    if isinstance(v, dict) and "median_ms" in v:
        print(f"{k:40s} {v['median_ms']:.3f} ms  {v['images_per_s_median'] / 1e6:.3f} M img/s", file=sys.stderr)
for k, v in alts.items():
    print(f"{k:40s} {v['median_ms']:.3f} ms  {v['images_per_s_median'] / 1e6:.3f} M img/s", file=sys.stderr)
print("h2d alone", res["h2d_alone_ms"], file=sys.stderr)
for name in ("timeline_fp32", "timeline_uint8"):
    print(name, file=sys.stderr)
    for r in res[name]:
        print("  " + "  ".join(f"{k}={v:.3f}" if isinstance(v, float) else f"{k}={v}" for k, v in r.items()), file=sys.stderr)
