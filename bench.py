#!/usr/bin/env python
"""Headline benchmark: images/sec of the static-PTQ int8 SimpleConvNet forward (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--batch B] [--impl reference]

* ours (default): one process per GPU (torchrun for N>1).  A step = one ``b200q_static_forward`` over a batch of
  B synthetic images resident in HBM (fp32 NCHW, the reference's input contract).  ``value`` = all ranks' images /
  max-over-ranks CUDA-event time.  ``e2e`` = the same through the reference-facing model object
  (``StaticPTQModel().quantize()`` -> ``model(x_cpu)``) with pinned HOST input and HOST logits, copies in the timed
  region.  ``roofline`` = the dominant kernel's achieved int8 TOP/s (or GB/s) from per-kernel CUDA events of the same
  forward, against MEASURED_PEAKS.json.  ``cpu_baseline`` = torch/fbgemm CPU oracle on this box's host cores
  (rank 0, N=1, bounded sample).
* ``--impl reference``: the reference's CPU implementation of the path — torch's fbgemm quantized ops applied to
  the reference's fused SimpleConvNet (``oracle/torch_oracle.py``; the reference has no native code of its own) —
  on all host threads, batch 64 (BASELINE config 1), each step a bounded sample.
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import sys
import threading
import time
import warnings

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
warnings.filterwarnings("ignore")

import torch  # noqa: E402

METRIC = "int8 ConvNet images/sec (static-PTQ SimpleConvNet forward)"
UNIT = "images/s"
OPS_PER_IMAGE = 309_733_376  # 154 866 688 MACs x 2

# Algorithmic work per image and stage; "tensor" stages in int8 ops, "hbm" stages in bytes (in+out).
STAGE_WORK = {
    "quant_conv1": ("hbm", 12288 + 65536),
    "conv2": ("tensor", 75_497_472),
    "pool1": ("hbm", 81_920),
    "conv3": ("tensor", 37_748_736),
    "conv4": ("tensor", 75_497_472),
    "pool2": ("hbm", 40_960),
    "conv5": ("tensor", 37_748_736),
    "conv6": ("tensor", 75_497_472),
    "pool3": ("hbm", 20_480),
    "fc1": ("tensor", 4_194_304),
    "fc2_dequant": ("hbm", 512 + 40),
    # fused variants (pool folded into the producing conv): same ops, fewer bytes
    "conv2_pool": ("tensor", 75_497_472),
    "conv1_conv2_pool": ("tensor", 3_538_944 + 75_497_472),  # conv1 + conv2 in one kernel (conv12_fused.cu)
    "conv4_pool": ("tensor", 75_497_472),
    "conv6_pool": ("tensor", 75_497_472),
}


def load_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        p = json.load(open(path))
        return {"hbm_gbs": float(p["hbm_gbs"]), "bf16_burst": float(p["bf16_tflops"]),
                "bf16_sustained": float(p.get("bf16_tflops_sustained", p["bf16_tflops"])), "source": "measured"}
    return {"hbm_gbs": 6650.0, "bf16_burst": 1590.0, "bf16_sustained": 1400.0, "source": "fallback"}


# ----------------------------------------------------------------------------------------------- clocks
class ClockSampler:
    """SM clock + throttle reasons sampled through NVML (what nvidia-smi prints) every 20 ms while running."""
    REASONS = {0x4: "sw_power_cap", 0x8: "hw_slowdown", 0x20: "sw_thermal_slowdown", 0x40: "hw_thermal_slowdown",
               0x80: "hw_power_brake_slowdown", 0x2: "applications_clocks_setting", 0x10: "sync_boost"}

    def __init__(self, device_index: int):
        self.samples, self.mask, self.max_mhz, self.ok = [], 0, None, False
        self._stop = threading.Event()
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            try:
                uuid = str(torch.cuda.get_device_properties(device_index).uuid)
                self.h = pynvml.nvmlDeviceGetHandleByUUID(("GPU-" + uuid).encode())
            except Exception:
                self.h = pynvml.nvmlDeviceGetHandleByIndex(device_index)
            self.max_mhz = int(pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM))
            self.ok = True
        except Exception as e:  # pragma: no cover
            self.err = repr(e)
        self.t = threading.Thread(target=self._run, daemon=True)

    def _run(self):
        nv = self.nv
        while not self._stop.is_set():
            try:
                self.samples.append(int(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM)))
                try:
                    self.mask |= int(nv.nvmlDeviceGetCurrentClocksEventReasons(self.h))
                except Exception:
                    self.mask |= int(nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h))
            except Exception:
                pass
            time.sleep(0.02)

    def __enter__(self):
        if self.ok:
            self.t.start()
        return self

    def __exit__(self, *exc):
        self._stop.set()
        if self.ok:
            self.t.join()

    def summary(self):
        if not self.ok or not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": self.max_mhz, "reasons": ["unavailable"]}
        return {"sm_mhz": int(statistics.median(self.samples)), "sm_max_mhz": self.max_mhz,
                "reasons": [n for bit, n in self.REASONS.items() if self.mask & bit], "samples": len(self.samples)}


# ----------------------------------------------------------------------------------------------- CPU (reference) leg
def build_cpu_oracle():
    from convnet_quantization_b200 import synth
    from convnet_quantization_b200.models.baseline_model import SimpleConvNet
    from oracle import torch_oracle
    net = SimpleConvNet()
    net.load_state_dict(synth.make_state_dict(0))
    return torch_oracle.build_static_oracle(net.eval(), synth.calibration_batches())


def time_cpu_oracle(q, batch: int, batches_per_step: int, steps: int, warmup: int):
    """images/s of the torch/fbgemm CPU static-PTQ forward: ``steps`` timed steps of ``batches_per_step`` x ``batch``."""
    from convnet_quantization_b200 import synth
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    torch.backends.quantized.engine = "fbgemm"
    x = synth.images_f32(batch, seed=11)
    with torch.no_grad():
        for _ in range(max(1, warmup)):
            q(x)
        t0 = time.perf_counter()
        for _ in range(steps):
            for _ in range(batches_per_step):
                q(x)
        dt = time.perf_counter() - t0
    return batch * batches_per_step * steps / dt, dt, cores


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    q = build_cpu_oracle()
    batch, per_step = 64, 16  # BASELINE config 1: batch 64 on CPU; a step = 16 such batches (1024 images)
    ips, dt, cores = time_cpu_oracle(q, batch, per_step, args.steps, args.warmup)
    sample = f"{args.steps} steps x {per_step} batches x {batch} images (torch {torch.__version__} fbgemm, {cores} threads)"
    line = {
        "impl": "reference", "metric": METRIC, "value": ips, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * dt / args.steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "u8", "data": "synthetic",
        "config": {"workload": "static-PTQ int8 SimpleConvNet forward, CIFAR-10 shape fp32 [B,3,32,32] -> logits [B,10]",
                   "batch": batch, "images_per_step": batch * per_step, "device": "host CPU"},
        "cpu_baseline": {"value": ips, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": ips, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)
    return 0


_REAL_STDOUT = None


def emit_line(obj) -> None:
    """Write the one JSON line to the process's original stdout (see run_ours)."""
    data = (json.dumps(obj) + "\n").encode()
    if _REAL_STDOUT is None:
        sys.stdout.write(data.decode())
        sys.stdout.flush()
    else:
        os.write(_REAL_STDOUT, data)


def ncu_traffic(stage: str, batch: int):
    """DRAM bytes (read + written) per launch of `stage`, from the committed `ncu --set full` capture of one forward
    at the same batch (profiles/r01_ncu_net.json, written by scripts/ncu_summary.py); None when no capture matches."""
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "profiles", "r01_ncu_net.json")
    try:
        with open(path) as f:
            d = json.load(f)
        if int(d["batch"]) != int(batch):
            return None
        return float(d["stages"][stage]["dram_bytes"])
    except (OSError, KeyError, ValueError, TypeError):
        return None


# ----------------------------------------------------------------------------------------------- GPU leg
def run_ours(args):
    import torch.distributed as dist
    from convnet_quantization_b200 import _lib, synth
    from convnet_quantization_b200.models.static_ptq_model import StaticPTQModel

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device (the quantized forward has no CPU fallback; use --impl reference)")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    # The contract is ONE JSON line on stdout.  NCCL prints its version banner on the C-level stdout when the first
    # communicator is created, so everything but that line is routed to stderr: fd 1 is pointed at fd 2 for the whole
    # run and the line is written to a private duplicate of the original stdout at the end (emit_line).
    global _REAL_STDOUT
    sys.stdout.flush()
    _REAL_STDOUT = os.dup(1)
    os.dup2(2, 1)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(v: float) -> float:
        if world == 1:
            return v
        t = torch.tensor([v], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    lib = _lib.load()
    model = StaticPTQModel(device=dev)
    model.fp32_model.load_state_dict(synth.make_state_dict(0))
    qmodel = model.quantize()  # fixed synthetic calibration set
    engine = qmodel.engine
    B = args.batch

    # synthetic shard of this rank, generated on the device: uint8 pixels -> /255 -> CIFAR normalise
    g = torch.Generator(device=dev).manual_seed(rank)
    x_u8 = torch.randint(0, 256, (B, 3, 32, 32), dtype=torch.uint8, device=dev, generator=g)
    x = synth.normalize(x_u8).contiguous()
    labels = torch.randint(0, 10, (B,), device=dev, generator=g)
    del x_u8

    for _ in range(max(args.warmup, 3)):
        logits = engine.forward(x)
    barrier()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    n0 = lib.b200q_launch_count()
    with ClockSampler(local) as clocks:
        barrier()
        ev0.record()
        for _ in range(args.steps):
            logits = engine.forward(x)
        ev1.record()
        barrier()
    launches = int(lib.b200q_launch_count() - n0)
    ms = max_over_ranks(ev0.elapsed_time(ev1))
    value = world * B * args.steps / (ms * 1e-3)

    # ---- per-kernel durations of the same forward (CUDA events between kernels, on the launching stream)
    prof_steps = max(3, min(args.steps, 20))
    acc = {}
    for i in range(prof_steps + 2):
        _, stage_ms = engine.forward_profiled(x)
        if i >= 2:
            for k, v in stage_ms.items():
                acc[k] = acc.get(k, 0.0) + v / prof_steps
    peaks = load_peaks()
    int8_peak = 2.0 * peaks["bf16_sustained"]  # tcgen05 kind::i8 issues at twice the bf16 rate (K=32 vs 16 per MMA)
    stages = {}
    for k, v in acc.items():
        bound, work = STAGE_WORK[k]
        if bound == "tensor":
            ach = work * B / (v * 1e-3) / 1e12
            stages[k] = {"ms": v, "bound": "tensor", "achieved": ach, "unit": "TOP/s", "frac": ach / int8_peak}
        else:
            ach = work * B / (v * 1e-3) / 1e9
            stages[k] = {"ms": v, "bound": "hbm", "achieved": ach, "unit": "GB/s", "frac": ach / peaks["hbm_gbs"]}
    total_stage_ms = sum(acc.values())
    dom = max(acc, key=acc.get)
    d = stages[dom]
    roofline = {"kernel": dom, "bound": d["bound"], "achieved": d["achieved"],
                "peak": int8_peak if d["bound"] == "tensor" else peaks["hbm_gbs"], "unit": d["unit"], "frac": d["frac"],
                "traffic": ncu_traffic(dom, B), "share_of_step": acc[dom] / total_stage_ms,
                "peak_source": (f"{peaks['source']}: 2 x bf16_tflops_sustained (int8 = 2x bf16 issue rate)"
                                if d["bound"] == "tensor" else f"{peaks['source']}: hbm_gbs"),
                "net_int8_tops": OPS_PER_IMAGE * B / (total_stage_ms * 1e-3) / 1e12,
                "stages": stages}

    if args.stages_only:  # kernel-timing experiments (scripts/gpu_debug_modes.sh): no e2e leg, no parity assertion
        if rank == 0:
            emit_line({"value": value, "ms_per_step": ms / args.steps, "roofline": roofline})
        return 0

    # ---- end to end through the reference-facing model object: pinned host input -> host logits
    x_host = x.cpu().pin_memory()
    e2e_steps = max(3, min(args.steps, 20))
    for _ in range(2):
        out_host = qmodel(x_host)
    barrier()
    ev0.record()
    for _ in range(e2e_steps):
        out_host = qmodel(x_host)
    ev1.record()
    barrier()
    e2e_ms = max_over_ranks(ev0.elapsed_time(ev1))
    e2e = {"value": world * B * e2e_steps / (e2e_ms * 1e-3), "unit": UNIT, "steps": e2e_steps,
           "h2d_bytes_per_step": x_host.numel() * x_host.element_size(),
           "d2h_bytes_per_step": out_host.numel() * out_host.element_size(),
           "api": "StaticPTQModel().quantize() -> model(x_cpu_pinned) -> cpu logits"}
    assert torch.equal(out_host, logits.cpu()), "e2e logits differ from the device-resident run"

    # ---- the same call on the uint8 data path (extra key, NOT the contract's e2e): raw uint8 NHWC
    #      pixels from pinned host memory -> host logits; a quarter of the bytes cross PCIe.  Its inputs are other images
    #      (the fp32 batch above was normalised on the device), so only determinism is checked here; bit-identity with
    #      the fp32 route is a parity test (tests/test_gpu_net.py).
    pix_host = torch.randint(0, 256, (B, 32, 32, 3), dtype=torch.uint8,
                             generator=torch.Generator().manual_seed(1000 + rank)).pin_memory()
    for _ in range(2):
        out_u8 = qmodel.forward_uint8(pix_host)
    barrier()
    ev0.record()
    for _ in range(e2e_steps):
        out_u8_2 = qmodel.forward_uint8(pix_host)
    ev1.record()
    barrier()
    u8_ms = max_over_ranks(ev0.elapsed_time(ev1))
    assert torch.equal(out_u8, out_u8_2)
    e2e_u8 = {"value": world * B * e2e_steps / (u8_ms * 1e-3), "unit": UNIT, "steps": e2e_steps,
              "h2d_bytes_per_step": pix_host.numel(), "d2h_bytes_per_step": out_u8.numel() * out_u8.element_size(),
              "api": "model.forward_uint8(pixels_cpu_pinned uint8 NHWC) -> cpu logits"}

    # ---- correct-count: the only collective, off the hot path (one NCCL all-reduce of 3 int64)
    from convnet_quantization_b200 import sharding
    counts = [int(v) for v in sharding.allreduce_counts(sharding.topk_counts(logits, labels)).tolist()]

    cpu_baseline = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        q = build_cpu_oracle()
        # bounded sample: ~10-20 s of CPU work at batch 64 (BASELINE config 1)
        ips0, _, cores = time_cpu_oracle(q, 64, 8, 2, 1)
        steps_cpu = max(2, int(12.0 * ips0 / (64 * 16)))
        ips, dt, cores = time_cpu_oracle(q, 64, 16, steps_cpu, 1)
        cpu_baseline = {"value": ips, "unit": UNIT, "cores": cores, "kind": "port",
                        "sample": f"{steps_cpu * 16} batches x 64 images in {dt:.1f} s, torch {torch.__version__} "
                                  f"fbgemm static-PTQ oracle, {cores} threads"}

    if rank == 0:
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "u8", "data": "synthetic",
            "config": {"workload": "static-PTQ int8 SimpleConvNet forward, CIFAR-10 shape fp32 [B,3,32,32] -> logits [B,10]",
                       "batch_per_gpu": B, "global_batch": B * world, "parallelism": f"batch-sharded x{world}",
                       "l2": f"input {x.numel() * 4 / 2**20:.0f} MiB + activations > 126 MiB L2 (no flush needed)"},
            "clocks": clocks.summary(), "e2e": e2e, "e2e_u8": e2e_u8, "gpu_launches": launches, "roofline": roofline,
            "cpu_baseline": cpu_baseline, "eval": {"top1": counts[0], "top5": counts[1], "total": counts[2]},
        }
        emit_line(line)
    if world > 1:
        dist.destroy_process_group()
    return 0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--batch", type=int, default=16384, help="images per GPU per step")
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--stages-only", action="store_true", help="print only the per-kernel table (timing experiments)")
    args = ap.parse_args()
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if args.impl == "ours" and world != args.gpus and args.gpus > 1:
        raise SystemExit(f"--gpus {args.gpus} needs torchrun with {args.gpus} ranks (WORLD_SIZE={world})")
    return run_reference(args) if args.impl == "reference" else run_ours(args)


if __name__ == "__main__":
    sys.exit(main())
