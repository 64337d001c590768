"""``nn.Module`` shells around the CUDA engines, shaped so the reference's drivers can use them unchanged
(``utils/inference_benchmark.py:19-28`` calls ``eval()/to(device)/model(data)``; ``utils/model_evaluator.py:19-33``
calls ``eval()/cpu()/model(cpu_images)`` and ``outputs.topk`` on the result):

* ``forward`` accepts fp32 NCHW ``[B,3,32,32]`` on CPU **or** CUDA and returns fp32 logits ``[B,10]`` on the input's
  device (H2D / D2H done here when the caller hands CPU tensors);
* ``.cpu()`` / ``.to('cpu')`` never tear the GPU engine down (the arithmetic has no CPU fallback);
* ``.to('cuda:N')`` re-homes the engine; everything mutates in place because the driver discards ``.to()``'s result.
"""
from __future__ import annotations

import os

import torch
import torch.nn as nn
import torch.nn.functional as F

from .. import _lib, ops
from ..engine import StaticEngine


def _default_device() -> torch.device:
    if not torch.cuda.is_available():
        raise _lib.B200QError("no CUDA device: the quantized forward has no CPU fallback")
    return torch.device("cuda", torch.cuda.current_device())


class _GpuResident(nn.Module):
    """Common device handling: the engine lives on ``self.engine_device`` regardless of ``.cpu()``."""

    quantized = True  # sniffed by utils/model_evaluator.py:25,66

    def __init__(self, device=None):
        super().__init__()
        self.engine_device = torch.device(device) if device is not None else _default_device()
        if self.engine_device.type != "cuda":
            raise _lib.B200QError("engine device must be CUDA")
        if self.engine_device.index is None:
            self.engine_device = torch.device("cuda", torch.cuda.current_device())

    def _rehome(self, device: torch.device):  # overridden
        raise NotImplementedError

    def to(self, *args, **kwargs):
        device = kwargs.get("device", args[0] if args else None)
        if isinstance(device, str):
            device = torch.device(device)
        if isinstance(device, torch.device) and device.type == "cuda":
            if device.index is None:
                device = torch.device("cuda", torch.cuda.current_device())
            if device != self.engine_device:
                self._rehome(device)
        return self  # 'cpu' (or dtype-only) requests: nothing to move, the engine stays on its GPU

    def cpu(self):
        return self

    def cuda(self, device=None):
        return self.to(torch.device("cuda", device) if isinstance(device, int) else (device or "cuda"))

    def _run(self, x: torch.Tensor, fn):
        if x.dtype != torch.float32:
            x = x.float()
        if x.is_cuda and x.device == self.engine_device:
            return fn(x.contiguous())
        y = fn(x.to(self.engine_device, non_blocking=True).contiguous())
        return y.to(x.device)


class B200StaticQuantizedNet(_GpuResident):
    """True static-PTQ int8 ``SimpleConvNet`` (all eight layers int8, fbgemm-exact requantisation) on a B200.

    Host inputs (what ``utils/model_evaluator.py`` hands over, and ``utils/inference_benchmark.py`` with
    ``device='cpu'``) are processed in chunks on two CUDA streams, so the host->device copy of chunk i+1 overlaps
    the kernels of chunk i; logits come back through a pinned staging buffer."""

    # images per pipelined chunk (B200Q_HOST_CHUNK overrides, for tuning).  The pipeline is bound by the host->device
    # copy (12 KiB of fp32 per image over PCIe), so what the chunk size controls is the un-overlapped tail: the kernels
    # of the LAST chunk.  2048 images = 24 MiB per copy, still far above the size where PCIe copies lose efficiency.
    HOST_CHUNK = int(os.environ.get("B200Q_HOST_CHUNK", "2048"))
    TAIL_MIN = int(os.environ.get("B200Q_TAIL_MIN", "256"))

    def __init__(self, qparams: dict, device=None):
        super().__init__(device)
        self.qparams = qparams
        self.engine = StaticEngine(qparams, self.engine_device)
        self._pipe = None

    def _rehome(self, device):
        self.engine = StaticEngine(self.qparams, device)
        self.engine_device = device
        self._pipe = None

    def _pipeline(self):
        if self._pipe is None:
            dev = self.engine_device
            self._pipe = {
                "streams": [torch.cuda.Stream(dev), torch.cuda.Stream(dev)],
                "x": [torch.empty((self.HOST_CHUNK, 3, 32, 32), dtype=torch.float32, device=dev) for _ in range(2)],
                "xu8": None,  # uint8 NHWC staging buffers, allocated on first use of forward_uint8
                "y": [torch.empty((self.HOST_CHUNK, 10), dtype=torch.float32, device=dev) for _ in range(2)],
                "out": None,
            }
        return self._pipe

    def _chunks(self, b: int, chunk: int | None = None):
        """(offset, size) of the pipelined chunks.  The copies run back to back, so the un-overlapped part of a call is
        the kernels of the LAST chunk; the final ``HOST_CHUNK`` images are therefore split 3/4 + 1/4 (one extra chunk:
        every chunk costs ~30 us of fixed overhead, and chunks much smaller than 512 images take longer to compute than
        to copy - measured, scripts/gpu_e2e_chunks.sh)."""
        chunk = chunk or self.HOST_CHUNK
        lo = 0
        while b - lo > chunk:
            yield lo, chunk
            lo += chunk
        rest = b - lo
        tail = rest // 4
        if tail >= self.TAIL_MIN:
            yield lo, rest - tail
            lo += rest - tail
        yield lo, b - lo

    def _forward_host(self, x: torch.Tensor, u8: bool = False) -> torch.Tensor:
        b = x.shape[0]
        x = x.contiguous()
        pipe = self._pipeline()
        # uint8 pixels are a quarter of the bytes: four times the images per chunk (same 24 MiB per copy), which also
        # keeps the kernels at an efficient batch size on a path that is compute- rather than PCIe-bound
        chunk = 4 * self.HOST_CHUNK if u8 else self.HOST_CHUNK
        if u8 and pipe["xu8"] is None:
            pipe["xu8"] = [torch.empty((chunk, 32, 32, 3), dtype=torch.uint8, device=self.engine_device) for _ in range(2)]
            pipe["yu8"] = [torch.empty((chunk, 10), dtype=torch.float32, device=self.engine_device) for _ in range(2)]
        if pipe["out"] is None or pipe["out"].shape[0] < b:
            pipe["out"] = torch.empty((max(b, self.HOST_CHUNK), 10), dtype=torch.float32).pin_memory()
        out = pipe["out"][:b]
        cur = torch.cuda.current_stream(self.engine_device)
        for s in pipe["streams"]:
            s.wait_stream(cur)
        for i, (lo, n) in enumerate(self._chunks(b, chunk)):
            k = i & 1
            with torch.cuda.stream(pipe["streams"][k]):  # per-stream buffers: reuse is ordered by the stream itself
                xin, yout = (pipe["xu8"] if u8 else pipe["x"])[k][:n], (pipe["yu8"] if u8 else pipe["y"])[k][:n]
                xin.copy_(x[lo:lo + n], non_blocking=True)
                if u8:
                    self.engine.forward_u8(xin, out=yout)
                else:
                    self.engine.forward(xin, out=yout)
                out[lo:lo + n].copy_(yout, non_blocking=True)
        for s in pipe["streams"]:
            s.synchronize()
        return out.clone()  # the pinned staging buffer is reused by the next call

    @torch.no_grad()
    def forward(self, x):
        if not x.is_cuda and x.dim() == 4 and x.shape[0] > 0:
            return self._forward_host(x.float())
        return self._run(x, self.engine.forward)

    @torch.no_grad()
    def forward_uint8(self, pixels: torch.Tensor) -> torch.Tensor:
        """uint8 data path: raw pixels, uint8 NHWC ``[B,32,32,3]`` as ``torchvision``'s CIFAR-10
        ``.data`` holds them, CPU or CUDA -> fp32 logits on the input's device.  Bit-identical to
        ``forward(Normalize(ToTensor(pixels)))``; a quarter of the bytes cross PCIe."""
        if pixels.dtype != torch.uint8 or pixels.dim() != 4 or tuple(pixels.shape[1:]) != (32, 32, 3):
            raise _lib.B200QError(f"forward_uint8 expects uint8 NHWC [B,32,32,3], got {pixels.dtype} {tuple(pixels.shape)}")
        if not pixels.is_cuda and pixels.shape[0] > 0:
            return self._forward_host(pixels, u8=True)
        if not pixels.is_cuda:
            return torch.empty((0, 10), dtype=torch.float32)
        return self.engine.forward_u8(pixels)

    @torch.no_grad()
    def forward_with_taps(self, x):
        """(logits, {layer: uint8 NHWC activations}) — parity-test hook."""
        return self.engine.forward(x.to(self.engine_device).float().contiguous(), taps=True)

    def state_dict(self, *args, **kwargs):
        """int8 weights + fp32 scales/bias, so ``get_model_size`` (torch.save of the state_dict,
        ``models/static_ptq_model.py:36-43``) reports the quantized size."""
        sd = {"quant.scale": torch.tensor(self.qparams["in_scale"]), "quant.zero_point": torch.tensor(self.qparams["in_zp"])}
        for name, L in self.qparams.items():
            if isinstance(L, dict):
                sd[f"{name}.weight"] = L["w_int8"]
                sd[f"{name}.weight_scales"] = L["w_scales"].float()
                sd[f"{name}.bias"] = L["bias"]
                sd[f"{name}.scale"] = torch.tensor(L["out_scale"])
                sd[f"{name}.zero_point"] = torch.tensor(L["out_zp"])
        return sd


class B200DynamicQuantizedNet(_GpuResident):
    """The reference's dynamic-PTQ net *as written* (SURVEY F3): ``quantize_dynamic`` only swaps ``fc1``/``fc2`` for
    ``DynamicQuantizedLinear``; the six BN-folded convs stay fp32.  Here the convs run as fp32 ATen/cuDNN ops (TF32 off;
    this is the tolerance path, not the product) and both linears run ``b200q_linear_dynamic`` (device-side min/max ->
    qparams -> quantize -> int8 GEMM -> fp32)."""

    def __init__(self, fused_fp32: nn.Module, fc_weights: dict, device=None, bn_after_fc1: nn.Module | None = None):
        """``bn_after_fc1``: eval-mode BatchNorm1d applied to fc1's output when fc1 was quantised WITHOUT the
        batch-norm folded in (``StaticPTQModel`` as written quantises the unfused net: fc1 -> bn7 -> relu)."""
        super().__init__(device)
        self._fused_cpu = fused_fp32
        self._fc_cpu = fc_weights  # name -> (w_int8, w_scale, bias)
        self._bn_cpu = None
        if bn_after_fc1 is not None:
            bn = bn_after_fc1
            scale = (bn.weight / torch.sqrt(bn.running_var + bn.eps)).detach().float()
            self._bn_cpu = (scale, (bn.bias - bn.running_mean * scale).detach().float())
        self._build(self.engine_device)

    def _build(self, device):
        self.convs = [(getattr(self._fused_cpu, f"conv{i}").weight.detach().to(device),
                       getattr(self._fused_cpu, f"conv{i}").bias.detach().to(device)) for i in range(1, 7)]
        self.fc = {n: ops.DynamicLinearWeights(w, s, b, device) for n, (w, s, b) in self._fc_cpu.items()}
        self.bn = None if self._bn_cpu is None else tuple(t.to(device) for t in self._bn_cpu)

    def _rehome(self, device):
        self._build(device)
        self.engine_device = device

    @torch.no_grad()
    def features(self, x):
        """fp32 conv stack (BN folded) -> ``[B,4096]`` in the reference's NCHW flatten order."""
        with torch.backends.cudnn.flags(enabled=True, allow_tf32=False):
            for i, (w, b) in enumerate(self.convs, start=1):
                x = F.relu(F.conv2d(x, w, b, padding=1))
                if i % 2 == 0:
                    x = F.max_pool2d(x, 2, 2)
        return x.reshape(x.shape[0], -1).contiguous()

    def _forward_dev(self, x):
        x = self.features(x)
        if self.bn is None:
            x = ops.linear_dynamic(x, self.fc["fc1"], relu=True)
        else:  # unfused fc1: batch-norm, then ReLU, in fp32 (tolerance path)
            x = F.relu(ops.linear_dynamic(x, self.fc["fc1"], relu=False) * self.bn[0] + self.bn[1]).contiguous()
        return ops.linear_dynamic(x, self.fc["fc2"], relu=False)

    @torch.no_grad()
    def forward(self, x):
        return self._run(x, self._forward_dev)

    def state_dict(self, *args, **kwargs):
        sd = {k: v for k, v in self._fused_cpu.state_dict().items() if not k.startswith(("fc1", "fc2"))}
        for n, (w, s, b) in self._fc_cpu.items():
            sd[f"{n}.weight"], sd[f"{n}.scale"], sd[f"{n}.bias"] = w, torch.tensor(s), b
        return sd
