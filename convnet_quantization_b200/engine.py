"""Device engine for the static-PTQ SimpleConvNet: packed weights + workspace + the single C-ABI forward call."""
from __future__ import annotations

import ctypes as C

import torch

from . import _lib
from .packing import PackedStaticNet

TAP_NAMES = ("quant", "conv1", "conv2", "pool1", "conv3", "conv4", "pool2", "conv5", "conv6", "pool3", "fc1", "fc2")
TAP_SHAPES = {  # per image, NHWC (quant is NHWC4: channel 3 is padding)
    "quant": (32, 32, 4), "conv1": (32, 32, 64), "conv2": (32, 32, 64), "pool1": (16, 16, 64),
    "conv3": (16, 16, 128), "conv4": (16, 16, 128), "pool2": (8, 8, 128), "conv5": (8, 8, 256),
    "conv6": (8, 8, 256), "pool3": (4, 4, 256), "fc1": (512,), "fc2": (10,),
}


class StaticEngine:
    """Runs ``b200q_static_forward`` on one CUDA device.  fp32 NCHW ``[B,3,32,32]`` (CUDA) -> fp32 logits ``[B,10]``."""

    def __init__(self, qparams: dict, device="cuda"):
        self.device = torch.device(device)
        if self.device.type != "cuda":
            raise _lib.B200QError("StaticEngine needs a CUDA device (no CPU fallback)")
        if self.device.index is None:
            self.device = torch.device("cuda", torch.cuda.current_device())
        self.lib = _lib.load()
        self.qparams = qparams
        with torch.cuda.device(self.device):
            self.packed = PackedStaticNet(qparams, self.device)
        self._ws = {}  # CUDA stream handle -> workspace (forwards on different streams must not share one)

    def _workspace(self, b: int) -> torch.Tensor:
        need = int(self.lib.b200q_static_workspace_bytes(b))
        key = torch.cuda.current_stream(self.device).cuda_stream
        ws = self._ws.get(key)
        if ws is None or ws.numel() < need:
            ws = torch.empty(need, dtype=torch.uint8, device=self.device)
            self._ws[key] = ws
        return ws

    @torch.no_grad()
    def forward(self, x: torch.Tensor, taps: bool = False, out: torch.Tensor | None = None):
        """``out`` (optional): preallocated fp32 ``[B,10]`` CUDA tensor to receive the logits."""
        if not x.is_cuda:
            raise _lib.B200QError("StaticEngine.forward expects a CUDA tensor")
        x = x.contiguous().float()
        if x.dim() != 4 or tuple(x.shape[1:]) != (3, 32, 32):
            raise _lib.B200QError(f"expected [B,3,32,32] input, got {tuple(x.shape)}")
        b = x.shape[0]
        with torch.cuda.device(self.device):
            logits = out if out is not None else torch.empty((b, 10), dtype=torch.float32, device=self.device)
            if b == 0:
                return (logits, {}) if taps else logits
            ws = self._workspace(b)
            tap_ptrs = None
            tap_tensors = {}
            if taps:
                arr = (C.c_void_p * len(TAP_NAMES))()
                for i, name in enumerate(TAP_NAMES):
                    t = torch.empty((b,) + TAP_SHAPES[name], dtype=torch.uint8, device=self.device)
                    tap_tensors[name] = t
                    arr[i] = t.data_ptr()
                tap_ptrs = arr
            rc = self.lib.b200q_static_forward(self.packed.ptr(), x.data_ptr(), logits.data_ptr(), b, ws.data_ptr(),
                                               ws.numel(), tap_ptrs, torch.cuda.current_stream().cuda_stream)
            _lib.check(rc, "static_forward")
        return (logits, tap_tensors) if taps else logits

    __call__ = forward

    @torch.no_grad()
    def forward_u8(self, x_u8: torch.Tensor, out: torch.Tensor | None = None) -> torch.Tensor:
        """uint8 data path: raw pixels, uint8 NHWC ``[B,32,32,3]`` (CUDA) -> fp32 logits ``[B,10]``; bit-identical to
        ``forward(synth.normalize(pixels))`` (``packing.input_lut``)."""
        if not x_u8.is_cuda or x_u8.dtype != torch.uint8:
            raise _lib.B200QError("StaticEngine.forward_u8 expects a CUDA uint8 tensor")
        x_u8 = x_u8.contiguous()
        if x_u8.dim() != 4 or tuple(x_u8.shape[1:]) != (32, 32, 3):
            raise _lib.B200QError(f"expected uint8 NHWC [B,32,32,3] input, got {tuple(x_u8.shape)}")
        b = x_u8.shape[0]
        with torch.cuda.device(self.device):
            logits = out if out is not None else torch.empty((b, 10), dtype=torch.float32, device=self.device)
            if b == 0:
                return logits
            ws = self._workspace(b)
            rc = self.lib.b200q_static_forward_u8(self.packed.ptr(), x_u8.data_ptr(), self.packed.input_lut.data_ptr(),
                                                  logits.data_ptr(), b, ws.data_ptr(), ws.numel(),
                                                  torch.cuda.current_stream().cuda_stream)
            _lib.check(rc, "static_forward_u8")
        return logits

    def stage_names(self):
        return [self.lib.b200q_static_stage_name(i).decode() for i in range(self.lib.b200q_static_num_stages())]

    @torch.no_grad()
    def forward_profiled(self, x: torch.Tensor):
        """(logits, {stage: ms}) — one forward with a CUDA event between kernels (measurement hook; synchronises)."""
        x = x.contiguous().float()
        b = x.shape[0]
        n = self.lib.b200q_static_num_stages()
        ms = (C.c_float * n)()
        with torch.cuda.device(self.device):
            logits = torch.empty((b, 10), dtype=torch.float32, device=self.device)
            ws = self._workspace(b)
            rc = self.lib.b200q_static_forward_profiled(self.packed.ptr(), x.data_ptr(), logits.data_ptr(), b,
                                                        ws.data_ptr(), ws.numel(), ms,
                                                        torch.cuda.current_stream().cuda_stream)
            _lib.check(rc, "static_forward_profiled")
        return logits, dict(zip(self.stage_names(), (float(v) for v in ms)))
