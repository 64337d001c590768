"""Operator layer: plain-tensor ops on CUDA uint8/int8/fp32 tensors that replace the ATen ``quantized::*`` CPU ops
the reference reaches.  Each function is a thin shim: argument checks, output allocation, one call
through the C ABI (``include/b200q.h``) on the current CUDA stream.  The same functions are registered under the
``torch.ops.b200q`` namespace (``torch.library``).  There is no CPU implementation: CPU tensors raise.

ATen originals, for orientation:
  aten::quantize_per_tensor(Tensor, float scale, int zero_point, ScalarType) -> Tensor
  quantized::conv2d.new(Tensor qx, Conv2dPackedParamsBase w, float output_scale, int output_zero_point) -> Tensor
  quantized::linear(Tensor X, LinearPackedParamsBase W, float Y_scale_i, int Y_zero_point_i) -> Tensor
  quantized::linear_dynamic(Tensor X, LinearPackedParamsBase W, bool reduce_range=False) -> Tensor
  aten::quantized_max_pool2d / aten::relu (quint8) / aten::dequantize
"""
from __future__ import annotations

import torch

from . import _lib
from .packing import PackedConv, PackedLinear

_U8 = torch.uint8


def _need_cuda(*tensors):
    for t in tensors:
        if not t.is_cuda:
            raise _lib.B200QError("b200q ops run on CUDA tensors only (no CPU fallback)")
        if not t.is_contiguous():
            raise _lib.B200QError("b200q ops need contiguous tensors")


def _stream() -> int:
    return torch.cuda.current_stream().cuda_stream


def _inv_scale(scale: float) -> float:
    return float(torch.tensor(1.0, dtype=torch.float32) / torch.tensor(float(scale), dtype=torch.float32))


def quantize_per_tensor(x: torch.Tensor, scale: float, zero_point: int, c_pad: int | None = None) -> torch.Tensor:
    """fp32 NCHW ``[B,C,H,W]`` -> uint8 NHWC ``[B,H,W,c_pad]`` (default c_pad=C; pad channels hold zero_point)."""
    _need_cuda(x)
    b, c, h, w = x.shape
    c_pad = c if c_pad is None else c_pad
    y = torch.empty((b, h, w, c_pad), dtype=_U8, device=x.device)
    _lib.check(_lib.load().b200q_quantize_nchw_to_nhwc(x.data_ptr(), y.data_ptr(), b, c, h, w, c_pad,
                                                      _inv_scale(scale), int(zero_point), _stream()), "quantize_per_tensor")
    return y


def quantize_flat(x: torch.Tensor, scale: float, zero_point: int) -> torch.Tensor:
    """Layout-preserving quantize of any contiguous fp32 tensor."""
    _need_cuda(x)
    y = torch.empty(x.shape, dtype=_U8, device=x.device)
    _lib.check(_lib.load().b200q_quantize_flat(x.data_ptr(), y.data_ptr(), x.numel(), _inv_scale(scale),
                                              int(zero_point), _stream()), "quantize_flat")
    return y


def dequantize(q: torch.Tensor, scale: float, zero_point: int) -> torch.Tensor:
    _need_cuda(q)
    y = torch.empty(q.shape, dtype=torch.float32, device=q.device)
    _lib.check(_lib.load().b200q_dequantize(q.data_ptr(), y.data_ptr(), q.numel(), float(scale), int(zero_point),
                                           _stream()), "dequantize")
    return y


def relu_q(q: torch.Tensor, zero_point: int) -> torch.Tensor:
    _need_cuda(q)
    y = torch.empty_like(q)
    _lib.check(_lib.load().b200q_relu_q(q.data_ptr(), y.data_ptr(), q.numel(), int(zero_point), _stream()), "relu_q")
    return y


def max_pool2d_q(x: torch.Tensor) -> torch.Tensor:
    """2x2/2 max-pool on uint8 NHWC."""
    _need_cuda(x)
    b, h, w, c = x.shape
    y = torch.empty((b, h // 2, w // 2, c), dtype=_U8, device=x.device)
    _lib.check(_lib.load().b200q_max_pool2x2_nhwc(x.data_ptr(), y.data_ptr(), b, h, w, c, _stream()), "max_pool2d_q")
    return y


def minmax(x: torch.Tensor) -> torch.Tensor:
    """Returns device tensor ``[min(x,0), max(x,0), scale, 1/scale, zero_point]`` (fbgemm reduce_range qparams)."""
    _need_cuda(x)
    out = torch.empty(8, dtype=torch.float32, device=x.device)
    scratch = torch.zeros(2 * 1024 + 8, dtype=torch.float32, device=x.device)
    _lib.check(_lib.load().b200q_minmax(x.data_ptr(), x.numel(), out.data_ptr(), scratch.data_ptr(), _stream()), "minmax")
    return out[:5]


def conv2d_q(x: torch.Tensor, w: PackedConv, pool2x2: bool = False, impl: str = "tc") -> torch.Tensor:
    """uint8 NHWC ``[B,img,img,cin]`` -> uint8 NHWC ``[B,img,img,cout]`` (requant + ReLU per the packed layer)."""
    _need_cuda(x)
    b = x.shape[0]
    if tuple(x.shape[1:]) != (w.img, w.img, w.cin):
        raise _lib.B200QError(f"conv2d_q: input {tuple(x.shape)} does not match layer {w.name} ({w.img},{w.img},{w.cin})")
    o = w.img // 2 if pool2x2 else w.img
    y = torch.empty((b, o, o, w.cout), dtype=_U8, device=x.device)
    lib = _lib.load()
    if impl == "tc":
        rc = lib.b200q_conv3x3_tc(x.data_ptr(), y.data_ptr(), b, w.ptr(), int(pool2x2), _stream())
    elif impl == "simt":
        rc = lib.b200q_conv3x3_simt(x.data_ptr(), y.data_ptr(), b, w.ptr(), _stream())
    elif impl == "first":
        rc = lib.b200q_conv3x3_first(x.data_ptr(), y.data_ptr(), b, w.ptr(), _stream())
    else:
        raise ValueError(impl)
    _lib.check(rc, f"conv2d_q[{impl}]")
    return y


def quantize_conv2d_first(x: torch.Tensor, scale: float, w: PackedConv) -> torch.Tensor:
    """Fused QuantStub + conv1(+ReLU): fp32 NCHW ``[B,3,32,32]`` -> uint8 NHWC ``[B,32,32,64]``."""
    _need_cuda(x)
    b = x.shape[0]
    y = torch.empty((b, w.img, w.img, w.cout), dtype=_U8, device=x.device)
    _lib.check(_lib.load().b200q_quantize_conv3x3_first(x.data_ptr(), y.data_ptr(), b, _inv_scale(scale), w.ptr(),
                                                       _stream()), "quantize_conv2d_first")
    return y


def quantize_conv2d_conv2d_pool(x: torch.Tensor, scale: float, w1: PackedConv, w2: PackedConv) -> torch.Tensor:
    """fp32 NCHW ``[B,3,32,32]`` -> quantize -> conv1+ReLU -> conv2+ReLU -> 2x2 max-pool -> uint8 NHWC ``[B,16,16,64]``
    in one kernel (``b200q_conv12_fused``)."""
    _need_cuda(x)
    b = x.shape[0]
    y = torch.empty((b, w2.img // 2, w2.img // 2, w2.cout), dtype=_U8, device=x.device)
    _lib.check(_lib.load().b200q_conv12_fused(x.data_ptr(), y.data_ptr(), b, _inv_scale(scale), w1.ptr(), w2.ptr(),
                                              _stream()), "conv12_fused")
    return y


def linear_q(x: torch.Tensor, w: PackedLinear, impl: str = "tc") -> torch.Tensor:
    _need_cuda(x)
    b = x.shape[0]
    y = torch.empty((b, w.n), dtype=_U8, device=x.device)
    lib = _lib.load()
    fn = lib.b200q_linear_tc if impl == "tc" else lib.b200q_linear_simt
    _lib.check(fn(x.data_ptr(), y.data_ptr(), b, w.ptr(), _stream()), f"linear_q[{impl}]")
    return y


def linear_dequant(x: torch.Tensor, w: PackedLinear, out_scale: float) -> torch.Tensor:
    _need_cuda(x)
    b = x.shape[0]
    y = torch.empty((b, w.n), dtype=torch.float32, device=x.device)
    _lib.check(_lib.load().b200q_linear_dequant(x.data_ptr(), y.data_ptr(), b, w.ptr(), float(out_scale), _stream()),
               "linear_dequant")
    return y


class DynamicLinearWeights:
    """Per-tensor symmetric qint8 weights of an ``nnqd.Linear`` (``default_dynamic_qconfig``), on device."""

    def __init__(self, w_int8: torch.Tensor, w_scale: float, bias: torch.Tensor, device):
        self.n, self.k = w_int8.shape
        self.w = w_int8.detach().to(torch.int8).contiguous().to(device)
        self.wsum = w_int8.detach().to(torch.int64).sum(dim=1).to(torch.int32).contiguous().to(device)
        self.w_scale = float(w_scale)
        self.bias = bias.detach().float().contiguous().to(device)
        self.scratch = torch.zeros(2 * 1024 + 16, dtype=torch.float32, device=device)


def linear_dynamic(x: torch.Tensor, w: DynamicLinearWeights, relu: bool = False) -> torch.Tensor:
    """``quantized::linear_dynamic(x, W, reduce_range=True)``: fp32 ``[B,K]`` -> fp32 ``[B,N]``; the activation
    scale/zero-point come from the min/max of the WHOLE input tensor, computed on device (no host sync)."""
    _need_cuda(x)
    b, k = x.shape
    y = torch.empty((b, w.n), dtype=torch.float32, device=x.device)
    xq = torch.empty((b, k), dtype=_U8, device=x.device)
    _lib.check(_lib.load().b200q_linear_dynamic(x.data_ptr(), y.data_ptr(), b, k, w.n, w.w.data_ptr(),
                                                w.wsum.data_ptr(), w.w_scale, w.bias.data_ptr(), int(relu),
                                                xq.data_ptr(), w.scratch.data_ptr(), _stream()), "linear_dynamic")
    return y


# ------------------------------------------------------------------ torch.library registration (b200q::*)
_registered = False


def register_torch_ops() -> None:
    """Expose the plain-tensor ops as ``torch.ops.b200q.*`` (CUDA dispatch key only — no CPU kernels)."""
    global _registered
    if _registered:
        return
    lib = torch.library.Library("b200q", "DEF")
    lib.define("quantize_per_tensor(Tensor x, float scale, int zero_point, int c_pad) -> Tensor")
    lib.define("dequantize(Tensor q, float scale, int zero_point) -> Tensor")
    lib.define("relu_q(Tensor q, int zero_point) -> Tensor")
    lib.define("max_pool2d_q(Tensor x) -> Tensor")
    lib.define("minmax(Tensor x) -> Tensor")
    lib.impl("quantize_per_tensor", lambda x, s, z, c: quantize_per_tensor(x, s, z, c), "CUDA")
    lib.impl("dequantize", dequantize, "CUDA")
    lib.impl("relu_q", relu_q, "CUDA")
    lib.impl("max_pool2d_q", max_pool2d_q, "CUDA")
    lib.impl("minmax", minmax, "CUDA")
    register_torch_ops._lib = lib  # keep alive
    _registered = True
