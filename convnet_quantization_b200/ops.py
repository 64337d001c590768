"""Operator layer: plain-tensor ops on CUDA uint8/int8/fp32 tensors that replace the ATen ``quantized::*`` CPU ops
the reference reaches.  Each function is a thin shim: argument checks, output allocation, one call
through the C ABI (``include/b200q.h``) on the current CUDA stream.  The same functions are registered under the
``torch.ops.b200q`` namespace (``torch.library``, CUDA dispatch key only) when the package is imported.  There is no
CPU implementation: CPU tensors raise.

ATen originals, for orientation:
  aten::quantize_per_tensor(Tensor, float scale, int zero_point, ScalarType) -> Tensor
  quantized::conv2d_prepack(Tensor weight, Tensor? bias, int[] stride, ...) -> Conv2dPackedParamsBase
  quantized::conv2d.new(Tensor qx, Conv2dPackedParamsBase w, float output_scale, int output_zero_point) -> Tensor
  quantized::linear_prepack(Tensor W, Tensor? B) -> LinearPackedParamsBase
  quantized::linear(Tensor X, LinearPackedParamsBase W, float Y_scale_i, int Y_zero_point_i) -> Tensor
  quantized::linear_dynamic(Tensor X, LinearPackedParamsBase W, bool reduce_range=False) -> Tensor
  aten::quantized_max_pool2d / aten::relu (quint8) / aten::dequantize
"""
from __future__ import annotations

import itertools
import weakref

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


def lut_u8(x: torch.Tensor, lut: torch.Tensor) -> torch.Tensor:
    """``y[i] = lut[x[i]]`` on uint8 data; ``lut``: uint8 ``[256]`` on the HOST (it travels as a kernel parameter)."""
    _need_cuda(x)
    if lut.is_cuda or lut.dtype != _U8 or lut.numel() != 256:
        raise _lib.B200QError("lut_u8: lut must be a CPU uint8 tensor of 256 entries")
    lut = lut.contiguous()
    y = torch.empty_like(x)
    _lib.check(_lib.load().b200q_lut_u8(x.data_ptr(), y.data_ptr(), x.numel(), lut.data_ptr(), _stream()), "lut_u8")
    return y


def _reduce_scratch(device) -> torch.Tensor:
    return torch.zeros(_lib.REDUCE_SCRATCH_BYTES // 4, dtype=torch.float32, device=device)


def minmax(x: torch.Tensor) -> torch.Tensor:
    """Returns device tensor ``[min(x,0), max(x,0), scale, 1/scale, zero_point]`` (fbgemm reduce_range qparams)."""
    _need_cuda(x)
    out = torch.empty(8, dtype=torch.float32, device=x.device)
    _lib.check(_lib.load().b200q_minmax(x.data_ptr(), x.numel(), out.data_ptr(), _reduce_scratch(x.device).data_ptr(),
                                        _stream()), "minmax")
    return out[:5]


def aminmax(x: torch.Tensor) -> torch.Tensor:
    """``torch.aminmax`` of an fp32 tensor as a device tensor ``[min, max]`` (calibration observers)."""
    _need_cuda(x)
    out = torch.empty(2, dtype=torch.float32, device=x.device)
    _lib.check(_lib.load().b200q_aminmax(x.data_ptr(), x.numel(), out.data_ptr(), _reduce_scratch(x.device).data_ptr(),
                                         _stream()), "aminmax")
    return out


def histc(x: torch.Tensor, bins: int, lo: float, hi: float) -> torch.Tensor:
    """``torch.histc(x, bins, min=lo, max=hi)`` with ATen's CPU binning rule; int64 counts on the device."""
    _need_cuda(x)
    lo, hi = float(lo), float(hi)
    if lo == hi:  # ATen widens a degenerate range the same way
        lo, hi = lo - 1.0, hi + 1.0
    hist = torch.zeros(int(bins), dtype=torch.int64, device=x.device)
    _lib.check(_lib.load().b200q_histc(x.data_ptr(), x.numel(), lo, hi, int(bins), hist.data_ptr(), _stream()), "histc")
    return hist


def conv2d_q(x: torch.Tensor, w: PackedConv, pool2x2: bool = False, impl: str = "tc") -> torch.Tensor:
    """uint8 NHWC ``[B,img,img,cin]`` -> uint8 NHWC ``[B,img,img,cout]`` (requant + ReLU per the packed layer).
    ``impl``: ``"tc"`` tensor cores (product), ``"first"`` the CUDA-core first layer, ``"simt"`` the bring-up
    cross-check kernel of the DEVELOPMENT library (tests only)."""
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
        lib = _lib.load_dev()
        rc = lib.b200q_conv3x3_simt(x.data_ptr(), y.data_ptr(), b, w.ptr(), _stream())
    elif impl == "first":
        rc = lib.b200q_conv3x3_first(x.data_ptr(), y.data_ptr(), b, w.ptr(), _stream())
    else:
        raise ValueError(impl)
    _lib.check(rc, f"conv2d_q[{impl}]", lib)
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
    """DEVELOPMENT library only (``b200q_conv12_fused``, measured slower than the two kernels it replaces):
    fp32 NCHW ``[B,3,32,32]`` -> quantize -> conv1+ReLU -> conv2+ReLU -> 2x2 max-pool -> uint8 NHWC ``[B,16,16,64]``."""
    _need_cuda(x)
    b = x.shape[0]
    y = torch.empty((b, w2.img // 2, w2.img // 2, w2.cout), dtype=_U8, device=x.device)
    lib = _lib.load_dev()
    _lib.check(lib.b200q_conv12_fused(x.data_ptr(), y.data_ptr(), b, _inv_scale(scale), w1.ptr(), w2.ptr(), _stream()),
               "conv12_fused", lib)
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
        self.scratch = _reduce_scratch(device)  # B200Q_REDUCE_SCRATCH_BYTES, zeroed once, self-resetting

    def last_qparams(self) -> torch.Tensor:
        """Device view ``[min, max, scale, 1/scale, zero_point]`` of the activation qparams of the last call."""
        o = _lib.REDUCE_QPARAMS_OFFSET // 4
        return self.scratch[o:o + 5]


def linear_dynamic(x: torch.Tensor, w: DynamicLinearWeights, relu: bool = False) -> torch.Tensor:
    """``quantized::linear_dynamic(x, W, reduce_range=True)``: fp32 ``[B,K]`` -> fp32 ``[B,N]``; the activation
    scale/zero-point come from the min/max of the WHOLE input tensor, computed on device (no host sync), and the
    input is quantised inside the tcgen05 GEMM's producer warps."""
    _need_cuda(x)
    b, k = x.shape
    y = torch.empty((b, w.n), dtype=torch.float32, device=x.device)
    if b == 0:
        return y
    _lib.check(_lib.load().b200q_linear_dynamic(x.data_ptr(), y.data_ptr(), b, k, w.n, w.w.data_ptr(),
                                                w.wsum.data_ptr(), w.w_scale, w.bias.data_ptr(), int(relu),
                                                w.scratch.data_ptr(), w.scratch.numel() * 4, _stream()), "linear_dynamic")
    return y


# ------------------------------------------------------------------ torch.library registration (b200q::*)
# Packed parameters travel through the dispatcher as an opaque int64 handle tensor (what ATen does with
# Conv2dPackedParamsBase / LinearPackedParamsBase custom-class objects): the handle indexes a registry that keeps the
# PackedConv / PackedLinear / DynamicLinearWeights alive until the handle tensor is garbage-collected.
_handles: dict[int, object] = {}
_next_handle = itertools.count(1)


def _new_handle(obj) -> torch.Tensor:
    key = next(_next_handle)
    _handles[key] = obj
    h = torch.tensor([key], dtype=torch.int64)
    weakref.finalize(h, _handles.pop, key, None)
    return h


def _packed(handle: torch.Tensor, kind):
    obj = _handles.get(int(handle.item()))
    if not isinstance(obj, kind):
        raise _lib.B200QError(f"b200q: handle does not refer to a live {kind.__name__}")
    return obj


def conv_prepack(weight: torch.Tensor, w_scales: torch.Tensor, bias: torch.Tensor, in_scale: float, in_zero_point: int,
                 out_scale: float, out_zero_point: int, relu: bool, device: str) -> torch.Tensor:
    """int8 OIHW ``[Cout,Cin,3,3]`` weights (``qconv.weight().int_repr()``), per-channel scales, fp32 bias and the
    activation qparams either side -> handle of a :class:`PackedConv` on ``device``."""
    cout, cin = int(weight.shape[0]), int(weight.shape[1])
    name = {(64, 3): "conv1", (64, 64): "conv2", (128, 64): "conv3", (128, 128): "conv4", (256, 128): "conv5",
            (256, 256): "conv6"}.get((cout, cin))
    if name is None:
        raise _lib.B200QError(f"conv_prepack: no kernel for a {cin}->{cout} 3x3 layer")
    layer = {"w_int8": weight.cpu(), "w_scales": w_scales.cpu(), "bias": bias.cpu(), "out_scale": float(out_scale),
             "out_zp": int(out_zero_point)}
    return _new_handle(PackedConv(name, layer, float(in_scale), int(in_zero_point), torch.device(device), relu=bool(relu)))


def linear_prepack(weight: torch.Tensor, w_scales: torch.Tensor, bias: torch.Tensor, in_scale: float, in_zero_point: int,
                   out_scale: float, out_zero_point: int, relu: bool, nhwc_from: list[int], device: str) -> torch.Tensor:
    """int8 ``[N,K]`` weights -> handle of a :class:`PackedLinear`; ``nhwc_from = [c,h,w]`` permutes the K columns from
    the reference's NCHW flatten order to the engine's NHWC order (``[]``: keep)."""
    layer = {"w_int8": weight.cpu(), "w_scales": w_scales.cpu(), "bias": bias.cpu(), "out_scale": float(out_scale),
             "out_zp": int(out_zero_point)}
    return _new_handle(PackedLinear("linear", layer, float(in_scale), int(in_zero_point), torch.device(device),
                                    relu=bool(relu), nhwc_from=tuple(nhwc_from) if len(nhwc_from) else None))


def linear_dynamic_prepack(weight: torch.Tensor, w_scale: float, bias: torch.Tensor, device: str) -> torch.Tensor:
    """Per-tensor symmetric int8 ``[N,K]`` weights of an ``nnqd.Linear`` -> handle of :class:`DynamicLinearWeights`."""
    return _new_handle(DynamicLinearWeights(weight.cpu(), float(w_scale), bias.cpu(), torch.device(device)))


_registered = False
_torch_lib = None


def register_torch_ops() -> None:
    """Expose the ops as ``torch.ops.b200q.*``.  Compute ops are registered for the CUDA dispatch key only (no CPU
    kernels: a CPU tensor raises torch's own "no kernel for backend CPU" error); the ``*_prepack`` ops take host
    tensors like their ATen counterparts and are backend-independent."""
    global _registered, _torch_lib
    if _registered:
        return
    lib = torch.library.Library("b200q", "DEF")
    lib.define("quantize_per_tensor(Tensor x, float scale, int zero_point, int c_pad) -> Tensor")
    lib.define("quantize_flat(Tensor x, float scale, int zero_point) -> Tensor")
    lib.define("dequantize(Tensor q, float scale, int zero_point) -> Tensor")
    lib.define("relu_q(Tensor q, int zero_point) -> Tensor")
    lib.define("max_pool2d_q(Tensor x) -> Tensor")
    lib.define("lut_u8(Tensor x, Tensor lut) -> Tensor")
    lib.define("minmax(Tensor x) -> Tensor")
    lib.define("aminmax(Tensor x) -> Tensor")
    lib.define("histc(Tensor x, int bins, float lo, float hi) -> Tensor")
    lib.define("conv_prepack(Tensor weight, Tensor w_scales, Tensor bias, float in_scale, int in_zero_point, "
               "float out_scale, int out_zero_point, bool relu, str device) -> Tensor")
    lib.define("linear_prepack(Tensor weight, Tensor w_scales, Tensor bias, float in_scale, int in_zero_point, "
               "float out_scale, int out_zero_point, bool relu, int[] nhwc_from, str device) -> Tensor")
    lib.define("linear_dynamic_prepack(Tensor weight, float w_scale, Tensor bias, str device) -> Tensor")
    lib.define("conv2d_q(Tensor x, Tensor packed, bool pool2x2) -> Tensor")
    lib.define("quantize_conv2d_first(Tensor x, float scale, Tensor packed) -> Tensor")
    lib.define("linear_q(Tensor x, Tensor packed) -> Tensor")
    lib.define("linear_dequant(Tensor x, Tensor packed, float out_scale) -> Tensor")
    lib.define("linear_dynamic(Tensor x, Tensor packed, bool relu) -> Tensor")
    lib.impl("quantize_per_tensor", lambda x, s, z, c: quantize_per_tensor(x, s, z, c), "CUDA")
    lib.impl("quantize_flat", quantize_flat, "CUDA")
    lib.impl("dequantize", dequantize, "CUDA")
    lib.impl("relu_q", relu_q, "CUDA")
    lib.impl("max_pool2d_q", max_pool2d_q, "CUDA")
    lib.impl("lut_u8", lut_u8, "CUDA")
    lib.impl("minmax", minmax, "CUDA")
    lib.impl("aminmax", aminmax, "CUDA")
    lib.impl("histc", histc, "CUDA")
    lib.impl("conv_prepack", conv_prepack, "CompositeExplicitAutograd")
    lib.impl("linear_prepack", linear_prepack, "CompositeExplicitAutograd")
    lib.impl("linear_dynamic_prepack", linear_dynamic_prepack, "CompositeExplicitAutograd")
    lib.impl("conv2d_q", lambda x, h, pool: conv2d_q(x, _packed(h, PackedConv), bool(pool)), "CUDA")
    lib.impl("quantize_conv2d_first", lambda x, s, h: quantize_conv2d_first(x, s, _packed(h, PackedConv)), "CUDA")
    lib.impl("linear_q", lambda x, h: linear_q(x, _packed(h, PackedLinear)), "CUDA")
    lib.impl("linear_dequant", lambda x, h, s: linear_dequant(x, _packed(h, PackedLinear), s), "CUDA")
    lib.impl("linear_dynamic", lambda x, h, relu: linear_dynamic(x, _packed(h, DynamicLinearWeights), bool(relu)), "CUDA")
    _torch_lib = lib  # keep alive
    _registered = True
