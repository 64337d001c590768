"""ORACLE (test infrastructure only — never imported by the product path).

Pure-integer numpy restatement of the torch/FBGEMM quantized CPU ops the
reference reaches (call sites in
``models/dynamic_ptq_model.py:289-306``, ``models/static_ptq_model.py:28``).
The arithmetic lives in PyTorch 2.11.0 (ATen QuantizedCPU + FBGEMM), which is
not vendored in /root/reference, so each function restates the *published
behaviour* of the op and is pinned by ``tests/test_oracle.py`` against the live
torch ops (``oracle/torch_oracle.py``) and the frozen ``tests/golden`` vectors.

All "f32" steps are IEEE binary32, round-to-nearest-even, no FMA contraction:
numpy float32 scalar/array ops round after every operation, which is what is
needed.  int32 accumulators are computed exactly through float64 BLAS
(|acc| < 2^53).

Layout: activations are uint8 **NHWC**; conv weights int8 ``[Cout,Cin,kh,kw]``.
"""
from __future__ import annotations

import numpy as np

F32 = np.float32


def quantize_per_tensor(x: np.ndarray, scale: float, zp: int) -> np.ndarray:
    """aten::quantize_per_tensor -> quint8:  clamp(rne(x * (1/s)) + zp, 0, 255)."""
    inv = F32(1.0) / F32(scale)
    q = np.rint(x.astype(F32) * inv).astype(np.int64) + int(zp)
    return np.clip(q, 0, 255).astype(np.uint8)


def dequantize(q: np.ndarray, scale: float, zp: int) -> np.ndarray:
    """aten::dequantize:  f32(int32(q) - zp) * f32(s)."""
    return (q.astype(np.int32) - int(zp)).astype(F32) * F32(scale)


def requant_params(s_x: float, w_scales: np.ndarray, bias: np.ndarray, s_out: float):
    """Per-output-channel fp32 constants of the fbgemm requantisation (Appendix A)."""
    atw = F32(s_x) * np.asarray(w_scales).astype(F32)           # act_times_w_scale
    mult = (atw / F32(s_out)).astype(F32)
    bdiv = (np.asarray(bias).astype(F32) / atw).astype(F32)
    return mult, bdiv


def requantize(acc: np.ndarray, mult: np.ndarray, bdiv: np.ndarray, zp_out: int, relu: bool) -> np.ndarray:
    """s32 accumulator -> quint8:  t=f32(acc)+bdiv; t*=mult; clamp(rne(t)+zp, relu?zp:0, 255).  Channel is the last axis."""
    t = acc.astype(F32) + bdiv
    t = t * mult
    q = np.rint(t).astype(np.int64) + int(zp_out)
    lo = int(zp_out) if relu else 0
    return np.clip(q, lo, 255).astype(np.uint8)


def conv3x3_acc(x_u8: np.ndarray, zp_x: int, w_int8: np.ndarray) -> np.ndarray:
    """Exact s32 accumulator of a 3x3/s1/p1 conv with *real-domain* zero padding:
    acc[b,h,w,co] = sum_{valid taps} (x_q - zp_x) * w_q   (x NHWC, w [Cout,Cin,3,3])."""
    B, H, W, C = x_u8.shape
    cout = w_int8.shape[0]
    xc = x_u8.astype(np.float64) - float(zp_x)
    xp = np.zeros((B, H + 2, W + 2, C), dtype=np.float64)      # zero == real zero after centring
    xp[:, 1:H + 1, 1:W + 1, :] = xc
    acc = np.zeros((B * H * W, cout), dtype=np.float64)
    for kh in range(3):
        for kw in range(3):
            a = xp[:, kh:kh + H, kw:kw + W, :].reshape(B * H * W, C)
            acc += a @ w_int8[:, :, kh, kw].astype(np.float64).T
    return acc.reshape(B, H, W, cout).astype(np.int64)


def conv2d_q(x_u8, s_x, zp_x, w_int8, w_scales, bias, s_out, zp_out, relu=True):
    """quantized::conv2d (+ aten::relu):  uint8 NHWC in -> uint8 NHWC out."""
    acc = conv3x3_acc(x_u8, zp_x, w_int8)
    mult, bdiv = requant_params(s_x, w_scales, bias, s_out)
    return requantize(acc, mult, bdiv, zp_out, relu)


def max_pool2x2(x_u8: np.ndarray) -> np.ndarray:
    """aten::quantized_max_pool2d k2 s2 on raw uint8 values (qparams pass through); NHWC."""
    B, H, W, C = x_u8.shape
    return x_u8.reshape(B, H // 2, 2, W // 2, 2, C).max(axis=(2, 4))


def relu_q(x_u8: np.ndarray, zp: int) -> np.ndarray:
    """aten::relu on quint8 = max(q, zp)."""
    return np.maximum(x_u8, np.uint8(zp))


def linear_acc(x_u8: np.ndarray, zp_x: int, w_int8: np.ndarray) -> np.ndarray:
    xc = x_u8.astype(np.float64) - float(zp_x)
    return (xc @ w_int8.astype(np.float64).T).astype(np.int64)


def linear_q(x_u8, s_x, zp_x, w_int8, w_scales, bias, s_out, zp_out, relu=False):
    """quantized::linear (+relu): uint8 [B,K] -> uint8 [B,N]; w_scales per-channel [N] or scalar."""
    acc = linear_acc(x_u8, zp_x, w_int8)
    ws = np.broadcast_to(np.asarray(w_scales, dtype=np.float64), (w_int8.shape[0],))
    mult, bdiv = requant_params(s_x, ws, bias, s_out)
    return requantize(acc, mult, bdiv, zp_out, relu)


def flatten_nchw(x_nhwc: np.ndarray) -> np.ndarray:
    """The oracle's ``clone(NCHW) + view(-1, 4096)``: feature index = c*H*W + h*W + w."""
    B = x_nhwc.shape[0]
    return np.ascontiguousarray(x_nhwc.transpose(0, 3, 1, 2)).reshape(B, -1)


SMALL_SCALE_THRESHOLD = F32(6.1e-5)


def dynamic_qparams(mn: float, mx: float, qmin: int = 0, qmax: int = 127):
    """ChooseQuantizationParams(min, max, 0, 255, reduce_range=True) (ATen/native/quantized/cpu/QuantUtils.h; not under
    /root/reference - torch 2.11 is the de-facto pin) as quantized::linear_dynamic calls it: fp32 scale, nudged
    zero-point.  Pinned against ``torch._choose_qparams_per_tensor`` in tests/test_oracle.py."""
    mn = F32(min(F32(mn), F32(0.0)))
    mx = F32(max(F32(mx), F32(0.0)))
    scale = (np.float64(mx) - np.float64(mn)) / (qmax - qmin)
    with np.errstate(divide="ignore", over="ignore"):
        if F32(scale) == 0.0 or np.isinf(F32(1.0) / F32(scale)):
            scale = np.float64(0.1)
    if scale < np.float64(SMALL_SCALE_THRESHOLD):
        org_scale = F32(scale)
        scale = np.float64(SMALL_SCALE_THRESHOLD)
        if mn == 0.0:
            mx = F32(SMALL_SCALE_THRESHOLD * F32(qmax - qmin))
        elif mx == 0.0:
            mn = F32(-(SMALL_SCALE_THRESHOLD * F32(qmax - qmin)))
        else:
            amplifier = F32(SMALL_SCALE_THRESHOLD / org_scale)
            mn = F32(mn * amplifier)
            mx = F32(mx * amplifier)
    zp_from_min = qmin - np.float64(mn) / scale
    zp_from_max = qmax - np.float64(mx) / scale
    err_min = abs(qmin) - abs(np.float64(mn) / scale)
    err_max = abs(qmax) - abs(np.float64(mx) / scale)
    izp = zp_from_min if err_min < err_max else zp_from_max
    if izp < qmin:
        zp = qmin
    elif izp > qmax:
        zp = qmax
    else:
        zp = int(np.rint(izp))
    return F32(scale), zp


def linear_dynamic(x_f32: np.ndarray, w_int8: np.ndarray, w_scale: float, bias: np.ndarray) -> np.ndarray:
    """quantized::linear_dynamic(x, W, reduce_range=True): fp32 [B,K] -> fp32 [B,N].
    Per-tensor min/max over the WHOLE input; weights per-tensor symmetric qint8.  Activations are quantised the way
    fbgemm's PackAWithQuantRowOffset does it: the zero-point is added in fp32 before the rounding, as one fused
    multiply-add, ``q = clamp(rne(fma(x, 1/s, zp)), 0, 255)`` (the float64 product of two binary32 numbers is exact, so
    rounding the float64 sum to binary32 IS the single rounding of an fma).  tests/test_oracle.py pins this form
    against the live op (``rne(x/s)+zp`` differs on ~3 elements per million)."""
    x = x_f32.astype(F32)
    s_x, zp = dynamic_qparams(x.min(), x.max())
    inv = F32(1.0) / s_x
    xq = np.clip(np.rint((x.astype(np.float64) * np.float64(inv) + np.float64(zp)).astype(F32)).astype(np.int64), 0, 255)
    acc = ((xq - zp).astype(np.float64) @ w_int8.astype(np.float64).T)
    # output stage: ONE fused multiply-add fma(f32(acc), fl32(s_x * s_w), bias) (|acc| < 2^29, so the float64 product and
    # sum below are exact and the final cast is the fma's single rounding); bit-identical to the live op
    s_xw = np.float64(F32(s_x * F32(w_scale)))
    return (acc.astype(F32).astype(np.float64) * s_xw + bias.astype(F32).astype(np.float64)).astype(F32)


def static_forward(x_f32_nchw: np.ndarray, qp: dict, taps: dict | None = None) -> np.ndarray:
    """Whole static-PTQ SimpleConvNet forward in integers (call order of
    ``models/baseline_model.py:58-83`` on the converted model); returns fp32 logits [B,10]."""
    def tap(n, v):
        if taps is not None:
            taps[n] = v
        return v

    x = np.ascontiguousarray(np.asarray(x_f32_nchw, dtype=F32).transpose(0, 2, 3, 1))  # NHWC
    s, zp = qp["in_scale"], qp["in_zp"]
    x = tap("quant", quantize_per_tensor(x, s, zp))
    for i in range(1, 7):
        L = qp[f"conv{i}"]
        x = tap(f"conv{i}", conv2d_q(x, s, zp, np.asarray(L["w_int8"]), np.asarray(L["w_scales"]),
                                     np.asarray(L["bias"]), L["out_scale"], L["out_zp"], relu=True))
        s, zp = L["out_scale"], L["out_zp"]
        if i % 2 == 0:
            x = tap(f"pool{i // 2}", max_pool2x2(x))
    x = flatten_nchw(x)
    for name, relu in (("fc1", True), ("fc2", False)):
        L = qp[name]
        x = tap(name, linear_q(x, s, zp, np.asarray(L["w_int8"]), np.asarray(L["w_scales"]),
                               np.asarray(L["bias"]), L["out_scale"], L["out_zp"], relu=relu))
        s, zp = L["out_scale"], L["out_zp"]
    return dequantize(x, s, zp)


def histc(x: np.ndarray, bins: int, lo: float, hi: float) -> np.ndarray:
    """torch.histc (ATen CPU, linear bins, no local search): bin = int(((x - lo) * bins) / (hi - lo)) evaluated in fp32
    in that order; bin == bins -> bins - 1; values outside [lo, hi] are dropped; a degenerate range is widened by 1
    on both sides.  Rule found by experiment against torch 2.11 (tests/test_oracle.py pins it)."""
    lo, hi = F32(lo), F32(hi)
    if lo == hi:
        lo, hi = lo - F32(1), hi + F32(1)
    x = np.asarray(x, dtype=F32).reshape(-1)
    x = x[(x >= lo) & (x <= hi)]
    pos = (((x - lo) * F32(bins)) / (hi - lo)).astype(np.int64)
    pos[pos >= bins] = bins - 1
    return np.bincount(pos, minlength=bins).astype(np.int64)


def sandwich_lut(out_scale: float, out_zp: int, next_scale: float, next_zp: int) -> np.ndarray:
    """DeQuantStub -> fp32 ReLU -> next QuantStub as a uint8 -> uint8 table (custom variant as intended,
    ``models/custom_quantization_model.py:34-58``): lut[q] = quantize(max(dequantize(q), 0))."""
    v = np.maximum(dequantize(np.arange(256, dtype=np.uint8), out_scale, out_zp), F32(0.0))
    return quantize_per_tensor(v, next_scale, next_zp)


def sandwich_forward(x_f32_nchw: np.ndarray, sp: dict, taps: dict | None = None) -> np.ndarray:
    """The per-layer sandwich net in integers (call order of ``models/custom_quantization_model.py:233-261``):
    ``sp[layer]`` = {in_scale, in_zp, w_int8, w_scales, bias, out_scale, out_zp} for conv1..conv6 and fc1, ``sp["fc2"]``
    = {weight, bias} fp32.  ReLU / max-pool between the sandwiches run on the quantized values through the monotone
    boundary table (equal to running them in fp32 on the dequantised tensors).  Returns fp32 logits; taps[layer] = uint8
    NHWC output of each int8 layer, before ReLU."""
    names = ("conv1", "conv2", "conv3", "conv4", "conv5", "conv6", "fc1")
    x = np.ascontiguousarray(np.asarray(x_f32_nchw, dtype=F32).transpose(0, 2, 3, 1))
    L = sp["conv1"]
    a = quantize_per_tensor(x, L["in_scale"], L["in_zp"])
    for i, name in enumerate(names[:6]):
        L = sp[name]
        a = conv2d_q(a, L["in_scale"], L["in_zp"], np.asarray(L["w_int8"]), np.asarray(L["w_scales"]),
                     np.asarray(L["bias"]), L["out_scale"], L["out_zp"], relu=False)
        if taps is not None:
            taps[name] = a
        nxt = sp[names[i + 1]]
        a = sandwich_lut(L["out_scale"], L["out_zp"], nxt["in_scale"], nxt["in_zp"])[a]
        if i % 2 == 1:
            a = max_pool2x2(a)
    L = sp["fc1"]
    h = linear_q(flatten_nchw(a), L["in_scale"], L["in_zp"], np.asarray(L["w_int8"]), np.asarray(L["w_scales"]),
                 np.asarray(L["bias"]), L["out_scale"], L["out_zp"], relu=False)
    if taps is not None:
        taps["fc1"] = h
    hf = np.maximum(dequantize(h, L["out_scale"], L["out_zp"]), F32(0.0))
    return (hf.astype(np.float64) @ np.asarray(sp["fc2"]["weight"], dtype=np.float64).T
            + np.asarray(sp["fc2"]["bias"], dtype=np.float64)).astype(F32)
