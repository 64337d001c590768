"""ORACLE (test infrastructure only — never imported by the product path).

Live CPU oracle for the quantized ``SimpleConvNet`` forward: PyTorch's own
eager-mode static PTQ ops with ``engine='fbgemm'``.  These ATen/FBGEMM
``QuantizedCPU`` kernels (``quantized::conv2d``, ``quantized::linear``,
``aten::quantize_per_tensor``, ``aten::quantized_max_pool2d``, ``aten::relu``,
``aten::dequantize``) are the third-party code the reference reaches through
``torch.quantization`` (call sites: ``models/dynamic_ptq_model.py:289-306``,
``models/static_ptq_model.py:28``, ``models/custom_quantization_model.py:180``);
they are not vendored under /root/reference and the reference pins no version,
so the de-facto pin is this image's torch 2.11.0+cu128.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline /
``--impl reference`` legs may import this module.

Parity pin: the reference holds no golden vectors for this path,
so the pin is (i) this live oracle, (ii) ``tests/golden/*.npz`` frozen from it
by ``tests/golden/make_golden.py`` and (iii) the integer restatement in
``oracle/int_ops.py`` asserted equal to (i).
"""
from __future__ import annotations

import copy

import torch
import torch.nn as nn
import torch.nn.functional as F
from torch.ao.quantization import (DeQuantStub, QuantStub, convert, fuse_modules,
                                   get_default_qconfig, prepare)

# fusion list: models/dynamic_ptq_model.py:289-299, models/custom_quantization_model.py:180-190
FUSE_LIST = [["conv1", "bn1"], ["conv2", "bn2"], ["conv3", "bn3"], ["conv4", "bn4"],
             ["conv5", "bn5"], ["conv6", "bn6"], ["fc1", "bn7"]]
LAYER_ORDER = ("quant", "conv1", "conv2", "pool1", "conv3", "conv4", "pool2",
               "conv5", "conv6", "pool3", "fc1", "fc2")


class StaticWrap(nn.Module):
    """QuantStub -> fused SimpleConvNet -> DeQuantStub.

    Forward restates ``models/baseline_model.py:58-83`` with ``reshape`` instead of
    ``view`` (a really-quantized conv output is channels-last); dropout
    is identity in eval mode and omitted.
    """

    def __init__(self, fused: nn.Module):
        super().__init__()
        self.quant = QuantStub()
        self.m = fused
        self.dequant = DeQuantStub()

    def forward(self, x, taps: dict | None = None):
        m = self.m

        def tap(name, t):
            if taps is not None:
                taps[name] = t
            return t

        x = tap("quant", self.quant(x))
        x = tap("conv1", F.relu(m.conv1(x)))
        x = tap("conv2", F.relu(m.conv2(x)))
        x = tap("pool1", m.pool1(x))
        x = tap("conv3", F.relu(m.conv3(x)))
        x = tap("conv4", F.relu(m.conv4(x)))
        x = tap("pool2", m.pool2(x))
        x = tap("conv5", F.relu(m.conv5(x)))
        x = tap("conv6", F.relu(m.conv6(x)))
        x = tap("pool3", m.pool3(x))
        x = x.reshape(-1, 256 * 4 * 4)
        x = tap("fc1", F.relu(m.fc1(x)))
        x = tap("fc2", m.fc2(x))
        return self.dequant(x)


def build_static_oracle(fp32_net: nn.Module, calib_batches) -> nn.Module:
    """fuse -> wrap -> prepare(fbgemm qconfig) -> calibrate -> convert.  Returns the CPU int8 model."""
    torch.backends.quantized.engine = "fbgemm"
    fused = fuse_modules(copy.deepcopy(fp32_net).cpu().eval(), FUSE_LIST, inplace=False)
    w = StaticWrap(fused).eval()
    w.qconfig = get_default_qconfig("fbgemm")
    p = prepare(w, inplace=False)
    nthreads = torch.get_num_threads()
    torch.set_num_threads(1)  # observed ranges must not depend on the fp32 summation order of a thread count
    try:
        with torch.no_grad():
            for xb in calib_batches:
                p(xb)
    finally:
        torch.set_num_threads(nthreads)
    return convert(p, inplace=False).eval()


def override_activation_qparams(qmodel: nn.Module, act: dict) -> nn.Module:
    """Replace the calibrated activation scales / zero-points of a converted model with frozen ones.

    ``act`` maps ``"in"`` and each layer name to ``(scale, zero_point)``.  Calibration runs fp32 MKL-DNN kernels
    whose last-ulp results depend on the host ISA, so golden-vector tests pin the observed ranges instead of
    re-deriving them on whatever CPU the test runs on; everything downstream is integer-exact.
    """
    s, z = act["in"]
    qmodel.quant.scale.fill_(float(s))
    qmodel.quant.zero_point.fill_(int(z))
    for name in ("conv1", "conv2", "conv3", "conv4", "conv5", "conv6", "fc1", "fc2"):
        s, z = act[name]
        mod = getattr(qmodel.m, name)
        mod.scale, mod.zero_point = float(s), int(z)
    return qmodel


@torch.no_grad()
def run_static_oracle(qmodel: nn.Module, x: torch.Tensor):
    """Returns ``(logits fp32 [B,10], taps)``; taps[name] = uint8 tensor in logical NCHW/[B,F] order."""
    torch.backends.quantized.engine = "fbgemm"
    taps: dict = {}
    logits = qmodel(x.cpu(), taps)
    return logits, {k: v.int_repr() for k, v in taps.items()}


def extract_qparams(qmodel: nn.Module) -> dict:
    """All integers/scales of the converted model, as plain tensors (for packing and golden files)."""
    out = {"in_scale": float(qmodel.quant.scale), "in_zp": int(qmodel.quant.zero_point)}
    for name in ("conv1", "conv2", "conv3", "conv4", "conv5", "conv6", "fc1", "fc2"):
        mod = getattr(qmodel.m, name)
        w = mod.weight()
        out[name] = {
            "w_int8": w.int_repr().clone(),
            "w_scales": w.q_per_channel_scales().clone(),  # float64
            "w_zps": w.q_per_channel_zero_points().clone(),
            "bias": mod.bias().detach().clone(),
            "out_scale": float(mod.scale),
            "out_zp": int(mod.zero_point),
        }
    return out


# ------------------------------------------------------------------ dynamic PTQ (as written in the reference)
def build_dynamic_oracle(fp32_net: nn.Module) -> nn.Module:
    """What ``DynamicPTQModel.quantize`` (``models/dynamic_ptq_model.py:281-308``) builds: BN folded, then
    ``quantize_dynamic({Linear, Conv2d}, qint8)`` - which converts fc1 / fc2 only."""
    torch.backends.quantized.engine = "fbgemm"
    fused = fuse_modules(copy.deepcopy(fp32_net).cpu().eval(), FUSE_LIST, inplace=False)
    return torch.ao.quantization.quantize_dynamic(fused, {nn.Linear, nn.Conv2d}, dtype=torch.qint8).eval()


# ------------------------------------------------------------------ the custom variant "as intended" (sandwiches)
SANDWICH_LAYERS = ("conv1", "conv2", "conv3", "conv4", "conv5", "conv6", "fc1")


class _Sandwich(nn.Module):
    """``CustomQuantizedConv2d`` / ``CustomQuantizedLinear`` (``models/custom_quantization_model.py:34-58``)."""

    def __init__(self, layer):
        super().__init__()
        self.quant = QuantStub()
        self.dequant = DeQuantStub()
        self.layer = layer

    def forward(self, x):
        return self.dequant(self.layer(self.quant(x)))


class SandwichWrap(nn.Module):
    """``CustomQuantizedSimpleConvNet`` (``models/custom_quantization_model.py:202-261``) in the one form in which
    ``prepare``/``convert`` can be applied to it (survey probe P3): no outer QuantStub/DeQuantStub (they double-quantise,
    SURVEY F5), ``reshape`` instead of ``view``, ``fc2`` fp32 (``:219``); dropout is identity in eval."""

    def __init__(self, fused: nn.Module):
        super().__init__()
        for name in SANDWICH_LAYERS:
            setattr(self, name, _Sandwich(getattr(fused, name)))
        self.fc2 = fused.fc2
        self.fc2.qconfig = None

    def forward(self, x, taps: dict | None = None):
        def run(name, t):
            sw = getattr(self, name)
            q = sw.layer(sw.quant(t))
            if taps is not None:
                taps[name] = q
            return F.relu(sw.dequant(q))

        x = run("conv2", run("conv1", x))
        x = F.max_pool2d(x, 2, 2)
        x = run("conv4", run("conv3", x))
        x = F.max_pool2d(x, 2, 2)
        x = run("conv6", run("conv5", x))
        x = F.max_pool2d(x, 2, 2)
        x = run("fc1", x.reshape(-1, 256 * 4 * 4))
        return self.fc2(x)


def build_sandwich_oracle(fp32_net: nn.Module, calib_batches) -> nn.Module:
    torch.backends.quantized.engine = "fbgemm"
    fused = fuse_modules(copy.deepcopy(fp32_net).cpu().eval(), FUSE_LIST, inplace=False)
    w = SandwichWrap(fused).eval()
    w.qconfig = get_default_qconfig("fbgemm")
    w.fc2.qconfig = None
    p = prepare(w, inplace=False)
    nthreads = torch.get_num_threads()
    torch.set_num_threads(1)
    try:
        with torch.no_grad():
            for xb in calib_batches:
                p(xb)
    finally:
        torch.set_num_threads(nthreads)
    return convert(p, inplace=False).eval()


@torch.no_grad()
def run_sandwich_oracle(qmodel: nn.Module, x: torch.Tensor):
    """``(logits fp32 [B,10], taps)``; taps[name] = uint8 output of the sandwich's int8 layer (pre-ReLU), logical
    NCHW / [B,F] order."""
    torch.backends.quantized.engine = "fbgemm"
    taps: dict = {}
    logits = qmodel(x.cpu(), taps)
    return logits, {k: v.int_repr() for k, v in taps.items()}


def sandwich_activation_qparams(qmodel: nn.Module) -> dict:
    """``{layer: (in_scale, in_zp, out_scale, out_zp)}`` of a converted sandwich model."""
    out = {}
    for name in SANDWICH_LAYERS:
        sw = getattr(qmodel, name)
        out[name] = (float(sw.quant.scale), int(sw.quant.zero_point), float(sw.layer.scale), int(sw.layer.zero_point))
    return out


def override_sandwich_qparams(qmodel: nn.Module, act: dict) -> nn.Module:
    """Frozen activation qparams (see ``override_activation_qparams``) for the sandwich model."""
    for name in SANDWICH_LAYERS:
        s_in, z_in, s_out, z_out = act[name]
        sw = getattr(qmodel, name)
        sw.quant.scale.fill_(float(s_in))
        sw.quant.zero_point.fill_(int(z_in))
        sw.layer.scale, sw.layer.zero_point = float(s_out), int(z_out)
    return qmodel
