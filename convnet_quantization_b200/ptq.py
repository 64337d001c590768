"""Static post-training quantisation *preparation* (host side, runs once per model).

This is model conversion, not the forward hot path: BN folding, calibration observers and weight
quantisation are done with the same ``torch.ao.quantization`` eager-mode calls the reference uses
(``fuse_modules`` at ``models/dynamic_ptq_model.py:289-299`` / ``models/custom_quantization_model.py:180-190``;
``QuantStub``/``DeQuantStub``/``prepare``/``convert``/``get_default_qconfig`` imported at
``models/custom_quantization_model.py:5``), so the resulting integers are identical to what a torch CPU
model would hold.  The converted torch module is thrown away; only its integers/scales are kept and
packed for the CUDA engine (``packing.py``).  GPU-side calibration is SURVEY.md §8(f) rank 2 ("next").
"""
from __future__ import annotations

import copy
import warnings

import torch
import torch.nn as nn
import torch.nn.functional as F

FUSE_LIST = [[f"conv{i}", f"bn{i}"] for i in range(1, 7)] + [["fc1", "bn7"]]
QUANT_LAYERS = ("conv1", "conv2", "conv3", "conv4", "conv5", "conv6", "fc1", "fc2")


def select_engine() -> str:
    """The reference's global side effect (``models/dynamic_ptq_model.py:227-232``): prefer fbgemm."""
    engines = torch.backends.quantized.supported_engines
    if "fbgemm" in engines:
        torch.backends.quantized.engine = "fbgemm"
    elif "qnnpack" in engines:
        torch.backends.quantized.engine = "qnnpack"
    else:
        raise RuntimeError("No supported quantization engine found")
    return torch.backends.quantized.engine


def fuse_bn(fp32_net: nn.Module) -> nn.Module:
    """conv+bn x6 and fc1+bn7 folded (eval mode), on CPU, as a copy."""
    net = copy.deepcopy(fp32_net).cpu().eval()
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        return torch.ao.quantization.fuse_modules(net, FUSE_LIST, inplace=False)


def fold_identity(fp32_net: nn.Module) -> nn.Module:
    """BN folded into the preceding conv/linear for *execution* only (numerically the unfused eval-mode net up to
    fp32 rounding); used by ``StaticPTQModel(mode="as_written")`` whose reference code does not fuse."""
    return fuse_bn(fp32_net)


class _CalibWrap(nn.Module):
    """QuantStub -> BN-folded SimpleConvNet -> DeQuantStub; forward = ``models/baseline_model.py:58-83``
    with ``reshape`` in place of ``view`` and dropout (identity in eval) dropped."""

    def __init__(self, fused):
        super().__init__()
        self.quant = torch.ao.quantization.QuantStub()
        self.m = fused
        self.dequant = torch.ao.quantization.DeQuantStub()

    def forward(self, x):
        m = self.m
        x = self.quant(x)
        x = m.pool1(F.relu(m.conv2(F.relu(m.conv1(x)))))
        x = m.pool2(F.relu(m.conv4(F.relu(m.conv3(x)))))
        x = m.pool3(F.relu(m.conv6(F.relu(m.conv5(x)))))
        x = x.reshape(x.shape[0], -1)
        return self.dequant(m.fc2(F.relu(m.fc1(x))))


class single_thread:
    """Calibration runs the fp32 net on the CPU; its summation order (and with it the last ulp of the observed
    activation ranges) depends on the intra-op thread count.  Pinning one thread makes the derived scales a function
    of (weights, calibration images) only."""

    def __enter__(self):
        self.n = torch.get_num_threads()
        torch.set_num_threads(1)

    def __exit__(self, *exc):
        torch.set_num_threads(self.n)


def calibrate_static(fp32_net: nn.Module, calib_batches) -> dict:
    """Returns the static-PTQ parameter dict:
    ``{"in_scale", "in_zp", layer: {"w_int8", "w_scales"(f64), "bias"(f32), "out_scale", "out_zp"}}``."""
    select_engine()
    with warnings.catch_warnings(), single_thread():
        warnings.simplefilter("ignore")
        wrap = _CalibWrap(fuse_bn(fp32_net)).eval()
        wrap.qconfig = torch.ao.quantization.get_default_qconfig("fbgemm")
        prepared = torch.ao.quantization.prepare(wrap, inplace=False)
        with torch.no_grad():
            for xb in calib_batches:
                prepared(xb.detach().cpu().float())
        q = torch.ao.quantization.convert(prepared, inplace=False)
    out = {"in_scale": float(q.quant.scale), "in_zp": int(q.quant.zero_point)}
    for name in QUANT_LAYERS:
        mod = getattr(q.m, name)
        w = mod.weight()
        if int(w.q_per_channel_zero_points().abs().max()) != 0:
            raise RuntimeError(f"{name}: expected symmetric per-channel weights")
        out[name] = {"w_int8": w.int_repr().clone(), "w_scales": w.q_per_channel_scales().clone(),
                     "bias": mod.bias().detach().clone().float(), "out_scale": float(mod.scale),
                     "out_zp": int(mod.zero_point)}
    return out
