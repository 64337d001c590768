"""``DynamicPTQModel`` — the reference's dynamic-PTQ wrapper (``models/dynamic_ptq_model.py:218-317``) with the same
methods (``load_state_dict/eval/cpu/to/forward/__call__/quantize/get_model_size``) and attributes (``fp32_model``,
``quantized_model``), executing on a B200.

``quantize()`` follows the reference: fold conv+bn x6 and fc1+bn7 (``:289-299``), then ``quantize_dynamic`` over
``{Linear, Conv2d}`` (``:302-306``) — which in PyTorch only converts the two ``Linear`` layers (per-tensor symmetric
qint8 weights); the convolutions stay fp32.  Weight quantisation is host-side preparation and reuses torch's
observer; the forward runs ``b200q_linear_dynamic`` on device.
"""
import os
import tempfile
import warnings

import torch

from .. import ptq
from ._gpu_modules import B200DynamicQuantizedNet
from .baseline_model import SimpleConvNet


def dynamic_linear_weights(net: torch.nn.Module, fused: bool = True) -> dict:
    """``{name: (w_int8, w_scale, bias)}`` exactly as ``quantize_dynamic(..., dtype=qint8)`` packs fc1 / fc2."""
    ptq.select_engine()
    out = {}
    for name in ("fc1", "fc2"):
        lin = getattr(net, name)
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            q = torch.ao.quantization.quantize_dynamic(torch.nn.Sequential(lin), {torch.nn.Linear}, dtype=torch.qint8)[0]
        w = q.weight()
        out[name] = (w.int_repr().clone(), float(w.q_scale()), q.bias().detach().clone().float())
    return out


class DynamicPTQModel:
    def __init__(self, device=None):
        self.device = device
        self.fp32_model = SimpleConvNet()
        self.quantized_model = None
        ptq.select_engine()  # same global side effect as the reference (:227-232)

    def load_state_dict(self, state_dict):
        self.fp32_model.load_state_dict(state_dict)

    def eval(self):
        (self.quantized_model if self.quantized_model is not None else self.fp32_model).eval()
        return self

    def cpu(self):
        if self.quantized_model is not None:
            self.quantized_model = self.quantized_model.cpu()  # no-op: the engine stays on its GPU
        else:
            self.fp32_model = self.fp32_model.cpu()
        return self

    def to(self, device):
        if self.quantized_model is not None:
            self.quantized_model = self.quantized_model.to(device)
        else:
            self.fp32_model = self.fp32_model.to(device)
        return self

    def forward(self, x):
        if self.quantized_model is not None:
            return self.quantized_model(x)
        return self.fp32_model(x)

    def __call__(self, x):
        return self.forward(x)

    def quantize(self):
        self.fp32_model = self.fp32_model.cpu().eval()
        self.fp32_model = ptq.fuse_bn(self.fp32_model)
        self.quantized_model = B200DynamicQuantizedNet(self.fp32_model, dynamic_linear_weights(self.fp32_model),
                                                       self.device)
        return self.quantized_model

    def get_model_size(self):
        with tempfile.NamedTemporaryFile(suffix=".pth", delete=False) as f:
            path = f.name
        try:
            torch.save(self.quantized_model.state_dict(), path)
            return os.path.getsize(path) / (1024 * 1024)
        finally:
            os.remove(path)
