"""``CustomQuantizationModel`` — mirror of ``models/custom_quantization_model.py:145-261`` (SimpleConvNet part only;
the ResNet50 wrappers in that file are out of scope).

As written, the reference wraps every layer in ``QuantStub``/``DeQuantStub`` but never calls ``prepare``/``convert``,
so the stubs are identity and the model is the BN-folded **fp32** net.  ``mode="as_written"`` (default)
reproduces exactly that (fp32 ATen ops on whatever device the driver moves the module to — the tolerance path);
``mode="sandwich"`` is the wrapper *as intended*: ``prepare``/``convert`` really applied to the per-layer
QuantStub -> int8 layer -> DeQuantStub sandwiches (with the two repairs without which the converted reference class
crashes), ReLU / pool in fp32, ``fc2`` fp32 - executed on the GPU by ``B200SandwichQuantizedNet`` and
bit-exact, layer by layer, against that converted torch model;
``mode="int8"`` is the fully fused reading (per-channel int8 weights with fused ReLU, activations stay int8 between
layers), i.e. the static-PTQ CUDA path.
"""
import torch
import torch.nn as nn
import torch.nn.functional as F

from .. import ptq, synth
from ._gpu_modules import B200SandwichQuantizedNet, B200StaticQuantizedNet
from .baseline_model import CONV_PLAN, FC_IN, SimpleConvNet


class _Sandwich(nn.Module):
    """quant -> layer -> dequant with un-converted (identity) stubs: ``CustomQuantizedConv2d`` / ``CustomQuantizedLinear``."""

    def __init__(self, layer, attr):
        super().__init__()
        self.quant = torch.ao.quantization.QuantStub()
        self.dequant = torch.ao.quantization.DeQuantStub()
        setattr(self, attr, layer)
        self._attr = attr

    def forward(self, x):
        return self.dequant(getattr(self, self._attr)(self.quant(x)))


class CustomQuantizedConv2d(_Sandwich):
    def __init__(self, conv_layer):
        super().__init__(conv_layer, "conv")


class CustomQuantizedLinear(_Sandwich):
    def __init__(self, linear_layer):
        super().__init__(linear_layer, "linear")


class CustomQuantizedSimpleConvNet(nn.Module):
    """Reference ``:202-261``: conv1..6 and fc1 in sandwiches, fc2 left fp32, ReLU/pool in fp32."""

    def __init__(self, model):
        super().__init__()
        self.quant = torch.ao.quantization.QuantStub()
        self.dequant = torch.ao.quantization.DeQuantStub()
        self.is_custom_quantized = True
        for i, _, _, pooled in CONV_PLAN:
            setattr(self, f"conv{i}", CustomQuantizedConv2d(getattr(model, f"conv{i}")))
            if pooled:
                setattr(self, f"pool{i // 2}", getattr(model, f"pool{i // 2}"))
                setattr(self, f"dropout{i // 2}", getattr(model, f"dropout{i // 2}"))
        self.fc1 = CustomQuantizedLinear(model.fc1)
        self.fc2 = model.fc2
        self.dropout4 = model.dropout4

    def forward(self, x):
        x = self.quant(x)
        for i, _, _, pooled in CONV_PLAN:
            x = F.relu(getattr(self, f"conv{i}")(x))
            if pooled:
                x = getattr(self, f"dropout{i // 2}")(getattr(self, f"pool{i // 2}")(x))
        x = x.reshape(-1, FC_IN)
        x = self.dropout4(F.relu(self.fc1(x)))
        return self.dequant(self.fc2(x))


class CustomQuantizationModel(nn.Module):
    def __init__(self, mode: str = "as_written", device=None):
        super().__init__()
        if mode not in ("as_written", "sandwich", "int8"):
            raise ValueError(f"unknown mode {mode!r}")
        self.mode = mode
        self.device = device
        self.model = SimpleConvNet()
        self.quantized_model = None
        self.is_custom_quantized = True
        ptq.select_engine()

    def load_state_dict(self, state_dict):
        self.model.load_state_dict(state_dict)

    def quantize(self, calibration_data_loader=None):
        self.model.eval()
        self.model = self.model.cpu()
        if self.mode in ("int8", "sandwich"):
            batches = synth.calibration_batches() if calibration_data_loader is None else (
                (b[0] if isinstance(b, (tuple, list)) else b) for b in calibration_data_loader)
            if self.mode == "int8":
                self.quantized_model = B200StaticQuantizedNet(ptq.calibrate_static(self.model, batches), self.device)
            else:
                self.quantized_model = B200SandwichQuantizedNet(ptq.calibrate_sandwich(self.model, batches), self.device)
            self.quantized_model.is_custom_quantized = True
        else:
            self.model = ptq.fuse_bn(self.model)
            self.quantized_model = CustomQuantizedSimpleConvNet(self.model)
        return self.quantized_model

    def forward(self, x):
        if self.quantized_model is not None:
            return self.quantized_model(x)
        return self.model(x)
