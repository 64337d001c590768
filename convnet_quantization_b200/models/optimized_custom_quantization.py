"""``OptimizedCustomQuantization`` — class and ``quantize(model)`` / ``get_model_size(model)`` signatures of
``models/optimized_custom_quantization.py:7-136``.

The reference body is ResNet50/ImageNet-only and not runnable (pretrained download at ``:13``; ``qconfig_dict=`` kwarg
``TypeError`` at ``:41-45``; all three "importance" branches assign the same qconfig, ``:118-126``).  Here the
evident intent — Conv-BN-ReLU fusion + per-channel int8 — is applied to ``SimpleConvNet`` (BASELINE.json config 4):
``quantize(model)`` returns the static-PTQ CUDA module with ``.quantized`` and ``.is_custom_quantized`` set (``:48-49``).
Parity for this class is unpinned by the reference; it is pinned to the same fbgemm oracle as ``StaticPTQModel``.
"""
import os
import tempfile

import torch

from .. import ptq, synth
from ._gpu_modules import B200StaticQuantizedNet
from .baseline_model import SimpleConvNet


class OptimizedCustomQuantization:
    def __init__(self, device=None):
        self.device = device
        self.fp32_model = SimpleConvNet()  # reference builds a pretrained ResNet50 here (:13); no network, not the north-star net
        self.quantized_model = None
        ptq.select_engine()

    def quantize(self, model, calibration_data_loader=None):
        if not isinstance(model, SimpleConvNet) and not all(hasattr(model, f"conv{i}") for i in range(1, 7)):
            raise TypeError("OptimizedCustomQuantization.quantize expects a SimpleConvNet-shaped model")
        model = model.cpu().eval()
        batches = synth.calibration_batches() if calibration_data_loader is None else (
            (b[0] if isinstance(b, (tuple, list)) else b) for b in calibration_data_loader)
        quantized_model = B200StaticQuantizedNet(ptq.calibrate_static(model, batches), self.device)
        quantized_model.quantized = True
        quantized_model.is_custom_quantized = True
        self.quantized_model = quantized_model
        return quantized_model

    def get_model_size(self, model):
        with tempfile.NamedTemporaryFile(suffix=".pth", delete=False) as f:
            path = f.name
        try:
            torch.save(model.state_dict(), path)
            return os.path.getsize(path) / (1024 * 1024)
        finally:
            os.remove(path)
