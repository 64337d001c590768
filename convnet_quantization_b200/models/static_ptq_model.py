"""``StaticPTQModel`` — same class / attributes / ``quantize`` signature as the reference
(``models/static_ptq_model.py:7-43``), re-targeted at the B200 engine.

The reference's method is static PTQ in name only: it calls ``quantize_dynamic`` and ignores its
``calibration_data_loader`` argument.  Two modes are therefore offered:

* ``mode="static"`` (default, the north-star path): true static PTQ — BN folded, activations calibrated with the
  fbgemm qconfig on ``calibration_data_loader`` (or, when it is ``None``, on the fixed synthetic calibration set),
  all eight layers int8, executed by the CUDA engine bit-exactly like torch's fbgemm CPU ops;
* ``mode="as_written"``: what the reference code really does — dynamic int8 on ``fc1``/``fc2`` of the *unfused* net,
  fp32 convolutions.
"""
import os
import tempfile

import torch
import torch.nn as nn

from .. import ptq, synth
from ._gpu_modules import B200DynamicQuantizedNet, B200StaticQuantizedNet
from .baseline_model import SimpleConvNet


def _calibration_tensors(loader, limit=None):
    n = 0
    for batch in loader:
        images = batch[0] if isinstance(batch, (tuple, list)) else batch
        yield images
        n += 1
        if limit is not None and n >= limit:
            return


class StaticPTQModel:
    def __init__(self, mode: str = "static", device=None):
        if mode not in ("static", "as_written"):
            raise ValueError(f"unknown mode {mode!r}")
        self.mode = mode
        self.device = device
        self.fp32_model = SimpleConvNet()
        self.quantized_model = None
        self.qparams = None

    def quantize(self, calibration_data_loader=None, calibration_device=None):
        """``calibration_device="cuda"``: run the calibration forward and the observers' reductions on the GPU
        (``ptq.B200HistogramObserver``).  Default: the host, one thread (reproducible scales)."""
        self.fp32_model.eval()
        if self.mode == "as_written":
            from .dynamic_ptq_model import dynamic_linear_weights
            net = self.fp32_model.cpu()
            # convolutions: BN folded for execution only (numerically the unfused eval-mode net); fc1 is quantised from
            # the UNFUSED weights as the reference does, so its batch-norm runs after the dynamic linear
            self.quantized_model = B200DynamicQuantizedNet(ptq.fuse_bn(net), dynamic_linear_weights(net, fused=False),
                                                           self.device, bn_after_fc1=net.bn7)
            return self.quantized_model
        batches = (synth.calibration_batches() if calibration_data_loader is None
                   else _calibration_tensors(calibration_data_loader))
        self.qparams = ptq.calibrate_static(self.fp32_model, batches, device=calibration_device)
        self.quantized_model = B200StaticQuantizedNet(self.qparams, self.device)
        return self.quantized_model

    def get_model_size(self, model):
        """Size in MB of ``torch.save(model.state_dict())`` (reference: ``models/static_ptq_model.py:36-43``)."""
        with tempfile.NamedTemporaryFile(suffix=".pth", delete=False) as f:
            path = f.name
        try:
            torch.save(model.state_dict(), path)
            return os.path.getsize(path) / (1024 * 1024)
        finally:
            os.remove(path)
