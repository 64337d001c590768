import os
import sys
import warnings

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

warnings.filterwarnings("ignore")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA B200 device (run with -m gpu on the GPU box)")


def pytest_collection_modifyitems(config, items):
    if torch.cuda.is_available():
        return
    skip = pytest.mark.skip(reason="no CUDA device")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


@pytest.fixture(scope="session")
def fp32_net():
    from convnet_quantization_b200 import synth
    from convnet_quantization_b200.models.baseline_model import SimpleConvNet
    net = SimpleConvNet()
    net.load_state_dict(synth.make_state_dict(0))
    return net.eval()


@pytest.fixture(scope="session")
def qparams(fp32_net):
    """Static-PTQ parameters as the PRODUCT derives them (ptq.calibrate_static)."""
    from convnet_quantization_b200 import ptq, synth
    return ptq.calibrate_static(fp32_net, synth.calibration_batches())


@pytest.fixture(scope="session")
def oracle_model(fp32_net):
    """Live torch/fbgemm CPU oracle (oracle/torch_oracle.py)."""
    from convnet_quantization_b200 import synth
    from oracle import torch_oracle
    return torch_oracle.build_static_oracle(fp32_net, synth.calibration_batches())


def qparams_to_numpy(qp):
    out = {}
    for k, v in qp.items():
        if isinstance(v, dict):
            out[k] = {kk: (vv.detach().cpu().numpy() if torch.is_tensor(vv) else vv) for kk, vv in v.items()}
        else:
            out[k] = v
    return out


@pytest.fixture(scope="session")
def qparams_np(qparams):
    return qparams_to_numpy(qparams)


@pytest.fixture(scope="session")
def lib():
    from convnet_quantization_b200 import _lib
    return _lib.load(build_if_missing=True)
