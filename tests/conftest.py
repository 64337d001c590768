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


GOLDEN_PATH = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "convnet_golden.npz")
_LAYERS = ("conv1", "conv2", "conv3", "conv4", "conv5", "conv6", "fc1", "fc2")


def golden_activation_qparams(g) -> dict:
    """Frozen activation (scale, zero_point) pairs of the golden file: calibration is fp32 CPU work whose last ulp
    depends on the host ISA, so golden comparisons pin the observed ranges (the rest is integer-exact)."""
    act = {"in": (float(g["in_scale"]), int(g["in_zp"]))}
    for n in _LAYERS:
        act[n] = (float(g[f"{n}_out_scale"]), int(g[f"{n}_out_zp"]))
    return act


@pytest.fixture(scope="session")
def golden():
    return np.load(GOLDEN_PATH)


@pytest.fixture(scope="session")
def golden_qparams(fp32_net, golden):
    """Product-derived weights/bias + the golden file's activation qparams."""
    import copy
    from convnet_quantization_b200 import ptq, synth
    qp = copy.deepcopy(ptq.calibrate_static(fp32_net, synth.calibration_batches()))
    act = golden_activation_qparams(golden)
    qp["in_scale"], qp["in_zp"] = act["in"]
    for n in _LAYERS:
        qp[n]["out_scale"], qp[n]["out_zp"] = act[n]
    return qp


@pytest.fixture(scope="session")
def golden_oracle(fp32_net, golden):
    """torch/fbgemm oracle with the golden file's activation qparams."""
    from convnet_quantization_b200 import synth
    from oracle import torch_oracle
    q = torch_oracle.build_static_oracle(fp32_net, synth.calibration_batches())
    return torch_oracle.override_activation_qparams(q, golden_activation_qparams(golden))
