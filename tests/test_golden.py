"""Oracle vs the frozen golden vectors (tests/golden/convnet_golden.npz, made by tests/golden/make_golden.py from the
reference's own classes and torch's fbgemm ops).  CPU only."""
import hashlib
import os

import numpy as np
import pytest
import torch

from convnet_quantization_b200 import synth
from oracle import int_ops as IO
from oracle import torch_oracle as TO
from tests.conftest import qparams_to_numpy

def sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def test_synthetic_inputs_and_weights_are_reproducible(golden):
    assert np.array_equal(synth.images_u8(16, seed=42).numpy(), golden["x_u8"])
    h = hashlib.sha256()
    sd = synth.make_state_dict(0)
    for k in sorted(sd):
        h.update(k.encode())
        h.update(sd[k].numpy().tobytes())
    assert h.hexdigest() == str(golden["sd_sha"])


def test_static_oracles_match_golden(golden, golden_oracle):
    x = synth.normalize(torch.from_numpy(golden["x_u8"])).contiguous()
    logits, taps = TO.run_static_oracle(golden_oracle, x)
    assert np.array_equal(logits.numpy(), golden["static_logits"])
    qp = qparams_to_numpy(TO.extract_qparams(golden_oracle))
    assert qp["in_scale"] == float(golden["in_scale"]) and qp["in_zp"] == int(golden["in_zp"])
    mine = {}
    out = IO.static_forward(x.numpy(), qp, mine)
    assert np.array_equal(out, golden["static_logits"])
    for k in TO.LAYER_ORDER:
        assert sha(mine[k]) == str(golden[f"static_{k}_sha"]), k
        assert np.array_equal(mine[k].reshape(-1)[:256], golden[f"static_{k}_head"]), k
    for name in ("conv1", "conv2", "conv3", "conv4", "conv5", "conv6", "fc1", "fc2"):
        assert qp[name]["out_scale"] == float(golden[f"{name}_out_scale"])
        assert qp[name]["out_zp"] == int(golden[f"{name}_out_zp"])
        assert sha(qp[name]["w_int8"]) == str(golden[f"{name}_w_sha"])


def test_fp32_mirror_matches_reference_class(golden, fp32_net):
    """Our SimpleConvNet mirror == the reference's class output (frozen), CPU fp32."""
    x = synth.normalize(torch.from_numpy(golden["x_u8"])).contiguous()
    with torch.no_grad():
        got = fp32_net(x).numpy()
    np.testing.assert_allclose(got, golden["ref_fp32"], rtol=1e-4, atol=1e-4)


def test_custom_as_written_matches_reference_class(golden):
    """CustomQuantizationModel as written is the BN-folded fp32 net; runs on CPU via torch."""
    from convnet_quantization_b200.models.custom_quantization_model import CustomQuantizationModel
    m = CustomQuantizationModel()
    m.load_state_dict(synth.make_state_dict(0))
    m.quantize()
    x = synth.normalize(torch.from_numpy(golden["x_u8"])).contiguous()
    with torch.no_grad():
        got = m(x).numpy()
    np.testing.assert_allclose(got, golden["ref_custom"], rtol=1e-4, atol=1e-4)
    assert m.is_custom_quantized and m.quantized_model.is_custom_quantized
