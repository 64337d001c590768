"""Pin the oracle: the numpy integer restatement (oracle/int_ops.py) must equal the live torch/fbgemm CPU ops
(oracle/torch_oracle.py) bit for bit, on every intermediate activation and on the logits, including
out-of-calibration (saturating) inputs; and the product-side calibration must yield the oracle's integers."""
import numpy as np
import pytest
import torch

from convnet_quantization_b200 import synth
from oracle import int_ops as IO
from oracle import torch_oracle as TO
from tests.conftest import qparams_to_numpy


@pytest.mark.parametrize("seed,gain", [(5, 1.0), (6, 3.0)])
def test_int_restatement_equals_torch(oracle_model, seed, gain):
    x = synth.images_f32(8, seed) * gain
    logits, taps = TO.run_static_oracle(oracle_model, x)
    qp = qparams_to_numpy(TO.extract_qparams(oracle_model))
    mine = {}
    out = IO.static_forward(x.numpy(), qp, mine)
    for k in TO.LAYER_ORDER:
        a = taps[k].numpy()
        if a.ndim == 4:
            a = a.transpose(0, 2, 3, 1)
        assert np.array_equal(a, mine[k]), k
    assert np.array_equal(out, logits.numpy())


def test_product_calibration_equals_oracle(qparams, oracle_model):
    ref = TO.extract_qparams(oracle_model)
    assert qparams["in_scale"] == ref["in_scale"] and qparams["in_zp"] == ref["in_zp"]
    for name in ("conv1", "conv2", "conv3", "conv4", "conv5", "conv6", "fc1", "fc2"):
        a, b = qparams[name], ref[name]
        assert torch.equal(a["w_int8"], b["w_int8"]), name
        assert torch.equal(a["w_scales"], b["w_scales"]), name
        assert torch.equal(a["bias"], b["bias"].float()), name
        assert a["out_scale"] == b["out_scale"] and a["out_zp"] == b["out_zp"], name


def test_elementwise_restatements_vs_torch():
    g = torch.Generator().manual_seed(0)
    x = torch.randn(1 << 16, generator=g) * 3
    x[:5] = torch.tensor([0.5, 1.5, 2.5, -0.5, -2.5]) * 0.05
    q = torch.quantize_per_tensor(x, 0.05, 31, torch.quint8)
    mine = IO.quantize_per_tensor(x.numpy(), 0.05, 31)
    assert np.array_equal(mine, q.int_repr().numpy())
    assert np.array_equal(IO.dequantize(mine, 0.05, 31), q.dequantize().numpy())
    assert np.array_equal(IO.relu_q(mine, 31), torch.relu(q).int_repr().numpy())
    q4 = torch.quantize_per_tensor(torch.randn(2, 16, 8, 8, generator=g), 0.03, 100, torch.quint8)
    pooled = torch.nn.functional.max_pool2d(q4, 2, 2).int_repr().numpy().transpose(0, 2, 3, 1)
    assert np.array_equal(IO.max_pool2x2(q4.int_repr().numpy().transpose(0, 2, 3, 1)), pooled)


@pytest.mark.parametrize("b,k,n", [(4, 4096, 512), (64, 512, 10)])
def test_linear_dynamic_restatement(b, k, n):
    torch.backends.quantized.engine = "fbgemm"
    g = torch.Generator().manual_seed(k)
    lin = torch.nn.Linear(k, n)
    with torch.no_grad():
        lin.weight.copy_(torch.randn(n, k, generator=g) * 0.05)
    qlin = torch.ao.quantization.quantize_dynamic(torch.nn.Sequential(lin), {torch.nn.Linear}, dtype=torch.qint8)[0]
    x = torch.randn(b, k, generator=g).abs()
    want = qlin(x).numpy()
    w = qlin.weight()
    got = IO.linear_dynamic(x.numpy(), w.int_repr().numpy(), w.q_scale(), qlin.bias().detach().numpy())
    np.testing.assert_allclose(got, want, rtol=1e-5, atol=1e-5 * np.abs(want).max())
