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
    assert np.array_equal(got, want)  # BIT-exact: fma activation quantisation + fma output stage (fbgemm's own forms)


def test_dynamic_qparams_restatement_vs_torch():
    """ChooseQuantizationParams incl. the small-scale cut-off and one-sided / tiny / degenerate ranges
    (ADVICE r1): the restatement must equal torch's own binding of the ATen routine."""
    g = torch.Generator().manual_seed(0)
    cases = [(0.0, 0.0), (1e-30, 2e-30), (-1e-38, 1e-38), (-3.0, 1e-7), (-1e-7, 3.0), (-2.43, 2.75)]
    for e in np.linspace(-9, 3, 120):
        a, b = (float(10 ** e) * float(torch.rand(1, generator=g) + 0.1) for _ in range(2))
        cases += [(-a, b), (0.0, b), (-a, 0.0), (0.5 * a, a), (-a, -0.5 * a)]
    for lo, hi in cases:
        x = torch.tensor([lo, hi], dtype=torch.float32)
        s, z = torch._choose_qparams_per_tensor(x, True)
        s2, z2 = IO.dynamic_qparams(float(x.min()), float(x.max()))
        assert np.float32(s) == s2 and z == z2, (lo, hi)


def test_histc_restatement_vs_torch():
    g = torch.Generator().manual_seed(3)
    for trial in range(40):
        n = int(torch.randint(1, 100000, (1,), generator=g))
        bins = (2048, 100, 37, 4096)[trial % 4]
        if trial % 2:
            lo = float(torch.randn(1, generator=g))
            hi = lo + float(torch.rand(1, generator=g)) * 10 + 0.1
            k = torch.randint(0, bins + 1, (n,), generator=g).float()  # values on / next to the bin edges
            x = (lo + (hi - lo) * k / bins + torch.randint(-2, 3, (n,), generator=g).float() * 1e-7 * max(abs(lo), abs(hi))).float()
        else:
            x = torch.relu(torch.randn(n, generator=g) * float(10 ** torch.empty(1).uniform_(-3, 2, generator=g)))
            lo, hi = float(x.min()), float(x.max()) * (1.0 if trial % 4 else 1.3)
        want = torch.histc(x, bins, min=lo, max=hi).numpy().astype(np.int64)
        assert np.array_equal(IO.histc(x.numpy(), bins, lo, hi), want), (trial, bins)
    same = torch.full((100,), 2.5)
    assert np.array_equal(IO.histc(same.numpy(), 16, 2.5, 2.5), torch.histc(same, 16, min=2.5, max=2.5).numpy().astype(np.int64))


def _sandwich_np(sp):
    out = {}
    for k, v in sp.items():
        out[k] = {kk: (vv.detach().cpu().numpy() if torch.is_tensor(vv) else vv) for kk, vv in v.items()}
    return out


def test_sandwich_restatement_and_product_calibration_vs_torch(fp32_net):
    """Custom variant as intended: product calibration == oracle calibration, and the integer restatement (int8 layers +
    monotone boundary tables, pool on the quantized values) == the converted torch model, layer by layer."""
    from convnet_quantization_b200 import ptq
    q = TO.build_sandwich_oracle(fp32_net, synth.calibration_batches())
    sp = ptq.calibrate_sandwich(fp32_net, synth.calibration_batches())
    act = TO.sandwich_activation_qparams(q)
    for n in TO.SANDWICH_LAYERS:
        assert act[n] == (sp[n]["in_scale"], sp[n]["in_zp"], sp[n]["out_scale"], sp[n]["out_zp"]), n
        assert np.array_equal(ptq.sandwich_boundary_lut(0.043, 70, 0.0185, 0).numpy(), IO.sandwich_lut(0.043, 70, 0.0185, 0))
    for seed, gain in ((5, 1.0), (6, 3.0)):
        x = synth.images_f32(6, seed) * gain
        logits, taps = TO.run_sandwich_oracle(q, x)
        mine = {}
        out = IO.sandwich_forward(x.numpy(), _sandwich_np(sp), mine)
        for k in TO.SANDWICH_LAYERS:
            a = taps[k].numpy()
            if a.ndim == 4:
                a = a.transpose(0, 2, 3, 1)
            assert np.array_equal(a, mine[k]), k
        np.testing.assert_allclose(out, logits.numpy(), rtol=1e-4, atol=1e-4 * np.abs(logits.numpy()).max())


def test_sandwich_golden_from_reference_classes(fp32_net):
    """tests/golden/sandwich_golden.npz was produced by the REFERENCE's own CustomQuantizedSimpleConvNet, calibrated
    through its own forward (make_golden_sandwich.py); the oracle with those activation qparams reproduces it."""
    import hashlib
    import os
    g = np.load(os.path.join(os.path.dirname(__file__), "golden", "sandwich_golden.npz"))
    q = TO.build_sandwich_oracle(fp32_net, synth.calibration_batches())
    act = {n: (float(g[f"{n}_qparams"][0]), int(g[f"{n}_qparams"][1]), float(g[f"{n}_qparams"][2]), int(g[f"{n}_qparams"][3]))
           for n in TO.SANDWICH_LAYERS}
    TO.override_sandwich_qparams(q, act)
    x = synth.normalize(torch.from_numpy(g["x_u8"])).contiguous()
    logits, taps = TO.run_sandwich_oracle(q, x)
    for n in TO.SANDWICH_LAYERS:
        a = taps[n].numpy()
        if a.ndim == 4:
            a = a.transpose(0, 2, 3, 1)
        assert hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest() == str(g[f"sandwich_{n}_sha"]), n
    np.testing.assert_allclose(logits.numpy(), g["sandwich_logits"], rtol=1e-5, atol=1e-5)


def test_dynamic_activation_quantisation_is_a_fused_multiply_add():
    """Which rounding does fbgemm use for the activations of quantized::linear_dynamic?  An identity-like weight matrix
    exposes every quantised activation in the output; over 2 M elements the fma form must reproduce ALL of them."""
    torch.backends.quantized.engine = "fbgemm"
    k = 256
    lin = torch.nn.Linear(k, k, bias=False)
    with torch.no_grad():
        lin.weight.copy_(torch.eye(k))
    qlin = torch.ao.quantization.quantize_dynamic(torch.nn.Sequential(lin), {torch.nn.Linear}, dtype=torch.qint8)[0]
    w = qlin.weight()
    assert int(w.int_repr().diagonal().min()) == 127
    g = torch.Generator().manual_seed(0)
    mism = {"fma": 0, "mul_then_add_int": 0}
    for it in range(4):
        x = torch.randn(2048, k, generator=g) * (0.3 + 0.4 * it) + 0.3 * it
        y = qlin(x).numpy()
        xn = x.numpy()
        s_x, zp = IO.dynamic_qparams(xn.min(), xn.max())
        inv = np.float32(1.0) / s_x
        xq = np.rint(y.astype(np.float64) / (np.float64(np.float32(s_x * np.float32(w.q_scale()))) * 127)).astype(np.int64) + zp
        fma = np.clip(np.rint((xn.astype(np.float64) * np.float64(inv) + zp).astype(np.float32)).astype(np.int64), 0, 255)
        mul = np.clip(np.rint(xn * inv).astype(np.int64) + zp, 0, 255)
        mism["fma"] += int((fma != xq).sum())
        mism["mul_then_add_int"] += int((mul != xq).sum())
        got = IO.linear_dynamic(xn, w.int_repr().numpy(), w.q_scale(), np.zeros(k, np.float32))
        assert np.array_equal(got, y)
    assert mism["fma"] == 0 and mism["mul_then_add_int"] > 0, mism
