"""Drop-in model classes on the GPU: same API as the reference's models/, driven by reference-protocol drivers."""
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu
GOLDEN = os.path.join(os.path.dirname(__file__), "golden", "convnet_golden.npz")


@pytest.fixture(scope="module")
def want_static(golden, oracle_model):
    """Static-PTQ logits of the live torch/fbgemm oracle, calibrated on this host like the product's ``quantize()``
    (the golden file's logits belong to the golden activation scales; see conftest.golden_activation_qparams)."""
    from oracle import torch_oracle as TO
    return TO.run_static_oracle(oracle_model, _x(golden))[0].numpy()


@pytest.fixture(scope="module")
def sd():
    from convnet_quantization_b200 import synth
    return synth.make_state_dict(0)


def _x(golden):
    from convnet_quantization_b200 import synth
    return synth.normalize(torch.from_numpy(golden["x_u8"])).contiguous()


def test_static_ptq_model_cpu_and_cuda_inputs(golden, sd, want_static):
    from convnet_quantization_b200.models.static_ptq_model import StaticPTQModel
    m = StaticPTQModel()
    m.fp32_model.load_state_dict(sd)
    q = m.quantize()  # calibration_data_loader=None -> fixed synthetic calibration set
    assert q is m.quantized_model and hasattr(q, "quantized")
    x = _x(golden)
    out_cpu = q.eval().cpu()(x)                    # ModelEvaluator protocol: .cpu() then CPU images
    assert out_cpu.device.type == "cpu"
    assert np.array_equal(out_cpu.numpy(), want_static)
    q.to("cuda")
    out_gpu = q(x.cuda())                          # InferenceBenchmark(device='cuda') protocol
    assert out_gpu.is_cuda and np.array_equal(out_gpu.cpu().numpy(), want_static)
    assert 2.5 < m.get_model_size(q) < 4.5         # ~3.25 M int8 weights + scales


def test_static_ptq_model_with_calibration_loader(golden, sd, want_static):
    from convnet_quantization_b200 import synth
    from convnet_quantization_b200.models.static_ptq_model import StaticPTQModel
    loader = [(b, torch.zeros(b.shape[0], dtype=torch.long)) for b in synth.calibration_batches()]
    m = StaticPTQModel()
    m.fp32_model.load_state_dict(sd)
    q = m.quantize(loader)
    assert np.array_equal(q(_x(golden)).numpy(), want_static)


def _assert_dynamic_close(got, want):
    """Tolerance for the WHOLE dynamic-PTQ model (BASELINE north_star: 1e-3 relative on logits, identical argmax).

    The int8 part is bit-exact (test_dynamic_linears_identical_inputs, tests/test_gpu_conv.py).  What is left is the fp32
    convolution stack in front of it - cuDNN here, MKL-DNN in the reference, different summation orders, ~1e-6 relative
    apart - and the reference algorithm is discontinuous in those features: a feature that crosses a rounding boundary of
    the per-batch activation quantiser moves a logit by a whole quantisation step.  Measured over 4 096 images / 40 960
    logits (profiles/r02_dynamic_flip_stats.json, scripts/dynamic_flip_stats.py): 93.1 % of the logits within 1e-3 of
    the logit range, the worst at 1.09e-2, 0.005 % beyond 1e-2, argmax equal on 99.93 % of the images; with the CPU's
    own features fed to the GPU linears every logit is bit-identical.  The bounds below are those measurements with
    margin: every element within 2e-2 of the logit range, at least 90 % within 1e-3, identical argmax on these images."""
    scale = np.abs(want).max()
    err = np.abs(got - want)
    assert err.max() <= 2e-2 * scale, err.max()
    assert (err <= 1e-3 * scale).mean() >= 0.9
    assert np.array_equal(got.argmax(1), want.argmax(1))


def test_dynamic_ptq_model_vs_reference_class(golden, sd):
    from convnet_quantization_b200.models.dynamic_ptq_model import DynamicPTQModel
    m = DynamicPTQModel()
    m.load_state_dict(sd)
    m.quantize()
    x = _x(golden)
    got = m.eval().cpu()(x).numpy()
    _assert_dynamic_close(got, golden["ref_dynamic"])
    # dynamic activation scale is per batch tensor: per-image batches are a different (also pinned) result
    got1 = np.concatenate([m(x[i:i + 1]).numpy() for i in range(4)])
    _assert_dynamic_close(got1, golden["ref_dynamic_b1"])
    assert m.get_model_size() > 1.0


def test_dynamic_linears_identical_inputs(golden, sd):
    """The int8 part in isolation: torch's CPU DynamicQuantizedLinear layers fed OUR conv features (so both sides
    quantize the same fp32 tensor), layer by layer - bit-exact."""
    import torch.nn.functional as F
    from convnet_quantization_b200 import ops, ptq
    from convnet_quantization_b200.models.baseline_model import SimpleConvNet
    from convnet_quantization_b200.models.dynamic_ptq_model import DynamicPTQModel
    m = DynamicPTQModel()
    m.load_state_dict(sd)
    q = m.quantize()
    net = SimpleConvNet()
    net.load_state_dict(sd)
    fused = ptq.fuse_bn(net)
    torch.backends.quantized.engine = "fbgemm"
    qd = torch.ao.quantization.quantize_dynamic(fused, {torch.nn.Linear}, dtype=torch.qint8)
    feats = q.features(_x(golden).cuda())
    with torch.no_grad():
        want_feats = _x(golden)
        for i in range(1, 7):
            want_feats = F.relu(getattr(fused, f"conv{i}")(want_feats))
            if i % 2 == 0:
                want_feats = F.max_pool2d(want_feats, 2, 2)
        torch.testing.assert_close(feats.cpu(), want_feats.reshape(16, -1), rtol=1e-4, atol=1e-4)
        a_want = F.relu(qd.fc1(feats.cpu()))
        a_got = ops.linear_dynamic(feats, q.fc["fc1"], relu=True)
        assert torch.equal(a_got.cpu(), a_want)  # the int8 part is bit-exact; only the fp32 conv features are not
        y_want = qd.fc2(a_want)
        y_got = ops.linear_dynamic(a_got, q.fc["fc2"], relu=False)
        assert torch.equal(y_got.cpu(), y_want)


def test_static_as_written_equals_dynamic_on_unfused(golden, sd):
    from convnet_quantization_b200.models.static_ptq_model import StaticPTQModel
    m = StaticPTQModel(mode="as_written")
    m.fp32_model.load_state_dict(sd)
    q = m.quantize()
    got = q(_x(golden)).numpy()
    # as written == quantize_dynamic on the unfused net: compare against torch's own CPU result
    from convnet_quantization_b200.models.baseline_model import SimpleConvNet
    ref = SimpleConvNet()
    ref.load_state_dict(sd)
    torch.backends.quantized.engine = "fbgemm"
    qref = torch.ao.quantization.quantize_dynamic(ref.eval(), {torch.nn.Linear, torch.nn.Conv2d}, dtype=torch.qint8)
    with torch.no_grad():
        want = qref(_x(golden)).numpy()
    _assert_dynamic_close(got, want)


def test_fp32_and_custom_on_cuda(golden, sd):
    from convnet_quantization_b200.models.baseline_model import SimpleConvNet
    from convnet_quantization_b200.models.custom_quantization_model import CustomQuantizationModel
    x = _x(golden).cuda()
    with torch.backends.cudnn.flags(enabled=True, allow_tf32=False), torch.no_grad():
        net = SimpleConvNet()
        net.load_state_dict(sd)
        got = net.eval().cuda()(x).cpu().numpy()
        cm = CustomQuantizationModel()
        cm.load_state_dict(sd)
        cm.quantize()
        got_c = cm.eval().cuda()(x).cpu().numpy()
    for g, w in ((got, golden["ref_fp32"]), (got_c, golden["ref_custom"])):
        np.testing.assert_allclose(g, w, rtol=1e-3, atol=1e-3 * np.abs(w).max())
        assert np.array_equal(g.argmax(1), w.argmax(1))


def test_custom_int8_and_optimized_custom(golden, sd, want_static):
    from convnet_quantization_b200.models.baseline_model import SimpleConvNet
    from convnet_quantization_b200.models.custom_quantization_model import CustomQuantizationModel
    from convnet_quantization_b200.models.optimized_custom_quantization import OptimizedCustomQuantization
    cm = CustomQuantizationModel(mode="int8")
    cm.load_state_dict(sd)
    cm.quantize()
    assert np.array_equal(cm(_x(golden)).numpy(), want_static)
    net = SimpleConvNet()
    net.load_state_dict(sd)
    oq = OptimizedCustomQuantization()
    qm = oq.quantize(net)
    assert qm.quantized and qm.is_custom_quantized
    assert np.array_equal(qm(_x(golden)).numpy(), want_static)
    assert oq.get_model_size(qm) < 4.5


def test_drivers_protocol(sd):
    """The repo's drivers restate utils/inference_benchmark.py / utils/model_evaluator.py; run them on the GPU model."""
    from convnet_quantization_b200 import synth
    from convnet_quantization_b200.drivers import InferenceBenchmark, ModelEvaluator
    from convnet_quantization_b200.models.baseline_model import SimpleConvNet
    from convnet_quantization_b200.models.static_ptq_model import StaticPTQModel
    net = SimpleConvNet()
    net.load_state_dict(sd)
    loader = synth.SyntheticLoader(256, 64, seed=5, label_model=net)
    m = StaticPTQModel()
    m.fp32_model.load_state_dict(sd)
    q = m.quantize()
    top1, top5 = ModelEvaluator(loader).evaluate_accuracy(q, verbose=False)
    assert top1 > 60.0 and top5 >= top1  # labels are the fp32 argmax; int8 agrees on most images
    thr = InferenceBenchmark(loader, device="cuda").measure_throughput(q, batch_size=32, num_iterations=20, verbose=False)
    assert thr > 0


def test_custom_sandwich_vs_reference_golden(golden, sd, fp32_net):
    """the custom variant as intended (per-layer QuantStub -> int8 -> DeQuantStub, fp32 ReLU / pool,
    fc2 fp32).  Every int8 layer output bit-exact against (a) the golden taken from the REFERENCE's own wrapper class
    (tests/golden/make_golden_sandwich.py) and (b) the live converted torch model; logits within 1e-3 (fc2 is an fp32
    GEMM on both sides) with identical argmax."""
    import hashlib
    from convnet_quantization_b200 import ptq, synth
    from convnet_quantization_b200.models._gpu_modules import B200SandwichQuantizedNet
    from convnet_quantization_b200.models.custom_quantization_model import CustomQuantizationModel
    from oracle import torch_oracle as TO
    g = np.load(os.path.join(os.path.dirname(__file__), "golden", "sandwich_golden.npz"))
    x = _x(golden)
    assert np.array_equal(g["x_u8"], golden["x_u8"])
    # (a) golden activation qparams pinned (calibration is fp32 CPU work), weights from the product's own calibration
    sp = ptq.calibrate_sandwich(fp32_net, synth.calibration_batches())
    for n in TO.SANDWICH_LAYERS:
        sp[n]["in_scale"], sp[n]["in_zp"] = float(g[f"{n}_qparams"][0]), int(g[f"{n}_qparams"][1])
        sp[n]["out_scale"], sp[n]["out_zp"] = float(g[f"{n}_qparams"][2]), int(g[f"{n}_qparams"][3])
    net = B200SandwichQuantizedNet(sp, "cuda")
    logits, taps = net.forward_with_taps(x)
    for n in TO.SANDWICH_LAYERS:
        a = np.ascontiguousarray(taps[n].cpu().numpy())
        assert hashlib.sha256(a.tobytes()).hexdigest() == str(g[f"sandwich_{n}_sha"]), n
    want = g["sandwich_logits"]
    np.testing.assert_allclose(logits.cpu().numpy(), want, rtol=1e-3, atol=1e-3 * np.abs(want).max())
    assert np.array_equal(logits.cpu().numpy().argmax(1), want.argmax(1))
    fused = net(x)  # production route: pools fused into conv2/4/6, tables on the pooled tensors; CPU in -> CPU out
    assert not fused.is_cuda and torch.equal(fused, logits.cpu())
    # (b) through the model class, against the live oracle calibrated on this host; odd batch, saturating inputs
    cm = CustomQuantizationModel(mode="sandwich")
    cm.load_state_dict(sd)
    q = cm.quantize()
    assert q.is_custom_quantized and cm.is_custom_quantized
    oq = TO.build_sandwich_oracle(fp32_net, synth.calibration_batches())
    xb = synth.images_f32(37, seed=8) * 2.5
    want_l, want_t = TO.run_sandwich_oracle(oq, xb)
    got_l, got_t = q.forward_with_taps(xb)
    for n in TO.SANDWICH_LAYERS:
        w = want_t[n].numpy()
        w = w.transpose(0, 2, 3, 1) if w.ndim == 4 else w
        assert np.array_equal(got_t[n].cpu().numpy(), w), n
    torch.testing.assert_close(got_l.cpu(), want_l, rtol=1e-3, atol=1e-3 * float(want_l.abs().max()))
    assert torch.equal(cm(xb.cuda()).cpu(), got_l.cpu())
    assert tuple(q(xb[:0]).shape) == (0, 10)


def test_custom_sandwich_host_pipeline_is_chunking_invariant(sd):
    """Host input to the sandwich net goes through the chunked two-stream pipeline (several chunks + split tail).  The
    int8 layers are per-image arithmetic, so chunking changes no bit of them; fc2 is an fp32 cuBLAS GEMM that may pick
    another kernel (summation order) for another batch size, hence fp32-rounding tolerance on the logits."""
    from convnet_quantization_b200 import synth
    from convnet_quantization_b200.models.custom_quantization_model import CustomQuantizationModel
    cm = CustomQuantizationModel(mode="sandwich")
    cm.load_state_dict(sd)
    q = cm.quantize()
    x = synth.images_f32(5000, seed=13)
    want = q(x.cuda()).cpu()
    got = q(x)
    tol = dict(rtol=1e-5, atol=1e-5 * float(want.abs().max()))
    assert not got.is_cuda
    torch.testing.assert_close(got, want, **tol)
    assert torch.equal(got.argmax(1), want.argmax(1))
    assert torch.equal(q(x.pin_memory()), got)  # second call: staging buffers reused, same chunking -> same bytes
    torch.testing.assert_close(q(x[:300]), want[:300], **tol)


def test_gpu_side_calibration(sd, fp32_net, oracle_model):
    """``StaticPTQModel.quantize(loader, calibration_device='cuda')`` runs the calibration forward and
    the observers' reductions on the GPU.  The observers are exact (test_histogram_observer_on_gpu_equals_cpu_observer);
    what differs from host calibration is the fp32 conv arithmetic that feeds them (cuDNN vs MKL-DNN summation order),
    so the resulting activation scales agree to ~1e-4 relative, the weights are identical, and the model built from
    them is bit-exact against the torch oracle GIVEN THOSE qparams."""
    from convnet_quantization_b200 import synth
    from convnet_quantization_b200.models.static_ptq_model import StaticPTQModel
    from oracle import torch_oracle as TO
    m = StaticPTQModel()
    m.fp32_model.load_state_dict(sd)
    loader = [(b, torch.zeros(b.shape[0], dtype=torch.long)) for b in synth.calibration_batches()]
    q = m.quantize(loader, calibration_device="cuda")
    host = TO.extract_qparams(oracle_model)
    qp = m.qparams
    assert abs(qp["in_scale"] / host["in_scale"] - 1) < 1e-6 and qp["in_zp"] == host["in_zp"]  # input: same tensor on both
    for n in ("conv1", "conv2", "conv3", "conv4", "conv5", "conv6", "fc1", "fc2"):
        assert torch.equal(qp[n]["w_int8"], host[n]["w_int8"]) and torch.equal(qp[n]["w_scales"], host[n]["w_scales"])
        assert abs(qp[n]["out_scale"] / host[n]["out_scale"] - 1) < 5e-3, (n, qp[n]["out_scale"], host[n]["out_scale"])
        assert abs(qp[n]["out_zp"] - host[n]["out_zp"]) <= 1, n
    import copy
    oq = copy.deepcopy(oracle_model)
    act = {"in": (qp["in_scale"], qp["in_zp"])}
    act.update({n: (qp[n]["out_scale"], qp[n]["out_zp"]) for n in ("conv1", "conv2", "conv3", "conv4", "conv5", "conv6", "fc1", "fc2")})
    TO.override_activation_qparams(oq, act)
    x = synth.images_f32(64, seed=12)
    assert torch.equal(q(x), TO.run_static_oracle(oq, x)[0])


def test_forward_synchronises_and_int_device(sd):
    """The drop-in modules finish the work before returning for CUDA inputs (the reference's benchmark reads the wall
    clock right after ``model(data)``); ``sync_on_forward = False`` opts out.  ``model.to(0)`` is ``model.to('cuda:0')``."""
    from convnet_quantization_b200 import synth
    from convnet_quantization_b200.models.static_ptq_model import StaticPTQModel
    m = StaticPTQModel()
    m.fp32_model.load_state_dict(sd)
    q = m.quantize()
    x = synth.images_f32(4096, seed=1).cuda()
    q(x)
    torch.cuda.synchronize()
    q(x)
    assert torch.cuda.current_stream().query(), "forward returned with work still queued"
    q.sync_on_forward = False
    q(x)
    pending = not torch.cuda.current_stream().query()
    torch.cuda.synchronize()
    assert pending, "opt-out must leave the forward asynchronous"
    q.sync_on_forward = True
    assert q.to(0) is q and q.engine_device == torch.device("cuda", 0)
    assert q.cuda(0) is q and q.cpu() is q and q.to("cpu") is q and q.to(torch.float32) is q
    if torch.cuda.device_count() > 1:
        q.to(1)
        assert q.engine_device == torch.device("cuda", 1)
        q.to(0)


def test_dynamic_net_empty_batch(sd):
    from convnet_quantization_b200.models.dynamic_ptq_model import DynamicPTQModel
    m = DynamicPTQModel()
    m.load_state_dict(sd)
    m.quantize()
    assert tuple(m(torch.empty(0, 3, 32, 32)).shape) == (0, 10)
    assert tuple(m(torch.empty(0, 3, 32, 32, device="cuda")).shape) == (0, 10)
