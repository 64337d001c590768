"""The reference's OWN drivers, unmodified, against this package's models (SURVEY 8a row a8 / VERDICT r1 item 6).

``utils/inference_benchmark.py`` and ``utils/model_evaluator.py`` are imported from ``/root/reference`` (build
container) or from ``baseline/_ref`` (the GPU box: ``__graft_entry__.build()`` stages the two files there, git-ignored,
so they travel with the gpurun snapshot without entering the history).  No line of them is edited or monkey-patched.
"""
import importlib
import os
import sys

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _reference_root():
    for cand in ("/root/reference", os.path.join(ROOT, "baseline", "_ref")):
        if os.path.isfile(os.path.join(cand, "utils", "inference_benchmark.py")):
            return cand
    return None


@pytest.fixture(scope="module")
def ref_drivers():
    root = _reference_root()
    if root is None:
        pytest.skip("reference drivers not available (neither /root/reference nor baseline/_ref)")
    saved_path, saved_utils = list(sys.path), {k: v for k, v in sys.modules.items() if k == "utils" or k.startswith("utils.")}
    for k in saved_utils:
        del sys.modules[k]
    sys.path.insert(0, root)
    try:
        bench = importlib.import_module("utils.inference_benchmark")
        evalr = importlib.import_module("utils.model_evaluator")
        assert os.path.realpath(bench.__file__).startswith(os.path.realpath(root))
        yield bench.InferenceBenchmark, evalr.ModelEvaluator
    finally:
        sys.path[:] = saved_path
        for k in [k for k in sys.modules if k == "utils" or k.startswith("utils.")]:
            del sys.modules[k]
        sys.modules.update(saved_utils)


@pytest.fixture(scope="module")
def sd():
    from convnet_quantization_b200 import synth
    return synth.make_state_dict(0)


@pytest.fixture(scope="module")
def loader(sd):
    from convnet_quantization_b200 import synth
    from convnet_quantization_b200.models.baseline_model import SimpleConvNet
    net = SimpleConvNet()
    net.load_state_dict(sd)
    return synth.SyntheticLoader(256, 64, seed=5, label_model=net)


def test_restated_drivers_agree_with_the_reference_drivers_on_cpu(ref_drivers, sd, loader, capsys):
    """``convnet_quantization_b200.drivers`` (used where the reference is absent) computes what the real thing computes."""
    from convnet_quantization_b200 import drivers
    from convnet_quantization_b200.models.baseline_model import SimpleConvNet
    RefBench, RefEval = ref_drivers
    net = SimpleConvNet()
    net.load_state_dict(sd)
    net.eval()
    ref_top1, ref_top5 = RefEval(loader).evaluate_accuracy(net)
    top1, top5 = drivers.ModelEvaluator(loader).evaluate_accuracy(net, verbose=False)
    assert (top1, top5) == (ref_top1, ref_top5) == (100.0, 100.0)  # labels are this net's own argmax
    thr = RefBench(loader, device="cpu").measure_throughput(net, batch_size=32, num_iterations=3)
    assert thr > 0
    capsys.readouterr()


@pytest.mark.gpu
def test_reference_drivers_drive_the_gpu_models_unchanged(ref_drivers, sd, loader, oracle_model, capsys):
    """evaluate_accuracy (forces ``model.cpu()`` + CPU images, ``model_evaluator.py:15-55``), measure_throughput and
    compare_models (``inference_benchmark.py:81-157``; ``device='cuda'``, wall clock without a device synchronize) on
    StaticPTQModel / DynamicPTQModel / CustomQuantizationModel of this package."""
    from convnet_quantization_b200 import _lib
    from convnet_quantization_b200.models.custom_quantization_model import CustomQuantizationModel
    from convnet_quantization_b200.models.dynamic_ptq_model import DynamicPTQModel
    from convnet_quantization_b200.models.static_ptq_model import StaticPTQModel
    from oracle import torch_oracle as TO
    RefBench, RefEval = ref_drivers
    m = StaticPTQModel()
    m.fp32_model.load_state_dict(sd)
    q = m.quantize()
    dyn = DynamicPTQModel()
    dyn.load_state_dict(sd)
    dyn.quantize()
    cus = CustomQuantizationModel(mode="sandwich")
    cus.load_state_dict(sd)
    cus.quantize()

    # accuracy driver: the static net must score exactly what the CPU oracle's logits score
    top1, top5 = RefEval(loader).evaluate_accuracy(q)
    hits1 = hits5 = n = 0
    for images, labels in loader:
        logits, _ = TO.run_static_oracle(oracle_model, images)
        top = logits.topk(5, 1).indices
        hits1 += int((top[:, 0] == labels).sum())
        hits5 += int((top == labels.view(-1, 1)).sum())
        n += labels.numel()
    assert top1 == pytest.approx(100.0 * hits1 / n) and top5 == pytest.approx(100.0 * hits5 / n)
    d1, d5 = RefEval(loader).evaluate_accuracy(dyn)   # DynamicPTQModel: plain class with eval()/cpu()/__call__
    c1, c5 = RefEval(loader).evaluate_accuracy(cus)
    assert d1 > 90.0 and c1 > 60.0 and d5 >= d1 and c5 >= c1

    # throughput driver: launches must be counted inside its timed loop, and the unsynchronised wall clock it reads
    # must cover the work (forward synchronises for CUDA inputs): compare with a device-timed run of the same loop
    lib = _lib.load()
    bench = RefBench(loader, device="cuda")
    bench.warm_up(q)
    n0 = lib.b200q_launch_count()
    thr = bench.measure_throughput(q, batch_size=32, num_iterations=50)
    assert lib.b200q_launch_count() - n0 >= 50 * 7
    x = next(iter(loader))[0][:32].cuda()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    ev0.record()
    for _ in range(50):
        q(x)
    ev1.record()
    torch.cuda.synchronize()
    device_thr = 32 * 50 / (ev0.elapsed_time(ev1) * 1e-3)
    assert thr <= 1.25 * device_thr, (thr, device_thr)  # an asynchronous return would report many times more
    res = bench.compare_models({"static int8 (B200)": q, "dynamic (B200)": dyn, "custom sandwich (B200)": cus},
                               batch_size=32, num_iterations=10)
    assert set(res) == {"static int8 (B200)", "dynamic (B200)", "custom sandwich (B200)"}
    assert all(v > 0 for v in res.values())
    capsys.readouterr()
