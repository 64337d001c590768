"""Whole-network GPU parity: b200q_static_forward vs the live torch/fbgemm CPU oracle and the golden vectors —
bit-exact uint8 activations at every layer and bit-exact fp32 logits."""
import hashlib
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu
GOLDEN = os.path.join(os.path.dirname(__file__), "golden", "convnet_golden.npz")


@pytest.fixture(scope="module")
def engine(qparams):
    from convnet_quantization_b200.engine import StaticEngine
    return StaticEngine(qparams, "cuda")


def _oracle_taps_nhwc(taps):
    out = {}
    for k, v in taps.items():
        a = v.numpy()
        out[k] = a.transpose(0, 2, 3, 1) if a.ndim == 4 else a
    return out


@pytest.mark.parametrize("b,seed,gain", [(1, 1, 1.0), (2, 2, 1.0), (7, 3, 1.0), (64, 4, 1.0), (33, 5, 3.0)])
def test_static_forward_taps_bit_exact(engine, oracle_model, b, seed, gain):
    from convnet_quantization_b200 import synth
    from oracle import torch_oracle as TO
    x = synth.images_f32(b, seed) * gain  # gain 3: out-of-calibration, saturating activations
    want_logits, want = TO.run_static_oracle(oracle_model, x)
    want = _oracle_taps_nhwc(want)
    logits, taps = engine.forward(x.cuda(), taps=True)
    torch.cuda.synchronize()
    for k in TO.LAYER_ORDER:
        got = taps[k].cpu().numpy()
        if k == "quant":
            got = got[..., :3]
        bad = got != want[k]
        assert not bad.any(), f"{k}: {int(bad.sum())}/{bad.size} mismatches (b={b})"
    assert torch.equal(logits.cpu(), want_logits)
    # the no-taps path (fused quantize+conv1, no copies) must give the same logits
    assert torch.equal(engine.forward(x.cuda()).cpu(), want_logits)


def test_static_forward_matches_golden(golden_qparams):
    from convnet_quantization_b200 import synth
    from convnet_quantization_b200.engine import StaticEngine
    engine = StaticEngine(golden_qparams, "cuda")
    g = np.load(GOLDEN)
    x = synth.normalize(torch.from_numpy(g["x_u8"])).contiguous()
    logits, taps = engine.forward(x.cuda(), taps=True)
    assert np.array_equal(logits.cpu().numpy(), g["static_logits"])
    for k, t in taps.items():
        a = t.cpu().numpy()
        if k == "quant":
            a = np.ascontiguousarray(a[..., :3])
        assert hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest() == str(g[f"static_{k}_sha"]), k


def test_static_forward_large_batch_properties(engine, oracle_model):
    """Full-size run (8192 images): per-image independence (batch composition must not matter) and spot parity."""
    from convnet_quantization_b200 import synth
    from oracle import torch_oracle as TO
    n = 8192
    x = synth.images_f32(n, seed=9).cuda()
    big = engine.forward(x)
    idx = torch.tensor([0, 1, 127, 128, 4095, 4096, n - 2, n - 1])
    small = engine.forward(x[idx.cuda()].contiguous())
    assert torch.equal(big[idx.cuda()], small)
    perm = torch.randperm(n, generator=torch.Generator().manual_seed(0)).cuda()
    assert torch.equal(engine.forward(x[perm].contiguous()), big[perm])
    want, _ = TO.run_static_oracle(oracle_model, x[:256].cpu())
    assert torch.equal(big[:256].cpu(), want)


def test_empty_batch(engine):
    out = engine.forward(torch.empty(0, 3, 32, 32, device="cuda"))
    assert tuple(out.shape) == (0, 10)


@pytest.mark.parametrize("b", [1, 5, 300])
def test_uint8_data_path_is_bit_identical(engine, oracle_model, b):
    """Raw uint8 NHWC pixels through the look-up-table quantiser == the fp32 route (ToTensor + Normalize on the CPU, then
    QuantStub), logits bit for bit."""
    from convnet_quantization_b200 import synth
    g = torch.Generator().manual_seed(40 + b)
    pix = torch.randint(0, 256, (b, 32, 32, 3), dtype=torch.uint8, generator=g)
    pix[0, :4] = 0
    pix[0, 4:8] = 255
    x = synth.normalize(pix.permute(0, 3, 1, 2)).contiguous()
    want = engine.forward(x.cuda())
    got = engine.forward_u8(pix.cuda())
    torch.cuda.synchronize()
    assert torch.equal(got, want)
    from oracle import torch_oracle as TO
    assert torch.equal(got.cpu(), TO.run_static_oracle(oracle_model, x)[0])  # and directly against the CPU oracle


def test_uint8_data_path_host_pipeline(golden_qparams):
    from convnet_quantization_b200 import synth
    from convnet_quantization_b200.models._gpu_modules import B200StaticQuantizedNet
    net = B200StaticQuantizedNet(golden_qparams, "cuda")
    g = torch.Generator().manual_seed(9)
    pix = torch.randint(0, 256, (5000, 32, 32, 3), dtype=torch.uint8, generator=g)
    want = net(synth.normalize(pix.permute(0, 3, 1, 2)).contiguous())
    got = net.forward_uint8(pix)
    assert not got.is_cuda and torch.equal(got, want)
    assert torch.equal(net.forward_uint8(pix.cuda()).cpu(), want)


def test_second_device_in_the_same_process(golden_qparams):
    """One process, two GPUs (``model.to('cuda:1')`` in the drop-in API): per-device kernel attributes (opt-in shared
    memory size) and workspaces must follow the device; logits are bit-identical on both."""
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs in one process")
    from convnet_quantization_b200 import synth
    from convnet_quantization_b200.engine import StaticEngine
    x = synth.images_f32(200, seed=21)
    e0 = StaticEngine(golden_qparams, "cuda:0")
    y0 = e0.forward(x.to("cuda:0")).cpu()
    e1 = StaticEngine(golden_qparams, "cuda:1")
    with torch.cuda.device(1):
        y1 = e1.forward(x.to("cuda:1")).cpu()
    y0b = e0.forward(x.to("cuda:0")).cpu()
    assert torch.equal(y0, y1) and torch.equal(y0, y0b)


def _oracle_sample(n: int, count: int) -> torch.Tensor:
    """``count`` image indices of a batch of ``n``: first and last 128 (tile / band / chunk boundaries) + a stride."""
    idx = torch.cat([torch.arange(min(128, n)), torch.arange(max(n - 128, 0), n), torch.arange(0, n, max(1, n // count))])
    return torch.unique(idx)


@pytest.mark.parametrize("n", [16384, 65536])
def test_benchmarked_batches_are_bit_exact_vs_oracle(engine, oracle_model, qparams, n):
    """BASELINE config 2 at the sizes bench.py times (16 384 per step; 65 536 = top of the sweep): >= 1 024 images of the
    batch - first, last, strided - against the fbgemm CPU oracle, through engine.forward AND through the drop-in module
    fed HOST tensors (chunked two-stream pipeline), plus whole-batch consistency of the two routes."""
    from convnet_quantization_b200 import synth
    from convnet_quantization_b200.models._gpu_modules import B200StaticQuantizedNet
    from oracle import torch_oracle as TO
    x = synth.images_f32(n, seed=1000 + n)
    idx = _oracle_sample(n, 1024)
    assert idx.numel() >= 1024
    want, _ = TO.run_static_oracle(oracle_model, x[idx])
    got = engine.forward(x.cuda())
    torch.cuda.synchronize()
    assert tuple(got.shape) == (n, 10)
    assert torch.equal(got.cpu()[idx], want)
    qmodel = B200StaticQuantizedNet(qparams, "cuda")
    host = qmodel(x)  # CPU input -> CPU logits
    assert not host.is_cuda and torch.equal(host, got.cpu())


@pytest.mark.parametrize("b", [1, 2, 7, 32, 128, 500])
@pytest.mark.parametrize("pdl", [True, False])
def test_graph_executor_is_bit_exact(qparams, oracle_model, b, pdl):
    """Whole-network executor: CUDA-graph replay with / without programmatic dependent launch ==
    eager launches == CPU oracle, across replays with DIFFERENT contents in the captured buffer (stale data from an
    overlapped predecessor kernel would show up here) and when replays are issued back to back."""
    from convnet_quantization_b200 import synth
    from convnet_quantization_b200.engine import StaticEngine
    from oracle import torch_oracle as TO
    eng = StaticEngine(qparams, "cuda", pdl=pdl)
    buf = torch.empty(b, 3, 32, 32, device="cuda")
    outs = []
    for rep in range(4):
        x = synth.images_f32(b, seed=50 * b + rep) * (3.0 if rep == 2 else 1.0)
        buf.copy_(x)
        y = eng.forward(buf)  # rep 0: eager (first sighting), rep 1: capture + replay, rep 2..: replay
        want, _ = TO.run_static_oracle(oracle_model, x)
        assert torch.equal(y.cpu(), want), (b, rep)
        assert torch.equal(eng.forward(buf, graph=False), y)
        outs.append(y)
    assert len(eng._graphs) == 1
    out = torch.empty(b, 10, device="cuda")
    for rep in range(3):  # (batch, input, output) keyed graph: writes straight into the caller's buffer
        assert eng.forward(buf, out=out) is out
    assert torch.equal(out, outs[-1])
    ys = [eng.forward(buf) for _ in range(20)]  # back-to-back replays (each clones the static logits)
    torch.cuda.synchronize()
    assert all(torch.equal(y, outs[-1]) for y in ys)


def test_graph_executor_launch_accounting_and_cache(qparams):
    from convnet_quantization_b200 import _lib, synth
    from convnet_quantization_b200.engine import StaticEngine
    lib = _lib.load()
    eng = StaticEngine(qparams, "cuda")
    x = synth.images_f32(8, seed=1).cuda()
    n0 = lib.b200q_launch_count()
    eng.forward(x, graph=False)
    per_forward = lib.b200q_launch_count() - n0  # 7 at this batch (fused small-batch head), 8 for large batches
    assert per_forward in (7, 8)
    eng.forward(x)
    eng.forward(x)
    n0 = lib.b200q_launch_count()
    for _ in range(5):
        eng.forward(x)
    assert lib.b200q_launch_count() - n0 == 5 * per_forward  # replays are counted like the launches they stand for
    for i in range(StaticEngine.GRAPH_CACHE + 3):  # more distinct buffers than the cache holds: LRU, no growth
        xi = synth.images_f32(3, seed=i).cuda()
        eng.forward(xi)
        eng.forward(xi)
    assert len(eng._graphs) == StaticEngine.GRAPH_CACHE
    big = synth.images_f32(StaticEngine.GRAPH_MAX_BATCH + 1, seed=2).cuda()
    eng.forward(big)
    eng.forward(big)
    assert all(k[0] <= StaticEngine.GRAPH_MAX_BATCH for k in eng._graphs)
    eng.release()
    assert not eng._graphs and not eng._ws


def test_engine_rejects_foreign_tensors(qparams):
    """wrong device / dtype / shape / stride must raise, not fault."""
    from convnet_quantization_b200 import _lib
    from convnet_quantization_b200.engine import StaticEngine
    eng = StaticEngine(qparams, "cuda")
    x = torch.zeros(4, 3, 32, 32, device="cuda")
    with pytest.raises(_lib.B200QError):
        eng.forward(torch.zeros(4, 3, 32, 32))
    with pytest.raises(_lib.B200QError):
        eng.forward(x, out=torch.zeros(4, 10))
    with pytest.raises(_lib.B200QError):
        eng.forward(x, out=torch.zeros(4, 10, device="cuda", dtype=torch.float64))
    with pytest.raises(_lib.B200QError):
        eng.forward(x, out=torch.zeros(5, 10, device="cuda"))
    with pytest.raises(_lib.B200QError):
        eng.forward(x, out=torch.zeros(4, 20, device="cuda")[:, ::2])
    with pytest.raises(_lib.B200QError):
        eng.forward_u8(torch.zeros(4, 32, 32, 3, device="cuda"))  # fp32 where uint8 is expected
    if torch.cuda.device_count() > 1:
        with pytest.raises(_lib.B200QError):
            eng.forward(torch.zeros(4, 3, 32, 32, device="cuda:1"))


@pytest.mark.parametrize("b", [18, 19, 32, 33, 37, 38, 74, 75, 148, 149, 443, 444])
def test_kernel_selection_thresholds_are_bit_exact(engine, oracle_model, b):
    """Every batch size at which b200q_conv3x3_tc / b200q_linear_tc / the forward switch kernels (small-batch N-tile-64
    tiles <-> band-resident kernels per layer: 37|38, 74|75, 148|149; one <-> three images per conv3 band: 443|444; fused
    head <-> two-kernel head: 32|33; fc1 tile shape) - logits vs the CPU oracle, eager and as a CUDA graph."""
    from convnet_quantization_b200 import synth
    from oracle import torch_oracle as TO
    x = synth.images_f32(b, seed=900 + b)
    want, _ = TO.run_static_oracle(oracle_model, x)
    xd = x.cuda()
    assert torch.equal(engine.forward(xd, graph=False).cpu(), want)
    assert torch.equal(engine.forward(xd, graph=True).cpu(), want)
