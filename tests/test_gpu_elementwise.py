"""GPU parity (bit-exact) of the memory-bound ops against the numpy integer restatement (oracle/int_ops.py)."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def _ops():
    from convnet_quantization_b200 import ops
    return ops


@pytest.mark.parametrize("b", [1, 3, 64])
def test_quantize_nchw_to_nhwc(b):
    from oracle import int_ops as IO
    ops = _ops()
    g = torch.Generator().manual_seed(b)
    x = torch.randn(b, 3, 32, 32, generator=g) * 2.0
    x.view(-1)[:7] = torch.tensor([0.5, 1.5, 2.5, -0.5, -1.5, 1e6, -1e6]) * 0.04  # exact ties + saturation
    scale, zp = 0.04, 60
    want = IO.quantize_per_tensor(x.numpy().transpose(0, 2, 3, 1), scale, zp)
    got4 = ops.quantize_per_tensor(x.cuda(), scale, zp, c_pad=4).cpu().numpy()
    assert np.array_equal(got4[..., :3], want)
    assert (got4[..., 3] == zp).all()
    got3 = ops.quantize_per_tensor(x.cuda(), scale, zp).cpu().numpy()  # generic path, c_pad = 3
    assert np.array_equal(got3, want)


@pytest.mark.parametrize("n", [1, 15, 16, 33, 4096 * 7 + 5])
def test_quantize_flat_and_dequantize(n):
    from oracle import int_ops as IO
    ops = _ops()
    g = torch.Generator().manual_seed(n)
    x = torch.randn(n, generator=g) * 3
    q = ops.quantize_flat(x.cuda(), 0.0371, 17)
    want = IO.quantize_per_tensor(x.numpy(), 0.0371, 17)
    assert np.array_equal(q.cpu().numpy(), want)
    back = ops.dequantize(q, 0.0371, 17).cpu().numpy()
    assert np.array_equal(back, IO.dequantize(want, 0.0371, 17))
    tq = torch.quantize_per_tensor(x, 0.0371, 17, torch.quint8)  # live torch op as second witness
    assert np.array_equal(q.cpu().numpy(), tq.int_repr().numpy())
    assert np.array_equal(back, tq.dequantize().numpy())


@pytest.mark.parametrize("n", [5, 16, 1000003])
def test_relu_q(n):
    from oracle import int_ops as IO
    ops = _ops()
    q = torch.randint(0, 256, (n,), dtype=torch.uint8, generator=torch.Generator().manual_seed(n))
    got = ops.relu_q(q.cuda(), 70).cpu().numpy()
    assert np.array_equal(got, IO.relu_q(q.numpy(), 70))


@pytest.mark.parametrize("shape", [(1, 2, 2, 16), (3, 32, 32, 64), (2, 16, 16, 128), (5, 8, 8, 256)])
def test_max_pool(shape):
    from oracle import int_ops as IO
    ops = _ops()
    x = torch.randint(0, 256, shape, dtype=torch.uint8, generator=torch.Generator().manual_seed(sum(shape)))
    got = ops.max_pool2d_q(x.cuda()).cpu().numpy()
    assert np.array_equal(got, IO.max_pool2x2(x.numpy()))


@pytest.mark.parametrize("n", [4, 1001, 64 * 4096, 3 * 1000 * 1000 + 3])
def test_minmax_and_dynamic_qparams(n):
    from oracle import int_ops as IO
    ops = _ops()
    x = torch.randn(n, generator=torch.Generator().manual_seed(n)) * 1.7 + 0.3
    out = ops.minmax(x.cuda()).cpu().numpy()
    mn, mx = min(float(x.min()), 0.0), max(float(x.max()), 0.0)
    assert out[0] == np.float32(mn) and out[1] == np.float32(mx)
    s, zp = IO.dynamic_qparams(mn, mx)
    assert out[2] == s and int(out[4]) == zp
    assert out[3] == np.float32(1.0) / s
    # second call on the same scratch semantics (counter self-reset) is covered by linear_dynamic tests


def test_errors_are_exceptions():
    from convnet_quantization_b200 import _lib
    ops = _ops()
    with pytest.raises(_lib.B200QError):
        ops.max_pool2d_q(torch.zeros(1, 3, 3, 16, dtype=torch.uint8, device="cuda"))  # odd h/w
    with pytest.raises(_lib.B200QError):
        ops.relu_q(torch.zeros(16, dtype=torch.uint8), 3)  # CPU tensor: no fallback


@pytest.mark.parametrize("case", ["zeros", "tiny", "tiny_pos", "tiny_neg", "neg_only", "pos_only", "mixed", "huge"])
def test_dynamic_qparams_edge_cases_vs_torch(case):
    """The device port of ChooseQuantizationParams (small-scale cut-off, one-sided and degenerate ranges)
    against torch's own binding of the ATen routine."""
    ops = _ops()
    g = torch.Generator().manual_seed(len(case))
    x = {"zeros": torch.zeros(1000), "tiny": torch.randn(1000, generator=g) * 1e-6,
         "tiny_pos": torch.rand(1000, generator=g) * 3e-4, "tiny_neg": -torch.rand(1000, generator=g) * 3e-4,
         "neg_only": -torch.rand(1000, generator=g) - 0.5, "pos_only": torch.rand(1000, generator=g) + 0.5,
         "mixed": torch.randn(1000, generator=g) * 2.0, "huge": torch.randn(1000, generator=g) * 1e30}[case]
    out = ops.minmax(x.cuda()).cpu()
    s, z = torch._choose_qparams_per_tensor(x, True)
    assert out[2].item() == np.float32(s) and int(out[4]) == z, (case, out.tolist(), s, z)
    assert out[0].item() == min(float(x.min()), 0.0) and out[1].item() == max(float(x.max()), 0.0)


@pytest.mark.parametrize("n", [1, 7, 4096, 1000003])
def test_aminmax(n):
    ops = _ops()
    x = torch.randn(n, generator=torch.Generator().manual_seed(n)) * 3 + 5.0  # strictly positive for small n: no zero extension
    out = ops.aminmax(x.cuda()).cpu()
    mn, mx = torch.aminmax(x)
    assert out[0] == mn and out[1] == mx


@pytest.mark.parametrize("n,bins", [(1, 2048), (1000, 2048), (300007, 2048), (4 * 1000 * 1000, 2048), (50000, 37), (50000, 4096)])
def test_histc_equals_torch_cpu(n, bins):
    """b200q_histc == torch.histc on the CPU (the HistogramObserver's counts), incl. values on the bin edges, half the
    tensor in one bin (post-ReLU zeros), a widened range and out-of-range values."""
    ops = _ops()
    g = torch.Generator().manual_seed(n + bins)
    x = torch.relu(torch.randn(n, generator=g) * 1.3)
    lo, hi = float(x.min()), float(x.max())
    for a, b in ((lo, hi), (lo - 0.37, hi * 1.21 + 0.1), (0.1, max(hi * 0.5, 0.2))):
        want = torch.histc(x, bins, min=a, max=b).to(torch.int64)
        got = ops.histc(x.cuda(), bins, a, b).cpu()
        assert torch.equal(got, want), (a, b, int((got - want).abs().sum()))
    k = torch.randint(0, bins + 1, (n,), generator=g).float()
    edges = (-1.5 + 4.0 * k / bins).float()
    assert torch.equal(ops.histc(edges.cuda(), bins, -1.5, 2.5).cpu(), torch.histc(edges, bins, min=-1.5, max=2.5).to(torch.int64))
    same = torch.full((max(n, 4),), 2.5)
    assert torch.equal(ops.histc(same.cuda(), bins, 2.5, 2.5).cpu(), torch.histc(same, bins, min=2.5, max=2.5).to(torch.int64))


@pytest.mark.parametrize("n", [3, 16, 100003, 8 * 1000 * 1000])
def test_lut_u8(n):
    ops = _ops()
    g = torch.Generator().manual_seed(n)
    x = torch.randint(0, 256, (n,), dtype=torch.uint8, generator=g)
    lut = torch.randint(0, 256, (256,), dtype=torch.uint8, generator=g)
    got = ops.lut_u8(x.cuda(), lut).cpu()
    assert torch.equal(got, lut[x.long()])


def test_histogram_observer_on_gpu_equals_cpu_observer():
    """the GPU-side observer must leave exactly the state (histogram, min, max) and the qparams torch's
    CPU HistogramObserver derives from the same tensors - first observation, same-range update, widened range."""
    from convnet_quantization_b200 import ptq
    g = torch.Generator().manual_seed(11)
    batches = [torch.relu(torch.randn(64, 64, 16, 16, generator=g) * s + m) for s, m in ((1.0, 0.2), (0.5, 0.1), (2.5, -0.3), (1.0, 0.0))]
    cpu = torch.ao.quantization.HistogramObserver(reduce_range=True)
    gpu = ptq.B200HistogramObserver(reduce_range=True)
    gpu = gpu.cuda()  # must stay on the host
    assert not gpu.histogram.is_cuda
    for xb in batches:
        cpu(xb)
        gpu(xb.cuda())
        assert torch.equal(cpu.histogram, gpu.histogram)
        assert cpu.min_val == gpu.min_val and cpu.max_val == gpu.max_val
    s0, z0 = cpu.calculate_qparams()
    s1, z1 = gpu.calculate_qparams()
    assert torch.equal(s0, s1) and torch.equal(z0, z1)


def test_c_abi_argument_errors_are_status_codes_not_faults():
    """Bad arguments at the C boundary return a negative status with a message (include/b200q.h conventions)."""
    import ctypes as C
    from convnet_quantization_b200 import _lib
    lib = _lib.load()
    x = torch.zeros(128, 4096, device="cuda")
    y = torch.zeros(128, 512, device="cuda")
    w = torch.zeros(512, 4096, dtype=torch.int8, device="cuda")
    ws = torch.zeros(512, dtype=torch.int32, device="cuda")
    bias = torch.zeros(512, device="cuda")
    scratch = torch.zeros(_lib.REDUCE_SCRATCH_BYTES // 4, device="cuda")
    s = torch.cuda.current_stream().cuda_stream
    args = (x.data_ptr(), y.data_ptr(), 128, 4096, 512, w.data_ptr(), ws.data_ptr(), 0.01, bias.data_ptr(), 0, scratch.data_ptr())
    assert lib.b200q_linear_dynamic(*args, _lib.REDUCE_SCRATCH_BYTES, s) == 0
    assert lib.b200q_linear_dynamic(*args, 8208, s) == -1                       # the size round 1's header documented: too small
    assert b"scratch too small" in lib.b200q_last_error()
    bad_n = list(args)
    bad_n[4] = 100
    assert lib.b200q_linear_dynamic(*bad_n, _lib.REDUCE_SCRATCH_BYTES, s) == -1  # n must be 512 or <= 16
    hist = torch.zeros(8192, dtype=torch.int64, device="cuda")
    assert lib.b200q_histc(x.data_ptr(), x.numel(), 0.0, 1.0, 8192, hist.data_ptr(), s) == -1  # bins > 4096
    assert lib.b200q_histc(x.data_ptr(), x.numel(), 1.0, 1.0, 16, hist.data_ptr(), s) == -1    # lo == hi
    assert lib.b200q_lut_u8(x.data_ptr(), y.data_ptr(), 16, None, s) == -1
    g = C.c_void_p()
    assert lib.b200q_graph_create(None, x.data_ptr(), y.data_ptr(), 4, scratch.data_ptr(), 1 << 20, 0, s, C.byref(g)) == -1
    assert lib.b200q_graph_launch(None, s) == -1
    assert lib.b200q_graph_destroy(None) == 0
    torch.cuda.synchronize()


def test_graph_capture_refuses_the_legacy_default_stream(qparams):
    import ctypes as C
    from convnet_quantization_b200 import _lib
    from convnet_quantization_b200.engine import StaticEngine
    eng = StaticEngine(qparams, "cuda")
    x = torch.zeros(4, 3, 32, 32, device="cuda")
    y = torch.zeros(4, 10, device="cuda")
    ws = torch.empty(int(eng.lib.b200q_static_workspace_bytes(4)), dtype=torch.uint8, device="cuda")
    g = C.c_void_p()
    rc = eng.lib.b200q_graph_create(eng.packed.ptr(), x.data_ptr(), y.data_ptr(), 4, ws.data_ptr(), ws.numel(), 0, None, C.byref(g))
    assert rc == -1 and b"legacy default stream" in eng.lib.b200q_last_error() and not g.value
    rc = eng.lib.b200q_graph_create(eng.packed.ptr(), x.data_ptr(), y.data_ptr(), 4, ws.data_ptr(), 100, _lib.GRAPH_PDL,
                                    torch.cuda.Stream().cuda_stream, C.byref(g))
    assert rc == -1 and b"workspace too small" in eng.lib.b200q_last_error()
    torch.cuda.synchronize()
