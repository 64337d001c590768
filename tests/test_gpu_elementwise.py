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
