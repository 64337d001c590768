"""Out-of-bounds writes: every C-ABI entry point writes ONLY the bytes its contract names.

compute-sanitizer is not available on the GPU pool, so the check is made with guard bands: each output (and the
whole-net workspace) sits inside a larger allocation pre-filled with a sentinel; after the call the bands either side
must be untouched and the payload must equal what the same entry produces into an ordinary allocation.  Ragged batch
sizes (partial M tiles, partial CTA pairs, partial warps) are where a stray store would show.
"""
import ctypes as C

import pytest
import torch

pytestmark = pytest.mark.gpu

GUARD = 8192  # bytes either side; a multiple of every alignment the kernels ask for
SENTINEL = 0xA5


class Guarded:
    def __init__(self, shape, dtype):
        self.nbytes = int(torch.empty(shape, dtype=dtype).numel() * torch.empty((), dtype=dtype).element_size())
        self.raw = torch.full((2 * GUARD + self.nbytes,), SENTINEL, dtype=torch.uint8, device="cuda")
        self.view = self.raw[GUARD:GUARD + self.nbytes].view(dtype).view(shape)

    def intact(self) -> bool:
        torch.cuda.synchronize()
        front, back = self.raw[:GUARD], self.raw[GUARD + self.nbytes:]
        return bool((front == SENTINEL).all()) and bool((back == SENTINEL).all())


def _stream():
    return torch.cuda.current_stream().cuda_stream


def _u8(shape, seed):
    return torch.randint(0, 256, shape, dtype=torch.uint8, generator=torch.Generator().manual_seed(seed)).cuda()


def _conv(qparams, name):
    from convnet_quantization_b200.packing import PackedConv
    order = ["in", "conv1", "conv2", "conv3", "conv4", "conv5", "conv6"]
    prev = order[order.index(name) - 1]
    s, zp = (qparams["in_scale"], qparams["in_zp"]) if prev == "in" else (qparams[prev]["out_scale"], qparams[prev]["out_zp"])
    return PackedConv(name, qparams[name], s, zp, "cuda"), s, zp


@pytest.mark.parametrize("name,pool", [("conv2", True), ("conv2", False), ("conv3", False), ("conv4", True), ("conv5", False),
                                       ("conv6", True), ("conv6", False)])
@pytest.mark.parametrize("b", [1, 3, 37, 149, 297])
def test_conv_tc_writes_only_its_output(qparams, name, pool, b):
    from convnet_quantization_b200 import _lib, ops
    pc, _, _ = _conv(qparams, name)
    x = _u8((b, pc.img, pc.img, pc.cin), 3 * b + len(name))
    o = pc.img // 2 if pool else pc.img
    g = Guarded((b, o, o, pc.cout), torch.uint8)
    _lib.check(_lib.load().b200q_conv3x3_tc(x.data_ptr(), g.view.data_ptr(), b, pc.ptr(), int(pool), _stream()), "conv3x3_tc")
    assert g.intact(), f"{name} pool={pool} b={b}: wrote outside its output"
    assert torch.equal(g.view, ops.conv2d_q(x, pc, pool2x2=pool))


@pytest.mark.parametrize("b", [1, 5, 131, 300])
def test_first_layer_writes_only_its_output(qparams, b):
    from convnet_quantization_b200 import _lib, ops, synth
    pc, s, zp = _conv(qparams, "conv1")
    x = (synth.images_f32(b, seed=b) * 1.3).cuda().contiguous()
    g = Guarded((b, 32, 32, 64), torch.uint8)
    _lib.check(_lib.load().b200q_quantize_conv3x3_first(x.data_ptr(), g.view.data_ptr(), b, ops._inv_scale(s), pc.ptr(),
                                                       _stream()), "quantize_conv3x3_first")
    assert g.intact()
    assert torch.equal(g.view, ops.quantize_conv2d_first(x, s, pc))
    xq = ops.quantize_per_tensor(x, s, zp, c_pad=4)
    g2 = Guarded((b, 32, 32, 64), torch.uint8)
    _lib.check(_lib.load().b200q_conv3x3_first(xq.data_ptr(), g2.view.data_ptr(), b, pc.ptr(), _stream()), "conv3x3_first")
    assert g2.intact()
    assert torch.equal(g2.view, g.view)


@pytest.mark.parametrize("b", [1, 7, 33, 129, 1000])
def test_linears_write_only_their_outputs(qparams, b):
    from convnet_quantization_b200 import _lib, ops
    from convnet_quantization_b200.packing import PackedLinear
    lib = _lib.load()
    fc1 = PackedLinear("fc1", qparams["fc1"], qparams["conv6"]["out_scale"], qparams["conv6"]["out_zp"], "cuda", relu=True,
                       nhwc_from=(256, 4, 4))
    fc2 = PackedLinear("fc2", qparams["fc2"], qparams["fc1"]["out_scale"], qparams["fc1"]["out_zp"], "cuda", relu=False)
    x = _u8((b, 4096), b)
    for fn, what in ((lib.b200q_linear_tc, "linear_tc"), (lib.b200q_linear_simt, "linear_simt")):
        g = Guarded((b, 512), torch.uint8)
        _lib.check(fn(x.data_ptr(), g.view.data_ptr(), b, fc1.ptr(), _stream()), what)
        assert g.intact(), f"{what} b={b}"
        assert torch.equal(g.view, ops.linear_q(x, fc1))
    h = ops.linear_q(x, fc1)
    g = Guarded((b, 10), torch.float32)
    _lib.check(lib.b200q_linear_dequant(h.data_ptr(), g.view.data_ptr(), b, fc2.ptr(), float(qparams["fc2"]["out_scale"]),
                                        _stream()), "linear_dequant")
    assert g.intact(), f"linear_dequant b={b}"
    assert torch.equal(g.view, ops.linear_dequant(h, fc2, qparams["fc2"]["out_scale"]))


@pytest.mark.parametrize("k,n", [(4096, 512), (512, 10), (64, 16), (128, 3)])
@pytest.mark.parametrize("b", [1, 127, 129, 300])
def test_linear_dynamic_writes_only_its_output(k, n, b):
    from convnet_quantization_b200 import _lib, ops
    gen = torch.Generator().manual_seed(k + n + b)
    w = torch.randint(-127, 128, (n, k), dtype=torch.int8, generator=gen)
    W = ops.DynamicLinearWeights(w, 0.0123, torch.randn(n, generator=gen), "cuda")
    x = (torch.randn(b, k, generator=gen) * 2).cuda()
    g = Guarded((b, n), torch.float32)
    _lib.check(_lib.load().b200q_linear_dynamic(x.data_ptr(), g.view.data_ptr(), b, k, n, W.w.data_ptr(), W.wsum.data_ptr(),
                                                W.w_scale, W.bias.data_ptr(), 1, W.scratch.data_ptr(), W.scratch.numel() * 4,
                                                _stream()), "linear_dynamic")
    assert g.intact(), f"linear_dynamic k={k} n={n} b={b}"
    assert torch.equal(g.view, ops.linear_dynamic(x, W, relu=True))


@pytest.mark.parametrize("n", [1, 31, 4097, 1_000_003])
def test_elementwise_entries_write_only_their_outputs(n):
    from convnet_quantization_b200 import _lib, ops
    lib = _lib.load()
    gen = torch.Generator().manual_seed(n)
    x = (torch.randn(n, generator=gen) * 3).cuda()
    q = _u8((n,), n)
    g = Guarded((n,), torch.uint8)
    _lib.check(lib.b200q_quantize_flat(x.data_ptr(), g.view.data_ptr(), n, 1.0 / 0.05, 128, _stream()), "quantize_flat")
    assert g.intact() and torch.equal(g.view, ops.quantize_flat(x, 0.05, 128))
    g = Guarded((n,), torch.float32)
    _lib.check(lib.b200q_dequantize(q.data_ptr(), g.view.data_ptr(), n, 0.05, 3, _stream()), "dequantize")
    assert g.intact() and torch.equal(g.view, ops.dequantize(q, 0.05, 3))
    g = Guarded((n,), torch.uint8)
    _lib.check(lib.b200q_relu_q(q.data_ptr(), g.view.data_ptr(), n, 77, _stream()), "relu_q")
    assert g.intact() and torch.equal(g.view, ops.relu_q(q, 77))
    lut = torch.randint(0, 256, (256,), dtype=torch.uint8, generator=gen)
    g = Guarded((n,), torch.uint8)
    _lib.check(lib.b200q_lut_u8(q.data_ptr(), g.view.data_ptr(), n, lut.data_ptr(), _stream()), "lut_u8")
    assert g.intact() and torch.equal(g.view, ops.lut_u8(q, lut))


@pytest.mark.parametrize("b,h,c", [(1, 2, 16), (3, 32, 64), (5, 8, 256), (7, 6, 48)])
def test_max_pool_and_layout_entries_write_only_their_outputs(b, h, c):
    from convnet_quantization_b200 import _lib, ops
    lib = _lib.load()
    x = _u8((b, h, h, c), b + h + c)
    g = Guarded((b, h // 2, h // 2, c), torch.uint8)
    _lib.check(lib.b200q_max_pool2x2_nhwc(x.data_ptr(), g.view.data_ptr(), b, h, h, c, _stream()), "max_pool2x2_nhwc")
    assert g.intact() and torch.equal(g.view, ops.max_pool2d_q(x))
    assert torch.equal(g.view.cpu(), torch.nn.functional.max_pool2d(x.cpu().permute(0, 3, 1, 2).float(), 2).permute(0, 2, 3, 1).to(torch.uint8))
    xf = torch.randn(b, 3, h, h, generator=torch.Generator().manual_seed(b)).cuda()
    g = Guarded((b, h, h, 4), torch.uint8)
    _lib.check(lib.b200q_quantize_nchw_to_nhwc(xf.data_ptr(), g.view.data_ptr(), b, 3, h, h, 4, 1.0 / 0.02, 120, _stream()),
               "quantize_nchw_to_nhwc")
    assert g.intact() and torch.equal(g.view, ops.quantize_per_tensor(xf, 0.02, 120, c_pad=4))


@pytest.mark.parametrize("b", [1, 7, 32, 33, 149, 300, 1025])
def test_whole_net_stays_inside_workspace_and_logits(qparams, b):
    """b200q_static_forward / _u8: the workspace is exactly b200q_static_workspace_bytes(b); nothing lands outside it or
    outside the b x 10 logits (batches either side of the fused-head and small-tile thresholds)."""
    from convnet_quantization_b200 import _lib, synth
    from convnet_quantization_b200.engine import StaticEngine
    lib = _lib.load()
    eng = StaticEngine(qparams, "cuda", use_graphs=False)
    x = synth.images_f32(b, seed=40 + b).cuda().contiguous()
    want = eng.forward(x, graph=False)
    ws_bytes = int(lib.b200q_static_workspace_bytes(b))
    ws = Guarded((ws_bytes,), torch.uint8)
    ws.view.zero_()
    out = Guarded((b, 10), torch.float32)
    rc = lib.b200q_static_forward(eng.packed.ptr(), x.data_ptr(), out.view.data_ptr(), b, ws.view.data_ptr(), ws_bytes, None,
                                  _stream())
    _lib.check(rc, "static_forward")
    assert ws.intact(), f"b={b}: wrote outside the workspace"
    assert out.intact(), f"b={b}: wrote outside the logits"
    assert torch.equal(out.view, want)
    pix = _u8((b, 32, 32, 3), b)
    want8 = eng.forward_u8(pix)
    out8 = Guarded((b, 10), torch.float32)
    ws.view.zero_()
    rc = lib.b200q_static_forward_u8(eng.packed.ptr(), pix.data_ptr(), eng.packed.input_lut.data_ptr(), out8.view.data_ptr(), b,
                                     ws.view.data_ptr(), ws_bytes, _stream())
    _lib.check(rc, "static_forward_u8")
    assert ws.intact() and out8.intact()
    assert torch.equal(out8.view, want8)


@pytest.mark.parametrize("b", [1, 19, 64])
@pytest.mark.parametrize("pdl", [0, 1])
def test_graph_executor_stays_inside_workspace_and_logits(qparams, b, pdl):
    from convnet_quantization_b200 import _lib, synth
    from convnet_quantization_b200.engine import StaticEngine
    lib = _lib.load()
    eng = StaticEngine(qparams, "cuda", use_graphs=False)
    x = synth.images_f32(b, seed=90 + b).cuda().contiguous()
    want = eng.forward(x, graph=False)
    ws_bytes = int(lib.b200q_static_workspace_bytes(b))
    ws = Guarded((ws_bytes,), torch.uint8)
    ws.view.zero_()
    out = Guarded((b, 10), torch.float32)
    side = torch.cuda.Stream()
    side.wait_stream(torch.cuda.current_stream())
    gh = C.c_void_p()
    with torch.cuda.stream(side):
        rc = lib.b200q_graph_create(eng.packed.ptr(), x.data_ptr(), out.view.data_ptr(), b, ws.view.data_ptr(), ws_bytes,
                                    _lib.GRAPH_PDL if pdl else 0, side.cuda_stream, C.byref(gh))
        _lib.check(rc, "graph_create")
        for _ in range(3):
            _lib.check(lib.b200q_graph_launch(gh, side.cuda_stream), "graph_launch")
    side.synchronize()
    try:
        assert ws.intact() and out.intact()
        assert torch.equal(out.view, want)
    finally:
        _lib.check(lib.b200q_graph_destroy(gh), "graph_destroy")
