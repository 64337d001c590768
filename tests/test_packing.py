"""CPU tests of the host-side pre-pack: the kernels' algebra (zero-filled taps + border-class correction, K-major
weight order, fc1 column permutation, fp32 requant constants) emulated in numpy must reproduce the oracle."""
import numpy as np
import pytest
import torch

from convnet_quantization_b200.packing import (CONV_GEOMETRY, PackedConv, PackedLinear, conv_border_corr,
                                               requant_constants)
from oracle import int_ops as IO


def _prev(qp, name):
    order = ["in", "conv1", "conv2", "conv3", "conv4", "conv5", "conv6", "fc1", "fc2"]
    p = order[order.index(name) - 1]
    return (qp["in_scale"], qp["in_zp"]) if p == "in" else (qp[p]["out_scale"], qp[p]["out_zp"])


def _emulate_tc_conv(x_u8, pc: PackedConv):
    """What igemm_tc computes: raw = sum over taps with OUT-OF-IMAGE = 0 (TMA fill) of x_q * w, K ordered
    (kh, kw, cin); then acc = raw - corr[border class]."""
    B, H, W, C = x_u8.shape
    w = pc.w.numpy().astype(np.float64).reshape(pc.cout, 9 * pc.cin)  # [Cout][K]
    xp = np.zeros((B, H + 2, W + 2, C))
    xp[:, 1:-1, 1:-1] = x_u8
    cols = [xp[:, kh:kh + H, kw:kw + W, :] for kh in range(3) for kw in range(3)]
    a = np.concatenate(cols, axis=-1).reshape(B * H * W, 9 * C)  # im2col, K = (tap, c)
    raw = (a @ w.T).reshape(B, H, W, pc.cout).astype(np.int64)
    hh = np.arange(H)
    cls = np.where(hh == 0, 0, np.where(hh == H - 1, 2, 1))
    cfg = cls[:, None] * 3 + cls[None, :]  # [H, W]
    corr = pc.corr.numpy().astype(np.int64)  # [9, Cout]
    return raw - corr[cfg][None]


@pytest.mark.parametrize("name", list(CONV_GEOMETRY))
def test_conv_pack_algebra(qparams, qparams_np, name):
    s, zp = _prev(qparams, name)
    pc = PackedConv(name, qparams[name], s, zp, "cpu")
    cin, cout, img = CONV_GEOMETRY[name]
    rng = np.random.default_rng(3)
    x = rng.integers(0, 256, (2, img, img, cin), dtype=np.uint8)
    L = qparams_np[name]
    acc_true = IO.conv3x3_acc(x, zp, L["w_int8"])
    xk = x
    if cin == 3:  # kernel input is NHWC4; pad channel value is irrelevant (weight 0) -> use garbage
        xk = np.concatenate([x, rng.integers(0, 256, (2, img, img, 1), dtype=np.uint8)], axis=-1)
    assert np.array_equal(_emulate_tc_conv(xk, pc), acc_true)
    mult, bdiv = IO.requant_params(s, L["w_scales"], L["bias"], L["out_scale"])
    assert np.array_equal(pc.mult.numpy(), mult) and np.array_equal(pc.bdiv.numpy(), bdiv)


def test_fc1_column_permutation(qparams, qparams_np):
    s, zp = _prev(qparams, "fc1")
    pl = PackedLinear("fc1", qparams["fc1"], s, zp, "cpu", relu=True, nhwc_from=(256, 4, 4))
    rng = np.random.default_rng(5)
    x = rng.integers(0, 256, (3, 4, 4, 256), dtype=np.uint8)  # NHWC as the engine stores pool3
    want = IO.linear_acc(IO.flatten_nchw(x), zp, qparams_np["fc1"]["w_int8"])
    raw = x.reshape(3, 4096).astype(np.float64) @ pl.w.numpy().astype(np.float64).T
    assert np.array_equal(raw.astype(np.int64) - pl.corr.numpy().astype(np.int64)[None], want)


def test_border_corr_classes():
    w = torch.arange(-13, 14, dtype=torch.int8).view(1, 3, 3, 3).permute(0, 3, 1, 2).contiguous()  # [1, Cin=3, 3, 3]
    corr = conv_border_corr(w, 2)
    ws = w.to(torch.int64).sum(1)[0]  # [3,3]
    assert int(corr[4, 0]) == 2 * int(ws.sum())                       # interior: all taps
    assert int(corr[0, 0]) == 2 * int(ws[1:, 1:].sum())               # top-left corner
    assert int(corr[8, 0]) == 2 * int(ws[:2, :2].sum())               # bottom-right corner
    assert int(corr[1, 0]) == 2 * int(ws[1:, :].sum())                # top edge
    assert int(corr[5, 0]) == 2 * int(ws[:, :2].sum())                # right edge


def test_requant_constants_match_numpy():
    g = torch.Generator().manual_seed(0)
    ws = torch.rand(64, generator=g, dtype=torch.float64) * 0.01 + 1e-4
    bias = torch.randn(64, generator=g)
    mult, bdiv = requant_constants(0.0407894998788833, ws, bias, 0.0422101989388465)
    m2, b2 = IO.requant_params(0.0407894998788833, ws.numpy(), bias.numpy(), 0.0422101989388465)
    assert np.array_equal(mult.numpy(), m2) and np.array_equal(bdiv.numpy(), b2)


def test_input_lut_is_the_reference_pipeline_on_every_pixel_value():
    """lut[c][v] == QuantStub(Normalize(ToTensor(v))) for all 3 x 256 inputs, via an actual image tensor."""
    import torch
    from convnet_quantization_b200 import synth
    from convnet_quantization_b200.packing import input_lut
    scale, zp = 0.0412, 59
    lut = input_lut(scale, zp)
    assert lut.shape == (3, 256) and lut.dtype == torch.uint8
    img = torch.arange(256, dtype=torch.uint8).view(1, 1, 16, 16).expand(1, 3, 16, 16).contiguous()
    q = torch.quantize_per_tensor(synth.normalize(img), scale, zp, torch.quint8).int_repr()
    for c in range(3):
        assert torch.equal(q[0, c].flatten(), lut[c])
    # monotone in the pixel value, and the explicit-mean/std form agrees with the default
    assert bool((lut[:, 1:].int() >= lut[:, :-1].int()).all())
    assert torch.equal(input_lut(scale, zp, synth.CIFAR_MEAN, synth.CIFAR_STD), lut)


def test_prepack_handles_keep_their_objects_alive_until_collected(qparams):
    """torch.ops.b200q.*_prepack return opaque int64 handle tensors (like ATen's packed-params objects); the registry
    entry lives exactly as long as the handle tensor."""
    import gc
    from convnet_quantization_b200 import _lib, ops
    from convnet_quantization_b200.packing import PackedConv, PackedLinear
    L = qparams["conv2"]
    h = torch.ops.b200q.conv_prepack(L["w_int8"], L["w_scales"], L["bias"], qparams["conv1"]["out_scale"],
                                     qparams["conv1"]["out_zp"], L["out_scale"], L["out_zp"], True, "cpu")
    assert h.dtype == torch.int64 and h.numel() == 1
    pc = ops._packed(h, PackedConv)
    assert (pc.cin, pc.cout, pc.img) == (64, 64, 32) and pc.c.rq.relu == 1
    with pytest.raises(_lib.B200QError):
        ops._packed(h, PackedLinear)  # wrong kind of handle
    key = int(h.item())
    assert key in ops._handles
    del h, pc
    gc.collect()
    assert key not in ops._handles
    with pytest.raises(_lib.B200QError):
        torch.ops.b200q.conv_prepack(torch.zeros(7, 5, 3, 3, dtype=torch.int8), torch.ones(7, dtype=torch.float64),
                                     torch.zeros(7), 0.1, 0, 0.1, 0, True, "cpu")  # no kernel for that geometry
