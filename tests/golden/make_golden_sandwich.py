"""Regenerate tests/golden/sandwich_golden.npz.  Run in the build container only (needs /root/reference):

    PYTHONDONTWRITEBYTECODE=1 python tests/golden/make_golden_sandwich.py

Golden for the custom variant *as intended* (SURVEY 8f rank 3).  Everything that can be is the REFERENCE's own code:
the reference's ``CustomQuantizedSimpleConvNet`` (``models/custom_quantization_model.py:202-261``) is instantiated on
the reference's fused ``SimpleConvNet``, given the fbgemm qconfig with the two repairs of survey probe P3 (outer
QuantStub / DeQuantStub and fc2 without qconfig), calibrated THROUGH ITS OWN ``forward`` and converted by torch.  Only
the inference pass over the converted model restates the reference forward, because its ``x.view(-1, 4096)`` (``:255``)
cannot run on the channels-last tensors a really-quantized conv produces: same statements, ``reshape``.

Content: ``x_u8`` (the 16 images of convnet_golden.npz), per layer ``<layer>_qparams`` = (in_scale, in_zp, out_scale,
out_zp), ``sandwich_logits``, ``sandwich_<layer>_sha`` = sha256 of the uint8 NHWC output of each int8 layer.
"""
import hashlib
import os
import sys
import warnings

import numpy as np
import torch
import torch.nn.functional as F

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
REF = "/root/reference"
warnings.filterwarnings("ignore")
LAYERS = ("conv1", "conv2", "conv3", "conv4", "conv5", "conv6", "fc1")


def sha(a: np.ndarray) -> str:
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def main():
    from convnet_quantization_b200 import synth
    from oracle import torch_oracle as TO

    sys.path.insert(0, REF)
    from models.baseline_model import SimpleConvNet as RefNet  # noqa: E402
    from models.custom_quantization_model import CustomQuantizedSimpleConvNet as RefCQ  # noqa: E402

    torch.backends.quantized.engine = "fbgemm"
    torch.set_num_threads(1)
    sd = synth.make_state_dict(0)
    x_u8 = synth.images_u8(16, seed=42)
    x = synth.normalize(x_u8).contiguous()
    ref = RefNet()
    ref.load_state_dict(sd)
    ref.eval()
    fused = torch.quantization.fuse_modules(ref, [[f"conv{i}", f"bn{i}"] for i in range(1, 7)] + [["fc1", "bn7"]],
                                            inplace=False)  # custom_quantization_model.py:180-190
    net = RefCQ(fused).eval()
    net.qconfig = torch.ao.quantization.get_default_qconfig("fbgemm")
    net.quant.qconfig = None
    net.dequant.qconfig = None
    net.fc2.qconfig = None
    prepared = torch.ao.quantization.prepare(net, inplace=False)
    with torch.no_grad():
        for xb in synth.calibration_batches():
            prepared(xb)  # the reference's own forward (fp32 tensors: its .view works here)
    q = torch.ao.quantization.convert(prepared, inplace=False).eval()

    taps = {}

    def run(name, t):
        sw = getattr(q, name)
        y = (sw.conv if hasattr(sw, "conv") else sw.linear)(sw.quant(t))
        taps[name] = y.int_repr()
        return F.relu(sw.dequant(y))

    with torch.no_grad():  # custom_quantization_model.py:233-261, view -> reshape
        h = q.quant(x)
        h = run("conv2", run("conv1", h))
        h = q.dropout1(q.pool1(h))
        h = run("conv4", run("conv3", h))
        h = q.dropout2(q.pool2(h))
        h = run("conv6", run("conv5", h))
        h = q.dropout3(q.pool3(h))
        h = run("fc1", h.reshape(-1, 256 * 4 * 4))
        logits = q.dequant(q.fc2(q.dropout4(h)))

    out = {"x_u8": x_u8.numpy(), "sandwich_logits": logits.numpy()}
    act = {}
    for name in LAYERS:
        sw = getattr(q, name)
        mod = sw.conv if hasattr(sw, "conv") else sw.linear
        act[name] = (float(sw.quant.scale), int(sw.quant.zero_point), float(mod.scale), int(mod.zero_point))
        out[f"{name}_qparams"] = np.array(act[name], dtype=np.float64)
        a = taps[name].numpy()
        if a.ndim == 4:
            a = a.transpose(0, 2, 3, 1)
        out[f"sandwich_{name}_sha"] = np.array(sha(a))

    # the oracle restatement (oracle/torch_oracle.py SandwichWrap) must agree with the reference-class run bit for bit
    oq = TO.build_sandwich_oracle(ref, synth.calibration_batches())
    assert TO.sandwich_activation_qparams(oq) == act, "oracle calibration differs from the reference class"
    ol, ot = TO.run_sandwich_oracle(oq, x)
    assert torch.equal(ol, logits)
    for name in LAYERS:
        assert torch.equal(ot[name], taps[name]), name

    path = os.path.join(HERE, "sandwich_golden.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path), "bytes")


if __name__ == "__main__":
    main()
