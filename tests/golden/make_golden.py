"""Regenerate tests/golden/convnet_golden.npz.  Run in the build container only (needs /root/reference):

    PYTHONDONTWRITEBYTECODE=1 python tests/golden/make_golden.py

Golden content (all from seed-fixed synthetic weights ``synth.make_state_dict(0)`` and images
``synth.images_u8(16, seed=42)``, because the reference ships neither trained_model.pth nor CIFAR-10):

* ``x_u8``                       the 16 input images (uint8 NCHW)
* ``ref_fp32 / ref_dynamic / ref_custom``  logits of the REFERENCE's own classes imported from /root/reference
  (``models.baseline_model.SimpleConvNet``, ``models.dynamic_ptq_model.DynamicPTQModel``,
  ``models.custom_quantization_model.CustomQuantizationModel``) on CPU
* ``static_logits``, ``static_<layer>_sha`` (sha256 of each uint8 NHWC activation), ``static_<layer>_head`` (first
  256 bytes) of the torch/fbgemm static-PTQ oracle (oracle/torch_oracle.py), plus every scale / zero-point
* ``sd_sha``                      sha256 over the synthetic state_dict (detects RNG drift across torch versions)
"""
import hashlib
import os
import sys
import warnings

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
REF = "/root/reference"
warnings.filterwarnings("ignore")


def sha(a: np.ndarray) -> str:
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def state_dict_sha(sd) -> str:
    h = hashlib.sha256()
    for k in sorted(sd):
        h.update(k.encode())
        h.update(sd[k].detach().cpu().numpy().tobytes())
    return h.hexdigest()


def main():
    from convnet_quantization_b200 import synth
    from oracle import torch_oracle as TO

    sys.path.insert(0, REF)
    from models.baseline_model import SimpleConvNet as RefNet          # noqa: E402
    from models.custom_quantization_model import CustomQuantizationModel as RefCustom  # noqa: E402
    from models.dynamic_ptq_model import DynamicPTQModel as RefDynamic  # noqa: E402

    torch.set_num_threads(1)  # deterministic fp32 summation order for the float references
    sd = synth.make_state_dict(0)
    x_u8 = synth.images_u8(16, seed=42)
    x = synth.normalize(x_u8).contiguous()
    out = {"x_u8": x_u8.numpy(), "sd_sha": np.array(state_dict_sha(sd))}

    with torch.no_grad():
        ref = RefNet()
        ref.load_state_dict(sd)
        ref.eval()
        out["ref_fp32"] = ref(x).numpy()

        dyn = RefDynamic()
        dyn.load_state_dict(sd)
        dyn.quantize()
        out["ref_dynamic"] = dyn(x).numpy()
        out["ref_dynamic_b1"] = torch.cat([dyn(x[i:i + 1]) for i in range(4)]).numpy()  # per-image batches

        cus = RefCustom()
        cus.load_state_dict(sd)
        cus.quantize()
        out["ref_custom"] = cus(x).numpy()

    q = TO.build_static_oracle(ref, synth.calibration_batches())
    logits, taps = TO.run_static_oracle(q, x)
    out["static_logits"] = logits.numpy()
    for k, v in taps.items():
        a = v.numpy()
        if a.ndim == 4:
            a = a.transpose(0, 2, 3, 1)
        out[f"static_{k}_sha"] = np.array(sha(a))
        out[f"static_{k}_head"] = np.ascontiguousarray(a).reshape(-1)[:256].copy()
    qp = TO.extract_qparams(q)
    out["in_scale"], out["in_zp"] = np.float64(qp["in_scale"]), np.int64(qp["in_zp"])
    for name in ("conv1", "conv2", "conv3", "conv4", "conv5", "conv6", "fc1", "fc2"):
        out[f"{name}_out_scale"] = np.float64(qp[name]["out_scale"])
        out[f"{name}_out_zp"] = np.int64(qp[name]["out_zp"])
        out[f"{name}_w_sha"] = np.array(sha(qp[name]["w_int8"].numpy()))
        out[f"{name}_w_scales"] = qp[name]["w_scales"].numpy()
    path = os.path.join(HERE, "convnet_golden.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path), "bytes")


if __name__ == "__main__":
    main()
