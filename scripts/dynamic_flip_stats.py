"""Dynamic-PTQ tolerance, quantified: logits of the GPU DynamicPTQModel against the CPU model the
reference's class builds (oracle/torch_oracle.build_dynamic_oracle = fuse + quantize_dynamic), over N images in
batches of 64 (dynamic quantisation is per batch tensor).  Three comparisons:
  full         GPU convs (cuDNN fp32) + GPU int8 linears     vs  CPU convs (MKL-DNN fp32) + fbgemm linears
  linears_only GPU int8 linears fed the CPU's conv features  vs  the same CPU model (isolates b200q_linear_dynamic)
  cpu_self     CPU model with 1 thread vs all threads (how much the REFERENCE moves under its own summation order)
    python scripts/dynamic_flip_stats.py [--images 4096] > profiles/r02_dynamic_flip_stats.json"""
import argparse, json, os, sys, warnings
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
warnings.filterwarnings("ignore")
import numpy as np
import torch
import torch.nn.functional as F
from convnet_quantization_b200 import ops, synth
from convnet_quantization_b200.models.baseline_model import SimpleConvNet
from convnet_quantization_b200.models.dynamic_ptq_model import DynamicPTQModel
from oracle import torch_oracle

ap = argparse.ArgumentParser()
ap.add_argument("--images", type=int, default=4096)
a = ap.parse_args()
sd = synth.make_state_dict(0)
net = SimpleConvNet()
net.load_state_dict(sd)
cpu = torch_oracle.build_dynamic_oracle(net.eval())
m = DynamicPTQModel()
m.load_state_dict(sd)
gpu = m.quantize()


def cpu_features(x):
    for i in range(1, 7):
        x = F.relu(getattr(cpu, f"conv{i}")(x))
        if i % 2 == 0:
            x = F.max_pool2d(x, 2, 2)
    return x.reshape(x.shape[0], -1)


def stats(err, scale, agree):
    err = np.concatenate(err) / scale
    return {"logits": int(err.size), "max_err_over_logit_range": float(err.max()), "frac_outside_1e-3": float((err > 1e-3).mean()),
            "frac_outside_1e-4": float((err > 1e-4).mean()), "frac_outside_1e-2": float((err > 1e-2).mean()),
            "images_with_a_logit_outside_1e-3": float(np.mean(err.reshape(-1, 10).max(1) > 1e-3)),
            "argmax_agreement": float(np.mean(np.concatenate(agree)))}


full, lin, selfc = ([], []), ([], []), ([], [])
scale = 0.0
threads = torch.get_num_threads()
with torch.no_grad():
    for b0 in range(0, a.images, 64):
        x = synth.images_f32(64, seed=5000 + b0)
        want = cpu(x)
        scale = max(scale, float(want.abs().max()))
        got = gpu(x)
        full[0].append((got - want).abs().numpy()); full[1].append((got.argmax(1) == want.argmax(1)).numpy())
        f = cpu_features(x)
        h = ops.linear_dynamic(f.cuda().contiguous(), gpu.fc["fc1"], relu=True)
        y = ops.linear_dynamic(h, gpu.fc["fc2"], relu=False).cpu()
        lin[0].append((y - want).abs().numpy()); lin[1].append((y.argmax(1) == want.argmax(1)).numpy())
        torch.set_num_threads(1)
        w1 = cpu(x)
        torch.set_num_threads(threads)
        selfc[0].append((w1 - want).abs().numpy()); selfc[1].append((w1.argmax(1) == want.argmax(1)).numpy())
out = {"what": "dynamic-PTQ logits, GPU vs the CPU model the reference builds; errors relative to the largest |logit| seen",
       "images": a.images, "batch": 64, "logit_range": scale, "cpu_threads": threads,
       "full_gpu_vs_cpu": stats(full[0], scale, full[1]),
       "gpu_linears_on_cpu_features_vs_cpu": stats(lin[0], scale, lin[1]),
       "cpu_1_thread_vs_cpu_all_threads": stats(selfc[0], scale, selfc[1])}
json.dump(out, sys.stdout, indent=1)
