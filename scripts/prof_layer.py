"""Run one layer of the static-PTQ net in isolation (for ncu): python scripts/prof_layer.py conv2 --pool --batch 8192"""
import argparse, os, sys, warnings
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
warnings.filterwarnings("ignore")
import torch
from convnet_quantization_b200 import ops, ptq, synth
from convnet_quantization_b200.models.baseline_model import SimpleConvNet
from convnet_quantization_b200.packing import PackedStaticNet

ap = argparse.ArgumentParser()
ap.add_argument("layer")
ap.add_argument("--pool", action="store_true")
ap.add_argument("--batch", type=int, default=8192)
ap.add_argument("--iters", type=int, default=5)
a = ap.parse_args()
net = SimpleConvNet(); net.load_state_dict(synth.make_state_dict(0))
packed = PackedStaticNet(ptq.calibrate_static(net.eval(), synth.calibration_batches()), "cuda")
g = torch.Generator(device="cuda").manual_seed(0)
if a.layer == "conv1":
    x = synth.normalize(torch.randint(0, 256, (a.batch, 3, 32, 32), dtype=torch.uint8, device="cuda", generator=g)).contiguous()
    fn = lambda: ops.quantize_conv2d_first(x, packed.in_scale, packed.convs[0])
elif a.layer == "fc1":
    x = torch.randint(0, 256, (a.batch, 4096), dtype=torch.uint8, device="cuda", generator=g)
    fn = lambda: ops.linear_q(x, packed.fc1)
else:
    # realistic activations: the layer's true input from a forward of the whole net (taps), tiled up to the batch
    from convnet_quantization_b200.engine import StaticEngine
    idx = int(a.layer[-1])
    pc = packed.convs[idx - 1]
    src = {2: "conv1", 3: "pool1", 4: "conv3", 5: "pool2", 6: "conv5"}[idx]
    eng = StaticEngine(ptq.calibrate_static(net.eval(), synth.calibration_batches()), "cuda")
    nb = min(a.batch, 1024)
    _, taps = eng.forward(synth.images_f32(nb, seed=5).cuda(), taps=True)
    x = taps[src].repeat((a.batch + nb - 1) // nb, 1, 1, 1)[:a.batch].contiguous()
    del eng, taps
    fn = lambda: ops.conv2d_q(x, pc, pool2x2=a.pool)
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
fn(); torch.cuda.synchronize()
e0.record()
for _ in range(a.iters):
    y = fn()
e1.record(); torch.cuda.synchronize()
print(f"{a.layer} pool={a.pool} batch={a.batch}: {e0.elapsed_time(e1) / a.iters:.4f} ms/iter, out {tuple(y.shape)}")
