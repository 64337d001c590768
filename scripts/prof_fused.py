"""Run the fused conv1+conv2 kernel in isolation (for ncu)."""
import os, sys, warnings
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
warnings.filterwarnings("ignore")
import torch
from convnet_quantization_b200 import ops, ptq, synth
from convnet_quantization_b200.models.baseline_model import SimpleConvNet
from convnet_quantization_b200.packing import PackedStaticNet
b = int(sys.argv[1]) if len(sys.argv) > 1 else 8192
net = SimpleConvNet(); net.load_state_dict(synth.make_state_dict(0))
packed = PackedStaticNet(ptq.calibrate_static(net.eval(), synth.calibration_batches()), "cuda")
g = torch.Generator(device="cuda").manual_seed(0)
x = synth.normalize(torch.randint(0, 256, (b, 3, 32, 32), dtype=torch.uint8, device="cuda", generator=g)).contiguous()
for _ in range(3):
    y = ops.quantize_conv2d_conv2d_pool(x, packed.in_scale, packed.convs[0], packed.convs[1])
torch.cuda.synchronize()
print("done", tuple(y.shape))
