"""One static-PTQ forward at the bench batch, for ncu: the 8 kernels of the LAST forward are the ones to capture
(`ncu -k regex:'conv1_tc|conv_halo|igemm_tc|linear_simt' -s 16 -c 8 python scripts/prof_net.py`)."""
import argparse, os, sys, warnings
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
warnings.filterwarnings("ignore")
import torch
from convnet_quantization_b200 import synth
from convnet_quantization_b200.models.static_ptq_model import StaticPTQModel

ap = argparse.ArgumentParser()
ap.add_argument("--batch", type=int, default=16384)
ap.add_argument("--forwards", type=int, default=3)
a = ap.parse_args()
dev = torch.device("cuda", 0)
model = StaticPTQModel(device=dev)
model.fp32_model.load_state_dict(synth.make_state_dict(0))
engine = model.quantize().engine
g = torch.Generator(device=dev).manual_seed(0)
x = synth.normalize(torch.randint(0, 256, (a.batch, 3, 32, 32), dtype=torch.uint8, device=dev, generator=g)).contiguous()
for _ in range(a.forwards):
    y = engine.forward(x)
torch.cuda.synchronize()
print("forwards done", tuple(y.shape))
