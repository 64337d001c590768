"""B200-native quantized ``SimpleConvNet`` forward (drop-in for his0si/ConvNet-Quantization's model classes).

Importing the package registers the operator layer as ``torch.ops.b200q.*`` (``ops.register_torch_ops``); the CUDA
library itself (``libb200q.so``) is loaded on first use and there is no CPU fallback."""
from . import ops as _ops

_ops.register_torch_ops()
