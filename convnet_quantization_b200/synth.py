"""Synthetic stand-ins for the blobs the reference needs but does not ship.

``trained_model.pth`` is missing from the reference (``.MISSING_LARGE_BLOBS``)
and CIFAR-10 cannot be downloaded, so weights and images are synthesised with
fixed seeds:

* checkpoint: seeded ``SimpleConvNet`` init (``models/baseline_model.py:45-56``)
  plus randomised BatchNorm statistics so that BN folding matters, stored in
  the dict schema ``main.py:22-26`` reads (``model_state_dict``, ``best_accuracy``);
* images: uniform uint8 pixels, ``/255`` then the CIFAR-10 normalisation of
  ``utils/dataset_manager.py:41-44``; fp32 NCHW ``[B,3,32,32]``.
"""
from __future__ import annotations

import torch

from .models.baseline_model import SimpleConvNet, IMAGE_SHAPE

CIFAR_MEAN = (0.4914, 0.4822, 0.4465)
CIFAR_STD = (0.2023, 0.1994, 0.2010)
CALIB_SEED = 1234
CALIB_BATCHES = 4
CALIB_BATCH = 64


def make_state_dict(seed: int = 0) -> dict:
    """Seeded fp32 weights with non-trivial BN running stats / affine."""
    with torch.random.fork_rng(devices=[]):
        torch.manual_seed(seed)
        net = SimpleConvNet()
        g = torch.Generator().manual_seed(seed + 7919)
        for m in net.modules():
            if isinstance(m, (torch.nn.BatchNorm2d, torch.nn.BatchNorm1d)):
                n = m.num_features
                m.running_mean.copy_(torch.randn(n, generator=g) * 0.1)
                m.running_var.copy_(torch.rand(n, generator=g) + 0.5)
                m.weight.data.copy_(torch.rand(n, generator=g) + 0.5)
                m.bias.data.copy_(torch.randn(n, generator=g) * 0.1)
            elif isinstance(m, (torch.nn.Conv2d, torch.nn.Linear)):
                m.bias.data.copy_(torch.randn(m.bias.shape, generator=g) * 0.05)
    return {k: v.detach().clone() for k, v in net.state_dict().items()}


def make_checkpoint(seed: int = 0) -> dict:
    """Checkpoint dict in the reference's schema (``main.py:22-26``)."""
    return {"epoch": 0, "model_state_dict": make_state_dict(seed), "best_accuracy": 0.0}


def images_u8(n: int, seed: int = 0, device: str | torch.device = "cpu") -> torch.Tensor:
    """``[n,3,32,32]`` uint8 pixels, U{0..255}; generated on CPU for cross-device determinism."""
    g = torch.Generator().manual_seed(seed)
    x = torch.randint(0, 256, (n,) + IMAGE_SHAPE, dtype=torch.uint8, generator=g)
    return x.to(device)


def normalize(x_u8: torch.Tensor) -> torch.Tensor:
    """uint8 pixels -> fp32 NCHW exactly as ToTensor()+Normalize() would."""
    mean = torch.tensor(CIFAR_MEAN, dtype=torch.float32, device=x_u8.device).view(1, 3, 1, 1)
    std = torch.tensor(CIFAR_STD, dtype=torch.float32, device=x_u8.device).view(1, 3, 1, 1)
    return ((x_u8.to(torch.float32) / 255.0) - mean) / std


def images_f32(n: int, seed: int = 0, device: str | torch.device = "cpu") -> torch.Tensor:
    return normalize(images_u8(n, seed, device)).contiguous()


def calibration_batches(seed: int = CALIB_SEED, batches: int = CALIB_BATCHES, batch: int = CALIB_BATCH):
    """The fixed calibration set (4 x 64 images) every static-PTQ model here is prepared on."""
    return [images_f32(batch, seed + i) for i in range(batches)]


class SyntheticLoader:
    """Minimal DataLoader look-alike: iterable of ``(images fp32 NCHW, labels int64)`` on CPU.

    Labels are the fp32 net's argmax so that "accuracy" is meaningful without CIFAR-10.
    """

    def __init__(self, n: int, batch_size: int, seed: int = 0, label_model: torch.nn.Module | None = None):
        self.batch_size = batch_size
        self.images = images_f32(n, seed)
        if label_model is None:
            g = torch.Generator().manual_seed(seed + 1)
            self.labels = torch.randint(0, 10, (n,), generator=g)
        else:
            label_model = label_model.eval()
            with torch.no_grad():
                self.labels = torch.cat([label_model(self.images[i:i + 256]).argmax(1)
                                         for i in range(0, n, 256)])

    def __len__(self):
        return (self.images.shape[0] + self.batch_size - 1) // self.batch_size

    def __iter__(self):
        for i in range(0, self.images.shape[0], self.batch_size):
            yield self.images[i:i + self.batch_size], self.labels[i:i + self.batch_size]
