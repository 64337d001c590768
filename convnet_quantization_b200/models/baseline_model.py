"""fp32 ``SimpleConvNet`` — host-side mirror of the reference's baseline net.

Same class name, attribute names and ``forward`` contract as the reference
(``models/baseline_model.py:5-83`` in his0si/ConvNet-Quantization), so a
``state_dict`` saved by the reference's trainer (``train_model.py:93-99``)
loads here unchanged and the reference drivers can call it unchanged.  The
topology is table-driven here; the arithmetic is plain ``torch.nn`` (on a GPU
box: cuDNN/cuBLAS via ATen) because the fp32 net is only the tolerance
baseline of the hot path, not the product.
"""
import torch
import torch.nn as nn
import torch.nn.functional as F

# (index, Cin, Cout, pool-after) for the six 3x3/s1/p1 convolutions
# reference: models/baseline_model.py:13-34
CONV_PLAN = (
    (1, 3, 64, False), (2, 64, 64, True),
    (3, 64, 128, False), (4, 128, 128, True),
    (5, 128, 256, False), (6, 256, 256, True),
)
FC_IN = 256 * 4 * 4
FC_HIDDEN = 512
NUM_CLASSES = 10
IMAGE_SHAPE = (3, 32, 32)


class SimpleConvNet(nn.Module):
    """6x(conv3x3+BN+ReLU) with 3 max-pools, then fc 4096->512(+BN+ReLU)->10."""

    def __init__(self):
        super().__init__()
        for i, cin, cout, pooled in CONV_PLAN:
            setattr(self, f"conv{i}", nn.Conv2d(cin, cout, kernel_size=3, padding=1))
            setattr(self, f"bn{i}", nn.BatchNorm2d(cout))
            if pooled:
                setattr(self, f"pool{i // 2}", nn.MaxPool2d(2, 2))
                setattr(self, f"dropout{i // 2}", nn.Dropout(0.25))
        self.fc1 = nn.Linear(FC_IN, FC_HIDDEN)
        self.bn7 = nn.BatchNorm1d(FC_HIDDEN)
        self.dropout4 = nn.Dropout(0.5)
        self.fc2 = nn.Linear(FC_HIDDEN, NUM_CLASSES)
        self._initialize_weights()

    def _initialize_weights(self):
        # reference: models/baseline_model.py:45-56 (kaiming fan_out, zero bias, unit BN)
        for m in self.modules():
            if isinstance(m, (nn.Conv2d, nn.Linear)):
                nn.init.kaiming_normal_(m.weight, mode="fan_out", nonlinearity="relu")
                if m.bias is not None:
                    nn.init.zeros_(m.bias)
            elif isinstance(m, nn.BatchNorm2d):
                nn.init.ones_(m.weight)
                nn.init.zeros_(m.bias)

    def forward(self, x):
        for i, _, _, pooled in CONV_PLAN:
            x = F.relu(getattr(self, f"bn{i}")(getattr(self, f"conv{i}")(x)))
            if pooled:
                x = getattr(self, f"dropout{i // 2}")(getattr(self, f"pool{i // 2}")(x))
        # reshape (not view): also valid for channels-last producers
        x = x.reshape(-1, FC_IN)
        x = self.dropout4(F.relu(self.bn7(self.fc1(x))))
        return self.fc2(x)


def test_model():
    model = SimpleConvNet()
    y = model(torch.randn(1, *IMAGE_SHAPE))
    print(f"Input shape: {(1,) + IMAGE_SHAPE}  Output shape: {tuple(y.shape)}")
    return model


if __name__ == "__main__":
    test_model()
