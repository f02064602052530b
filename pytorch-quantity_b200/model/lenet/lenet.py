"""LeNet-style MNIST classifier in tracer-compatible form (architecture of the reference's
quantity/model/lenet/lenet.py:12-31: two conv + ReLU + max-pool stages, a ``View`` flatten and three
stacked ``nn.Linear``).  No BatchNorm, a single input channel, 3x3 and 5x5 kernels and output widths that are
not multiples of 16: the shapes the ResNets do not exercise."""
import torch
from torch import nn

from common.quantity import View


class Cnn(nn.Module):
    def __init__(self, in_dim=1, n_class=10):
        super().__init__()
        self.conv = nn.Sequential(
            nn.Conv2d(in_dim, 6, 3, stride=1, padding=1), nn.ReLU(False), nn.MaxPool2d(2, 2),
            nn.Conv2d(6, 16, 5, stride=1, padding=0), nn.ReLU(False), nn.MaxPool2d(2, 2))
        self.review = View()
        self.fc = nn.Sequential(nn.Linear(400, 120), nn.Linear(120, 84), nn.Linear(84, n_class))

    def forward(self, x):
        return self.fc(self.review(self.conv(x)))


def lenet_batches(n_batches, batch, seed=1):
    """Synthetic MNIST-like calibration batches: (1x28x28 images in [0, 1), label None) tuples, the loader
    form ``PRE_PROCESS.IMG: 1`` expects (quantity/test/lenet_quantity.py:12-20)."""
    out = []
    for i in range(n_batches):
        g = torch.Generator().manual_seed(seed + i)
        out.append((torch.rand(batch, 1, 28, 28, generator=g), None))
    return out
