"""A small tracer-compatible network that exercises every cared op type (Conv2d,
Linear, Eltwise, Concat) and all three merge-group rules of
quantity/tools/pytorch_quantizer.py:396-465.  Test helper; imports ``common.quantity``
by name so it runs on the reference package (golden generation) and on this repo's."""
import torch
import torch.nn as nn

from common.quantity import Concat, Eltwise, View


class TinyFabuNet(nn.Module):
    def __init__(self):
        super().__init__()
        self.conv0 = nn.Sequential(nn.Conv2d(3, 8, 3, padding=1, bias=False),
                                   nn.BatchNorm2d(8), nn.ReLU(False))
        self.pool = nn.MaxPool2d(2)
        self.branch_a = nn.Sequential(nn.Conv2d(8, 8, 3, padding=1, bias=False), nn.BatchNorm2d(8))
        self.branch_b = nn.Sequential(nn.Conv2d(8, 8, 1, bias=False), nn.BatchNorm2d(8))
        self.Concat = Concat()                      # group [branch_a, branch_b]: shared interval + summed hist
        self.relu1 = nn.ReLU(False)
        self.conv_c = nn.Sequential(nn.Conv2d(16, 8, 3, padding=1, bias=False), nn.BatchNorm2d(8))
        self.Eltwise1 = Eltwise()                   # group [conv_c, conv0] (pool/relu pruned): no Eltwise member
        self.relu2 = nn.ReLU(False)
        self.conv_d = nn.Sequential(nn.Conv2d(8, 8, 3, stride=1, padding=1, bias=True), nn.BatchNorm2d(8))
        self.Eltwise2 = Eltwise()                   # group [conv_d, Eltwise1]: bit(conv_d) := bit(Eltwise1)
        self.relu3 = nn.ReLU(False)
        self.avgpool = nn.AvgPool2d(8)
        self.view = View()
        self.fc = nn.Linear(8, 10)

    def forward(self, x):
        p = self.pool(self.conv0(x))
        c = self.relu1(self.Concat(self.branch_a(p), self.branch_b(p)))
        e1 = self.relu2(self.Eltwise1(self.conv_c(c), p))
        e2 = self.relu3(self.Eltwise2(self.conv_d(e1), e1))
        return self.fc(self.view(self.avgpool(e2)))


def build_tiny(seed=0):
    torch.manual_seed(seed)
    net = TinyFabuNet().eval()
    g = torch.Generator().manual_seed(seed + 7)
    for m in net.modules():
        if isinstance(m, nn.BatchNorm2d):
            m.running_mean.copy_(torch.randn(m.running_mean.shape, generator=g) * 0.1)
            m.running_var.copy_(torch.rand(m.running_var.shape, generator=g) + 0.5)
            m.weight.data.copy_(torch.rand(m.weight.shape, generator=g) + 0.5)
            m.bias.data.copy_(torch.randn(m.bias.shape, generator=g) * 0.1)
    return net


def tiny_batches(n_batches=3, batch=2, seed=1):
    out = []
    for i in range(n_batches):
        g = torch.Generator().manual_seed(seed + i)
        out.append((torch.randn(batch, 3, 16, 16, generator=g), None))
    return out
