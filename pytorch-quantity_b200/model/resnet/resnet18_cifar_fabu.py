"""CIFAR-10 ResNet-18 in tracer-compatible form, with the module names of the reference's
quantity/model/resnet/ResNet_18_fabu.py:11-70 (``conv1``, ``layer1..4[i].left / .shortcut``, ``avePool2d``,
``view``, ``fc``) so that a ``resnet18.pth`` trained with the reference's train-18.py loads with
``load_state_dict``.  3x3 stride-1 stem, no max-pool, 4x4 average pool: 32x32 inputs."""
import torch.nn as nn

from common.quantity import View

from .resnet_fabu import BasicUnit, _conv_bn


class CifarResNet18(nn.Module):
    def __init__(self, num_classes=10):
        super().__init__()
        self.conv1 = nn.Sequential(*_conv_bn(3, 64, 3, 1, 1), nn.ReLU(False))
        stages, cin = [], 64
        for i, width in enumerate((64, 128, 256, 512)):
            stages.append(nn.Sequential(BasicUnit(cin, width, 1 if i == 0 else 2), BasicUnit(width, width, 1)))
            cin = width
        self.layer1, self.layer2, self.layer3, self.layer4 = stages
        self.avePool2d = nn.AvgPool2d(4)
        self.fc = nn.Linear(512, num_classes)
        self.view = View()

    def forward(self, x):
        x = self.layer4(self.layer3(self.layer2(self.layer1(self.conv1(x)))))
        return self.fc(self.view(self.avePool2d(x)))


def ResNet18(num_classes=10):
    return CifarResNet18(num_classes)
