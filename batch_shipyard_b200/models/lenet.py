"""LeNet-style MNIST network: the model behind the MNIST examples most of the reference's single-framework recipes launch
(Caffe `lenet_solver`, Keras `mnist_cnn.py`, Chainer `train_mnist.py`, Lua Torch / PyTorch / TensorFlow MNIST demos:
/root/reference/recipes/{Caffe,Keras+Theano,Chainer,Torch,PyTorch,TensorFlow}-CPU/README.md).  Two 5x5 convolutions with 2x2 max-pooling
and two fully connected layers, ~431 k parameters; plain PyTorch ops (the layers are too small for the tcgen05 paths), trained by the same
fused data-parallel trainer as ResNet-50 (flat parameters, one fused all-reduce + SGD per step)."""
from __future__ import annotations

import torch
import torch.nn as nn
import torch.nn.functional as F


class LeNet(nn.Module):
    s2d_stem = False          # the trainer feeds plain NHWC images (no space-to-depth stem)

    def __init__(self, num_classes: int = 10, in_channels: int = 1):
        super().__init__()
        self.conv1 = nn.Conv2d(in_channels, 20, 5)
        self.conv2 = nn.Conv2d(20, 50, 5)
        self.fc1 = nn.Linear(50 * 4 * 4, 500)
        self.fc2 = nn.Linear(500, num_classes)

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        x = F.max_pool2d(self.conv1(x), 2)
        x = F.max_pool2d(self.conv2(x), 2)
        x = F.relu(self.fc1(x.flatten(1)))
        return self.fc2(x)


def lenet(num_classes: int = 10) -> LeNet:
    return LeNet(num_classes)


def synthetic_digits(n: int, seed: int, num_classes: int = 10, size: int = 28):
    """MNIST-shaped synthetic data with a learnable signal: every class is a fixed random stroke template (shared by all ranks, seed 0)
    drawn with per-sample shift, intensity jitter and pixel noise.  Returns uint8 NHWC images [n, size, size, 1] and int64 labels."""
    g0 = torch.Generator().manual_seed(1234)
    templates = (torch.rand(num_classes, size, size, generator=g0) > 0.82).float()
    templates = F.avg_pool2d(templates.unsqueeze(1), 3, 1, 1).squeeze(1)                     # thicken the strokes
    templates = templates / templates.amax(dim=(1, 2), keepdim=True).clamp_min(1e-6)
    g = torch.Generator().manual_seed(seed)
    y = torch.randint(0, num_classes, (n,), generator=g)
    x = templates[y]
    dx, dy = torch.randint(-2, 3, (n,), generator=g), torch.randint(-2, 3, (n,), generator=g)
    x = torch.stack([torch.roll(img, (int(a), int(b)), (0, 1)) for img, a, b in zip(x, dy, dx)])
    x = x * (0.6 + 0.4 * torch.rand(n, 1, 1, generator=g)) + 0.25 * torch.rand(n, size, size, generator=g)
    return (x.clamp(0, 1) * 255).to(torch.uint8).unsqueeze(-1), y
