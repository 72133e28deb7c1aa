"""Model helpers with the reference's `models/wrapper.py` surface.

Functions (same names, argument meaning and return values; /root/reference/models/wrapper.py:10-50):
    get_nb_parameters(model), save(model, filename), load(model, filename),
    train(model, device, data, criterion, optimizer, epochs),
    eval(model, device, data) -> (softmax predictions [N, C] on `device`, LongTensor targets on CPU),
    accuracy(predictions, labels) -> float
Models (wrapper.py:53-119): BaseNet_750 (748 parameters), BaseNet_15k (15 080 parameters), both with
`weight_init_gaussian(std)` / `weight_init_uniform(lim)`.  Additions used by the BASELINE configs:
`LeNet5` and `MLP` (the reference has no MLP).

`eval` collects per-batch logits in a list and concatenates once (the reference re-concatenates the
growing tensor every batch, which is quadratic in the number of batches; the result is identical).
"""
from __future__ import annotations

from typing import Iterable, Sequence

import numpy as np
import torch
import torch.nn as nn
import torch.nn.functional as F
from torch.nn import init


def get_nb_parameters(model):
    total = sum(p.numel() for p in model.parameters())
    print('Total params: %.2fK' % (total / 1000.0))
    return total


def save(model, filename):
    print('Writting %s\n' % filename)
    torch.save(model.state_dict(), filename)


def load(model, filename):
    print('Reading %s\n' % filename)
    model.load_state_dict(torch.load(filename))


def train(model, device, data, criterion, optimizer, epochs):
    model.train()
    for _ in range(epochs):
        for images, labels in data:
            loss = criterion(model(images.to(device)), labels.to(device))
            model.zero_grad()
            loss.backward()
            optimizer.step()


def eval(model, device, data):  # noqa: A001  (name fixed by the reference API)
    model.eval()
    chunks, targets = [], []
    with torch.no_grad():
        for images, labels in data:
            chunks.append(model(images.to(device)))
            targets.append(labels)
    if not chunks:
        return torch.Tensor().to(device), torch.LongTensor()
    logits = torch.cat(chunks)
    return F.softmax(logits, dim=1), torch.cat(targets).long().cpu()


def accuracy(predictions, labels):
    acc = 100 * np.mean(np.argmax(predictions.cpu().numpy(), axis=1) == labels.numpy())
    print(f"Accuracy: {acc:.2f}%")
    return acc


class _InitMixin:
    """weight_init_gaussian / weight_init_uniform of the reference nets (wrapper.py:68-84, 103-119):
    weights ~ N(0, std) or U(-lim, lim), biases zero, for every Linear / Conv2d."""

    def _init_layers(self) -> Iterable[nn.Module]:
        for layer in self.modules():
            name = layer.__class__.__name__
            if name in ('Linear', 'Conv2d'):
                yield layer
            elif name == 'MultiheadAttention':
                raise NotImplementedError

    def weight_init_gaussian(self, std):
        for layer in self._init_layers():
            init.normal_(layer.weight, 0, std)
            layer.bias.data.fill_(0)

    def weight_init_uniform(self, lim):
        for layer in self._init_layers():
            init.uniform_(layer.weight, -lim, lim)
            layer.bias.data.fill_(0)


class BaseNet_750(_InitMixin, nn.Module):
    """conv(1->3,k3) / pool / conv(3->6,k3,s2) / pool / fc 54->10.  wrapper.py:53-66."""

    def __init__(self):
        super().__init__()
        self.conv1 = nn.Conv2d(1, 3, kernel_size=3, stride=1)
        self.pool = nn.MaxPool2d(2, 2)
        self.conv2 = nn.Conv2d(3, 6, kernel_size=3, stride=2)
        self.fc1 = nn.Linear(6 * 3 * 3, 10)

    def forward(self, x):
        x = self.pool(F.relu(self.conv1(x)))
        x = self.pool(F.relu(self.conv2(x)))
        return self.fc1(torch.flatten(x, 1))


class BaseNet_15k(_InitMixin, nn.Module):
    """conv(1->5,k5) / pool / conv(5->10,k5) / pool / fc 160->80 / fc 80->10.  wrapper.py:86-101."""

    def __init__(self):
        super().__init__()
        self.conv1 = nn.Conv2d(1, 5, 5)
        self.pool = nn.MaxPool2d(2, 2)
        self.conv2 = nn.Conv2d(5, 10, 5)
        self.fc1 = nn.Linear(10 * 4 * 4, 80)
        self.fc2 = nn.Linear(80, 10)

    def forward(self, x):
        x = self.pool(F.relu(self.conv1(x)))
        x = self.pool(F.relu(self.conv2(x)))
        x = F.relu(self.fc1(torch.flatten(x, 1)))
        return self.fc2(x)


class LeNet5(_InitMixin, nn.Module):
    """True LeNet-5 shapes for BASELINE config 4: conv(1->6,k5,pad2) / pool / conv(6->16,k5) / pool /
    fc 400->120->84->10."""

    def __init__(self):
        super().__init__()
        self.conv1 = nn.Conv2d(1, 6, 5, padding=2)
        self.pool = nn.MaxPool2d(2, 2)
        self.conv2 = nn.Conv2d(6, 16, 5)
        self.fc1 = nn.Linear(16 * 5 * 5, 120)
        self.fc2 = nn.Linear(120, 84)
        self.fc3 = nn.Linear(84, 10)

    def forward(self, x):
        x = self.pool(F.relu(self.conv1(x)))
        x = self.pool(F.relu(self.conv2(x)))
        x = F.relu(self.fc1(torch.flatten(x, 1)))
        x = F.relu(self.fc2(x))
        return self.fc3(x)


class MLP(_InitMixin, nn.Module):
    """Linear/ReLU stack, e.g. MLP([784, 1024, 1024, 10]) (BASELINE config 1) or
    MLP([4096, 4096, 4096, 4096, 10]) (config 5); inputs are flattened to [N, sizes[0]]."""

    def __init__(self, sizes: Sequence[int]):
        super().__init__()
        self.sizes = list(sizes)
        self.layers = nn.ModuleList(nn.Linear(a, b) for a, b in zip(sizes[:-1], sizes[1:]))

    def forward(self, x):
        x = torch.flatten(x, 1)
        for i, layer in enumerate(self.layers):
            x = layer(x)
            if i + 1 < len(self.layers):
                x = F.relu(x)
        return x
