"""Layer builders with the reference's module/parameter names (pointnet2_lib/pointnet2/pytorch_utils.py).

Only what the hot path's callers need -- SharedMLP, Conv1d, Conv2d, BatchNorm1d/2d, FC -- built so that
`state_dict()` keys match the reference exactly (e.g. `layer0.conv.weight`, `layer0.bn.bn.running_mean`):
checkpoints move between the two code bases unchanged.
"""
from typing import List, Optional

import torch.nn as nn


def _named_bn(kind, channels: int, name: str = "") -> nn.Sequential:
    """The reference wraps every BatchNorm in a one-element Sequential called `<name>bn` (:101-122)."""
    wrapper = nn.Sequential()
    bn = kind(channels)
    nn.init.constant_(bn.weight, 1.0)
    nn.init.constant_(bn.bias, 0)
    wrapper.add_module(name + "bn", bn)
    return wrapper


class BatchNorm1d(nn.Sequential):
    def __init__(self, in_size: int, *, name: str = ""):
        super().__init__()
        for k, v in _named_bn(nn.BatchNorm1d, in_size, name).named_children():
            self.add_module(k, v)


class BatchNorm2d(nn.Sequential):
    def __init__(self, in_size: int, name: str = ""):
        super().__init__()
        for k, v in _named_bn(nn.BatchNorm2d, in_size, name).named_children():
            self.add_module(k, v)


class _ConvBlock(nn.Sequential):
    """conv -> [bn] -> [activation] (or bn/activation first when preact), names as in :35-101."""

    _conv = None
    _bn = None
    _inorm = None

    def __init__(self, in_size, out_size, *, kernel_size, stride, padding, activation=nn.ReLU(inplace=True),
                 bn=False, init=nn.init.kaiming_normal_, bias=True, preact=False, name="", instance_norm=False):
        super().__init__()
        conv = self._conv(in_size, out_size, kernel_size=kernel_size, stride=stride, padding=padding,
                          bias=bias and not bn)
        init(conv.weight)
        if conv.bias is not None:
            nn.init.constant_(conv.bias, 0)
        norm_width = in_size if preact else out_size
        extras = []
        if bn:
            extras.append((name + "bn", self._bn(norm_width)))
        if activation is not None:
            extras.append((name + "activation", activation))
        if not bn and instance_norm:
            extras.append((name + "in", self._inorm(norm_width, affine=False, track_running_stats=False)))
        order = extras + [(name + "conv", conv)] if preact else [(name + "conv", conv)] + extras
        for key, mod in order:
            self.add_module(key, mod)


class Conv1d(_ConvBlock):
    _conv, _bn, _inorm = nn.Conv1d, BatchNorm1d, nn.InstanceNorm1d

    def __init__(self, in_size: int, out_size: int, *, kernel_size: int = 1, stride: int = 1, padding: int = 0, **kw):
        super().__init__(in_size, out_size, kernel_size=kernel_size, stride=stride, padding=padding, **kw)


class Conv2d(_ConvBlock):
    _conv, _bn, _inorm = nn.Conv2d, BatchNorm2d, nn.InstanceNorm2d

    def __init__(self, in_size: int, out_size: int, *, kernel_size=(1, 1), stride=(1, 1), padding=(0, 0), **kw):
        super().__init__(in_size, out_size, kernel_size=kernel_size, stride=stride, padding=padding, **kw)


class SharedMLP(nn.Sequential):
    """[1x1 Conv2d (no bias under BN) -> BN2d -> ReLU] per consecutive pair of `args` (:5-32)."""

    def __init__(self, args: List[int], *, bn: bool = False, activation=nn.ReLU(inplace=True), preact: bool = False,
                 first: bool = False, name: str = "", instance_norm: bool = False):
        super().__init__()
        for i in range(len(args) - 1):
            plain_first = first and preact and i == 0
            self.add_module(
                name + "layer{}".format(i),
                Conv2d(args[i], args[i + 1], bn=bn and not plain_first,
                       activation=None if plain_first else activation, preact=preact, instance_norm=instance_norm))


class FC(nn.Sequential):
    def __init__(self, in_size: int, out_size: int, *, activation=nn.ReLU(inplace=True), bn: bool = False,
                 init=None, preact: bool = False, name: str = ""):
        super().__init__()
        fc = nn.Linear(in_size, out_size, bias=not bn)
        if init is not None:
            init(fc.weight)
        if not bn:
            nn.init.constant_(fc.bias, 0)
        extras = []
        if bn:
            extras.append((name + "bn", BatchNorm1d(in_size if preact else out_size)))
        if activation is not None:
            extras.append((name + "activation", activation))
        order = extras + [(name + "fc", fc)] if preact else [(name + "fc", fc)] + extras
        for key, mod in order:
            self.add_module(key, mod)
