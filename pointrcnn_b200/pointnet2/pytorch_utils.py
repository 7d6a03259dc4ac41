"""Mirror of pointnet2_lib/pointnet2/pytorch_utils.py: SharedMLP / Conv1d / Conv2d / FC / BatchNorm blocks.

Same class names, constructor arguments and -- what checkpoints depend on -- the same sub-module names, so
state_dict keys are identical to the reference's:
    SharedMLP:  layer{i}.conv.{weight,bias}, layer{i}.bn.bn.{weight,bias,running_mean,running_var,num_batches_tracked}
    Conv1d/2d:  conv.*, bn.bn.*, (activation), (in)
    FC:         fc.*, bn.bn.*
(reference: pytorch_utils.py:5-236).  The blocks are plain torch.nn; the fused tensor-core path of
pointnet2_modules.py reads their parameters, it does not change them.
"""
from typing import List, Tuple

import torch.nn as nn


def _norm_block(kind, channels, name):
    """BatchNorm wrapper whose inner module is called 'bn' (-> keys '<name>bn.bn.*'), weight=1 / bias=0"""
    block = nn.Sequential()
    block.add_module(name + "bn", kind(channels))
    nn.init.constant_(block[0].weight, 1.0)
    nn.init.constant_(block[0].bias, 0)
    return block


class BatchNorm1d(nn.Sequential):
    def __init__(self, in_size: int, *, name: str = ""):
        super().__init__()
        for k, m in _norm_block(nn.BatchNorm1d, in_size, name).named_children():
            self.add_module(k, m)


class BatchNorm2d(nn.Sequential):
    def __init__(self, in_size: int, name: str = ""):
        super().__init__()
        for k, m in _norm_block(nn.BatchNorm2d, in_size, name).named_children():
            self.add_module(k, m)


class _ConvNd(nn.Sequential):
    """conv (+bias iff no bn) -> [bn] -> [activation] -> [instance norm], or the pre-activation order"""
    _conv = None
    _bn = None
    _inorm = None

    def __init__(self, in_size, out_size, *, kernel_size, stride, padding, activation=nn.ReLU(inplace=True), bn=False,
                 init=nn.init.kaiming_normal_, bias=True, preact=False, name="", instance_norm=False):
        super().__init__()
        conv = self._conv(in_size, out_size, kernel_size=kernel_size, stride=stride, padding=padding, bias=bias and not bn)
        init(conv.weight)
        if conv.bias is not None:
            nn.init.constant_(conv.bias, 0)
        norm_ch = in_size if preact else out_size
        tail = []
        if bn:
            tail.append((name + "bn", self._bn(norm_ch)))
        if activation is not None:
            tail.append((name + "activation", activation))
        if not bn and instance_norm:
            tail.append((name + "in", self._inorm(norm_ch, affine=False, track_running_stats=False)))
        order = tail + [(name + "conv", conv)] if preact else [(name + "conv", conv)] + tail
        for k, m in order:
            self.add_module(k, m)


class Conv1d(_ConvNd):
    _conv, _bn, _inorm = nn.Conv1d, BatchNorm1d, nn.InstanceNorm1d

    def __init__(self, in_size: int, out_size: int, *, kernel_size: int = 1, stride: int = 1, padding: int = 0,
                 activation=nn.ReLU(inplace=True), bn: bool = False, init=nn.init.kaiming_normal_, bias: bool = True,
                 preact: bool = False, name: str = "", instance_norm=False):
        super().__init__(in_size, out_size, kernel_size=kernel_size, stride=stride, padding=padding, activation=activation,
                         bn=bn, init=init, bias=bias, preact=preact, name=name, instance_norm=instance_norm)


class Conv2d(_ConvNd):
    _conv, _bn, _inorm = nn.Conv2d, BatchNorm2d, nn.InstanceNorm2d

    def __init__(self, in_size: int, out_size: int, *, kernel_size: Tuple[int, int] = (1, 1), stride: Tuple[int, int] = (1, 1),
                 padding: Tuple[int, int] = (0, 0), activation=nn.ReLU(inplace=True), bn: bool = False,
                 init=nn.init.kaiming_normal_, bias: bool = True, preact: bool = False, name: str = "", instance_norm=False):
        super().__init__(in_size, out_size, kernel_size=kernel_size, stride=stride, padding=padding, activation=activation,
                         bn=bn, init=init, bias=bias, preact=preact, name=name, instance_norm=instance_norm)


class SharedMLP(nn.Sequential):
    def __init__(self, args: List[int], *, bn: bool = False, activation=nn.ReLU(inplace=True), preact: bool = False,
                 first: bool = False, name: str = "", instance_norm: bool = False):
        super().__init__()
        for i in range(len(args) - 1):
            plain = (not first) or (not preact) or (i != 0)
            self.add_module(name + "layer{}".format(i),
                            Conv2d(args[i], args[i + 1], bn=plain and bn, activation=activation if plain else None,
                                   preact=preact, instance_norm=instance_norm))


class FC(nn.Sequential):
    def __init__(self, in_size: int, out_size: int, *, activation=nn.ReLU(inplace=True), bn: bool = False, init=None,
                 preact: bool = False, name: str = ""):
        super().__init__()
        fc = nn.Linear(in_size, out_size, bias=not bn)
        if init is not None:
            init(fc.weight)
        if not bn:
            nn.init.constant_(fc.bias, 0)
        tail = []
        if bn:
            tail.append((name + "bn", BatchNorm1d(in_size if preact else out_size)))
        if activation is not None:
            tail.append((name + "activation", activation))
        order = tail + [(name + "fc", fc)] if preact else [(name + "fc", fc)] + tail
        for k, m in order:
            self.add_module(k, m)
