"""Benchmark / example harness around the hot path (NOT part of the re-implemented scope): a CIFAR ResNet-20
like the reference's ``models/cifar10/resnet.py`` (option-A shortcuts, ``resnet.py:73-75``) and the model
surgery of ``utils/wrapper/replace_module.py`` (every ``nn.Conv2d`` -> ``Conv2dLSQCiM``, first conv forced
to 8-bit weights/activations, ``replace_module.py:83-95``), so that the ResNet-20 configurations of
BASELINE.json can be run on the GPU box where the reference tree is not available.  When the reference IS
available, ``cim_quantization_b200.dropin.install()`` lets its own ``main_lsq.py`` drive these kernels.
"""
import torch
import torch.nn as nn
import torch.nn.functional as F

from . import functional as CF
from .modules import Conv2dLSQCiM

# fused batch norm (+ shortcut) + ReLU kernels of libcimq (csrc/bn_fused.cu) instead of cuDNN's spatial batch norm,
# which runs one block per channel and is 30 % of the ResNet-20 step at 16-64 channels; False restores torch's
FUSED_BN = True


# the fused batch norm also writes the activation codes of the convolution that consumes its output
# (one read of the fp32 activation and one launch less per layer); False keeps the separate quantiser kernel
FUSED_QUANT = True


def _bn_act(x, bn, residual=None, relu=True, next_conv=None):
    """-> (y, codes-or-None)"""
    if FUSED_BN and x.is_cuda:
        if (FUSED_QUANT and next_conv is not None and hasattr(next_conv, "accepts_codes") and next_conv.accepts_codes()):
            return CF.batch_norm_act(x, bn, residual, relu, next_conv=next_conv)
        return CF.batch_norm_act(x, bn, residual, relu), None
    out = bn(x)
    if residual is not None:
        out = out + residual
    return (F.relu(out) if relu else out), None


def _conv(conv, x, codes):
    return conv(x, xcodes=codes) if codes is not None else conv(x)


class _OptionA(nn.Module):
    def __init__(self, planes):
        super().__init__()
        self.pad = planes // 4

    def forward(self, x):
        return F.pad(x[:, :, ::2, ::2], (0, 0, 0, 0, self.pad, self.pad), "constant", 0)


class BasicBlock(nn.Module):
    def __init__(self, in_planes, planes, stride=1):
        super().__init__()
        self.conv1 = nn.Conv2d(in_planes, planes, 3, stride, 1, bias=False)
        self.bn1 = nn.BatchNorm2d(planes)
        self.conv2 = nn.Conv2d(planes, planes, 3, 1, 1, bias=False)
        self.bn2 = nn.BatchNorm2d(planes)
        self.shortcut = _OptionA(planes) if (stride != 1 or in_planes != planes) else nn.Sequential()

    next_conv = None  # the convolution that consumes this block's output (set by ResNetCifar), for the fused quantiser

    def forward(self, x, xcodes=None):
        """-> (out, codes of out for ``self.next_conv`` or None)"""
        out, c1 = _bn_act(_conv(self.conv1, x, xcodes), self.bn1, next_conv=self.conv2)
        return _bn_act(_conv(self.conv2, out, c1), self.bn2, self.shortcut(x), next_conv=self.next_conv)


class ResNetCifar(nn.Module):
    def __init__(self, num_blocks=(3, 3, 3), width=1, num_classes=10):
        super().__init__()
        p = 16 * width
        self.in_planes = p
        self.conv1 = nn.Conv2d(3, p, 3, 1, 1, bias=False)
        self.bn1 = nn.BatchNorm2d(p)
        self.layer1 = self._make_layer(p, num_blocks[0], 1)
        self.layer2 = self._make_layer(2 * p, num_blocks[1], 2)
        self.layer3 = self._make_layer(4 * p, num_blocks[2], 2)
        self.linear = nn.Linear(4 * p, num_classes)
        for m in self.modules():
            if isinstance(m, (nn.Conv2d, nn.Linear)):
                nn.init.kaiming_normal_(m.weight)

    def _make_layer(self, planes, n, stride):
        layers = []
        for s in [stride] + [1] * (n - 1):
            layers.append(BasicBlock(self.in_planes, planes, s))
            self.in_planes = planes
        return nn.Sequential(*layers)

    def _blocks(self):
        return [b for layer in (self.layer1, self.layer2, self.layer3) for b in layer]

    def forward(self, x):
        blocks = self._blocks()
        for b, nxt in zip(blocks, blocks[1:] + [None]):  # (object.__setattr__: not a registered submodule)
            object.__setattr__(b, "next_conv", nxt.conv1 if nxt is not None else None)
        out, codes = _bn_act(self.conv1(x), self.bn1, next_conv=blocks[0].conv1)
        for b in blocks:
            out, codes = b(out, codes)
        out = F.adaptive_avg_pool2d(out, 1).flatten(1)
        return self.linear(out)


def resnet20(width=1):
    """CIFAR ResNet-20 (19 convs + 1 linear); ``width=4`` is the 'wide' variant of BASELINE.json config 5."""
    return ResNetCifar((3, 3, 3), width)


def convert_to_cim(model, nbits_w=3, nbits_a=3, nbits_alpha=8, wbitslice=1, abitslice=1, xbar=128, adcbits=1.5,
                   stochastic_quant=False):
    """Replace every ``nn.Conv2d`` by ``Conv2dLSQCiM`` (weights copied), first conv at 8/8 bits -- the
    behaviour of ``ReplaceModuleTool(model, {'Conv2d': [Conv2dLSQCiM]}, True, ...)`` in the reference."""
    state = {"first": True}

    def recurse(mod):
        for name, child in list(mod._modules.items()):
            if isinstance(child, nn.Conv2d) and not isinstance(child, Conv2dLSQCiM):
                bw, ba = (8, 8) if state["first"] else (nbits_w, nbits_a)
                state["first"] = False
                new = Conv2dLSQCiM(child.in_channels, child.out_channels, child.kernel_size, child.stride,
                                   child.padding, child.dilation, groups=child.groups, bias=child.bias is not None,
                                   nbits_w=bw, nbits_a=ba, nbits_alpha=nbits_alpha, wbitslice=wbitslice,
                                   abitslice=abitslice, xbar=xbar, adcbits=adcbits, signed_xbar=False,
                                   stochastic_quant=stochastic_quant)
                new.weight.data.copy_(child.weight.data)
                if child.bias is not None:
                    new.bias.data.copy_(child.bias.data)
                mod._modules[name] = new.to(child.weight.device)
            elif child is not None and len(child._modules) > 0:
                recurse(child)

    recurse(model)
    return model


def sgd_param_groups(model, weight_decay=1e-4):
    """No weight decay on the learned step sizes (names containing 'alpha'), examples/__init__.py:181-196."""
    decay, no_decay = [], []
    for n, p in model.named_parameters():
        (no_decay if "alpha" in n else decay).append(p)
    return [{"params": decay, "weight_decay": weight_decay}, {"params": no_decay, "weight_decay": 0.0}]


class HostBatchPipeline:
    """Feeds batches that live in pinned host memory to the GPU through a side stream, ``depth`` batches ahead
    of the consumer, so the PCIe copy of step n+1 overlaps the kernels of step n (the CiM step moves 134 MB of
    fp32 activations / gradients per 1.8 ms of compute at the microbench shape: unpipelined, the copy dominates).

        pipe = HostBatchPipeline(device)
        for host_batch in loader:            # tuples of pinned CPU tensors
            pipe.submit(host_batch)          # asynchronous H2D on the copy stream
            ...
            dev_batch = pipe.get()           # the oldest submitted batch, ready on the current stream
            ... forward / backward ...
            pipe.release()                   # its device buffers may be overwritten by a later submit
    """

    def __init__(self, device, depth: int = 2):
        self.device = torch.device(device)
        self.depth = depth
        self.stream = torch.cuda.Stream(self.device)
        self.slots = [None] * depth          # device tensors per slot, allocated on first use
        self.ready = [torch.cuda.Event() for _ in range(depth)]
        self.freed = [torch.cuda.Event() for _ in range(depth)]
        self.n_submitted = self.n_got = self.n_released = 0

    def can_submit(self) -> bool:
        return self.n_submitted - self.n_released < self.depth

    def submit(self, host_tensors) -> None:
        if not self.can_submit():
            raise RuntimeError("HostBatchPipeline: all slots in flight; release() one first")
        k = self.n_submitted % self.depth
        if self.slots[k] is None:
            self.slots[k] = [torch.empty(t.shape, dtype=t.dtype, device=self.device) for t in host_tensors]
        with torch.cuda.stream(self.stream):
            if self.n_submitted >= self.depth:
                self.stream.wait_event(self.freed[k])  # the consumer of the previous batch in this slot is done
            for d, h in zip(self.slots[k], host_tensors):
                d.copy_(h, non_blocking=True)
            self.ready[k].record(self.stream)
        self.n_submitted += 1

    def get(self):
        if self.n_got >= self.n_submitted:
            raise RuntimeError("HostBatchPipeline: nothing submitted")
        k = self.n_got % self.depth
        torch.cuda.current_stream(self.device).wait_event(self.ready[k])
        self.n_got += 1
        return self.slots[k]

    def release(self) -> None:
        k = self.n_released % self.depth
        self.freed[k].record(torch.cuda.current_stream(self.device))
        self.n_released += 1
