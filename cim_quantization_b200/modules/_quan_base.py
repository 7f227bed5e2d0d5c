"""Base classes of the quantized modules -- the state-dict / attribute contract of the reference's
``models/_modules/_quan_base.py`` (class names, constructor keywords, parameter and buffer names and
shapes), so checkpoints and ``utils/wrapper/replace_module.py`` work unchanged.

Only what the CiM / LSQ hot path needs is here; the reference's unused legacy helpers
(``truncation``, ``linear_quantize*``, ``get_sparsity_mask``, ``FunStopGradient``; _quan_base.py:31-103)
are out of scope.
"""
import math
from enum import Enum

import torch
import torch.nn as nn
from torch.nn.parameter import Parameter

__all__ = ['Qmodes', 'Qmodes_cim', '_Conv2dQ', '_LinearQ', '_ActQ', '_Conv2dQCiM', '_LinearQCiM', 'round_pass',
           'grad_scale']


class Qmodes(Enum):
    layer_wise = 1
    kernel_wise = 2


class Qmodes_cim(Enum):
    column_wise = 1
    bit_wise = 2


def grad_scale(x, scale):
    """Value ``x``, gradient ``scale`` (lsq.py:23-26): written as the reference writes it because the
    fp32 value ``(x - x*scale) + x*scale`` is not always bitwise ``x``."""
    y_grad = x * scale
    return x.detach() - y_grad.detach() + y_grad


def round_pass(x):
    """Round with a straight-through gradient (lsq.py:29-32)."""
    return x.round().detach() - x.detach() + x


def _default_kwargs_q(kwargs_q, layer):
    """Defaults the reference injects (``get_default_kwargs_q``, _quan_base.py:106-137)."""
    default = {'nbits': 4}
    if isinstance(layer, (_Conv2dQCiM, _LinearQCiM)):
        default['cimmode'] = Qmodes_cim.bit_wise
    if isinstance(layer, (_Conv2dQ, _Conv2dQCiM, _LinearQCiM)):
        default['mode'] = Qmodes.layer_wise
    for k, v in default.items():
        kwargs_q.setdefault(k, v)
    return kwargs_q


def _init_cim_state(layer, kwargs_q, flattened_dim, out_channels):
    """Attributes, parameters and buffers of a CiM layer (_quan_base.py:199-237)."""
    layer.kwargs_q = _default_kwargs_q(kwargs_q, layer)
    for key in ('nbits_w', 'nbits_a', 'nbits_alpha', 'wbitslice', 'abitslice', 'xbar', 'stochastic_quant',
                'adcbits'):
        setattr(layer, key, kwargs_q[key])
    if layer.nbits_w < 0:
        layer.register_parameter('alpha', None)
        layer.register_parameter('alpha_cim', None)
        return
    layer.q_mode = kwargs_q['mode']
    layer.num_xbars = int(math.ceil(flattened_dim / layer.xbar))
    layer.num_bit_slice_weight = int(layer.nbits_w / layer.wbitslice)
    layer.num_bit_slice_act = int(layer.nbits_a / layer.abitslice)
    nsw, nsa = layer.num_bit_slice_weight, layer.num_bit_slice_act
    mask = torch.empty(nsw, nsa, dtype=torch.int8)
    for i in range(nsa):
        for j in range(nsw):
            mask[j, i] = _wrap_int8(((2 ** layer.abitslice) ** i) * ((2 ** layer.wbitslice) ** j))
    layer.binary_mask = mask.view(1, 1, nsw, nsa, 1, 1)
    if layer.adcbits == 1.5 or layer.adcbits == 1:
        layer.alpha_cim = Parameter(torch.ones(1, layer.num_xbars, nsw, nsa, 1, out_channels))
    else:  # adcbits == 0 or > 1.5: no partial-sum scale factor
        layer.alpha_cim = None
    layer.alpha_weight = Parameter(torch.ones(1))
    layer.alpha_act = Parameter(torch.ones(1))
    layer.register_buffer('init_state', torch.zeros(1))
    layer.register_buffer('signed_act', torch.zeros(1))
    layer.register_buffer('init_state_cim', torch.zeros(1))
    layer._flags_stale = True  # host mirror of the three buffers (avoids a device sync per forward)


class _QBase:
    def add_param(self, param_k, param_v):
        self.kwargs_q[param_k] = param_v

    def set_bit(self, nbits):
        self.kwargs_q['nbits'] = nbits


class _Conv2dQ(nn.Conv2d, _QBase):
    """Plain LSQ conv base (_quan_base.py:140-172): ``alpha`` [1] (or [Cout]) + ``init_state``."""

    def __init__(self, in_channels, out_channels, kernel_size, stride=1, padding=0, dilation=1, groups=1,
                 bias=True, **kwargs_q):
        super().__init__(in_channels, out_channels, kernel_size, stride=stride, padding=padding,
                         dilation=dilation, groups=groups, bias=bias)
        self.kwargs_q = _default_kwargs_q(kwargs_q, self)
        self.nbits = kwargs_q['nbits']
        if self.nbits < 0:
            self.register_parameter('alpha', None)
            self.register_parameter('alpha_cim', None)
            return
        self.q_mode = kwargs_q['mode']
        n = out_channels if self.q_mode == Qmodes.kernel_wise else 1
        self.alpha = Parameter(torch.Tensor(n))
        self.register_buffer('init_state', torch.zeros(1))

    def extra_repr(self):
        s_prefix = super().extra_repr()
        if self.alpha is None:
            return '{}, fake'.format(s_prefix)
        return '{}, {}'.format(s_prefix, self.kwargs_q)


def _wrap_int8(v: int) -> int:
    """Two's-complement wrap of ``float -> int8``: the reference stores ``binary_mask`` as int8
    (_quan_base.py:214), so 128 -> -128 and larger powers of two -> 0 (matters for 8x8 slices)."""
    return ((int(v) + 128) % 256) - 128


class _Conv2dQCiM(nn.Conv2d, _QBase):
    """CiM conv base (_quan_base.py:174-249).

    Parameters: ``alpha_cim`` ``[1, NX, NSW, NSA, 1, Cout]`` (only for adcbits 1 / 1.5), ``alpha_weight`` [1],
    ``alpha_act`` [1].  Buffers: ``init_state``, ``signed_act``, ``init_state_cim``.  ``binary_mask`` is a plain
    int8 attribute ``[1,1,NSW,NSA,1,1]`` (not in the state dict), exactly as in the reference.
    """

    def __init__(self, in_channels, out_channels, kernel_size, stride=1, padding=0, dilation=1, groups=1,
                 bias=True, **kwargs_q):
        super().__init__(in_channels, out_channels, kernel_size, stride=stride, padding=padding,
                         dilation=dilation, groups=groups, bias=bias)
        ks = self.kernel_size  # nn.Conv2d normalises ints to tuples; the reference indexes the ctor arg
        _init_cim_state(self, kwargs_q, in_channels * ks[0] * ks[1], out_channels)

    def _load_from_state_dict(self, *args, **kwargs):
        super()._load_from_state_dict(*args, **kwargs)
        self._flags_stale = True

    def extra_repr(self):
        return '{}, {}'.format(super().extra_repr(), self.kwargs_q)


class _LinearQCiM(nn.Linear, _QBase):
    """CiM linear base: the parameter / buffer set of ``_Conv2dQCiM`` on an ``nn.Linear`` (a crossbar sees a
    fully connected layer as a 1x1 convolution over one pixel).  Not in the reference, which quantises only
    convolutions with the crossbar model; the state-dict names follow ``_Conv2dQCiM``."""

    def __init__(self, in_features, out_features, bias=True, **kwargs_q):
        super().__init__(in_features=in_features, out_features=out_features, bias=bias)
        _init_cim_state(self, kwargs_q, in_features, out_features)

    def _load_from_state_dict(self, *args, **kwargs):
        super()._load_from_state_dict(*args, **kwargs)
        self._flags_stale = True

    def extra_repr(self):
        return '{}, {}'.format(super().extra_repr(), self.kwargs_q)


class _LinearQ(nn.Linear, _QBase):
    """_quan_base.py:252-270."""

    def __init__(self, in_features, out_features, bias=True, **kwargs_q):
        super().__init__(in_features=in_features, out_features=out_features, bias=bias)
        self.kwargs_q = _default_kwargs_q(kwargs_q, self)
        self.nbits = kwargs_q['nbits']
        if self.nbits < 0:
            self.register_parameter('alpha', None)
            return
        self.alpha = Parameter(torch.Tensor(1))
        self.register_buffer('init_state', torch.zeros(1))

    def extra_repr(self):
        s_prefix = super().extra_repr()
        if self.alpha is None:
            return '{}, fake'.format(s_prefix)
        return '{}, {}'.format(s_prefix, self.kwargs_q)


class _ActQ(nn.Module, _QBase):
    """_quan_base.py:273-296."""

    def __init__(self, **kwargs_q):
        super().__init__()
        self.kwargs_q = _default_kwargs_q(kwargs_q, self)
        self.nbits = kwargs_q['nbits']
        if self.nbits < 0:
            self.register_parameter('alpha', None)
            return
        self.alpha = Parameter(torch.Tensor(1))
        self.register_buffer('init_state', torch.zeros(1))
        self.register_buffer('signed', torch.zeros(1))

    def extra_repr(self):
        if self.alpha is None:
            return 'fake'
        return '{}'.format(self.kwargs_q)
