"""LSQ / CiM quantized modules with the reference's public surface (``models/_modules/lsq.py:21``):
``Conv2dLSQ``, ``LinearLSQ``, ``ActLSQ``, ``Conv2dLSQCiM`` -- same constructor arguments, attributes,
parameters and buffers -- computing through the sm_100a kernels of ``libcimq.so``.

There is no CPU path: CPU inputs raise.
"""
import math

import torch
import torch.nn.functional as F

from .. import _lib
from .. import distributed as CD
from .. import functional as CF
from ._quan_base import _ActQ, _Conv2dQ, _Conv2dQCiM, _LinearQ, _LinearQCiM, grad_scale, round_pass

__all__ = ['Conv2dLSQ', 'LinearLSQ', 'ActLSQ', 'Conv2dLSQCiM', 'LinearLSQCiM', 'get_cim_output_signed']

get_cim_output_signed = CF.get_cim_output_signed


class _CiMForward:
    """What Conv2dLSQCiM and LinearLSQCiM share: lazy initialisation (lsq.py:532-563), the alpha_cim quantiser
    (lsq.py:566-571) and the fused CiM convolution on ``(x [B,C,H,W], w [Cout,C,k,k])``."""

    # -- host mirror of the init buffers -----------------------------------------------------------
    def _flags(self):
        """(init_state != 0, init_state_cim != 0) without a device read per forward: the host copy is refreshed only
        when a buffer was written since the last read (tensor version counters catch fill_ / copy_ / load_state_dict /
        a DDP broadcast; the reference reads both buffers on every forward, lsq.py:532, 557)."""
        seen = (self.init_state._version, self.init_state_cim._version, self.init_state.data_ptr(),
                self.init_state_cim.data_ptr())
        if self._flags_stale or seen != getattr(self, "_flags_seen", None):
            self._init_done = bool(self.init_state.item() != 0)
            self._init_cim_done = bool(self.init_state_cim.item() != 0)
            self._flags_stale = False
            self._flags_seen = seen
        return self._init_done, self._init_cim_done

    @torch.no_grad()
    def _lazy_init(self, x, w):
        """First training batch: data-dependent step sizes (lsq.py:532-542)."""
        qp_a = 2 ** self.nbits_a - 1
        qp_w = 2 ** (self.nbits_w - 1) - 1
        # with several ranks the statistics are reduced over the global batch; no-ops otherwise
        if CD.global_min_(x.min()) < -1e-5:
            self.signed_act.data.fill_(1)
        self.alpha_act.data.copy_(2 * CD.global_mean_(x.abs().mean()) / math.sqrt(qp_a))
        self.alpha_weight.data.copy_(2 * w.abs().mean() / math.sqrt(qp_w))
        self.init_state.fill_(1)
        self._init_done = True
        self._note_flags()

    def _note_flags(self):
        self._flags_seen = (self.init_state._version, self.init_state_cim._version, self.init_state.data_ptr(),
                            self.init_state_cim.data_ptr())

    @torch.no_grad()
    def _lazy_init_cim(self, x, w, stride, padding):
        """First training batch: ``alpha_cim = 2*mean|psum|/sqrt(Qp_adc)`` (lsq.py:557-563)."""
        spec = CF._make_spec(x.shape, w.shape, stride, padding, self.nbits_a, self.abitslice, self.nbits_w,
                             self.wbitslice, self.xbar, self.adcbits)
        qp_a = 2 ** self.nbits_a - 1
        qn_w, qp_w = -(2 ** (self.nbits_w - 1)), 2 ** (self.nbits_w - 1) - 1
        ga = 1.0 / math.sqrt(x.numel() * qp_a)
        gw = 1.0 / math.sqrt(w.numel() * qp_w)
        s = _lib.step_sizes(self.alpha_act.data, self.alpha_weight.data, ga, gw)
        xcodes = _lib.lsq_quantize(x.contiguous(), s[0:1], 0, qp_a)
        wcodes = _lib.lsq_quantize(w.data.contiguous(), s[1:2], qn_w, qp_w)
        self.alpha_cim.data.copy_(CF.alpha_cim_initial_value(spec, xcodes, wcodes, s, qp_adc=1.0,
                                                             reduce_sums=CD.global_sum_))
        self.init_state_cim.fill_(1)
        self._init_cim_done = True
        self._note_flags()

    def _alpha_q(self):
        """``nbits_alpha``-bit range quantiser of alpha_cim, inside autograd (lsq.py:566-571): (alpha_q, step)."""
        return CF.alpha_quantize(self.alpha_cim, self.nbits_alpha)

    def accepts_codes(self, x_shape=None) -> bool:
        """True when this layer can take its activation codes from the producer of its input
        (``functional.batch_norm_act(..., next_conv=self)``): step sizes initialised, unsigned activations, and the
        fused v2 path in use.  Checked on the host mirror of the init flags, no device read."""
        init_done, init_cim_done = self._flags()
        if not init_done or (self.alpha_cim is not None and not init_cim_done) or self.stochastic_quant:
            return False
        return self.nbits_alpha <= 11 and not (self.kernel_flags & _lib.FLAG_FORCE_SIMT) and self.abitslice == 1

    def _cim_forward(self, x, w, stride, padding, xcodes=None):
        """[B,C,H,W] x [Cout,C,k,k] -> [B,Cout,OH,OW] through the crossbar model (adcbits != 0)."""
        if not x.is_cuda:
            raise RuntimeError(f"{type(self).__name__} (cim_quantization_b200) needs CUDA tensors; "
                               "there is no CPU fallback")
        init_done, init_cim_done = self._flags()
        if self.training and not init_done:
            self._lazy_init(x, w)
        if self.binary_mask.device != x.device:
            self.binary_mask = self.binary_mask.to(x.device)
        if self.training and not init_cim_done and self.alpha_cim is not None:
            self._lazy_init_cim(x, w, stride, padding)
        # v2 kernels (covered shape, alpha_q / scale an fp16 integer < 2048, deterministic read-out): the raw
        # alpha_cim goes in and one launch prepares everything that depends only on the parameters
        spec = CF._make_spec(x.shape, w.shape, stride, padding, self.nbits_a, self.abitslice, self.nbits_w,
                             self.wbitslice, self.xbar, self.adcbits)
        if (self.nbits_alpha <= 11 and not self.stochastic_quant and not (self.kernel_flags & _lib.FLAG_FORCE_SIMT)
                and _lib.v2_usable(spec, self.alpha_cim is not None, True, self.kernel_flags)):
            return CF.cim_conv2d_v2(x, w, self.alpha_act, self.alpha_weight, self.alpha_cim, self.binary_mask, stride,
                                    padding, self.nbits_a, self.abitslice, self.nbits_w, self.wbitslice, self.xbar,
                                    self.adcbits, self.nbits_alpha, self.kernel_flags,
                                    xcodes=xcodes if init_done and (init_cim_done or self.alpha_cim is None) else None)
        alpha_q, alpha_scale = self._alpha_q() if self.alpha_cim is not None else (None, None)
        if self.nbits_alpha > 11:  # the v2 kernels carry alpha_q / scale as an fp16 integer (< 2048)
            alpha_scale = None
        return CF.cim_conv2d(x, w, self.alpha_act, self.alpha_weight, alpha_q, self.binary_mask, stride, padding,
                             self.nbits_a, self.abitslice, self.nbits_w, self.wbitslice, self.xbar, self.adcbits,
                             self.kernel_flags, stochastic=bool(self.stochastic_quant), alpha_scale=alpha_scale)


class Conv2dLSQCiM(_Conv2dQCiM, _CiMForward):
    """Crossbar-aware, bit-sliced, partial-sum-quantized convolution (lsq.py:511-588)."""

    def __init__(self, in_channels, out_channels, kernel_size, stride=1, padding=0, dilation=1, groups=1,
                 bias=True, nbits_w=8, nbits_a=8, nbits_alpha=8, wbitslice=1, abitslice=1, xbar=64, adcbits=6,
                 stochastic_quant=False, **kwargs):
        # unknown keywords (e.g. signed_xbar from main_lsq.py:56) are accepted and ignored, as in the reference
        super().__init__(in_channels=in_channels, out_channels=out_channels, kernel_size=kernel_size,
                         stride=stride, padding=padding, dilation=dilation, groups=groups, bias=bias,
                         nbits_w=nbits_w, nbits_a=nbits_a, nbits_alpha=nbits_alpha, wbitslice=wbitslice,
                         abitslice=abitslice, xbar=xbar, adcbits=adcbits, stochastic_quant=stochastic_quant)
        self.kernel_flags = 0  # _lib.FLAG_FORCE_SIMT selects the CUDA-core kernels (tests)

    def _spec(self, x):
        return CF._make_spec(x.shape, self.weight.shape, self.stride, self.padding, self.nbits_a, self.abitslice,
                             self.nbits_w, self.wbitslice, self.xbar, self.adcbits)

    def forward(self, x, xcodes=None):
        """``xcodes`` (extension, optional): the uint8 activation codes of ``x`` already written by its producer
        (``functional.batch_norm_act(..., next_conv=self)``); ignored unless :meth:`accepts_codes`."""
        if not x.is_cuda:
            raise RuntimeError("Conv2dLSQCiM (cim_quantization_b200) needs CUDA tensors; there is no CPU fallback")
        if self.dilation[0] != 1 or self.groups != 1:
            raise ValueError("Conv2dLSQCiM supports dilation=1, groups=1 (reference envelope, lsq.py:141,153)")
        if self.adcbits != 0:
            out = self._cim_forward(x, self.weight, self.stride, self.padding, xcodes=xcodes)
            if self.bias is not None:
                out = out + self.bias  # same broadcast as lsq.py:582-583
            return out
        # adcbits == 0: no crossbar model, plain convolution of the fake-quantised operands (lsq.py:584-585)
        init_done, _ = self._flags()
        if self.training and not init_done:
            self._lazy_init(x, self.weight)
        qp_a = 2 ** self.nbits_a - 1
        qn_w, qp_w = -(2 ** (self.nbits_w - 1)), 2 ** (self.nbits_w - 1) - 1
        x_q, _ = CF.lsq_fake_quant(x, self.alpha_act, 1.0 / math.sqrt(x.numel() * qp_a), 0, qp_a, True)
        w_q, _ = CF.lsq_fake_quant(self.weight, self.alpha_weight, 1.0 / math.sqrt(self.weight.numel() * qp_w),
                                   qn_w, qp_w, True)
        return F.conv2d(x_q, w_q, self.bias, self.stride, self.padding, self.dilation)


class LinearLSQCiM(_LinearQCiM, _CiMForward):
    """Crossbar-aware fully connected layer: ``Conv2dLSQCiM`` with a 1x1 kernel over one pixel per sample, on
    an ``nn.Linear`` surface so ``replace_map={'Linear': [LinearLSQCiM]}`` works with the reference's
    ``ReplaceModuleTool`` (utils/wrapper/replace_module.py:34-64).  Same quantisation, crossbar chunking, ADC
    and straight-through backward (lsq.py:92-386); the reference itself has no CiM linear."""

    def __init__(self, in_features, out_features, bias=True, nbits_w=8, nbits_a=8, nbits_alpha=8, wbitslice=1,
                 abitslice=1, xbar=64, adcbits=6, stochastic_quant=False, **kwargs):
        super().__init__(in_features=in_features, out_features=out_features, bias=bias, nbits_w=nbits_w,
                         nbits_a=nbits_a, nbits_alpha=nbits_alpha, wbitslice=wbitslice, abitslice=abitslice,
                         xbar=xbar, adcbits=adcbits, stochastic_quant=stochastic_quant)
        self.kernel_flags = 0

    def forward(self, x):
        if self.adcbits == 0:
            raise ValueError("LinearLSQCiM needs a crossbar ADC model (adcbits != 0); use LinearLSQ otherwise")
        lead = x.shape[:-1]
        x4 = x.reshape(-1, self.in_features, 1, 1)
        w4 = self.weight.view(self.out_features, self.in_features, 1, 1)
        out = self._cim_forward(x4, w4, (1, 1), (0, 0)).reshape(*lead, self.out_features)
        if self.bias is not None:
            out = out + self.bias
        return out


class Conv2dLSQ(_Conv2dQ):
    """Plain LSQ conv on integer activation codes (lsq.py:389-436): input is ``(codes, act_scale)``."""

    def __init__(self, in_channels, out_channels, kernel_size, stride=1, padding=0, dilation=1, groups=1,
                 bias=True, nbits_w=8, **kwargs):
        super().__init__(in_channels=in_channels, out_channels=out_channels, kernel_size=kernel_size,
                         stride=stride, padding=padding, dilation=dilation, groups=groups, bias=bias,
                         nbits=nbits_w)

    def forward(self, x):
        if self.alpha is None:
            return F.conv2d(x, self.weight, self.bias, self.stride, self.padding, self.dilation, self.groups)
        x_q, act_scaling_factor = x
        qn = -2 ** (self.nbits - 1)
        qp = 2 ** (self.nbits - 1) - 1
        if self.training and self.init_state == 0:
            self.alpha.data.copy_(2 * self.weight.abs().mean() / math.sqrt(qp))
            self.init_state.fill_(1)
        g = 1.0 / math.sqrt(self.weight.numel() * qp)
        w_codes, _ = CF.lsq_fake_quant(self.weight, self.alpha, g, qn, qp, False)
        s_w = grad_scale(self.alpha, g)  # lsq.py:426
        return F.conv2d(x_q, w_codes, self.bias, self.stride, self.padding, self.dilation,
                        self.groups) * act_scaling_factor * s_w


class LinearLSQ(_LinearQ):
    """LSQ fake-quantised weights + ``F.linear`` (lsq.py:591-617)."""

    def __init__(self, in_features, out_features, bias=True, nbits_w=4, **kwargs):
        super().__init__(in_features=in_features, out_features=out_features, bias=bias, nbits=nbits_w)

    def forward(self, x):
        if self.alpha is None:
            return F.linear(x, self.weight, self.bias)
        qn = -2 ** (self.nbits - 1)
        qp = 2 ** (self.nbits - 1) - 1
        if self.training and self.init_state == 0:
            self.alpha.data.copy_(2 * self.weight.abs().mean() / math.sqrt(qp))
            self.init_state.fill_(1)
        g = 1.0 / math.sqrt(self.weight.numel() * qp)
        w_q, _ = CF.lsq_fake_quant(self.weight, self.alpha, g, qn, qp, True)
        return F.linear(x, w_q, self.bias)


class ActLSQ(_ActQ):
    """LSQ activation quantiser returning ``(codes, step size)`` (lsq.py:620-662)."""

    def __init__(self, nbits_a=4, **kwargs):
        super().__init__(nbits=nbits_a)

    def _range(self):
        if self.signed == 1:
            return -2 ** (self.nbits - 1), 2 ** (self.nbits - 1) - 1
        return 0, 2 ** self.nbits - 1

    def forward(self, x):
        if self.alpha is None:
            return x
        if self.training and self.init_state == 0:
            if x.min() < -1e-5:
                self.signed.data.fill_(1)
            _, qp = self._range()
            self.alpha.data.copy_(2 * x.abs().mean() / math.sqrt(qp))
            self.init_state.fill_(1)
        qn, qp = self._range()
        g = 1.0 / math.sqrt(x.numel() * qp)
        codes, _ = CF.lsq_fake_quant(x, self.alpha, g, qn, qp, False)
        return codes, grad_scale(self.alpha, g)
