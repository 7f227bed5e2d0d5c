"""autograd.Function surface of the CiM conv path, backed by ``libcimq.so``.

* ``get_cim_output_signed`` keeps the reference's 17-argument signature and its 17-tuple of
  gradients (``models/_modules/lsq.py:89-386``) so existing callers work unchanged.
* ``cim_conv2d`` is the fused path ``Conv2dLSQCiM.forward`` uses: it goes from fp32 activations /
  weights and the learned step sizes straight to integer codes (the fake-quant floats of
  lsq.py:549/555 are never materialised) and its backward also produces the LSQ step-size
  gradients that the reference obtains through autograd.
"""
from __future__ import annotations

import math

import torch
from torch.autograd import Function

from . import _lib
from ._lib import LayerSpec


def _pair0(v):
    return int(v[0]) if isinstance(v, (tuple, list)) else int(v)


def _as_mask_2d(binary_mask, nsw, nsa, device):
    """binary_mask arrives as int8 ``[1,1,NSW,NSA,1,1]`` (``_quan_base.py:207-214``)."""
    m = binary_mask.to(device=device, dtype=torch.int8).reshape(nsw, nsa).contiguous()
    return m


def _make_spec(x_shape, w_shape, stride, padding, nbits_a, abitslice, nbits_w, wbitslice, xbar, adcbits) -> LayerSpec:
    b, cin, h, w = x_shape
    cout, cin_w, kh, kw = w_shape
    if h != w or kh != kw:
        raise ValueError("CiM conv supports square images and kernels only (reference envelope, lsq.py:123,369)")
    if cin_w != cin:
        raise ValueError("groups != 1 is not supported by the CiM conv (lsq.py:153)")
    return LayerSpec(batch=int(b), in_channels=int(cin), in_hw=int(h), out_channels=int(cout), kernel=int(kh),
                     stride=_pair0(stride), padding=_pair0(padding), nbits_a=int(nbits_a), abitslice=int(abitslice),
                     nbits_w=int(nbits_w), wbitslice=int(wbitslice), xbar=int(xbar), adcbits=adcbits)


def _require_cuda(*tensors, f32: bool = True):
    for t in tensors:
        if t is None:
            continue
        if not t.is_cuda:
            raise RuntimeError("cim_quantization_b200 runs on CUDA tensors only; there is no CPU fallback")
        if f32 and t.dtype != torch.float32:  # the kernels read raw fp32; the reference is fp32 end to end as well
            raise TypeError(f"cim_quantization_b200 expects float32 tensors, got {t.dtype}")


class get_cim_output_signed(Function):
    """Drop-in for the reference Function (lsq.py:89): same arguments, output ``[B, L, Cout]``,
    gradients at positions 0 (x_q), 1 (w_q) and 12 (alpha_cim).

    The integer codes are recovered from the fake-quant floats as ``rint(x_q / s)``: identical to
    the reference wherever its own ``x_q / s`` is exact, and free of its truncation
    artefacts elsewhere.
    """

    @staticmethod
    def forward(ctx, x, w, conv_stride, conv_padding, conv_dilation, act_bits, act_bit_slice, weight_bits,
                weight_bit_slice, adc_bits, arr, binary_mask, alpha_cim, weight_scaling_factor,
                act_scaling_factor, stochastic, signed_act):
        if stochastic and (adc_bits != 1.5 or alpha_cim is None):
            stochastic = False  # the reference only samples in the adcbits 1.5 branch (lsq.py:203-205)
        if _pair0(conv_dilation) != 1:
            raise ValueError("dilation != 1 is not supported (the reference's Unfold ignores it, lsq.py:141)")
        _require_cuda(x, w, f32=False)  # converted with .float() below
        spec = _make_spec(x.shape, w.shape, conv_stride, conv_padding, act_bits, act_bit_slice, weight_bits,
                          weight_bit_slice, arr, adc_bits)
        info = _lib.layer_info(spec)
        dev = x.device
        s = torch.cat([act_scaling_factor.detach().reshape(1), weight_scaling_factor.detach().reshape(1)]).to(
            device=dev, dtype=torch.float32)
        qp_a = 2 ** int(act_bits) - 1
        qn_w, qp_w = -(2 ** (int(weight_bits) - 1)), 2 ** (int(weight_bits) - 1) - 1
        xcodes = _lib.lsq_quantize(x.detach().contiguous().float(), s[0:1], 0, qp_a, from_fakequant=True)
        wcodes = _lib.lsq_quantize(w.detach().contiguous().float(), s[1:2], qn_w, qp_w, from_fakequant=True)
        mask = _as_mask_2d(binary_mask, info.NSW, info.NSA, dev)
        has_alpha = adc_bits in (1, 1.5)
        alpha_q = alpha_cim.detach().contiguous().float() if has_alpha else None
        table = _lib.adc_table(spec, s, alpha_q, mask)
        need_bwd = any(ctx.needs_input_grad)
        wdigits, wtiles = _lib.weight_prepare(spec, wcodes, want_digits=need_bwd and not info.tc_backward)
        if stochastic:  # sampled read-out, lsq.py:205-220
            out, state = _lib.conv_forward_stochastic(spec, xcodes, wcodes, table, s, alpha_q, _stochastic_seed(),
                                                      save_state=need_bwd)
        else:
            # an arbitrary alpha_cim tensor carries no quantiser step, so only the multi-bit ADC can take the v2 kernels
            flags = _lib.FLAG_V2 if _lib.v2_usable(spec, has_alpha, None) else 0
            out, state = _lib.conv_forward(spec, xcodes, wcodes, wtiles, table, s, mask, save_state=need_bwd,
                                           flags=flags)
        ctx.spec, ctx.has_alpha, ctx.w_shape = spec, has_alpha, tuple(w.shape)
        ctx.save_for_backward(xcodes, wdigits, wtiles, state, s, mask)
        return out.transpose(1, 2)  # [B, L, Cout] like lsq.py:233 (a view of the NCHW buffer)

    @staticmethod
    def backward(ctx, grad_output):
        xcodes, wdigits, wtiles, state, s, mask = ctx.saved_tensors
        go = grad_output.transpose(1, 2).contiguous().float()  # [B, Cout, L]
        gxq, gwq, galpha = _lib.conv_backward(ctx.spec, go, xcodes, wdigits, wtiles, state, s, mask,
                                              need_alpha=ctx.has_alpha and ctx.needs_input_grad[12],
                                              need_input=ctx.needs_input_grad[0])
        gwq = gwq.view(ctx.w_shape)
        return (gxq, gwq, None, None, None, None, None, None, None, None, None, None, galpha, None, None, None,
                None)


class _CimConv2dFused(Function):
    """Fused ``Conv2dLSQCiM`` hot path: LSQ quantisation of x and w, CiM conv, and the complete
    backward (grad x, grad w, the three step-size gradients)."""

    @staticmethod
    def forward(ctx, x, weight, alpha_act, alpha_weight, alpha_q, binary_mask, stride, padding, nbits_a,
                abitslice, nbits_w, wbitslice, xbar, adcbits, flags, stochastic_seed=None, alpha_scale=None):
        _require_cuda(x, weight, alpha_act, alpha_weight)
        x = x.contiguous()
        weight = weight.contiguous()
        spec = _make_spec(x.shape, weight.shape, stride, padding, nbits_a, abitslice, nbits_w, wbitslice, xbar,
                          adcbits)
        info = _lib.layer_info(spec)
        qp_a = 2 ** spec.nbits_a - 1  # activations always clamp to [0, Qp_a], lsq.py:537-538
        qn_w, qp_w = -(2 ** (spec.nbits_w - 1)), 2 ** (spec.nbits_w - 1) - 1
        ga = 1.0 / math.sqrt(x.numel() * qp_a)  # lsq.py:547
        gw = 1.0 / math.sqrt(weight.numel() * qp_w)  # lsq.py:553
        s = _lib.step_sizes(alpha_act.detach(), alpha_weight.detach(), ga, gw)
        xcodes = _lib.lsq_quantize(x.detach(), s[0:1], 0, qp_a)
        wcodes = _lib.lsq_quantize(weight.detach(), s[1:2], qn_w, qp_w)
        mask = _as_mask_2d(binary_mask, info.NSW, info.NSA, x.device)
        has_alpha = alpha_q is not None
        sampled = stochastic_seed is not None and has_alpha and adcbits == 1.5
        use_v2 = not sampled and _lib.v2_usable(spec, has_alpha, alpha_scale, flags)
        table = _lib.adc_table(spec, s, alpha_q.detach().contiguous() if has_alpha else None, mask,
                               alpha_scale=alpha_scale if use_v2 else None)
        if use_v2:
            flags |= _lib.FLAG_V2
        need_bwd = any(ctx.needs_input_grad)
        simt = bool(flags & _lib.FLAG_FORCE_SIMT)
        wdigits, wtiles = _lib.weight_prepare(spec, wcodes, want_digits=need_bwd and (simt or not info.tc_backward))
        if sampled:  # lsq.py:203-220
            out, state = _lib.conv_forward_stochastic(spec, xcodes, wcodes, table, s, alpha_q.detach(),
                                                      stochastic_seed, save_state=need_bwd)
        else:
            out, state = _lib.conv_forward(spec, xcodes, wcodes, wtiles, table, s, mask, save_state=need_bwd,
                                           flags=flags)
        ctx.spec, ctx.has_alpha, ctx.flags = spec, has_alpha, flags
        ctx.consts = (qp_a, qn_w, qp_w, ga, gw)
        ctx.save_for_backward(x, weight, xcodes, wdigits, wtiles, state, s, mask)
        return out.view(spec.batch, spec.out_channels, info.out_hw, info.out_hw)

    @staticmethod
    def backward(ctx, grad_y):
        x, weight, xcodes, wdigits, wtiles, state, s, mask = ctx.saved_tensors
        spec = ctx.spec
        qp_a, qn_w, qp_w, ga, gw = ctx.consts
        go = grad_y.contiguous().float().view(spec.batch, spec.out_channels, -1)
        need_x = ctx.needs_input_grad[0] or ctx.needs_input_grad[2]
        gxq, gwq, galpha = _lib.conv_backward(spec, go, xcodes, wdigits, wtiles, state, s, mask,
                                              need_alpha=ctx.has_alpha and ctx.needs_input_grad[4],
                                              need_input=need_x, flags=ctx.flags)
        gx = g_aa = None
        if need_x:
            gx, g_aa = _lib.lsq_backward(gxq, x, s[0:1], 0, qp_a, ga)
        gwt, g_aw = _lib.lsq_backward(gwq.view_as(weight), weight, s[1:2], qn_w, qp_w, gw)
        return (gx, gwt, g_aa, g_aw, galpha) + (None,) * 12


class _CimConv2dFusedV2(Function):
    """``Conv2dLSQCiM`` step on the v2 kernels with the parameter-only work in ONE launch: layer_prepare (step sizes,
    weight codes, alpha quantiser, ADC table / constants, weight tiles), activation quantiser, CiM conv -- three
    launches forward.  The backward returns the gradients of x, weight, alpha_act, alpha_weight and alpha_cim (the
    alpha quantiser's backward, lsq.py:566-571, is applied here instead of in a separate autograd node)."""

    @staticmethod
    def forward(ctx, x, weight, alpha_act, alpha_weight, alpha_cim, binary_mask, stride, padding, nbits_a, abitslice,
                nbits_w, wbitslice, xbar, adcbits, nbits_alpha, flags, xcodes_in=None):
        _require_cuda(x, weight, alpha_act, alpha_weight, alpha_cim)
        x = x.contiguous()
        weight = weight.contiguous()
        spec = _make_spec(x.shape, weight.shape, stride, padding, nbits_a, abitslice, nbits_w, wbitslice, xbar,
                          adcbits)
        info = _lib.layer_info(spec)
        qp_a = 2 ** spec.nbits_a - 1
        qn_w, qp_w = -(2 ** (spec.nbits_w - 1)), 2 ** (spec.nbits_w - 1) - 1
        ga = 1.0 / math.sqrt(x.numel() * qp_a)  # lsq.py:547
        gw = 1.0 / math.sqrt(weight.numel() * qp_w)  # lsq.py:553
        mask = _as_mask_2d(binary_mask, info.NSW, info.NSA, x.device)
        has_alpha = alpha_cim is not None
        a_cim = alpha_cim.detach().contiguous() if has_alpha else None
        s, wcodes, alpha_q, aux, table, wtiles = _lib.layer_prepare(
            spec, weight.detach(), alpha_act.detach(), alpha_weight.detach(), ga, gw, a_cim, 1, 2 ** nbits_alpha - 1,
            mask)
        # codes already written by the producer of x (batch_norm_act(..., next_conv=...))?
        xcodes = xcodes_in if xcodes_in is not None else _lib.lsq_quantize(x.detach(), s[0:1], 0, qp_a)
        need_bwd = any(ctx.needs_input_grad)
        out, state = _lib.conv_forward(spec, xcodes, wcodes, wtiles, table, s, mask, save_state=need_bwd,
                                       flags=flags | _lib.FLAG_V2)
        ctx.spec, ctx.has_alpha, ctx.flags = spec, has_alpha, flags
        ctx.consts = (qp_a, qn_w, qp_w, ga, gw, nbits_alpha)
        ctx.save_for_backward(x, weight, xcodes, wtiles, state, s, mask, a_cim, aux)
        return out.view(spec.batch, spec.out_channels, info.out_hw, info.out_hw)

    @staticmethod
    def backward(ctx, grad_y):
        x, weight, xcodes, wtiles, state, s, mask, a_cim, aux = ctx.saved_tensors
        spec = ctx.spec
        qp_a, qn_w, qp_w, ga, gw, nbits_alpha = ctx.consts
        go = grad_y.contiguous().float().view(spec.batch, spec.out_channels, -1)
        need_x = ctx.needs_input_grad[0] or ctx.needs_input_grad[2]
        need_alpha = ctx.has_alpha and ctx.needs_input_grad[4]
        gxq, gwq, galpha_q = _lib.conv_backward(spec, go, xcodes, None, wtiles, state, s, mask, need_alpha=need_alpha,
                                                need_input=need_x, flags=ctx.flags)
        gx = g_aa = galpha = None
        if need_x:
            gx, g_aa = _lib.lsq_backward(gxq, x, s[0:1], 0, qp_a, ga)
        gwt, g_aw = _lib.lsq_backward(gwq.view_as(weight), weight, s[1:2], qn_w, qp_w, gw)
        if need_alpha:
            galpha = _lib.alpha_quantize_backward(a_cim, galpha_q.view_as(a_cim), 1, 2 ** nbits_alpha - 1, aux)
        return (gx, gwt, g_aa, g_aw, galpha) + (None,) * 12


def cim_conv2d_v2(x, weight, alpha_act, alpha_weight, alpha_cim, binary_mask, stride, padding, nbits_a, abitslice,
                  nbits_w, wbitslice, xbar, adcbits, nbits_alpha, flags: int = 0, xcodes=None):
    """Fused CiM convolution of ``Conv2dLSQCiM`` (lsq.py:546-581) from the RAW ``alpha_cim`` parameter, for layers the
    v2 kernels cover (``_lib.layer_info(spec).tc_v2``, ``nbits_alpha <= 11``)."""
    return _CimConv2dFusedV2.apply(x, weight, alpha_act, alpha_weight, alpha_cim, binary_mask, stride, padding,
                                   nbits_a, abitslice, nbits_w, wbitslice, xbar, adcbits, nbits_alpha, flags, xcodes)


def _stochastic_seed() -> int:
    """A fresh 62-bit seed from torch's CPU generator (so torch.manual_seed makes the sampling reproducible)."""
    return int(torch.randint(0, 2 ** 62, (1,)).item())


def cim_conv2d(x, weight, alpha_act, alpha_weight, alpha_q, binary_mask, stride, padding, nbits_a, abitslice,
               nbits_w, wbitslice, xbar, adcbits, flags: int = 0, stochastic: bool = False, alpha_scale=None):
    """Fused CiM convolution of ``Conv2dLSQCiM`` (lsq.py:546-581): returns ``[B, Cout, OH, OW]``.
    ``stochastic``: the sampled near-ADC-less read-out of lsq.py:205-220 (adcbits 1.5 only; CUDA-core kernel).
    ``alpha_scale``: the step of the alpha quantiser that produced ``alpha_q`` (second output of
    :func:`alpha_quantize`); with it the layer runs on the v2 kernels where they cover it."""
    return _CimConv2dFused.apply(x, weight, alpha_act, alpha_weight, alpha_q, binary_mask, stride, padding, nbits_a,
                                 abitslice, nbits_w, wbitslice, xbar, adcbits, flags,
                                 _stochastic_seed() if stochastic else None, alpha_scale)


class _LsqFakeQuant(Function):
    """``round_pass(clamp(x / grad_scale(alpha, g), qn, qp)) [* s]`` with its STE / step-size backward
    (lsq.py:23-32 and the ``Method1`` blocks at lsq.py:426-427, 608-609, 653-654)."""

    @staticmethod
    def forward(ctx, x, alpha, g, qn, qp, rescale):
        _require_cuda(x, alpha)
        x = x.contiguous()
        s = _lib.step_sizes(alpha.detach(), alpha.detach(), g, g)
        y = _lib.lsq_fakequant(x.detach(), s[0:1], qn, qp, rescale)
        ctx.save_for_backward(x, s)
        ctx.consts = (g, qn, qp, rescale)
        s1 = s[0:1].clone()  # the tensor that is returned is the one marked (a view of `s` would stay differentiable)
        ctx.mark_non_differentiable(s1)
        return y, s1

    @staticmethod
    def backward(ctx, grad_y, _grad_s):
        x, s = ctx.saved_tensors
        g, qn, qp, rescale = ctx.consts
        gy = grad_y.contiguous().float()
        if rescale:
            gx, galpha = _lib.lsq_backward(gy, x, s[0:1], qn, qp, g)
        else:
            # y = round_pass(clamp(x/s)):  dy/dx = mask / s,  dy/ds = -(x/s) * mask / s   (tiny-use path: torch ops)
            u = x / s[0:1]
            inside = (u >= qn) & (u <= qp)
            gx = gy * inside / s[0:1]
            galpha = (-(gy * u * inside).sum() * g / s[0:1]).reshape(1)
        return gx, galpha, None, None, None, None


def lsq_fake_quant(x, alpha, g: float, qn: int, qp: int, rescale: bool = True):
    """Returns ``(y, s)`` with ``s = grad_scale(alpha, g)`` (value only, as a 1-element tensor)."""
    return _LsqFakeQuant.apply(x, alpha, g, qn, qp, rescale)


class _AlphaQuant(Function):
    """``nbits_alpha``-bit range quantiser of alpha_cim with torch-autograd-equivalent backward
    (lsq.py:566-571), one kernel each way instead of ~25 elementwise launches."""

    @staticmethod
    def forward(ctx, alpha, qn, qp):
        _require_cuda(alpha)
        a = alpha.detach().contiguous()
        aq, aux = _lib.alpha_quantize(a, qn, qp)
        ctx.save_for_backward(a, aux)
        ctx.q = (qn, qp)
        scale = aux[0:1].clone()
        ctx.mark_non_differentiable(scale)
        return aq, scale

    @staticmethod
    def backward(ctx, grad_aq, _grad_scale):
        a, aux = ctx.saved_tensors
        return _lib.alpha_quantize_backward(a, grad_aq.contiguous().float(), ctx.q[0], ctx.q[1], aux), None, None


def alpha_quantize(alpha, nbits_alpha: int):
    """``(alpha_q, scale)``: the quantised alpha_cim (differentiable, lsq.py:566-571) and the quantiser's step
    ``(max - min) / (Qp - 1)`` as a 1-element tensor (not differentiable; alpha_q = n * scale, n integer)."""
    return _AlphaQuant.apply(alpha, 1, 2 ** nbits_alpha - 1)


def alpha_cim_initial_value(spec: LayerSpec, xcodes, wcodes, s, qp_adc: float = 1.0, reduce_sums=None):
    """Data-dependent initial ``alpha_cim`` (lsq.py:557-563): ``2*mean|psum*s_w*s_a| / sqrt(Qp_adc)``
    over (batch, pixel), zeros replaced by ``s_w*s_a``.  The |psum| sums are exact integers.
    ``reduce_sums(t) -> (t, ranks)`` may sum them over data-parallel ranks (distributed.global_sum_)."""
    info = _lib.layer_info(spec)
    sums = _lib.conv_psum_abs_sums(spec, xcodes, wcodes)  # int64 [1,NX,NSW,NSA,1,Cout]
    ranks = 1
    if reduce_sums is not None:
        sums, ranks = reduce_sums(sums)
    sa, sw = s[0].double(), s[1].double()
    t = 2.0 * (sums.double() / float(ranks * spec.batch * info.L)) * sw * sa / math.sqrt(qp_adc)
    t = torch.where(t == 0, sw * sa, t)
    return t.float()


class _BatchNormAct(torch.autograd.Function):
    """BatchNorm2d (+ residual) (+ ReLU) in two launches per direction (csrc/bn_fused.cu)."""

    @staticmethod
    def forward(ctx, x, residual, weight, bias, running_mean, running_var, training, momentum, eps, relu):
        x = x.contiguous()
        res = residual.contiguous() if residual is not None else None
        y, mean, invstd = _lib.bn_forward(x, res, weight, bias, running_mean, running_var, training, momentum, eps,
                                          relu)
        if not training:
            mean, invstd = running_mean, torch.rsqrt(running_var + eps)
        ctx.save_for_backward(x, y if relu else None, weight, mean, invstd)
        ctx.training, ctx.relu, ctx.has_res = training, relu, residual is not None
        return y

    @staticmethod
    def backward(ctx, gy):
        x, y, weight, mean, invstd = ctx.saved_tensors
        gx, gres, gw, gb = _lib.bn_backward(gy.contiguous(), x, y, weight, mean, invstd, ctx.training, ctx.relu,
                                            ctx.has_res and ctx.needs_input_grad[1])
        return (gx, gres, gw if weight is not None else None, gb if ctx.needs_input_grad[3] else None, None, None,
                None, None, None, None)


class _BatchNormActQuant(torch.autograd.Function):
    """:class:`_BatchNormAct` whose forward kernel also writes the activation codes of the layer that consumes its
    output (resnet.py:83-86 + lsq.py:547-549 in one pass).  Returns ``(y, codes)``; codes carry no
    gradient -- the consumer's quantiser backward works on y, as before."""

    @staticmethod
    def forward(ctx, x, residual, weight, bias, running_mean, running_var, training, momentum, eps, relu, next_alpha,
                next_gscale, next_qp):
        x = x.contiguous()
        res = residual.contiguous() if residual is not None else None
        y, mean, invstd, codes = _lib.bn_forward(x, res, weight, bias, running_mean, running_var, training, momentum,
                                                 eps, relu, next_quant=(next_alpha.detach(), next_gscale, next_qp))
        if not training:
            mean, invstd = running_mean, torch.rsqrt(running_var + eps)
        ctx.save_for_backward(x, y if relu else None, weight, mean, invstd)
        ctx.training, ctx.relu, ctx.has_res = training, relu, residual is not None
        ctx.mark_non_differentiable(codes)
        return y, codes

    @staticmethod
    def backward(ctx, gy, _gcodes):
        x, y, weight, mean, invstd = ctx.saved_tensors
        gx, gres, gw, gb = _lib.bn_backward(gy.contiguous(), x, y, weight, mean, invstd, ctx.training, ctx.relu,
                                            ctx.has_res and ctx.needs_input_grad[1])
        return (gx, gres, gw if weight is not None else None, gb if ctx.needs_input_grad[3] else None) + (None,) * 9


def batch_norm_act(x, bn: torch.nn.BatchNorm2d, residual=None, relu: bool = False, next_conv=None):
    """``relu?(bn(x) + residual?)`` with ``bn``'s parameters, buffers and train/eval mode.  With ``next_conv`` (a
    ``Conv2dLSQCiM`` that will consume the result and whose step sizes are initialised) the return value is
    ``(y, codes)``: pass ``codes`` to ``next_conv(y, xcodes=codes)``."""
    _require_cuda(x, residual)
    training = bn.training or not bn.track_running_stats
    momentum = 0.0 if bn.momentum is None else bn.momentum
    if training and bn.track_running_stats and bn.num_batches_tracked is not None:
        bn.num_batches_tracked.add_(1)
        if bn.momentum is None:
            momentum = 1.0 / float(bn.num_batches_tracked)
    if next_conv is not None:
        qp = 2 ** next_conv.nbits_a - 1
        gscale = 1.0 / math.sqrt(x.numel() * qp)  # lsq.py:547 (the consumer sees a tensor of the same size)
        return _BatchNormActQuant.apply(x, residual, bn.weight, bn.bias,
                                        bn.running_mean if bn.track_running_stats else None,
                                        bn.running_var if bn.track_running_stats else None, training, momentum, bn.eps,
                                        relu, next_conv.alpha_act, gscale, qp)
    return _BatchNormAct.apply(x, residual, bn.weight, bn.bias, bn.running_mean if bn.track_running_stats else None,
                               bn.running_var if bn.track_running_stats else None, training, momentum, bn.eps, relu)
