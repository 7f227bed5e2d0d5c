"""CPU oracle for the CiM-aware quantized convolution path.  TEST INFRASTRUCTURE ONLY.

This file is a numpy restatement of the algorithm in the reference repository
(UtkarshSaxena1/CiM_Quantization, ``models/_modules/lsq.py`` and ``_quan_base.py``).
Citations below are ``file:line`` relative to the reference tree.

Who may import this module: ``tests/``, ``__graft_entry__.smoke()`` and the ``cpu_baseline`` /
``--impl reference`` legs of ``bench.py`` -- as the *checker* or the *timed CPU baseline*, never
as part of the product path.  The product (``cim_quantization_b200``) never imports it and has
no CPU fallback.

Parity status: PINNED.  ``tests/golden/*.npz`` were produced by importing the reference's own
Python modules (``tests/golden/make_golden.py``, run in the build container where
``/root/reference`` exists) and ``tests/test_oracle_golden.py`` checks this restatement
against them: integer codes / partial sums / ADC codes bit-exactly, floating-point outputs and
gradients to 1e-5.

Arithmetic conventions
----------------------
* every elementwise fp32 operation is done in ``np.float32`` one IEEE operation at a time, in
  the same order as the reference (``x / s``, ``clamp``, ``rint`` = round-half-to-even, ``* s``);
* integer quantities (codes, digit planes, partial sums) are held as integers; the reference
  holds them as fp32/fp16 tensors whose values are exact small integers (lsq.py:169-177);
* "exact recovery"  is assumed: the reference recovers codes as ``x_q / s``
  (lsq.py:97-98) which is exact only for step sizes where ``fl(fl(k*s)/s) == k`` for every code
  ``k``; ``snap_step_size`` finds such step sizes and the golden cases use them.
"""
from __future__ import annotations

import math
from dataclasses import dataclass

import numpy as np

F32 = np.float32


# --------------------------------------------------------------------------------------
# configuration
# --------------------------------------------------------------------------------------
@dataclass(frozen=True)
class CimConfig:
    """Static description of one CiM conv layer (``_quan_base.py:174-237``)."""

    in_channels: int
    out_channels: int
    kernel: int
    stride: int = 1
    padding: int = 0
    nbits_w: int = 3
    nbits_a: int = 3
    nbits_alpha: int = 8
    wbitslice: int = 1
    abitslice: int = 1
    xbar: int = 128
    adcbits: float = 1.5

    @property
    def flatdim(self) -> int:  # _quan_base.py:199
        return self.in_channels * self.kernel * self.kernel

    @property
    def num_xbars(self) -> int:  # _quan_base.py:201
        return int(math.ceil(self.flatdim / self.xbar))

    @property
    def nsw(self) -> int:  # _quan_base.py:203
        return int(self.nbits_w / self.wbitslice)

    @property
    def nsa(self) -> int:  # _quan_base.py:204
        return int(self.nbits_a / self.abitslice)

    @property
    def qn_w(self) -> int:  # lsq.py:523
        return -(2 ** (self.nbits_w - 1))

    @property
    def qp_w(self) -> int:  # lsq.py:524
        return 2 ** (self.nbits_w - 1) - 1

    @property
    def qp_a(self) -> int:  # lsq.py:537-538 (activations are always clamped to [0, Qp_a])
        return 2 ** self.nbits_a - 1

    @property
    def adc_range(self) -> tuple[float, float]:  # lsq.py:125-129, 525-531
        if self.adcbits in (1, 1.5):
            return -1, 1
        return -(2 ** (self.adcbits - 1)), 2 ** (self.adcbits - 1) - 1

    @property
    def has_alpha_cim(self) -> bool:  # _quan_base.py:218-224
        return self.adcbits in (1, 1.5)

    def out_hw(self, in_hw: int) -> int:  # lsq.py:123 (square images / kernels, dilation 1)
        return int((in_hw - self.kernel + 2 * self.padding) / self.stride + 1)

    def binary_mask(self) -> np.ndarray:
        """Shift-and-add weights ``[NSW, NSA]`` (``_quan_base.py:207-214``).

        The reference stores the mask as **int8** (``_quan_base.py:214``), so ``2^(abs*i+wbs*j)``
        wraps: 128 -> -128 and every larger power of two -> 0.  This only bites the 8-bit first
        layer (8x8 slice pairs) but it is the reference's arithmetic, so it is reproduced.
        """
        m = np.ones((self.nsw, self.nsa), dtype=np.float32)
        for i in range(self.nsa):
            for j in range(self.nsw):
                v = ((2 ** self.abitslice) ** i) * ((2 ** self.wbitslice) ** j)
                m[j, i] = ((v + 128) % 256) - 128  # two's-complement wrap of the int8 cast
        return m


# --------------------------------------------------------------------------------------
# STE primitives (values only)
# --------------------------------------------------------------------------------------
def grad_scale_value(alpha, g: float) -> np.ndarray:
    """Forward value of ``grad_scale`` (lsq.py:23-26): ``(a - a*g) + a*g`` in fp32.

    The scale ``g`` is a Python double in the reference; ``tensor * python_float`` keeps the
    tensor dtype, i.e. the scalar is rounded to fp32 first.
    """
    a = np.asarray(alpha, dtype=F32)
    ag = a * F32(g)
    return (a - ag) + ag


def lsq_codes(x: np.ndarray, s, qn: int, qp: int) -> np.ndarray:
    """Integer codes of the LSQ fake-quantiser (lsq.py:549, 555): ``rint(clamp(x / s))``."""
    u = np.asarray(x, dtype=F32) / F32(s)
    return np.rint(np.clip(u, F32(qn), F32(qp))).astype(np.int32)


def lsq_fake_quant(x: np.ndarray, s, qn: int, qp: int) -> np.ndarray:
    """Forward value of ``round_pass(clamp(x/s)) * s`` (lsq.py:29-32, 549)."""
    u = np.clip(np.asarray(x, dtype=F32) / F32(s), F32(qn), F32(qp))
    r = np.rint(u)
    # round_pass value is (r - u) + u which is exact (Sterbenz), i.e. r
    return (((r - u) + u) * F32(s)).astype(F32)


def recovery_is_exact(s, qn: int, qp: int) -> bool:
    """True when ``fl(fl(k*s)/s) == k`` for every code (lsq.py:97-98)."""
    k = np.arange(qn, qp + 1, dtype=F32)
    return bool(np.all((k * F32(s)) / F32(s) == k))


def snap_step_size(s, qn: int, qp: int) -> np.float32:
    """Nudge ``s`` upwards by ulps until code recovery by division is exact."""
    s = F32(s)
    for _ in range(4096):
        if recovery_is_exact(s, qn, qp):
            return s
        s = np.nextafter(s, F32(np.inf), dtype=F32)
    raise RuntimeError("no exact-recovery step size found")


# --------------------------------------------------------------------------------------
# bit slicing (lsq.py:438-509)
# --------------------------------------------------------------------------------------
def slice_unsigned(codes: np.ndarray, bits: int, bit_slice: int) -> np.ndarray:
    """LSB-first digit planes of non-negative codes (``slicing_act``, lsq.py:466-480).

    Returns ``[bits/bit_slice, *codes.shape]`` int32.
    """
    n = int(bits / bit_slice)
    base = 2 ** bit_slice
    c = np.asarray(codes, dtype=np.int64)
    planes = [np.remainder(np.floor_divide(c, base ** i), base) for i in range(n)]
    return np.stack(planes).astype(np.int32)


def slice_signed(codes: np.ndarray, bits: int, bit_slice: int) -> np.ndarray:
    """Sign-magnitude digit planes (``slicing_weights_signed`` / ``slicing_act_signed``,
    lsq.py:438-464, 483-509): positive and negative parts are sliced separately and
    subtracted, so a digit carries the sign of its code."""
    c = np.asarray(codes, dtype=np.int64)
    pos = np.where(c > 0, c, 0)
    neg = np.where(c < 0, -c, 0)
    return slice_unsigned(pos, bits, bit_slice) - slice_unsigned(neg, bits, bit_slice)


# --------------------------------------------------------------------------------------
# im2col / col2im (nn.Unfold / nn.Fold as used at lsq.py:141, 290, 382)
# --------------------------------------------------------------------------------------
def unfold(x: np.ndarray, k: int, pad: int, stride: int) -> np.ndarray:
    """``nn.Unfold(k, padding=pad, stride=stride)(x).transpose(1, 2)`` -> ``[B, L, C*k*k]``.

    Column index is ``c*k*k + ky*k + kx`` (channel-major), output pixels row-major.
    """
    b, c, h, w = x.shape
    oh = (h + 2 * pad - k) // stride + 1
    ow = (w + 2 * pad - k) // stride + 1
    xp = np.zeros((b, c, h + 2 * pad, w + 2 * pad), dtype=x.dtype)
    xp[:, :, pad:pad + h, pad:pad + w] = x
    cols = np.empty((b, c, k, k, oh, ow), dtype=x.dtype)
    for ky in range(k):
        for kx in range(k):
            cols[:, :, ky, kx] = xp[:, :, ky:ky + stride * oh:stride, kx:kx + stride * ow:stride]
    return cols.reshape(b, c * k * k, oh * ow).transpose(0, 2, 1)


def fold(cols: np.ndarray, out_hw: tuple[int, int], k: int, pad: int, stride: int) -> np.ndarray:
    """``nn.Fold(out_hw, k, 1, pad, stride)`` applied to ``cols`` of shape ``[B, C*k*k, L]``."""
    b, ckk, _ = cols.shape
    c = ckk // (k * k)
    h, w = out_hw
    oh = (h + 2 * pad - k) // stride + 1
    ow = (w + 2 * pad - k) // stride + 1
    cols6 = cols.reshape(b, c, k, k, oh, ow)
    xp = np.zeros((b, c, h + 2 * pad, w + 2 * pad), dtype=cols.dtype)
    for ky in range(k):
        for kx in range(k):
            xp[:, :, ky:ky + stride * oh:stride, kx:kx + stride * ow:stride] += cols6[:, :, ky, kx]
    return xp[:, :, pad:pad + h, pad:pad + w]


# --------------------------------------------------------------------------------------
# integer partial sums (lsq.py:141-192)
# --------------------------------------------------------------------------------------
def sliced_operands(cfg: CimConfig, x_codes: np.ndarray, w_codes: np.ndarray):
    """Digit-plane operands of the crossbar GEMMs.

    x_codes ``[B, Cin, H, W]`` ints in ``[0, Qp_a]``; w_codes ``[Cout, Cin, k, k]`` ints.
    Returns ``x_sl [B, NSA, L, F]`` and ``w_sl [NSW, F, Cout]`` (fp32 holding small ints).
    Non-negative activation codes make ``slicing_act_signed`` identical to ``slicing_act``
    (lsq.py:146-149), so the ``signed_act`` flag does not change the forward value.
    """
    x_unf = unfold(np.asarray(x_codes, dtype=np.int32), cfg.kernel, cfg.padding, cfg.stride)
    x_sl = slice_signed(x_unf, cfg.nbits_a, cfg.abitslice).transpose(1, 0, 2, 3)  # [B,NSA,L,F]
    w_unf = np.asarray(w_codes, dtype=np.int32).reshape(cfg.out_channels, -1).T  # [F,Cout], lsq.py:153
    w_sl = slice_signed(w_unf, cfg.nbits_w, cfg.wbitslice)  # [NSW,F,Cout]
    return x_sl.astype(F32), w_sl.astype(F32)


def chunk_bounds(cfg: CimConfig):
    """Row ranges of the crossbars: full ``xbar``-row chunks then the remainder (lsq.py:172-185)."""
    f = cfg.flatdim
    return [(i * cfg.xbar, min((i + 1) * cfg.xbar, f)) for i in range(cfg.num_xbars)]


def integer_psums(cfg: CimConfig, x_codes: np.ndarray, w_codes: np.ndarray) -> np.ndarray:
    """Integer partial sums ``[B, NX, NSW, NSA, L, Cout]`` (``ctx.ps_int``, lsq.py:169-192).

    One fp32 matmul per (crossbar, act slice, weight slice); the products and sums are small
    integers so fp32 is exact.
    """
    x_sl, w_sl = sliced_operands(cfg, x_codes, w_codes)
    b, _, l, _ = x_sl.shape
    out = np.empty((b, cfg.num_xbars, cfg.nsw, cfg.nsa, l, cfg.out_channels), dtype=F32)
    for i, (lo, hi) in enumerate(chunk_bounds(cfg)):
        for j in range(cfg.nsa):
            for k in range(cfg.nsw):
                out[:, i, k, j] = np.matmul(x_sl[:, j, :, lo:hi], w_sl[k, lo:hi, :])
    return np.rint(out).astype(np.int32)


# --------------------------------------------------------------------------------------
# ADC / partial-sum quantisation (lsq.py:195-233)
# --------------------------------------------------------------------------------------
def _scaled_psums(ps_int: np.ndarray, s_w, s_a) -> np.ndarray:
    """fp16 storage then ``* s_w * s_a`` promoted to fp32 (lsq.py:169, 195)."""
    p16 = ps_int.astype(np.float16).astype(F32)
    return (p16 * F32(s_w)) * F32(s_a)


def adc_codes(cfg: CimConfig, ps_int: np.ndarray, s_w, s_a, alpha_q) -> np.ndarray:
    """Integer ADC codes of every partial sum (lsq.py:197-230), same shape as ``ps_int``."""
    qn, qp = cfg.adc_range
    v = _scaled_psums(ps_int, s_w, s_a)
    if cfg.adcbits == 1:
        return np.sign(v).astype(np.int32)
    if cfg.adcbits == 1.5:
        return np.clip(np.rint(v / np.asarray(alpha_q, dtype=F32)), qn, qp).astype(np.int32)
    return np.clip(np.rint(v / (F32(s_w) * F32(s_a))), qn, qp).astype(np.int32)


def cim_forward(cfg: CimConfig, x_codes, w_codes, s_w, s_a, alpha_q, return_internals=False):
    """``get_cim_output_signed.forward`` (lsq.py:92-237) -> ``[B, L, Cout]`` fp32."""
    ps_int = integer_psums(cfg, x_codes, w_codes)
    codes = adc_codes(cfg, ps_int, s_w, s_a, alpha_q)
    if cfg.adcbits in (1, 1.5):
        adc_out = codes.astype(F32) * np.asarray(alpha_q, dtype=F32)  # lsq.py:202, 225
    else:
        adc_out = (codes.astype(F32) * F32(s_w)) * F32(s_a)  # lsq.py:230
    mask = cfg.binary_mask().reshape(1, 1, cfg.nsw, cfg.nsa, 1, 1)
    out = np.sum(adc_out * mask, axis=(1, 2, 3), dtype=F32)  # lsq.py:233
    if return_internals:
        return out, ps_int, codes
    return out


# --------------------------------------------------------------------------------------
# backward of the Function (lsq.py:244-386)
# --------------------------------------------------------------------------------------
def int8_saved_codes(x_codes: np.ndarray) -> np.ndarray:
    """What ``ctx.x_int = x_int.type(torch.int8)`` (lsq.py:99) keeps of the activation codes: two's-complement
    wrap-around, i.e. codes 128..255 of an 8-bit layer come back as -128..-1 in the backward."""
    return ((np.asarray(x_codes, dtype=np.int32) + 128) % 256) - 128


def cim_backward(cfg: CimConfig, grad_out, x_codes, w_codes, s_w, s_a, alpha_q, in_hw, signed_act=False,
                 int8_save=False):
    """``get_cim_output_signed.backward``.

    grad_out ``[B, L, Cout]``.  Returns ``(grad_xq [B,Cin,H,W], grad_wq [Cout,Cin,k,k],
    grad_alpha_q [1,NX,NSW,NSA,1,Cout] or None)``.

    ``int8_save`` + ``signed_act`` reproduce the reference's int8 save of the activation codes (lsq.py:99) for
    layers that slice with ``slicing_act_signed`` (lsq.py:291-292, the 8-bit first conv of a model fed normalised
    images): a code >= 128 re-enters the backward as ``code - 256`` and is sliced as minus the digits of
    ``256 - code``, so its contribution to grad_w differs from the forward's digits (hazard H6).  With the unsigned
    ``slicing_act`` the wrap is invisible (floor-division / remainder recover the same low bits), and codes < 128
    are unaffected either way.  Default ``False``: the arithmetic the forward used (what the CUDA path computes).
    """
    qn, qp = cfg.adc_range
    s_w = F32(s_w)
    s_a = F32(s_a)
    ps_int = integer_psums(cfg, x_codes, w_codes)
    x_sl, w_sl = sliced_operands(cfg, x_codes, w_codes)
    if int8_save and signed_act:
        x_sl, _ = sliced_operands(cfg, int8_saved_codes(x_codes), w_codes)
    w_sl = w_sl * s_w  # lsq.py:252
    x_sl = x_sl * s_a  # lsq.py:295
    if cfg.adcbits in (1, 1.5):
        ps = _scaled_psums(ps_int, s_w, s_a) / np.asarray(alpha_q, dtype=F32)  # lsq.py:257-264
    else:
        ps = ps_int.astype(np.float16).astype(F32)  # lsq.py:267
    b, nx, nsw, nsa, l, cout = ps.shape
    mask = cfg.binary_mask().reshape(1, 1, nsw, nsa, 1, 1)
    g = np.broadcast_to(np.asarray(grad_out, dtype=F32)[:, None, None, None], ps.shape) * mask
    g_after_adc = g.copy()  # lsq.py:307
    greater = ps >= F32(qp + 1e-5)  # lsq.py:310
    lesser = ps <= F32(qn - 1e-5)  # lsq.py:311
    g = np.where(greater | lesser, F32(0), g)  # lsq.py:313

    grad_alpha = None
    if cfg.adcbits == 1:  # lsq.py:321-325
        ga = np.sign(ps) * F32(1.0 / math.sqrt(ps.size * qp))
        grad_alpha = np.sum(ga * g_after_adc, axis=(0, 4), keepdims=True, dtype=F32)
    elif cfg.adcbits == 1.5:  # lsq.py:326-332
        ga = np.rint(ps)
        ga = np.where(greater, F32(qp), ga)
        ga = np.where(lesser, F32(qn), ga)
        ga = ga * F32(1.0 / math.sqrt(ps.size * qp))
        grad_alpha = np.sum(ga * g_after_adc, axis=(0, 4), keepdims=True, dtype=F32)

    f = cfg.flatdim
    grad_in = np.zeros((b, nsw, nsa, l, f), dtype=F32)
    grad_w = np.zeros((nsw, nsa, f, cout), dtype=F32)
    for i, (lo, hi) in enumerate(chunk_bounds(cfg)):  # lsq.py:338-356
        for j in range(nsa):
            for k in range(nsw):
                gt = g[:, i, k, j]  # [B,L,Cout]
                grad_in[:, k, j, :, lo:hi] = np.matmul(gt, w_sl[k, lo:hi, :].T)
                grad_w[k, j, lo:hi, :] = np.matmul(x_sl[:, j, :, lo:hi].transpose(0, 2, 1), gt).sum(0)

    gw = grad_w.sum(axis=1)  # lsq.py:362
    for k in range(1, nsw):
        gw[k] = gw[k] / F32((2 ** cfg.wbitslice) ** k)  # lsq.py:363-364
    gw = gw.mean(axis=0, dtype=F32)  # lsq.py:366
    gw = gw.T.reshape(cfg.out_channels, cfg.in_channels, cfg.kernel, cfg.kernel)  # lsq.py:369

    gi = grad_in.sum(axis=1)  # lsq.py:372
    for j in range(1, nsa):
        gi[:, j] = gi[:, j] / F32((2 ** cfg.abitslice) ** j)  # lsq.py:373-374
    gi = gi.mean(axis=1, dtype=F32)  # lsq.py:376   [B,L,F]
    gx = fold(np.ascontiguousarray(gi.transpose(0, 2, 1)), (in_hw, in_hw), cfg.kernel,
              cfg.padding, cfg.stride)  # lsq.py:380-382
    return gx.astype(F32), gw.astype(F32), grad_alpha


# --------------------------------------------------------------------------------------
# module level: Conv2dLSQCiM.forward and its full backward (lsq.py:522-588)
# --------------------------------------------------------------------------------------
def init_step_size(t: np.ndarray, qp: int) -> np.float32:
    """``2 * mean|t| / sqrt(Qp)`` (lsq.py:540-541)."""
    return F32(F32(2) * np.mean(np.abs(np.asarray(t, dtype=F32)), dtype=F32) / F32(math.sqrt(qp)))


def analog_partial_sums(cfg: CimConfig, x_codes, w_codes, s_w, s_a) -> np.ndarray:
    """``get_analog_partial_sums_signed`` (lsq.py:35-87): un-quantised fp32 psums."""
    ps_int = integer_psums(cfg, x_codes, w_codes).astype(F32)
    return (ps_int * F32(s_w)) * F32(s_a)


def init_alpha_cim(cfg: CimConfig, x_codes, w_codes, s_w, s_a) -> np.ndarray:
    """Data-dependent initial value of ``alpha_cim`` (lsq.py:557-563)."""
    _, qp = cfg.adc_range
    ps = analog_partial_sums(cfg, x_codes, w_codes, s_w, s_a)
    t = F32(2.0) * np.mean(np.abs(ps), axis=(0, 4), keepdims=True, dtype=F32) / F32(math.sqrt(qp))
    t = np.where(t == 0, F32(1.0) * F32(s_w) * F32(s_a), t)
    return t.astype(F32)


def quantize_alpha(cfg: CimConfig, alpha: np.ndarray):
    """8-bit range quantiser of ``alpha_cim`` (lsq.py:566-571).

    Returns ``(alpha_q, aux)``; ``aux`` holds what the backward needs.
    """
    qp_al = 2 ** cfg.nbits_alpha - 1
    qn_al = 1
    a = np.asarray(alpha, dtype=F32)
    scale = F32((a.max() - a.min()) / F32(qp_al - qn_al))
    t = a / scale
    r = np.rint(t)
    rp = (r - t) + t  # round_pass value
    c = np.clip(rp, F32(qn_al), F32(qp_al))
    aq = c * scale
    return aq.astype(F32), dict(scale=scale, t=t, rp=rp, c=c, qn=qn_al, qp=qp_al)


def quantize_alpha_backward(alpha: np.ndarray, grad_aq: np.ndarray, aux) -> np.ndarray:
    """Autograd of ``quantize_alpha`` w.r.t. ``alpha`` (through ``max``/``min``, the STE of
    ``round_pass`` and the inclusive clamp mask)."""
    a = np.asarray(alpha, dtype=F32)
    g = np.asarray(grad_aq, dtype=F32)
    scale = aux["scale"]
    inside = (aux["rp"] >= aux["qn"]) & (aux["rp"] <= aux["qp"])
    g_t = np.where(inside, g * scale, F32(0))  # d/dt through clamp (STE through round)
    g_alpha = g_t / scale
    g_scale = np.sum(g * aux["c"], dtype=F32) + np.sum(g_t * (-a / (scale * scale)), dtype=F32)
    g_range = g_scale / F32(aux["qp"] - aux["qn"])
    mx = a == a.max()
    mn = a == a.min()
    g_alpha = g_alpha + np.where(mx, g_range / F32(mx.sum()), F32(0))
    g_alpha = g_alpha - np.where(mn, g_range / F32(mn.sum()), F32(0))
    return g_alpha.astype(F32)


def lsq_backward(x: np.ndarray, s, qn: int, qp: int, g: float, grad_xq: np.ndarray):
    """Autograd of ``round_pass(clamp(x/s, qn, qp)) * s`` with ``s = grad_scale(alpha, g)``
    (lsq.py:547-555).  Returns ``(grad_x, grad_alpha)``.

    ``grad_x = grad_xq * 1[qn <= x/s <= qp]`` (torch's clamp mask is inclusive) and
    ``grad_alpha = g * sum(grad_xq * (q - (x/s) * 1[...]))``.
    """
    s = F32(s)
    u = np.asarray(x, dtype=F32) / s
    inside = (u >= F32(qn)) & (u <= F32(qp))
    q = np.rint(np.clip(u, F32(qn), F32(qp)))
    gq = np.asarray(grad_xq, dtype=F32)
    grad_x = np.where(inside, (gq * s) / s, F32(0)).astype(F32)
    # fp64 accumulation: the two terms cancel
    d = q.astype(np.float64) - np.where(inside, u, F32(0)).astype(np.float64)
    grad_alpha = float(g) * float(np.sum(gq.astype(np.float64) * d))
    return grad_x, grad_alpha


def module_forward_backward(cfg: CimConfig, x, weight, alpha_act, alpha_weight, alpha_cim, grad_y):
    """One training-mode forward+backward of ``Conv2dLSQCiM`` with initialised step sizes.

    x ``[B,Cin,H,W]``, grad_y ``[B,Cout,OH,OW]`` (gradient w.r.t. the module output).
    Returns a dict with the output and all gradients (lsq.py:522-588 + autograd).
    """
    x = np.asarray(x, dtype=F32)
    weight = np.asarray(weight, dtype=F32)
    ga = 1.0 / math.sqrt(x.size * cfg.qp_a)  # lsq.py:547
    gw = 1.0 / math.sqrt(weight.size * cfg.qp_w)  # lsq.py:553
    s_a = grad_scale_value(alpha_act, ga).reshape(())
    s_w = grad_scale_value(alpha_weight, gw).reshape(())
    x_codes = lsq_codes(x, s_a, 0, cfg.qp_a)
    w_codes = lsq_codes(weight, s_w, cfg.qn_w, cfg.qp_w)
    alpha_q, aux = (None, None)
    if cfg.has_alpha_cim:
        alpha_q, aux = quantize_alpha(cfg, alpha_cim)
    out, ps_int, codes = cim_forward(cfg, x_codes, w_codes, s_w, s_a, alpha_q, return_internals=True)
    b = x.shape[0]
    oh = cfg.out_hw(x.shape[-1])
    y = out.transpose(0, 2, 1).reshape(b, cfg.out_channels, oh, oh)  # lsq.py:581
    go = np.asarray(grad_y, dtype=F32).reshape(b, cfg.out_channels, oh * oh).transpose(0, 2, 1)
    gxq, gwq, gaq = cim_backward(cfg, go, x_codes, w_codes, s_w, s_a, alpha_q, x.shape[-1])
    grad_x, g_alpha_act = lsq_backward(x, s_a, 0, cfg.qp_a, ga, gxq)
    grad_w, g_alpha_w = lsq_backward(weight, s_w, cfg.qn_w, cfg.qp_w, gw, gwq)
    res = dict(y=y, x_codes=x_codes, w_codes=w_codes, ps_int=ps_int, adc_codes=codes, s_a=s_a, s_w=s_w,
               alpha_q=alpha_q, grad_xq=gxq, grad_wq=gwq, grad_alpha_q=gaq, grad_x=grad_x,
               grad_weight=grad_w, grad_alpha_act=F32(g_alpha_act), grad_alpha_weight=F32(g_alpha_w))
    if cfg.has_alpha_cim:
        res["grad_alpha_cim"] = quantize_alpha_backward(alpha_cim, gaq, aux)
    return res


def stochastic_code_expectation(cfg: CimConfig, ps_int, s_w, s_a, alpha_q, num_iter: int = 50, sharpness: float = 0.01):
    """Mean and variance of the sampled ternary code of lsq.py:205-220 for every partial sum.

    The reference draws ``n1 ~ Binomial(50, sigmoid((v - alpha/2)/0.01))`` and ``n2 ~ Binomial(50, sigmoid((v +
    alpha/2)/0.01))`` (``ceil(p - U)`` is 1 iff ``U < p``) and outputs ``clamp(round(n1/50 + n2/50 - 1), -1, 1)``:
    +1 iff ``n1 + n2 > 75``, -1 iff ``n1 + n2 < 25`` (round half to even).  The distribution of ``n1 + n2`` is the
    convolution of the two binomials.  Returns ``(mean, var)`` with the shape of ``ps_int``
    ``[B,NX,NSW,NSA,L,Cout]``."""
    from math import comb
    v = _scaled_psums(ps_int, s_w, s_a)
    a = np.asarray(alpha_q, dtype=F32).reshape(1, cfg.num_xbars, cfg.nsw, cfg.nsa, 1, cfg.out_channels)
    with np.errstate(over="ignore"):
        p1 = 1.0 / (1.0 + np.exp(-((v - F32(0.5) * a) / F32(sharpness)).astype(np.float64)))
        p2 = 1.0 / (1.0 + np.exp(-((v + F32(0.5) * a) / F32(sharpness)).astype(np.float64)))
    k = np.arange(num_iter + 1)
    binom = np.array([comb(num_iter, int(i)) for i in k], dtype=np.float64)

    def pmf(p):  # [..., num_iter+1]
        p = np.clip(p, 0.0, 1.0)[..., None]
        return binom * p ** k * (1.0 - p) ** (num_iter - k)

    flat1, flat2 = pmf(p1).reshape(-1, num_iter + 1), pmf(p2).reshape(-1, num_iter + 1)
    # P(n1 + n2 = t): row-wise convolution
    conv = np.zeros((flat1.shape[0], 2 * num_iter + 1))
    for i in range(num_iter + 1):
        conv[:, i:i + num_iter + 1] += flat1[:, i:i + 1] * flat2
    hi = int(round(1.5 * num_iter))  # 75
    lo = int(round(0.5 * num_iter))  # 25
    p_plus, p_minus = conv[:, hi + 1:].sum(1), conv[:, :lo].sum(1)
    mean = (p_plus - p_minus).reshape(v.shape)
    var = (p_plus + p_minus).reshape(v.shape) - mean ** 2
    return mean, var
