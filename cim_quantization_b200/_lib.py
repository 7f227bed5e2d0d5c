"""ctypes binding of ``libcimq.so`` (C ABI declared in ``include/cimq.h``).

There is no CPU fallback: importing this module without the built library raises, and every
compute wrapper requires CUDA tensors.  All wrappers enqueue on the current torch CUDA stream and
never synchronise, so they can be captured in CUDA graphs.
"""
from __future__ import annotations

import ctypes as C
import os
from dataclasses import dataclass

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libcimq.so")

ADC_MULTIBIT, ADC_BINARY, ADC_TERNARY = 0, 1, 2
FLAG_FORCE_SIMT = 1
FLAG_DETERMINISTIC = 2  # backward: fixed-order fold of grad_x instead of fp32 reductions in the dgrad epilogue
FLAG_V2 = 4  # second-generation kernels + uint8 ADC-state planes (include/cimq.h); set by conv_forward(v2=True)
_deterministic = False


def set_deterministic(enabled: bool) -> None:
    """Make every conv_backward bit-reproducible run to run (also implied by
    torch.use_deterministic_algorithms(True)); costs the separate col2im pass."""
    global _deterministic
    _deterministic = bool(enabled)


def _backward_flags(flags: int) -> int:
    if _deterministic or torch.are_deterministic_algorithms_enabled():
        flags |= FLAG_DETERMINISTIC
    return flags



class CimqLayer(C.Structure):
    """``cimq_layer_t``."""
    _fields_ = [(n, C.c_int32) for n in (
        "batch", "in_channels", "in_hw", "out_channels", "kernel", "stride", "padding",
        "nbits_a", "abitslice", "nbits_w", "wbitslice", "xbar", "adc_mode", "adc_qn", "adc_qp")]


class CimqInfo(C.Structure):
    """``cimq_info_t``."""
    _fields_ = ([(n, C.c_int32) for n in ("out_hw", "L", "M", "F", "NX", "NSW", "NSA", "pairs", "state_words",
                                          "tc_forward", "tc_backward", "tc_v2")] +
                [(n, C.c_int64) for n in ("state_bytes", "table_bytes", "wdigits_bytes", "wtiles_bytes",
                                          "bwd_workspace_bytes", "psum_count", "state_v2_bytes")])


EXPORTS = {
    # name: (restype, argtypes)
    "cimq_version": (C.c_int, []),
    "cimq_last_error": (C.c_char_p, []),
    "cimq_layer_info": (C.c_int, [C.POINTER(CimqLayer), C.POINTER(CimqInfo)]),
    "cimq_step_sizes": (C.c_int, [C.c_void_p, C.c_void_p, C.c_float, C.c_float, C.c_void_p, C.c_void_p]),
    "cimq_lsq_quantize": (C.c_int, [C.c_void_p, C.c_int64, C.c_void_p, C.c_int32, C.c_int32, C.c_void_p,
                                    C.c_void_p]),
    "cimq_codes_from_fakequant": (C.c_int, [C.c_void_p, C.c_int64, C.c_void_p, C.c_int32, C.c_int32, C.c_void_p,
                                            C.c_void_p]),
    "cimq_lsq_fakequant": (C.c_int, [C.c_void_p, C.c_int64, C.c_void_p, C.c_int32, C.c_int32, C.c_int32,
                                     C.c_void_p, C.c_void_p]),
    "cimq_lsq_backward_workspace_bytes": (C.c_int64, [C.c_int64]),
    "cimq_lsq_backward": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p, C.c_int32, C.c_int32,
                                    C.c_float, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    "cimq_alpha_quantize": (C.c_int, [C.c_void_p, C.c_int64, C.c_int32, C.c_int32, C.c_void_p, C.c_void_p, C.c_void_p]),
    "cimq_alpha_quantize_backward": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int64, C.c_int32, C.c_int32, C.c_void_p,
                                               C.c_void_p, C.c_void_p]),
    "cimq_adc_table": (C.c_int, [C.POINTER(CimqLayer), C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p,
                                 C.c_void_p, C.c_void_p]),
    "cimq_adc_table2": (C.c_int, [C.POINTER(CimqLayer), C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p,
                                  C.c_void_p, C.c_void_p]),
    "cimq_layer_prepare": (C.c_int, [C.POINTER(CimqLayer), C.c_void_p, C.c_void_p, C.c_void_p, C.c_float, C.c_float,
                                     C.c_void_p, C.c_int32, C.c_int32, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p,
                                     C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    "cimq_weight_prepare": (C.c_int, [C.POINTER(CimqLayer), C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    "cimq_conv_forward": (C.c_int, [C.POINTER(CimqLayer), C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p,
                                    C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint32, C.c_void_p]),
    "cimq_conv_forward_stochastic": (C.c_int, [C.POINTER(CimqLayer), C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p,
                                               C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint64, C.c_void_p]),
    "cimq_conv_backward": (C.c_int, [C.POINTER(CimqLayer), C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p,
                                     C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p,
                                     C.c_void_p, C.c_uint32, C.c_void_p]),
    "cimq_conv_psums": (C.c_int, [C.POINTER(CimqLayer), C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    "cimq_conv_psum_abs_sums": (C.c_int, [C.POINTER(CimqLayer), C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    "cimq_bn_workspace_bytes": (C.c_int64, [C.c_int32, C.c_int32]),
    "cimq_bn_forward": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int32,
                                  C.c_float, C.c_float, C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.c_void_p,
                                  C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    "cimq_bn_forward_quant": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int32,
                                        C.c_float, C.c_float, C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.c_void_p,
                                        C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_float, C.c_int32, C.c_void_p,
                                        C.c_void_p]),
    "cimq_bn_backward": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int32,
                                   C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.c_void_p, C.c_void_p, C.c_void_p,
                                   C.c_void_p, C.c_void_p, C.c_void_p]),
}

_lib = None
launch_counter = 0  # kernels launched by this binding since the caller last reset it (bench.py: gpu_launches)


def _count(n: int):
    global launch_counter
    launch_counter += n


def load():
    """Load ``libcimq.so`` (built in-tree by ``make`` / ``__graft_entry__.build()``).  Fails loudly."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError(f"{LIB_PATH} not found: build it with `make` (or __graft_entry__.build()); "
                           "cim_quantization_b200 has no CPU / PyTorch fallback")
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in EXPORTS.items():
        fn = getattr(lib, name)  # AttributeError if the library lacks a declared symbol
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


def _check(rc: int):
    if rc != 0:
        raise RuntimeError("libcimq: " + load().cimq_last_error().decode())


def _ptr(t):
    if t is None:
        return None
    if not t.is_cuda:
        raise RuntimeError("libcimq operates on CUDA tensors only (no CPU fallback)")
    if not t.is_contiguous():
        raise RuntimeError("libcimq expects contiguous tensors")
    return C.c_void_p(t.data_ptr())


def _stream():
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def _on_tensor_device(fn):
    """Run a wrapper inside the CUDA device context of its tensor arguments (they must share one device): kernels
    are launched on that device's current stream and outputs allocated there, whatever the caller's current device
    is -- a model moved with ``.cuda(1)`` works without ``torch.cuda.set_device``."""
    import functools

    @functools.wraps(fn)
    def wrapper(*args, **kwargs):
        dev = None
        for a in list(args) + list(kwargs.values()):
            if isinstance(a, torch.Tensor) and a.is_cuda:
                if dev is None:
                    dev = a.device
                elif a.device != dev:
                    raise RuntimeError(f"libcimq: tensor arguments on different devices ({dev} and {a.device})")
        if dev is None or dev.index == torch.cuda.current_device():
            return fn(*args, **kwargs)
        with torch.cuda.device(dev):
            return fn(*args, **kwargs)

    return wrapper


def adc_mode_of(adcbits) -> tuple[int, int, int]:
    """Map the reference's ``adcbits`` (a float from protobuf) to (mode, qn, qp) -- lsq.py:125-129."""
    if adcbits == 1:
        return ADC_BINARY, -1, 1
    if adcbits == 1.5:
        return ADC_TERNARY, -1, 1
    if adcbits <= 0:
        raise ValueError("adcbits == 0 (no ADC) is the plain F.conv2d path, not a CiM layer")
    qp = int(2 ** (adcbits - 1) - 1)
    qn = int(-(2 ** (adcbits - 1)))
    return ADC_MULTIBIT, qn, qp


@dataclass(frozen=True)
class LayerSpec:
    """Hashable description of a CiM conv layer instance (one batch size / image size)."""
    batch: int
    in_channels: int
    in_hw: int
    out_channels: int
    kernel: int
    stride: int
    padding: int
    nbits_a: int
    abitslice: int
    nbits_w: int
    wbitslice: int
    xbar: int
    adcbits: float

    def c_layer(self) -> CimqLayer:
        mode, qn, qp = adc_mode_of(self.adcbits)
        return CimqLayer(self.batch, self.in_channels, self.in_hw, self.out_channels, self.kernel, self.stride,
                         self.padding, self.nbits_a, self.abitslice, self.nbits_w, self.wbitslice, self.xbar,
                         mode, qn, qp)


_info_cache: dict = {}


def layer_info(spec: LayerSpec) -> CimqInfo:
    info = _info_cache.get(spec)
    if info is None:
        info = CimqInfo()
        layer = spec.c_layer()
        _check(load().cimq_layer_info(C.byref(layer), C.byref(info)))
        _info_cache[spec] = info
    return info


# ---- thin wrappers (tensors in, tensors out; all allocation happens here, in torch) ------------------
@_on_tensor_device
def step_sizes(alpha_act, alpha_weight, ga: float, gw: float):
    s = torch.empty(2, dtype=torch.float32, device=alpha_act.device)
    _check(load().cimq_step_sizes(_ptr(alpha_act), _ptr(alpha_weight), ga, gw, _ptr(s), _stream()))
    _count(1)
    return s


@_on_tensor_device
def lsq_quantize(x, s_elem, qn: int, qp: int, from_fakequant: bool = False):
    """x fp32 (contiguous) -> one-byte codes (uint8 if qn >= 0 else int8); s_elem: 1-element CUDA tensor."""
    codes = torch.empty(x.shape, dtype=torch.uint8 if qn >= 0 else torch.int8, device=x.device)
    fn = load().cimq_codes_from_fakequant if from_fakequant else load().cimq_lsq_quantize
    _check(fn(_ptr(x), x.numel(), _ptr(s_elem), qn, qp, _ptr(codes), _stream()))
    _count(1 if x.numel() % 16 == 0 else 2)
    return codes


@_on_tensor_device
def lsq_fakequant(x, s_elem, qn: int, qp: int, rescale: bool):
    y = torch.empty_like(x)
    _check(load().cimq_lsq_fakequant(_ptr(x), x.numel(), _ptr(s_elem), qn, qp, int(rescale), _ptr(y), _stream()))
    _count(1)
    return y


@_on_tensor_device
def lsq_backward(grad_xq, x, s_elem, qn: int, qp: int, g: float):
    gx = torch.empty_like(x)
    galpha = torch.empty(1, dtype=torch.float32, device=x.device)
    ws = torch.empty(load().cimq_lsq_backward_workspace_bytes(x.numel()), dtype=torch.uint8, device=x.device)
    _check(load().cimq_lsq_backward(_ptr(grad_xq), _ptr(x), x.numel(), _ptr(s_elem), qn, qp, g, _ptr(gx),
                                    _ptr(galpha), _ptr(ws), _stream()))
    _count(2)
    return gx, galpha


@_on_tensor_device
def alpha_quantize(alpha, qn: int, qp: int):
    aq = torch.empty_like(alpha)
    aux = torch.empty(8, dtype=torch.float32, device=alpha.device)
    _check(load().cimq_alpha_quantize(_ptr(alpha), alpha.numel(), qn, qp, _ptr(aq), _ptr(aux), _stream()))
    _count(1)
    return aq, aux


@_on_tensor_device
def alpha_quantize_backward(alpha, grad_aq, qn: int, qp: int, aux):
    ga = torch.empty_like(alpha)
    _check(load().cimq_alpha_quantize_backward(_ptr(alpha), _ptr(grad_aq), alpha.numel(), qn, qp, _ptr(aux), _ptr(ga),
                                               _stream()))
    _count(1)
    return ga


@_on_tensor_device
def adc_table(spec: LayerSpec, s, alpha_q, binary_mask, status=None, alpha_scale=None):
    """ADC thresholds / amplitudes.  ``alpha_scale`` (1-element CUDA tensor: the step of the alpha quantiser,
    ``aux[0]`` of :func:`alpha_quantize`) additionally builds the constants of the v2 kernels (always built for the
    multi-bit ADC, which has no alpha)."""
    info = layer_info(spec)
    table = torch.empty(info.table_bytes, dtype=torch.uint8, device=s.device)
    layer = spec.c_layer()
    _check(load().cimq_adc_table2(C.byref(layer), _ptr(s), _ptr(alpha_q), _ptr(alpha_scale), _ptr(binary_mask),
                                  _ptr(table), _ptr(status), _stream()))
    _count(2 if (info.tc_v2 and (alpha_scale is not None or alpha_q is None)) else 1)
    return table


def v2_usable(spec: LayerSpec, has_alpha: bool, alpha_scale, flags: int = 0) -> bool:
    """Whether a step of this layer can run on the v2 kernels: covered shape, not forced onto the CUDA-core
    kernels, and -- for the binary / ternary ADC -- the alpha quantiser's step at hand."""
    if flags & FLAG_FORCE_SIMT or os.environ.get("CIMQ_DISABLE_V2"):
        return False
    return bool(layer_info(spec).tc_v2) and (not has_alpha or alpha_scale is not None)


@_on_tensor_device
def layer_prepare(spec: LayerSpec, weight, alpha_act, alpha_weight, ga: float, gw: float, alpha_cim, aq_qn: int,
                  aq_qp: int, binary_mask, status=None):
    """One launch for everything that depends only on the parameters (v2 layers): returns
    ``(s, wcodes, alpha_q, aux, table, wtiles)``; ``alpha_q`` / ``aux`` are None for the multi-bit ADC."""
    info = layer_info(spec)
    dev = weight.device
    s = torch.empty(2, dtype=torch.float32, device=dev)
    wcodes = torch.empty((spec.out_channels, info.F), dtype=torch.int8, device=dev)
    alpha_q = aux = None
    if alpha_cim is not None:
        alpha_q = torch.empty_like(alpha_cim)
        aux = torch.empty(8, dtype=torch.float32, device=dev)
    table = torch.empty(info.table_bytes, dtype=torch.uint8, device=dev)
    wtiles = torch.empty(info.wtiles_bytes, dtype=torch.uint8, device=dev)
    layer = spec.c_layer()
    _check(load().cimq_layer_prepare(C.byref(layer), _ptr(weight), _ptr(alpha_act), _ptr(alpha_weight), ga, gw,
                                     _ptr(alpha_cim), aq_qn, aq_qp, _ptr(binary_mask), _ptr(s), _ptr(wcodes),
                                     _ptr(alpha_q), _ptr(aux), _ptr(table), _ptr(wtiles), _ptr(status), _stream()))
    _count(1)
    return s, wcodes, alpha_q, aux, table, wtiles


@_on_tensor_device
def weight_prepare(spec: LayerSpec, wcodes, want_digits=True, want_tiles=True):
    info = layer_info(spec)
    dev = wcodes.device
    wdigits = torch.empty(info.wdigits_bytes // 4, dtype=torch.float32, device=dev) if want_digits else None
    wtiles = (torch.empty(info.wtiles_bytes, dtype=torch.uint8, device=dev)
              if (want_tiles and info.wtiles_bytes > 0) else None)
    layer = spec.c_layer()
    _check(load().cimq_weight_prepare(C.byref(layer), _ptr(wcodes), _ptr(wdigits), _ptr(wtiles), _stream()))
    _count((1 if wdigits is not None else 0) + (2 if wtiles is not None else 0))
    return wdigits, wtiles


@_on_tensor_device
def conv_forward(spec: LayerSpec, xcodes, wcodes, wtiles, table, s, binary_mask, save_state: bool, flags: int = 0):
    """``flags & FLAG_V2``: v2 kernel; the state then is a uint8 tensor (v2 planes) instead of int32 words, which is
    how :func:`conv_backward` tells the two formats apart."""
    info = layer_info(spec)
    dev = xcodes.device
    out = torch.empty((spec.batch, spec.out_channels, info.L), dtype=torch.float32, device=dev)
    if not save_state:
        state = None
    elif flags & FLAG_V2:
        state = torch.empty(info.state_v2_bytes, dtype=torch.uint8, device=dev)
    else:
        state = torch.empty(info.state_bytes // 4, dtype=torch.int32, device=dev)
    layer = spec.c_layer()
    _check(load().cimq_conv_forward(C.byref(layer), _ptr(xcodes), _ptr(wcodes), _ptr(wtiles), _ptr(table), _ptr(s),
                                    _ptr(binary_mask), _ptr(out), _ptr(state), flags, _stream()))
    _count(1)
    return out, state


@_on_tensor_device
def conv_forward_stochastic(spec: LayerSpec, xcodes, wcodes, table, s, alpha_q, seed: int, save_state: bool):
    """Forward with the stochastic near-ADC-less read-out (lsq.py:205-220); CUDA-core kernel."""
    info = layer_info(spec)
    dev = xcodes.device
    out = torch.empty((spec.batch, spec.out_channels, info.L), dtype=torch.float32, device=dev)
    state = torch.empty(info.state_bytes // 4, dtype=torch.int32, device=dev) if save_state else None
    layer = spec.c_layer()
    _check(load().cimq_conv_forward_stochastic(C.byref(layer), _ptr(xcodes), _ptr(wcodes), _ptr(table), _ptr(s),
                                               _ptr(alpha_q.contiguous()), _ptr(out), _ptr(state),
                                               C.c_uint64(seed & 0xFFFFFFFFFFFFFFFF), _stream()))
    _count(1)
    return out, state


@_on_tensor_device
def conv_backward(spec: LayerSpec, grad_out, xcodes, wdigits, wtiles, state, s, binary_mask, need_alpha: bool,
                  need_input: bool = True, flags: int = 0, need_weight: bool = True):
    info = layer_info(spec)
    dev = grad_out.device
    gxq = (torch.empty((spec.batch, spec.in_channels, spec.in_hw, spec.in_hw), dtype=torch.float32, device=dev)
           if need_input else None)
    gwq = torch.empty((spec.out_channels, info.F), dtype=torch.float32, device=dev) if need_weight else None
    galpha = (torch.empty((1, info.NX, info.NSW, info.NSA, 1, spec.out_channels), dtype=torch.float32, device=dev)
              if need_alpha else None)
    ws = torch.empty(info.bwd_workspace_bytes, dtype=torch.uint8, device=dev)
    layer = spec.c_layer()
    if state.dtype == torch.uint8:  # v2 state planes
        flags |= FLAG_V2
    _check(load().cimq_conv_backward(C.byref(layer), _ptr(grad_out), _ptr(xcodes), _ptr(wdigits), _ptr(wtiles),
                                     _ptr(state), _ptr(s), _ptr(binary_mask), _ptr(gxq), _ptr(gwq), _ptr(galpha), _ptr(ws),
                                     _backward_flags(flags), _stream()))
    bflags = _backward_flags(flags)
    fused_fold = bool(info.tc_backward) and not (bflags & (FLAG_FORCE_SIMT | FLAG_DETERMINISTIC))
    # kernels launched: wgrad + finish, dgrad (+ col2im unless the fold is fused into its epilogue), alpha-grad + finish
    # (+ the grad_out scale pre-pass of the v2 dgrad / wgrad)
    _count((2 if need_weight else 0) + ((1 if fused_fold else 2) if need_input else 0) +
           (2 if galpha is not None else 0) + (1 if (flags & FLAG_V2) and (need_weight or need_input) else 0))
    return gxq, gwq, galpha


@_on_tensor_device
def conv_psums(spec: LayerSpec, xcodes, wcodes):
    info = layer_info(spec)
    ps = torch.empty((spec.batch, info.NX, info.NSW, info.NSA, info.L, spec.out_channels), dtype=torch.int32,
                     device=xcodes.device)
    layer = spec.c_layer()
    _check(load().cimq_conv_psums(C.byref(layer), _ptr(xcodes), _ptr(wcodes), _ptr(ps), _stream()))
    return ps


@_on_tensor_device
def conv_psum_abs_sums(spec: LayerSpec, xcodes, wcodes):
    info = layer_info(spec)
    sums = torch.zeros((1, info.NX, info.NSW, info.NSA, 1, spec.out_channels), dtype=torch.int64,
                       device=xcodes.device)
    layer = spec.c_layer()
    _check(load().cimq_conv_psum_abs_sums(C.byref(layer), _ptr(xcodes), _ptr(wcodes), _ptr(sums), _stream()))
    return sums


# ---- batch norm (+ residual) (+ ReLU) ------------------------------------------------------------------------
@_on_tensor_device
def bn_forward(x, residual, weight, bias, running_mean, running_var, training: bool, momentum: float, eps: float,
               relu: bool, next_quant=None):
    """y, save_mean, save_invstd (None, None in inference) [, codes].  ``next_quant = (alpha_act, grad_scale, qp)`` of
    the layer that consumes y: its activation codes are written by the same kernel (cimq_bn_forward_quant)."""
    b, c = x.shape[0], x.shape[1]
    hw = x.numel() // (b * c)
    y = torch.empty_like(x)
    mean = invstd = ws = None
    if training:
        mean = torch.empty(c, dtype=torch.float32, device=x.device)
        invstd = torch.empty(c, dtype=torch.float32, device=x.device)
        ws = torch.empty(load().cimq_bn_workspace_bytes(b, c), dtype=torch.uint8, device=x.device)
    if next_quant is None:
        _check(load().cimq_bn_forward(_ptr(x), _ptr(residual), _ptr(weight), _ptr(bias), _ptr(running_mean),
                                      _ptr(running_var), int(training), float(momentum), float(eps), int(relu), b, c, hw,
                                      _ptr(y), _ptr(mean), _ptr(invstd), _ptr(ws), _stream()))
        _count(2 if training else 1)
        return y, mean, invstd
    alpha_act, gscale, qp = next_quant
    codes = torch.empty(x.shape, dtype=torch.uint8, device=x.device)
    _check(load().cimq_bn_forward_quant(_ptr(x), _ptr(residual), _ptr(weight), _ptr(bias), _ptr(running_mean),
                                        _ptr(running_var), int(training), float(momentum), float(eps), int(relu), b, c,
                                        hw, _ptr(y), _ptr(mean), _ptr(invstd), _ptr(ws), _ptr(alpha_act),
                                        float(gscale), int(qp), _ptr(codes), _stream()))
    _count(2 if training else 1)
    return y, mean, invstd, codes


@_on_tensor_device
def bn_backward(grad_y, x, y, weight, mean, invstd, training: bool, relu: bool, need_residual: bool):
    """grad_x, grad_residual (or None), grad_weight, grad_bias."""
    b, c = x.shape[0], x.shape[1]
    hw = x.numel() // (b * c)
    gx = torch.empty_like(x)
    gres = torch.empty_like(x) if need_residual else None
    gw = torch.empty(c, dtype=torch.float32, device=x.device)
    gb = torch.empty(c, dtype=torch.float32, device=x.device)
    ws = torch.empty(load().cimq_bn_workspace_bytes(b, c), dtype=torch.uint8, device=x.device)
    _check(load().cimq_bn_backward(_ptr(grad_y), _ptr(x), _ptr(y), _ptr(weight), _ptr(mean), _ptr(invstd),
                                   int(training), int(relu), b, c, hw, _ptr(gx), _ptr(gres), _ptr(gw), _ptr(gb),
                                   _ptr(ws), _stream()))
    _count(2)
    return gx, gres, gw, gb
