"""GPU parity tests: the CUDA path (through the C ABI / the module surface) against the oracle and the
golden vectors produced by the reference.  Integer quantities bit-exact, fp32 within 1e-5 relative
(north_star).  Run on the B200 box with ``pytest -m gpu``."""
import math

import numpy as np
import pytest
import torch

from oracle import cim_oracle as O
from tests._util import golden_names, load_golden, rel_err

pytestmark = pytest.mark.gpu

TOL = 1e-5


def _lib():
    from cim_quantization_b200 import _lib
    return _lib


def _spec(cfg, hw, batch):
    return _lib().LayerSpec(batch=batch, in_channels=cfg.in_channels, in_hw=hw, out_channels=cfg.out_channels,
                            kernel=cfg.kernel, stride=cfg.stride, padding=cfg.padding, nbits_a=cfg.nbits_a,
                            abitslice=cfg.abitslice, nbits_w=cfg.nbits_w, wbitslice=cfg.wbitslice, xbar=cfg.xbar,
                            adcbits=cfg.adcbits)


def _cuda(a, dtype=None):
    t = torch.from_numpy(np.ascontiguousarray(a)).cuda()
    return t.to(dtype) if dtype is not None else t


def _mask(cfg):
    return _cuda(cfg.binary_mask().astype(np.int8))


def unpack_state(state, cfg, info, batch):
    """ADC state words -> (codes, clip) arrays shaped like ps_int ``[B,NX,NSW,NSA,L,Cout]``."""
    nx, cout, sw, m = info.NX, cfg.out_channels, info.state_words, info.M
    st = state.cpu().numpy().view(np.uint32).reshape(nx, cout, sw, m)
    pairs = info.pairs
    multibit = cfg.adcbits not in (1, 1.5)

    def bit(pos):
        return ((st[:, :, pos >> 5, :] >> np.uint32(pos & 31)) & 1).astype(np.int32)  # [NX,Cout,M]

    codes = np.zeros((batch, nx, info.NSW, info.NSA, info.L, cout), dtype=np.int32)
    clip = np.zeros_like(codes)
    for k in range(info.NSW):
        for j in range(info.NSA):
            q = j * info.NSW + k  # state pair index (activation-slice major), see include/cimq.h
            if multibit:
                cl = bit(q)
                cd = np.zeros_like(cl)
            else:
                cd = bit(q) - bit(pairs + q)
                cl = bit(2 * pairs + q)
            # [NX,Cout,M] -> [B,NX,L,Cout]
            codes[:, :, k, j] = cd.reshape(nx, cout, batch, info.L).transpose(2, 0, 3, 1)
            clip[:, :, k, j] = cl.reshape(nx, cout, batch, info.L).transpose(2, 0, 3, 1)
    return codes, clip


def oracle_clip(cfg, ps_int, s_w, s_a, alpha_q):
    qn, qp = cfg.adc_range
    if cfg.adcbits in (1, 1.5):
        ps = O._scaled_psums(ps_int, s_w, s_a) / np.asarray(alpha_q, dtype=np.float32)
    else:
        ps = ps_int.astype(np.float16).astype(np.float32)
    return ((ps >= np.float32(qp + 1e-5)) | (ps <= np.float32(qn - 1e-5))).astype(np.int32)


# ----------------------------------------------------------------------------------------------
# LSQ quantiser
# ----------------------------------------------------------------------------------------------
@pytest.mark.parametrize("n", [1, 15, 16, 1000, 4099, 1 << 20])
@pytest.mark.parametrize("signed", [False, True])
def test_lsq_quantize_bit_exact(n, signed):
    L = _lib()
    rng = np.random.default_rng(n + signed)
    x = (rng.standard_normal(n) * 1.5).astype(np.float32)
    x[:: 7] = 0.0
    s = np.float32(0.1234567)
    # exact ties and clamp edges
    if n >= 16:
        x[1] = np.float32(2.5) * s
        x[2] = np.float32(3.5) * s
        x[3] = np.float32(-0.5) * s
        x[4] = np.float32(1e9)
        x[5] = np.float32(-1e9)
    qn, qp = (-4, 3) if signed else (0, 7)
    codes = L.lsq_quantize(_cuda(x), _cuda(np.array([s], dtype=np.float32)), qn, qp)
    ref = O.lsq_codes(x, s, qn, qp)
    np.testing.assert_array_equal(codes.cpu().numpy().astype(np.int32), ref)


@pytest.mark.parametrize("s", [0.1234567, 0.3333333, 1.0, 7.1e-4, 3.0])
def test_lsq_quantize_adversarial_ties(s):
    """Inputs within a few ulps of every rounding tie (k+0.5)*s and of the clamp bounds: the hoisted-reciprocal
    fast path must fall back to the exact division there and still match the IEEE chain bit-for-bit."""
    L = _lib()
    s = np.float32(s)
    ks = np.arange(-6, 10, dtype=np.float32)
    base = np.concatenate([(ks + np.float32(0.5)) * s, ks * s])
    xs = [base]
    for _ in range(6):  # walk ulps up and down around every critical point
        xs.append(np.nextafter(xs[-1], np.float32(np.inf), dtype=np.float32))
    lo = [base]
    for _ in range(6):
        lo.append(np.nextafter(lo[-1], np.float32(-np.inf), dtype=np.float32))
    x = np.concatenate(xs + lo + [np.array([np.inf, -np.inf, 3.4e38, -3.4e38, 1e-45, -1e-45, 0.0, -0.0], np.float32)])
    x = np.tile(x, 5).astype(np.float32)
    sd = _cuda(np.array([s], dtype=np.float32))
    for qn, qp in ((0, 7), (-4, 3), (0, 255), (-128, 127)):
        codes = L.lsq_quantize(_cuda(x), sd, qn, qp)
        np.testing.assert_array_equal(codes.cpu().numpy().astype(np.int32), O.lsq_codes(x, s, qn, qp))
        fin = np.isfinite(x)
        g = np.ones_like(x)
        gx, _ = L.lsq_backward(_cuda(g), _cuda(np.where(fin, x, 0).astype(np.float32)), sd, qn, qp, 1.0)
        u = np.where(fin, x, 0).astype(np.float32) / s
        np.testing.assert_array_equal(gx.cpu().numpy(), ((u >= qn) & (u <= qp)).astype(np.float32))


@pytest.mark.parametrize("n", [37, 1 << 16, (1 << 18) + 5])
def test_lsq_backward(n):
    L = _lib()
    rng = np.random.default_rng(n)
    x = (rng.standard_normal(n) * 1.5).astype(np.float32)
    gq = rng.standard_normal(n).astype(np.float32)
    s = np.float32(0.31)
    g = 1.0 / math.sqrt(n * 7)
    gx, ga = L.lsq_backward(_cuda(gq), _cuda(x), _cuda(np.array([s], dtype=np.float32)), 0, 7, g)
    gx_ref, ga_ref = O.lsq_backward(x, s, 0, 7, g, gq)
    assert rel_err(gx.cpu().numpy(), gx_ref) < TOL
    u = x / s
    terms = np.abs(gq * (np.rint(np.clip(u, 0, 7)) - np.where((u >= 0) & (u <= 7), u, 0)))
    assert abs(float(ga.item()) - ga_ref) <= TOL * g * float(terms.sum())


def test_alpha_quantizer_matches_oracle():
    """Fused alpha_cim range quantiser (lsq.py:566-571) and its autograd-equivalent backward."""
    L = _lib()
    rng = np.random.default_rng(5)
    cfg = O.CimConfig(in_channels=64, out_channels=64, kernel=3, padding=1, xbar=128)
    for shape in ((1, 5, 3, 3, 1, 64), (1, 1, 8, 8, 1, 16), (1, 2, 2, 2, 1, 7)):
        alpha = rng.uniform(0.01, 0.4, size=shape).astype(np.float32)
        alpha.flat[3] = alpha.max()  # a tie at the maximum shares the range gradient
        g = rng.standard_normal(shape).astype(np.float32)
        aq_ref, aux = O.quantize_alpha(cfg, alpha)
        ga_ref = O.quantize_alpha_backward(alpha, g, aux)
        aq, auxd = L.alpha_quantize(_cuda(alpha), 1, 255)
        np.testing.assert_array_equal(aq.cpu().numpy(), aq_ref)
        ga = L.alpha_quantize_backward(_cuda(alpha), _cuda(g), 1, 255, auxd)
        assert rel_err(ga.cpu().numpy(), ga_ref) < TOL


def test_step_sizes_match_grad_scale():
    L = _lib()
    rng = np.random.default_rng(0)
    for _ in range(20):
        a, w = np.float32(rng.uniform(0.01, 2)), np.float32(rng.uniform(0.001, 0.5))
        ga, gw = 1.0 / math.sqrt(rng.integers(100, 10 ** 7) * 7), 1.0 / math.sqrt(rng.integers(100, 10 ** 5) * 3)
        s = L.step_sizes(_cuda(np.array([a])), _cuda(np.array([w])), ga, gw).cpu().numpy()
        assert s[0] == O.grad_scale_value(a, ga) and s[1] == O.grad_scale_value(w, gw)


# ----------------------------------------------------------------------------------------------
# golden cases through the C ABI
# ----------------------------------------------------------------------------------------------
@pytest.mark.parametrize("name", golden_names())
def test_psums_bit_exact(name):
    L = _lib()
    cfg, d, hw, batch = load_golden(name)
    spec = _spec(cfg, hw, batch)
    ps = L.conv_psums(spec, _cuda(d["x_codes"], torch.uint8), _cuda(d["w_codes"], torch.int8))
    np.testing.assert_array_equal(ps.cpu().numpy(), d["ps_int"].astype(np.int32))


@pytest.mark.parametrize("force_simt", [True, False])
@pytest.mark.parametrize("name", golden_names())
def test_conv_forward_backward_cabi(name, force_simt):
    L = _lib()
    cfg, d, hw, batch = load_golden(name)
    spec = _spec(cfg, hw, batch)
    info = L.layer_info(spec)
    if not force_simt and not info.tc_forward:
        pytest.skip("layer not covered by the tcgen05 kernel")
    flags = L.FLAG_FORCE_SIMT if force_simt else 0
    s = _cuda(np.array([d["s_a"].reshape(()), d["s_w"].reshape(())], dtype=np.float32))
    xc, wc = _cuda(d["x_codes"], torch.uint8), _cuda(d["w_codes"], torch.int8).reshape(cfg.out_channels, -1)
    aq = _cuda(d["alpha_q"]) if cfg.has_alpha_cim else None
    mask = _mask(cfg)
    status = torch.zeros(1, dtype=torch.int32, device="cuda")
    table = L.adc_table(spec, s, aq, mask, status)
    wdigits, wtiles = L.weight_prepare(spec, wc)
    out, state = L.conv_forward(spec, xc, wc, wtiles, table, s, mask, save_state=True, flags=flags)
    torch.cuda.synchronize()
    assert status.item() == 0
    # per-chunk partial-sum ADC codes and STE clip mask: bit-exact
    codes, clip = unpack_state(state, cfg, info, batch)
    ps_int = d["ps_int"].astype(np.int32)
    ref_clip = oracle_clip(cfg, ps_int, d["s_w"].reshape(()), d["s_a"].reshape(()), d.get("alpha_q"))
    np.testing.assert_array_equal(clip, ref_clip)
    if cfg.has_alpha_cim:
        ref_codes = O.adc_codes(cfg, ps_int, d["s_w"].reshape(()), d["s_a"].reshape(()), d["alpha_q"])
        np.testing.assert_array_equal(codes, ref_codes)
    # output [B,Cout,L] vs the reference's [B,L,Cout]
    assert rel_err(out.cpu().numpy().transpose(0, 2, 1), d["fn_out"]) < TOL
    # backward
    oh = cfg.out_hw(hw)
    go = _cuda(d["grad_y"].reshape(batch, cfg.out_channels, oh * oh))
    gxq, gwq, galpha = L.conv_backward(spec, go, xc, wdigits, wtiles, state, s, mask, need_alpha=cfg.has_alpha_cim,
                                       flags=flags)
    assert rel_err(gxq.cpu().numpy(), d["fn_grad_xq"]) < TOL
    assert rel_err(gwq.cpu().numpy().reshape(d["fn_grad_wq"].shape), d["fn_grad_wq"]) < TOL
    if cfg.has_alpha_cim:
        assert rel_err(galpha.cpu().numpy(), d["fn_grad_alpha_q"]) < TOL


@pytest.mark.parametrize("name", [n for n in golden_names() if load_golden(n)[0].has_alpha_cim])
def test_alpha_cim_init(name):
    from cim_quantization_b200 import functional as CF
    cfg, d, hw, batch = load_golden(name)
    spec = _spec(cfg, hw, batch)
    s = _cuda(np.array([d["s_a"].reshape(()), d["s_w"].reshape(())], dtype=np.float32))
    a0 = CF.alpha_cim_initial_value(spec, _cuda(d["x_codes"], torch.uint8),
                                    _cuda(d["w_codes"], torch.int8).reshape(cfg.out_channels, -1), s)
    assert rel_err(a0.cpu().numpy(), d["alpha_cim_init"]) < TOL


# ----------------------------------------------------------------------------------------------
# module / Function surface
# ----------------------------------------------------------------------------------------------
def _build_module(cfg, d, force_simt):
    import cim_quantization_b200 as cq
    L = _lib()
    m = cq.Conv2dLSQCiM(cfg.in_channels, cfg.out_channels, (cfg.kernel, cfg.kernel), (cfg.stride, cfg.stride),
                        (cfg.padding, cfg.padding), (1, 1), 1, False, nbits_w=cfg.nbits_w, nbits_a=cfg.nbits_a,
                        nbits_alpha=8, wbitslice=cfg.wbitslice, abitslice=cfg.abitslice, xbar=cfg.xbar,
                        adcbits=cfg.adcbits, signed_xbar=False, stochastic_quant=False)
    sd = {"weight": torch.from_numpy(d["weight"]), "alpha_act": torch.from_numpy(d["alpha_act"]),
          "alpha_weight": torch.from_numpy(d["alpha_weight"]), "init_state": torch.ones(1),
          "signed_act": torch.from_numpy(d["signed_act"]), "init_state_cim": torch.ones(1)}
    if cfg.has_alpha_cim:
        sd["alpha_cim"] = torch.from_numpy(d["alpha_cim"])
    m.load_state_dict(sd)
    m.kernel_flags = L.FLAG_FORCE_SIMT if force_simt else 0
    return m.cuda().train()


@pytest.mark.parametrize("force_simt", [True, False])
@pytest.mark.parametrize("name", golden_names())
def test_module_matches_reference(name, force_simt):
    cfg, d, hw, batch = load_golden(name)
    m = _build_module(cfg, d, force_simt)
    x = _cuda(d["x"]).requires_grad_(True)
    y = m(x)
    y.backward(_cuda(d["grad_y"]))
    assert rel_err(y.detach().cpu().numpy(), d["y"]) < TOL
    assert rel_err(x.grad.cpu().numpy(), d["grad_x"]) < TOL
    assert rel_err(m.weight.grad.cpu().numpy(), d["grad_weight"]) < TOL
    if cfg.has_alpha_cim:
        assert rel_err(m.alpha_cim.grad.cpu().numpy(), d["grad_alpha_cim"]) < TOL
    # step-size gradients: sums of cancelling terms -> tolerance relative to sum |terms|
    r = O.module_forward_backward(cfg, d["x"], d["weight"], d["alpha_act"], d["alpha_weight"], d.get("alpha_cim"),
                                  d["grad_y"])
    for key, p, xs, s, qn, qp, gq in (
            ("grad_alpha_act", m.alpha_act, d["x"], r["s_a"], 0, cfg.qp_a, r["grad_xq"]),
            ("grad_alpha_weight", m.alpha_weight, d["weight"], r["s_w"], cfg.qn_w, cfg.qp_w, r["grad_wq"])):
        u = xs / s
        terms = np.abs(gq * (np.rint(np.clip(u, qn, qp)) - np.where((u >= qn) & (u <= qp), u, 0)))
        g = 1.0 / math.sqrt(xs.size * qp)
        assert abs(float(p.grad.item()) - float(d[key].reshape(()))) <= TOL * g * float(terms.sum()) + 1e-12


@pytest.mark.parametrize("name", ["tern_c16o16_x128_s2", "bin_c16o8_x128", "adc3_w4a4_c8o16_x32"])
def test_function_17_arg_surface(name):
    """``get_cim_output_signed.apply`` keeps the reference's 17 arguments and 17 gradients (lsq.py:92-93, 386)."""
    import cim_quantization_b200 as cq
    cfg, d, hw, batch = load_golden(name)
    s_a, s_w = _cuda(d["s_a"]), _cuda(d["s_w"])
    x_q = (_cuda(d["x_codes"], torch.float32) * s_a).requires_grad_(True)
    w_q = (_cuda(d["w_codes"], torch.float32) * s_w).requires_grad_(True)
    aq = _cuda(d["alpha_q"]).requires_grad_(True) if cfg.has_alpha_cim else None
    mask = _cuda(cfg.binary_mask().astype(np.int8)).view(1, 1, cfg.nsw, cfg.nsa, 1, 1)
    out = cq.get_cim_output_signed.apply(x_q, w_q, (cfg.stride,) * 2, (cfg.padding,) * 2, (1, 1), cfg.nbits_a,
                                         cfg.abitslice, cfg.nbits_w, cfg.wbitslice, cfg.adcbits, cfg.xbar, mask, aq,
                                         s_w, s_a, False, torch.zeros(1))
    assert tuple(out.shape) == d["fn_out"].shape
    assert rel_err(out.detach().cpu().numpy(), d["fn_out"]) < TOL
    oh = cfg.out_hw(hw)
    go = _cuda(d["grad_y"].reshape(batch, cfg.out_channels, oh * oh)).transpose(1, 2)
    out.backward(go)
    assert rel_err(x_q.grad.cpu().numpy(), d["fn_grad_xq"]) < TOL
    assert rel_err(w_q.grad.cpu().numpy(), d["fn_grad_wq"]) < TOL
    if cfg.has_alpha_cim:
        assert rel_err(aq.grad.cpu().numpy(), d["fn_grad_alpha_q"]) < TOL


def test_lazy_init_matches_reference_formulas():
    """First training batch initialises alpha_act / alpha_weight / alpha_cim (lsq.py:532-563)."""
    import cim_quantization_b200 as cq
    cfg, d, hw, batch = load_golden("tern_c16o16_x128_s2")
    m = cq.Conv2dLSQCiM(cfg.in_channels, cfg.out_channels, (3, 3), (2, 2), (1, 1), (1, 1), 1, False, nbits_w=3,
                        nbits_a=3, nbits_alpha=8, wbitslice=1, abitslice=1, xbar=128, adcbits=1.5).cuda().train()
    with torch.no_grad():
        m.weight.copy_(_cuda(d["weight"]))
    x = _cuda(d["x"])
    m(x)
    assert m.init_state.item() == 1 and m.init_state_cim.item() == 1 and m.signed_act.item() == 0
    a_act = O.init_step_size(d["x"], cfg.qp_a)
    a_w = O.init_step_size(d["weight"], cfg.qp_w)
    assert abs(m.alpha_act.item() - a_act) <= 1e-6 * a_act and abs(m.alpha_weight.item() - a_w) <= 1e-6 * a_w
    # alpha_cim against the oracle run with the module's own (device-computed) step sizes
    ga = 1.0 / math.sqrt(d["x"].size * cfg.qp_a)
    gw = 1.0 / math.sqrt(d["weight"].size * cfg.qp_w)
    s_a = O.grad_scale_value(np.float32(m.alpha_act.item()), ga)
    s_w = O.grad_scale_value(np.float32(m.alpha_weight.item()), gw)
    xc = O.lsq_codes(d["x"], s_a, 0, cfg.qp_a)
    wc = O.lsq_codes(d["weight"], s_w, cfg.qn_w, cfg.qp_w)
    assert rel_err(m.alpha_cim.detach().cpu().numpy(), O.init_alpha_cim(cfg, xc, wc, s_w, s_a)) < TOL


# ----------------------------------------------------------------------------------------------
# larger random cases: tcgen05 kernel vs CUDA-core kernel vs oracle
# ----------------------------------------------------------------------------------------------
RANDOM_CASES = [
    # cin, cout, hw, batch, stride, nbits, xbar, adc
    (32, 32, 16, 4, 1, 3, 128, 1.5),
    (64, 64, 8, 3, 1, 3, 64, 1.5),
    (64, 64, 8, 2, 1, 3, 256, 1.5),
    (16, 32, 16, 2, 2, 3, 128, 1),
    (32, 64, 8, 2, 1, 4, 128, 3),
    (16, 16, 12, 5, 1, 2, 128, 1),
    (3, 16, 16, 2, 1, 8, 128, 1.5),
    # Cout = 128: beyond the register-resident operand paths of dgrad (Cout <= 64) and wgrad (Cout <= 72)
    (32, 128, 16, 3, 1, 3, 128, 1.5),
    (64, 128, 8, 2, 1, 3, 128, 1),
    # Cout = 256: the backward runs as two blocks of 128 output channels
    (32, 256, 8, 2, 1, 3, 128, 1.5),
    # several 128-pixel tiles per CTA-less grid, W = 32 rows staged by cp.async, remainder crossbar of 64 rows
    (64, 32, 32, 2, 1, 3, 128, 1.5),
]


@pytest.mark.parametrize("case", RANDOM_CASES)
def test_random_layer_against_oracle(case):
    L = _lib()
    cin, cout, hw, batch, stride, nbits, xbar, adc = case
    cfg = O.CimConfig(in_channels=cin, out_channels=cout, kernel=3, stride=stride, padding=1, nbits_w=nbits,
                      nbits_a=nbits, wbitslice=1, abitslice=1, xbar=xbar, adcbits=adc)
    rng = np.random.default_rng(hash(case) % (2 ** 31))
    spec = _spec(cfg, hw, batch)
    info = L.layer_info(spec)
    xc = rng.integers(0, cfg.qp_a + 1, size=(batch, cin, hw, hw)).astype(np.uint8)
    xc[rng.random(xc.shape) < 0.4] = 0
    wc = rng.integers(cfg.qn_w, cfg.qp_w + 1, size=(cout, cin, 3, 3)).astype(np.int8)
    s_a, s_w = np.float32(0.173), np.float32(0.0421)
    ps_int = O.integer_psums(cfg, xc, wc)
    aq = None
    if cfg.has_alpha_cim:
        a0 = O.init_alpha_cim(cfg, xc, wc, s_w, s_a)
        a0 = a0 * rng.uniform(0.6, 1.4, size=a0.shape).astype(np.float32)
        aq, _ = O.quantize_alpha(cfg, a0)
    ref_out = O.cim_forward(cfg, xc, wc, s_w, s_a, aq)
    oh = cfg.out_hw(hw)
    go = rng.standard_normal((batch, oh * oh, cout)).astype(np.float32)
    ref_gx, ref_gw, ref_ga = O.cim_backward(cfg, go, xc, wc, s_w, s_a, aq, hw)
    ref_clip = oracle_clip(cfg, ps_int, s_w, s_a, aq)

    s = _cuda(np.array([s_a, s_w], dtype=np.float32))
    xcd, wcd = _cuda(xc), _cuda(wc).reshape(cout, -1)
    aqd = _cuda(aq) if aq is not None else None
    mask = _mask(cfg)
    table = L.adc_table(spec, s, aqd, mask)
    wdigits, wtiles = L.weight_prepare(spec, wcd)
    np.testing.assert_array_equal(L.conv_psums(spec, xcd, wcd).cpu().numpy(), ps_int)
    for flags in ([L.FLAG_FORCE_SIMT, 0, L.FLAG_DETERMINISTIC] if info.tc_forward else [L.FLAG_FORCE_SIMT]):
        out, state = L.conv_forward(spec, xcd, wcd, wtiles, table, s, mask, save_state=True, flags=flags)
        codes, clip = unpack_state(state, cfg, info, batch)
        np.testing.assert_array_equal(clip, ref_clip)
        if cfg.has_alpha_cim:
            np.testing.assert_array_equal(codes, O.adc_codes(cfg, ps_int, s_w, s_a, aq))
        assert rel_err(out.cpu().numpy().transpose(0, 2, 1), ref_out) < TOL
        gxq, gwq, galpha = L.conv_backward(spec, _cuda(np.ascontiguousarray(go.transpose(0, 2, 1))), xcd, wdigits,
                                           wtiles, state, s, mask, need_alpha=cfg.has_alpha_cim, flags=flags)
        assert rel_err(gxq.cpu().numpy(), ref_gx) < TOL
        assert rel_err(gwq.cpu().numpy().reshape(ref_gw.shape), ref_gw) < TOL
        if cfg.has_alpha_cim:
            assert rel_err(galpha.cpu().numpy(), ref_ga) < TOL


# ----------------------------------------------------------------------------------------------
# BASELINE.json full-size microbench layer: size-independent properties
# ----------------------------------------------------------------------------------------------
def test_wide_adc_equals_dense_conv_full_size():
    """P1 (reference test/test_cim.py:39-57): with an ADC wider than any partial sum the CiM conv is the
    dense integer convolution, exactly; its gradients are the plain conv gradients.  Full microbench
    shape: 3x3, 64->64, 32x32, batch 256, xbar 128."""
    L = _lib()
    B, C, HW = 256, 64, 32
    cfg = O.CimConfig(in_channels=C, out_channels=C, kernel=3, stride=1, padding=1, nbits_w=3, nbits_a=3,
                      wbitslice=1, abitslice=1, xbar=128, adcbits=12)
    spec = _spec(cfg, HW, B)
    g = torch.Generator(device="cuda").manual_seed(0)
    xc = torch.randint(0, 8, (B, C, HW, HW), device="cuda", generator=g, dtype=torch.uint8)
    wc = torch.randint(-4, 4, (C, C, 3, 3), device="cuda", generator=g, dtype=torch.int8)
    s = torch.ones(2, device="cuda")
    mask = _mask(cfg)
    table = L.adc_table(spec, s, None, mask)
    wdigits, wtiles = L.weight_prepare(spec, wc.reshape(C, -1))
    # exact reference: im2col + fp32 SGEMM on small integers (cuDNN may pick an inexact Winograd kernel)
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    cols = torch.nn.functional.unfold(xc.float(), 3, padding=1)  # [B, F, L]
    dense = torch.matmul(wc.float().view(C, -1), cols).view(B, C, HW, HW)
    del cols
    for flags in (0, L.FLAG_FORCE_SIMT):
        out, state = L.conv_forward(spec, xc, wc.reshape(C, -1), wtiles, table, s, mask, save_state=True,
                                    flags=flags)
        assert int((out.view_as(dense) != dense).sum()) == 0
        assert int(state.count_nonzero()) == 0  # nothing clipped
    go = torch.randn(B, C, HW * HW, device="cuda", generator=g)
    # float64 reference: an fp32 reduction over 262144 pixels carries ~1e-5 of rounding of its own
    xf = xc.double().requires_grad_(True)
    wf = wc.double().requires_grad_(True)
    torch.nn.functional.conv2d(xf, wf, padding=1).backward(go.double().view(B, C, HW, HW))
    for flags in (0, L.FLAG_FORCE_SIMT):  # tcgen05 dgrad/wgrad, then the CUDA-core kernels
        gxq, gwq, _ = L.conv_backward(spec, go, xc, wdigits, wtiles, state, s, mask, need_alpha=False, flags=flags)
        assert rel_err(gxq.cpu().numpy(), xf.grad.cpu().numpy()) < TOL
        assert rel_err(gwq.view_as(wf).cpu().numpy(), wf.grad.cpu().numpy()) < TOL


@pytest.mark.parametrize("xbar,adc", [(64, 1.5), (128, 1.5), (128, 1), (128, 3)])
def test_full_size_tc_equals_simt(xbar, adc):
    """Full microbench shape: the tcgen05 kernel and the CUDA-core kernel agree bit-for-bit on every ADC
    code / clip bit (state words) and to fp32 rounding on the output."""
    L = _lib()
    B, C, HW = 256, 64, 32
    cfg = O.CimConfig(in_channels=C, out_channels=C, kernel=3, stride=1, padding=1, nbits_w=3, nbits_a=3,
                      wbitslice=1, abitslice=1, xbar=xbar, adcbits=adc)
    spec = _spec(cfg, HW, B)
    info = L.layer_info(spec)
    assert info.tc_forward
    g = torch.Generator(device="cuda").manual_seed(1)
    xc = (torch.randint(0, 8, (B, C, HW, HW), device="cuda", generator=g, dtype=torch.uint8) *
          (torch.rand(B, C, HW, HW, device="cuda", generator=g) < 0.5)).to(torch.uint8)
    wc = torch.randint(-4, 4, (C, C * 9), device="cuda", generator=g, dtype=torch.int8)
    s = torch.tensor([0.21, 0.037], device="cuda")
    mask = _mask(cfg)
    aq = None
    if cfg.has_alpha_cim:
        sums = L.conv_psum_abs_sums(spec, xc, wc).double()
        aq = (2.0 * sums / (B * HW * HW) * float(s[0]) * float(s[1])).float().clamp_min(1e-4)
        aq = aq * (0.6 + 0.8 * torch.rand(aq.shape, device="cuda", generator=g))
    table = L.adc_table(spec, s, aq, mask)
    _, wtiles = L.weight_prepare(spec, wc)
    out_tc, st_tc = L.conv_forward(spec, xc, wc, wtiles, table, s, mask, save_state=True, flags=0)
    out_si, st_si = L.conv_forward(spec, xc, wc, wtiles, table, s, mask, save_state=True,
                                   flags=L.FLAG_FORCE_SIMT)
    assert torch.equal(st_tc, st_si)
    assert rel_err(out_tc.cpu().numpy(), out_si.cpu().numpy()) < TOL


@pytest.mark.parametrize("name", ["tern_c16o16_x128_s2", "tern_c32o32_x64", "tern_c16o64_x128"])
def test_backward_deterministic_fold(name):
    """CIMQ_FLAG_DETERMINISTIC: the separate fixed-order fold gives the golden grad_x and is bit-identical run to
    run; the default (fp32 reductions in the dgrad epilogue) agrees with it to rounding."""
    L = _lib()
    cfg, d, hw, batch = load_golden(name)
    spec = _spec(cfg, hw, batch)
    info = L.layer_info(spec)
    if not info.tc_backward:
        pytest.skip("layer not covered by the tcgen05 backward kernels")
    s = _cuda(np.array([d["s_a"].reshape(()), d["s_w"].reshape(())], dtype=np.float32))
    xc, wc = _cuda(d["x_codes"], torch.uint8), _cuda(d["w_codes"], torch.int8).reshape(cfg.out_channels, -1)
    aq = _cuda(d["alpha_q"]) if cfg.has_alpha_cim else None
    mask = _mask(cfg)
    table = L.adc_table(spec, s, aq, mask)
    wdigits, wtiles = L.weight_prepare(spec, wc)
    _, state = L.conv_forward(spec, xc, wc, wtiles, table, s, mask, save_state=True)
    oh = cfg.out_hw(hw)
    go = _cuda(d["grad_y"].reshape(batch, cfg.out_channels, oh * oh))
    runs = [L.conv_backward(spec, go, xc, wdigits, wtiles, state, s, mask, need_alpha=False,
                            flags=L.FLAG_DETERMINISTIC)[0].cpu().numpy() for _ in range(2)]
    np.testing.assert_array_equal(runs[0], runs[1])
    assert rel_err(runs[0], d["fn_grad_xq"]) < TOL
    fused = L.conv_backward(spec, go, xc, wdigits, wtiles, state, s, mask, need_alpha=False)[0].cpu().numpy()
    assert rel_err(fused, d["fn_grad_xq"]) < TOL
    assert rel_err(fused, runs[0]) < 1e-6


@pytest.mark.parametrize("xbar", [128, 256])
def test_full_size_backward_tc_equals_simt(xbar):
    """Full microbench shape (3x3 64->64 32x32 B=256 w3a3 adc1.5): the tcgen05 dgrad / wgrad kernels
    (register-resident operands, staged rows, fused fold; xbar 256 as 128-row virtual chunks sharing the ADC state)
    against the CUDA-core backward on the same ADC state."""
    L = _lib()
    B, C, HW = 256, 64, 32
    cfg = O.CimConfig(in_channels=C, out_channels=C, kernel=3, stride=1, padding=1, nbits_w=3, nbits_a=3,
                      wbitslice=1, abitslice=1, xbar=xbar, adcbits=1.5)
    spec = _spec(cfg, HW, B)
    info = L.layer_info(spec)
    assert info.tc_backward
    g = torch.Generator(device="cuda").manual_seed(2)
    xc = (torch.randint(0, 8, (B, C, HW, HW), device="cuda", generator=g, dtype=torch.uint8) *
          (torch.rand(B, C, HW, HW, device="cuda", generator=g) < 0.5)).to(torch.uint8)
    wc = torch.randint(-4, 4, (C, C * 9), device="cuda", generator=g, dtype=torch.int8)
    s = torch.tensor([0.21, 0.037], device="cuda")
    mask = _mask(cfg)
    sums = L.conv_psum_abs_sums(spec, xc, wc).double()
    aq = (2.0 * sums / (B * HW * HW) * float(s[0]) * float(s[1])).float().clamp_min(1e-4)
    table = L.adc_table(spec, s, aq, mask)
    wdigits, wtiles = L.weight_prepare(spec, wc)
    _, state = L.conv_forward(spec, xc, wc, wtiles, table, s, mask, save_state=True)
    go = torch.randn(B, C, HW * HW, device="cuda", generator=g)
    ref = L.conv_backward(spec, go, xc, wdigits, wtiles, state, s, mask, need_alpha=True, flags=L.FLAG_FORCE_SIMT)
    for flags in (0, L.FLAG_DETERMINISTIC):
        got = L.conv_backward(spec, go, xc, wdigits, wtiles, state, s, mask, need_alpha=True, flags=flags)
        for a, b, name in zip(got, ref, ("grad_xq", "grad_wq", "grad_alpha_q")):
            assert rel_err(a.cpu().numpy(), b.cpu().numpy()) < TOL, (name, flags)
