"""BASELINE.json config 2 as a parity matrix: the microbench layer's channel counts (64 -> 64, 3x3) at a mid size
(16x16 images, batch 4) for every crossbar depth {64, 128, 256} x ADC resolution {1, 1.5, 2, 3, 4}, each point checked
against the oracle -- ADC codes / STE clip mask bit-exact, output and all gradients to 1e-5 (max-normalised AND
per-element with the RMS floor) -- on every kernel generation that covers it (CUDA-core, tcgen05 v1, v2)."""
import numpy as np
import pytest
import torch

from oracle import cim_oracle as O
from tests._util import rel_err, rel_err_elem
from tests.test_gpu_parity import TOL, _cuda, _lib, _mask, _spec, oracle_clip, unpack_state
from tests.test_gpu_v2 import split_state_v2, v2_planes_reference

pytestmark = pytest.mark.gpu

POINTS = [(xbar, adc) for xbar in (64, 128, 256) for adc in (1, 1.5, 2, 3, 4)]


@pytest.mark.parametrize("xbar,adc", POINTS)
def test_config2_point_against_oracle(xbar, adc):
    L = _lib()
    cin = cout = 64
    hw, batch, nbits = 16, 4, 3
    cfg = O.CimConfig(in_channels=cin, out_channels=cout, kernel=3, stride=1, padding=1, nbits_w=nbits, nbits_a=nbits,
                      wbitslice=1, abitslice=1, xbar=xbar, adcbits=adc)
    rng = np.random.default_rng(1000 * xbar + int(10 * adc))
    spec = _spec(cfg, hw, batch)
    info = L.layer_info(spec)
    xc = rng.integers(0, cfg.qp_a + 1, size=(batch, cin, hw, hw)).astype(np.uint8)
    xc[rng.random(xc.shape) < 0.4] = 0
    wc = rng.integers(cfg.qn_w, cfg.qp_w + 1, size=(cout, cin, 3, 3)).astype(np.int8)
    s_a, s_w = np.float32(0.173), np.float32(0.0421)
    ps_int = O.integer_psums(cfg, xc, wc)
    aq = scale = None
    if cfg.has_alpha_cim:
        a0 = O.init_alpha_cim(cfg, xc, wc, s_w, s_a)
        a0 = a0 * rng.uniform(0.6, 1.4, size=a0.shape).astype(np.float32)
        aq, aux = O.quantize_alpha(cfg, a0)
        scale = np.float32(aux["scale"])
    ref_out = O.cim_forward(cfg, xc, wc, s_w, s_a, aq)
    go = rng.standard_normal((batch, hw * hw, cout)).astype(np.float32)
    ref_gx, ref_gw, ref_ga = O.cim_backward(cfg, go, xc, wc, s_w, s_a, aq, hw)
    ref_clip = oracle_clip(cfg, ps_int, s_w, s_a, aq)
    ref_codes = O.adc_codes(cfg, ps_int, s_w, s_a, aq) if cfg.has_alpha_cim else None

    s = _cuda(np.array([s_a, s_w], dtype=np.float32))
    xcd, wcd = _cuda(xc), _cuda(wc).reshape(cout, -1)
    aqd = _cuda(aq) if aq is not None else None
    sc = _cuda(np.array([scale], dtype=np.float32)) if scale is not None else None
    mask = _mask(cfg)
    status = torch.zeros(1, dtype=torch.int32, device="cuda")
    table = L.adc_table(spec, s, aqd, mask, status, alpha_scale=sc)
    wdigits, wtiles = L.weight_prepare(spec, wcd)
    god = _cuda(np.ascontiguousarray(go.transpose(0, 2, 1)))
    gens = [("simt", L.FLAG_FORCE_SIMT)]
    if info.tc_forward:
        gens.append(("tc_v1", 0))
    if info.tc_v2:
        gens.append(("tc_v2", L.FLAG_V2))
    assert info.tc_forward, "every point of the matrix must be covered by a tensor-core forward"
    for name, flags in gens:
        out, state = L.conv_forward(spec, xcd, wcd, wtiles, table, s, mask, save_state=True, flags=flags)
        torch.cuda.synchronize()
        assert status.item() == 0, name
        if flags & L.FLAG_V2:
            d, w, c = split_state_v2(state, info, cfg)
            rd, rw, rc = v2_planes_reference(cfg, ref_codes, ref_clip)
            np.testing.assert_array_equal(d, rd, err_msg=name)
            np.testing.assert_array_equal(w, rw, err_msg=name)
            if cfg.has_alpha_cim:
                np.testing.assert_array_equal(c, rc, err_msg=name)
        else:
            codes, clip = unpack_state(state, cfg, info, batch)
            np.testing.assert_array_equal(clip, ref_clip, err_msg=name)
            if cfg.has_alpha_cim:
                np.testing.assert_array_equal(codes, ref_codes, err_msg=name)
        o = out.cpu().numpy().transpose(0, 2, 1)
        assert rel_err(o, ref_out) < TOL and rel_err_elem(o, ref_out) < TOL, name
        gxq, gwq, galpha = L.conv_backward(spec, god, xcd, wdigits, wtiles, state, s, mask,
                                           need_alpha=cfg.has_alpha_cim, flags=flags & ~L.FLAG_V2)
        gx, gw = gxq.cpu().numpy(), gwq.cpu().numpy().reshape(ref_gw.shape)
        assert rel_err(gx, ref_gx) < TOL and rel_err_elem(gx, ref_gx) < 2 * TOL, name
        assert rel_err(gw, ref_gw) < TOL and rel_err_elem(gw, ref_gw) < 2 * TOL, name
        if cfg.has_alpha_cim:
            ga = galpha.cpu().numpy()
            assert rel_err(ga, ref_ga) < TOL and rel_err_elem(ga, ref_ga) < 2 * TOL, name


def test_multibit_clip_at_the_bound_itself():
    """lsq.py:310-311 compares in fp32 against Qp + 1e-5 / Qn - 1e-5; for Qn = -256 (adcbits 9) the 1e-5 is below half
    an ulp, so a partial sum EQUAL to Qn is already clipped (STE mask off).  All-ones activations against all-minus-one
    weights on a 256-row crossbar produce exactly -256 at the interior pixels."""
    L = _lib()
    cin, cout, hw, batch = 32, 16, 6, 1
    cfg = O.CimConfig(in_channels=cin, out_channels=cout, kernel=3, stride=1, padding=1, nbits_w=3, nbits_a=3,
                      wbitslice=1, abitslice=1, xbar=256, adcbits=9)
    xc = np.ones((batch, cin, hw, hw), dtype=np.uint8)
    wc = -np.ones((cout, cin, 3, 3), dtype=np.int8)
    s_a, s_w = np.float32(0.173), np.float32(0.0421)
    ps_int = O.integer_psums(cfg, xc, wc)
    assert ps_int.min() == -256
    ref_clip = oracle_clip(cfg, ps_int, s_w, s_a, None)
    assert ref_clip[ps_int == -256].all() and not ref_clip[ps_int == -255].any()
    rng = np.random.default_rng(3)
    go = rng.standard_normal((batch, hw * hw, cout)).astype(np.float32)
    ref_out = O.cim_forward(cfg, xc, wc, s_w, s_a, None)
    ref_gx, ref_gw, _ = O.cim_backward(cfg, go, xc, wc, s_w, s_a, None, hw)
    spec = _spec(cfg, hw, batch)
    info = L.layer_info(spec)
    s = _cuda(np.array([s_a, s_w], dtype=np.float32))
    xcd, wcd, mask = _cuda(xc), _cuda(wc).reshape(cout, -1), _mask(cfg)
    table = L.adc_table(spec, s, None, mask)
    wdigits, wtiles = L.weight_prepare(spec, wcd)
    god = _cuda(np.ascontiguousarray(go.transpose(0, 2, 1)))
    for flags in ([L.FLAG_FORCE_SIMT, 0] if info.tc_forward else [L.FLAG_FORCE_SIMT]):
        out, state = L.conv_forward(spec, xcd, wcd, wtiles, table, s, mask, save_state=True, flags=flags)
        _, clip = unpack_state(state, cfg, info, batch)
        np.testing.assert_array_equal(clip, ref_clip)
        assert rel_err(out.cpu().numpy().transpose(0, 2, 1), ref_out) < TOL
        gxq, gwq, _ = L.conv_backward(spec, god, xcd, wdigits, wtiles, state, s, mask, need_alpha=False, flags=flags)
        assert rel_err(gxq.cpu().numpy(), ref_gx) < TOL
        assert rel_err(gwq.cpu().numpy().reshape(ref_gw.shape), ref_gw) < TOL
