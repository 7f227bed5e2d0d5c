"""GPU parity tests of the second-generation ("v2") kernels: fp16 partial sums in tensor memory, packed-half ADC
epilogue, tensor-core shift-and-add, uint8 ADC-state planes (include/cimq.h, csrc/cim_v2.cuh).

Everything is checked through the C ABI against the numpy oracle (oracle/cim_oracle.py, pinned to the reference by
tests/test_oracle_golden.py): ADC codes and STE clip counts bit-exact, outputs and gradients to 1e-5."""
import numpy as np
import pytest
import torch

from oracle import cim_oracle as O
from tests._util import golden_names, load_golden, rel_err
from tests.test_gpu_parity import TOL, _cuda, _lib, _mask, _spec, oracle_clip

pytestmark = pytest.mark.gpu


def v2_planes_reference(cfg, codes, clip):
    """(codes, clip) shaped ``[B,NX,NSW,NSA,L,Cout]`` -> the three v2 state planes as the kernels lay them out:
    D ``[NX,M,Cout]``, W ``[NX,M,Cout]``, C ``[NX,NSA,M,Cout]`` (M = B*L, pixel-major then channel)."""
    b, nx, nsw, nsa, l, cout = clip.shape
    passed = 1 - clip
    d = sum((4 ** k) * passed[:, :, k].sum(axis=2) for k in range(nsw))       # [B,NX,L,Cout]
    w = sum((4 ** j) * passed[:, :, :, j].sum(axis=2) for j in range(nsa))
    d = d.transpose(1, 0, 2, 3).reshape(nx, b * l, cout).astype(np.uint8)
    w = w.transpose(1, 0, 2, 3).reshape(nx, b * l, cout).astype(np.uint8)
    c = None
    if codes is not None:
        c = sum((4 ** k) * (codes[:, :, k] + 1) for k in range(nsw))          # [B,NX,NSA,L,Cout]
        c = c.transpose(1, 2, 0, 3, 4).reshape(nx, nsa, b * l, cout).astype(np.uint8)
    return d, w, c


def split_state_v2(state, info, cfg):
    st = state.cpu().numpy()
    n = info.NX * info.M * cfg.out_channels
    d = st[:n].reshape(info.NX, info.M, cfg.out_channels)
    w = st[n:2 * n].reshape(info.NX, info.M, cfg.out_channels)
    c = None
    if cfg.has_alpha_cim:
        c = st[2 * n:].reshape(info.NX, info.NSA, info.M, cfg.out_channels)
    return d, w, c


def _random_case(case):
    cin, cout, hw, batch, stride, nbits, xbar, adc, k = case
    cfg = O.CimConfig(in_channels=cin, out_channels=cout, kernel=k, stride=stride, padding=k // 2, nbits_w=nbits,
                      nbits_a=nbits, wbitslice=1, abitslice=1, xbar=xbar, adcbits=adc)
    rng = np.random.default_rng(abs(hash(case)) % (2 ** 31))
    xc = rng.integers(0, cfg.qp_a + 1, size=(batch, cin, hw, hw)).astype(np.uint8)
    xc[rng.random(xc.shape) < 0.4] = 0
    wc = rng.integers(cfg.qn_w, cfg.qp_w + 1, size=(cout, cin, k, k)).astype(np.int8)
    s_a, s_w = np.float32(0.173), np.float32(0.0421)
    aq = scale = None
    if cfg.has_alpha_cim:
        a0 = O.init_alpha_cim(cfg, xc, wc, s_w, s_a)
        a0 = a0 * rng.uniform(0.6, 1.4, size=a0.shape).astype(np.float32)
        aq, aux = O.quantize_alpha(cfg, a0)
        scale = np.float32(aux["scale"])
    return cfg, rng, xc, wc, s_a, s_w, aq, scale


# (cin, cout, hw, batch, stride, nbits, xbar, adcbits, kernel)
V2_CASES = [
    (16, 16, 8, 2, 1, 3, 128, 1.5, 3),    # Cout 16: one epilogue warpgroup, remainder crossbar of 16 rows
    (32, 32, 8, 2, 1, 3, 64, 1.5, 3),     # CT 32, 64-row crossbars
    (64, 64, 8, 2, 1, 3, 128, 1.5, 3),    # the microbench layer at a small image (one tile)
    (64, 64, 32, 2, 1, 3, 128, 1.5, 3),   # several tiles per CTA, rows staged by cp.async
    (16, 32, 16, 2, 2, 3, 128, 1.5, 3),   # stride 2
    (64, 64, 8, 2, 1, 3, 128, 1, 3),      # binary ADC
    (64, 64, 8, 2, 1, 3, 128, 3, 3),      # multi-bit ADC (clamp), no alpha
    (64, 64, 8, 2, 1, 3, 64, 2, 3),
    (32, 64, 8, 2, 1, 2, 64, 1, 3),       # two digit planes
    (64, 128, 8, 3, 1, 3, 128, 1.5, 3),   # two channel tiles, ragged last pixel tile (192 pixels)
    (48, 32, 8, 2, 1, 3, 128, 1.5, 1),    # 1x1 kernel: generic producer
    (16, 16, 4, 8, 1, 3, 128, 1.5, 3),    # 4x4 images (L = 16): generic gather / fold, alpha-grad on the plane-per-thread kernel
]


@pytest.mark.parametrize("case", V2_CASES)
def test_v2_forward_backward_against_oracle(case):
    L = _lib()
    cfg, rng, xc, wc, s_a, s_w, aq, scale = _random_case(case)
    cin, cout, hw, batch = case[0], case[1], case[2], case[3]
    spec = _spec(cfg, hw, batch)
    info = L.layer_info(spec)
    assert info.tc_v2, "case is meant to be covered by the v2 kernels"
    ps_int = O.integer_psums(cfg, xc, wc)
    ref_out = O.cim_forward(cfg, xc, wc, s_w, s_a, aq)
    ref_clip = oracle_clip(cfg, ps_int, s_w, s_a, aq)
    ref_codes = O.adc_codes(cfg, ps_int, s_w, s_a, aq) if cfg.has_alpha_cim else None
    rd, rw, rc = v2_planes_reference(cfg, ref_codes, ref_clip)

    s = _cuda(np.array([s_a, s_w], dtype=np.float32))
    xcd, wcd = _cuda(xc), _cuda(wc).reshape(cout, -1)
    aqd = _cuda(aq) if aq is not None else None
    sc = _cuda(np.array([scale], dtype=np.float32)) if scale is not None else None
    mask = _mask(cfg)
    status = torch.zeros(1, dtype=torch.int32, device="cuda")
    table = L.adc_table(spec, s, aqd, mask, status, alpha_scale=sc)
    _, wtiles = L.weight_prepare(spec, wcd, want_digits=False)
    oh = cfg.out_hw(hw)
    go = rng.standard_normal((batch, oh * oh, cout)).astype(np.float32)
    ref_gx, ref_gw, ref_ga = O.cim_backward(cfg, go, xc, wc, s_w, s_a, aq, hw)
    god = _cuda(np.ascontiguousarray(go.transpose(0, 2, 1)))
    for save in (False, True):
        out, state = L.conv_forward(spec, xcd, wcd, wtiles, table, s, mask, save_state=save, flags=L.FLAG_V2)
        torch.cuda.synchronize()
        assert status.item() == 0
        assert rel_err(out.cpu().numpy().transpose(0, 2, 1), ref_out) < TOL
        if save:
            assert state.dtype == torch.uint8
            d, w, c = split_state_v2(state, info, cfg)
            np.testing.assert_array_equal(d, rd)
            np.testing.assert_array_equal(w, rw)
            if cfg.has_alpha_cim:
                np.testing.assert_array_equal(c, rc)
            # backward on the v2 planes (the uint8 state selects CIMQ_FLAG_V2): fused fold and the deterministic path
            for flags in (0, L.FLAG_DETERMINISTIC):
                gxq, gwq, galpha = L.conv_backward(spec, god, xcd, None, wtiles, state, s, mask,
                                                   need_alpha=cfg.has_alpha_cim, flags=flags)
                assert rel_err(gxq.cpu().numpy(), ref_gx) < TOL
                assert rel_err(gwq.cpu().numpy().reshape(ref_gw.shape), ref_gw) < TOL
                if cfg.has_alpha_cim:
                    assert rel_err(galpha.cpu().numpy(), ref_ga) < TOL


def test_full_size_v2_backward_equals_v1():
    """The microbench layer at FULL size (B=256: 2048 pixel tiles, the non-split grad_out scale pass, many tiles per
    CTA in every kernel): the oracle is too slow there, so the v2 backward (byte planes, two fp16 pieces of grad_out)
    is compared with the v1 backward (uint32 state words, three bf16 terms) on the same inputs -- two independent
    kernel generations that both pass the oracle at small sizes."""
    L = _lib()
    B, C, HW = 256, 64, 32
    spec = L.LayerSpec(B, C, HW, C, 3, 1, 1, 3, 1, 3, 1, 128, 1.5)
    g = torch.Generator(device="cuda").manual_seed(5)
    x = torch.relu(torch.randn(B, C, HW, HW, device="cuda", generator=g))
    w = torch.randn(C, C * 9, device="cuda", generator=g) * (2.0 / (C * 9)) ** 0.5
    s = torch.stack([2 * x.abs().mean() / 7 ** 0.5, 2 * w.abs().mean() / 3 ** 0.5]).float()
    xc = L.lsq_quantize(x, s[0:1], 0, 7)
    wc = L.lsq_quantize(w, s[1:2], -4, 3)
    mask = torch.tensor([[1, 2, 4], [2, 4, 8], [4, 8, 16]], dtype=torch.int8, device="cuda")
    sums = L.conv_psum_abs_sums(spec, xc, wc).double()
    a0 = (2.0 * sums / (B * HW * HW) * float(s[0]) * float(s[1])).float().clamp_min(1e-6).contiguous()
    aq, aux = L.alpha_quantize(a0, 1, 255)
    table = L.adc_table(spec, s, aq, mask, alpha_scale=aux[0:1].clone())
    wdig, wtiles = L.weight_prepare(spec, wc)
    go = torch.randn(B, C, HW * HW, device="cuda", generator=g)
    res = {}
    for gen, flags in (("v1", 0), ("v2", L.FLAG_V2)):
        out, state = L.conv_forward(spec, xc, wc, wtiles, table, s, mask, save_state=True, flags=flags)
        res[gen] = (out,) + tuple(L.conv_backward(spec, go, xc, wdig, wtiles, state, s, mask, need_alpha=True))
        del state
    torch.cuda.synchronize()
    for name, a, b in zip(("out", "grad_x", "grad_w", "grad_alpha"), res["v1"], res["v2"]):
        err = ((a.double() - b.double()).abs().max() / a.double().abs().max()).item()
        assert err < 2e-5, (name, err)


@pytest.mark.parametrize("name", golden_names())
def test_v2_module_matches_reference_goldens(name):
    """The module surface picks the v2 kernels by itself where they cover the layer: replay the reference's golden
    vectors (tests/golden/make_golden.py) through Conv2dLSQCiM and compare outputs and all gradients."""
    from tests.test_gpu_parity import _build_module
    L = _lib()
    cfg, d, hw, batch = load_golden(name)
    if not L.layer_info(_spec(cfg, hw, batch)).tc_v2:
        pytest.skip("layer not covered by the v2 kernels")
    m = _build_module(cfg, d, force_simt=False)
    x = _cuda(d["x"]).requires_grad_(True)
    y = m(x)
    y.backward(_cuda(d["grad_y"]))
    assert rel_err(y.detach().cpu().numpy(), d["y"]) < TOL
    assert rel_err(x.grad.cpu().numpy(), d["grad_x"]) < TOL
    assert rel_err(m.weight.grad.cpu().numpy(), d["grad_weight"]) < TOL


def test_init_state_reset_reinitialises_alpha_cim():
    """The reference reads ``init_state_cim`` on every forward (lsq.py:557): zeroing the buffer in place must make the
    next training forward recompute alpha_cim (the host mirror of the flags follows the buffers' version counters)."""
    import cim_quantization_b200 as cq
    torch.manual_seed(0)
    m = cq.Conv2dLSQCiM(16, 16, (3, 3), (1, 1), (1, 1), (1, 1), 1, False, nbits_w=3, nbits_a=3, nbits_alpha=8,
                        wbitslice=1, abitslice=1, xbar=128, adcbits=1.5).cuda().train()
    x = torch.relu(torch.randn(2, 16, 8, 8, device="cuda"))
    m(x)
    a0 = m.alpha_cim.detach().clone()
    assert m.init_state_cim.item() == 1
    with torch.no_grad():
        m.alpha_cim.fill_(123.0)
    m.init_state_cim.fill_(0)
    m(x)
    assert m.init_state_cim.item() == 1
    torch.testing.assert_close(m.alpha_cim.detach(), a0)


def test_tensors_on_another_device_than_current():
    """Kernels run in the device context of their tensors, not the caller's current device."""
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    L = _lib()
    x = torch.randn(1024, device="cuda:1")
    s = torch.tensor([0.1], device="cuda:1")
    with torch.cuda.device(0):
        codes = L.lsq_quantize(x, s, -4, 3)
    assert codes.device == x.device
    torch.testing.assert_close(codes.float(), torch.clamp(torch.round(x / s), -4, 3))


def _module_on_golden(name):
    from tests.test_gpu_parity import _build_module
    cfg, d, hw, batch = load_golden(name)
    m = _build_module(cfg, d, force_simt=False)
    x = _cuda(d["x"]).requires_grad_(True)
    y = m(x)
    y.backward(_cuda(d["grad_y"]))
    clean = O.module_forward_backward(cfg, d["x"], d["weight"], d["alpha_act"], d["alpha_weight"], d.get("alpha_cim"),
                                      d["grad_y"])
    return cfg, d, m, x, y, clean


def test_h6_int8_saved_codes_deviation_is_only_in_grad_weight():
    """Known deviation: the reference's backward re-reads 8-bit activation codes through an int8 save
    (lsq.py:99), which turns codes >= 128 of a signed_act layer into negative digits (the oracle reproduces
    it, tests/test_oracle_golden.py).  The CUDA path uses the digits the forward used.  On a reference-generated case
    with 13 % of the codes >= 128: output, grad_x and grad_alpha_cim equal the reference's to 1e-5; grad_weight equals
    the un-wrapped oracle gradient to 1e-5 and differs from the reference's wrapped one by the (large) artefact."""
    cfg, d, m, x, y, clean = _module_on_golden("h6_first_w8a8_c3o16_x128")
    assert rel_err(y.detach().cpu().numpy(), d["y"]) < TOL
    assert rel_err(x.grad.cpu().numpy(), d["grad_x"]) < TOL
    assert rel_err(m.alpha_cim.grad.cpu().numpy(), d["grad_alpha_cim"]) < TOL
    gw = m.weight.grad.cpu().numpy()
    assert rel_err(gw, clean["grad_weight"]) < TOL
    dev = rel_err(gw, d["grad_weight"])
    print(f"H6 deviation of grad_weight from the reference (codes >= 128: "
          f"{float((d['x_codes'] >= 128).mean()):.1%}): {dev:.3f}")
    assert dev > 0.05  # the artefact is real; if this ever drops to ~0 the reference behaviour is being reproduced


def test_h1_unsnapped_step_sizes_deviation_is_only_in_grad_x():
    """Known deviation: with step sizes whose ``fl(fl(k*s)/s) != k`` for some weight code the
    reference's backward truncates 0.9999998 digits to 0 (int8 save, lsq.py:160, 249); its grad_x is then off by ~10 %
    from what its own forward implies.  The CUDA path never divides fake-quant floats: output, grad_weight and
    grad_alpha_cim equal the reference's to 1e-5, grad_x equals the artefact-free oracle value."""
    cfg, d, m, x, y, clean = _module_on_golden("h1_tern_c16o16_x128_unsnapped")
    assert rel_err(y.detach().cpu().numpy(), d["y"]) < TOL
    assert rel_err(m.weight.grad.cpu().numpy(), d["grad_weight"]) < TOL
    assert rel_err(m.alpha_cim.grad.cpu().numpy(), d["grad_alpha_cim"]) < TOL
    gx = x.grad.cpu().numpy()
    assert rel_err(gx, clean["grad_x"]) < TOL
    dev = rel_err(gx, d["grad_x"])
    print(f"H1 deviation of grad_x from the reference on un-snapped step sizes: {dev:.3f}")
    assert 0.01 < dev < 0.5


@pytest.mark.parametrize("case", [V2_CASES[0], V2_CASES[2], V2_CASES[4], V2_CASES[6], V2_CASES[8], V2_CASES[9]])
def test_layer_prepare_equals_separate_entry_points(case):
    """cimq_layer_prepare (one launch) writes the same bytes as cimq_step_sizes + cimq_lsq_quantize +
    cimq_alpha_quantize + cimq_adc_table2 + cimq_weight_prepare for everything the v2 kernels read."""
    L = _lib()
    cfg, rng, xc, wc, s_a, s_w, aq, scale = _random_case(case)
    cin, cout, hw, batch, k = case[0], case[1], case[2], case[3], case[8]
    spec = _spec(cfg, hw, batch)
    info = L.layer_info(spec)
    w = (rng.standard_normal((cout, cin * k * k)) * 0.2).astype(np.float32)
    aa, aw = np.float32(0.173), np.float32(0.0891)
    ga, gw = 1.0 / np.sqrt(batch * cin * hw * hw * cfg.qp_a), 1.0 / np.sqrt(w.size * cfg.qp_w)
    alpha = None
    if cfg.has_alpha_cim:
        alpha = (rng.uniform(0.05, 2.0, size=(1, info.NX, info.NSW, info.NSA, 1, cout))).astype(np.float32)
    wd, aad, awd = _cuda(w), _cuda(np.array([aa])), _cuda(np.array([aw]))
    ald = _cuda(alpha) if alpha is not None else None
    mask = _mask(cfg)
    st1 = torch.zeros(1, dtype=torch.int32, device="cuda")
    s1, wc1, aq1, aux1, tab1, wt1 = L.layer_prepare(spec, wd, aad, awd, ga, gw, ald, 1, 255, mask, st1)
    # the separate path
    s2 = L.step_sizes(aad, awd, ga, gw)
    wc2 = L.lsq_quantize(wd, s2[1:2], cfg.qn_w, cfg.qp_w)
    aq2 = aux2 = None
    if alpha is not None:
        aq2, aux2 = L.alpha_quantize(ald, 1, 255)
    st2 = torch.zeros(1, dtype=torch.int32, device="cuda")
    tab2 = L.adc_table(spec, s2, aq2, mask, st2, alpha_scale=aux2[0:1].clone() if aux2 is not None else None)
    _, wt2 = L.weight_prepare(spec, wc2, want_digits=False)
    torch.cuda.synchronize()
    assert st1.item() == 0 and st2.item() == 0
    np.testing.assert_array_equal(s1.cpu().numpy(), s2.cpu().numpy())
    np.testing.assert_array_equal(wc1.cpu().numpy(), wc2.cpu().numpy())
    if alpha is not None:
        np.testing.assert_array_equal(aq1.cpu().numpy(), aq2.cpu().numpy())
        np.testing.assert_array_equal(aux1.cpu().numpy()[:5], aux2.cpu().numpy()[:5])
    # the sections the v2 kernels (and the CUDA-core forward: the AoS table) read; the v1-only sections are not written
    n_e = info.NX * info.pairs * cout
    t1, t2 = tab1.cpu().numpy(), tab2.cpu().numpy()
    np.testing.assert_array_equal(t1[:16 * n_e], t2[:16 * n_e])
    v2_off = ((16 * n_e + 255) // 256 * 256 + 12 * n_e + 255) // 256 * 256
    np.testing.assert_array_equal(t1[v2_off:v2_off + 8], t2[v2_off:v2_off + 8])      # header {o0, o1}
    np.testing.assert_array_equal(t1[v2_off + 256:], t2[v2_off + 256:])              # constants blocks
    # weight tiles: run the forward + backward on both and compare results bit for bit (covers fwd8, bwd2, lut)
    xcd = _cuda(xc)
    go = _cuda(rng.standard_normal((batch, cout, cfg.out_hw(hw) ** 2)).astype(np.float32))
    res = []
    for (s_, wc_, tab_, wt_) in ((s1, wc1, tab1, wt1), (s2, wc2, tab2, wt2)):
        out, state = L.conv_forward(spec, xcd, wc_, wt_, tab_, s_, mask, save_state=True, flags=L.FLAG_V2)
        gxq, gwq, ga_ = L.conv_backward(spec, go, xcd, None, wt_, state, s_, mask, need_alpha=alpha is not None,
                                        flags=L.FLAG_DETERMINISTIC)
        res.append((out, state, gxq, gwq, ga_))
    for a_, b_ in zip(res[0], res[1]):
        if a_ is not None:
            assert torch.equal(a_, b_)
