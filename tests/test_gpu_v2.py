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
]


@pytest.mark.parametrize("case", V2_CASES)
def test_v2_forward_against_oracle(case):
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
