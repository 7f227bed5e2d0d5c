"""Stochastic near-ADC-less read-out (lsq.py:205-220): not bit-comparable with the reference (its
torch RNG), so parity is statistical -- the mean of the CUDA output over many seeds against the exact expectation
of the reference's sampling scheme (oracle.stochastic_code_expectation), a z-test per output element -- plus
reproducibility for a seed and an unchanged backward."""
import numpy as np
import pytest
import torch

from oracle import cim_oracle as O
from tests.test_gpu_parity import _cuda, _lib, _mask, _spec

pytestmark = pytest.mark.gpu


def _setup():
    L = _lib()
    cfg = O.CimConfig(in_channels=16, out_channels=8, kernel=3, stride=1, padding=1, nbits_w=3, nbits_a=3,
                      wbitslice=1, abitslice=1, xbar=64, adcbits=1.5)
    hw, batch = 6, 2
    rng = np.random.default_rng(7)
    xc = rng.integers(0, cfg.qp_a + 1, size=(batch, 16, hw, hw)).astype(np.uint8)
    xc[rng.random(xc.shape) < 0.4] = 0
    wc = rng.integers(cfg.qn_w, cfg.qp_w + 1, size=(8, 16, 3, 3)).astype(np.int8)
    s_a, s_w = np.float32(0.11), np.float32(0.023)   # v = p * 0.0025: partial sums sit well inside the sigmoids
    aq, _ = O.quantize_alpha(cfg, O.init_alpha_cim(cfg, xc, wc, s_w, s_a))
    spec = _spec(cfg, hw, batch)
    s = _cuda(np.array([s_a, s_w], dtype=np.float32))
    xcd, wcd, aqd = _cuda(xc), _cuda(wc).reshape(8, -1), _cuda(aq)
    mask = _mask(cfg)
    table = L.adc_table(spec, s, aqd, mask)
    return L, cfg, spec, (xc, wc, s_w, s_a, aq), (xcd, wcd, aqd, s, mask, table), hw, batch


def test_stochastic_readout_statistics():
    L, cfg, spec, host, dev, hw, batch = _setup()
    xc, wc, s_w, s_a, aq = host
    xcd, wcd, aqd, s, mask, table = dev
    ps_int = O.integer_psums(cfg, xc, wc)
    mean, var = O.stochastic_code_expectation(cfg, ps_int, s_w, s_a, aq)
    amp = (np.asarray(aq, dtype=np.float64).reshape(1, cfg.num_xbars, 3, 3, 1, 8) *
           cfg.binary_mask().astype(np.float64).reshape(1, 1, 3, 3, 1, 1))
    exp_out = (mean * amp).sum(axis=(1, 2, 3))            # [B, L, Cout]
    var_out = (var * amp ** 2).sum(axis=(1, 2, 3))
    assert float((var > 1e-3).mean()) > 0.05, "the test layer must have partial sums inside the sigmoid transition"
    runs = 400
    acc = torch.zeros(batch, 8, hw * hw, device="cuda", dtype=torch.float64)
    for r in range(runs):
        out, _ = L.conv_forward_stochastic(spec, xcd, wcd, table, s, aqd, seed=1000 + r, save_state=False)
        acc += out.double()
    got = (acc / runs).cpu().numpy().transpose(0, 2, 1)
    z = (got - exp_out) / np.sqrt(var_out / runs + 1e-12)
    # 576 output elements: a correct sampler keeps |z| below ~4.5 and the z-scores standard normal on the whole
    assert np.abs(z).max() < 5.0, float(np.abs(z).max())
    assert abs(float(z.mean())) < 0.25 and 0.8 < float(z.std()) < 1.2, (float(z.mean()), float(z.std()))
    # same seed -> same output; different seed -> different output
    a, _ = L.conv_forward_stochastic(spec, xcd, wcd, table, s, aqd, seed=5, save_state=False)
    b, _ = L.conv_forward_stochastic(spec, xcd, wcd, table, s, aqd, seed=5, save_state=False)
    c, _ = L.conv_forward_stochastic(spec, xcd, wcd, table, s, aqd, seed=6, save_state=False)
    assert torch.equal(a, b) and not torch.equal(a, c)


def test_stochastic_state_is_the_deterministic_one():
    """The backward of the reference does not see the sampled code (lsq.py:310-332): same ADC state as the
    deterministic forward."""
    L, cfg, spec, host, dev, hw, batch = _setup()
    xcd, wcd, aqd, s, mask, table = dev
    _, st_s = L.conv_forward_stochastic(spec, xcd, wcd, table, s, aqd, seed=3, save_state=True)
    _, st_d = L.conv_forward(spec, xcd, wcd, None, table, s, mask, save_state=True, flags=L.FLAG_FORCE_SIMT)
    assert torch.equal(st_s, st_d)


def test_module_stochastic_quant_runs_and_trains():
    import cim_quantization_b200 as cq
    torch.manual_seed(0)
    m = cq.Conv2dLSQCiM(16, 16, 3, 1, 1, 1, 1, False, nbits_w=3, nbits_a=3, nbits_alpha=8, wbitslice=1, abitslice=1,
                        xbar=64, adcbits=1.5, stochastic_quant=True).cuda().train()
    x = torch.relu(torch.randn(2, 16, 8, 8, device="cuda")).requires_grad_(True)
    y1 = m(x)
    y1.sum().backward()
    assert x.grad is not None and m.alpha_cim.grad is not None and torch.isfinite(m.weight.grad).all()
    y2 = m(x.detach())
    assert y1.shape == y2.shape and not torch.equal(y1, y2)  # fresh draws every call
