import sys, numpy as np, torch
sys.path.insert(0, "/root/repo")
from tests.test_gpu_v2 import _random_case, V2_CASES
from tests.test_gpu_parity import _lib, _spec, _cuda, _mask
from tests._util import rel_err
from oracle import cim_oracle as O
L = _lib()
for case in [V2_CASES[10], (48, 32, 8, 2, 1, 3, 128, 1.5, 3)]:
    cfg, rng, xc, wc, s_a, s_w, aq, scale = _random_case(case)
    cin, cout, hw, batch = case[:4]
    spec = _spec(cfg, hw, batch)
    s = _cuda(np.array([s_a, s_w], dtype=np.float32)); xcd, wcd = _cuda(xc), _cuda(wc).reshape(cout, -1)
    aqd = _cuda(aq); sc = _cuda(np.array([scale], dtype=np.float32)); mask = _mask(cfg)
    table = L.adc_table(spec, s, aqd, mask, alpha_scale=sc)
    wdig, wtiles = L.weight_prepare(spec, wcd)
    oh = cfg.out_hw(hw)
    go = rng.standard_normal((batch, oh * oh, cout)).astype(np.float32)
    ref_gx, ref_gw, ref_ga = O.cim_backward(cfg, go, xc, wc, s_w, s_a, aq, hw)
    god = _cuda(np.ascontiguousarray(go.transpose(0, 2, 1)))
    for name, ff, bf in (("simt", L.FLAG_FORCE_SIMT, L.FLAG_FORCE_SIMT), ("v1", 0, 0), ("v2", L.FLAG_V2, 0)):
        out, state = L.conv_forward(spec, xcd, wcd, wtiles, table, s, mask, save_state=True, flags=ff)
        gxq, gwq, ga = L.conv_backward(spec, god, xcd, wdig, wtiles, state, s, mask, need_alpha=True, flags=bf)
        print(case, name, "gx", rel_err(gxq.cpu().numpy(), ref_gx), "gw", rel_err(gwq.cpu().numpy().reshape(ref_gw.shape), ref_gw), "ga", rel_err(ga.cpu().numpy(), ref_ga))
