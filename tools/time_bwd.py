"""Device timing of the conv backward parts at the microbench shape: v1 (uint32 state words) vs v2 (byte planes)."""
import argparse, math, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from cim_quantization_b200 import _lib as L

p = argparse.ArgumentParser()
p.add_argument("--xbar", type=int, default=128)
p.add_argument("--adcbits", type=float, default=1.5)
p.add_argument("--batch", type=int, default=256)
p.add_argument("--cin", type=int, default=64)
p.add_argument("--cout", type=int, default=64)
p.add_argument("--hw", type=int, default=32)
p.add_argument("--iters", type=int, default=5)
p.add_argument("--only", default="")
a = p.parse_args()
adc = int(a.adcbits) if a.adcbits == int(a.adcbits) else a.adcbits
B, C, CO, HW = a.batch, a.cin, a.cout, a.hw
spec = L.LayerSpec(B, C, HW, CO, 3, 1, 1, 3, 1, 3, 1, a.xbar, adc)
info = L.layer_info(spec)
g = torch.Generator(device="cuda").manual_seed(0)
x = torch.relu(torch.randn(B, C, HW, HW, device="cuda", generator=g))
w = torch.randn(CO, C * 9, device="cuda", generator=g) * math.sqrt(2.0 / (C * 9))
s = torch.stack([2 * x.abs().mean() / math.sqrt(7), 2 * w.abs().mean() / math.sqrt(3)]).float()
xc = L.lsq_quantize(x, s[0:1], 0, 7)
wc = L.lsq_quantize(w, s[1:2], -4, 3)
mask = torch.tensor([[1, 2, 4], [2, 4, 8], [4, 8, 16]], dtype=torch.int8, device="cuda")
aq = sc = None
if adc in (1, 1.5):
    sums = L.conv_psum_abs_sums(spec, xc, wc).double()
    a0 = (2.0 * sums / (B * HW * HW) * float(s[0]) * float(s[1])).float().clamp_min(1e-6).contiguous()
    aq, aux = L.alpha_quantize(a0, 1, 255)
    sc = aux[0:1].clone()
table = L.adc_table(spec, s, aq, mask, alpha_scale=sc)
wdig, wtiles = L.weight_prepare(spec, wc)
OH = info.out_hw
go = torch.randn(B, CO, OH * OH, device="cuda", generator=g)
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
res = {}
for gen, flags in (("v1", 0), ("v2", L.FLAG_V2)):
    if flags and not info.tc_v2:
        continue
    if a.only and a.only != gen:
        continue
    out, state = L.conv_forward(spec, xc, wc, wtiles, table, s, mask, save_state=True, flags=flags)
    for name, kw in (("wgrad", dict(need_alpha=False, need_input=False)),
                     ("dgrad", dict(need_alpha=False, need_weight=False)),
                     ("alpha", dict(need_alpha=True, need_input=False, need_weight=False)),
                     ("all", dict(need_alpha=aq is not None))):
        if name == "alpha" and aq is None:
            continue
        ts = []
        for it in range(a.iters + 2):
            flush.zero_()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            r = L.conv_backward(spec, go, xc, wdig, wtiles, state, s, mask, **kw)
            e1.record()
            torch.cuda.synchronize()
            if it >= 2:
                ts.append(e0.elapsed_time(e1))
        res[(gen, name)] = (sum(ts) / len(ts), r)
        print(f"{gen} {name}: avg {res[(gen, name)][0] * 1e3:.1f} us")
if ("v1", "all") in res and ("v2", "all") in res:
    for idx, nm in enumerate(("grad_x", "grad_w", "grad_alpha")):
        r1, r2 = res[("v1", "all")][1][idx], res[("v2", "all")][1][idx]
        if r1 is not None:
            print(nm, "v1 vs v2 max-normalised diff:", ((r1 - r2).abs().max() / r1.abs().max()).item())
