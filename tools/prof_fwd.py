"""Run only the conv forward (and optionally backward) kernels of the microbench layer: ncu target."""
import argparse, math, sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from cim_quantization_b200 import _lib as L

p = argparse.ArgumentParser()
p.add_argument("--xbar", type=int, default=128)
p.add_argument("--adcbits", type=float, default=1.5)
p.add_argument("--batch", type=int, default=256)
p.add_argument("--iters", type=int, default=3)
p.add_argument("--bwd", action="store_true")
p.add_argument("--quant", action="store_true")
p.add_argument("--no-state", action="store_true")
p.add_argument("--timers", action="store_true")
p.add_argument("--timers-wgrad", action="store_true", help="with --timers --bwd: label the counters of the wgrad kernel (it runs last)")
a = p.parse_args()
adc = int(a.adcbits) if a.adcbits == int(a.adcbits) else a.adcbits
B, C, HW = a.batch, 64, 32
spec = L.LayerSpec(B, C, HW, C, 3, 1, 1, 3, 1, 3, 1, a.xbar, adc)
info = L.layer_info(spec)
g = torch.Generator(device="cuda").manual_seed(0)
x = torch.relu(torch.randn(B, C, HW, HW, device="cuda", generator=g))
w = torch.randn(C, C * 9, device="cuda", generator=g) * math.sqrt(2.0 / (C * 9))
s = torch.stack([2 * x.abs().mean() / math.sqrt(7), 2 * w.abs().mean() / math.sqrt(3)]).float()
xc = L.lsq_quantize(x, s[0:1], 0, 7)
wc = L.lsq_quantize(w, s[1:2], -4, 3)
mask = torch.tensor([[1, 2, 4], [2, 4, 8], [4, 8, 16]], dtype=torch.int8, device="cuda")
aq = None
if adc in (1, 1.5):
    sums = L.conv_psum_abs_sums(spec, xc, wc).double()
    aq = (2.0 * sums / (B * HW * HW) * float(s[0]) * float(s[1])).float().clamp_min(1e-6).contiguous()
table = L.adc_table(spec, s, aq, mask)
wdig, wtiles = L.weight_prepare(spec, wc)
go = torch.randn(B, C, HW * HW, device="cuda", generator=g)
dbg = None
if a.timers:
    import ctypes
    dbg = torch.zeros(16, dtype=torch.int64, device="cuda")
    L.load().cimq_debug_set_timers(ctypes.c_void_p(dbg.data_ptr()))
ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
for it in range(a.iters + 1):
    if it == 1:
        torch.cuda.synchronize(); ev[0].record()
    out, state = L.conv_forward(spec, xc, wc, wtiles, table, s, mask, save_state=not a.no_state)
    if a.bwd:
        L.conv_backward(spec, go, xc, wdig, wtiles, state, s, mask, need_alpha=aq is not None)
    if a.quant:
        L.lsq_quantize(x, s[0:1], 0, 7)
        L.lsq_backward(go.view_as(x), x, s[0:1], 0, 7, 1e-3)
ev[1].record(); torch.cuda.synchronize()
if dbg is not None:
    d = dbg.cpu().tolist()
    names = {0: "producer wait empty", 1: "producer produce", 4: "mma wait full", 5: "mma wait tmem-empty", 6: "mma total",
             8: "epilogue wait tmem-full", 9: "epilogue compute", 10: "epilogue table+barrier", 11: "epilogue state store"}
    if a.bwd:  # the dgrad kernel runs after the forward and overwrites these slots with its own counters
        names = {0: "dgrad producer wait empty", 1: "dgrad producer produce", 4: "dgrad mma wait full",
                 5: "dgrad mma wait tmem-empty", 6: "dgrad mma total", 8: "dgrad epilogue wait tmem-full",
                 9: "dgrad epilogue compute"}
    if a.bwd and a.timers_wgrad:
        names = {0: "wgrad producer wait empty", 1: "wgrad producer stage rows", 2: "wgrad producer X tile",
                 3: "wgrad producer G' tile", 4: "wgrad mma wait full", 6: "wgrad mma total"}
    for k, n in names.items():
        print(f"  timer {n:28s} {d[k] / 1e3:10.1f} kcycles")
print("avg ms per iter:", ev[0].elapsed_time(ev[1]) / a.iters, "codes nonzero frac", float((xc != 0).float().mean()))
