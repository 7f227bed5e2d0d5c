"""One forward + one backward of the microbench layer on the v2 kernels: ncu target (tools/gpu_ncu_v2.sh)."""
import argparse, math, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from cim_quantization_b200 import _lib as L

p = argparse.ArgumentParser()
p.add_argument("--xbar", type=int, default=128)
p.add_argument("--adcbits", type=float, default=1.5)
p.add_argument("--batch", type=int, default=256)
p.add_argument("--iters", type=int, default=2)
p.add_argument("--v1", action="store_true")
p.add_argument("--channels", type=int, default=64)
p.add_argument("--hw", type=int, default=32)
p.add_argument("--time", action="store_true", help="print CUDA-event times of the forward and the backward call")
p.add_argument("--timers", action="store_true", help="library built with `make TIMERS=1`: per-role cycle counters of the wgrad kernel (it runs last)")
a = p.parse_args()
adc = int(a.adcbits) if a.adcbits == int(a.adcbits) else a.adcbits
B, C, HW = a.batch, a.channels, a.hw
spec = L.LayerSpec(B, C, HW, C, 3, 1, 1, 3, 1, 3, 1, a.xbar, adc)
g = torch.Generator(device="cuda").manual_seed(0)
x = torch.relu(torch.randn(B, C, HW, HW, device="cuda", generator=g))
w = torch.randn(C, C * 9, device="cuda", generator=g) * math.sqrt(2.0 / (C * 9))
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
go = torch.randn(B, C, HW * HW, device="cuda", generator=g)
flags = 0 if a.v1 else L.FLAG_V2
dbg = None
if a.timers:
    import ctypes
    dbg = torch.zeros(32, dtype=torch.int64, device="cuda")
    L.load().cimq_debug_set_timers(ctypes.c_void_p(dbg.data_ptr()))
ev = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
for it in range(a.iters):
    ev[0].record()
    out, state = L.conv_forward(spec, xc, wc, wtiles, table, s, mask, save_state=True, flags=flags)
    ev[1].record()
    L.conv_backward(spec, go, xc, wdig, wtiles, state, s, mask, need_alpha=aq is not None)
    ev[2].record()
torch.cuda.synchronize()
if a.time:
    print(f"forward {ev[0].elapsed_time(ev[1]) * 1e3:.1f} us   backward {ev[1].elapsed_time(ev[2]) * 1e3:.1f} us")
if dbg is not None:
    d = dbg.cpu().tolist()
    for k, n in {0: "producer wait empty", 1: "producer chunk start (rows, gather)", 2: "producer X tile", 3: "producer G' tile", 7: "producer chunk-end barrier", 16: "producer tile start (pieces, tables)", 17: "producer tile-start barrier", 4: "mma wait full", 6: "mma total"}.items():
        print(f"  wgrad timer {n:36s} {d[k] / 1e3:10.1f} kcycles")
    if d[10]:
        t0 = d[10]
        print("  wgrad block 0 (cycles after entry): prologue done %d, MMA issuer starts %d, last commit %d, epilogue starts %d, all done %d"
              % (d[11] - t0, d[12] - t0, d[13] - t0, d[14] - t0, d[15] - t0))
print("done")
