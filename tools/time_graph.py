"""In-graph time of the parts of one layer's backward / forward (no CPU launch overhead): a CUDA graph of `reps`
back-to-back calls, replayed; per-call time vs batch size separates fixed (per launch) from per-tile cost."""
import argparse, math, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from cim_quantization_b200 import _lib as L

p = argparse.ArgumentParser()
p.add_argument("--channels", type=int, default=64)
p.add_argument("--hw", type=int, default=8)
p.add_argument("--batches", default="64,128,256,512,1024")
p.add_argument("--reps", type=int, default=20)
a = p.parse_args()
C, HW = a.channels, a.hw
mask = torch.tensor([[1, 2, 4], [2, 4, 8], [4, 8, 16]], dtype=torch.int8, device="cuda")
for B in [int(b) for b in a.batches.split(",")]:
    spec = L.LayerSpec(B, C, HW, C, 3, 1, 1, 3, 1, 3, 1, 128, 1.5)
    g = torch.Generator(device="cuda").manual_seed(0)
    x = torch.relu(torch.randn(B, C, HW, HW, device="cuda", generator=g))
    w = torch.randn(C, C * 9, device="cuda", generator=g) * math.sqrt(2.0 / (C * 9))
    s = torch.stack([2 * x.abs().mean() / math.sqrt(7), 2 * w.abs().mean() / math.sqrt(3)]).float()
    xc = L.lsq_quantize(x, s[0:1], 0, 7)
    wc = L.lsq_quantize(w, s[1:2], -4, 3)
    sums = L.conv_psum_abs_sums(spec, xc, wc).double()
    a0 = (2.0 * sums / (B * HW * HW) * float(s[0]) * float(s[1])).float().clamp_min(1e-6).contiguous()
    aq, aux = L.alpha_quantize(a0, 1, 255)
    table = L.adc_table(spec, s, aq, mask, alpha_scale=aux[0:1].clone())
    wdig, wtiles = L.weight_prepare(spec, wc)
    go = torch.randn(B, C, HW * HW, device="cuda", generator=g)
    out, state = L.conv_forward(spec, xc, wc, wtiles, table, s, mask, save_state=True, flags=L.FLAG_V2)
    parts = {
        "fwd": lambda: L.conv_forward(spec, xc, wc, wtiles, table, s, mask, save_state=True, flags=L.FLAG_V2),
        "wgrad": lambda: L.conv_backward(spec, go, xc, wdig, wtiles, state, s, mask, need_alpha=False, need_input=False),
        "dgrad": lambda: L.conv_backward(spec, go, xc, wdig, wtiles, state, s, mask, need_alpha=False, need_weight=False),
        "alpha": lambda: L.conv_backward(spec, go, xc, wdig, wtiles, state, s, mask, need_alpha=True, need_input=False, need_weight=False),
    }
    line = f"ch={C} hw={HW} B={B:5d} (tiles {B * HW * HW // 128:5d}):"
    for name, fn in parts.items():
        fn(); torch.cuda.synchronize()
        st = torch.cuda.Stream()
        with torch.cuda.stream(st):
            gr = torch.cuda.CUDAGraph()
            with torch.cuda.graph(gr, stream=st):
                for _ in range(a.reps):
                    fn()
            gr.replay(); torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(st); gr.replay(); gr.replay(); gr.replay(); e1.record(st); torch.cuda.synchronize()
        line += f"  {name} {e0.elapsed_time(e1) * 1e3 / (3 * a.reps):7.1f} us"
    print(line)
