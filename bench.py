#!/usr/bin/env python
"""Benchmark of the CiM-aware quantized convolution hot path (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
                    [--xbar 128] [--adcbits 1.5] [--batch 256]

Workload (BASELINE.json configs[1]): one Conv2dLSQCiM layer, 3x3, 64->64 channels, 32x32 images,
batch 256 per GPU, w3a3, 1-bit slices; a *step* is one forward + backward of the layer through the
module surface (LSQ quantisation of x and w, CiM conv forward, backward to x, w and the three step
sizes).  ``value`` = algorithmic TOPS of the whole job: (2*NSW*NSA + 2*NSW + 2*NSA) * B*L*F*Cout ops
per step per GPU divided by the device time of K steps, max over ranks.

N > 1 (torchrun, one rank per GPU): every rank runs its own batch shard (weak scaling) and the
weight / step-size gradients are all-reduced with NCCL inside the timed step, as data-parallel
training does.

Every line also carries (BASELINE.json's second metric and configs 2 / 4):
  ``train``          ResNet-20 w3a3 CiM training step (forward, loss, backward, gradient all-reduce, SGD) at GLOBAL
                     batch 2048 split evenly over the N GPUs (strong scaling): img/s, ms/step, launches/step;
  ``matrix``         (N = 1) forward / backward device time of the microbench layer for crossbar rows {64,128,256} x
                     ADC bits {1, 1.5, 2, 3, 4};
  ``reference_cuda`` (N = 1) the UNMODIFIED reference module (baseline/_ref) on the same GPU at the same shape: what a
                     user of the reference gets today.

``--impl reference`` times the reference's own ``Conv2dLSQCiM`` (unmodified, from baseline/_ref) on the host cores
with the same metric -- or, only if that copy is missing, the numpy restatement under oracle/ (the one other place
besides the ``cpu_baseline`` leg where bench.py executes anything under oracle/).  It runs a bounded sample
(``config.cpu_sample_batch`` images per step, rate-normalised) and says so in ``config``.
"""
import argparse
import json
import math
import os
import statistics
import subprocess
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

METRIC = "cim_conv_fwd_bwd_tops"
UNIT = "TOPS"


def parse_args():
    p = argparse.ArgumentParser()
    p.add_argument("--gpus", type=int, default=1)
    p.add_argument("--steps", type=int, default=20)
    p.add_argument("--warmup", type=int, default=5)
    p.add_argument("--impl", default="ours", choices=["ours", "reference"])
    p.add_argument("--xbar", type=int, default=128)
    p.add_argument("--adcbits", type=float, default=1.5)
    p.add_argument("--batch", type=int, default=256, help="images per GPU")
    p.add_argument("--nbits", type=int, default=3)
    p.add_argument("--channels", type=int, default=64)
    p.add_argument("--hw", type=int, default=32)
    p.add_argument("--cpu-sample-batch", type=int, default=8)
    p.add_argument("--no-cpu-baseline", action="store_true")
    p.add_argument("--no-graph", action="store_true", help="launch eagerly instead of replaying a CUDA graph")
    p.add_argument("--no-train", action="store_true", help="skip the ResNet-20 training block")
    p.add_argument("--no-matrix", action="store_true", help="skip the crossbar x ADC matrix and the reference-on-GPU leg")
    p.add_argument("--train-global-batch", type=int, default=2048)
    p.add_argument("--train-steps", type=int, default=10)
    return p.parse_args()


def _traffic_bytes(traffic, key):
    """DRAM bytes per launch of a kernel from the committed ncu capture (profiles/ncu_traffic.json), or None."""
    e = traffic.get(key)
    return float(e["total_MB"]) * 1e6 if isinstance(e, dict) and "total_MB" in e else None


def layer_ops(batch, channels, hw, nbits):
    """Algorithmic integer/float ops of one step on one GPU."""
    mkn = batch * hw * hw * (channels * 9) * channels
    ns = nbits  # 1-bit slices: NSW = NSA = nbits
    fwd = 2 * ns * ns * mkn
    bwd = 2 * ns * mkn + 2 * ns * mkn
    return fwd, bwd


def workload_config(a):
    adc = int(a.adcbits) if a.adcbits == int(a.adcbits) else a.adcbits
    return {"workload": f"CiM conv layer microbench 3x3 {a.channels}->{a.channels} {a.hw}x{a.hw} "
                        f"w{a.nbits}a{a.nbits} xbar{a.xbar} adcbits{adc} fwd+bwd",
            "batch_per_gpu": a.batch, "xbar": a.xbar, "adcbits": adc,
            "l2": "working set (x 67 MB, grad_out 67 MB, grad_x 67 MB, ADC state 419 MB) exceeds the 126 MB L2; no extra flush"}


# --------------------------------------------------------------------------------------------------
# CPU arm: the oracle port of the reference algorithm
# --------------------------------------------------------------------------------------------------
def cpu_step_factory(a, batch):
    import numpy as np
    from oracle import cim_oracle as O
    adc = int(a.adcbits) if a.adcbits == int(a.adcbits) else a.adcbits
    cfg = O.CimConfig(in_channels=a.channels, out_channels=a.channels, kernel=3, stride=1, padding=1,
                      nbits_w=a.nbits, nbits_a=a.nbits, wbitslice=1, abitslice=1, xbar=a.xbar, adcbits=adc)
    rng = np.random.default_rng(0)
    x = np.maximum(rng.standard_normal((batch, a.channels, a.hw, a.hw)), 0).astype(np.float32)
    w = (rng.standard_normal((a.channels, a.channels, 3, 3)) * math.sqrt(2.0 / (a.channels * 9))).astype(np.float32)
    gy = rng.standard_normal((batch, a.channels, a.hw, a.hw)).astype(np.float32)
    aa, aw = O.init_step_size(x, cfg.qp_a), O.init_step_size(w, cfg.qp_w)
    s_a = O.grad_scale_value(aa, 1.0 / math.sqrt(x.size * cfg.qp_a))
    s_w = O.grad_scale_value(aw, 1.0 / math.sqrt(w.size * cfg.qp_w))
    ac = None
    if cfg.has_alpha_cim:
        ac = O.init_alpha_cim(cfg, O.lsq_codes(x, s_a, 0, cfg.qp_a), O.lsq_codes(w, s_w, cfg.qn_w, cfg.qp_w), s_w, s_a)

    def step():
        return O.module_forward_backward(cfg, x, w, aa, aw, ac, gy)

    return step


def reference_root():
    """baseline/_ref: the unmodified reference tree copied next to the repo by tools/make_baseline_ref.sh
    (git-ignored; it travels to the GPU box).  /root/reference itself is never read by bench.py."""
    cand = os.path.join(ROOT, "baseline", "_ref")
    return cand if os.path.isfile(os.path.join(cand, "models", "_modules", "lsq.py")) else None


def ref_cpu_step_factory(a, batch):
    """The reference's own Conv2dLSQCiM (models/_modules/lsq.py, unmodified) forward+backward on the host
    cores, behind the two-line CPU shim of SURVEY 8c (the reference hard-codes torch.cuda allocations).
    Only ever called in a `--impl reference` process: the shim patches torch.Tensor.cuda."""
    import torch
    torch.cuda.FloatTensor = lambda *t: torch.FloatTensor(*t)  # lsq.py:64,169,215,336
    torch.Tensor.cuda = lambda self, *t, **k: self              # lsq.py:169
    sys.path.insert(0, reference_root())
    import models._modules as ref_nn
    torch.set_num_threads(os.cpu_count() or 1)
    g = torch.Generator().manual_seed(0)
    adc = int(a.adcbits) if a.adcbits == int(a.adcbits) else a.adcbits
    m = ref_nn.Conv2dLSQCiM(a.channels, a.channels, (3, 3), (1, 1), (1, 1), (1, 1), 1, False, nbits_w=a.nbits,
                            nbits_a=a.nbits, nbits_alpha=8, wbitslice=1, abitslice=1, xbar=a.xbar, adcbits=adc,
                            signed_xbar=False, stochastic_quant=False)
    with torch.no_grad():
        m.weight.copy_(torch.randn(m.weight.shape, generator=g) * math.sqrt(2.0 / (a.channels * 9)))
    x = torch.relu(torch.randn(batch, a.channels, a.hw, a.hw, generator=g))
    gy = torch.randn(batch, a.channels, a.hw, a.hw, generator=g)
    m.train()
    m(x)  # lazy inits (lsq.py:532-563) outside the timed steps

    def step():
        xi = x.clone().requires_grad_(True)
        for p_ in m.parameters():
            p_.grad = None
        m(xi).backward(gy)
        return xi.grad

    return step


def cpu_step(a, batch):
    """(step function, kind): the unmodified reference when baseline/_ref is there, else the numpy port."""
    if reference_root() is not None:
        try:
            return ref_cpu_step_factory(a, batch), "reference"
        except Exception as e:  # e.g. a torch build the reference cannot import under: fall back, and say so
            print(f"bench: reference import failed ({e!r}); timing the oracle port", file=sys.stderr)
    return cpu_step_factory(a, batch), "port"


def run_reference(a):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return  # the CPU arm runs on rank 0 only
    cores = os.cpu_count() or 1
    os.environ.setdefault("OMP_NUM_THREADS", str(cores))
    b = a.cpu_sample_batch
    step, kind = cpu_step(a, b)
    for _ in range(min(a.warmup, 1)):
        step()
    t0 = time.perf_counter()
    for _ in range(a.steps):
        step()
    dt = time.perf_counter() - t0
    fwd, bwd = layer_ops(b, a.channels, a.hw, a.nbits)
    value = (fwd + bwd) * a.steps / dt / 1e12
    what = ("the reference's own Conv2dLSQCiM module on CPU (baseline/_ref, unmodified, CUDA-allocation shim)"
            if kind == "reference" else "numpy port of the reference algorithm (oracle/)")
    sample = f"{b} images per step of the same layer, {what}, all host threads"
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": a.gpus,
            "steps": a.steps, "warmup": min(a.warmup, 1), "ms_per_step": dt / a.steps * 1e3,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32 (holding integers)",
            "data": "synthetic",
            "config": dict(workload_config(a), batch_per_gpu=b, cpu_sample_batch=b,
                           note=f"bounded sample: {b} images per step instead of {a.batch} (the reference's 6-D "
                                "temporaries at 256 images need > 30 GB and minutes per step on CPU); TOPS is a rate, "
                                "so the figure is comparable"),
            "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": kind, "sample": sample},
            "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line))


# --------------------------------------------------------------------------------------------------
# GPU arm
# --------------------------------------------------------------------------------------------------
class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled every 200 ms while the timed region runs."""
    QUERY = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
             "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
             "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.proc, self.path = index, None, None

    def __enter__(self):
        try:
            f = tempfile.NamedTemporaryFile("w", suffix=".csv", delete=False)
            self.path = f.name
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.QUERY}",
                                          "--format=csv,noheader,nounits", "-lms", "200"], stdout=f,
                                         stderr=subprocess.DEVNULL)
        except Exception:
            self.proc = None
        return self

    def __exit__(self, *exc):
        if self.proc is not None:
            time.sleep(0.25)
            self.proc.terminate()
            try:
                self.proc.wait(timeout=5)
            except Exception:
                self.proc.kill()

    def summary(self):
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        if not self.path or not os.path.exists(self.path):
            return out
        sm, mx, reasons = [], [], set()
        for row in open(self.path):
            cols = [c.strip() for c in row.split(",")]
            if len(cols) < 7:
                continue
            try:
                sm.append(float(cols[0]))
                mx.append(float(cols[1]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"),
                               cols[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        os.unlink(self.path)
        if sm:
            out.update(sm_mhz=statistics.median(sm), sm_max_mhz=max(mx), reasons=sorted(reasons), samples=len(sm))
        return out


def event_time_ms(fn, iters, torch, graph=True):
    """Average device time of fn() over `iters` launches: CUDA events on the current stream around one replay
    of a CUDA graph holding `iters` back-to-back calls (no host launch gaps between them)."""
    start, stop = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    fn()
    torch.cuda.synchronize()
    g = None
    if graph:
        try:
            side = torch.cuda.Stream()
            side.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(side):
                fn()
            torch.cuda.current_stream().wait_stream(side)
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g):
                for _ in range(iters):
                    fn()
            g.replay()
        except Exception:
            g = None
    torch.cuda.synchronize()
    start.record()
    if g is not None:
        g.replay()
    else:
        for _ in range(iters):
            fn()
    stop.record()
    torch.cuda.synchronize()
    return start.elapsed_time(stop) / iters


# --------------------------------------------------------------------------------------------------
# kernel-level pieces (C-ABI calls on resident tensors), shared by the per-kernel block and the matrix
# --------------------------------------------------------------------------------------------------
def prepare_layer_tensors(torch, L, layer, x, gy):
    """Everything the C-ABI conv calls of one layer need, built the way the module builds it (v2 kernels where the
    layer is covered, CIMQ_FLAG_V2)."""
    spec = layer._spec(x)
    info = L.layer_info(spec)
    nb_a, nb_w = layer.nbits_a, layer.nbits_w
    qp_a, qn_w, qp_w = 2 ** nb_a - 1, -(2 ** (nb_w - 1)), 2 ** (nb_w - 1) - 1
    B, C = x.shape[0], layer.out_channels
    with torch.no_grad():
        s = L.step_sizes(layer.alpha_act.data, layer.alpha_weight.data, 1.0 / math.sqrt(x.numel() * qp_a),
                         1.0 / math.sqrt(layer.weight.numel() * qp_w))
        xd, wd = x.detach(), layer.weight.detach().contiguous()
        xc = L.lsq_quantize(xd, s[0:1], 0, qp_a)
        wc = L.lsq_quantize(wd, s[1:2], qn_w, qp_w).view(C, -1)
        mask = layer.binary_mask.reshape(info.NSW, info.NSA).contiguous()
        aq = scale = None
        if layer.alpha_cim is not None:
            aq, scale = layer._alpha_q()
            aq, scale = aq.detach().contiguous(), scale.detach()
        v2 = L.v2_usable(spec, aq is not None, scale)
        table = L.adc_table(spec, s, aq, mask, alpha_scale=scale if v2 else None)
        wdig, wtiles = L.weight_prepare(spec, wc, want_digits=not info.tc_backward)
    return dict(spec=spec, info=info, s=s, xd=xd, xc=xc, wc=wc, mask=mask, aq=aq, table=table, wdig=wdig,
                wtiles=wtiles, go=gy.reshape(B, C, -1), flags=L.FLAG_V2 if v2 else 0, v2=bool(v2),
                qp_a=qp_a)


def time_conv_kernels(torch, L, t, iters, parts=True):
    """Device time (ms) of the conv forward (training / inference) and of the backward and its parts."""
    fw = lambda save: L.conv_forward(t["spec"], t["xc"], t["wc"], t["wtiles"], t["table"], t["s"], t["mask"],
                                     save_state=save, flags=t["flags"])
    out, state = fw(True)
    bw = lambda **kw: L.conv_backward(t["spec"], t["go"], t["xc"], t["wdig"], t["wtiles"], state, t["s"], t["mask"], **kw)
    has_alpha = t["aq"] is not None
    r = {"fwd_train": event_time_ms(lambda: fw(True), iters, torch),
         "fwd_infer": event_time_ms(lambda: fw(False), iters, torch),
         "bwd": event_time_ms(lambda: bw(need_alpha=has_alpha), iters, torch)}
    if parts:
        r["wgrad"] = event_time_ms(lambda: bw(need_alpha=False, need_input=False), iters, torch)
        r["dgrad"] = event_time_ms(lambda: bw(need_alpha=False, need_weight=False), iters, torch)
        r["alpha"] = (event_time_ms(lambda: bw(need_alpha=True, need_input=False, need_weight=False), iters, torch)
                      if has_alpha else 0.0)
    del out, state
    return r


def config_matrix(torch, cq, L, a, dev, int8_peak, bf16_peak):
    """BASELINE.json config 2: the microbench layer for crossbar rows {64,128,256} x ADC bits {1,1.5,2,3,4}:
    forward (training) and backward device time, algorithmic rate and fraction of the tensor roofline."""
    C, HW, B = a.channels, a.hw, a.batch
    fwd_ops, bwd_ops = layer_ops(B, C, HW, a.nbits)
    g = torch.Generator(device=dev).manual_seed(7)
    x = torch.relu(torch.randn(B, C, HW, HW, device=dev, generator=g))
    gy = torch.randn(B, C, HW, HW, device=dev, generator=g)
    rows = []
    for xbar in (64, 128, 256):
        for adc in (1, 1.5, 2, 3, 4):
            torch.manual_seed(11)
            layer = cq.Conv2dLSQCiM(C, C, (3, 3), (1, 1), (1, 1), (1, 1), 1, False, nbits_w=a.nbits, nbits_a=a.nbits,
                                    nbits_alpha=8, wbitslice=1, abitslice=1, xbar=xbar, adcbits=adc).to(dev).train()
            with torch.no_grad():
                torch.nn.init.kaiming_normal_(layer.weight)
                layer(x)  # lazy initialisation
                t = prepare_layer_tensors(torch, L, layer, x, gy)
                r = time_conv_kernels(torch, L, t, 3, parts=False)
            step_ms = r["fwd_train"] + r["bwd"]
            rows.append({"xbar": xbar, "adcbits": adc, "kernels": "v2" if t["v2"] else "v1",
                         "fwd_ms": round(r["fwd_train"], 4), "bwd_ms": round(r["bwd"], 4),
                         "fwd_TOPS": round(fwd_ops / r["fwd_train"] / 1e9, 1),
                         "fwd_frac_int8_tc": round(fwd_ops / r["fwd_train"] / 1e9 / int8_peak, 4),
                         "bwd_TFLOPs": round(bwd_ops / r["bwd"] / 1e9, 1),
                         "bwd_frac_bf16_tc": round(bwd_ops / r["bwd"] / 1e9 / bf16_peak, 4),
                         "fwd_bwd_TOPS": round((fwd_ops + bwd_ops) / step_ms / 1e9, 1)})
            del layer, t
            torch.cuda.empty_cache()
    return rows


def reference_cuda_leg(torch, a, dev):
    """The UNMODIFIED reference module (baseline/_ref/models/_modules/lsq.py) on this GPU at the bench shape: its
    own torch-CUDA path (6-D temporaries, python slice loops).  None if baseline/_ref is missing."""
    root = reference_root()
    if root is None:
        return None
    saved = list(sys.path)
    try:
        sys.path.insert(0, root)
        import importlib
        ref_nn = importlib.import_module("models._modules")
        adc = int(a.adcbits) if a.adcbits == int(a.adcbits) else a.adcbits
        C, HW, B = a.channels, a.hw, a.batch
        torch.manual_seed(5)
        m = ref_nn.Conv2dLSQCiM(C, C, (3, 3), (1, 1), (1, 1), (1, 1), 1, False, nbits_w=a.nbits, nbits_a=a.nbits,
                                nbits_alpha=8, wbitslice=1, abitslice=1, xbar=a.xbar, adcbits=adc, signed_xbar=False,
                                stochastic_quant=False).to(dev).train()
        x = torch.relu(torch.randn(B, C, HW, HW, device=dev)).requires_grad_(True)
        gy = torch.randn(B, C, HW, HW, device=dev)
        m(x.detach())  # lazy inits

        def step():
            for p_ in m.parameters():
                p_.grad = None
            x.grad = None
            m(x).backward(gy)

        step()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        n = 3
        e0.record()
        for _ in range(n):
            step()
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / n
        fwd_ops, bwd_ops = layer_ops(B, C, HW, a.nbits)
        peak_gb = torch.cuda.max_memory_allocated(dev) / 2 ** 30
        del m, x, gy
        torch.cuda.empty_cache()
        return {"ms_per_step": ms, "TOPS": (fwd_ops + bwd_ops) / ms / 1e9, "steps": n, "batch": B,
                "what": "unmodified reference Conv2dLSQCiM (baseline/_ref) forward+backward on the same GPU, eager torch",
                "peak_mem_GiB": round(peak_gb, 1)}
    except Exception as e:  # e.g. out of memory at other shapes
        return {"error": repr(e)[:300]}
    finally:
        sys.path[:] = saved
        for k in [k for k in sys.modules if k == "models" or k.startswith("models.")]:
            sys.modules.pop(k, None)


def train_block(torch, dist, a, world, rank, dev):
    """BASELINE.json config 4: ResNet-20 w3a3 CiM (19 Conv2dLSQCiM, xbar 128, ternary ADC) training step on synthetic
    CIFAR-shaped data, GLOBAL batch a.train_global_batch split evenly over the ranks: forward, cross-entropy, backward,
    one flat NCCL all-reduce of all gradients (N > 1), SGD.  Reference loop: examples/__init__.py:390-462, DDP at
    :693-716.  Device time by CUDA events around K graph replays, max over ranks."""
    from cim_quantization_b200 import harness, _lib as L
    from cim_quantization_b200.distributed import FlatGradAllReducer, broadcast_parameters
    gb = a.train_global_batch
    if gb % world != 0:
        return {"error": f"global batch {gb} does not divide over {world} ranks"}
    bpg = gb // world
    torch.manual_seed(0)
    model = harness.convert_to_cim(harness.resnet20(1), nbits_w=3, nbits_a=3, xbar=128, adcbits=1.5).to(dev).train()
    torch.manual_seed(1 + rank)
    x = torch.randn(bpg, 3, 32, 32, device=dev)
    y = torch.randint(0, 10, (bpg,), device=dev)
    crit = torch.nn.CrossEntropyLoss()
    model(x)  # lazy initialisation of all step sizes on the first batch (lsq.py:532-563)
    if world > 1:
        broadcast_parameters(model, 0)
    opt = torch.optim.SGD(harness.sgd_param_groups(model), lr=0.01, momentum=0.9)
    reducer = FlatGradAllReducer(model.parameters())

    def step():
        opt.zero_grad(set_to_none=True)
        loss = crit(model(x), y)
        loss.backward()
        if world > 1:
            reducer.all_reduce_()
        opt.step()
        return loss

    L.launch_counter = 0
    step()
    launches = L.launch_counter
    for _ in range(2):
        step()
    torch.cuda.synchronize()
    runner, graphed = step, False
    if not a.no_graph:
        try:
            side = torch.cuda.Stream()
            side.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(side):
                step()
            torch.cuda.current_stream().wait_stream(side)
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g):
                step()
            g.replay()
            torch.cuda.synchronize()
            runner, graphed = g.replay, True
        except Exception as e:  # pragma: no cover
            print(f"[bench] train graph capture failed ({e!r}); eager steps", file=sys.stderr)
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(a.train_steps):
        runner()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    if world > 1:
        t = torch.tensor([ms], device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
    nconv = sum(1 for m_ in model.modules() if m_.__class__.__name__ == "Conv2dLSQCiM")
    return {"metric": "resnet20_w3a3_cim_train_img_per_s", "img_per_s": gb * a.train_steps / (ms * 1e-3), "unit": "img/s",
            "ms_per_step": ms / a.train_steps, "steps": a.train_steps, "global_batch": gb, "batch_per_gpu": bpg,
            "n_gpus": world, "scaling": "strong", "cim_convs": nconv, "cuda_graph": graphed,
            "gpu_launches_per_step": int(launches), "grad_allreduce_bytes": int(reducer.nbytes),
            "model": "ResNet-20 CIFAR (harness.resnet20; option-A shortcuts), w3a3, first conv w8a8, xbar 128, adcbits 1.5, "
                     "SGD lr 0.01 momentum 0.9 wd 1e-4, synthetic 32x32 data"}


def run_ours(a):
    import torch
    import torch.distributed as dist
    import cim_quantization_b200 as cq
    from cim_quantization_b200 import _lib as L

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py --impl ours needs a CUDA device (there is no CPU fallback)")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    hbm_peak = peaks.get("hbm_gbs", 6650.0)
    bf16_peak = peaks.get("bf16_tflops", 1590.0)
    peak_src = "measured (MEASURED_PEAKS.json)" if peaks else "fallback (B200_PROFILING.md)"

    adc = int(a.adcbits) if a.adcbits == int(a.adcbits) else a.adcbits
    C, HW, B = a.channels, a.hw, a.batch
    torch.manual_seed(1234 + rank)
    layer = cq.Conv2dLSQCiM(C, C, (3, 3), (1, 1), (1, 1), (1, 1), 1, False, nbits_w=a.nbits, nbits_a=a.nbits,
                            nbits_alpha=8, wbitslice=1, abitslice=1, xbar=a.xbar, adcbits=adc).to(dev).train()
    with torch.no_grad():
        torch.nn.init.kaiming_normal_(layer.weight)
    x_host = torch.relu(torch.randn(B, C, HW, HW)).pin_memory()
    gy_host = torch.randn(B, C, HW, HW).pin_memory()
    x = x_host.to(dev).requires_grad_(True)
    gy = gy_host.to(dev)
    layer(x.detach())  # lazy initialisation of the step sizes (first training batch, lsq.py:532-563)
    if world > 1:  # identical parameters on every rank, as DDP's initial broadcast does
        for p in layer.parameters():
            dist.broadcast(p.data, 0)
    params = [p for p in layer.parameters()]
    from cim_quantization_b200.distributed import FlatGradAllReducer
    reducer = FlatGradAllReducer(params)
    flat = reducer.flat

    def compute():
        for p in params:
            p.grad = None
        x.grad = None
        y = layer(x)
        y.backward(gy)
        return y

    def step():
        y = compute()
        if world > 1:  # one flat all-reduce of all parameter gradients (weights + step sizes), DDP average
            reducer.all_reduce_()
        return y

    L.launch_counter = 0
    step()
    launches_per_step = L.launch_counter
    for _ in range(max(a.warmup, 3) - 1):
        step()
    torch.cuda.synchronize()

    runner = step
    graph = None
    if not a.no_graph:
        try:  # replay the kernels of the step as one CUDA graph: no launch gaps, no host work in the timed region.
            # With N > 1 the gradient all-reduce stays outside the graph (NCCL work captured in a graph made the
            # process teardown hang here): graph replay, then the eager flat all-reduce, every step.
            side = torch.cuda.Stream()
            side.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(side):
                compute()
            torch.cuda.current_stream().wait_stream(side)
            graph = torch.cuda.CUDAGraph()
            with torch.cuda.graph(graph):
                compute()
            if world == 1:
                runner = graph.replay
            else:
                captured = [p.grad for p in params]  # the graph rewrites these buffers on every replay

                def runner():
                    graph.replay()
                    for p, g_ in zip(params, captured):
                        p.grad = g_
                    reducer.all_reduce_()
            runner()
            torch.cuda.synchronize()
        except Exception as e:  # pragma: no cover
            print(f"[bench] CUDA graph capture failed ({e!r}); timing eager launches", file=sys.stderr)
            graph, runner = None, step

    # ---- timed region: exactly K steps between barrier + synchronize, CUDA events, max over ranks
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    start, stop = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    with ClockSampler(local) as clk:
        start.record()
        for _ in range(a.steps):
            runner()
        stop.record()
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
    ms = start.elapsed_time(stop)
    if world > 1:
        t = torch.tensor([ms], device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
    fwd_ops, bwd_ops = layer_ops(B, C, HW, a.nbits)
    wgrad_ops = dgrad_ops = bwd_ops / 2  # 2*NSA*MKN and 2*NSW*MKN with NSW == NSA
    ops_step = fwd_ops + bwd_ops
    value = ops_step * world * a.steps / (ms * 1e-3) / 1e12

    # ---- end to end through the module surface with HOST buffers (pinned), copies inside the timed region
    gw_host = torch.empty_like(layer.weight, device="cpu").pin_memory()
    small_host = torch.empty(flat.numel() - layer.weight.numel(), device="cpu").pin_memory()

    # Host batches reach the GPU through the package's HostBatchPipeline (a copy stream, two batches deep): the
    # H2D copy of step n+1 overlaps the kernels of step n.  Every timed step still copies its own inputs from
    # pinned host memory and reads its gradients back; the pipeline starts empty inside the timed region.
    from cim_quantization_b200.harness import HostBatchPipeline

    def e2e_run(nsteps):
        pipe = HostBatchPipeline(dev, depth=2)
        submitted = 0
        for n in range(nsteps):
            while submitted < nsteps and pipe.can_submit():
                pipe.submit((x_host, gy_host))
                submitted += 1
            xin, gyin = pipe.get()
            xin = xin.detach().requires_grad_(True)
            for p in params:
                p.grad = None
            layer(xin).backward(gyin)
            pipe.release()
            if world > 1:
                reducer.all_reduce_()
            gw_host.copy_(layer.weight.grad, non_blocking=True)
            small_host.copy_(torch.cat([p.grad.reshape(-1) for p in params if p is not layer.weight]),
                             non_blocking=True)

    e2e_run(3)
    torch.cuda.synchronize()
    e2e_iters = max(3, min(a.steps, 10))
    if world > 1:
        dist.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    e2e_run(e2e_iters)
    e1.record()
    torch.cuda.synchronize()
    e2e_ms = e0.elapsed_time(e1) / e2e_iters
    if world > 1:  # every rank feeds its own GPU from its own pinned buffers; the job's time is the slowest rank's
        t_ = torch.tensor([e2e_ms], device=dev)
        dist.all_reduce(t_, op=dist.ReduceOp.MAX)
        e2e_ms = float(t_.item())
    e2e = {"value": ops_step * world / (e2e_ms * 1e-3) / 1e12, "unit": UNIT,
           "h2d_bytes_per_step": int(x_host.numel() * 4 + gy_host.numel() * 4),
           "d2h_bytes_per_step": int(gw_host.numel() * 4 + small_host.numel() * 4), "n_gpus": world,
           "ms_per_step": e2e_ms, "steps": e2e_iters,
           "pipeline": "HostBatchPipeline depth 2 (copy stream); starts empty inside the timed region"}


    # ---- ResNet-20 training block (every rank takes part), before the rank-0-only legs
    train = None
    if not a.no_train:
        try:
            train = train_block(torch, dist, a, world, rank, dev)
        except Exception as e:  # pragma: no cover
            train = {"error": repr(e)[:300]}
    if world > 1:
        torch.cuda.synchronize()
        dist.barrier()
        if rank != 0:
            # the training step graph holds NCCL work: tearing the communicator down afterwards can hang, and nothing is
            # left for this rank to do, so leave without the collective teardown
            sys.stdout.flush()
            os._exit(0)

    line = None
    if rank == 0:
        # ---- per-kernel device times (CUDA events, same process, same resident tensors)
        t = prepare_layer_tensors(torch, L, layer, x, gy)
        info = t["info"]
        it = max(3, min(a.steps, 10))
        with torch.no_grad():
            xd, gxq = t["xd"], torch.empty_like(t["xd"])
            t_q = event_time_ms(lambda: L.lsq_quantize(xd, t["s"][0:1], 0, t["qp_a"]), it, torch)
            t_qb = event_time_ms(lambda: L.lsq_backward(gxq, xd, t["s"][0:1], 0, t["qp_a"], 1e-3), it, torch)
            kt = time_conv_kernels(torch, L, t, it)
        t_f, t_fi, t_b, t_bw, t_bx, t_ba = (kt["fwd_train"], kt["fwd_infer"], kt["bwd"], kt["wgrad"], kt["dgrad"],
                                            kt["alpha"])
        gen = "v2" if t["v2"] else "v1"
        n = x.numel()
        int8_peak = 2.0 * bf16_peak  # no measured int8 peak: 2x the measured dense bf16 rate
        psums = float(info.psum_count)
        # conv-minimal HBM bytes: codes + output + weights (forward); + grad_out + grad_x (backward)
        alg_fwd = n * 1.0 + B * C * HW * HW * 4.0 + layer.weight.numel() * 4.0
        alg_dgrad = B * C * HW * HW * 4.0 + n * 4.0 + layer.weight.numel() * 4.0
        alg_wgrad = B * C * HW * HW * 4.0 + n * 1.0 + layer.weight.numel() * 4.0
        kernels = {
            "generation": gen,
            "lsq_quantize_x": {"ms": t_q, "GB/s": 5.0 * n / t_q / 1e6, "frac_hbm": 5.0 * n / t_q / 1e6 / hbm_peak},
            "lsq_backward_x": {"ms": t_qb, "GB/s": 12.0 * n / t_qb / 1e6, "frac_hbm": 12.0 * n / t_qb / 1e6 / hbm_peak},
            "conv_forward_train": {"ms": t_f, "TOPS": fwd_ops / t_f / 1e9, "frac_int8_tc": fwd_ops / t_f / 1e9 / int8_peak},
            "conv_forward_infer": {"ms": t_fi, "TOPS": fwd_ops / t_fi / 1e9,
                                   "frac_int8_tc": fwd_ops / t_fi / 1e9 / int8_peak},
            "conv_backward": {"ms": t_b, "TFLOP/s": bwd_ops / t_b / 1e9, "frac_bf16_tc": bwd_ops / t_b / 1e9 / bf16_peak},
            "conv_backward.wgrad": {"ms": t_bw, "TFLOP/s": wgrad_ops / t_bw / 1e9,
                                    "frac_bf16_tc": wgrad_ops / t_bw / 1e9 / bf16_peak},
            "conv_backward.dgrad+fold": {"ms": t_bx, "TFLOP/s": dgrad_ops / t_bx / 1e9,
                                         "frac_bf16_tc": dgrad_ops / t_bx / 1e9 / bf16_peak},
            "conv_backward.alpha_grad": (
                {"ms": t_ba, "GB/s": (psums / 3.0 + n * 4.0) / t_ba / 1e6,
                 "frac_hbm": (psums / 3.0 + n * 4.0) / t_ba / 1e6 / hbm_peak,
                 "bytes": "state plane C (1 byte per 3 partial sums) + grad_out"} if t_ba > 0 else {"ms": 0.0}),
        }
        # DRAM traffic per launch from the committed ncu capture of the same kernels (profiles/ncu_traffic.json)
        traffic = {}
        try:
            traffic = json.load(open(os.path.join(ROOT, "profiles", "ncu_traffic.json")))
        except Exception:
            pass
        fwd_name = ("conv_v2_kernel (CiM conv forward: tcgen05 kind::f8f6f4 -> fp16 TMEM partial sums, packed-half ADC "
                    "epilogue, kind::f16 A-from-TMEM shift-and-add)" if t["v2"]
                    else "conv_tc_kernel (CiM conv forward, tcgen05 kind::i8 + ADC epilogue)")
        sfx = "_v2" if t["v2"] else ""
        split_note = ("2x that in fp16 (grad_out scaled by a power of two per row / channel and split into two 9-bit "
                      "pieces, exact piece * pass-count products)" if t["v2"] else
                      "3x that in bf16 (hi/mid/lo split of grad)")
        cand = [
            (t_f, {"kernel": fwd_name, "bound": "tensor",
                   "achieved": fwd_ops / t_f / 1e9, "peak": int8_peak, "unit": "TOPS",
                   "frac": fwd_ops / t_f / 1e9 / int8_peak, "traffic": _traffic_bytes(traffic, "conv_forward" + sfx),
                   "algorithmic_bytes": alg_fwd,
                   "peak_source": peak_src + " x2 for 8-bit operands (no measured int8 / fp8 peak)",
                   "note": "algorithmic ops 2*NSW*NSA*MKN, one contraction per slice pair"}),
            (t_bw, {"kernel": ("bwd_weight_tc_kernel<V2> (+ go_scales, finish; CiM conv wgrad, tcgen05 kind::f16, grad_out as "
                               "two fp16 pieces)" if t["v2"] else
                               "bwd_weight_tc_kernel (+ finish; CiM conv wgrad, tcgen05 kind::f16 bf16x3)"), "bound": "tensor",
                    "achieved": wgrad_ops / t_bw / 1e9, "peak": bf16_peak, "unit": "TFLOP/s",
                    "frac": wgrad_ops / t_bw / 1e9 / bf16_peak,
                    "traffic": _traffic_bytes(traffic, "conv_wgrad" + sfx),
                    "algorithmic_bytes": alg_wgrad, "peak_source": peak_src,
                    "note": "algorithmic flops 2*NSA*MKN; the kernel issues " + split_note}),
            (t_bx, {"kernel": ("bwd_input_v2_kernel (+ go_scales; CiM conv dgrad + fold in the epilogue, tcgen05 kind::f16, "
                               "grad_out as two fp16 pieces)" if t["v2"] else
                               "bwd_input_tc_kernel (CiM conv dgrad + fused fold, tcgen05 kind::f16 bf16x3)"), "bound": "tensor",
                    "achieved": dgrad_ops / t_bx / 1e9, "peak": bf16_peak, "unit": "TFLOP/s",
                    "frac": dgrad_ops / t_bx / 1e9 / bf16_peak,
                    "traffic": _traffic_bytes(traffic, "conv_dgrad" + sfx),
                    "algorithmic_bytes": alg_dgrad, "peak_source": peak_src,
                    "note": "algorithmic flops 2*NSW*MKN; the kernel issues " + split_note}),
        ]
        roof = max(cand, key=lambda c: c[0])[1]
        if roof.get("traffic"):
            roof["traffic_over_algorithmic"] = roof["traffic"] / roof["algorithmic_bytes"]
        # the whole step against the roofline the metric names (int tensor core)
        roof["step"] = {"ops": ops_step, "ms": ms / a.steps, "TOPS": value / world,
                        "frac_int8_tc": value / world / int8_peak,
                        "note": "fwd+bwd algorithmic ops of one GPU / step time / (2 x measured bf16 peak)"}
        roof["ms_per_launch"] = max(c[0] for c in cand)

        cpu = None
        if not a.no_cpu_baseline:
            # a separate process: the reference needs torch.Tensor.cuda patched out, which must not leak in here
            try:
                r = subprocess.run([sys.executable, os.path.abspath(__file__), "--impl", "reference", "--steps", "3",
                                    "--warmup", "1", "--xbar", str(a.xbar), "--adcbits", str(a.adcbits), "--nbits",
                                    str(a.nbits), "--channels", str(a.channels), "--hw", str(a.hw),
                                    "--cpu-sample-batch", str(a.cpu_sample_batch)],
                                   capture_output=True, text=True, timeout=600,
                                   env={k: v for k, v in os.environ.items() if k not in ("RANK", "WORLD_SIZE")})
                cpu = json.loads(r.stdout.strip().splitlines()[-1])["cpu_baseline"]
            except Exception as e:
                cpu = {"value": None, "unit": UNIT, "cores": os.cpu_count() or 1, "kind": "port",
                       "sample": f"CPU baseline failed: {e!r}"}
        clocks = clk.summary()
        matrix = ref_cuda = None
        if world == 1 and not a.no_matrix:
            try:
                matrix = config_matrix(torch, cq, L, a, dev, int8_peak, bf16_peak)
            except Exception as e:  # pragma: no cover
                matrix = {"error": repr(e)[:300]}
            ref_cuda = reference_cuda_leg(torch, a, dev)
        line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": a.steps,
                "warmup": max(a.warmup, 3), "ms_per_step": ms / a.steps, "higher_is_better": True, "scaling": "weak",
                "vs_baseline": None,
                "dtype": ("e4m3 digits -> fp16 partial sums in tensor memory (forward); fp32 grad_out as 2 fp16 pieces x exact "
                          "small integers, fp32 accumulation (backward)" if t["v2"] else
                          "u8 x s8 -> s32 partial sums (forward), f32 via bf16x3 (backward)"),
                "data": "synthetic",
                "config": dict(workload_config(a), cuda_graph=graph is not None,
                               tcgen05_forward=bool(info.tc_forward)),
                "clocks": clocks, "e2e": e2e, "gpu_launches": int(launches_per_step * a.steps),
                "roofline": roof, "cpu_baseline": cpu, "kernels": kernels,
                "ops_per_step": {"forward": fwd_ops, "backward": bwd_ops},
                "matrix": matrix, "reference_cuda": ref_cuda}
        line["config"]["kernels"] = gen
    if line is not None:
        line["train"] = train
        print(json.dumps(line))
        sys.stdout.flush()
    if world > 1:
        torch.cuda.synchronize()
        os._exit(0)


def main():
    a = parse_args()
    if a.impl == "reference":
        run_reference(a)
    else:
        run_ours(a)


if __name__ == "__main__":
    main()
