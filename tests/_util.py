"""Shared helpers for the test-suite (golden loading, tolerances)."""
import glob
import os

import numpy as np

from oracle import cim_oracle as O

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def golden_names():
    """Single-layer golden cases (the whole-model vector resnet20_*.npz has its own test)."""
    names = (os.path.splitext(os.path.basename(p))[0] for p in glob.glob(os.path.join(GOLDEN_DIR, "*.npz")))
    return sorted(n for n in names if not n.startswith(("resnet", "h6_", "h1_")))


def load_golden(name):
    d = dict(np.load(os.path.join(GOLDEN_DIR, name + ".npz")))
    cin, cout, k, stride, pad, hw, batch, nbw, nba, wbs, abs_, xbar = (int(v) for v in d["cfg"])
    adc = float(d["adcbits"])
    adc = int(adc) if adc == int(adc) else adc
    cfg = O.CimConfig(in_channels=cin, out_channels=cout, kernel=k, stride=stride, padding=pad, nbits_w=nbw,
                      nbits_a=nba, nbits_alpha=8, wbitslice=wbs, abitslice=abs_, xbar=xbar, adcbits=adc)
    return cfg, d, hw, batch


def rel_err(a, b):
    """max |a-b| / max |b| -- the 'relative tolerance' used for fp32 tensors (north_star: 1e-5)."""
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    den = max(float(np.max(np.abs(b))), 1e-30)
    return float(np.max(np.abs(a - b))) / den


def rel_err_elem(a, b):
    """Per-element relative error with the tensor's RMS as the floor of the denominator:
    ``max_i |a_i - b_i| / max(|b_i|, rms(b))``.  Stricter than :func:`rel_err` (which divides every element's error
    by the LARGEST magnitude): an element of typical size must itself agree to the tolerance; only elements far
    below the RMS -- sums that cancelled -- are measured against the RMS instead of against themselves."""
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    rms = max(float(np.sqrt(np.mean(b * b))), 1e-30)
    return float(np.max(np.abs(a - b) / np.maximum(np.abs(b), rms)))
