"""Whole-model parity (BASELINE.json config 1): ResNet-20 w3a3 CiM, one training forward+backward, against a
golden vector produced by the reference's own model zoo + model surgery + Conv2dLSQCiM
(tests/golden/make_golden_resnet.py).

The model is loaded from the reference's state dict (same keys).  Representative CiM layers -- the 8-bit first
conv, a 16-channel layer, both stride-2 layers and the last 64-channel layer on 8x8 images -- are checked with
teacher forcing: the layer gets the reference's input and grad_output and must reproduce the reference's
output, grad_input and parameter gradients to the per-layer tolerance (1e-5).  End-to-end logits are printed
for information only: an fp32 rounding difference can flip one integer activation code, batch norm over a batch
of 2 spreads it to every pixel, and 19 quantised layers amplify it, so the two runs are not expected to agree
to 1e-5 at the output."""
import os

import numpy as np
import pytest
import torch

from tests._util import GOLDEN_DIR, rel_err

pytestmark = pytest.mark.gpu
TOL = 1e-5


def test_resnet20_w3a3_layers_match_reference():
    from cim_quantization_b200 import harness
    d = dict(np.load(os.path.join(GOLDEN_DIR, "resnet20_w3a3_x128_tern.npz")))
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    model = harness.convert_to_cim(harness.resnet20(), nbits_w=3, nbits_a=3, xbar=128, adcbits=1.5)
    state = {k[len("state/"):]: torch.from_numpy(v) for k, v in d.items() if k.startswith("state/")}
    model.load_state_dict(state, strict=True)  # same keys as the reference model
    model = model.cuda().train()
    mods = dict(model.named_modules())
    layers = sorted({k.split("/")[1] for k in d if k.startswith("layer/")})
    assert len(layers) == 5
    for name in layers:
        m = mods[name]
        for p in m.parameters():
            p.grad = None
        xin = torch.from_numpy(d[f"layer/{name}/in"]).cuda().requires_grad_(f"layer/{name}/grad_in" in d)
        out = m(xin)
        assert rel_err(out.detach().cpu().numpy(), d[f"layer/{name}/out"]) < TOL, name
        out.backward(torch.from_numpy(d[f"layer/{name}/grad_out"]).cuda())
        if f"layer/{name}/grad_in" in d:
            assert rel_err(xin.grad.cpu().numpy(), d[f"layer/{name}/grad_in"]) < TOL, name
        assert rel_err(m.weight.grad.cpu().numpy(), d[f"grad/{name}.weight"]) < TOL, name
        assert rel_err(m.alpha_cim.grad.cpu().numpy(), d[f"grad/{name}.alpha_cim"]) < TOL, name
        for a in ("alpha_act", "alpha_weight"):  # sums of cancelling terms: tolerance relative to the larger of the two
            got, ref = float(getattr(m, a).grad.item()), float(d[f"grad/{name}.{a}"].reshape(()))
            assert abs(got - ref) <= 1e-3 * max(abs(ref), abs(got)) + 1e-7, (name, a, got, ref)
    # end to end, information only
    for p in model.parameters():
        p.grad = None
    logits = model(torch.from_numpy(d["x"]).cuda())
    loss = torch.nn.functional.cross_entropy(logits, torch.from_numpy(d["y"]).cuda())
    print(f"end-to-end: loss {loss.item():.4f} (reference {float(d['loss']):.4f}), logits rel err "
          f"{rel_err(logits.detach().cpu().numpy(), d['logits']):.3e}")
    assert torch.isfinite(loss)
