"""Whole-model golden vector: the REFERENCE ResNet-20 (models/cifar10/resnet.py) after the reference's own
surgery (utils/wrapper/replace_module.py) with the shipped prototxt's CiM settings (w3a3, xbar 128, adc 1.5),
one training-mode forward+backward on a seeded batch of 4 images.  Run in the build container only.

    python tests/golden/make_golden_resnet.py
"""
import math
import os
import sys

import numpy as np
import torch

REF = os.environ.get("CIMQ_REFERENCE", "/root/reference")
HERE = os.path.dirname(os.path.abspath(__file__))
torch.cuda.FloatTensor = lambda *a: torch.FloatTensor(*a)
torch.Tensor.cuda = lambda self, *a, **k: self
sys.path.insert(0, REF)
import models._modules as ref_nn  # noqa: E402
import models.cifar10 as zoo  # noqa: E402
from models._modules import lsq as ref_lsq  # noqa: E402
from utils import wrapper  # noqa: E402


def exact(s, qn, qp):
    k = torch.arange(qn, qp + 1, dtype=torch.float32)
    return bool(((k * s) / s == k).all())


def snap(alpha, g, qn, qp):
    a = alpha.clone()
    for _ in range(8192):
        if exact(ref_lsq.grad_scale(a, g).detach(), qn, qp):
            return a
        a = torch.nextafter(a, torch.full_like(a, float("inf")))
    raise RuntimeError("no exact-recovery step size")


torch.manual_seed(7)
model = zoo.resnet20(pretrained=False)
wrapper.ReplaceModuleTool(model, {'Conv2d': [ref_nn.Conv2dLSQCiM]}, True, nbits_w=3, nbits_a=3, nbits_alpha=8,
                          wbitslice=1, abitslice=1, xbar=128, adcbits=1.5, signed_xbar=False,
                          stochastic_quant=False).replace()
g = torch.Generator().manual_seed(11)
x = torch.randn(2, 3, 32, 32, generator=g)
y = torch.randint(0, 10, (2,), generator=g)
model.train()
convs = [(n, m) for n, m in model.named_modules() if isinstance(m, ref_nn.Conv2dLSQCiM)]
numel = {}
hooks = [m.register_forward_hook(lambda mod, inp, out, n=n: numel.__setitem__(n, inp[0].numel())) for n, m in convs]
model(x)  # lazy init everywhere
with torch.no_grad():
    for n, m in convs:  # exact-recovery step sizes
        qp_a, qn_w, qp_w = 2 ** m.nbits_a - 1, -2 ** (m.nbits_w - 1), 2 ** (m.nbits_w - 1) - 1
        m.alpha_act.copy_(snap(m.alpha_act.data, 1.0 / math.sqrt(numel[n] * qp_a), 0, qp_a))
        m.alpha_weight.copy_(snap(m.alpha_weight.data, 1.0 / math.sqrt(m.weight.numel() * qp_w), qn_w, qp_w))
        m.init_state_cim.fill_(0)
model(x)  # alpha_cim re-initialised by the reference with the snapped step sizes
for h in hooks:
    h.remove()
state = {k: v.detach().clone().numpy() for k, v in model.state_dict().items()}
model.zero_grad()
# teacher-forcing probes: input / output / grad_output / grad_input of representative CiM layers.  (An fp32
# rounding difference can flip an integer code and batch norm then spreads it, so whole-model outputs are
# only comparable layer by layer on identical inputs.)
probe = ["conv1", "layer1.0.conv1", "layer2.0.conv1", "layer3.0.conv1", "layer3.2.conv2"]
mods = dict(model.named_modules())
rec = {}
hooks = []
for n in probe:
    hooks.append(mods[n].register_forward_hook(
        lambda mod, inp, out, n=n: rec.update({f"layer/{n}/in": inp[0].detach().clone().contiguous().numpy(),
                                               f"layer/{n}/out": out.detach().clone().contiguous().numpy()})))
    hooks.append(mods[n].register_full_backward_hook(
        lambda mod, gin, gout, n=n: rec.update({f"layer/{n}/grad_out": gout[0].detach().clone().contiguous().numpy(),
                                                **({f"layer/{n}/grad_in": gin[0].detach().clone().contiguous().numpy()}
                                                   if gin[0] is not None else {})})))
logits = model(x)
loss = torch.nn.functional.cross_entropy(logits, y)
loss.backward()
for h in hooks:
    h.remove()
keep = tuple(probe) + ("linear", "bn1")
grads = {n: p.grad.detach().numpy() for n, p in model.named_parameters() if n.rsplit(".", 1)[0] in keep}
out = {"x": x.numpy(), "y": y.numpy(), "logits": logits.detach().numpy(), "loss": np.float32(loss.item())}
out.update({"state/" + k: v for k, v in state.items()})
out.update({"grad/" + k: v for k, v in grads.items()})
out.update(rec)
np.savez_compressed(os.path.join(HERE, "resnet20_w3a3_x128_tern.npz"), **out)
print("loss", loss.item(), "logits", logits[0, :4].tolist(), "probes", sorted(rec)[:4])
