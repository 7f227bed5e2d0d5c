"""Plain LSQ modules of the reference's public surface (ActLSQ / LinearLSQ / Conv2dLSQ, lsq.py:389-436, 591-662)
on the fused quantiser kernels, against an inline torch restatement of the reference forward (its grad_scale /
round_pass composition) differentiated by autograd."""
import math

import numpy as np
import pytest
import torch

from tests._util import rel_err

pytestmark = pytest.mark.gpu
TOL = 1e-5


def _grad_scale(x, s):
    return x.detach() - (x * s).detach() + x * s


def _round_pass(x):
    return x.round().detach() - x.detach() + x


def test_act_lsq_matches_reference_formula():
    import cim_quantization_b200 as cq
    torch.manual_seed(0)
    x = torch.randn(8, 16, 12, 12, device="cuda").relu_().requires_grad_(True)
    m = cq.ActLSQ(nbits_a=4).cuda().train()
    codes, s = m(x)  # lazy init on the first batch (lsq.py:628-641)
    assert m.init_state.item() == 1 and m.signed.item() == 0
    g = 1.0 / math.sqrt(x.numel() * 15)
    go = torch.randn_like(codes)
    (codes * go).sum().backward()
    gx, ga = x.grad.clone(), m.alpha.grad.clone()
    # reference composition (lsq.py:650-662)
    xr = x.detach().clone().requires_grad_(True)
    ar = m.alpha.detach().clone().requires_grad_(True)
    sr = _grad_scale(ar, g)
    cr = _round_pass((xr / sr).clamp(0, 15))
    (cr * go).sum().backward()
    assert torch.equal(codes.detach(), cr.detach()) and torch.equal(s.detach(), sr.detach())
    assert rel_err(gx.cpu().numpy(), xr.grad.cpu().numpy()) < TOL
    assert abs(ga.item() - ar.grad.item()) <= 1e-4 * abs(ar.grad.item()) + 1e-7


def test_linear_lsq_matches_reference_formula():
    import cim_quantization_b200 as cq
    torch.manual_seed(1)
    torch.backends.cuda.matmul.allow_tf32 = False
    m = cq.LinearLSQ(64, 10, bias=True, nbits_w=4).cuda().train()
    x = torch.randn(32, 64, device="cuda")
    y = m(x)
    go = torch.randn_like(y)
    (y * go).sum().backward()
    wr = m.weight.detach().clone().requires_grad_(True)
    ar = m.alpha.detach().clone().requires_grad_(True)
    g = 1.0 / math.sqrt(wr.numel() * 7)
    sr = _grad_scale(ar, g)
    wq = _round_pass((wr / sr).clamp(-8, 7)) * sr
    yr = torch.nn.functional.linear(x, wq, m.bias.detach())
    (yr * go).sum().backward()
    assert rel_err(y.detach().cpu().numpy(), yr.detach().cpu().numpy()) < TOL
    assert rel_err(m.weight.grad.cpu().numpy(), wr.grad.cpu().numpy()) < TOL
    assert abs(m.alpha.grad.item() - ar.grad.item()) <= 1e-4 * abs(ar.grad.item()) + 1e-7


def test_conv2d_lsq_matches_reference_formula():
    import cim_quantization_b200 as cq
    torch.manual_seed(2)
    torch.backends.cudnn.allow_tf32 = False
    act = cq.ActLSQ(nbits_a=4).cuda().train()
    conv = cq.Conv2dLSQ(8, 16, 3, padding=1, bias=False, nbits_w=4).cuda().train()
    x = torch.randn(4, 8, 10, 10, device="cuda").relu_()
    y = conv(act(x))
    go = torch.randn_like(y)
    (y * go).sum().backward()
    # reference composition (lsq.py:398-436 on top of lsq.py:624-662)
    aa = act.alpha.detach().clone().requires_grad_(True)
    aw = conv.alpha.detach().clone().requires_grad_(True)
    wr = conv.weight.detach().clone().requires_grad_(True)
    sa = _grad_scale(aa, 1.0 / math.sqrt(x.numel() * 15))
    xq = _round_pass((x / sa).clamp(0, 15))
    sw = _grad_scale(aw, 1.0 / math.sqrt(wr.numel() * 7))
    wq = _round_pass((wr / sw).clamp(-8, 7))
    yr = torch.nn.functional.conv2d(xq, wq, None, 1, 1) * sa * sw
    (yr * go).sum().backward()
    assert rel_err(y.detach().cpu().numpy(), yr.detach().cpu().numpy()) < TOL
    assert rel_err(conv.weight.grad.cpu().numpy(), wr.grad.cpu().numpy()) < TOL
    for got, ref in ((act.alpha.grad, aa.grad), (conv.alpha.grad, aw.grad)):
        assert abs(got.item() - ref.item()) <= 1e-4 * abs(ref.item()) + 1e-6


def test_linear_cim_equals_1x1_conv_and_oracle():
    """LinearLSQCiM is Conv2dLSQCiM with a 1x1 kernel over one pixel: identical to it bit for bit,
    and to the oracle's restatement of lsq.py:522-588 at 1e-5.  in=96 over xbar 64 -> a 32-row remainder crossbar."""
    import numpy as np
    import cim_quantization_b200 as cq
    from oracle import cim_oracle as O
    from tests._util import rel_err
    torch.manual_seed(3)
    B, fin, fout = 64, 96, 32
    kw = dict(nbits_w=3, nbits_a=3, nbits_alpha=8, wbitslice=1, abitslice=1, xbar=64, adcbits=1.5)
    lin = cq.LinearLSQCiM(fin, fout, bias=True, signed_xbar=False, **kw).cuda().train()
    conv = cq.Conv2dLSQCiM(fin, fout, 1, 1, 0, 1, 1, False, **kw).cuda().train()  # (the reference's conv bias add
    # broadcasts over the last axis, lsq.py:582-583: not comparable on a 1x1 map)
    with torch.no_grad():
        conv.weight.copy_(lin.weight.view(fout, fin, 1, 1))
    assert isinstance(lin, torch.nn.Linear)
    assert {k: tuple(v.shape) for k, v in lin.state_dict().items()} == {
        "weight": (fout, fin), "bias": (fout,), "alpha_cim": (1, 2, 3, 3, 1, fout), "alpha_weight": (1,),
        "alpha_act": (1,), "init_state": (1,), "signed_act": (1,), "init_state_cim": (1,)}
    x = torch.relu(torch.randn(B, fin, device="cuda"))
    gy = torch.randn(B, fout, device="cuda")
    xl = x.clone().requires_grad_(True)
    xc = x.clone().requires_grad_(True)
    yl = lin(xl)
    yl.backward(gy)
    yc = conv(xc[:, :, None, None])
    yc.backward(gy[:, :, None, None])
    assert torch.equal(yl, yc.flatten(1) + lin.bias)
    assert torch.equal(xl.grad, xc.grad)
    assert torch.equal(lin.weight.grad, conv.weight.grad.view(fout, fin))
    assert torch.allclose(lin.bias.grad, gy.sum(0), rtol=1e-6, atol=1e-6)
    for n in ("alpha_cim", "alpha_act", "alpha_weight"):
        assert torch.equal(getattr(lin, n).grad, getattr(conv, n).grad), n
    # second step (initialised) against the oracle
    cfg = O.CimConfig(in_channels=fin, out_channels=fout, kernel=1, stride=1, padding=0, **{k: v for k, v in kw.items()
                                                                                        if k != "nbits_alpha"})
    for p_ in lin.parameters():
        p_.grad = None
    xl = x.clone().requires_grad_(True)
    y = lin(xl)
    y.backward(gy)
    r = O.module_forward_backward(cfg, x.cpu().numpy()[:, :, None, None], lin.weight.detach().cpu().numpy()[:, :, None, None],
                                  lin.alpha_act.detach().cpu().numpy(), lin.alpha_weight.detach().cpu().numpy(),
                                  lin.alpha_cim.detach().cpu().numpy(), gy.cpu().numpy()[:, :, None, None])
    assert rel_err((y - lin.bias).detach().cpu().numpy(), r["y"].reshape(B, fout)) < 1e-5
    assert rel_err(xl.grad.cpu().numpy(), r["grad_x"].reshape(B, fin)) < 1e-5
    assert rel_err(lin.weight.grad.cpu().numpy(), r["grad_weight"].reshape(fout, fin)) < 1e-5
    assert rel_err(lin.alpha_cim.grad.cpu().numpy(), r["grad_alpha_cim"]) < 1e-5
