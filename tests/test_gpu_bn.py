"""Fused batch norm (+ residual) (+ ReLU) kernels (csrc/bn_fused.cu) against torch.nn.BatchNorm2d
evaluated in float64: outputs, all gradients, running statistics, training and inference."""
import pytest
import torch

from tests._util import rel_err

pytestmark = pytest.mark.gpu
TOL = 1e-5


def _ref(x, res, bn64, relu, gy):
    x64 = x.double().requires_grad_(True)
    r64 = res.double().requires_grad_(True) if res is not None else None
    y = bn64(x64)
    if r64 is not None:
        y = y + r64
    if relu:
        y = torch.relu(y)
    y.backward(gy.double())
    return y, x64.grad, (r64.grad if r64 is not None else None)


@pytest.mark.parametrize("shape", [(8, 16, 8, 8), (256, 16, 32, 32), (4, 3, 5, 5), (6, 64, 3, 3)])
@pytest.mark.parametrize("residual,relu", [(False, False), (False, True), (True, True)])
@pytest.mark.parametrize("training", [True, False])
def test_fused_bn_matches_float64_batchnorm(shape, residual, relu, training):
    from cim_quantization_b200 import functional as CF
    torch.manual_seed(hash((shape, residual, relu, training)) % 1000)
    b, c, h, w = shape
    x = (torch.randn(shape, device="cuda") * 1.7 + 0.4)
    res = torch.randn(shape, device="cuda") if residual else None
    gy = torch.randn(shape, device="cuda")
    bn = torch.nn.BatchNorm2d(c).cuda()
    with torch.no_grad():
        bn.weight.uniform_(0.5, 1.5)
        bn.bias.uniform_(-0.3, 0.3)
        bn.running_mean.uniform_(-0.2, 0.2)
        bn.running_var.uniform_(0.6, 1.4)
    bn64 = torch.nn.BatchNorm2d(c).cuda().double()
    bn64.load_state_dict({k: v.double() if v.is_floating_point() else v for k, v in bn.state_dict().items()})
    bn.train(training)
    bn64.train(training)
    xg = x.clone().requires_grad_(True)
    rg = res.clone().requires_grad_(True) if residual else None
    y = CF.batch_norm_act(xg, bn, rg, relu)
    y.backward(gy)
    y_ref, gx_ref, gr_ref = _ref(x, res, bn64, relu, gy)
    assert rel_err(y.detach().cpu().numpy(), y_ref.detach().cpu().numpy()) < TOL
    assert rel_err(xg.grad.cpu().numpy(), gx_ref.cpu().numpy()) < 2e-5  # fp32 cancellation in dy - mean(dy) - ...
    if residual:
        assert rel_err(rg.grad.cpu().numpy(), gr_ref.cpu().numpy()) < TOL
    assert rel_err(bn.weight.grad.cpu().numpy(), bn64.weight.grad.cpu().numpy()) < TOL
    assert rel_err(bn.bias.grad.cpu().numpy(), bn64.bias.grad.cpu().numpy()) < TOL
    assert rel_err(bn.running_mean.cpu().numpy(), bn64.running_mean.cpu().numpy()) < TOL
    assert rel_err(bn.running_var.cpu().numpy(), bn64.running_var.cpu().numpy()) < TOL
    assert int(bn.num_batches_tracked) == int(bn64.num_batches_tracked)


@pytest.mark.parametrize("shape,res,relu", [((8, 16, 32, 32), False, True), ((4, 32, 16, 16), True, True),
                                             ((3, 64, 7, 7), True, True), ((2, 16, 8, 8), False, False)])
def test_bn_forward_quant_codes_equal_the_separate_quantiser(shape, res, relu):
    """cimq_bn_forward_quant writes the consumer's activation codes in the batch-norm epilogue; they are
    the bytes cimq_lsq_quantize computes from the fp32 output with the consumer's step size (lsq.py:547-549)."""
    from cim_quantization_b200 import _lib as L
    torch.manual_seed(3)
    b, c, h, w = shape
    x = torch.randn(shape, device="cuda") * 2.0 + 0.3
    r = torch.randn(shape, device="cuda") if res else None
    wt, bs = torch.rand(c, device="cuda") + 0.5, torch.randn(c, device="cuda") * 0.1
    rm, rv = torch.zeros(c, device="cuda"), torch.ones(c, device="cuda")
    for qp, alpha in ((7, 0.31), (15, 0.173), (255, 0.0123)):
        alpha_act = torch.tensor([alpha], device="cuda")
        g = 1.0 / (x.numel() * qp) ** 0.5
        y0, _, _ = L.bn_forward(x, r, wt, bs, rm.clone(), rv.clone(), True, 0.1, 1e-5, relu)
        y1, _, _, codes = L.bn_forward(x, r, wt, bs, rm.clone(), rv.clone(), True, 0.1, 1e-5, relu,
                                       next_quant=(alpha_act, g, qp))
        assert torch.equal(y0, y1)
        s = L.step_sizes(alpha_act, alpha_act, g, g)
        ref = L.lsq_quantize(y0, s[0:1], 0, qp)
        assert codes.dtype == torch.uint8 and torch.equal(codes, ref.view_as(codes))


def test_resnet20_step_identical_with_and_without_fused_quantiser():
    """The harness ResNet-20 with the batch-norm kernels writing the next layer's codes (harness.FUSED_QUANT) produces
    the same logits and parameter gradients, bit for bit, as with the separate quantiser launches."""
    from cim_quantization_b200 import harness
    torch.manual_seed(0)
    model = harness.convert_to_cim(harness.resnet20(), nbits_w=3, nbits_a=3, xbar=128, adcbits=1.5).cuda().train()
    x = torch.randn(16, 3, 32, 32, device="cuda")
    y = torch.randint(0, 10, (16,), device="cuda")
    model(x)  # lazy initialisation of the step sizes
    bn_state = {k: v.clone() for k, v in model.state_dict().items()}
    results = []
    from cim_quantization_b200 import _lib as L
    L.set_deterministic(True)  # fixed-order fold instead of fp32 atomics in the dgrad epilogues
    for fused in (False, True):
        harness.FUSED_QUANT = fused
        model.load_state_dict(bn_state)
        for p in model.parameters():
            p.grad = None
        logits = model(x)
        torch.nn.functional.cross_entropy(logits, y).backward()
        results.append((logits.detach().clone(), [p.grad.clone() for p in model.parameters()]))
    harness.FUSED_QUANT = True
    L.set_deterministic(False)
    assert torch.equal(results[0][0], results[1][0])
    for g0, g1 in zip(results[0][1], results[1][1]):
        assert torch.equal(g0, g1)
