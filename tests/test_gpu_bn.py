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
