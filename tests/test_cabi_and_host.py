"""CPU-side checks: the C-ABI library loads and exports every symbol include/cimq.h declares (no compute
calls without a GPU), derived layer sizes, and the host mirror of the reference's module surface."""
import ctypes
import os
import re

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def lib():
    import __graft_entry__ as ge
    ge.build()
    from cim_quantization_b200 import _lib
    return _lib


def test_header_symbols_exported(lib):
    header = open(os.path.join(ROOT, "include", "cimq.h")).read()
    declared = set(re.findall(r"\b(cimq_[a-z0-9_]+)\s*\(", header))
    assert declared, "no declarations found"
    dll = ctypes.CDLL(lib.LIB_PATH)
    for name in sorted(declared):
        assert hasattr(dll, name), f"libcimq.so does not export {name}"
    assert declared == set(lib.EXPORTS), "ctypes binding table and header disagree"
    assert lib.load().cimq_version() == 200


def test_layer_info_microbench(lib):
    spec = lib.LayerSpec(batch=256, in_channels=64, in_hw=32, out_channels=64, kernel=3, stride=1, padding=1,
                         nbits_a=3, abitslice=1, nbits_w=3, wbitslice=1, xbar=128, adcbits=1.5)
    i = lib.layer_info(spec)
    assert (i.out_hw, i.L, i.M, i.F, i.NX, i.NSW, i.NSA, i.pairs) == (32, 1024, 262144, 576, 5, 3, 3, 9)
    assert i.state_words == 1 and i.tc_forward == 1
    assert i.psum_count == 256 * 5 * 9 * 1024 * 64
    for xbar, nx in ((64, 9), (256, 3)):
        spec2 = lib.LayerSpec(256, 64, 32, 64, 3, 1, 1, 3, 1, 3, 1, xbar, 1.5)
        assert lib.layer_info(spec2).NX == nx


def test_bad_layer_reports_error(lib):
    layer = lib.CimqLayer(0, 64, 32, 64, 3, 1, 1, 3, 1, 3, 1, 128, 2, -1, 1)
    info = lib.CimqInfo()
    assert lib.load().cimq_layer_info(ctypes.byref(layer), ctypes.byref(info)) != 0
    assert b"bad layer shape" in lib.load().cimq_last_error()


def test_adc_mode_mapping(lib):
    assert lib.adc_mode_of(1) == (lib.ADC_BINARY, -1, 1)
    assert lib.adc_mode_of(1.5) == (lib.ADC_TERNARY, -1, 1)
    assert lib.adc_mode_of(3.0) == (lib.ADC_MULTIBIT, -4, 3)
    assert lib.adc_mode_of(4) == (lib.ADC_MULTIBIT, -8, 7)


def test_module_surface_matches_reference_contract():
    """State-dict keys, parameter shapes and attributes of _quan_base.py:174-237 / lsq.py:512-519."""
    import cim_quantization_b200 as cq
    m = cq.Conv2dLSQCiM(64, 64, (3, 3), (1, 1), (1, 1), (1, 1), 1, False, nbits_w=3, nbits_a=3, nbits_alpha=8,
                        wbitslice=1, abitslice=1, xbar=128, adcbits=1.5, signed_xbar=False, stochastic_quant=False)
    assert isinstance(m, torch.nn.Conv2d)
    assert list(m.state_dict().keys()) == ['weight', 'alpha_cim', 'alpha_weight', 'alpha_act', 'init_state',
                                           'signed_act', 'init_state_cim']
    assert tuple(m.alpha_cim.shape) == (1, 5, 3, 3, 1, 64)
    assert (m.num_xbars, m.num_bit_slice_weight, m.num_bit_slice_act) == (5, 3, 3)
    assert m.binary_mask.dtype == torch.int8 and tuple(m.binary_mask.shape) == (1, 1, 3, 3, 1, 1)
    assert m.binary_mask.flatten().tolist() == [1, 2, 4, 2, 4, 8, 4, 8, 16]
    assert all('alpha' in n for n, _ in m.named_parameters() if n != 'weight')  # optimizer no-decay filter
    m4 = cq.Conv2dLSQCiM(3, 16, (3, 3), 1, 1, 1, 1, True, nbits_w=8, nbits_a=8, xbar=128, adcbits=4)
    assert m4.alpha_cim is None and 'alpha_cim' not in m4.state_dict() and 'bias' in m4.state_dict()
    # int8 wrap of the 8x8 shift-and-add mask (_quan_base.py:214)
    bm = m4.binary_mask.view(8, 8)
    assert bm[7, 0].item() == -128 and bm[4, 4].item() == 0 and bm[3, 3].item() == 64
    for cls in (cq.Conv2dLSQ, cq.LinearLSQ, cq.ActLSQ):
        assert cls.__name__ in ('Conv2dLSQ', 'LinearLSQ', 'ActLSQ')
    assert list(cq.ActLSQ(nbits_a=4).state_dict().keys()) == ['alpha', 'init_state', 'signed']
    assert list(cq.LinearLSQ(8, 4, nbits_w=4).state_dict().keys()) == ['weight', 'bias', 'alpha', 'init_state']


def test_no_cpu_fallback():
    import cim_quantization_b200 as cq
    m = cq.Conv2dLSQCiM(4, 16, (3, 3), 1, 1, 1, 1, False, nbits_w=3, nbits_a=3, xbar=64, adcbits=1.5)
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        m(torch.zeros(1, 4, 8, 8))


def test_dropin_install_resolves_reference_import():
    from cim_quantization_b200 import dropin
    import sys
    saved = {k: sys.modules.get(k) for k in ('models', 'models._modules', 'models._modules.lsq',
                                              'models._modules._quan_base')}
    try:
        mods = dropin.install()
        import models._modules as my_nn  # the reference's import (examples/__init__.py:12)
        assert my_nn is mods and my_nn.Conv2dLSQCiM.__module__.startswith('cim_quantization_b200')
        from models._modules import _Conv2dQ, Qmodes, _LinearQ, _ActQ, _Conv2dQCiM  # lsq.py:16  # noqa: F401
    finally:
        for k, v in saved.items():
            if v is None:
                sys.modules.pop(k, None)
            else:
                sys.modules[k] = v


@pytest.mark.skipif(not os.path.isdir("/root/reference/utils/wrapper"), reason="reference tree not available")
def test_reference_model_surgery_accepts_dropin_class():
    """The reference's own ReplaceModuleTool (utils/wrapper/replace_module.py:70-115) builds our Conv2dLSQCiM
    from its model zoo exactly as main_lsq.py:53-56 does (construction only: no GPU here)."""
    import importlib.util
    import cim_quantization_b200 as cq
    spec = importlib.util.spec_from_file_location("ref_replace_module", "/root/reference/utils/wrapper/replace_module.py")
    rm = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(rm)
    spec2 = importlib.util.spec_from_file_location("ref_resnet", "/root/reference/models/cifar10/resnet.py")
    zoo = importlib.util.module_from_spec(spec2)
    spec2.loader.exec_module(zoo)
    model = zoo.resnet20(pretrained=False)
    rm.ReplaceModuleTool(model, {'Conv2d': [cq.Conv2dLSQCiM]}, True, nbits_w=3, nbits_a=3, nbits_alpha=8, wbitslice=1,
                         abitslice=1, xbar=128, adcbits=1.5, signed_xbar=False, stochastic_quant=False).replace()
    convs = [m for m in model.modules() if isinstance(m, cq.Conv2dLSQCiM)]
    assert len(convs) == 19 and (convs[0].nbits_w, convs[0].nbits_a) == (8, 8) and convs[1].nbits_w == 3
    assert sum(p.numel() for p in model.parameters()) == 293536
    assert sum(p.numel() for n, p in model.named_parameters() if 'alpha' in n) == 23814


@pytest.mark.skipif(not os.path.isdir("/root/reference/utils/wrapper"), reason="reference tree not available")
def test_reference_model_surgery_accepts_cim_linear():
    """replace_map={'Conv2d': [...], 'Linear': [LinearLSQCiM]}: the reference's ReplaceModuleTool also swaps the
    classifier (utils/wrapper/replace_module.py:34-64 needs an nn.Linear subclass taking the same keywords)."""
    import importlib.util
    import torch
    import cim_quantization_b200 as cq
    spec = importlib.util.spec_from_file_location("ref_replace_module2", "/root/reference/utils/wrapper/replace_module.py")
    rm = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(rm)
    spec2 = importlib.util.spec_from_file_location("ref_resnet2", "/root/reference/models/cifar10/resnet.py")
    zoo = importlib.util.module_from_spec(spec2)
    spec2.loader.exec_module(zoo)
    model = zoo.resnet20(pretrained=False)
    fc_w = model.linear.weight.detach().clone()
    rm.ReplaceModuleTool(model, {'Conv2d': [cq.Conv2dLSQCiM], 'Linear': [cq.LinearLSQCiM]}, True, nbits_w=3, nbits_a=3,
                         nbits_alpha=8, wbitslice=1, abitslice=1, xbar=128, adcbits=1.5, signed_xbar=False,
                         stochastic_quant=False).replace()
    assert isinstance(model.linear, cq.LinearLSQCiM) and isinstance(model.linear, torch.nn.Linear)
    assert torch.equal(model.linear.weight.detach(), fc_w)
    assert tuple(model.linear.alpha_cim.shape) == (1, 1, 3, 3, 1, 10) and model.linear.num_xbars == 1
    assert set(dict(model.linear.named_buffers())) == {"init_state", "signed_act", "init_state_cim"}
    with pytest.raises(RuntimeError):
        model.linear(torch.zeros(2, 64))  # no CPU fallback


def test_header_is_plain_c(tmp_path):
    """include/cimq.h is the C ABI: it must compile as C99 (no C++ or torch types in the signatures)."""
    import shutil
    import subprocess
    if shutil.which("gcc") is None:
        pytest.skip("gcc not available")
    src = tmp_path / "t.c"
    src.write_text('#include "cimq.h"\nint main(void){ cimq_layer_t l; cimq_info_t i; (void)l; (void)i; return 0; }\n')
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    r = subprocess.run(["gcc", "-std=c99", "-Wall", "-Wextra", "-pedantic", "-Werror", "-I", os.path.join(root, "include"),
                        "-fsyntax-only", str(src)], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
