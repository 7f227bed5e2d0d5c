"""cim_quantization_b200 -- B200 (sm_100a) implementation of the CiM-aware quantized conv path of
UtkarshSaxena1/CiM_Quantization (``models/_modules/lsq.py``).

* ``cim_quantization_b200.modules``  mirrors the reference's ``models._modules`` (same classes / Function).
* ``cim_quantization_b200.dropin.install()`` makes ``import models._modules`` resolve to it, so the
  reference's ``examples/classifier_cifar10/main_lsq.py`` runs on these kernels unchanged.
* ``libcimq.so`` (C ABI in ``include/cimq.h``) holds the hand-written CUDA kernels; there is no CPU or
  PyTorch fallback -- importing the compute path without the built library fails.
"""
from . import _lib  # noqa: F401
from ._lib import set_deterministic  # noqa: F401
from .functional import cim_conv2d, get_cim_output_signed, lsq_fake_quant  # noqa: F401
from .modules import ActLSQ, Conv2dLSQ, Conv2dLSQCiM, LinearLSQ, LinearLSQCiM  # noqa: F401

__all__ = ['ActLSQ', 'Conv2dLSQ', 'Conv2dLSQCiM', 'LinearLSQ', 'LinearLSQCiM', 'cim_conv2d', 'get_cim_output_signed',
           'lsq_fake_quant', 'set_deterministic']
