"""Make the reference's import ``import models._modules as my_nn`` (examples/__init__.py:12) resolve to
this package, so ``replace_map={'Conv2d': [my_nn.Conv2dLSQCiM]}`` (main_lsq.py:53-56) picks up the
sm_100a kernels with ``main_lsq.py`` and ``utils/wrapper/replace_module.py`` unchanged."""
import sys
import types


def install(force: bool = True):
    from . import modules
    from .modules import _quan_base, lsq
    pkg = sys.modules.get('models')
    if pkg is None:
        try:
            import models as pkg  # the reference tree (or a copy) is on sys.path: keep its model zoo
        except Exception:
            pkg = types.ModuleType('models')
            pkg.__path__ = []
            sys.modules['models'] = pkg
    if force or 'models._modules' not in sys.modules:
        sys.modules['models._modules'] = modules
        sys.modules['models._modules._quan_base'] = _quan_base
        sys.modules['models._modules.lsq'] = lsq
        setattr(pkg, '_modules', modules)
    return modules
