"""Mirror of the reference's ``models/_modules`` package (``models/_modules/__init__.py:1-2``)."""
from ._quan_base import *  # noqa: F401,F403
from .lsq import *  # noqa: F401,F403
from ._quan_base import _Conv2dQ, _LinearQ, _ActQ, _Conv2dQCiM, _LinearQCiM  # noqa: F401  (underscore names, like the reference)
