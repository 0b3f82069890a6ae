"""Model registry: every lower-case callable exported here is a ``--model`` choice of main.py, exactly like
the reference's ``models/__init__.py`` + ``main.py:24-26``.  Scope of this package is the ResNet family of
the north-star hot path: resnet, resnext, mobilenet_v2 -- plus the neighbour families of section
8(f) row 4 that run on the same kernels: resnet_se / resnext_se (squeeze-excitation blocks), mobilenet (v1)."""
from .resnet import *  # noqa: F401,F403
from .resnext import *  # noqa: F401,F403
from .mobilenet_v2 import *  # noqa: F401,F403
from .mobilenet import *  # noqa: F401,F403
