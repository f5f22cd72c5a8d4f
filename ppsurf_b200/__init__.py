"""ppsurf_b200: the PPSurf occupancy hot path (FKAConv encoder -> exact kNN -> per-query decode) on hand-written
sm_100a CUDA kernels behind a C ABI (``include/ppsurf_b200.h``), wrapped in the reference's module interface.

Importing the package loads ``libppsurf_b200.so`` (built in-tree by ``python ppsurf_b200/build.py``); there is no CPU
or PyTorch fallback."""
from . import _lib  # noqa: F401  (raises ImportError when the shared library is missing)
from . import ops, packing  # noqa: F401
from .network import PPSurfNetwork  # noqa: F401
from .model import PPSurfModel  # noqa: F401
from . import data_pipeline, mesh, sharding  # noqa: F401
