"""imomd_b200 -- host-side Python mirror of the B200-native IMOMD-RRT* hot path.

Python here is plumbing only (ctypes over the C ABI in include/imomd_b200.h, graph
generators, the multi-GPU launcher).  The product is the CUDA library under
``csrc/`` and the C++ facade under ``host/``.
"""
from . import graphs  # noqa: F401
