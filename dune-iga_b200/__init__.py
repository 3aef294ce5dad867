"""dune-iga_b200: B200-native batched NURBS basis/geometry evaluation and element-matrix assembly behind the
dune-iga local-finite-element / geometry API.  Host side only here; the compute lives in csrc/ (CUDA, sm_100a)
behind the C-ABI declared in include/igab200.h."""
from . import partition, patchdata, quadrature  # noqa: F401
