"""mpm_b200 — B200-native MLS-MPM substep (drop-in for the hot path of kekeblom/mpm).

The product is the CUDA library `libmpm_b200.so` behind the C ABI in include/mpm_b200.h; this
package is the Python host-side mirror used by tests and bench.py.  There is no CPU fallback:
loading fails loudly if the library is missing and creating a simulation fails without a GPU.
"""
from .api import (  # noqa: F401
    FIXED_COROTATED, SNOW, JELLY, MODEL_USER, SVD_EXACT, SVD_FAST, P2G_RUNS, P2G_DIRECT, G2P_TILE, G2P_DIRECT, PIPE_HANDOVER, PIPE_CLASSIC, GRAPH_AUTO, GRAPH_OFF, GRAPH_ON, STAGES, PARTICLE_DTYPE, MpmError, Sim, lib, make_material,
    svd3_batch, polar_batch, determinant_batch, dinv_batch, comm_unique_id,
)
