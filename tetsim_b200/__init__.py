"""tetsim_b200 -- B200-native XPBD tetrahedral-FEM substep solver behind the SoftBody / SoftBodyGPU
API of zalo/TetSim.  The product is libtetsim_b200.so (hand-written sm_100a CUDA behind a C ABI,
include/tetsim_b200.h); this package is the host-side mirror of the reference's two classes plus
mesh inputs.  No CPU fallback: using a body without the built library and a B200 raises.
"""
from ._capi import (ARITH_BITEXACT, ARITH_FAST_F32, NH_GS_COLOR, NH_GS_EXACT, NH_JACOBI, POLAR_JACOBI, TetSimError,
                    greedy_colors, level_schedule)
from .softbody import DEFAULT_PHYSICS_PARAMS, SoftBody, SoftBodyGPU, default_physics_params
from . import mesh

__all__ = ["SoftBody", "SoftBodyGPU", "DEFAULT_PHYSICS_PARAMS", "TetSimError", "mesh", "level_schedule",
           "greedy_colors", "NH_GS_EXACT", "NH_GS_COLOR", "NH_JACOBI", "POLAR_JACOBI", "ARITH_FAST_F32",
           "ARITH_BITEXACT"]
