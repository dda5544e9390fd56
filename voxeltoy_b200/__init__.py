"""voxeltoy_b200 -- B200-native implementation of voxelToy's progressive path tracer and
mesh voxelizer (hand-written sm_100a CUDA kernels behind a C ABI, include/voxeltoy_b200.h).

There is no CPU, OpenGL or Triton fallback: importing works anywhere, but creating a
Context needs the built library (python -m voxeltoy_b200.build) and a CUDA device."""
from ._capi import (Context, VtError, load, LIB_PATH, SIGNATURES,  # noqa: F401
                    VT_LENS_PINHOLE, VT_LENS_THIN, VT_LENS_ORTHO,
                    VT_INTEGRATOR_PATHTRACER, VT_INTEGRATOR_EDIT_MODE,
                    VT_PART_NONE, VT_PART_TILES, VT_PART_SAMPLES)

from . import host, scenes  # noqa: E402,F401

__version__ = "0.1"
