"""glome_b200 -- B200 (sm_100a) backend for GlomeTrace's ray-cast hot path.

The product is libglomecuda.so (glome_b200/csrc, C-ABI in include/glome_cuda.h); this package is
the ctypes plumbing tests and bench.py use.  Importing it never touches oracle/.
"""
from . import _lib  # noqa: F401
from .scene import (Scene, SceneBuilder, MultiScene, FlatView, camera, camera_rays, compose, deg, render_opts, rotate, scale,  # noqa: F401
                    tile_rects, translate, HIT_DTYPE, bih_build, mesh_build)
