"""ctypes binding of tests/tools/gen_host.cpp: the scene-graph machine (glome_b200/csrc/glome_gen.cuh) compiled for the
CPU.  TEST INFRASTRUCTURE ONLY -- lets the not-gpu suite check the machine's control flow against the oracle."""
import ctypes as C
import os
import subprocess

import numpy as np

from glome_b200 import _lib as L
from glome_b200.scene import HIT_DTYPE

_ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SRC = os.path.join(_ROOT, "tests", "tools", "gen_host.cpp")
OUT = os.path.join(_ROOT, "tests", "tools", "_build", "libgenhost.so")
DEPS = [SRC] + [os.path.join(_ROOT, "glome_b200", "csrc", f) for f in
                ("glome_gen.cuh", "glome_device.cuh", "glome_math.h", "glome_tagmap.h")] + [os.path.join(_ROOT, "include", "glome_cuda.h")]
_lib = None
_vp = C.c_void_p
NCPU = os.cpu_count() or 1


def build():
    os.makedirs(os.path.dirname(OUT), exist_ok=True)
    if os.path.exists(OUT) and all(os.path.getmtime(OUT) >= os.path.getmtime(d) for d in DEPS):
        return
    subprocess.check_call(["g++", "-O2", "-std=c++17", "-fPIC", "-ffp-contract=off", "-fno-fast-math", "-Wall",
                           "-Wno-unused-function", "-Wno-unused-variable", "-Wno-unknown-pragmas", "-pthread", "-shared",
                           "-o", OUT, SRC])


def load():
    global _lib
    if _lib is None:
        build()
        lib = C.CDLL(OUT)
        lib.genh_create.restype = _vp
        lib.genh_create.argtypes = [C.POINTER(L.GlomeFlatScene)]
        lib.genh_destroy.argtypes = [_vp]
        lib.genh_rayint_batch.argtypes = [_vp, C.c_int64, _vp, _vp, C.c_int, _vp, C.c_int]
        lib.genh_shadow_batch.argtypes = [_vp, C.c_int64, _vp, _vp, C.c_int, _vp, C.c_int]
        lib.genh_inside_batch.argtypes = [_vp, C.c_int64, _vp, _vp, C.c_int]
        lib.genh_debug_count_batch.argtypes = [_vp, C.c_int64, _vp, _vp, C.c_int, _vp, C.c_int]
        lib.genh_trace_batch.argtypes = [_vp, C.c_int64, _vp, _vp, C.c_int, C.c_int, _vp, _vp, _vp, _vp, _vp, C.c_int]
        _lib = lib
    return _lib


def _p(a):
    return a.ctypes.data_as(_vp)


def _f64(a, shape=None):
    a = np.ascontiguousarray(a, dtype=np.float64)
    return a.reshape(shape) if shape is not None else a


class HostGenScene:
    def __init__(self, flat):
        self.lib = load()
        self.h = self.lib.genh_create(C.byref(flat))
        if not self.h:
            raise RuntimeError("scene over a machine limit")

    def close(self):
        if self.h:
            self.lib.genh_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    @staticmethod
    def _tmax(tmax):
        t = _f64(np.atleast_1d(tmax))
        return (t, 0) if t.size == 1 else (t, 1)

    def rayint(self, rays, tmax=1000000.0, threads=NCPU):
        rays = _f64(rays, (-1, 6))
        t, stride = self._tmax(tmax)
        out = np.zeros(len(rays), dtype=HIT_DTYPE)
        self.lib.genh_rayint_batch(self.h, len(rays), _p(rays), _p(t), stride, _p(out), threads)
        return out

    def shadow(self, rays, tmax=1000000.0, threads=NCPU):
        rays = _f64(rays, (-1, 6))
        t, stride = self._tmax(tmax)
        out = np.zeros(len(rays), dtype=np.uint8)
        self.lib.genh_shadow_batch(self.h, len(rays), _p(rays), _p(t), stride, _p(out), threads)
        return out

    def inside(self, pts, threads=NCPU):
        pts = _f64(pts, (-1, 3))
        out = np.zeros(len(pts), dtype=np.uint8)
        self.lib.genh_inside_batch(self.h, len(pts), _p(pts), _p(out), threads)
        return out

    def debug_count(self, rays, tmax=1000000.0, threads=NCPU):
        rays = _f64(rays, (-1, 6))
        t, stride = self._tmax(tmax)
        out = np.zeros(len(rays), dtype=np.int32)
        self.lib.genh_debug_count_batch(self.h, len(rays), _p(rays), _p(t), stride, _p(out), threads)
        return out

    def trace(self, rays, tmax=1000000.0, recurs=3, want_tags=False, threads=NCPU):
        rays = _f64(rays, (-1, 6))
        t, stride = self._tmax(tmax)
        rgba = np.zeros((len(rays), 4))
        depth = np.zeros(len(rays))
        hits = np.zeros(len(rays), dtype=HIT_DTYPE)
        tags = np.zeros((len(rays), 17), dtype=np.int32) if want_tags else None
        cnt = np.zeros(8, dtype=np.int64)
        self.lib.genh_trace_batch(self.h, len(rays), _p(rays), _p(t), stride, int(recurs), _p(rgba), _p(depth), _p(hits),
                                  _p(tags) if want_tags else None, _p(cnt), threads)
        return rgba, depth, hits, tags, cnt
