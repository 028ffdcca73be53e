"""ctypes binding of the CPU oracle (oracle/glome_oracle.cpp).  TEST INFRASTRUCTURE ONLY.

Nothing under glome_b200/ imports this module; it is used by tests/, __graft_entry__.smoke() and
bench.py's cpu_baseline / --impl reference legs as the checker.
"""
import ctypes as C
import os
import subprocess

import numpy as np

from glome_b200 import _lib as L
from glome_b200.scene import HIT_DTYPE

_ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ORACLE_DIR = os.path.join(_ROOT, "oracle")
ORACLE_LIB = os.path.join(ORACLE_DIR, "_build", "libglome_oracle.so")

_lib = None
_vp = C.c_void_p


def build():
    subprocess.check_call(["make", "-C", ORACLE_DIR], stdout=subprocess.DEVNULL)


def load():
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(ORACLE_LIB):
        build()
    lib = C.CDLL(ORACLE_LIB)
    lib.orc_scene_create.restype = _vp
    lib.orc_scene_create.argtypes = [C.POINTER(L.GlomeFlatScene)]
    lib.orc_scene_destroy.argtypes = [_vp]
    lib.orc_rayint_batch.argtypes = [_vp, C.c_int64, _vp, _vp, C.c_int, _vp, C.c_int]
    lib.orc_shadow_batch.argtypes = [_vp, C.c_int64, _vp, _vp, C.c_int, _vp, C.c_int]
    lib.orc_debug_count_batch.argtypes = [_vp, C.c_int64, _vp, _vp, C.c_int, _vp, C.c_int]
    lib.orc_inside_batch.argtypes = [_vp, C.c_int64, _vp, _vp, C.c_int]
    lib.orc_trace_batch.argtypes = [_vp, C.c_int64, _vp, _vp, C.c_int, C.c_int, _vp, _vp, _vp, _vp, C.c_int]
    lib.orc_tile_count.argtypes = [C.c_int, C.c_int, C.c_int]
    lib.orc_tile_rects.argtypes = [C.c_int, C.c_int, C.c_int, _vp]
    lib.orc_render.argtypes = [_vp, C.POINTER(L.GlomeCamera), C.c_int, C.c_int, C.POINTER(L.GlomeRenderOpts), _vp, _vp,
                               C.c_int, C.c_int]
    lib.orc_stats.argtypes = [_vp, _vp, C.c_int]
    lib.orc_bih_build.argtypes = [C.c_int64, _vp, C.POINTER(C.POINTER(L.GlomeBihNode)), C.POINTER(C.c_int32),
                                  C.POINTER(C.POINTER(C.c_int32)), C.POINTER(C.c_int32),
                                  C.POINTER(C.POINTER(C.c_int32)), C.POINTER(C.c_int32), C.POINTER(C.c_double)]
    lib.orc_mesh_build.argtypes = [C.c_int64, _vp, C.c_int64, _vp, C.POINTER(C.POINTER(L.GlomeBvhNode)),
                                   C.POINTER(C.c_int32), C.POINTER(C.POINTER(C.c_int32)), C.POINTER(C.c_int32),
                                   C.POINTER(C.POINTER(C.c_int32)), C.POINTER(C.c_int32), C.POINTER(C.c_int32),
                                   C.POINTER(C.c_double)]
    lib.orc_free.argtypes = [_vp]
    lib.orc_bbclip_ub.argtypes = [_vp, _vp, _vp]
    lib.orc_getcoords.argtypes = [C.c_int, C.c_int, C.c_double, C.c_double, _vp]
    lib.orc_perlin.restype = C.c_double
    lib.orc_perlin.argtypes = [_vp]
    lib.orc_triangle_wave.restype = C.c_double
    lib.orc_triangle_wave.argtypes = [C.c_double]
    lib.orc_rgbf.restype = C.c_uint32
    lib.orc_rgbf.argtypes = [C.c_double, C.c_double, C.c_double]
    lib.orc_ccmp.restype = C.c_double
    lib.orc_ccmp.argtypes = [_vp, _vp]
    _lib = lib
    return lib


def _ptr(a):
    return a.ctypes.data_as(_vp)


def _f64(a, shape=None):
    a = np.ascontiguousarray(a, dtype=np.float64)
    return a.reshape(shape) if shape is not None else a


STAT_NAMES = ["bih_branch", "bvh_branch", "bih_leaf_items", "tri", "trinorm", "instance", "rays_primary",
              "rays_shadow", "rays_secondary", "overflow", "perlin_range"] + ["node_%d" % i for i in range(21)]

NCPU = os.cpu_count() or 1


class OracleScene:
    def __init__(self, flat):
        self.lib = load()
        self.h = self.lib.orc_scene_create(C.byref(flat))

    def close(self):
        if self.h:
            self.lib.orc_scene_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    @staticmethod
    def _tmax(tmax, n):
        t = _f64(np.atleast_1d(tmax))
        return (t, 0) if t.size == 1 else (t, 1)

    def rayint(self, rays, tmax=1000000.0, threads=NCPU):
        rays = _f64(rays, (-1, 6))
        t, stride = self._tmax(tmax, len(rays))
        out = np.zeros(len(rays), dtype=HIT_DTYPE)
        self.lib.orc_rayint_batch(self.h, len(rays), _ptr(rays), _ptr(t), stride, _ptr(out), threads)
        return out

    def shadow(self, rays, tmax=1000000.0, threads=NCPU):
        rays = _f64(rays, (-1, 6))
        t, stride = self._tmax(tmax, len(rays))
        out = np.zeros(len(rays), dtype=np.uint8)
        self.lib.orc_shadow_batch(self.h, len(rays), _ptr(rays), _ptr(t), stride, _ptr(out), threads)
        return out

    def inside(self, pts, threads=NCPU):
        pts = _f64(pts, (-1, 3))
        out = np.zeros(len(pts), dtype=np.uint8)
        self.lib.orc_inside_batch(self.h, len(pts), _ptr(pts), _ptr(out), threads)
        return out

    def trace(self, rays, tmax=1000000.0, recurs=3, want_hits=False, want_tags=False, threads=NCPU):
        rays = _f64(rays, (-1, 6))
        t, stride = self._tmax(tmax, len(rays))
        rgba = np.zeros((len(rays), 4))
        depth = np.zeros(len(rays))
        hits = np.zeros(len(rays), dtype=HIT_DTYPE) if want_hits else None
        tags = np.zeros((len(rays), 17), dtype=np.int32) if want_tags else None
        self.lib.orc_trace_batch(self.h, len(rays), _ptr(rays), _ptr(t), stride, int(recurs), _ptr(rgba), _ptr(depth),
                                 _ptr(hits) if want_hits else None, _ptr(tags) if want_tags else None, threads)
        res = [rgba, depth]
        if want_hits:
            res.append(hits)
        if want_tags:
            res.append(tags)
        return tuple(res)

    def debug_count(self, rays, tmax=1000000.0, threads=NCPU):
        rays = _f64(rays, (-1, 6))
        t, stride = self._tmax(tmax, len(rays))
        out = np.zeros(len(rays), dtype=np.int32)
        self.lib.orc_debug_count_batch(self.h, len(rays), _ptr(rays), _ptr(t), stride, _ptr(out), threads)
        return out

    def render(self, cam, width, height, opts, want_rgb8=False, threads=NCPU, max_tiles=0, out=None):
        tc = out if out is not None else np.zeros((height, width, 5))
        rgb = np.zeros((height, width), dtype=np.uint32) if want_rgb8 else None
        self.lib.orc_render(self.h, C.byref(cam), width, height, C.byref(opts), _ptr(tc),
                            _ptr(rgb) if want_rgb8 else None, threads, max_tiles)
        return tc, rgb

    def stats(self, reset=True):
        out = np.zeros(64, dtype=np.int64)
        k = self.lib.orc_stats(self.h, _ptr(out), 1 if reset else 0)
        return dict(zip(STAT_NAMES, out[:k].tolist()))


def bih_build(bboxes):
    lib = load()
    bboxes = _f64(bboxes, (-1, 6))
    nodes = C.POINTER(L.GlomeBihNode)()
    leaves = C.POINTER(C.c_int32)()
    order = C.POINTER(C.c_int32)()
    nn, nl, root = C.c_int32(), C.c_int32(), C.c_int32()
    bb = (C.c_double * 6)()
    rc = lib.orc_bih_build(len(bboxes), _ptr(bboxes), C.byref(nodes), C.byref(nn), C.byref(leaves), C.byref(nl),
                           C.byref(order), C.byref(root), bb)
    if rc != 0:
        raise RuntimeError("bih: infinite bounding box")
    node_dt = np.dtype([("lsplit", "<f8"), ("rsplit", "<f8"), ("axis", "<i4"), ("left", "<i4"), ("right", "<i4"),
                        ("pad", "<i4")])
    res = dict(
        nodes=np.frombuffer(C.string_at(nodes, nn.value * 32), dtype=node_dt).copy(),
        leaves=np.frombuffer(C.string_at(leaves, nl.value * 8), dtype=np.int32).copy().reshape(-1, 2),
        order=np.frombuffer(C.string_at(order, len(bboxes) * 4), dtype=np.int32).copy(),
        root=root.value, bb=np.array(bb[:]))
    for p in (nodes, leaves, order):
        lib.orc_free(C.cast(p, _vp))
    return res


def mesh_build(verts, tris):
    lib = load()
    verts = _f64(verts, (-1, 3))
    tris = np.ascontiguousarray(tris, dtype=np.int32).reshape(-1, 8)
    nodes = C.POINTER(L.GlomeBvhNode)()
    leafpool = C.POINTER(C.c_int32)()
    leafoff = C.POINTER(C.c_int32)()
    nn, nlp, nl, root = C.c_int32(), C.c_int32(), C.c_int32(), C.c_int32()
    bb = (C.c_double * 6)()
    lib.orc_mesh_build(len(verts), _ptr(verts), len(tris), _ptr(tris), C.byref(nodes), C.byref(nn), C.byref(leafpool),
                       C.byref(nlp), C.byref(leafoff), C.byref(nl), C.byref(root), bb)
    node_dt = np.dtype([("lbb", "<f8", (6,)), ("rbb", "<f8", (6,)), ("left", "<i4"), ("right", "<i4"),
                        ("pad", "<i4", (6,))])
    res = dict(
        nodes=np.frombuffer(C.string_at(nodes, nn.value * 128), dtype=node_dt).copy(),
        leafpool=np.frombuffer(C.string_at(leafpool, nlp.value * 4), dtype=np.int32).copy(),
        leafoff=np.frombuffer(C.string_at(leafoff, nl.value * 4), dtype=np.int32).copy(),
        root=root.value, bb=np.array(bb[:]))
    for p in (nodes, leafpool, leafoff):
        lib.orc_free(C.cast(p, _vp))
    return res


def tile_rects(width, height, blocksize=65):
    lib = load()
    n = lib.orc_tile_count(width, height, blocksize)
    out = np.zeros((n, 4), dtype=np.int32)
    lib.orc_tile_rects(width, height, blocksize, _ptr(out))
    return out
