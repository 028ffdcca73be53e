"""Python face of the host mirror + device scene (plumbing over the C-ABI, include/glome_cuda.h).

`SceneBuilder` exposes the GlomeTrace constructors under their reference names (sphere, box, cone,
difference, bih, mesh, tex, transform ...), implemented in C++ (glome_b200/csrc/host_builder.cpp).
`Scene` is a FlatScene uploaded to one B200; its methods are the batch forms of the reference's
`rayint` / `shadow` / `inside` / `trace` / `renderTiles`.
"""
import ctypes as C

import numpy as np

from . import _lib as L

HIT_DTYPE = np.dtype([("t", "<f8"), ("pos", "<f8", (3,)), ("norm", "<f8", (3,)), ("hit", "<i4"), ("prim", "<i4"),
                      ("sub", "<i4"), ("ntex", "<i4"), ("ntag", "<i4"), ("flags", "<i4"),
                      ("tex", "<i4", (L.GLOME_MAX_STACK,)), ("tag", "<i4", (L.GLOME_MAX_STACK,))])
assert HIT_DTYPE.itemsize == C.sizeof(L.GlomeHit)


def _d3(v):
    return (C.c_double * 3)(*[float(x) for x in v])


def _ptr(a):
    return a.ctypes.data_as(C.c_void_p)


def _f64(a, shape=None):
    a = np.ascontiguousarray(a, dtype=np.float64)
    if shape is not None:
        a = a.reshape(shape)
    return a


def _i32(a):
    return np.ascontiguousarray(a, dtype=np.int32)


def translate(v):
    out = (C.c_double * 24)()
    L.check(L.load().glome_xfm_translate(_d3(v), out))
    return np.array(out[:], dtype=np.float64)


def scale(v):
    out = (C.c_double * 24)()
    L.check(L.load().glome_xfm_scale(_d3(v), out))
    return np.array(out[:], dtype=np.float64)


def rotate(axis, angle):
    out = (C.c_double * 24)()
    L.check(L.load().glome_xfm_rotate(_d3(axis), float(angle), out))
    return np.array(out[:], dtype=np.float64)


def compose(xfms):
    xs = _f64(np.stack([_f64(x, (24,)) for x in xfms]) if len(xfms) else np.zeros((0, 24)))
    out = (C.c_double * 24)()
    L.check(L.load().glome_xfm_compose(len(xfms), _ptr(xs), out))
    return np.array(out[:], dtype=np.float64)


def deg(x):
    """Vec.hs:17 (pi truncated to 3.1415926535897)."""
    return (x * 3.1415926535897) / 180


def camera(pos, at, up, angle_deg):
    cam = L.GlomeCamera()
    L.check(L.load().glome_camera(_d3(pos), _d3(at), _d3(up), float(angle_deg), C.byref(cam)))
    return cam


def render_opts(**kw):
    o = L.GlomeRenderOpts()
    L.load().glome_render_opts_default(C.byref(o))
    for k, v in kw.items():
        if k == "thresholds":
            for i in range(4):
                o.thresholds[i] = float(v[i])
        else:
            setattr(o, k, v)
    return o


def tile_rects(width, height, blocksize=65):
    lib = L.load()
    n = lib.glome_tile_count(width, height, blocksize)
    out = np.zeros((n, 4), dtype=np.int32)
    for i in range(n):
        L.check(lib.glome_tile_rect(width, height, blocksize, i, out[i].ctypes.data_as(C.POINTER(C.c_int32))))
    return out


class FlatView:
    """numpy views of a GlomeFlatScene (valid while the owning builder is alive and un-reflattened)."""

    def __init__(self, fs):
        self.fs = fs

        def arr(ptr, n, dtype, shape=None):
            if n == 0:
                return np.zeros((0,) if shape is None else shape, dtype=dtype)
            a = np.ctypeslib.as_array(C.cast(ptr, C.POINTER(C.c_uint8)), shape=(n * np.dtype(dtype).itemsize,))
            return a.view(dtype)

        self.nodes = arr(fs.nodes, fs.n_nodes, np.dtype([("type", "<i4"), ("a", "<i4"), ("b", "<i4"), ("c", "<i4")]))
        self.bihnodes = arr(fs.bihnodes, fs.n_bihnodes,
                            np.dtype([("lsplit", "<f8"), ("rsplit", "<f8"), ("axis", "<i4"), ("left", "<i4"),
                                      ("right", "<i4"), ("pad", "<i4")]))
        self.bvhnodes = arr(fs.bvhnodes, fs.n_bvhnodes,
                            np.dtype([("lbb", "<f8", (6,)), ("rbb", "<f8", (6,)), ("left", "<i4"), ("right", "<i4"),
                                      ("pad", "<i4", (6,))]))
        self.ipool = arr(fs.ipool, fs.n_ipool, np.int32)
        self.dpool = arr(fs.dpool, fs.n_dpool, np.float64)
        self.root = fs.root
        self.scene_class = fs.scene_class
        self.max_depth = fs.max_depth


BIH_NODE_DTYPE = np.dtype([("lsplit", "<f8"), ("rsplit", "<f8"), ("axis", "<i4"), ("left", "<i4"), ("right", "<i4"),
                           ("pad", "<i4")])


def bih_build(bboxes, device=None):
    """The tree `bih` builds over objects with these bounding boxes (Bih.hs:211-324), as arrays:
    nodes (pre-order GlomeBihNode records), leaves ({first, count} into order), order (the leaf-ordered
    permutation), root ref, bb.  device=None: host builder; device=k: built on GPU k (same arrays), with
    timings_ms = (H2D, device build, D2H)."""
    lib = L.load()
    bboxes = _f64(bboxes, (-1, 6))
    nodes = C.POINTER(L.GlomeBihNode)()
    leaves = C.POINTER(C.c_int32)()
    order = C.POINTER(C.c_int32)()
    nn, nl, root = C.c_int32(), C.c_int32(), C.c_int32()
    bb = (C.c_double * 6)()
    tm = (C.c_double * 3)()
    if device is None:
        L.check(lib.glome_bih_build(len(bboxes), _ptr(bboxes), C.byref(nodes), C.byref(nn), C.byref(leaves),
                                    C.byref(nl), C.byref(order), C.byref(root), bb))
    else:
        L.check(lib.glome_bih_build_gpu(len(bboxes), _ptr(bboxes), int(device), C.byref(nodes), C.byref(nn),
                                        C.byref(leaves), C.byref(nl), C.byref(order), C.byref(root), bb, tm))
    res = dict(nodes=np.frombuffer(C.string_at(nodes, nn.value * 32), dtype=BIH_NODE_DTYPE).copy(),
               leaves=np.frombuffer(C.string_at(leaves, nl.value * 8), dtype=np.int32).copy().reshape(-1, 2),
               order=np.frombuffer(C.string_at(order, len(bboxes) * 4), dtype=np.int32).copy(), root=root.value,
               bb=np.array(bb[:]), timings_ms=tuple(tm[:]))
    for p in (nodes, leaves, order):
        lib.glome_free(C.cast(p, C.c_void_p))
    return res


BVH_NODE_DTYPE = np.dtype([("lbb", "<f8", (6,)), ("rbb", "<f8", (6,)), ("left", "<i4"), ("right", "<i4"), ("pad", "<i4", (6,))])


def mesh_build(verts, tris, device=None):
    """The BVH `mesh` builds (Mesh.hs:50-134): nodes (pre-order GlomeBvhNode), leafpool ({count, tri...} records),
    leafoff (leaf -> offset), root ref, bb.  tris = n x 8 int32 {a,b,c,na,nb,nc,tex,tag}.  device as in bih_build."""
    lib = L.load()
    verts = _f64(verts, (-1, 3))
    tris = _i32(tris).reshape(-1, 8)
    nodes = C.POINTER(L.GlomeBvhNode)()
    leafpool = C.POINTER(C.c_int32)()
    leafoff = C.POINTER(C.c_int32)()
    nn, nlp, nl, root = C.c_int32(), C.c_int32(), C.c_int32(), C.c_int32()
    bb = (C.c_double * 6)()
    tm = (C.c_double * 3)()
    if device is None:
        L.check(lib.glome_mesh_build(len(verts), _ptr(verts), len(tris), _ptr(tris), C.byref(nodes), C.byref(nn),
                                     C.byref(leafpool), C.byref(nlp), C.byref(leafoff), C.byref(nl), C.byref(root), bb))
    else:
        L.check(lib.glome_mesh_build_gpu(len(verts), _ptr(verts), len(tris), _ptr(tris), int(device), C.byref(nodes),
                                         C.byref(nn), C.byref(leafpool), C.byref(nlp), C.byref(leafoff), C.byref(nl),
                                         C.byref(root), bb, tm))
    res = dict(nodes=np.frombuffer(C.string_at(nodes, nn.value * 128), dtype=BVH_NODE_DTYPE).copy(),
               leafpool=np.frombuffer(C.string_at(leafpool, nlp.value * 4), dtype=np.int32).copy(),
               leafoff=np.frombuffer(C.string_at(leafoff, nl.value * 4), dtype=np.int32).copy(), root=root.value,
               bb=np.array(bb[:]), timings_ms=tuple(tm[:]))
    for p in (nodes, leafpool, leafoff):
        lib.glome_free(C.cast(p, C.c_void_p))
    return res


def bih_preorder_stream(tree):
    """The pre-order stream glome_sb_bih_prebuilt takes (what a Haskell `Flatten` instance emits while walking
    `BihBranch lsplit rsplit axis l r | BihLeaf [s]`, Bih.hs:51-57), from the arrays of bih_build: (kinds, splits,
    leaf-ordered item positions)."""
    nodes, leaves, order = tree["nodes"], tree["leaves"], tree["order"]
    kinds, splits, items = [], [], []
    stack = [tree["root"]]
    while stack:
        ref = stack.pop()
        if ref < 0:
            first, count = leaves[~ref]
            kinds.append(-(int(count) + 1))
            splits.append((0.0, 0.0))
            items.extend(int(x) for x in order[first:first + count])
        else:
            nd = nodes[ref]
            kinds.append(int(nd["axis"]))
            splits.append((float(nd["lsplit"]), float(nd["rsplit"])))
            stack.append(int(nd["right"]))
            stack.append(int(nd["left"]))
    return np.array(kinds, np.int32), np.array(splits, np.float64).reshape(-1, 2), np.array(items, np.int32)


def mesh_preorder_stream(tree):
    """Same for `Branch lbb rbb l r | Leaf [Tri]` (Mesh.hs:36-42), from the arrays of mesh_build: (kinds, boxes, leaf_tris)."""
    nodes, leafpool, leafoff = tree["nodes"], tree["leafpool"], tree["leafoff"]
    kinds, boxes, tris = [], [], []
    stack = [tree["root"]]
    while stack:
        ref = stack.pop()
        if ref < 0:
            off = leafoff[~ref]
            count = int(leafpool[off])
            kinds.append(-(count + 1))
            boxes.append(np.zeros(12))
            tris.extend(int(x) for x in leafpool[off + 1:off + 1 + count])
        else:
            nd = nodes[ref]
            kinds.append(0)
            boxes.append(np.concatenate([nd["lbb"], nd["rbb"]]))
            stack.append(int(nd["right"]))
            stack.append(int(nd["left"]))
    return np.array(kinds, np.int32), np.array(boxes, np.float64).reshape(-1, 12), np.array(tris, np.int32)


class SceneBuilder:
    """Host mirror of the GlomeTrace construction API.  Items are int ids."""

    def __init__(self):
        self.lib = L.load()
        h = C.c_void_p()
        L.check(self.lib.glome_builder_create(C.byref(h)))
        self.h = h

    def close(self):
        if self.h:
            self.lib.glome_builder_destroy(self.h)
            self.h = None

    def set_build_device(self, device):
        """device >= 0: `bih` builds its tree on that GPU (same tree); -1: on the host (default)."""
        L.check(self.lib.glome_builder_set_build_device(self.h, int(device)))

    def load_nff(self, text):
        """Spd.hs:1-261: NFF / SPD scene text -> (root item, camera, background rgb, bytes consumed)."""
        if isinstance(text, str):
            text = text.encode("ascii")
        cam = L.GlomeCamera()
        bg = (C.c_double * 3)()
        used = C.c_int64()
        root = L.check(self.lib.glome_sb_load_nff(self.h, text, len(text), C.byref(cam), bg, C.byref(used)))
        return root, cam, tuple(bg[:]), used.value

    def last_build_ms(self):
        """(H2D, device build, D2H, wall) of the last `bih`, milliseconds."""
        out = (C.c_double * 4)()
        L.check(self.lib.glome_builder_last_build_ms(self.h, out))
        return tuple(out[:])

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # -- primitives --
    def void(self):
        return L.check(self.lib.glome_sb_void(self.h))

    def sphere(self, c, r):
        return L.check(self.lib.glome_sb_sphere(self.h, _d3(c), float(r)))

    def spheres(self, centers, radii):
        centers = _f64(centers, (-1, 3))
        radii = _f64(radii, (-1,))
        ids = np.zeros(len(radii), dtype=np.int32)
        L.check(self.lib.glome_sb_spheres(self.h, len(radii), _ptr(centers), _ptr(radii), _ptr(ids)))
        return ids

    def triangle(self, p1, p2, p3):
        a = (C.c_double * 9)(*[float(x) for x in (*p1, *p2, *p3)])
        return L.check(self.lib.glome_sb_triangle(self.h, a))

    def trianglenorm(self, p1, p2, p3, n1, n2, n3):
        a = (C.c_double * 18)(*[float(x) for x in (*p1, *p2, *p3, *n1, *n2, *n3)])
        return L.check(self.lib.glome_sb_trianglenorm(self.h, a))

    def box(self, p1, p2):
        return L.check(self.lib.glome_sb_box(self.h, _d3(p1), _d3(p2)))

    def plane(self, orig, norm):
        return L.check(self.lib.glome_sb_plane(self.h, _d3(orig), _d3(norm)))

    def plane_offset(self, norm, off):
        return L.check(self.lib.glome_sb_plane_offset(self.h, _d3(norm), float(off)))

    def disc(self, pos, norm, r):
        return L.check(self.lib.glome_sb_disc(self.h, _d3(pos), _d3(norm), float(r)))

    def cylinder(self, p1, p2, r):
        return L.check(self.lib.glome_sb_cylinder(self.h, _d3(p1), _d3(p2), float(r)))

    def cone(self, p1, r1, p2, r2):
        return L.check(self.lib.glome_sb_cone(self.h, _d3(p1), float(r1), _d3(p2), float(r2)))

    def cylinder_z(self, r, h1, h2):
        return L.check(self.lib.glome_sb_cylinder_z(self.h, float(r), float(h1), float(h2)))

    def cone_z(self, r, h1, h2, height):
        return L.check(self.lib.glome_sb_cone_z(self.h, float(r), float(h1), float(h2), float(height)))

    # -- composites --
    def group(self, items):
        a = _i32(items)
        return L.check(self.lib.glome_sb_group(self.h, len(a), _ptr(a)))

    def bih(self, items):
        a = _i32(items)
        return L.check(self.lib.glome_sb_bih(self.h, len(a), _ptr(a)))

    def mesh(self, verts, norms, tris, texs=(), tags=()):
        verts = _f64(verts, (-1, 3))
        norms = _f64(norms, (-1, 3)) if len(norms) else np.zeros((0, 3))
        tris = _i32(tris).reshape(-1, 8)
        texs, tags = _i32(texs), _i32(tags)
        return L.check(self.lib.glome_sb_mesh(self.h, len(verts), _ptr(verts), len(norms), _ptr(norms), len(tris),
                                              _ptr(tris), len(texs), _ptr(texs), len(tags), _ptr(tags)))

    # raw forms (what the Haskell `flatten` instances call: stored fields, no constructor logic)
    def list_raw(self, items):
        a = _i32(items)
        return L.check(self.lib.glome_sb_list(self.h, len(a), _ptr(a)))

    def instance_raw(self, item, xfm):
        x = _f64(xfm, (24,))
        return L.check(self.lib.glome_sb_instance(self.h, item, x.ctypes.data_as(C.POINTER(C.c_double))))

    def disc_raw(self, pos, norm, rsqr):
        return L.check(self.lib.glome_sb_disc_raw(self.h, _d3(pos), _d3(norm), float(rsqr)))

    def difference_ex(self, sa, sb, useatex):
        return L.check(self.lib.glome_sb_difference_ex(self.h, sa, sb, 1 if useatex else 0))

    def bih_prebuilt(self, items, kinds, splits, bb):
        """A Bih whose tree the caller built (Bih.hs:51-57), as the pre-order stream include/glome_cuda.h describes."""
        a, kinds, splits, bb = _i32(items), _i32(kinds), _f64(splits, (-1, 2)), _f64(bb, (6,))
        return L.check(self.lib.glome_sb_bih_prebuilt(self.h, len(a), _ptr(a), len(kinds), _ptr(kinds), _ptr(splits), _ptr(bb)))

    def mesh_prebuilt(self, verts, norms, tris, texs, tags, kinds, boxes, leaf_tris, bb):
        """A Mesh whose BVH the caller built (Mesh.hs:36-42), same pre-order stream with two boxes per branch."""
        verts = _f64(verts, (-1, 3))
        norms = _f64(norms, (-1, 3)) if len(norms) else np.zeros((0, 3))
        tris = _i32(tris).reshape(-1, 8)
        texs, tags, kinds, leaf_tris = _i32(texs), _i32(tags), _i32(kinds), _i32(leaf_tris)
        boxes, bb = _f64(boxes, (-1, 12)), _f64(bb, (6,))
        return L.check(self.lib.glome_sb_mesh_prebuilt(self.h, len(verts), _ptr(verts), len(norms), _ptr(norms), len(tris),
                                                       _ptr(tris), len(texs), _ptr(texs), len(tags), _ptr(tags), len(kinds),
                                                       _ptr(kinds), _ptr(boxes), len(leaf_tris), _ptr(leaf_tris), _ptr(bb)))

    def difference(self, sa, sb):
        return L.check(self.lib.glome_sb_difference(self.h, sa, sb))

    def intersection(self, items):
        a = _i32(items)
        return L.check(self.lib.glome_sb_intersection(self.h, len(a), _ptr(a)))

    def tex(self, item, texture):
        return L.check(self.lib.glome_sb_tex(self.h, item, texture))

    def tag(self, item, tag):
        return L.check(self.lib.glome_sb_tag(self.h, item, tag))

    def noshadow(self, item):
        return L.check(self.lib.glome_sb_noshadow(self.h, item))

    def onlyshadow(self, item):
        return L.check(self.lib.glome_sb_onlyshadow(self.h, item))

    def bound_object(self, sa, sb):
        return L.check(self.lib.glome_sb_bound_object(self.h, sa, sb))

    def innerbound(self, sa, sb):
        return L.check(self.lib.glome_sb_innerbound(self.h, sa, sb))

    def transform(self, item, xfms):
        xs = _f64(np.stack([_f64(x, (24,)) for x in xfms]))
        return L.check(self.lib.glome_sb_transform(self.h, item, len(xfms), _ptr(xs)))

    def flatten_transform_bih(self, item):
        return L.check(self.lib.glome_sb_flatten_transform_bih(self.h, item))

    def bound(self, item):
        out = (C.c_double * 6)()
        L.check(self.lib.glome_sb_bound(self.h, item, out))
        return np.array(out[:])

    # -- materials, textures, lights --
    def mat_surface(self, rgb, alpha=1.0, amb=0.2, kd=1.0, ks=0.0, shine=0.0):
        return L.check(self.lib.glome_sb_mat_surface(self.h, _d3(rgb), alpha, amb, kd, ks, shine))

    def mat_reflect(self, refl):
        return L.check(self.lib.glome_sb_mat_reflect(self.h, float(refl)))

    def mat_refract(self, refl, refr, ior):
        return L.check(self.lib.glome_sb_mat_refract(self.h, float(refl), float(refr), float(ior)))

    def mat_warp(self, frame, scene, lightset, xfm):
        x = (C.c_double * 24)(*[float(v) for v in xfm])
        return L.check(self.lib.glome_sb_mat_warp(self.h, frame, scene, lightset, x))

    def mat_warp_set_scene(self, mat, scene):
        return L.check(self.lib.glome_sb_mat_warp_set_scene(self.h, mat, scene))

    def mat_additive(self, mats):
        a = _i32(mats)
        return L.check(self.lib.glome_sb_mat_additive(self.h, len(a), _ptr(a)))

    def mat_blend(self, ma, mb, w):
        return L.check(self.lib.glome_sb_mat_blend(self.h, ma, mb, float(w)))

    def tex_uniform(self, mat):
        return L.check(self.lib.glome_sb_tex_uniform(self.h, mat))

    def tex_stripe_blend(self, ma, mb, axis):
        return L.check(self.lib.glome_sb_tex_stripe_blend(self.h, ma, mb, _d3(axis)))

    def tex_perlin_blend(self, ma, mb, scale_):
        return L.check(self.lib.glome_sb_tex_perlin_blend(self.h, ma, mb, float(scale_)))

    def t_matte(self, rgb):
        """TestScene.hs:236-240"""
        return self.tex_uniform(self.mat_surface(rgb, 1.0, 0.2, 1.0, 0.0, 0.0))

    def light(self, pos, color):
        return L.check(self.lib.glome_sb_light(self.h, _d3(pos), _d3(color)))

    def lightset(self, lights):
        a = _i32(lights)
        return L.check(self.lib.glome_sb_lightset(self.h, len(a), _ptr(a)))

    def config_scene(self, config, n=0, seed=0):
        """BASELINE.json config scenes -> (root item, GlomeCamera, recurs)."""
        cam = L.GlomeCamera()
        rec = C.c_int32(0)
        root = L.check(self.lib.glome_sb_config_scene(self.h, config, int(n), int(seed), C.byref(cam), C.byref(rec)))
        return root, cam, rec.value

    def flatten(self, root):
        fs = L.GlomeFlatScene()
        L.check(self.lib.glome_sb_flatten(self.h, root, C.byref(fs)))
        return fs


class Scene:
    """A FlatScene resident on one B200.  No CPU fallback: creation fails without a CUDA device."""

    def __init__(self, flat, device=0, precision=64):
        """precision=32: the optional FP32 mode (Flt = Float, Vec.hs:7-9); everything else about the handle is the same."""
        self.lib = L.load()
        h = C.c_void_p()
        if precision not in (32, 64):
            raise ValueError("precision is 64 (the reference's Flt = Double) or 32")
        create = self.lib.glome_scene_create if precision == 64 else self.lib.glome_scene_create_f32
        L.check(create(C.byref(flat), int(device), C.byref(h)))
        self.h = h
        self.device = device
        self.precision = precision
        self.scene_class = flat.scene_class

    def set_option(self, option, value):
        """L.OPT_SEG_CONCURRENT / L.OPT_AA_SPECULATE (include/glome_cuda.h): measurement switches, never the frame."""
        L.check(self.lib.glome_scene_set_option(self.h, int(option), int(value)))

    def close(self):
        if getattr(self, "h", None):
            self.lib.glome_scene_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    @staticmethod
    def _tmax(tmax, n):
        t = _f64(np.atleast_1d(tmax))
        if t.size == 1:
            return t, 0
        assert t.size == n
        return t, 1

    def rayint(self, rays, tmax=1000000.0):
        rays = _f64(rays, (-1, 6))
        t, stride = self._tmax(tmax, len(rays))
        out = np.zeros(len(rays), dtype=HIT_DTYPE)
        L.check(self.lib.glome_rayint_batch(self.h, len(rays), _ptr(rays), _ptr(t), stride, _ptr(out)))
        return out

    def shadow(self, rays, tmax=1000000.0):
        rays = _f64(rays, (-1, 6))
        t, stride = self._tmax(tmax, len(rays))
        out = np.zeros(len(rays), dtype=np.uint8)
        L.check(self.lib.glome_shadow_batch(self.h, len(rays), _ptr(rays), _ptr(t), stride, _ptr(out)))
        return out

    def inside(self, pts):
        pts = _f64(pts, (-1, 3))
        out = np.zeros(len(pts), dtype=np.uint8)
        L.check(self.lib.glome_inside_batch(self.h, len(pts), _ptr(pts), _ptr(out)))
        return out

    def trace(self, rays, tmax=1000000.0, recurs=3, want_hits=False):
        rays = _f64(rays, (-1, 6))
        t, stride = self._tmax(tmax, len(rays))
        rgba = np.zeros((len(rays), 4))
        depth = np.zeros(len(rays))
        hits = np.zeros(len(rays), dtype=HIT_DTYPE) if want_hits else None
        L.check(self.lib.glome_trace_batch(self.h, len(rays), _ptr(rays), _ptr(t), stride, int(recurs), _ptr(rgba),
                                           _ptr(depth), _ptr(hits) if want_hits else None))
        return (rgba, depth, hits) if want_hits else (rgba, depth)

    def debug_count(self, rays, tmax=1000000.0):
        """rayint_debug's box count per ray (Bih.hs:378-412)."""
        rays = _f64(rays, (-1, 6))
        t, stride = self._tmax(tmax, len(rays))
        out = np.zeros(len(rays), dtype=np.int32)
        L.check(self.lib.glome_debug_count_batch(self.h, len(rays), _ptr(rays), _ptr(t), stride, _ptr(out)))
        return out

    def get_tags(self, cam, width, height, px, py, recurs=3):
        """getTags' (Glome.hs:410-414): (tag ids of the trace result under the pixel, truncated flag, primary hit)."""
        cap = 16
        tags = (C.c_int32 * cap)()
        n, trunc = C.c_int32(), C.c_int32()
        hit = L.GlomeHit()
        L.check(self.lib.glome_get_tags(self.h, C.byref(cam), width, height, px, py, recurs, tags, cap,
                                        C.byref(n), C.byref(trunc), C.byref(hit)))
        return list(tags[:min(n.value, cap)]), bool(trunc.value), hit

    def render(self, cam, width, height, opts=None, want_rgb8=False, out=None, rgb8_out=None):
        """renderTiles: returns (tcolor[h,w,5], rgb8[h,w] or None, stats)."""
        if opts is None:
            opts = render_opts()
        tc = out if out is not None else np.zeros((height, width, 5))
        rgb = rgb8_out if rgb8_out is not None else (np.zeros((height, width), dtype=np.uint32) if want_rgb8 else None)
        st = L.GlomeRenderStats()
        L.check(self.lib.glome_render(self.h, C.byref(cam), width, height, C.byref(opts), _ptr(tc),
                                      _ptr(rgb) if rgb is not None else None, C.byref(st)))
        return tc, rgb, st

    def render_ptr(self, cam, width, height, opts, tcolor_ptr, rgb8_ptr, dev=False, stream=0, want_stats=True):
        """Raw-pointer form used by bench.py (pinned host buffers, or device buffers when dev=True)."""
        st = L.GlomeRenderStats()
        if dev:
            L.check(self.lib.glome_render_dev(self.h, C.byref(cam), width, height, C.byref(opts), C.c_void_p(tcolor_ptr),
                                              C.c_void_p(rgb8_ptr) if rgb8_ptr else None,
                                              C.byref(st) if want_stats else None, C.c_void_p(stream)))
        else:
            L.check(self.lib.glome_render(self.h, C.byref(cam), width, height, C.byref(opts), C.c_void_p(tcolor_ptr),
                                          C.c_void_p(rgb8_ptr) if rgb8_ptr else None, C.byref(st)))
        return st


class MultiScene:
    """A FlatScene replicated on several GPUs of this process (glome_multi_*): tile i is rendered by device i % N and
    the tiles are gathered on the first device by peer copies.  No torch, no NCCL: the C-ABI a Haskell caller uses."""

    def __init__(self, flat, devices):
        self.lib = L.load()
        self.flat = flat
        devs = (C.c_int32 * len(devices))(*devices)
        h = C.c_void_p()
        L.check(self.lib.glome_multi_create(C.byref(flat), len(devices), devs, C.byref(h)))
        self.h = h

    def close(self):
        if self.h:
            self.lib.glome_multi_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def render(self, cam, width, height, opts=None, want_rgb8=False, want_tcolor=True):
        opts = opts or render_opts()
        tc = np.zeros((height, width, 5)) if want_tcolor else None
        rgb = np.zeros((height, width), dtype=np.uint32) if want_rgb8 else None
        st = L.GlomeRenderStats()
        L.check(self.lib.glome_multi_render(self.h, C.byref(cam), width, height, C.byref(opts),
                                            _ptr(tc) if want_tcolor else None, _ptr(rgb) if want_rgb8 else None, C.byref(st)))
        return tc, rgb, st


def camera_rays(cam, width, height, xs, ys):
    """get_rayint's ray for pixel coordinates (Glome.hs:27-33, 119-128); numpy, for tests."""
    xs = np.asarray(xs, dtype=np.float64)
    ys = np.asarray(ys, dtype=np.float64)
    wf, hf = float(width), float(height)
    xc = (((xs / wf) * 2) - 1) * (wf / hf)
    yc = -(((ys / hf) * 2) - 1)
    pos, fwd = np.array(cam.pos[:]), np.array(cam.fwd[:])
    up, right = np.array(cam.up[:]), np.array(cam.right[:])
    d = fwd[None, :] + right[None, :] * (-xc)[:, None] + up[None, :] * yc[:, None]
    inv = 1.0 / np.sqrt((d[:, 0] * d[:, 0]) + (d[:, 1] * d[:, 1]) + (d[:, 2] * d[:, 2]))
    d = d * inv[:, None]
    rays = np.empty((len(xs), 6))
    rays[:, :3] = pos[None, :]
    rays[:, 3:] = d
    return rays
