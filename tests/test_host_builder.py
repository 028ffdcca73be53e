"""Host mirror of the GlomeTrace construction API (glome_b200/csrc/host_builder.cpp): the product's
index-based bih / mesh builders must produce exactly the tree of the oracle's literal restatement of
Bih.hs:211-324 / Mesh.hs:50-134, and the constructors must follow the reference's semantics."""
import ctypes as C
import re
import os

import numpy as np
import pytest

import glome_b200 as G
from glome_b200 import _lib as L
import oracle as O

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def product_bih(bboxes):
    lib = L.load()
    bboxes = np.ascontiguousarray(bboxes, dtype=np.float64).reshape(-1, 6)
    nodes = C.POINTER(L.GlomeBihNode)()
    leaves = C.POINTER(C.c_int32)()
    order = C.POINTER(C.c_int32)()
    nn, nl, root = C.c_int32(), C.c_int32(), C.c_int32()
    bb = (C.c_double * 6)()
    L.check(lib.glome_bih_build(len(bboxes), bboxes.ctypes.data_as(C.c_void_p), C.byref(nodes), C.byref(nn),
                                C.byref(leaves), C.byref(nl), C.byref(order), C.byref(root), bb))
    node_dt = np.dtype([("lsplit", "<f8"), ("rsplit", "<f8"), ("axis", "<i4"), ("left", "<i4"), ("right", "<i4"),
                        ("pad", "<i4")])
    res = dict(nodes=np.frombuffer(C.string_at(nodes, nn.value * 32), dtype=node_dt).copy(),
               leaves=np.frombuffer(C.string_at(leaves, nl.value * 8), dtype=np.int32).copy().reshape(-1, 2),
               order=np.frombuffer(C.string_at(order, len(bboxes) * 4), dtype=np.int32).copy(), root=root.value,
               bb=np.array(bb[:]))
    for p in (nodes, leaves, order):
        lib.glome_free(C.cast(p, C.c_void_p))
    return res


def product_mesh(verts, tris):
    lib = L.load()
    verts = np.ascontiguousarray(verts, dtype=np.float64).reshape(-1, 3)
    tris = np.ascontiguousarray(tris, dtype=np.int32).reshape(-1, 8)
    nodes = C.POINTER(L.GlomeBvhNode)()
    leafpool = C.POINTER(C.c_int32)()
    leafoff = C.POINTER(C.c_int32)()
    nn, nlp, nl, root = C.c_int32(), C.c_int32(), C.c_int32(), C.c_int32()
    bb = (C.c_double * 6)()
    L.check(lib.glome_mesh_build(len(verts), verts.ctypes.data_as(C.c_void_p), len(tris), tris.ctypes.data_as(C.c_void_p),
                                 C.byref(nodes), C.byref(nn), C.byref(leafpool), C.byref(nlp), C.byref(leafoff),
                                 C.byref(nl), C.byref(root), bb))
    node_dt = np.dtype([("lbb", "<f8", (6,)), ("rbb", "<f8", (6,)), ("left", "<i4"), ("right", "<i4"),
                        ("pad", "<i4", (6,))])
    res = dict(nodes=np.frombuffer(C.string_at(nodes, nn.value * 128), dtype=node_dt).copy(),
               leafpool=np.frombuffer(C.string_at(leafpool, nlp.value * 4), dtype=np.int32).copy(),
               leafoff=np.frombuffer(C.string_at(leafoff, nl.value * 4), dtype=np.int32).copy(), root=root.value,
               bb=np.array(bb[:]))
    for p in (nodes, leafpool, leafoff):
        lib.glome_free(C.cast(p, C.c_void_p))
    return res


def sphere_boxes(n, seed, half=10.0, rmax=0.5):
    rng = np.random.default_rng(seed)
    c = rng.uniform(-half, half, size=(n, 3))
    r = rng.uniform(0.05, rmax, size=(n, 1))
    return np.hstack([c - r, c + r])


@pytest.mark.parametrize("n,seed", [(1, 1), (3, 2), (4, 3), (17, 4), (1000, 5), (50000, 6)])
def test_bih_tree_equals_oracle(n, seed):
    bb = sphere_boxes(n, seed)
    a, b = product_bih(bb), O.bih_build(bb)
    assert a["root"] == b["root"]
    assert np.array_equal(a["bb"], b["bb"])
    assert np.array_equal(a["order"], b["order"])
    assert np.array_equal(a["leaves"], b["leaves"])
    assert a["nodes"].tobytes() == b["nodes"].tobytes()


def test_bih_big_small_split_and_mixed_sizes():
    # a few huge objects among many small ones exercises the big/small candidate (Bih.hs:223, 285)
    bb = sphere_boxes(2000, 7, rmax=0.2)
    bb[:5] = np.array([[-9, -9, -9, 9, 9, 9]] * 5) + np.arange(5)[:, None] * 0.01
    a, b = product_bih(bb), O.bih_build(bb)
    assert a["nodes"].tobytes() == b["nodes"].tobytes() and np.array_equal(a["order"], b["order"])


def test_bih_infinite_bbox_is_an_error():
    bb = np.array([[-1e6, -1e6, -1e6, 1e6, 1e6, 1e6], [0, 0, 0, 1, 1, 1]], dtype=np.float64)  # a bare plane
    with pytest.raises(L.GlomeError) as e:
        product_bih(bb)
    assert "infinite bounding box" in str(e.value)
    with pytest.raises(RuntimeError):
        O.bih_build(bb)


def grid_mesh(g, seed):
    rng = np.random.default_rng(seed)
    xs, zs = np.meshgrid(np.linspace(-5, 5, g + 1), np.linspace(-5, 5, g + 1), indexing="xy")
    ys = rng.uniform(0, 1.0, size=xs.shape)
    verts = np.stack([xs, ys, zs], axis=-1).reshape(-1, 3)
    tris = []
    for j in range(g):
        for i in range(g):
            v00 = j * (g + 1) + i
            v10, v01, v11 = v00 + 1, v00 + g + 1, v00 + g + 2
            tris.append([v00, v01, v10, -1, -1, -1, -1, -1])
            tris.append([v10, v01, v11, -1, -1, -1, -1, -1])
    return verts, np.array(tris, dtype=np.int32)


@pytest.mark.parametrize("g,seed", [(1, 1), (2, 2), (9, 3), (64, 4)])
def test_mesh_tree_equals_oracle(g, seed):
    verts, tris = grid_mesh(g, seed)
    a, b = product_mesh(verts, tris), O.mesh_build(verts, tris)
    assert a["root"] == b["root"] and np.array_equal(a["bb"], b["bb"])
    assert np.array_equal(a["leafpool"], b["leafpool"]) and np.array_equal(a["leafoff"], b["leafoff"])
    assert a["nodes"].tobytes() == b["nodes"].tobytes()


def test_group_flattens_and_drops_void():
    b = G.SceneBuilder()
    s1, s2, s3 = b.sphere((0, 0, 0), 1), b.sphere((3, 0, 0), 1), b.sphere((6, 0, 0), 1)
    assert b.group([s1]) == s1                      # group (sld:[]) = sld   (Solid.hs:295)
    g = b.group([b.group([s1, s2]), b.void(), s3])  # flatten_group: nested groups are smashed
    fv = G.FlatView(b.flatten(g))
    root = fv.nodes[fv.root]
    assert root["type"] == L.GROUP and root["b"] == 3
    assert [fv.nodes[root["a"] + i]["type"] for i in range(3)] == [L.SPHERE] * 3
    fv = G.FlatView(b.flatten(b.group([])))
    assert fv.nodes[fv.root]["type"] == L.VOID      # group [] = Void (Solid.hs:294)
    assert G.FlatView(b.flatten(b.bih([]))).nodes[0]["type"] == L.VOID  # bih [] = Void (Bih.hs:310)


def test_transform_of_instance_merges():
    b = G.SceneBuilder()
    s = b.sphere((0, 0, 0), 1)
    t1 = b.transform(s, [G.translate((1, 0, 0))])
    t2 = b.transform(t1, [G.scale((2, 2, 2))])      # Solid.hs:494-496: one Instance, composed matrix
    fv = G.FlatView(b.flatten(t2))
    root = fv.nodes[fv.root]
    assert root["type"] == L.INSTANCE and fv.nodes[root["a"]]["type"] == L.SPHERE
    m = fv.dpool[root["b"]:root["b"] + 12].reshape(3, 4)
    assert np.allclose(m, [[2, 0, 0, 2], [0, 2, 0, 0], [0, 0, 2, 0]])  # translate first, then scale


def test_transform_triangle_moves_vertices():
    b = G.SceneBuilder()
    t = b.transform(b.triangle((0, 0, 0), (1, 0, 0), (0, 1, 0)), [G.translate((0, 0, 5))])
    fv = G.FlatView(b.flatten(t))
    root = fv.nodes[fv.root]
    assert root["type"] == L.TRIANGLE and fv.dpool[root["a"] + 2] == 5.0  # Triangle.hs:164-168


def test_rotate_requires_unit_axis():
    with pytest.raises(L.GlomeError):
        G.rotate((1, 1, 0), 0.3)


def test_flatten_transform_bih_strips_bounds_and_pushes_xfms():
    # the oak idiom (TestScene.hs:110): Bound objects vanish, Instances of groups become per-leaf Instances
    b = G.SceneBuilder()
    leaf = b.tex(b.sphere((0, 0, 0), 0.4), b.t_matte((0, 1, 0)))
    inner = b.bound_object(b.sphere((0, 0.5, 0), 0.5), b.group([b.cone((0, 0, 0), 0.1, (0, 1, 0), 0.05),
                                                                b.transform(leaf, [G.translate((0, 1, 0))])]))
    outer = b.bound_object(b.sphere((0, 1, 0), 1), b.group([b.cone((0, 0, 0), 0.2, (0, 1, 0), 0.1),
                                                            b.transform(inner, [G.scale((0.9, 0.9, 0.9)),
                                                                                G.translate((0, 1, 0))])]))
    t = b.flatten_transform_bih(outer)
    fv = G.FlatView(b.flatten(t))
    root = fv.nodes[fv.root]
    assert root["type"] == L.BIH
    types = sorted(int(fv.nodes[i]["type"]) for i in range(len(fv.nodes)))
    assert L.BOUND not in types and L.GROUP not in types
    assert types.count(L.INSTANCE) == 3 and types.count(L.CONE) == 2 and types.count(L.TEX) == 1


def test_bound_of_primitives():
    b = G.SceneBuilder()
    assert b.bound(b.sphere((1, 2, 3), 0.5)).tolist() == [0.5, 1.5, 2.5, 1.5, 2.5, 3.5]
    assert b.bound(b.plane((0, 0, 0), (0, 1, 0))).tolist() == [-1e6] * 3 + [1e6] * 3
    tb = b.bound(b.triangle((0, 0, 0), (1, 0, 0), (0, 1, 0)))
    assert tb.tolist() == [-1e-4, -1e-4, -1e-4, 1 + 1e-4, 1 + 1e-4, 1e-4]
    assert b.bound(b.cylinder_z(2, 1, 3)).tolist() == [-2, -2, 1, 2, 2, 3]
    ib = b.bound(b.transform(b.sphere((0, 0, 0), 1), [G.translate((5, 0, 0))]))
    assert np.allclose(ib, [4 - 1e-4, -1 - 1e-4, -1 - 1e-4, 6 + 1e-4, 1 + 1e-4, 1 + 1e-4], atol=1e-12)


def test_scene_class():
    b = G.SceneBuilder()
    root, cam, rec = b.config_scene(2, 500)
    assert b.flatten(root).scene_class == L.CLASS_FLAT
    b = G.SceneBuilder()
    root, cam, rec = b.config_scene(4, 3)
    assert b.flatten(root).scene_class == L.CLASS_GENERAL and rec == 5
    b = G.SceneBuilder()
    root, cam, rec = b.config_scene(3, 2000)
    fs = b.flatten(root)
    assert fs.scene_class == L.CLASS_FLAT and fs.n_bvhnodes > 0


def test_testscene_inventory():
    # 9 261 lattice spheres, oak = 1 023 cones/cylinders + 1 024 leaf spheres, 64 chess boxes
    b = G.SceneBuilder()
    root, cam, rec = b.config_scene(1)
    fs = b.flatten(root)
    fv = G.FlatView(fs)
    t = fv.nodes["type"]
    assert rec == 3 and fs.scene_class == L.CLASS_GENERAL
    assert int((t == L.SPHERE).sum()) >= 9261 + 1024
    assert int(((t == L.CONE) | (t == L.CYLINDER)).sum()) >= 1023 + 1
    assert int((t == L.BOX).sum()) >= 64 + 3
    assert int((t == L.PLANE).sum()) == 12 + 20
    assert fs.n_lights == 2


def test_abi_exports_every_declared_symbol():
    hdr = open(os.path.join(ROOT, "include", "glome_cuda.h")).read()
    declared = set(re.findall(r"\b(glome_[a-z0-9_]+)\s*\(", hdr))
    declared -= set(re.findall(r"static inline [A-Za-z0-9_ ]*?\b(glome_[a-z0-9_]+)\s*\(", hdr))  # header-only helpers
    lib = L.load()
    missing = [n for n in sorted(declared) if not hasattr(lib, n)]
    assert not missing, missing
    assert declared == set(L.SIGNATURES), declared ^ set(L.SIGNATURES)


def test_no_device_is_an_error_not_a_fallback():
    lib = L.load()
    if lib.glome_device_count() > 0:
        pytest.skip("a CUDA device is present")
    b = G.SceneBuilder()
    fs = b.flatten(b.sphere((0, 0, 0), 1))
    with pytest.raises(L.GlomeError) as e:
        G.Scene(fs)
    assert e.value.code == L.ENODEV


def test_product_never_imports_oracle():
    for dirpath, _, files in os.walk(os.path.join(ROOT, "glome_b200")):
        for f in files:
            if f.endswith((".py", ".cpp", ".cu", ".cuh", ".h")):
                src = open(os.path.join(dirpath, f), errors="replace").read()
                assert "oracle" not in src.lower().replace("no oracle", "").replace("touches oracle/", "") or \
                    "import oracle" not in src and "glome_oracle" not in src, f


def _flat_bytes(fs):
    v = G.scene.FlatView(fs)
    return [v.nodes.tobytes(), v.bihnodes.tobytes(), v.bvhnodes.tobytes(), v.ipool.tobytes(), v.dpool.tobytes(), v.root,
            v.scene_class, v.max_depth]


@pytest.mark.parametrize("n,seed", [(1, 1), (2, 2), (7, 3), (500, 4), (20000, 5)])
def test_prebuilt_bih_import_equals_own_build(n, seed):
    """glome_sb_bih_prebuilt: a `Bih bb root` the caller already built (Bih.hs:51-57, 309-324), imported as a pre-order
    stream, flattens to exactly the FlatScene glome_sb_bih produces by building the tree itself."""
    rng = np.random.default_rng(seed)
    c, r = rng.uniform(-10, 10, size=(n, 3)), rng.uniform(0.05, 0.7, size=n)
    tree = G.scene.bih_build(np.hstack([c - r[:, None], c + r[:, None]]))
    kinds, splits, pos = G.scene.bih_preorder_stream(tree)
    a, b = G.SceneBuilder(), G.SceneBuilder()
    own = a.flatten(a.bih([a.tag(s, i) if i % 3 == 0 else s for i, s in enumerate(a.spheres(c, r))]))
    ids = [b.tag(s, i) if i % 3 == 0 else s for i, s in enumerate(b.spheres(c, r))]
    imp = b.flatten(b.bih_prebuilt([ids[p] for p in pos], kinds, splits, tree["bb"]))
    assert _flat_bytes(own) == _flat_bytes(imp)


def test_prebuilt_mesh_import_equals_own_build():
    """glome_sb_mesh_prebuilt: `Branch lbb rbb l r | Leaf [Tri]` (Mesh.hs:36-42) in pre-order == glome_sb_mesh's own tree."""
    for g, seed in [(1, 1), (3, 2), (40, 3)]:
        verts, tris = grid_mesh(g, seed)
        tris = tris.copy()
        tris[:, 7] = np.arange(len(tris)) % 5
        tree = G.scene.mesh_build(verts, tris)
        kinds, boxes, leaf_tris = G.scene.mesh_preorder_stream(tree)
        a, b = G.SceneBuilder(), G.SceneBuilder()
        own = a.flatten(a.mesh(verts, [], tris, [], list(range(5))))
        imp = b.flatten(b.mesh_prebuilt(verts, [], tris, [], list(range(5)), kinds, boxes, leaf_tris, tree["bb"]))
        assert _flat_bytes(own) == _flat_bytes(imp)


def test_raw_forms_equal_the_constructors():
    """glome_sb_list / _instance / _disc_raw / _difference_ex take the fields a constructed Haskell value stores
    ([s]; Instance s xfm, Solid.hs:386; Disc pos norm (r*r), Cone.hs:21; Difference a b Bool, Csg.hs:14)."""
    a, b = G.SceneBuilder(), G.SceneBuilder()
    x = G.compose([G.scale((1, 2, 3)), G.translate((1, 0, -2))])

    def scene(bl, raw):
        s1, s2 = bl.sphere((0, 0, 0), 1.0), bl.box((-1, -1, -1), (0.5, 2, 1))
        d = bl.disc_raw((0, 1, 0), (0, 1, 0), 0.7 * 0.7) if raw else bl.disc((0, 1, 0), (0, 1, 0), 0.7)
        df = bl.difference_ex(s1, s2, True) if raw else bl.difference(s1, s2)
        ins = bl.instance_raw(df, x) if raw else bl.transform(df, [x])
        return bl.flatten((bl.list_raw if raw else bl.group)([ins, d, bl.sphere((3, 0, 0), 0.5)]))

    assert _flat_bytes(scene(a, False)) == _flat_bytes(scene(b, True))
    c = G.SceneBuilder()
    keep = c.flatten(c.list_raw([c.list_raw([c.sphere((0, 0, 0), 1)]), c.void()]))  # no flattening, Void kept
    assert G.scene.FlatView(keep).nodes["type"].tolist().count(9) == 2  # GLOME_GROUP
    d0, d1 = G.SceneBuilder(), G.SceneBuilder()
    f0 = d0.flatten(d0.difference_ex(d0.sphere((0, 0, 0), 1), d0.sphere((1, 0, 0), 1), False))
    f1 = d1.flatten(d1.difference(d1.sphere((0, 0, 0), 1), d1.sphere((1, 0, 0), 1)))
    n0, n1 = G.scene.FlatView(f0).nodes, G.scene.FlatView(f1).nodes
    assert n0[f0.root]["c"] == 0 and n1[f1.root]["c"] == 1  # difference_retexture vs difference (Csg.hs:26-30)


def test_prebuilt_streams_are_validated():
    b = G.SceneBuilder()
    s = b.spheres(np.zeros((3, 3)), np.ones(3))
    bb = [-1, -1, -1, 1, 1, 1]
    for kinds, splits in [([0, -2], [[0, 0]] * 2),             # branch with one child
                          ([-3], [[0, 0]]),                    # leaf holds fewer items than passed
                          ([0, -3, -3], [[0, 0]] * 3),         # leaves hold more than passed
                          ([5, -2, -3], [[0, 0]] * 3),         # axis out of range
                          ([-4, -1], [[0, 0]] * 2)]:           # records after the tree
        with pytest.raises(L.GlomeError):
            b.bih_prebuilt(s, kinds, splits, bb)


def test_host_library_stands_alone():
    """libglomehost.so (scene construction) has no CUDA in it, and a process that sets GLOME_HOST_ONLY=1 -- bench.py's CPU
    reference arm -- builds its scene and runs the oracle without ever mapping libglomecuda.so."""
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    host = os.path.join(root, "glome_b200", "_build", "libglomehost.so")
    deps = subprocess.run(["ldd", host], capture_output=True, text=True).stdout
    assert "cuda" not in deps.lower() and "glomecuda" not in deps
    code = ("import os,sys; os.environ['GLOME_HOST_ONLY']='1'; sys.path.insert(0,%r); sys.path.insert(0,%r)\n"
            "import glome_b200 as G, oracle as O\n"
            "from glome_b200 import _lib as L\n"
            "b=G.SceneBuilder(); root,cam,rec=b.config_scene(4,3); fs=b.flatten(root)\n"
            "tc,_=O.OracleScene(fs).render(cam,32,24,G.render_opts(mode=L.MODE_ONE_RAY,recurs=rec))\n"
            "maps=open('/proc/self/maps').read()\n"
            "assert 'libglomehost' in maps and 'libglomecuda' not in maps and tc[...,3].max()>0\n"
            "assert not hasattr(L.load(),'glome_render')\n"
            "print('ok')\n") % (root, os.path.join(root, "tests"))
    out = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True)
    assert out.returncode == 0 and "ok" in out.stdout, out.stderr
