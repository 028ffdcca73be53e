"""The scene-graph machine (glome_b200/csrc/glome_gen.cuh) against the oracle, on the CPU.

The machine replaces the reference's mutual recursion (rayint / shadow / inside / get_metainfo / trace / materialShader)
by loops over explicit stacks; the product runs it only inside its sm_100a kernels.  tests/tools/gen_host.cpp compiles the
same header with g++ (test infrastructure), so that its control flow can be checked here, without a GPU, on random scene
graphs that use EVERY node kind and material of SURVEY.md section 8(a) -- Bound, InnerBound, NoShadow, OnlyShadow,
AdditiveLayers, Warp, Mesh inside a general scene, nested Instances ... -- and on the two general BASELINE.json scenes.
Everything is bit-exact (both sides use libm's pow)."""
import numpy as np
import pytest

import glome_b200 as G
import genhost as H
import oracle as O
import scenegen as SG

HIT_FIELDS = ("hit", "prim", "sub", "t", "pos", "norm", "ntex", "tex", "ntag", "tag", "flags")


def same_hits(g, o):
    for f in HIT_FIELDS:
        assert np.array_equal(g[f], o[f]), f


def check_scene(fs, cam, seed, recurs, w=48, h=32):
    osc, hs = O.OracleScene(fs), H.HostGenScene(fs)
    rays = SG.query_rays(cam, w, h, seed)
    same_hits(hs.rayint(rays), osc.rayint(rays))
    # a bounded distance per ray, and rays that are NOT normalised (the reference's refracted rays are not, Shader.hs:145)
    rng = np.random.default_rng(seed + 1)
    tmax = rng.uniform(0.5, 30.0, len(rays))
    same_hits(hs.rayint(rays, tmax), osc.rayint(rays, tmax))
    long_rays = rays.copy()
    long_rays[:, 3:] *= rng.uniform(0.3, 3.0, (len(rays), 1))
    same_hits(hs.rayint(long_rays), osc.rayint(long_rays))
    assert np.array_equal(hs.shadow(rays, 30.0), osc.shadow(rays, 30.0))
    assert np.array_equal(hs.shadow(long_rays, tmax), osc.shadow(long_rays, tmax))
    pts = rng.uniform(-5, 5, (3000, 3))
    assert np.array_equal(hs.inside(pts), osc.inside(pts))
    assert np.array_equal(hs.debug_count(rays), osc.debug_count(rays))
    ro, do, ho, to = osc.trace(rays, recurs=recurs, want_hits=True, want_tags=True)
    rg, dg, hg, tg, cnt = hs.trace(rays, recurs=recurs, want_tags=True)
    assert np.array_equal(rg, ro) and np.array_equal(dg, do)
    same_hits(hg, ho)
    assert np.array_equal(tg, to[:, :17])
    return osc, hs, (ro, ho, to, cnt)


@pytest.mark.parametrize("seed", range(12))
def test_random_scene_graphs(seed):
    b, fs, cam = SG.random_scene(seed)
    assert fs.scene_class == 0  # general
    check_scene(fs, cam, seed, recurs=4)


def test_every_node_kind_and_material_is_exercised():
    kinds, mats = set(), set()
    import glome_b200.scene as SC
    import ctypes as C
    for seed in range(12):
        b, fs, cam = SG.random_scene(seed)
        kinds |= set(int(t) for t in SC.FlatView(fs).nodes["type"])
        m = np.ctypeslib.as_array(C.cast(fs.materials, C.POINTER(C.c_int32)), shape=(fs.n_materials * 24,)).reshape(-1, 24)
        mats |= set(int(k) for k in m[:, 0])
    assert kinds == set(range(21)), sorted(set(range(21)) - kinds)
    assert mats == set(range(6))


@pytest.mark.parametrize("config,n,recurs", [(1, 0, 3), (4, 6, 5)])
def test_baseline_general_scenes(config, n, recurs):
    b = G.SceneBuilder()
    root, cam, rec = b.config_scene(config, n)
    assert rec == recurs
    fs = b.flatten(root)
    osc, hs, (ro, ho, to, cnt) = check_scene(fs, cam, 7, rec, w=96, h=64)
    assert cnt[1] > 0  # secondary rays were traced


def test_many_csg_crossings_fold_pending_offsets():
    """A ray that has to cross 140 slabs of the subtrahend before it reaches the minuend re-issues the Difference 280
    times (rayint_advance, Solid.hs:85-91): more pending depth offsets than the control stack holds.  The machine then
    folds the newest offset into the one on top and flags the ray (GLOME_HITFLAG_CSG_OVERFLOW): the hit is the oracle's,
    its depth agrees to rounding.  (Round 1's device recursion gave up after 48 crossings.)"""
    b = G.SceneBuilder()
    slabs = [b.box((-1, -1, 2 * i + 0.5), (1, 1, 2 * i + 1.5)) for i in range(140)]
    body = b.box((-0.5, -0.5, 300), (0.5, 0.5, 310))
    root = b.tex(b.difference(body, b.group(slabs)), b.t_matte((0.5, 0.5, 0.5)))
    b.light((0, 50, 0), (100, 100, 100))
    fs = b.flatten(root)
    osc, hs = O.OracleScene(fs), H.HostGenScene(fs)
    rays = np.array([[0.1, 0.2, -5.0, 1e-4, 2e-4, 1.0], [0.1, 0.2, 100.2, 1e-4, 2e-4, 1.0], [0.1, 0.2, 290.0, 1e-4, 2e-4, 1.0]])
    rays[:, 3:] /= np.linalg.norm(rays[:, 3:], axis=1)[:, None]
    o, g = osc.rayint(rays), hs.rayint(rays)
    assert o["hit"].all() and np.array_equal(g["hit"], o["hit"]) and np.array_equal(g["prim"], o["prim"])
    assert np.allclose(g["t"], o["t"], rtol=1e-12, atol=0)
    assert np.array_equal(g["pos"], o["pos"]) and np.array_equal(g["norm"], o["norm"])
    assert g["flags"][0] == 2 and g["flags"][2] == 0 and np.array_equal(g["t"][2:], o["t"][2:])
    st = osc.stats()


def _tie_group_scene(nested=False):
    """A plain `group` big enough for the implicit BIH (glome_tagmap.h), full of exact ties: a 6 x 6 board of boxes that
    share faces (a ray through a shared vertical face meets both at the same depth), every sphere listed twice with
    different tags, an instanced cone and an instanced box, one Void.  The reference's fold gives equal depths to the LATER
    list element (Solid.hs:37-44); the tags say which element won."""
    b = G.SceneBuilder()
    items, tag = [], 0
    mat = b.t_matte((0.6, 0.6, 0.6))
    for i in range(6):
        for j in range(6):
            items.append(b.tag(b.box((i, 0, j), (i + 1, 0.5 + 0.25 * ((i + j) % 3), j + 1)), tag)); tag += 1
    for k in range(5):
        c, r = (0.7 + 1.1 * k, 1.6, 2.5 + 0.3 * k), 0.45
        items.append(b.tag(b.sphere(c, r), tag)); tag += 1
        items.append(b.tag(b.sphere(c, r), tag)); tag += 1   # exact duplicate, listed later: it must win
    items.append(b.void())
    items.append(b.tag(b.transform(b.cone((0, 0, 0), 0.5, (0, 1.2, 0), 0.1), [G.translate((3, 1.0, 5.2))]), tag)); tag += 1
    items.append(b.tag(b.transform(b.box((-0.3, -0.3, -0.3), (0.3, 0.3, 0.3)), [G.rotate((0, 1, 0), 0.4), G.translate((5, 1.5, 1))]), tag)); tag += 1
    grp = b.group(items)
    root = b.tex(grp, mat)
    if nested:  # under an Instance and a Difference, like TestScene's chessboard
        root = b.tex(b.difference(b.transform(grp, [G.scale((1.5, 1.0, 1.5))]), b.sphere((4, 1.0, 4), 1.3)), mat)
    b.light((3, 9, -4), (60, 60, 60))
    b.light((-5, 7, 9), (40, 40, 50))
    return b, b.flatten(root)


def _tie_rays(seed):
    rng = np.random.default_rng(seed)
    o = np.column_stack([rng.uniform(-1, 7, 6000), rng.uniform(2.5, 6, 6000), rng.uniform(-1, 7, 6000)])
    d = np.column_stack([rng.normal(0, 0.4, 6000), -np.abs(rng.normal(1, 0.2, 6000)), rng.normal(0, 0.4, 6000)])
    # rays that run exactly inside shared faces (x = const planes), axis-parallel rays (zero direction components: the
    # machine takes the plain list for those), rays along the board
    k = np.arange(600)
    face = np.column_stack([1.0 + (k % 5), np.full(600, 4.0), 0.2 + 0.009 * k, np.zeros(600), -np.ones(600), 0.3 * np.sin(k)])
    down = np.column_stack([0.5 + 0.01 * k, np.full(600, 5.0), 0.5 + 0.009 * k, np.zeros(600), -np.ones(600), np.zeros(600)])
    flat = np.column_stack([np.full(600, -2.0), 0.3 + 0.001 * k, 0.05 + 0.0098 * k, np.ones(600), 0.01 * np.cos(k), 0.02 * np.sin(k)])
    rays = np.vstack([np.hstack([o, d]), face, down, flat])
    rays[:, 3:] /= np.linalg.norm(rays[:, 3:], axis=1)[:, None]
    return rays


@pytest.mark.parametrize("nested", [False, True])
def test_large_plain_group_keeps_the_list_folds_ties(nested):
    b, fs = _tie_group_scene(nested)
    osc, hs = O.OracleScene(fs), H.HostGenScene(fs)
    rays = _tie_rays(5)
    o, g = osc.rayint(rays), hs.rayint(rays)
    same_hits(g, o)
    assert o["hit"].mean() > 0.4
    if not nested:  # the duplicated spheres: always the later copy's tag (odd offsets 37, 39, ...)
        sph = o["hit"].astype(bool) & (o["tag"][:, 0] >= 36) & (o["tag"][:, 0] < 46)
        assert sph.sum() > 50 and np.all((o["tag"][sph, 0] - 36) % 2 == 1)
    tmax = np.random.default_rng(6).uniform(1.0, 8.0, len(rays))
    same_hits(hs.rayint(rays, tmax), osc.rayint(rays, tmax))
    assert np.array_equal(hs.shadow(rays, 6.0), osc.shadow(rays, 6.0))
    ro, do = osc.trace(rays, recurs=3)
    rg, dg = hs.trace(rays, recurs=3)[:2]
    assert np.array_equal(rg, ro) and np.array_equal(dg, do)
    assert np.array_equal(hs.debug_count(rays), osc.debug_count(rays))
