"""GPU <-> oracle on random scene graphs that use every node kind and material (tests/scenegen.py), through the C-ABI.

Covers what round 1 left untested on the device: Bound / InnerBound kept in the device scene, NoShadow / OnlyShadow
wrappers inside a bih and a group, AdditiveLayers, Warp with its own light set, a Mesh inside a general scene, light sets of
0 and of more than 8 lights, rays with more CSG crossings than the control stack holds, the tie rule of `nearest`
under drain-phase donation, and the frame buffers of glome_render / glome_render_dev at growing sizes."""
import ctypes as C

import numpy as np
import pytest

import glome_b200 as G
from glome_b200 import _lib as L
import oracle as O
import scenegen as SG

pytestmark = pytest.mark.gpu
RGB_TOL = 1e-6
HIT_FIELDS = ("hit", "prim", "sub", "t", "pos", "norm", "ntex", "tex", "ntag", "tag", "flags")


def same_hits(g, o):
    for f in HIT_FIELDS:
        assert np.array_equal(g[f], o[f]), f


@pytest.mark.parametrize("seed", range(8))
def test_random_scene_graphs(seed):
    b, fs, cam = SG.random_scene(seed)
    gs, osc = G.Scene(fs), O.OracleScene(fs)
    rays = SG.query_rays(cam, 64, 40, seed)
    rng = np.random.default_rng(seed + 1)
    tmax = rng.uniform(0.5, 30.0, len(rays))
    long_rays = rays.copy()
    long_rays[:, 3:] *= rng.uniform(0.3, 3.0, (len(rays), 1))  # not normalised, like the reference's refracted rays
    same_hits(gs.rayint(rays), osc.rayint(rays))
    same_hits(gs.rayint(rays, tmax), osc.rayint(rays, tmax))
    same_hits(gs.rayint(long_rays), osc.rayint(long_rays))
    assert np.array_equal(gs.shadow(rays, 30.0), osc.shadow(rays, 30.0))
    assert np.array_equal(gs.shadow(long_rays, tmax), osc.shadow(long_rays, tmax))
    pts = rng.uniform(-5, 5, (3000, 3))
    assert np.array_equal(gs.inside(pts), osc.inside(pts))
    assert np.array_equal(gs.debug_count(rays), osc.debug_count(rays))
    ro, do, ho = osc.trace(rays, recurs=4, want_hits=True)
    rg, dg, hg = gs.trace(rays, recurs=4, want_hits=True)
    assert np.array_equal(dg, do)
    same_hits(hg, ho)
    assert np.abs(rg - ro).max() <= RGB_TOL
    # the frame: one ray per pixel and adaptive AA
    for mode in (L.MODE_ONE_RAY, L.MODE_ADAPTIVE_AA):
        opts = G.render_opts(mode=mode, recurs=4)
        tg, _, st = gs.render(cam, 70, 66, opts)
        to, _ = osc.render(cam, 70, 66, opts)
        assert np.abs(tg[..., :4] - to[..., :4]).max() <= RGB_TOL and np.array_equal(tg[..., 4], to[..., 4])
        assert st.overflow_rays == 0


def test_pick_query_on_random_scenes():
    for seed in (1, 3, 5):
        b, fs, cam = SG.random_scene(seed)
        gs, osc = G.Scene(fs), O.OracleScene(fs)
        w, h = 48, 32
        ys, xs = np.mgrid[0:h:3, 0:w:3]
        rays = G.camera_rays(cam, w, h, xs.ravel(), ys.ravel())
        ro, do, to = osc.trace(rays, recurs=3, want_tags=True)
        n_with = 0
        for k, (x, y) in enumerate(zip(xs.ravel(), ys.ravel())):
            tags, trunc, hit = gs.get_tags(cam, w, h, int(x), int(y), recurs=3)
            want = to[k, 1:1 + min(to[k, 0], 16)].tolist()
            assert tags[:len(want)] == want, (seed, x, y, tags, want)
            n_with += len(want) > 0
        assert n_with > 0


@pytest.mark.parametrize("n_lights", [0, 11, 40])
def test_light_sets_of_0_and_more_than_8_lights(n_lights):
    b, fs, cam = SG.random_scene(2, n_lights=n_lights, with_warp=False)
    gs, osc = G.Scene(fs), O.OracleScene(fs)
    opts = G.render_opts(mode=L.MODE_ONE_RAY, recurs=3)
    tg, _, st = gs.render(cam, 64, 48, opts)
    to, _ = osc.render(cam, 64, 48, opts)
    assert np.abs(tg[..., :4] - to[..., :4]).max() <= RGB_TOL and np.array_equal(tg[..., 4], to[..., 4])
    assert (st.rays_shadow > 0) == (n_lights > 0)


def test_flat_scene_with_many_lights():
    """A flat-class scene (bih of spheres) with 11 and with 40 lights: the wavefront pipeline carries 32 occlusion bits
    per sample, beyond that the flat one-thread-per-sample tracer takes over; both equal the oracle."""
    for n_lights in (11, 40):
        b = G.SceneBuilder()
        rng = np.random.default_rng(n_lights)
        ids = b.spheres(rng.uniform(-8, 8, (3000, 3)), rng.uniform(0.1, 0.5, 3000))
        for _ in range(n_lights):
            b.light(rng.uniform(-30, 30, 3) + np.array([0, 40, 0]), rng.uniform(50, 200, 3))
        root = b.tex(b.bih(ids), b.t_matte((0.7, 0.6, 0.5)))
        fs = b.flatten(root)
        assert fs.scene_class == L.CLASS_FLAT
        cam = G.camera((14, 9, 17), (0, 0, 0), (0, 1, 0), 45)
        gs, osc = G.Scene(fs), O.OracleScene(fs)
        opts = G.render_opts(mode=L.MODE_ONE_RAY, recurs=3)
        tg, _, st = gs.render(cam, 96, 64, opts)
        to, _ = osc.render(cam, 96, 64, opts)
        assert np.abs(tg[..., :4] - to[..., :4]).max() <= RGB_TOL and np.array_equal(tg[..., 4], to[..., 4])


def test_more_than_64_lights_is_a_limit_not_a_fault():
    b = G.SceneBuilder()
    s = b.tex(b.sphere((0, 0, 0), 1), b.t_matte((1, 1, 1)))
    for i in range(65):
        b.light((i, 10, 0), (1, 1, 1))
    fs = b.flatten(s)
    with pytest.raises(L.GlomeError) as e:
        G.Scene(fs)
    assert e.value.code == L.ELIMIT


def test_many_csg_crossings():
    """280 rayint_advance re-issues on one ray (round 1 gave up after 48): same hit as the oracle, flagged."""
    b = G.SceneBuilder()
    slabs = [b.box((-1, -1, 2 * i + 0.5), (1, 1, 2 * i + 1.5)) for i in range(140)]
    body = b.box((-0.5, -0.5, 300), (0.5, 0.5, 310))
    root = b.tex(b.difference(body, b.group(slabs)), b.t_matte((0.5, 0.5, 0.5)))
    b.light((0, 50, 0), (100, 100, 100))
    fs = b.flatten(root)
    gs, osc = G.Scene(fs), O.OracleScene(fs)
    rays = np.array([[0.1, 0.2, -5.0, 1e-4, 2e-4, 1.0], [0.1, 0.2, 100.2, 1e-4, 2e-4, 1.0], [0.1, 0.2, 290.0, 1e-4, 2e-4, 1.0]])
    rays[:, 3:] /= np.linalg.norm(rays[:, 3:], axis=1)[:, None]
    o, g = osc.rayint(rays), gs.rayint(rays)
    assert o["hit"].all() and np.array_equal(g["hit"], o["hit"]) and np.array_equal(g["prim"], o["prim"])
    assert np.allclose(g["t"], o["t"], rtol=1e-12, atol=0)
    assert np.array_equal(g["pos"], o["pos"]) and np.array_equal(g["norm"], o["norm"])
    assert g["flags"][0] == L.HITFLAG_CSG_OVERFLOW and g["flags"][2] == 0


def mirrored_pairs_scene(npairs, seed):
    """Pairs of spheres (p, q, z) / (q, p, z) of one radius, each with its own colour, under one bih.  A ray whose origin and
    direction have x == y meets the two spheres of a pair at EXACTLY the same depth (the arithmetic is symmetric under
    the swap), so `nearest` has to decide: the hit met later by the reference's walk wins (Solid.hs:37-44, Bih.hs:350-366)."""
    b = G.SceneBuilder()
    rng = np.random.default_rng(seed)
    items = []
    for i in range(npairs):
        p, q = rng.uniform(-6, 6, 2)
        z, r = rng.uniform(-6, 6), rng.uniform(0.3, 0.9)
        for k, c in enumerate(((p, q, z), (q, p, z))):
            col = rng.uniform(0.1, 1.0, 3)
            items.append(b.tag(b.tex(b.sphere(c, r), b.t_matte(col)), 2 * i + k))
    b.light((3, 40, -30), (900, 900, 900))
    root = b.bih(items)
    cam = L.GlomeCamera()
    cam.pos[:] = (0.0, 0.0, -16.0)
    cam.fwd[:] = (0.0, 0.0, 1.0)
    cam.up[:] = (0.0, 0.5, 0.0)
    cam.right[:] = (0.5, 0.0, 0.0)
    return b, b.flatten(root), cam


@pytest.mark.parametrize("size", [24, 64, 512])
def test_tie_rule_of_nearest_under_donation(size):
    """Square frames: the rays of the diagonal pixels have d.x == d.y, so every mirrored pair they meet is an exact tie.
    Small frames are all drain phase (every ray is split among the lanes of its warp), the large one mixes both.  Ten
    renders each: the winner must be the oracle's every time."""
    b, fs, cam = mirrored_pairs_scene(300, 5)
    assert fs.scene_class == L.CLASS_FLAT
    gs, osc = G.Scene(fs), O.OracleScene(fs)
    # the ties are real: on the diagonal, the two spheres of a pair give the same depth bit for bit
    idx = np.arange(size)
    diag = G.camera_rays(cam, size, size, idx, idx)
    assert np.array_equal(diag[:, 3], diag[:, 4])
    opts = G.render_opts(mode=L.MODE_ONE_RAY, recurs=3)
    to, _ = osc.render(cam, size, size, opts)
    ho = osc.rayint(diag)
    assert ho["hit"].sum() > size // 4
    for rep in range(10):
        tg, _, st = gs.render(cam, size, size, opts)
        assert np.abs(tg[..., :4] - to[..., :4]).max() <= RGB_TOL, rep
        assert np.array_equal(tg[..., 4], to[..., 4])
    same_hits(gs.rayint(diag), ho)


def mirrored_mesh_scene(ntri, seed):
    """A Mesh of random triangles, each followed somewhere in the list by its mirror image under x <-> y (own texture and
    tag).  For a ray with o.x == o.y and d.x == d.y the two intersection computations are the same arithmetic with the x and
    y terms swapped (the cross products change sign as a whole, the three-term dot products only commute), so a triangle
    and its mirror image are met at EXACTLY the same depth -- in different leaves of the BVH, so that which one the
    reference's walk meets later (and `nearest` therefore keeps: Mesh.hs:172-176, Solid.hs:37-44) depends on the walk."""
    b = G.SceneBuilder()
    rng = np.random.default_rng(seed)
    c = rng.uniform(-6, 6, (ntri, 1, 3))
    tri = c + rng.uniform(-1.3, 1.3, (ntri, 3, 3))
    mir = tri[:, :, [1, 0, 2]]
    verts = np.concatenate([tri.reshape(-1, 3), mir.reshape(-1, 3)])
    norms = np.tile(np.array([[0.0, 0.0, -1.0]]), (len(verts), 1))
    order = rng.permutation(2 * ntri)
    tris = []
    for k in order:
        v = 3 * int(k)
        tris.append((v, v + 1, v + 2, v, v + 1, v + 2, (k % 4) + (4 if k >= ntri else 0), int(k >= ntri)))
    texs = [b.t_matte(col) for col in rng.uniform(0.1, 1.0, (8, 3))]
    mesh = b.mesh(verts, norms, np.array(tris, np.int32), texs, [500, 501])
    b.light((3, 40, -30), (900, 900, 900))
    root = b.group([mesh])
    cam = L.GlomeCamera()
    cam.pos[:] = (0.0, 0.0, -16.0)
    cam.fwd[:] = (0.0, 0.0, 1.0)
    cam.up[:] = (0.0, 0.5, 0.0)
    cam.right[:] = (0.5, 0.0, 0.0)
    return b, b.flatten(root), cam


@pytest.mark.parametrize("size", [24, 64, 384])
def test_tie_rule_of_nearest_under_donation_in_a_mesh(size):
    """k_bvh_closest's drain phase splits a ray's pending subtrees among the lanes of its warp.  The rays of the diagonal
    pixels of a square frame meet every mirrored pair at exactly the same depth; small frames are all drain phase, the large
    one mixes both.  Ten renders each: the winner must be the oracle's every time."""
    b, fs, cam = mirrored_mesh_scene(700, 11)
    assert fs.scene_class == L.CLASS_FLAT
    gs, osc = G.Scene(fs), O.OracleScene(fs)
    idx = np.arange(size)
    diag = G.camera_rays(cam, size, size, idx, idx)
    assert np.array_equal(diag[:, 3], diag[:, 4])
    ho = osc.rayint(diag)
    assert ho["hit"].sum() > size // 4
    assert len(np.unique(ho["tag"][ho["hit"] != 0, 0])) == 2   # both a triangle and a mirror image win somewhere
    opts = G.render_opts(mode=L.MODE_ONE_RAY, recurs=2)
    to, _ = osc.render(cam, size, size, opts)
    for rep in range(10):
        tg, _, st = gs.render(cam, size, size, opts)
        assert np.abs(tg[..., :4] - to[..., :4]).max() <= RGB_TOL, rep
        assert np.array_equal(tg[..., 4], to[..., 4])
    same_hits(gs.rayint(diag), ho)
    aa = G.render_opts(mode=L.MODE_ADAPTIVE_AA, recurs=2, blocksize=33)
    to, _ = osc.render(cam, size, size, aa)
    for rep in range(3):
        tg, _, st = gs.render(cam, size, size, aa)
        assert np.abs(tg[..., :4] - to[..., :4]).max() <= RGB_TOL, rep


def test_render_and_render_dev_buffers_at_growing_sizes():
    """glome_render and glome_render_dev size their work buffers separately: small host frame, then a large
    device frame with adaptive AA, then a large host frame."""
    import torch
    b, fs, cam = SG.random_scene(4)
    gs, osc = G.Scene(fs), O.OracleScene(fs)
    opts = G.render_opts(mode=L.MODE_ADAPTIVE_AA, recurs=3)
    small, _, _ = gs.render(cam, 40, 30, opts, want_rgb8=True)
    w, h = 200, 130
    dev = torch.zeros((h, w, 5), dtype=torch.float64, device="cuda")
    gs.render_ptr(cam, w, h, opts, dev.data_ptr(), 0, dev=True, stream=torch.cuda.current_stream().cuda_stream)
    torch.cuda.synchronize()
    big, rgb, _ = gs.render(cam, w, h, opts, want_rgb8=True)
    to, ro = osc.render(cam, w, h, opts, want_rgb8=True)
    assert np.abs(big[..., :4] - to[..., :4]).max() <= RGB_TOL
    assert np.array_equal(dev.cpu().numpy(), big)
    assert (rgb != ro).mean() < 1e-3


def test_recurs_above_the_tracer_limit_is_refused():
    b, fs, cam = SG.random_scene(0)
    gs = G.Scene(fs)
    with pytest.raises(L.GlomeError) as e:
        gs.render(cam, 16, 16, G.render_opts(mode=L.MODE_ONE_RAY, recurs=9))
    assert e.value.code == L.ELIMIT


@pytest.mark.parametrize("nested", [False, True])
def test_large_plain_group_keeps_the_list_folds_ties_on_the_device(nested):
    """The implicit BIH the device walks for a large plain `group` (glome_tagmap.h) must give the list fold's result,
    ties included (shared box faces, duplicated spheres, Void, instanced items; tests/test_gen_host.py builds the scene),
    in FP64 bit for bit; the FP32 twin walks the same structure."""
    from test_gen_host import _tie_group_scene, _tie_rays
    b, fs = _tie_group_scene(nested)
    gs, osc = G.Scene(fs), O.OracleScene(fs)
    rays = _tie_rays(5)
    same_hits(gs.rayint(rays), osc.rayint(rays))
    tmax = np.random.default_rng(6).uniform(1.0, 8.0, len(rays))
    same_hits(gs.rayint(rays, tmax), osc.rayint(rays, tmax))
    assert np.array_equal(gs.shadow(rays, 6.0), osc.shadow(rays, 6.0))
    ro, do = osc.trace(rays, recurs=3)
    rg, dg = gs.trace(rays, recurs=3)
    assert np.array_equal(dg, do) and np.abs(rg - ro).max() <= RGB_TOL
    assert np.array_equal(gs.debug_count(rays), osc.debug_count(rays))
    g32 = G.Scene(fs, precision=32).rayint(rays)
    o = osc.rayint(rays)
    assert ((g32["hit"] == o["hit"]) & (g32["prim"] == o["prim"])).mean() > 0.97
