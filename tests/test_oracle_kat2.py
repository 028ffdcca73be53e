"""More known-answer tests for the oracle AND the scene-graph machine (round 2; VERDICT r1 "pin the oracle harder").

Each expected value is worked out below from the Haskell formula it cites, independently of oracle/glome_oracle.cpp
(by hand where the geometry has a closed form, by a few lines of Python restating the .hs text where it does not).
Every KAT is applied to the oracle and to the host-compiled machine of glome_b200/csrc/glome_gen.cuh (tests/genhost.py),
so a misreading shared by both would have to be shared with this file too."""
import ctypes as C
import math

import numpy as np
import pytest

import glome_b200 as G
import glome_b200.scene as SC
import genhost as H
import oracle as O

INF = 1000000.0
S5 = math.sqrt(5.0)


def both(scene_fn):
    b = G.SceneBuilder()
    root = scene_fn(b)
    fs = b.flatten(root)
    return b, fs, (O.OracleScene(fs), H.HostGenScene(fs))


def ray(o, d):
    return np.array([[*o, *d]], dtype=np.float64)


def close(a, b, tol=1e-12):
    return np.allclose(np.asarray(a, float), np.asarray(b, float), rtol=0, atol=tol)


# ---- Cone.hs ----------------------------------------------------------------------------------
def cone_scene(b):
    # Cone r clip1 clip2 height (Cone.hs:23): radius r at z = clip1 ... 0 at z = height
    return b.cone_z(1.0, 0.0, 1.0, 2.0)


def test_cone_side_hit_and_normal():
    """Cone.hs:155-190.  r = 1, height = 2: the radius at height z is 1 - z/2.  The ray (5,0,0.5) + t(-1,0,0) meets the
    side at x = 0.75, t = 4.25.  Normal: invhyp = 1/sqrt(height^2 + r^2) = 1/sqrt 5, up = r invhyp, out = height invhyp,
    n = (x out/r_, y out/r_, up) with r_ = sqrt(x^2 + y^2) = 0.75  =>  n = (2/sqrt 5, 0, 1/sqrt 5)."""
    b, fs, scs = both(cone_scene)
    for sc in scs:
        h = sc.rayint(ray((5, 0, 0.5), (-1, 0, 0)))[0]
        assert h["hit"] == 1 and close(h["t"], 4.25) and close(h["pos"], (0.75, 0, 0.5))
        assert close(h["norm"], (2 / S5, 0, 1 / S5))


def test_cone_caps():
    """Cone.hs:191-200.  A ray along +z at x = 0.2 meets the (double) cone at z = 1.6, outside (clip1, clip2) = (0, 1); with
    dz > 0 and oz < clip1 only the bottom disc (z = clip1, radius r, normal (0,0,-1)) is tried: t = 3.  Coming down from
    z = 5 (dz < 0, oz > clip2) only the top disc: radius r (1 - (clip2 - clip1)/height) = 0.5, normal (0,0,1), t = 4.  At
    x = 0.7 the top disc is missed (0.7 > 0.5) and the other root is never tried: a miss, although the side is there."""
    b, fs, scs = both(cone_scene)
    for sc in scs:
        h = sc.rayint(ray((0.2, 0, -3), (0, 0, 1)))[0]
        assert h["hit"] == 1 and close(h["t"], 3.0) and close(h["norm"], (0, 0, -1)) and close(h["pos"], (0.2, 0, 0))
        h = sc.rayint(ray((0.2, 0, 5), (0, 0, -1)))[0]
        assert h["hit"] == 1 and close(h["t"], 4.0) and close(h["norm"], (0, 0, 1)) and close(h["pos"], (0.2, 0, 1))
        # x = 0.7: the quadratic's near root is the mirrored cone above the apex (z = 3.4), outside the clip range
        assert sc.rayint(ray((0.7, 0, 5), (0, 0, -1)))[0]["hit"] == 0
        assert sc.shadow(ray((0.2, 0, -3), (0, 0, 1)), 10.0)[0] == 1 and sc.shadow(ray((0.2, 0, -3), (0, 0, 1)), 2.5)[0] == 0


def test_cylinder_side_normal():
    """Cone.hs:104-128: Cylinder r h1 h2 around z; the normal of a side hit is (x/r, y/r, 0).  r = 2: the ray (5,0,0) +
    t(-1,0,0) hits at t = 3, n = (1,0,0); from (0, 5, 0.5) along (0,-1,0): t = 3, n = (0,1,0)."""
    b, fs, scs = both(lambda b: b.cylinder_z(2.0, -1.0, 1.0))
    for sc in scs:
        h = sc.rayint(ray((5, 0, 0), (-1, 0, 0)))[0]
        assert h["hit"] == 1 and close(h["t"], 3.0) and close(h["norm"], (1, 0, 0))
        h = sc.rayint(ray((0, 5, 0.5), (0, -1, 0)))[0]
        assert h["hit"] == 1 and close(h["t"], 3.0) and close(h["norm"], (0, 1, 0)) and close(h["pos"], (0, 2, 0.5))


# ---- Solid.hs: Instance -------------------------------------------------------------------------
def test_instance_with_non_uniform_scale():
    """Solid.hs:388-403.  Unit sphere under scale (2,1,1).  Ray (5,0,0) + t(-1,0,0): object space origin (2.5,0,0), direction
    (-0.5,0,0), lenscale 0.5; the unit-length object ray hits at 1.5, world depth 1.5 * (1/0.5) = 3; pos = fwd (1,0,0) =
    (2,0,0).  Ray (sqrt2, 5, 0) + t(0,-1,0): object origin (sqrt2/2, 5, 0), hit at y = sqrt2/2, t = 5 - sqrt2/2; object
    normal (sqrt2/2, sqrt2/2, 0); world normal = vnorm (inverse-transpose n) = vnorm (sqrt2/4, sqrt2/2, 0) = (1, 2, 0)/sqrt 5:
    NOT the direction from the centre to the hit point."""
    r2 = math.sqrt(2.0)
    b, fs, scs = both(lambda b: b.transform(b.sphere((0, 0, 0), 1.0), [G.scale((2, 1, 1))]))
    for sc in scs:
        h = sc.rayint(ray((5, 0, 0), (-1, 0, 0)))[0]
        assert h["hit"] == 1 and close(h["t"], 3.0) and close(h["pos"], (2, 0, 0)) and close(h["norm"], (1, 0, 0))
        h = sc.rayint(ray((r2, 5, 0), (0, -1, 0)))[0]
        assert h["hit"] == 1 and close(h["t"], 5 - r2 / 2) and close(h["pos"], (r2, r2 / 2, 0))
        assert close(h["norm"], (1 / S5, 2 / S5, 0))
        # the distance limit is scaled into object space too (d * lenscale): 2.9 is short of the hit, 3.1 reaches it
        assert sc.rayint(ray((5, 0, 0), (-1, 0, 0)), 2.9)[0]["hit"] == 0 and sc.rayint(ray((5, 0, 0), (-1, 0, 0)), 3.1)[0]["hit"] == 1
        assert sc.inside(np.array([[1.9, 0, 0]]))[0] == 1 and sc.inside(np.array([[0, 1.1, 0]]))[0] == 0


# ---- Shader.hs ----------------------------------------------------------------------------------
def test_refract_total_internal_reflection_branch():
    """Shader.hs:120-155 with recurs = 1: both secondary traces run with recurs 0 and are transparent misses (Trace.hs:60),
    so the result is b * refr, where b = ca_black = (0,0,0,1) ONLY in the cs2 < 0 branch (Shader.hs:139-141).  Refract 0.35 0.8
    1.5 on a unit sphere, eta = ior = 1.5 on the way in (n . eyedir > 0): cs2 = 1 - eta^2 (1 - c1^2).  A ray hitting where
    c1 = dir . n = -0.5 gives cs2 = 1 - 2.25 * 0.75 = -0.6875 < 0  =>  (0, 0, 0, 0.8).  A central ray (c1 = -1, cs2 = 1)
    refracts, its trace is a miss: (0, 0, 0, 0)."""
    def scene(b):
        return b.tex(b.sphere((0, 0, 0), 1.0), b.tex_uniform(b.mat_refract(0.35, 0.8, 1.5)))
    b, fs, scs = both(scene)
    x = math.sqrt(3.0) / 2  # the normal at (x, 0, -0.5) makes 60 degrees with -dir
    for sc in scs:
        rgba = sc.trace(ray((x, 0, -5), (0, 0, 1)), recurs=1)[0][0]
        assert close(rgba, (0, 0, 0, 0.8))
        assert close(sc.trace(ray((0, 0, -5), (0, 0, 1)), recurs=1)[0][0], (0, 0, 0, 0))
        assert close(sc.trace(ray((0, 0, -5), (0, 0, 1)), recurs=0)[0][0], (0, 0, 0, 0))


def flat(b, rgb, alpha=1.0):
    """a Surface that shows its own colour whatever the lights do: ambient 1, kd = ks = 0 (Shader.hs:90-105)"""
    return b.mat_surface(rgb, alpha, 1.0, 0.0, 0.0, 0.0)


def test_warp_nearer_wins():
    """Shader.hs:157-175: Warp frame scene' lights' xfm traces the ORIGINAL ray against `frame` and the transformed ray
    (from the hit point) against `scene'`, the latter limited to the frame's depth, and keeps the nearer of the two.
    Portal pane at z = 0, a red frame object 7 behind the ray's origin plane (depth 7 from the ray origin at z = -2:
    a box face at z = 5), the warped scene a green box whose face is 3 (resp. 9) beyond the pane: green, resp. red."""
    def scene(green_z):
        def f(b):
            frame = b.tex(b.box((-1, -1, 5), (1, 1, 6)), b.tex_uniform(flat(b, (1, 0, 0))))
            world = b.tex(b.box((-1, -1, green_z), (1, 1, green_z + 1)), b.tex_uniform(flat(b, (0, 1, 0))))
            ident = G.translate((0, 0, 0))
            warp = b.mat_warp(frame, world, 0, ident)
            return b.tex(b.box((-1, -1, -0.01), (1, 1, 0.01)), b.tex_uniform(warp))
        return f
    for green_z, want in ((3.0, (0, 1, 0, 1)), (9.0, (1, 0, 0, 1))):
        b, fs, scs = both(scene(green_z))
        d = np.array([0.01, 0.02, 1.0])  # (a +0.0 direction component would miss every box: SURVEY Appendix A3)
        d /= np.linalg.norm(d)
        for sc in scs:
            assert close(sc.trace(ray((0.1, 0.2, -2), d), recurs=3)[0][0], want)


def test_additive_layers_casum():
    """Shader.hs:177-179, Clr.hs:93-103: casum sums r*a per layer and gives alpha 1 - prod (1 - clamp a).  Layers (0.2,0.4,0.6)
    alpha 0.5 and (1,0.5,0.25) alpha 0.25: rgb = (0.1+0.25, 0.2+0.125, 0.3+0.0625), alpha = 1 - 0.5*0.75 = 0.625; trace then
    folds that over transparent black (cafold, Clr.hs:106): rgb * alpha, alpha."""
    def scene(b):
        m = b.mat_additive([flat(b, (0.2, 0.4, 0.6), 0.5), flat(b, (1.0, 0.5, 0.25), 0.25)])
        return b.tex(b.sphere((0, 0, 0), 1.0), b.tex_uniform(m))
    b, fs, scs = both(scene)
    a = 0.625
    want = (0.35 * a, 0.325 * a, 0.3625 * a, a)
    for sc in scs:
        assert close(sc.trace(ray((0, 0, -5), (0, 0, 1)), recurs=3)[0][0], want)


def test_blend_weights():
    """Shader.hs:181-184, Clr.hs:87: caweight a b w = a w + b (1 - w), per channel incl. alpha."""
    def scene(b):
        m = b.mat_blend(flat(b, (1, 0, 0)), flat(b, (0, 0, 1)), 0.25)
        return b.tex(b.sphere((0, 0, 0), 1.0), b.tex_uniform(m))
    b, fs, scs = both(scene)
    for sc in scs:
        assert close(sc.trace(ray((0, 0, -5), (0, 0, 1)), recurs=3)[0][0], (0.25, 0, 0.75, 1))


# ---- Texture.hs ---------------------------------------------------------------------------------
PHI = [3, 0, 2, 7, 4, 1, 5, 11, 8, 10, 9, 6]                                    # Texture.hs:57
GRAD = [v for v in ((x, y, z) for x in (-1, 0, 1) for y in (-1, 0, 1) for z in (-1, 0, 1))
        if 1.1 < math.sqrt(v[0] ** 2 + v[1] ** 2 + v[2] ** 2) < 1.5]            # Texture.hs:60-64


def hs_omega(t_):                                                               # Texture.hs:49-54
    t = -t_ if t_ < 0 else t_
    tsqr = t * t
    tcube = tsqr * t
    return (-6) * tcube * tsqr + 15 * tcube * t - 10 * tcube + 1


def hs_knot(i, j, k, v):                                                        # Texture.hs:66-77
    a = PHI[abs(k) % 12]
    b = PHI[abs(j + a) % 12]
    c = PHI[abs(i + b) % 12]
    g = GRAD[c]
    return hs_omega(v[0]) * hs_omega(v[1]) * hs_omega(v[2]) * ((g[0] * v[0]) + (g[1] * v[1]) + (g[2] * v[2]))


def hs_perlin(p):                                                               # Texture.hs:92-116
    x, y, z = p
    i, j, k = math.floor(x), math.floor(y), math.floor(z)
    u, v, w = x - i, y - j, z - k
    n = (hs_knot(i, j, k, (u, v, w)) + hs_knot(i + 1, j, k, (u - 1, v, w)) + hs_knot(i, j + 1, k, (u, v - 1, w)) +
         hs_knot(i, j, k + 1, (u, v, w - 1)) + hs_knot(i + 1, j + 1, k, (u - 1, v - 1, w)) +
         hs_knot(i + 1, j, k + 1, (u - 1, v, w - 1)) + hs_knot(i, j + 1, k + 1, (u, v - 1, w - 1)) +
         hs_knot(i + 1, j + 1, k + 1, (u - 1, v - 1, w - 1)))
    return (n + 1) * 0.5


def test_perlin_values():
    """Three perlin values restated from Texture.hs:49-116 in Python (same operation order, IEEE doubles, no FMA): the
    oracle must give the same bits.  The gradient table is rebuilt from the list comprehension of Texture.hs:60-64."""
    assert len(GRAD) == 12 and GRAD[0] == (-1, -1, 0) and GRAD[11] == (1, 1, 0)
    lib = O.load()
    for p in ((0.5, 0.5, 0.5), (1.25, 2.75, -0.5), (-3.6, 0.2, 7.9), (12.3, -4.4, 0.01)):
        q = np.array(p, dtype=np.float64)
        got = lib.orc_perlin(q.ctypes.data_as(C.c_void_p))
        assert got == hs_perlin(p), (p, got, hs_perlin(p))
    # and through a perlin-blend texture on both engines: Blend a b (perlin (pos * 3)) (TestScene.hs:214-220)
    def scene(b):
        t = b.tex_perlin_blend(flat(b, (1, 1, 1)), flat(b, (0, 0, 0)), 3.0)
        return b.tex(b.box((-1, -1, 0), (1, 1, 1)), t)
    b, fs, scs = both(scene)
    w = hs_perlin((0.3 * 3, 0.4 * 3, 0.0 * 3))
    for sc in scs:
        rgba = sc.trace(ray((0.3, 0.4, -2), (-0.0, -0.0, 1.0)), recurs=3)[0][0]  # (-0.0: a proper slab; +0.0 would miss the box)
        assert close(rgba, (w, w, w, 1.0), 1e-15)


# ---- Bih.hs: the builder's selection slip -------------------------------------------------------------
SLIP_BOXES = np.array([[1.4, 3.4, -2.1, 5.8, 3.6, 1.7], [1.3, 3.1, -3.6, 5.1, 3.9, -0.6], [1.0, 0.4, 0.1, 2.2, 3.2, 1.7],
                       [-1.0, -1.4, 0.8, 3.4, 3.0, 5.4], [0.2, -1.0, -1.5, 4.8, 3.8, -0.7]])


def test_bih_builder_costy_costb_slip():
    """Bih.hs:279-285 chooses x if it is strictly cheapest, else y if costy < costz && costy < costb, else z if
    `costy < costb` (sic: costz was meant), else big/small.  For these five boxes (box = bounding box of the list):
    costx = 1330.08, costy = 1507.44, costz = 1101.87 (the cheapest), costb = 1492.18, costorig = 1449.4: x is not the
    cheapest, y is not below z, and costy < costb is FALSE, so the reference splits big/small -- on the x planes of that
    partition (Bih.hs:231-232, 285): axis 0, lsplit = max x2 of the big boxes + delta = 3.4 + 1e-4, rsplit = min x1 of the
    small ones - delta = 0.2 - 1e-4 -- although the z split would have been cheaper."""
    for build in (O.bih_build, SC.bih_build):
        t = build(SLIP_BOXES)
        assert t["root"] == 0
        n = t["nodes"][0]
        assert n["axis"] == 0 and n["lsplit"] == 3.4 + 0.0001 and n["rsplit"] == 0.2 - 0.0001
    # the partition: "big" = surface area > 0.4 * the list's (Bih.hs:223): only box 3 (bbsa 90.88 of 289.88 * 0.4 = 115.95)? no:
    sa = lambda b: max(0.0, 2 * ((b[3] - b[0]) * (b[4] - b[1]) + (b[3] - b[0]) * (b[5] - b[2]) + (b[4] - b[1]) * (b[5] - b[2])))
    bb = np.concatenate([SLIP_BOXES[:, :3].min(0), SLIP_BOXES[:, 3:].max(0)])
    big = [i for i in range(5) if sa(SLIP_BOXES[i]) > sa(bb) * 0.4]
    assert max(SLIP_BOXES[i][3] for i in big) == 3.4 and min(SLIP_BOXES[i][0] for i in range(5) if i not in big) == 0.2


# ---- System.Random (TestScene's oak) ------------------------------------------------------------------
def test_stdgen_known_answers():
    """The mixing function of splitmix (Stafford's variant 13) is Vigna's splitmix64, whose published outputs pin it:
    state 0 -> 0xe220a8397b1dcdaf, state 1234567 -> 6457827717110365317.  `split` hands the first generator the advanced
    seed and the old gamma, and seeds the second one with mix64 of the intermediate seed (splitSMGen): its seed equals the
    parent's next Word64.  randomR (0, 0.5) stays in range."""
    from glome_b200 import _lib as L
    lib = L.load()
    out = (C.c_uint64 * 9)()
    d = (C.c_double * 3)()
    v = C.c_uint64()
    lib.glome_stdgen_probe(0, out, d, C.byref(v))
    assert v.value == 0xe220a8397b1dcdaf
    lib.glome_stdgen_probe(1234567, out, d, C.byref(v))
    assert v.value == 6457827717110365317
    lib.glome_stdgen_probe(42, out, d, C.byref(v))
    seed, gamma = out[0], out[1]
    assert gamma & 1 == 1                                    # mixGamma: always odd
    assert out[5] == (seed + 2 * gamma) % 2 ** 64 and out[6] == gamma    # left half of split
    assert out[7] == out[2]                                  # right half's seed = mix64 (seed + gamma) = the next output
    assert all(0.0 <= x <= 0.5 for x in d)
