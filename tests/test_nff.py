"""NFF / SPD reader (glome_b200/csrc/nff.cpp restates Spd.hs:1-261).  The reference has no tests or sample files
for it, so the checks are: (1) a scene text yields exactly the FlatScene that the same constructor calls yield, in
the order Spd.hs's cons-accumulation implies (groups and lights reversed, last camera / background win);
(2) the reader's quirks; (3) on a GPU, the loaded scene renders like the oracle renders the same FlatScene."""
import hashlib

import numpy as np
import pytest

import glome_b200 as G
from glome_b200 import _lib as L

T_SPHERE, T_TRIANGLE = 1, 2  # GlomeNodeType (include/glome_cuda.h)


def balls_nff(depth=1):
    """a small sphereflake in the shape SPD's `balls` writes (v / b / l / f / s lines, one polygon floor)"""
    out = ["v", "from 2.1 1.3 1.7", "at 0 0 0", "up 0 0 1", "angle 45", "hither 1", "resolution 512 512",
           "b 0.078 0.361 0.753", "l 4 3 2", "l 1 -4 4 0.9 0.8 0.7", "l -3 1 5",
           "f 1 0.75 0.33 0.8 0 100000 0 0", "p 4", "12 12 -0.5", "-12 12 -0.5", "-12 -12 -0.5", "12 -12 -0.5",
           "f 1 0.9 0.7 0.5 0.5 3 0 0"]
    rng = np.random.default_rng(5)

    def flake(c, r, d):
        out.append("s %.10g %.10g %.10g %.10g" % (c[0], c[1], c[2], r))
        if d > 0:
            for _ in range(9):
                v = rng.normal(size=3); v /= np.linalg.norm(v)
                flake(c + v * (r + r / 3), r / 3, d - 1)
    flake(np.zeros(3), 0.5, depth)
    out += ["f 0.2 0.4 0.9 0.7 0.3 20 0.25 1.5", "c", "0 0 0.9 0.2", "0 0 1.6 0.05",
            "pp 3", "1 1 0 0 0 1", "1.5 1 0 0 0.1 1", "1 1.5 0.2 0.1 0 1"]
    return "\n".join(out) + "\n"


def flat_sig(fs):
    fv = G.FlatView(fs)
    return hashlib.sha1(fv.nodes.tobytes() + fv.bihnodes.tobytes() + fv.ipool.tobytes() + fv.dpool.tobytes()).hexdigest()


def test_nff_scene_equals_explicit_construction():
    text = balls_nff(1)
    b = G.SceneBuilder()
    root, cam, bg, used = b.load_nff(text)
    assert used == len(text.rstrip("\n")) or used == len(text)
    assert bg == (0.078, 0.361, 0.753)
    ref = G.camera((2.1, 1.3, 1.7), (0, 0, 0), (0, 0, 1), 45)
    assert bytes(cam) == bytes(ref)
    fs = b.flatten(root)
    assert fs.n_lights == 3
    # the same scene through the constructors, in Spd.hs's order: groups reversed, lights reversed
    e = G.SceneBuilder()
    lines = [ln.split() for ln in text.splitlines()]
    m1 = e.mat_surface((1, 0.75, 0.33), alpha=1 - 0, amb=0, kd=0.8, ks=0, shine=100000)
    t1 = e.tex_uniform(m1)
    quad = [(12, 12, -0.5), (-12, 12, -0.5), (-12, -12, -0.5), (12, -12, -0.5)]
    g1 = e.tex(e.bih([e.group([e.triangle(quad[0], quad[1], quad[2]), e.triangle(quad[0], quad[2], quad[3])])]), t1)
    m2 = e.mat_surface((1, 0.9, 0.7), alpha=1, amb=0, kd=0.5, ks=0.5, shine=3)
    t2 = e.tex_uniform(m2)
    sph = [ln for ln in lines if ln and ln[0] == "s"]
    g2 = e.tex(e.bih([e.sphere([float(x) for x in s[1:4]], float(s[4])) for s in sph]), t2)
    m3 = e.mat_surface((0.2, 0.4, 0.9), alpha=1 - 0.25, amb=0, kd=0.7, ks=0.3, shine=20)
    t3 = e.tex_uniform(m3)
    cone = e.cone((0, 0, 0.9), 0.2, (0, 0, 1.6), 0.05)
    patch = e.group([e.trianglenorm((1, 1, 0), (1.5, 1, 0), (1, 1.5, 0.2), (0, 0, 1), (0, 0.1, 1), (0.1, 0, 1))])
    g3 = e.tex(e.bih([cone, patch]), t3)
    e.light((-3, 1, 5), (1, 1, 1)); e.light((1, -4, 4), (0.9, 0.8, 0.7)); e.light((4, 3, 2), (1, 1, 1))
    eroot = e.bih([g3, g2, g1])
    efs = e.flatten(eroot)
    assert flat_sig(fs) == flat_sig(efs)
    import ctypes as C
    assert C.string_at(fs.lights, 64 * 3) == C.string_at(efs.lights, 64 * 3)
    assert C.string_at(fs.materials, 96 * fs.n_materials) == C.string_at(efs.materials, 96 * efs.n_materials)


def test_nff_quirks():
    head = "v\nfrom 0 0 5\nat 0 0 0\nup 0 1 0\nangle 40\nhither 1\nresolution 64 64\nb 0 0 0\n"
    b = G.SceneBuilder()
    # a '#' comment swallows the rest of the input (Spd.hs:12-29): the sphere group after it is never read
    root, cam, bg, used = b.load_nff(head + "f 1 1 1 1 0 0 0 0\ns 0 0 0 1\n# comment\nf 1 0 0 1 0 0 0 0\ns 3 0 0 1\n")
    fv = G.FlatView(b.flatten(root))
    assert (fv.nodes["type"] == T_SPHERE).sum() == 1
    # primitives before any fill end the parse; no camera / no background is an error
    b2 = G.SceneBuilder()
    root2, _, _, used2 = b2.load_nff(head + "s 0 0 0 1\nf 1 1 1 1 0 0 0 0\ns 1 0 0 1\n")
    assert used2 <= len(head) and (G.FlatView(b2.flatten(root2)).nodes["type"] == T_SPHERE).sum() == 0
    with pytest.raises(L.GlomeError):
        G.SceneBuilder().load_nff("b 0 0 0\nf 1 1 1 1 0 0 0 0\ns 0 0 0 1\n")
    with pytest.raises(L.GlomeError):
        G.SceneBuilder().load_nff(head.replace("b 0 0 0\n", "") + "f 1 1 1 1 0 0 0 0\ns 0 0 0 1\n")
    # the last camera and the last background win; "p" ignores its count and fans the vertices it finds
    b3 = G.SceneBuilder()
    root3, cam3, bg3, _ = b3.load_nff(head + "b 1 0 0\n" + head.replace("from 0 0 5", "from 0 0 9").replace("b 0 0 0\n", "") +
                                      "f 1 1 1 1 0 0 0 0\np 3\n0 0 0\n1 0 0\n1 1 0\n0 1 0\n-1 1 0\n")
    assert bg3 == (1.0, 0.0, 0.0) and bytes(cam3) == bytes(G.camera((0, 0, 9), (0, 0, 0), (0, 1, 0), 40))
    assert (G.FlatView(b3.flatten(root3)).nodes["type"] == T_TRIANGLE).sum() == 3


@pytest.mark.gpu
def test_nff_scene_renders_like_the_oracle():
    import oracle as O
    b = G.SceneBuilder()
    root, cam, bg, _ = b.load_nff(balls_nff(2))
    fs = b.flatten(root)
    gs, osc = G.Scene(fs, 0), O.OracleScene(fs)
    for mode in (L.MODE_ONE_RAY, L.MODE_ADAPTIVE_AA):
        opts = G.render_opts(mode=mode, recurs=3)
        tg, _, st = gs.render(cam, 200, 150, opts)
        to, _ = osc.render(cam, 200, 150, opts)
        assert np.abs(tg[..., :4] - to[..., :4]).max() <= 1e-6
        assert np.array_equal(tg[..., 4], to[..., 4])
        assert st.overflow_rays == 0
