"""getTags' (Glome.hs:410-414): clicking a pixel returns the tag list of the trace result, `ts ++ tags`
(Trace.hs:82): tags gathered by Reflect / Refract / Warp recursion, then the hit's own tag stack."""
import numpy as np
import pytest

import glome_b200 as G
from glome_b200 import _lib as L
import oracle as O

pytestmark = pytest.mark.gpu


def test_get_tags_matches_the_oracle_on_testscene():
    b = G.SceneBuilder()
    root, cam, rec = b.config_scene(1, 0)
    fs = b.flatten(root)
    gs, osc = G.Scene(fs, 0), O.OracleScene(fs)
    w, h = 180, 120
    xs, ys = np.meshgrid(np.arange(3, w, 7), np.arange(2, h, 5))
    xs, ys = xs.ravel(), ys.ravel()
    rays = G.camera_rays(cam, w, h, xs, ys)
    _, _, ohits, otags = osc.trace(rays, recurs=rec, want_hits=True, want_tags=True)
    n_tagged = n_gathered = 0
    for i, (x, y) in enumerate(zip(xs, ys)):
        tags, truncated, hit = gs.get_tags(cam, w, h, int(x), int(y), rec)
        ref = list(otags[i, 1:1 + min(otags[i, 0], 16)])
        assert hit.hit == ohits["hit"][i] and hit.prim == ohits["prim"][i]
        assert tags == ref and not truncated
        n_tagged += bool(tags)
        n_gathered += len(tags) > hit.ntag  # tags that came back from a reflected / refracted / warped trace
    assert n_tagged > 20 and n_gathered > 0  # the scene has tagged objects and mirror / warp surfaces


@pytest.mark.parametrize("config,n", [(2, 30000), (1, 0), (3, 20000)])
def test_rayint_debug_count_and_heat_map(config, n):
    """rayint_debug's box count (Bih.hs:378-412 and the instances that forward or sum it) per camera ray, and the
    frame get_color_debug makes of it (Glome.hs:57-60), against the oracle's restatement."""
    b = G.SceneBuilder()
    root, cam, rec = b.config_scene(config, n)
    fs = b.flatten(root)
    gs, osc = G.Scene(fs, 0), O.OracleScene(fs)
    w, h = 96, 64
    ys, xs = np.mgrid[0:h, 0:w]
    rays = G.camera_rays(cam, w, h, xs.ravel(), ys.ravel())
    g, o = gs.debug_count(rays), osc.debug_count(rays)
    assert np.array_equal(g, o)
    assert g.max() > 0
    short = osc.debug_count(rays, 30.0)
    assert np.array_equal(gs.debug_count(rays, 30.0), short)
    opts = G.render_opts(mode=L.MODE_ONE_RAY, recurs=rec, debug_heatmap=1)
    tg, _, _ = gs.render(cam, w, h, opts)
    to, _ = osc.render(cam, w, h, opts)
    assert np.abs(tg[..., :4] - to[..., :4]).max() <= 1e-6
    plain, _, _ = gs.render(cam, w, h, G.render_opts(mode=L.MODE_ONE_RAY, recurs=rec))
    cnt = g.reshape(h, w)
    assert np.array_equal(tg[..., 1], plain[..., 1] + cnt / 1000.0)
    with pytest.raises(L.GlomeError):  # the heat map is per-pixel get_color_debug: not defined for the AA pipeline
        gs.render(cam, w, h, G.render_opts(mode=L.MODE_ADAPTIVE_AA, recurs=rec, debug_heatmap=1))
