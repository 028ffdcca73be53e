"""floor_probe.py -- developer harness: how long is the longest ray of the Mesh scene, and what does each traversal launch of a
1/N share of configs[4] cost when it runs alone (GLOME_OPT_SEG_CONCURRENT 0)?"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import torch
import glome_b200 as G
from glome_b200 import _lib as L
b = G.SceneBuilder()
b.set_build_device(0)
root, cam, rec = b.config_scene(3, 2000000, 3)
sc = G.Scene(b.flatten(root))
W, H = 3840, 2160
if len(sys.argv) > 1 and sys.argv[1] == "count":
    ys, xs = np.meshgrid(np.arange(0, H, 3), np.arange(0, W, 3), indexing="ij")
    rays = G.camera_rays(cam, W, H, xs.ravel() + 0.5, ys.ravel() + 0.5)
    c = sc.debug_count(rays)
    print("debug_count over %d rays: mean %.1f p50 %d p90 %d p99 %d p99.9 %d max %d" % (
        len(c), c.mean(), *np.percentile(c, [50, 90, 99, 99.9]).astype(int), c.max()))
for conc in (0, 1):
    sc.set_option(L.OPT_SEG_CONCURRENT, conc)
    for spec in (0, 1):
        sc.set_option(L.OPT_AA_SPECULATE, spec)
        for N in (1, 8):
            opts = G.render_opts(mode=L.MODE_ADAPTIVE_AA, recurs=rec, tile_first=0, tile_stride=N)
            buf = torch.zeros((H, W, 5), dtype=torch.float64, device="cuda")
            best = None
            for i in range(8):
                st = sc.render_ptr(cam, W, H, opts, buf.data_ptr(), 0, dev=True)
                if best is None or st.kernel_ms < best.kernel_ms: best = st
            print("%s concurrent %d speculate %d share 1/%d: kernel_ms %.3f launches %d family_ms %s family_launches %s rays %d/%d" % (
                os.path.basename(os.environ.get("GLOME_LIB", "default")), conc, spec, N, best.kernel_ms, best.launches,
                [round(x, 3) for x in best.family_ms], list(best.family_launches), best.rays_primary, best.rays_shadow), flush=True)
