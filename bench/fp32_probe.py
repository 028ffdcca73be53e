"""fp32_probe.py -- developer harness: configs[4] / configs[2] with Flt = Float, AA schedule forced, against the FP64 frame."""
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
fs = b.flatten(root)
for prec in (64, 32):
    sc = G.Scene(fs, 0, precision=prec)
    for W, H, mode, spec in ((3840, 2160, L.MODE_ADAPTIVE_AA, 0), (3840, 2160, L.MODE_ADAPTIVE_AA, 1), (1920, 1080, L.MODE_ONE_RAY, -1)):
        sc.set_option(L.OPT_AA_SPECULATE, spec)
        opts = G.render_opts(mode=mode, recurs=rec)
        buf = torch.zeros((H, W, 5), dtype=torch.float64, device="cuda")
        best = None
        for i in range(6):
            st = sc.render_ptr(cam, W, H, opts, buf.data_ptr(), 0, dev=True)
            if best is None or st.kernel_ms < best.kernel_ms: best = st
        print("prec %d %dx%d mode %d spec %d: kernel_ms %.3f launches %d rays %d/%d bvh %d tri %d bih %d prim %d family_ms %s" % (
            prec, W, H, mode, spec, best.kernel_ms, best.launches, best.rays_primary, best.rays_shadow, best.visits_bvh, best.tests_tri,
            best.visits_bih, best.tests_prim, [round(x, 3) for x in best.family_ms]), flush=True)
    sc.close()
