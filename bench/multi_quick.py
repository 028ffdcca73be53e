"""multi_quick.py -- developer harness: config 2 at 1080p through glome_multi_render (one process, N GPUs, peer copies)."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import glome_b200 as G
from glome_b200 import _lib as L
b = G.SceneBuilder()
b.set_build_device(0)
root, cam, rec = b.config_scene(2, 1000000)
fs = b.flatten(root)
ndev = L.load().glome_device_count()
for n in [k for k in (1, 2, 4, 8) if k <= ndev]:
    m = G.MultiScene(fs, list(range(n)))
    opts = G.render_opts(mode=L.MODE_ONE_RAY, recurs=rec)
    ms, wall = [], []
    for i in range(10):
        t0 = time.perf_counter()
        _, rgb, st = m.render(cam, 1920, 1080, opts, want_rgb8=True, want_tcolor=False)
        wall.append((time.perf_counter() - t0) * 1e3)
        ms.append(st.kernel_ms)
    rays = st.rays_primary + st.rays_shadow
    print("glome_multi_render N=%d: device %.3f ms (min) wall %.3f ms (min, incl. D2H + numpy alloc)  %.0f Mrays/s" % (n, min(ms[2:]), min(wall[2:]), rays / (min(ms[2:]) * 1e-3) / 1e6))
    m.close()
