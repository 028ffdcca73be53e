"""Quick timing of the general-scene configs (1: TestScene 720x480, 4: CSG grid 1280x720): device ms per frame."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import glome_b200 as G
from glome_b200 import _lib as L
which = [int(x) for x in sys.argv[1].split(",")] if len(sys.argv) > 1 else [1, 4]
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 5
for cfg, n, w, h in ((1, 0, 720, 480), (4, 16, 1280, 720)):
    if cfg not in which:
        continue
    b = G.SceneBuilder()
    root, cam, rec = b.config_scene(cfg, n)
    sc = G.Scene(b.flatten(root))
    for mode in (L.MODE_ONE_RAY, L.MODE_ADAPTIVE_AA):
        opts = G.render_opts(mode=mode, recurs=rec)
        ms = []
        for i in range(reps):
            tc, _, st = sc.render(cam, w, h, opts)
            ms.append(st.kernel_ms)
        rays = st.rays_primary + st.rays_shadow + st.rays_secondary
        print("config %d mode %d: %.3f ms (min %.3f)  %.2f Mrays  %.1f Mrays/s  bih %d prim %d inst %d csg %d" % (
            cfg, mode, float(np.median(ms)), min(ms), rays / 1e6, rays / (np.median(ms) * 1e-3) / 1e6,
            st.visits_bih, st.tests_prim, st.visits_instance, st.csg_steps))
