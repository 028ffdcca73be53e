"""floor_probe2.py -- developer harness: time of each traversal kernel alone (GLOME_OPT_SEG_CONCURRENT 0) against the number of
rays of a one-ray-per-pixel frame of the Mesh scene (configs[2]'s view at growing resolutions)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import torch
import glome_b200 as G
from glome_b200 import _lib as L
cfg = int(sys.argv[1]) if len(sys.argv) > 1 else 3
b = G.SceneBuilder()
b.set_build_device(0)
root, cam, rec = b.config_scene(cfg, 2000000 if cfg == 3 else 1000000, 3)
sc = G.Scene(b.flatten(root))
sc.set_option(L.OPT_SEG_CONCURRENT, 0)
for W, H in ((64, 36), (128, 72), (256, 144), (512, 288), (1024, 576), (2048, 1152), (3840, 2160)):
    opts = G.render_opts(mode=L.MODE_ONE_RAY, recurs=rec)
    buf = torch.zeros((H, W, 5), dtype=torch.float64, device="cuda")
    fam = []
    for i in range(8):
        st = sc.render_ptr(cam, W, H, opts, buf.data_ptr(), 0, dev=True)
        fam.append(list(st.family_ms))
    fam = np.array(fam)[2:]
    print("%s cfg %d %dx%d rays %d/%d kernel_ms %.3f  closest-bih %.1f us  any-bih %.1f us  bvh %.1f us (min over frames; bvh visits %d)" % (
        os.path.basename(os.environ.get("GLOME_LIB", "default")), cfg, W, H, st.rays_primary, st.rays_shadow, st.kernel_ms,
        1e3 * fam[:, 0].min(), 1e3 * fam[:, 1].min(), 1e3 * fam[:, 2].min(), st.visits_bvh), flush=True)
