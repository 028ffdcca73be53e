"""shares.py -- developer harness: one rank's share (tiles i with i % N == 0) of configs[4] / [1] / [2] on one GPU, N = 1, 2, 4, 8.
usage: shares.py [reps]   prints kernel_ms (min / median of the last half) and the frame's checksum; GLOME_LIB picks the library."""
import hashlib, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import torch
import glome_b200 as G
from glome_b200 import _lib as L
reps = int(sys.argv[1]) if len(sys.argv) > 1 else 12
cases = [(3, 2000000, 3840, 2160, L.MODE_ADAPTIVE_AA, "configs[4]"), (2, 1000000, 1920, 1080, L.MODE_ONE_RAY, "configs[1]"),
         (3, 2000000, 1920, 1080, L.MODE_ONE_RAY, "configs[2]")]
for cfg, n, W, H, mode, name in cases:
    b = G.SceneBuilder()
    b.set_build_device(0)
    root, cam, rec = b.config_scene(cfg, n, 3)
    sc = G.Scene(b.flatten(root))
    for N in (1, 2, 4, 8):
        opts = G.render_opts(mode=mode, recurs=rec, tile_first=0, tile_stride=N)
        buf = torch.zeros((H, W, 5), dtype=torch.float64, device="cuda")
        ms = []
        for i in range(reps):
            st = sc.render_ptr(cam, W, H, opts, buf.data_ptr(), 0, dev=True)
            ms.append(st.kernel_ms)
        torch.cuda.synchronize()
        sha = hashlib.sha1(buf.cpu().numpy().tobytes()).hexdigest()[:12]
        tail = ms[reps // 2:]
        print("%s %s share 1/%d: kernel_ms min %.3f med(last half) %.3f launches %d bvh %d tri %d family_ms %s sha %s" % (
            os.path.basename(os.environ.get("GLOME_LIB", "default")), name, N, min(ms), sorted(tail)[len(tail) // 2], st.launches,
            st.visits_bvh, st.tests_tri, [round(x, 3) for x in st.family_ms], sha), flush=True)
    del sc
