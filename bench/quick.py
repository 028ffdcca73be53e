"""quick.py -- developer A/B harness: time one config with the library named by GLOME_LIB and print a
checksum of the frame (all variants must print the same checksum).  Not part of the bench contract."""
import hashlib
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import glome_b200 as G
from glome_b200 import _lib as L

cfg = int(sys.argv[1]) if len(sys.argv) > 1 else 2
n = int(sys.argv[2]) if len(sys.argv) > 2 else 1000000
w = int(sys.argv[3]) if len(sys.argv) > 3 else 1920
h = int(sys.argv[4]) if len(sys.argv) > 4 else 1080
mode = int(sys.argv[5]) if len(sys.argv) > 5 else 0
reps = int(sys.argv[6]) if len(sys.argv) > 6 else 8
b = G.SceneBuilder()
t0 = time.time()
root, cam, rec = b.config_scene(cfg, n)
fs = b.flatten(root)
t1 = time.time()
sc = G.Scene(fs)
opts = G.render_opts(mode=mode, recurs=rec)
tc, _, st = sc.render(cam, w, h, opts)
ms = []
for i in range(reps):
    tc, _, st = sc.render(cam, w, h, opts)
    ms.append(st.kernel_ms)
rays = st.rays_primary + st.rays_shadow + st.rays_secondary
print("%-28s cfg%d n=%d %dx%d mode%d build %.1fs  kernel_ms min %.3f med %.3f  Mrays/s %.1f  rays %d/%d/%d launches %d "
      "visits bih %d prim %d bvh %d tri %d  sha %s" % (
          os.path.basename(os.environ.get("GLOME_LIB", "default")), cfg, n, w, h, mode, t1 - t0, min(ms), sorted(ms)[len(ms) // 2],
          rays / (min(ms) * 1e-3) / 1e6, st.rays_primary, st.rays_shadow, st.rays_secondary, st.launches, st.visits_bih,
          st.tests_prim, st.visits_bvh, st.tests_tri, hashlib.sha1(tc.tobytes()).hexdigest()[:12]))
