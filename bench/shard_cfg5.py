"""shard_cfg5.py -- developer harness: one rank's share (tiles i with i % N == 0) of BASELINE.json configs[4]
(3840x2160 adaptive AA on the 2M-triangle Mesh scene) on one GPU.  usage: shard_cfg5.py N [reps]"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import torch
import glome_b200 as G
from glome_b200 import _lib as L
N = int(sys.argv[1]) if len(sys.argv) > 1 else 8
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 8
b = G.SceneBuilder()
b.set_build_device(0)
root, cam, rec = b.config_scene(3, 2000000, 3)
sc = G.Scene(b.flatten(root))
W, H = 3840, 2160
opts = G.render_opts(mode=L.MODE_ADAPTIVE_AA, recurs=rec, tile_first=0, tile_stride=N)
buf = torch.zeros((H, W, 5), dtype=torch.float64, device="cuda")
ms = []
for i in range(reps):
    st = sc.render_ptr(cam, W, H, opts, buf.data_ptr(), 0, dev=True)
    ms.append(st.kernel_ms)
print("cfg5 stride %d: kernel_ms min %.3f med %.3f rays %d/%d launches %d  family_ms %s" % (
    N, min(ms), sorted(ms)[len(ms) // 2], st.rays_primary, st.rays_shadow, st.launches, [round(x, 3) for x in st.family_ms]))
