"""quick_shard.py -- developer harness: one rank's share (tile i with i % N == 0) of the bench frame on one GPU."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import glome_b200 as G
from glome_b200 import _lib as L
N = int(sys.argv[1]) if len(sys.argv) > 1 else 8
b = G.SceneBuilder()
root, cam, rec = b.config_scene(2, 1000000)
fs = b.flatten(root)
sc = G.Scene(fs)
tc = np.zeros((1080, 1920, 5))
opts = G.render_opts(mode=0, recurs=rec, tile_first=0, tile_stride=N)
import torch
buf = torch.zeros((1080, 1920, 5), dtype=torch.float64, device="cuda")
ms = []
for i in range(8):
    st = sc.render_ptr(cam, 1920, 1080, opts, buf.data_ptr(), 0, dev=True)
    ms.append(st.kernel_ms)
print("stride %d: kernel_ms min %.3f med %.3f rays %d/%d launches %d" % (N, min(ms), sorted(ms)[4], st.rays_primary, st.rays_shadow, st.launches))
