"""ab_mesh.py -- developer A/B harness for the Mesh walk (GLOME_LIB picks the library): configs[4] whole and 1/8 share, configs[2],
and the walk alone at 4K / 590 K rays / 147 K rays.  Prints the frame checksum (all variants must agree)."""
import hashlib, os, sys
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
name = os.path.basename(os.environ.get("GLOME_LIB", "default"))
out = []
for W, H, mode, N in ((3840, 2160, L.MODE_ADAPTIVE_AA, 1), (3840, 2160, L.MODE_ADAPTIVE_AA, 8), (1920, 1080, L.MODE_ONE_RAY, 1)):
    opts = G.render_opts(mode=mode, recurs=rec, tile_first=0, tile_stride=N)
    buf = torch.zeros((H, W, 5), dtype=torch.float64, device="cuda")
    ms = [sc.render_ptr(cam, W, H, opts, buf.data_ptr(), 0, dev=True).kernel_ms for _ in range(12)]
    torch.cuda.synchronize()
    out.append("%dx%d/%d %.3f (%s)" % (W, H, N, min(ms[4:]), hashlib.sha1(buf.cpu().numpy().tobytes()).hexdigest()[:8]))
sc.set_option(L.OPT_SEG_CONCURRENT, 0)
for W, H in ((3840, 2160), (1024, 576), (512, 288)):
    opts = G.render_opts(mode=L.MODE_ONE_RAY, recurs=rec)
    buf = torch.zeros((H, W, 5), dtype=torch.float64, device="cuda")
    fam = [sc.render_ptr(cam, W, H, opts, buf.data_ptr(), 0, dev=True).family_ms[2] for _ in range(8)]
    out.append("bvh@%dx%d %.1f us" % (W, H, 1e3 * min(fam[2:])))
print(name, " | ".join(out), flush=True)
