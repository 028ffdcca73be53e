"""traffic_probe.py -- run under `ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum`:
renders frames of one bench config (reference AA schedule, segments one after the other) so that the launch list of the
LAST frame gives every kernel's DRAM traffic and duration.  usage: traffic_probe.py <cfg 1..5> [frames]"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import glome_b200 as G
from glome_b200 import _lib as L
import bench as B
cfg = int(sys.argv[1])
frames = int(sys.argv[2]) if len(sys.argv) > 2 else 2
c = B.CFG[cfg]
b, fs, cam, rec = B.build_scene(G, cfg, 0 if c["scene"] in (2, 3, 5) else -1)
sc = G.Scene(fs)
sc.set_option(L.OPT_SEG_CONCURRENT, 0)
mode = L.MODE_ADAPTIVE_AA_STRICT if c["aa"] else L.MODE_ONE_RAY
opts = G.render_opts(mode=mode, recurs=rec)
buf = torch.zeros((c["h"], c["w"], 5), dtype=torch.float64, device="cuda")
for i in range(frames):
    st = sc.render_ptr(cam, c["w"], c["h"], opts, buf.data_ptr(), 0, dev=True)
print("cfg %d: %d launches per frame, kernel_ms %.3f" % (cfg, st.launches, st.kernel_ms))
