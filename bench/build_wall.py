"""build_wall.py -- developer harness: wall time of the GPU tree builders against their own components (H2D + device +
D2H), for 10^6 random spheres (bih) and a 2*10^6-triangle height field (mesh).  VERDICT r1 #9: wall <= 2x the components."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import glome_b200 as G

rng = np.random.default_rng(1)
n = 1000000
c, r = rng.uniform(-100, 100, size=(n, 3)), rng.uniform(0.05, 0.5, size=(n, 1))
bb = np.hstack([c - r, c + r])
for rep in range(4):
    t0 = time.perf_counter()
    t = G.scene.bih_build(bb, device=0)
    wall = (time.perf_counter() - t0) * 1e3
    comp = sum(t["timings_ms"])
    print("bih 1e6 spheres: wall %.1f ms (incl. %.1f ms of numpy copies of the result)  h2d %.2f device %.2f d2h %.2f  sum %.1f  wall/sum %.2f"
          % (wall, 0.0, *t["timings_ms"], comp, wall / comp))
b = G.SceneBuilder()
b.set_build_device(0)
for rep in range(3):
    t0 = time.perf_counter()
    root, cam, rec = b.config_scene(2, 1000000, 2)
    ms = b.last_build_ms()
    print("config_scene(2): last tree build  h2d %.2f device %.2f d2h %.2f wall %.2f  wall/sum %.2f   (whole scene construction %.0f ms)" % (
        ms[0], ms[1], ms[2], ms[3], ms[3] / max(1e-9, ms[0] + ms[1] + ms[2]), (time.perf_counter() - t0) * 1e3))
b2 = G.SceneBuilder()
b2.set_build_device(0)
for rep in range(3):
    root, cam, rec = b2.config_scene(3, 2000000, 3)
    ms = b2.last_build_ms()
    print("config_scene(3): last tree build  h2d %.2f device %.2f d2h %.2f wall %.2f  wall/sum %.2f" % (
        ms[0], ms[1], ms[2], ms[3], ms[3] / max(1e-9, ms[0] + ms[1] + ms[2])))
