#!/usr/bin/env python
"""make_golden.py -- writes tests/golden/frames.npz.

What these vectors are: frames and first-hit records computed by the ORACLE (oracle/glome_oracle.cpp, the CPU
restatement of GlomeTrace) on the four BASELINE.json scene families at small sizes.  They are regression pins: they
freeze what the oracle says today, so that (a) a later change to the oracle that alters any result is caught by
`-m "not gpu"` tests, and (b) the GPU path is also checked against committed numbers, not only against a
freshly built checker.  They are NOT outputs of the Haskell reference -- no GHC exists in this image, the reference
ships no fixtures of its own, and the oracle's parity status stays "unpinned against a running
reference" (oracle header).  Regenerate with:  python tests/golden/make_golden.py
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
sys.path.insert(0, os.path.dirname(HERE))
import glome_b200 as G
from glome_b200 import _lib as L
import oracle as O

W, H = 64, 48
CASES = [("testscene", 1, 0), ("spheres", 2, 20000), ("mesh", 3, 5000), ("csg", 4, 4)]


def build(config, n):
    b = G.SceneBuilder()
    root, cam, rec = b.config_scene(config, n)
    return b, b.flatten(root), cam, rec


def main():
    out = {}
    for name, config, n in CASES:
        b, fs, cam, rec = build(config, n)
        osc = O.OracleScene(fs)
        for mode, tag in ((L.MODE_ONE_RAY, "one"), (L.MODE_ADAPTIVE_AA, "aa")):
            tc, _ = osc.render(cam, W, H, G.render_opts(mode=mode, recurs=rec))
            out["%s_%s" % (name, tag)] = tc
        ys, xs = np.mgrid[0:H:2, 0:W:2]
        rays = G.camera_rays(cam, W, H, xs.ravel(), ys.ravel())
        hits = osc.rayint(rays)
        out["%s_prim" % name] = hits["prim"].astype(np.int32)
        out["%s_t" % name] = hits["t"]
        out["%s_shadow" % name] = osc.shadow(rays, 60.0).astype(np.uint8)
        out["%s_dbg" % name] = osc.debug_count(rays)
    np.savez_compressed(os.path.join(HERE, "frames.npz"), **out)
    print("wrote", os.path.join(HERE, "frames.npz"), sorted(out))


if __name__ == "__main__":
    main()
