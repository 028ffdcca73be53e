"""Committed golden vectors (tests/golden/frames.npz, written by tests/golden/make_golden.py from the oracle; see
that script for what they are and are not).  CPU: the oracle still reproduces them bit for bit.  GPU: the CUDA path
through the C-ABI reproduces them (ids / depths / counts bit-exact, RGB within 1e-6 where libm `pow` is involved)."""
import os
import sys

import numpy as np
import pytest

import glome_b200 as G
from glome_b200 import _lib as L
import oracle as O

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "golden"))
import make_golden as MG

GOLD = np.load(os.path.join(HERE, "golden", "frames.npz"))


def check(scene_like, cam, rec, name, exact):
    for mode, tag in ((L.MODE_ONE_RAY, "one"), (L.MODE_ADAPTIVE_AA, "aa")):
        tc = scene_like.render(cam, MG.W, MG.H, G.render_opts(mode=mode, recurs=rec))[0]
        ref = GOLD["%s_%s" % (name, tag)]
        # RGB passes through libm / CUDA `pow` (Shader.hs:98): last-ulp differences between pow implementations are
        # allowed (1e-9 for the oracle on another host's glibc, the 1e-6 of BASELINE.json for the device); depth never
        assert np.abs(tc[..., :4] - ref[..., :4]).max() <= (1e-9 if exact else 1e-6), (name, tag)
        assert np.array_equal(tc[..., 4], ref[..., 4]), (name, tag)
    ys, xs = np.mgrid[0:MG.H:2, 0:MG.W:2]
    rays = G.camera_rays(cam, MG.W, MG.H, xs.ravel(), ys.ravel())
    hits = scene_like.rayint(rays)
    assert np.array_equal(hits["prim"], GOLD["%s_prim" % name])
    assert np.array_equal(hits["t"], GOLD["%s_t" % name])
    assert np.array_equal(np.asarray(scene_like.shadow(rays, 60.0)).astype(np.uint8), GOLD["%s_shadow" % name])
    assert np.array_equal(scene_like.debug_count(rays), GOLD["%s_dbg" % name])


@pytest.mark.parametrize("name,config,n", MG.CASES)
def test_oracle_reproduces_the_golden_vectors(name, config, n):
    b, fs, cam, rec = MG.build(config, n)
    check(O.OracleScene(fs), cam, rec, name, exact=True)


@pytest.mark.gpu
@pytest.mark.parametrize("name,config,n", MG.CASES)
def test_gpu_reproduces_the_golden_vectors(name, config, n):
    b, fs, cam, rec = MG.build(config, n)
    check(G.Scene(fs, 0), cam, rec, name, exact=False)
