"""GPU construction of the BIH (glome_b200/csrc/glome_build.cu): the level-synchronous
device builder must return exactly the arrays of the host builder (which tests/test_host_builder.py pins to the
oracle's literal restatement of Bih.hs:211-324): same pre-order nodes, same leaves, same permutation, same box."""
import hashlib

import numpy as np
import pytest

import glome_b200 as G
from glome_b200 import _lib as L
import oracle as O

pytestmark = pytest.mark.gpu


def sphere_boxes(n, seed, lo=-100.0, hi=100.0, rmin=0.05, rmax=0.5):
    rng = np.random.default_rng(seed)
    c = rng.uniform(lo, hi, (n, 3))
    r = rng.uniform(rmin, rmax, (n, 1))
    d = 0.0001  # bound_sphere pads by delta (Sphere.hs:78-81)
    return np.hstack([c - r - d, c + r + d])


def same_tree(a, b):
    assert a["root"] == b["root"]
    assert np.array_equal(a["bb"], b["bb"])
    assert np.array_equal(a["order"], b["order"])
    assert np.array_equal(a["leaves"], b["leaves"])
    assert a["nodes"].tobytes() == b["nodes"].tobytes()


@pytest.mark.parametrize("n,seed", [(1, 1), (3, 2), (4, 3), (5, 4), (17, 5), (1000, 6), (50000, 7)])
def test_gpu_bih_equals_host_and_oracle(n, seed):
    bb = sphere_boxes(n, seed)
    g = G.bih_build(bb, device=0)
    h = G.bih_build(bb)
    same_tree(g, h)
    o = O.bih_build(bb)  # the literal list-recursion restatement
    assert np.array_equal(g["order"], o["order"]) and np.array_equal(g["leaves"], o["leaves"])
    assert g["nodes"].tobytes() == o["nodes"].tobytes() and g["root"] == o["root"]


def test_gpu_bih_full_size_config2():
    bb = sphere_boxes(1000000, 2)
    g = G.bih_build(bb, device=0)
    h = G.bih_build(bb)
    same_tree(g, h)
    # every item appears once; leaves tile the permutation
    assert np.array_equal(np.sort(g["order"]), np.arange(len(bb)))
    assert g["leaves"][:, 1].sum() == len(bb)


def test_gpu_bih_big_small_and_degenerate_inputs():
    rng = np.random.default_rng(11)
    small = sphere_boxes(4000, 12, -10, 10, 0.01, 0.05)
    big = sphere_boxes(40, 13, -10, 10, 4.0, 9.0)       # bbsa' > 0.4 * bbsa' bb: the big/small partition fires
    mixed = np.vstack([small, big])[rng.permutation(4040)]
    same_tree(G.bih_build(mixed, device=0), G.bih_build(mixed))
    same = np.tile(np.array([[0.0, 0, 0, 1, 1, 1]]), (500, 1))  # identical boxes: no partition helps -> one leaf
    g = G.bih_build(same, device=0)
    same_tree(g, G.bih_build(same))
    assert g["root"] < 0 and len(g["nodes"]) == 0
    line = np.zeros((3000, 6)); line[:, 0] = np.arange(3000); line[:, 3] = line[:, 0] + 0.5; line[:, 4:] = 0.5
    same_tree(G.bih_build(line, device=0), G.bih_build(line))       # 1-D input: y and z partitions are degenerate
    grid = np.stack(np.meshgrid(np.arange(21.0), np.arange(21.0), np.arange(21.0), indexing="ij"), -1).reshape(-1, 3)
    lattice = np.hstack([grid - 0.3, grid + 0.3])                    # TestScene's lattice: many exact ties with mid
    same_tree(G.bih_build(lattice, device=0), G.bih_build(lattice))


def test_gpu_bih_infinite_box_is_an_error():
    bb = sphere_boxes(10, 3)
    bb[4, 3] = 1000000.0  # a bare plane inside a bih (Bih.hs:319-322)
    with pytest.raises(L.GlomeError):
        G.bih_build(bb, device=0)
    with pytest.raises(L.GlomeError):
        G.bih_build(bb)


def test_scene_built_on_gpu_is_the_same_flat_scene_and_frame():
    frames = []
    for dev in (-1, 0):
        b = G.SceneBuilder()
        b.set_build_device(dev)
        root, cam, rec = b.config_scene(2, 200000)
        fs = b.flatten(root)
        fv = G.FlatView(fs)
        sig = hashlib.sha1(fv.bihnodes.tobytes() + fv.ipool.tobytes() + fv.dpool.tobytes() + fv.nodes.tobytes()).hexdigest()
        sc = G.Scene(fs, 0)
        tc, _, _ = sc.render(cam, 320, 180, G.render_opts(mode=L.MODE_ONE_RAY, recurs=rec))
        frames.append((sig, hashlib.sha1(tc.tobytes()).hexdigest()))
        if dev == 0:
            ms = b.last_build_ms()
            assert ms[1] > 0  # device build time was recorded
    assert frames[0] == frames[1]


def grid_mesh(nx, ny, seed):
    """height-field grid: (nx+1)*(ny+1) shared vertices, 2*nx*ny triangles (config 3's shape)"""
    rng = np.random.default_rng(seed)
    xs, ys = np.meshgrid(np.arange(nx + 1.0), np.arange(ny + 1.0), indexing="ij")
    verts = np.stack([xs, rng.uniform(0, 2, xs.shape), ys], -1).reshape(-1, 3)
    i, j = np.meshgrid(np.arange(nx), np.arange(ny), indexing="ij")
    a = (i * (ny + 1) + j).ravel(); b = a + 1; c = a + (ny + 1); d = c + 1
    t = np.concatenate([np.stack([a, b, c], -1), np.stack([b, d, c], -1)])
    tris = np.full((len(t), 8), -1, dtype=np.int32)
    tris[:, :3] = t
    return verts, tris


def same_mesh_tree(a, b):
    assert a["root"] == b["root"]
    assert np.array_equal(a["bb"], b["bb"])
    assert np.array_equal(a["leafoff"], b["leafoff"])
    assert np.array_equal(a["leafpool"], b["leafpool"])
    assert a["nodes"].tobytes() == b["nodes"].tobytes()


@pytest.mark.parametrize("nx,ny,seed", [(1, 1, 1), (2, 1, 2), (7, 5, 3), (60, 40, 4), (300, 200, 5)])
def test_gpu_mesh_bvh_equals_host_and_oracle(nx, ny, seed):
    verts, tris = grid_mesh(nx, ny, seed)
    g = G.mesh_build(verts, tris, device=0)
    same_mesh_tree(g, G.mesh_build(verts, tris))
    if len(tris) <= 5000:
        o = O.mesh_build(verts, tris)
        assert g["nodes"].tobytes() == o["nodes"].tobytes() and np.array_equal(g["leafpool"], o["leafpool"])


def test_gpu_mesh_bvh_full_size_config3_and_soup():
    verts, tris = grid_mesh(1000, 1000, 3)  # 2 000 000 triangles
    g = G.mesh_build(verts, tris, device=0)
    same_mesh_tree(g, G.mesh_build(verts, tris))
    rng = np.random.default_rng(9)     # triangle soup: random sizes, so the big/small partition fires
    v = rng.uniform(-50, 50, (30000, 3))
    v[::7] *= 3
    t = np.full((10000, 8), -1, dtype=np.int32)
    t[:, :3] = rng.permutation(30000).reshape(-1, 3)
    same_mesh_tree(G.mesh_build(v, t, device=0), G.mesh_build(v, t))


def test_mesh_scene_built_on_gpu_is_the_same_flat_scene():
    sigs = []
    for dev in (-1, 0):
        b = G.SceneBuilder()
        b.set_build_device(dev)
        root, cam, rec = b.config_scene(3, 20000)
        fv = G.FlatView(b.flatten(root))
        sigs.append(hashlib.sha1(fv.bvhnodes.tobytes() + fv.bihnodes.tobytes() + fv.ipool.tobytes() + fv.dpool.tobytes()).hexdigest())
    assert sigs[0] == sigs[1]
