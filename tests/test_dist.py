"""world_size-2 test of the N>1 path's host logic on CPU (gloo): tile -> rank assignment, slot packing,
all-gather, unpack.  The per-rank render is done by the oracle here (no GPU in this container); the GPU
version of the same path is tests/test_gpu_dist.py."""
import os
import socket
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, w, h, mode, q):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import torch
    import torch.distributed as dist
    import glome_b200 as G
    from glome_b200 import _lib as L
    from glome_b200.dist import pack_host, unpack_host, tile_slots
    import oracle as O
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    b = G.SceneBuilder()
    root, cam, rec = b.config_scene(2, 3000)
    fs = b.flatten(root)
    osc = O.OracleScene(fs)
    # this rank's tiles only
    frame = np.zeros((h, w, 5))
    osc.render(cam, w, h, G.render_opts(mode=mode, recurs=rec, tile_first=rank, tile_stride=world), out=frame, threads=2)
    packed = pack_host(frame, 65, rank, world)
    assert packed.shape[0] == tile_slots(w, h, 65, world)
    mine = torch.from_numpy(packed.reshape(-1).copy())
    gathered = torch.zeros(world * mine.numel(), dtype=torch.float64)
    dist.all_gather_into_tensor(gathered, mine)
    full = np.zeros((h, w, 5))
    g = gathered.numpy().reshape((world,) + packed.shape)
    for r in range(world):
        unpack_host(g[r], full, 65, r, world)
    if rank == 0:
        ref = np.zeros((h, w, 5))
        osc.render(cam, w, h, G.render_opts(mode=mode, recurs=rec), out=ref, threads=2)
        q.put(bool(np.array_equal(full, ref)) and float(np.abs(ref).sum()) > 0)
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("mode,w,h", [(0, 200, 140), (1, 137, 70)])
def test_two_rank_tile_gather_gloo(mode, w, h):
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, w, h, mode, q)) for r in range(2)]
    for p in procs:
        p.start()
    ok = q.get(timeout=180)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert ok


def test_pack_unpack_roundtrip_and_slots():
    import glome_b200 as G
    from glome_b200.dist import pack_host, unpack_host, tile_slots
    rng = np.random.default_rng(0)
    w, h = 333, 140
    frame = rng.normal(size=(h, w, 5))
    for world in (1, 2, 3, 8):
        out = np.zeros_like(frame)
        total = 0
        for r in range(world):
            p = pack_host(frame, 65, r, world)
            assert p.shape[0] == tile_slots(w, h, 65, world) == -(-len(G.tile_rects(w, h, 65)) // world)
            unpack_host(p, out, 65, r, world)
            total += p.shape[0]
        assert np.array_equal(out, frame)
    # the product's tile enumeration equals the oracle's (chunk, Glome.hs:371-384)
    import oracle as O
    for (ww, hh) in [(720, 480), (1920, 1080), (3840, 2160), (65, 65), (66, 1), (130, 131)]:
        assert np.array_equal(G.tile_rects(ww, hh, 65), O.tile_rects(ww, hh, 65))
