"""N-GPU tile sharding on real devices: run under `gpurun --gpus 2` (skipped with fewer than 2 GPUs)."""
import os
import socket
import sys

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    import torch
    import torch.distributed as dist
    import glome_b200 as G
    from glome_b200 import _lib as L
    from glome_b200.dist import ShardedRenderer
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    b = G.SceneBuilder()
    root, cam, rec = b.config_scene(2, 50000)
    fs = b.flatten(root)
    sc = G.Scene(fs, rank)
    ok = True
    for mode in (L.MODE_ONE_RAY, L.MODE_ADAPTIVE_AA):
        w, h = 640, 360
        rdr = ShardedRenderer(sc, cam, w, h, mode, rec, rank=rank, world=world, want_tcolor=True)
        rdr.render_frame_dev()
        torch.cuda.synchronize()
        full = rdr.tcolor.cpu().numpy()
        rgb, _ = rdr.render_frame_host()
        ref, ref_rgb, _ = sc.render(cam, w, h, G.render_opts(mode=mode, recurs=rec), want_rgb8=True)
        ok = ok and np.array_equal(full, ref) and np.array_equal(rgb.numpy().view(np.uint32), ref_rgb)
    if rank == 0:
        q.put(bool(ok))
    dist.barrier()
    dist.destroy_process_group()


def test_two_gpu_sharded_frame_is_bit_identical():
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    ok = q.get(timeout=300)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert ok


@pytest.mark.gpu
@pytest.mark.parametrize("config,n,w,h", [(2, 30000, 333, 201), (1, 0, 160, 110)])
def test_single_process_multi_gpu_frame_is_bit_identical(config, n, w, h):
    """glome_multi_*: one process, tile i on device i % N, tiles gathered on the first device by peer copies.
    Uses the box's GPUs when there are several; two handles on GPU 0 exercise the same code on a 1-GPU box."""
    import glome_b200 as G
    from glome_b200 import _lib as L
    ndev = L.load().glome_device_count()
    devices = list(range(min(ndev, 4))) if ndev >= 2 else [0, 0, 0]
    b = G.SceneBuilder()
    root, cam, rec = b.config_scene(config, n)
    fs = b.flatten(root)
    one = G.Scene(fs, 0)
    multi = G.MultiScene(fs, devices)
    for mode in (L.MODE_ONE_RAY, L.MODE_ADAPTIVE_AA):
        opts = G.render_opts(mode=mode, recurs=rec)
        t1, r1, s1 = one.render(cam, w, h, opts, want_rgb8=True)
        for _ in range(2):  # the second frame may take the speculative AA schedule
            tm, rm, sm = multi.render(cam, w, h, opts, want_rgb8=True)
            assert tm.tobytes() == t1.tobytes()
            assert rm.tobytes() == r1.tobytes()
        tm2, _, _ = multi.render(cam, w, h, opts, want_rgb8=False)
        assert tm2.tobytes() == t1.tobytes()
        _, rm2, _ = multi.render(cam, w, h, opts, want_rgb8=True, want_tcolor=False)
        assert rm2.tobytes() == r1.tobytes()
    assert sm.rays_primary >= s1.rays_primary and sm.kernel_ms > 0
    multi.close()
