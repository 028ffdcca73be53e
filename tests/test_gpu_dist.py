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
