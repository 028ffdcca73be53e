"""config5_scaling.py -- BASELINE.json configs[4]: 3840x2160 adaptive AA on the 2M-triangle Mesh scene, tile-sharded
across N GPUs (torchrun).  Prints one JSON line (rank 0): device ms/frame (CUDA events, max over ranks), fps,
Mrays/s, and the same frame end to end (render + all-gather + D2H of the packed image).  Not the bench contract's
headline (bench.py is); the numbers land in profiles/.
  python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 bench/config5_scaling.py"""
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import torch.distributed as dist
import glome_b200 as G
from glome_b200 import _lib as L
from glome_b200.dist import ShardedRenderer

W, H, STEPS, WARM = 3840, 2160, 10, 3
rank = int(os.environ.get("RANK", "0")); local = int(os.environ.get("LOCAL_RANK", "0")); world = int(os.environ.get("WORLD_SIZE", "1"))
torch.cuda.set_device(local)
if world > 1:
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
b = G.SceneBuilder()
b.set_build_device(local)
root, cam, rec = b.config_scene(3, 2000000)
scene = G.Scene(b.flatten(root), local)
rdr = ShardedRenderer(scene, cam, W, H, L.MODE_ADAPTIVE_AA, rec, rank=rank, world=world, want_tcolor=False)
flush = torch.empty(512 * 1024 * 1024 // 4, dtype=torch.int32, device="cuda")


def barrier():
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()


for _ in range(WARM):
    rdr.render_frame_dev()
rdr.render_frame_dev(want_stats=True)
st = rdr.last_stats
cnt = torch.tensor([st.rays_primary, st.rays_shadow, st.rays_secondary], dtype=torch.float64, device="cuda")
if world > 1:
    dist.all_reduce(cnt)
rays = cnt.sum().item()
barrier()
evs = []
for _ in range(STEPS):
    flush.fill_(1)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); rdr.render_frame_dev(); e1.record()
    evs.append((e0, e1))
barrier()
tot = torch.tensor([sum(a.elapsed_time(c) for a, c in evs)], dtype=torch.float64, device="cuda")
if world > 1:
    dist.all_reduce(tot, op=dist.ReduceOp.MAX)
ms = tot.item() / STEPS
e2e = []
for i in range(WARM + STEPS):
    flush.fill_(1)
    barrier()
    t0 = time.perf_counter()
    rdr.render_frame_host(copy_on=0)
    torch.cuda.synchronize()
    if i >= WARM:
        e2e.append((time.perf_counter() - t0) * 1e3)
t2 = torch.tensor([sum(e2e)], dtype=torch.float64, device="cuda")
if world > 1:
    dist.all_reduce(t2, op=dist.ReduceOp.MAX)
ms2 = t2.item() / STEPS
if rank == 0:
    print(json.dumps({"config": "configs[4]: 2M-triangle Mesh (+ occluder bih), 3840x2160, adaptive AA, tiles sharded over %d GPU(s)" % world,
                      "n_gpus": world, "ms_per_frame": ms, "fps": 1000 / ms, "Mrays_per_s": rays / (ms * 1e-3) / 1e6,
                      "rays_per_frame": {"primary": cnt[0].item(), "shadow": cnt[1].item(), "secondary": cnt[2].item()},
                      "e2e_ms_per_frame": ms2, "e2e_fps": 1000 / ms2, "steps": STEPS, "warmup": WARM, "dtype": "f64"}))
if world > 1:
    dist.barrier()
    dist.destroy_process_group()
