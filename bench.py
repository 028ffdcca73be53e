#!/usr/bin/env python
"""bench.py -- the headline measurement (BASELINE.json metric: Mrays/s and fps).

A "step" is one frame of BASELINE.json configs[1]: a `bih` over 1 000 000 random spheres, 1920x1080,
one camera ray per pixel plus a shadow ray per light and hit (2 point lights), FP64, Surface material.

  value    whole-job Mrays/s (primary + shadow + secondary rays resolved / device time), inputs
           (the flattened scene) resident in HBM, framebuffer left in HBM
  e2e      the same metric through the reference-facing C-ABI call `glome_render` with HOST buffers:
           camera/options in, packed 0x00RRGGBB frame copied back to pinned host memory every step
  roofline the persistent trace kernel against the measured HBM copy bandwidth (MEASURED_PEAKS.json)
  cpu_baseline / --impl reference: the C++ oracle (a literal restatement of GlomeTrace; the Haskell
           reference cannot be built here: no GHC) on the host cores, bounded sample of the same frame

N > 1: the frame is sharded by 65x65 tile (tile i -> rank i mod N), scene replicated per GPU, one NCCL
all-gather per frame: fixed total work, "scaling": "strong".
"""
import argparse
import json
import os
import subprocess
import sys
import tempfile
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

CONFIG = 2
N_SPHERES = 1000000
WIDTH, HEIGHT = 1920, 1080
SEED = 2


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return json.load(open(p)), "measured"
        except Exception:
            pass
    return {"hbm_gbs": 6650.0}, "fallback"


class ClockSampler:
    """nvidia-smi clocks + throttle reasons during the timed region (B200_PROFILING.md)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.idx = gpu_index
        self.f = tempfile.NamedTemporaryFile("w+", suffix=".csv", delete=False)
        self.p = None

    def start(self):
        try:
            self.p = subprocess.Popen(["nvidia-smi", "-i", str(self.idx), "--query-gpu=" + self.Q,
                                       "--format=csv,noheader,nounits", "-lms", "100"], stdout=self.f,
                                      stderr=subprocess.DEVNULL)
        except Exception:
            self.p = None

    def stop(self):
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": []}
        if self.p is not None:
            time.sleep(0.15)
            self.p.terminate()
            try:
                self.p.wait(timeout=5)
            except Exception:
                self.p.kill()
        try:
            self.f.flush()
            rows = [r.strip().split(", ") for r in open(self.f.name).read().strip().splitlines() if r.strip()]
            sm = sorted(float(r[1]) for r in rows if len(r) >= 9)
            if sm:
                out["sm_mhz"] = sm[len(sm) // 2]
                out["sm_max_mhz"] = float(rows[0][2])
                names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
                seen = set()
                for r in rows:
                    for k, nm in enumerate(names):
                        if r[5 + k].strip().lower().startswith("active"):
                            seen.add(nm)
                out["reasons"] = sorted(seen)
                out["samples"] = len(sm)
        except Exception:
            pass
        try:
            os.unlink(self.f.name)
        except Exception:
            pass
        return out


def build_scene(G, build_device=-1):
    """build_device >= 0: the scene's `bih` is built on that GPU (glome_build.cu), -1: on the host; same tree."""
    b = G.SceneBuilder()
    b.set_build_device(build_device)
    root, cam, recurs = b.config_scene(CONFIG, N_SPHERES, SEED)
    fs = b.flatten(root)
    return b, fs, cam, recurs


def cpu_sample(G, fs, cam, recurs, seconds_target, threads):
    """Time the oracle (CPU restatement) on a bounded sample of the frame: the first k tiles."""
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import oracle as O  # the checker / CPU baseline: the only place bench.py touches oracle/
    from glome_b200 import _lib as L
    import numpy as np
    osc = O.OracleScene(fs)
    opts = G.render_opts(mode=L.MODE_ONE_RAY, recurs=recurs)
    frame = np.zeros((HEIGHT, WIDTH, 5))
    ntiles = len(O.tile_rects(WIDTH, HEIGHT, 65))
    # probe, then size the sample for ~seconds_target of wall time; tiles are taken with a stride so the
    # sample covers the whole frame rather than one corner
    def run(k):
        o = G.render_opts(mode=L.MODE_ONE_RAY, recurs=recurs, tile_first=0, tile_stride=max(1, ntiles // k))
        osc.stats()
        t0 = time.perf_counter()
        osc.render(cam, WIDTH, HEIGHT, o, threads=threads, max_tiles=k, out=frame)
        dt = time.perf_counter() - t0
        st = osc.stats()
        rays = st["rays_primary"] + st["rays_shadow"] + st["rays_secondary"]
        return rays, dt, st
    k = max(threads, 8)
    rays, dt, st = run(min(k, ntiles))
    if dt < seconds_target / 2 and k < ntiles:
        k = int(min(ntiles, max(k, k * seconds_target / max(dt, 1e-3))))
        rays, dt, st = run(k)
    return {"rays": rays, "seconds": dt, "tiles": min(k, ntiles), "ntiles": ntiles, "stats": st, "osc": osc}


def reference_arm(args):
    """--impl reference: the reference's CPU implementation of the path.  The reference is Haskell and
    cannot be compiled here (no GHC in the image), so this is the C++ oracle, all host threads."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    import glome_b200 as G
    threads = os.cpu_count() or 1
    b, fs, cam, recurs = build_scene(G)
    per_step = max(2.0, min(20.0, 90.0 / max(1, args.steps + args.warmup)))
    for _ in range(args.warmup):
        cpu_sample(G, fs, cam, recurs, per_step / 4, threads)
    rays = 0
    secs = 0.0
    tiles = 0
    for _ in range(args.steps):
        r = cpu_sample(G, fs, cam, recurs, per_step, threads)
        rays += r["rays"]
        secs += r["seconds"]
        tiles = r["tiles"]
        ntiles = r["ntiles"]
    v = rays / secs / 1e6
    line = {"impl": "reference", "metric": "Mrays/s", "value": v, "unit": "Mrays/s", "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1000.0 * secs / max(1, args.steps),
            "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": workload_config(1),
            "cpu_baseline": {"value": v, "unit": "Mrays/s", "cores": threads, "kind": "port",
                             "sample": "%d of %d 65x65 tiles per step (strided over the frame), one ray per pixel + "
                                       "shadow rays; C++ restatement of GlomeTrace, not GHC" % (tiles, ntiles)},
            "e2e": {"value": v, "unit": "Mrays/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line))
    return 0


def workload_config(n):
    return {"workload": "configs[1]: bih of %d random spheres, %dx%d, 1 ray/pixel + shadow rays to 2 point lights"
                        % (N_SPHERES, WIDTH, HEIGHT),
            "mode": "one_ray_per_pixel", "recurs": 3, "blocksize": 65, "seed": SEED,
            "parallelism": "tiles%d" % n, "l2": "flushed between timed steps (512 MiB write)"}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="glome_b200")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--all-configs", action="store_true",
                    help="also time the other BASELINE.json configs (printed to stderr as a table; the JSON line is unchanged)")
    args = ap.parse_args()
    if args.impl == "reference":
        return reference_arm(args)
    if args.warmup < 3:
        args.warmup = 3

    import numpy as np
    import torch
    import torch.distributed as dist
    import glome_b200 as G
    from glome_b200 import _lib as L
    from glome_b200.dist import ShardedRenderer

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: there is no CPU fallback (use --impl reference for the CPU arm)")
    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    # scene set-up (timed separately by the reference too, Glome.hs:448-453): the bih is built on this rank's GPU;
    # rank 0 of a 1-GPU run also times the host builder on the same boxes for the report
    setup = None
    if rank == 0 and world == 1:
        hb = build_scene(G)[0]
        host_ms = hb.last_build_ms()[3]
        hb.close()
    if rank == 0 and world == 1:
        build_scene(G, local_rank)[0].close()  # untimed first build: context, module load, first cudaMalloc of the work space
    b, fs, cam, recurs = build_scene(G, local_rank)
    gm = b.last_build_ms()
    if rank == 0 and world == 1:
        setup = {"bih_items": N_SPHERES, "bih_build_gpu_ms": {"h2d": gm[0], "device": gm[1], "d2h": gm[2], "wall": gm[3]},
                 "bih_build_host_ms": host_ms, "host_threads": os.cpu_count(),
                 "note": "same tree either way (tests/test_gpu_build.py); not part of the timed frame"}
    scene = G.Scene(fs, local_rank)
    # N > 1: the gathered framebuffer is the packed 0x00RRGGBB image (what blitTile writes, Glome.hs:353-358)
    rdr = ShardedRenderer(scene, cam, WIDTH, HEIGHT, L.MODE_ONE_RAY, recurs, rank=rank, world=world, want_tcolor=False)
    flush = torch.empty(512 * 1024 * 1024 // 4, dtype=torch.int32, device="cuda")

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- warm-up ----
    for _ in range(args.warmup):
        rdr.render_frame_dev()
    barrier()
    # rays per frame (whole job) from the kernels' own counters
    rdr.render_frame_dev(want_stats=True)
    st = rdr.last_stats
    counts = torch.tensor([st.rays_primary, st.rays_shadow, st.rays_secondary, st.visits_bih, st.tests_prim,
                           st.visits_bvh, st.tests_tri, st.overflow_rays], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(counts)
    c = counts.tolist()
    rays_frame = c[0] + c[1] + c[2]
    barrier()

    # ---- timed: device-resident (value) ----
    sampler = ClockSampler(local_rank) if rank == 0 else None
    if sampler:
        sampler.start()
    rdr.launches = 0
    evs = []
    kernel_ms = []
    barrier()
    t_wall0 = time.perf_counter()
    for _ in range(args.steps):
        flush.fill_(1)  # L2 flush, outside the event pair
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        rdr.render_frame_dev()
        e1.record()
        evs.append((e0, e1))
    barrier()
    t_wall = time.perf_counter() - t_wall0
    step_ms = [a.elapsed_time(bb) for a, bb in evs]
    tot = torch.tensor([sum(step_ms)], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(tot, op=dist.ReduceOp.MAX)
    total_ms = tot.item()
    launches = rdr.launches
    ms_per_step = total_ms / args.steps
    value = rays_frame / (ms_per_step * 1e-3) / 1e6

    # ---- timed: the dominant kernel alone (roofline): the persistent BIH traversal kernel, closest-hit + any-hit
    # launches of one frame, bracketed by CUDA events on the launching stream inside glome_render_dev ----
    opts1 = G.render_opts(mode=L.MODE_ONE_RAY, recurs=recurs, tile_first=rank, tile_stride=world)
    kt, tt, tl = [], [], 0
    for _ in range(min(args.steps, 10)):
        flush.fill_(1)
        torch.cuda.synchronize()
        s1 = scene.render_ptr(cam, WIDTH, HEIGHT, opts1, rdr.tcolor.data_ptr(), 0, dev=True,
                              stream=torch.cuda.current_stream().cuda_stream)
        kt.append(s1.kernel_ms)
        tt.append(s1.traverse_ms)
        tl = s1.traverse_launches
    kern_ms = float(np.mean(kt))
    trav_ms = float(np.mean(tt))

    # ---- timed: end to end through the C-ABI with host buffers (e2e) ----
    e2e_ms = []
    pinned = torch.zeros((HEIGHT, WIDTH), dtype=torch.int32).pin_memory()
    opts_e2e = G.render_opts(mode=L.MODE_ONE_RAY, recurs=recurs)
    barrier()
    for i in range(args.warmup + args.steps):
        flush.fill_(1)
        barrier()
        t0 = time.perf_counter()
        if world == 1:
            scene.render_ptr(cam, WIDTH, HEIGHT, opts_e2e, None, pinned.data_ptr(), dev=False)
        else:
            rdr.render_frame_host(copy_on=0)
        torch.cuda.synchronize()
        dt = (time.perf_counter() - t0) * 1e3
        if i >= args.warmup:
            e2e_ms.append(dt)
    e2e_t = torch.tensor([sum(e2e_ms)], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(e2e_t, op=dist.ReduceOp.MAX)
    e2e_ms_per_step = e2e_t.item() / args.steps
    e2e_value = rays_frame / (e2e_ms_per_step * 1e-3) / 1e6
    clocks = sampler.stop() if sampler else None

    if rank == 0:
        peaks, which = measured_peaks()
        # algorithmic bytes of the traversal launches of one frame on this rank: BIH branch 32 B,
        # sphere record 32 B, BVH branch 128 B, triangle 32 B Tri + 72 B vertices, 24 B hit record out per ray
        stl = rdr.last_stats
        rays_rank = stl.rays_primary + stl.rays_shadow
        alg_bytes = (stl.visits_bih * 32 + stl.tests_prim * 32 + stl.visits_bvh * 128 + stl.tests_tri * 104 + rays_rank * 24)
        achieved = alg_bytes / (trav_ms * 1e-3) / 1e9
        traffic = None
        try:  # DRAM bytes of the same launches from the committed ncu capture (profiles/), N=1 only
            if world == 1:
                traffic = json.load(open(os.path.join(ROOT, "profiles", "r1_traffic.json")))["dram_bytes_per_frame_traversal"]
        except Exception:
            pass
        own = None
        try:  # own-measured L2 / FP64 denominators (bench/peaks.cu on this pool's B200)
            own = json.load(open(os.path.join(ROOT, "profiles", "own_peaks.json")))
        except Exception:
            pass
        line = {
            "metric": "Mrays/s", "value": value, "unit": "Mrays/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "strong",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic", "config": workload_config(world),
            "fps": 1000.0 / ms_per_step,
            "rays_per_frame": {"primary": c[0], "shadow": c[1], "secondary": c[2], "overflow": c[7]},
            "e2e": {"value": e2e_value, "unit": "Mrays/s", "h2d_bytes_per_step": 160,
                    "d2h_bytes_per_step": WIDTH * HEIGHT * 4, "ms_per_step": e2e_ms_per_step,
                    "fps": 1000.0 / e2e_ms_per_step,
                    "api": "glome_render (C-ABI, host buffers): camera+options in, 0x00RRGGBB frame to pinned host memory"
                           if world == 1 else "ShardedRenderer.render_frame_host: render + NCCL all-gather on every rank + D2H of the frame on rank 0 (the displaying rank)"},
            "gpu_launches": launches,
            "clocks": clocks,
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peaks["hbm_gbs"], "unit": "GB/s",
                         "frac": achieved / peaks["hbm_gbs"], "traffic": traffic, "peak_source": which,
                         "kernel": "k_bih_traverse (persistent BIH traversal): %d launches per frame, closest-hit + any-hit" % tl,
                         "kernel_ms": trav_ms, "share_of_step": trav_ms / kern_ms, "frame_kernels_ms": kern_ms,
                         "algorithmic_bytes_per_launch": alg_bytes,
                         "visits": {"bih_branch": stl.visits_bih, "sphere_tests": stl.tests_prim},
                         "note": "working set (~96 MB) is L2-resident: see profiles/ for L2 and issue-slot figures"},
            "wall_s_timed_region": t_wall,
        }
        if own:
            # the working set is L2-resident, so the meaningful memory ceiling is the L2's: algorithmic bytes per second
            # against the own-measured L2 streaming-read bandwidth (an upper bound on what reaches L2: L1 hits ~50 %)
            line["roofline"]["l2_own_measured"] = {"achieved": achieved, "peak": own["l2_read_gbs"], "unit": "GB/s",
                                                   "frac": achieved / own["l2_read_gbs"],
                                                   "fp64_nofma_peak_gops": own["fp64_nofma_gops"],
                                                   "l2_dependent_load_cycles": own["l2_dependent_load_cycles"],
                                                   "source": "profiles/own_peaks.json (bench/peaks.cu)"}
        if setup:
            line["scene_setup"] = setup
        if not args.no_cpu_baseline and world == 1:
            threads = os.cpu_count() or 1
            r = cpu_sample(G, fs, cam, recurs, 12.0, threads)
            line["cpu_baseline"] = {
                "value": r["rays"] / r["seconds"] / 1e6, "unit": "Mrays/s", "cores": threads, "kind": "port",
                "sample": "%d of %d tiles (strided over the frame), %.1f s; C++ restatement of GlomeTrace (oracle/), "
                          "not GHC: no Haskell toolchain in this image" % (r["tiles"], r["ntiles"], r["seconds"]),
                "ref_visits_per_ray": {"bih_branch": r["stats"]["bih_branch"] / max(1, r["rays"]),
                                       "sphere_tests": r["stats"]["node_1"] / max(1, r["rays"])}}
        if args.all_configs and world == 1:
            all_configs_table(G, L, not args.no_cpu_baseline)
        print(json.dumps(line))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    return 0


def all_configs_table(G, L, with_cpu):
    """The other BASELINE.json configs (parity-test cases, not bench lines): device time per frame on one GPU and,
    optionally, the oracle on the host cores for the same frame (bounded sample of tiles)."""
    import numpy as np
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    rows = [("1 TestScene 720x480, 1 ray/px", 1, 0, 720, 480, L.MODE_ONE_RAY),
            ("1 TestScene 720x480, adaptive AA", 1, 0, 720, 480, L.MODE_ADAPTIVE_AA),
            ("2 1M spheres 1920x1080, 1 ray/px", 2, 1000000, 1920, 1080, L.MODE_ONE_RAY),
            ("2 1M spheres 720x480, adaptive AA", 2, 1000000, 720, 480, L.MODE_ADAPTIVE_AA),
            ("3 2M-tri mesh 1920x1080, 1 ray/px", 3, 2000000, 1920, 1080, L.MODE_ONE_RAY),
            ("4 CSG grid 1280x720, recurs 5", 4, 16, 1280, 720, L.MODE_ONE_RAY),
            ("5 2M-tri mesh 3840x2160, adaptive AA", 5, 2000000, 3840, 2160, L.MODE_ADAPTIVE_AA)]
    print("%-40s %10s %8s %12s %12s %10s" % ("config", "ms/frame", "fps", "Mrays/frame", "GPU Mrays/s", "CPU Mrays/s"), file=sys.stderr)
    for name, cfg, n, w, h, mode in rows:
        b = G.SceneBuilder()
        root, cam, rec = b.config_scene(cfg, n)
        fs = b.flatten(root)
        sc = G.Scene(fs)
        opts = G.render_opts(mode=mode, recurs=rec)
        ms = []
        for i in range(6):
            tc, _, st = sc.render(cam, w, h, opts)
            if i >= 2:
                ms.append(st.kernel_ms)
        rays = st.rays_primary + st.rays_shadow + st.rays_secondary
        cpu = ""
        if with_cpu:
            import oracle as O
            osc = O.OracleScene(fs)
            nt = len(O.tile_rects(w, h, 65))
            k = min(nt, 48)
            o2 = G.render_opts(mode=mode, recurs=rec, tile_first=0, tile_stride=max(1, nt // k))
            frame = np.zeros((h, w, 5))
            t0 = time.perf_counter()
            osc.render(cam, w, h, o2, threads=os.cpu_count() or 1, max_tiles=k, out=frame)
            dt = time.perf_counter() - t0
            so = osc.stats()
            cpu = "%.2f" % ((so["rays_primary"] + so["rays_shadow"] + so["rays_secondary"]) / dt / 1e6)
            osc.close()
        m = float(np.median(ms))
        print("%-40s %10.3f %8.1f %12.3f %12.1f %10s" % (name, m, 1000.0 / m, rays / 1e6, rays / (m * 1e-3) / 1e6, cpu),
              file=sys.stderr)
        sc.close()


if __name__ == "__main__":
    sys.exit(main())
