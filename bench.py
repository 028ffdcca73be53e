#!/usr/bin/env python
"""bench.py -- the headline measurement (BASELINE.json metric: Mrays/s and fps at 720x480 and 4K).

Headline step = one frame of BASELINE.json configs[4]: the 2 000 000-triangle Mesh (shared vertex / normal
arrays, per-triangle texture / tag ids, a bih of occluder spheres above it, 2 point lights) rendered at
3840x2160 with GlomeView's adaptive anti-aliasing (1/8 subsample -> <= 2 rays/pixel, Glome.hs:226-323), FP64.
It is the largest single-GPU configuration and the one north_star shards across GPUs.

  value    whole-job Mrays/s (primary + shadow + secondary rays resolved / device time), the flattened scene
           resident in HBM, framebuffer left in HBM, L2 flushed between timed frames
  e2e      the same metric through the reference-facing C-ABI call `glome_render` with HOST buffers:
           camera / options in, packed 0x00RRGGBB frame copied back to pinned host memory every step
  roofline the dominant kernel of the frame against the measured HBM copy bandwidth (MEASURED_PEAKS.json)
  configs  at N = 1: the same figures (ms, fps, Mrays/s, e2e, roofline, cpu_baseline) for all five
           BASELINE.json configs, configs[0] (TestScene 720x480) included
  cpu_baseline / --impl reference: the C++ oracle (a literal restatement of GlomeTrace; the Haskell reference
           cannot be built here: no GHC) on the host cores, bounded sample of the same frame's tiles

N > 1: the frame is sharded by 65x65 tile (tile i -> rank i mod N), scene replicated per GPU, one NCCL
all-gather per frame: fixed total work, "scaling": "strong".  `frame_hash` is the SHA-1 of the gathered
0x00RRGGBB frame: it must be the same at N = 1, 2, 4, 8 (tiles never read across their edges).
"""
import argparse
import hashlib
import json
import os
import subprocess
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

HEADLINE = 5
# id -> BASELINE.json config.  scene = glome_sb_config_scene's id, n = its size argument
CFG = {
    1: dict(key="configs[0]", scene=1, n=0, w=720, h=480, aa=False,
            what="GlomeView TestScene.geom'' (TestScene.hs:183-197; oak PRNG: random-1.2 SplitMix StdGen), 720x480, "
                 "1 ray/pixel, maxdepth 3, 2 lights"),
    2: dict(key="configs[1]", scene=2, n=1000000, w=1920, h=1080, aa=False,
            what="bih of 1000000 random spheres, 1920x1080, 1 ray/pixel + shadow rays to 2 point lights"),
    3: dict(key="configs[2]", scene=3, n=2000000, w=1920, h=1080, aa=False,
            what="2000000-triangle Mesh (shared vertex/normal arrays, per-triangle tex/tag) + bih of 4096 occluder "
                 "spheres, 1920x1080, 1 ray/pixel + shadow rays to 2 point lights"),
    4: dict(key="configs[3]", scene=4, n=16, w=1280, h=720, aa=False,
            what="CSG grid: 256 x difference(intersection[box, sphere, cylinder], cone)[, sphere] under a bih, mirrors, "
                 "1280x720, 1 ray/pixel, recurs 5 (primary + 4 reflected generations)"),
    5: dict(key="configs[4]", scene=5, n=2000000, w=3840, h=2160, aa=True,
            what="2000000-triangle Mesh scene of configs[2] at 3840x2160, adaptive AA (1/8 subsample -> <= 2 rays/pixel, "
                 "65x65 tiles = 2040 tiles), tile-sharded across the GPUs"),
}
SEEDS = {1: 42, 2: 2, 3: 3, 4: 4, 5: 3}


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


def build_scene(G, cfg, build_device=-1):
    """build_device >= 0: `bih` / `mesh` trees are built on that GPU (glome_build.cu), -1: on the host; same tree."""
    c = CFG[cfg]
    b = G.SceneBuilder()
    b.set_build_device(build_device)
    root, cam, recurs = b.config_scene(c["scene"], c["n"], SEEDS[cfg])
    fs = b.flatten(root)
    return b, fs, cam, recurs


def mode_of(L, cfg):
    return L.MODE_ADAPTIVE_AA if CFG[cfg]["aa"] else L.MODE_ONE_RAY


def workload_config(cfg, n):
    c = CFG[cfg]
    return {"workload": "%s: %s" % (c["key"], c["what"]),
            "mode": "adaptive_aa" if c["aa"] else "one_ray_per_pixel", "width": c["w"], "height": c["h"],
            "blocksize": 65, "seed": SEEDS[cfg], "parallelism": "tiles%d" % n,
            "l2": "flushed between timed steps (512 MiB write)"}


def cpu_sample(G, cfg, fs, cam, recurs, seconds_target, threads):
    """Time the oracle (CPU restatement) on a bounded sample of the frame's tiles (strided over the frame)."""
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import oracle as O  # the checker / CPU baseline: the only place bench.py touches oracle/
    from glome_b200 import _lib as L
    import numpy as np
    c = CFG[cfg]
    w, h = c["w"], c["h"]
    osc = O.OracleScene(fs)
    frame = np.zeros((h, w, 5))
    ntiles = len(O.tile_rects(w, h, 65))

    def run(k):
        o = G.render_opts(mode=mode_of(L, cfg), recurs=recurs, tile_first=0, tile_stride=max(1, ntiles // k))
        osc.stats()
        t0 = time.perf_counter()
        osc.render(cam, w, h, o, threads=threads, max_tiles=k, out=frame)
        dt = time.perf_counter() - t0
        st = osc.stats()
        rays = st["rays_primary"] + st["rays_shadow"] + st["rays_secondary"]
        return rays, dt, st
    k = min(ntiles, max(threads, 8))
    rays, dt, st = run(k)
    if dt < seconds_target / 2 and k < ntiles:
        k = int(min(ntiles, max(k, k * seconds_target / max(dt, 1e-3))))
        rays, dt, st = run(k)
    osc.close()
    return {"rays": rays, "seconds": dt, "tiles": k, "ntiles": ntiles, "stats": st}


def cpu_baseline_entry(r, threads):
    return {"value": r["rays"] / r["seconds"] / 1e6, "unit": "Mrays/s", "cores": threads, "kind": "port",
            "sample": "%d of %d 65x65 tiles (strided over the frame), %.1f s; C++ restatement of GlomeTrace (oracle/), "
                      "not GHC: no Haskell toolchain in this image" % (r["tiles"], r["ntiles"], r["seconds"])}


def reference_arm(args):
    """--impl reference: the reference's CPU implementation of the path on the headline config.  The reference is
    Haskell and cannot be compiled here (no GHC in the image), so this is the C++ oracle on all host threads."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    os.environ["GLOME_HOST_ONLY"] = "1"  # scene construction only: this arm never maps the CUDA library
    import glome_b200 as G
    threads = os.cpu_count() or 1
    cfg = args.config
    b, fs, cam, recurs = build_scene(G, cfg)
    per_step = max(2.0, min(20.0, 90.0 / max(1, args.steps + args.warmup)))
    for _ in range(args.warmup):
        cpu_sample(G, cfg, fs, cam, recurs, per_step / 4, threads)
    rays, secs, tiles, ntiles = 0, 0.0, 0, 0
    for _ in range(args.steps):
        r = cpu_sample(G, cfg, fs, cam, recurs, per_step, threads)
        rays += r["rays"]
        secs += r["seconds"]
        tiles, ntiles = r["tiles"], r["ntiles"]
    v = rays / secs / 1e6
    line = {"impl": "reference", "metric": "Mrays/s", "value": v, "unit": "Mrays/s", "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1000.0 * secs / max(1, args.steps),
            "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": workload_config(cfg, 1),
            "cpu_baseline": {"value": v, "unit": "Mrays/s", "cores": threads, "kind": "port",
                             "sample": "%d of %d 65x65 tiles per step (strided over the frame); C++ restatement of "
                                       "GlomeTrace, not GHC" % (tiles, ntiles)},
            "e2e": {"value": v, "unit": "Mrays/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line))
    return 0


class Ctx:
    pass


def measure(X, cfg, steps, warmup, with_cpu, cpu_seconds, setup_report=False):
    """One config on this job's GPUs: device-resident frames (value), the dominant kernel (roofline), end to end
    through the C-ABI with host buffers (e2e), optionally the CPU baseline.  Returns a dict on rank 0."""
    G, L, torch, dist, np = X.G, X.L, X.torch, X.dist, X.np
    from glome_b200.dist import ShardedRenderer
    c = CFG[cfg]
    w, h = c["w"], c["h"]
    rank, world, local_rank = X.rank, X.world, X.local_rank
    mode = mode_of(L, cfg)

    setup = None
    use_gpu_build = c["scene"] in (2, 3, 5)
    if setup_report and rank == 0 and world == 1 and use_gpu_build:
        hb = build_scene(G, cfg)[0]
        host_ms = hb.last_build_ms()[3]
        hb.close()
        build_scene(G, cfg, local_rank)[0].close()  # untimed first build: module load, first cudaMalloc of the work space
    b, fs, cam, recurs = build_scene(G, cfg, local_rank if use_gpu_build else -1)
    if setup_report and rank == 0 and world == 1 and use_gpu_build:
        gm = b.last_build_ms()
        setup = {"last_tree_build_gpu_ms": {"h2d": gm[0], "device": gm[1], "d2h": gm[2], "wall": gm[3]},
                 "last_tree_build_host_ms": host_ms, "host_threads": os.cpu_count(),
                 "note": "the scene's last bih/mesh tree; same tree either way (tests/test_gpu_build.py); not part of the timed frame"}
    scene = G.Scene(fs, local_rank)
    rdr = ShardedRenderer(scene, cam, w, h, mode, recurs, rank=rank, world=world, want_tcolor=False)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(warmup):
        rdr.render_frame_dev()
    barrier()
    # The frame's ray count is the REFERENCE's: its adaptive schedule, passes 1-4 tracing only the samples their decisions
    # ask for (GLOME_MODE_ADAPTIVE_AA_STRICT).  The timed frames may trace every pixel centre up front instead when that
    # is faster on a rank (same frame bit for bit, more rays): those extra rays are never counted as work done.
    if mode == L.MODE_ADAPTIVE_AA:
        rdr.opts.mode = L.MODE_ADAPTIVE_AA_STRICT
    rdr.render_frame_dev(want_stats=True)
    rdr.opts.mode = mode
    st = rdr.last_stats
    for _ in range(8):  # let the schedule choice settle (each schedule is timed twice before the faster one is kept)
        rdr.render_frame_dev()
        barrier()
    counts = torch.tensor([st.rays_primary, st.rays_shadow, st.rays_secondary, st.overflow_rays], dtype=torch.float64,
                          device="cuda")
    if world > 1:
        dist.all_reduce(counts)
    cnt = counts.tolist()
    rays_frame = cnt[0] + cnt[1] + cnt[2]
    barrier()
    # the gathered frame's hash: must not depend on N
    frame_hash = hashlib.sha1(rdr.rgb8.cpu().numpy().tobytes()).hexdigest()[:16]

    # ---- timed: device-resident (value) ----
    rdr.launches = 0
    evs = []
    barrier()
    t_wall0 = time.perf_counter()
    for _ in range(steps):
        X.flush.fill_(1)  # L2 flush, outside the event pair
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        rdr.render_frame_dev()
        e1.record()
        evs.append((e0, e1))
    barrier()
    t_wall = time.perf_counter() - t_wall0
    tot = torch.tensor([sum(a.elapsed_time(bb) for a, bb in evs)], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(tot, op=dist.ReduceOp.MAX)
    ms_per_step = tot.item() / steps
    launches = rdr.launches
    value = rays_frame / (ms_per_step * 1e-3) / 1e6

    # ---- the kernels of this rank's part of the frame, timed by CUDA events on the launching stream inside
    # glome_render_dev (per kernel family) ----
    # (for the roofline the reference schedule is timed: its visit counters and its kernel times belong together)
    opts1 = G.render_opts(mode=L.MODE_ADAPTIVE_AA_STRICT if mode == L.MODE_ADAPTIVE_AA else mode, recurs=recurs,
                          tile_first=rank, tile_stride=world)
    fam_ms = np.zeros(4)
    fam_n = np.zeros(4)
    kt = []
    reps = min(steps, 8)
    scene.set_option(L.OPT_SEG_CONCURRENT, 0)  # each traversal kernel timed alone (they overlap in the timed frames above)
    for _ in range(reps):
        X.flush.fill_(1)
        torch.cuda.synchronize()
        s1 = scene.render_ptr(cam, w, h, opts1, rdr.tcolor.data_ptr(), 0, dev=True,
                              stream=torch.cuda.current_stream().cuda_stream)
        kt.append(s1.kernel_ms)
        for k in range(4):
            fam_ms[k] += s1.family_ms[k]
            fam_n[k] = s1.family_launches[k]
    fam_ms /= reps
    kern_ms = float(np.mean(kt))
    scene.set_option(L.OPT_SEG_CONCURRENT, 1)

    # ---- timed: end to end through the C-ABI with host buffers (e2e) ----
    e2e_ms = []
    pinned = torch.zeros((h, w), dtype=torch.int32).pin_memory()
    opts_e2e = G.render_opts(mode=mode, recurs=recurs)
    barrier()
    n_e2e = max(3, min(steps, 10))
    for i in range(3 + n_e2e):
        X.flush.fill_(1)
        barrier()
        t0 = time.perf_counter()
        if world == 1:
            scene.render_ptr(cam, w, h, opts_e2e, None, pinned.data_ptr(), dev=False)
        else:
            rdr.render_frame_host(copy_on=0)
        torch.cuda.synchronize()
        dt = (time.perf_counter() - t0) * 1e3
        if i >= 3:
            e2e_ms.append(dt)
    e2e_t = torch.tensor([sum(e2e_ms)], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(e2e_t, op=dist.ReduceOp.MAX)
    e2e_ms_per_step = e2e_t.item() / n_e2e
    e2e_value = rays_frame / (e2e_ms_per_step * 1e-3) / 1e6

    # ---- the optional FP32 mode, reported separately (north_star; never the headline): the same frame with
    # Flt = Float on this GPU, device-resident like `value`, and how far its frame is from the FP64 one ----
    fp32 = None
    if world == 1:
        try:
            s32 = G.Scene(fs, local_rank, precision=32)
            tc64 = torch.zeros((h, w, 5), dtype=torch.float64, device="cuda")
            tc32 = torch.zeros((h, w, 5), dtype=torch.float64, device="cuda")
            o1 = G.render_opts(mode=mode, recurs=recurs)
            cs = torch.cuda.current_stream().cuda_stream
            scene.render_ptr(cam, w, h, o1, tc64.data_ptr(), 0, dev=True, stream=cs)
            # warm-up with a host sync per frame: work buffers allocated, and (adaptive AA) both schedules timed twice so that
            # the timed frames below all run the one the tuner keeps -- as the FP64 frames above do
            for _ in range(10):
                st32 = s32.render_ptr(cam, w, h, o1, tc32.data_ptr(), 0, dev=True, stream=cs)
                torch.cuda.synchronize()
            ev32 = []
            n32 = max(3, min(steps, 10))
            for _ in range(n32):
                X.flush.fill_(1)
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                s32.render_ptr(cam, w, h, o1, tc32.data_ptr(), 0, dev=True, stream=cs, want_stats=False)
                e1.record()
                ev32.append((e0, e1))
            torch.cuda.synchronize()
            ms32 = sum(a.elapsed_time(bb) for a, bb in ev32) / n32
            rays32 = rays_frame  # the reference schedule's ray count of this frame (a speculated AA frame traces more; never counted)
            d = (tc32[..., :4] - tc64[..., :4]).abs().amax(dim=-1)
            fp32 = {"dtype": "f32", "ms_per_step": ms32, "fps": 1000.0 / ms32, "value": rays32 / (ms32 * 1e-3) / 1e6,
                    "unit": "Mrays/s", "speedup_over_f64": ms_per_step / ms32,
                    "rgba_within_1e-3_of_f64": float((d <= 1e-3).double().mean().item()),
                    "rgba_median_abs_diff": float(d.median().item()),
                    "note": "Flt = Float (Vec.hs:7-9): payloads rounded once at upload, 16-byte BIH nodes / 64-byte BVH nodes; "
                            "pixels beyond 1e-3 are silhouette and shadow-edge pixels where the FP32 walk picks another "
                            "primitive (tests/test_gpu_f32.py)"}
            s32.close()
            del tc64, tc32
        except Exception as e:
            fp32 = {"dtype": "f32", "error": repr(e)}

    res = None
    if rank == 0:
        peaks, which = measured_peaks()
        # algorithmic bytes of one frame's launches of the dominant kernel on this rank: BIH
        # branch 32 B, primitive record 32 B, BVH branch 128 B, triangle 32 B Tri + 72 B vertices, Instance 192 B,
        # 24 B hit record out per ray
        stl = s1
        names = ["k_bih_traverse<closest>", "k_bih_traverse<any>", "k_bvh_closest", "k_gen_trace"]
        bytes_f = [0.0] * 4
        bih_total = stl.visits_bih * 32 + stl.tests_prim * 32
        if fam_ms[3] > 0:  # general scenes: the one tracer does everything
            bytes_f[3] = bih_total + stl.visits_instance * 192 + (stl.rays_primary + stl.rays_shadow + stl.rays_secondary) * 24
        else:
            # the two BIH traversal launches share their visit counters: split the bytes by their share of the time
            tb = fam_ms[0] + fam_ms[1]
            for k in (0, 1):
                bytes_f[k] = (bih_total * (fam_ms[k] / tb) if tb > 0 else 0.0)
            bytes_f[0] += stl.rays_primary * 24
            bytes_f[1] += stl.rays_shadow * 4
            bytes_f[2] = stl.visits_bvh * 128 + stl.tests_tri * 104 + stl.rays_primary * 24
        # the dominant kernel: BIH closest + any are one kernel template (k_bih_traverse), reported together
        groups = {"k_bih_traverse (persistent BIH traversal, closest-hit + any-hit launches)": (0, 1),
                  "k_bvh_closest (persistent Mesh BVH traversal)": (2,),
                  "k_gen_trace (persistent general-scene tracer: iterative scene-graph machine + shading)": (3,)}
        gname, gidx = max(groups.items(), key=lambda kv: sum(fam_ms[i] for i in kv[1]))
        g_ms = float(sum(fam_ms[i] for i in gidx))
        g_bytes = float(sum(bytes_f[i] for i in gidx))
        g_launch = int(sum(fam_n[i] for i in gidx))
        achieved = g_bytes / (g_ms * 1e-3) / 1e9 if g_ms > 0 else 0.0
        traffic, traffic_src = None, None
        try:  # DRAM bytes of the same kernel from the committed ncu capture of this round (profiles/), N=1 only
            if world == 1:
                t = json.load(open(os.path.join(ROOT, "profiles", "r2_traffic.json")))
                e = t.get(c["key"])
                if e:
                    fam = e.get("families", {}).get(gname.split(" ")[0])  # DRAM bytes per frame of the dominant kernel family
                    if fam is None and gname.startswith(e.get("kernel", "?")):
                        fam = e["dram_bytes_per_frame"]
                    if fam is not None:
                        traffic = fam / max(1, g_launch)
                        traffic_src = "committed ncu pass (%s), per launch; not measured in this run" % e["source"]
        except Exception:
            pass
        res = {
            "workload": "%s: %s" % (c["key"], c["what"]),
            "ms_per_step": ms_per_step, "fps": 1000.0 / ms_per_step, "value": value, "unit": "Mrays/s",
            "rays_per_frame": {"primary": cnt[0], "shadow": cnt[1], "secondary": cnt[2], "overflow": cnt[3]},
            "frame_hash": frame_hash,
            "e2e": {"value": e2e_value, "unit": "Mrays/s", "h2d_bytes_per_step": 160,
                    "d2h_bytes_per_step": w * h * 4, "ms_per_step": e2e_ms_per_step, "fps": 1000.0 / e2e_ms_per_step,
                    "api": "glome_render (C-ABI, host buffers): camera+options in, 0x00RRGGBB frame to pinned host memory"
                           if world == 1 else "ShardedRenderer.render_frame_host: render + NCCL all-gather on every rank + D2H of the frame on rank 0 (the displaying rank)"},
            "gpu_launches": launches,
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peaks["hbm_gbs"], "unit": "GB/s",
                         "frac": achieved / peaks["hbm_gbs"], "traffic": traffic, "traffic_source": traffic_src,
                         "peak_source": which, "kernel": gname, "launches_per_frame": g_launch,
                         "kernel_ms_per_frame": g_ms, "avg_launch_ms": g_ms / max(1, g_launch),
                         "share_of_step": g_ms / kern_ms if kern_ms > 0 else None, "frame_kernels_ms": kern_ms,
                         "timing": "CUDA events around every traversal launch of the reference-schedule frame, segments run one "
                                   "after the other for this measurement (in the timed frames a Bih and a Mesh segment overlap)",
                         "algorithmic_bytes_per_frame": g_bytes,
                         "family_ms": {names[k]: float(fam_ms[k]) for k in range(4) if fam_n[k] > 0},
                         "visits": {"bih_branch": stl.visits_bih, "prim_tests": stl.tests_prim, "bvh_branch": stl.visits_bvh,
                                    "tri_tests": stl.tests_tri, "instance": stl.visits_instance}},
            "wall_s_timed_region": t_wall,
        }
        if setup:
            res["scene_setup"] = setup
        if fp32:
            res["fp32"] = fp32
        if with_cpu and world == 1:
            threads = os.cpu_count() or 1
            r = cpu_sample(G, cfg, fs, cam, recurs, cpu_seconds, threads)
            res["cpu_baseline"] = cpu_baseline_entry(r, threads)
            res["cpu_baseline"]["ref_visits_per_ray"] = {"bih_branch": r["stats"]["bih_branch"] / max(1, r["rays"]),
                                                         "bvh_branch": r["stats"]["bvh_branch"] / max(1, r["rays"])}
    del rdr
    scene.close()
    b.close()
    torch.cuda.empty_cache()
    return res


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="glome_b200")
    ap.add_argument("--config", type=int, default=HEADLINE, help="headline config id (1..5 = configs[0..4]); default 5")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-all-configs", action="store_true", help="skip the per-config `configs` object (N = 1 only)")
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

    X = Ctx()
    X.G, X.L, X.torch, X.dist, X.np = G, L, torch, dist, np
    X.rank = int(os.environ.get("RANK", "0"))
    X.local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    X.world = int(os.environ.get("WORLD_SIZE", "1"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: there is no CPU fallback (use --impl reference for the CPU arm)")
    torch.cuda.set_device(X.local_rank)
    if X.world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", X.local_rank))
    X.flush = torch.empty(512 * 1024 * 1024 // 4, dtype=torch.int32, device="cuda")

    sampler = ClockSampler(X.local_rank) if X.rank == 0 else None
    if sampler:
        sampler.start()
    head = measure(X, args.config, args.steps, args.warmup, not args.no_cpu_baseline, 10.0, setup_report=True)
    clocks = sampler.stop() if sampler else None

    others = {}
    if X.world == 1 and not args.no_all_configs:
        for cfg in sorted(CFG):
            if cfg == args.config:
                continue
            try:
                others[CFG[cfg]["key"]] = measure(X, cfg, min(args.steps, 10), 3, not args.no_cpu_baseline, 4.0)
            except Exception as e:  # a failing side config must not take the headline line with it
                others[CFG[cfg]["key"]] = {"error": repr(e)}

    if X.rank == 0:
        line = {
            "metric": "Mrays/s", "value": head["value"], "unit": "Mrays/s", "n_gpus": X.world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": head["ms_per_step"], "higher_is_better": True, "scaling": "strong",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic", "config": workload_config(args.config, X.world),
            "fps": head["fps"], "rays_per_frame": head["rays_per_frame"], "frame_hash": head["frame_hash"],
            "e2e": head["e2e"], "gpu_launches": head["gpu_launches"], "clocks": clocks, "roofline": head["roofline"],
            "wall_s_timed_region": head["wall_s_timed_region"],
        }
        for k in ("scene_setup", "cpu_baseline", "fp32"):
            if k in head:
                line[k] = head[k]
        if X.world == 1 and not args.no_all_configs:
            cf = dict(others)
            cf[CFG[args.config]["key"]] = {k: head[k] for k in head if k not in ("scene_setup",)}
            line["configs"] = {k: cf[k] for k in sorted(cf)}
        print(json.dumps(line))
    if X.world > 1:
        dist.barrier()
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
