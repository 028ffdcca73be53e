// glome_cuda.cu -- sm_100a kernels and the C-ABI of libglomecuda.so (include/glome_cuda.h).
//
// Kernels:
//   K1 traverse   rayint / shadow over BIH, Mesh BVH and the scene graph      (glome_device.cuh)
//   K2 shade      trace + materialShader (shadow rays, Blinn, reflect/refract/warp)
//   K3 adaptive AA: per pass, a decide kernel (edge detect + warp-ballot/prefix compaction into a
//      ray queue) and the persistent trace kernel that drains the queue
//   K4 pack       rgbf / blitTile
// The trace kernel is persistent: one grid sized to the SM count, warps pull 32-sample chunks
// from an atomic counter, so queue lengths never travel to the host.
// Precision twins.  This file is compiled twice: as it is (FP64, GlomeVec's `type Flt = Double`, Vec.hs:9) and through
// glome_cuda_f32.cu with GLOME_F32 defined (the Float instance Vec.hs:7-8 leaves as a to-do: the optional FP32 mode).
// The FP32 copy lives in its own namespaces and exports every scene-bound entry with an `_f32` suffix; the FP64 entries
// forward to them when the handle says so (GlomeScene.precision), so a caller only chooses at glome_scene_create[_f32].
#ifdef GLOME_F32
#define glm glm_f32
#define gdev gdev_f32
#define gwave gwave_f32
#define ggen ggen_f32
#define glome_tagmap glome_tagmap_f32
#define GlomeScene GlomeScene_f32
#define GlomeMulti GlomeMulti_f32
#define GLOME_API(n) n##_f32
#define GLOME_PREC_NS glome_prec_f32
#define GLOME_PRECISION 32
#else
#define GLOME_API(n) n
#define GLOME_PREC_NS glome_prec_f64
#define GLOME_PRECISION 64
#endif
#include <cuda_runtime.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <string>
#include <thread>
#include <vector>

#include <algorithm>
#include <mutex>

#include "glome_device.cuh"
#include "glome_gen.cuh"
#include "glome_tagmap.h"
#include "glome_wave.cuh"

#ifndef GLOME_AA_SPEC_RATIO
#define GLOME_AA_SPEC_RATIO 0.6  /* flat scenes: speculate AA passes 1-4 when the last frame traced this share of the pixel centres */
#endif
#ifndef GEN_THREADS
#define GEN_THREADS 256
#endif
#ifndef GEN_MINBLOCKS
#define GEN_MINBLOCKS 2
#endif

using namespace gdev;

// ---------------------------------------------------------------------------------------------
// error plumbing
// ---------------------------------------------------------------------------------------------
std::string& glome_err_ref();  // host_base.cpp (libglomehost.so): the C-ABI's thread-local error string
#define g_err (glome_err_ref())

#define CK(call)                                                                                   \
    do {                                                                                           \
        cudaError_t e_ = (call);                                                                   \
        if (e_ != cudaSuccess) {                                                                   \
            g_err = std::string(#call) + ": " + cudaGetErrorString(e_);                            \
            return GLOME_ECUDA;                                                                    \
        }                                                                                          \
    } while (0)

namespace GLOME_PREC_NS {  // this file's own kernels and types, once per precision

// ---------------------------------------------------------------------------------------------
// device-side helpers shared by kernels
// ---------------------------------------------------------------------------------------------
struct TileGeom {
    int width, height, bs, ntx, nty;
    int nbx, nby;          // 8x4 micro-tiles per tile
    int slots_per_tile;    // nbx*nby*32
};
__host__ __device__ inline TileGeom make_geom(int width, int height, int bs) {
    TileGeom g;
    g.width = width; g.height = height; g.bs = bs;
    g.ntx = (width + bs - 1) / bs;   // chunk (Glome.hs:371-377): tiles start at k*bs, last one partial
    g.nty = (height + bs - 1) / bs;
    g.nbx = (bs + 7) / 8; g.nby = (bs + 3) / 4;
    g.slots_per_tile = g.nbx * g.nby * 32;
    return g;
}
// tile index in renderTiles' enumeration order: x chunks outer, y chunks inner (Glome.hs:384)
__device__ __forceinline__ void tile_rect(const TileGeom& g, int ti, int& xt, int& yt, int& tw, int& th) {
    int tx = ti / g.nty, ty = ti % g.nty;
    xt = tx * g.bs; yt = ty * g.bs;
    tw = min(g.bs, g.width - xt); th = min(g.bs, g.height - yt);
}

struct TC { Flt r, g, b, a, d; };  // TColor (Glome.hs:153)
__device__ __forceinline__ TC ld_tc(const double* __restrict__ buf, size_t pix) {
    const double* p = buf + pix * 5;
    TC c; c.r = p[0]; c.g = p[1]; c.b = p[2]; c.a = p[3]; c.d = p[4];
    return c;
}
__device__ __forceinline__ void st_tc(double* __restrict__ buf, size_t pix, const TC& c) {
    double* p = buf + pix * 5;
    p[0] = c.r; p[1] = c.g; p[2] = c.b; p[3] = c.a; p[4] = c.d;
}
__device__ __forceinline__ Flt cCmp(const TC& p, const TC& q) {  // Glome.hs:179-189
    Flt md;
    if (p.d == 0 && q.d == 0) md = 0;
    else md = (p.d > q.d) ? (p.d / q.d) - 1 : (q.d / p.d) - 1;
    return fabs_(q.r - p.r) + fabs_(q.g - p.g) + fabs_(q.b - p.b) + fabs_(q.a - p.a) + md;
}
__device__ __forceinline__ TC cAvg(const TC& a, const TC& b, const TC& c, const TC& d) {  // Glome.hs:191-197
    TC r;
    r.r = (a.r + b.r + c.r + d.r) * FL(0.25); r.g = (a.g + b.g + c.g + d.g) * FL(0.25); r.b = (a.b + b.b + c.b + d.b) * FL(0.25);
    r.a = (a.a + b.a + c.a + d.a) * FL(0.25); r.d = (a.d + b.d + c.d + d.d) * FL(0.25);
    return r;
}
__device__ __forceinline__ TC cAvg2(const TC& a, const TC& b) {  // Glome.hs:199-205
    TC r;
    r.r = (a.r + b.r) * FL(0.5); r.g = (a.g + b.g) * FL(0.5); r.b = (a.b + b.b) * FL(0.5); r.a = (a.a + b.a) * FL(0.5); r.d = (a.d + b.d) * FL(0.5);
    return r;
}
__device__ __forceinline__ TC tc_init() { TC c; c.r = 0; c.g = 0; c.b = 0; c.a = 0; c.d = GLM_INFINITY; return c; }
// getc (Glome.hs:233-235): outside the TILE -> (0,0,0,0,infinity)
__device__ __forceinline__ TC getc(const double* __restrict__ v, int width, int xt, int yt, int tw, int th, int x, int y) {
    if ((x >= xt) && (x < xt + tw) && (y >= yt) && (y < yt + th)) return ld_tc(v, (size_t)y * width + x);
    return tc_init();
}
// pass-5 output (Glome.hs:309-316)
__device__ __forceinline__ TC pass5_combine(const TC& color, const TC& a, const TC& b, const TC& c, const TC& d, bool lastx,
                                            bool lasty) {
    if (lastx) {
        if (lasty) return color;
        return cAvg2(color, cAvg2(a, b));
    }
    if (lasty) return cAvg2(color, cAvg2(a, d));
    return cAvg2(color, cAvg(a, b, c, d));
}
__device__ __forceinline__ Flt cap1(Flt x) { return (x >= 1) ? 1 - GLM_DELTA : x; }  // Glome.hs:98
__device__ __forceinline__ uint32_t rgbf(Flt r, Flt g, Flt b) {                       // Glome.hs:107 (Word32 arithmetic)
    uint32_t R = (uint32_t)(long long)floor(cap1(r) * 256);
    uint32_t G = (uint32_t)(long long)floor(cap1(g) * 256);
    uint32_t B = (uint32_t)(long long)floor(cap1(b) * 256);
    return R * 65536u + G * 256u + B;
}

struct DevStats {  // accumulated with atomics; the first nine are gwave's stats slots
    unsigned long long primary, shadow, secondary, overflow, perlin_range, bih, prim, bvh, tri, inst, csg;
};

__device__ __forceinline__ void flush_counters(DevStats* st, unsigned int primary, const RayCounters& rc, unsigned int ovf,
                                               unsigned int inst = 0, unsigned int csg = 0) {
    // warp-reduce then one atomic per warp and counter
    unsigned int vals[11] = {primary, rc.shadow, rc.secondary, ovf, rc.perlin_range, rc.cnt.bih, rc.cnt.prim, rc.cnt.bvh, rc.cnt.tri, inst, csg};
#pragma unroll
    for (int k = 0; k < 11; k++) {
        unsigned int v = vals[k];
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
        vals[k] = v;
    }
    if ((threadIdx.x & 31) == 0) {
        if (vals[0]) atomicAdd(&st->primary, (unsigned long long)vals[0]);
        if (vals[1]) atomicAdd(&st->shadow, (unsigned long long)vals[1]);
        if (vals[2]) atomicAdd(&st->secondary, (unsigned long long)vals[2]);
        if (vals[3]) atomicAdd(&st->overflow, (unsigned long long)vals[3]);
        if (vals[4]) atomicAdd(&st->perlin_range, (unsigned long long)vals[4]);
        if (vals[5]) atomicAdd(&st->bih, (unsigned long long)vals[5]);
        if (vals[6]) atomicAdd(&st->prim, (unsigned long long)vals[6]);
        if (vals[7]) atomicAdd(&st->bvh, (unsigned long long)vals[7]);
        if (vals[8]) atomicAdd(&st->tri, (unsigned long long)vals[8]);
        if (vals[9]) atomicAdd(&st->inst, (unsigned long long)vals[9]);
        if (vals[10]) atomicAdd(&st->csg, (unsigned long long)vals[10]);
    }
}
__device__ __forceinline__ void flush_gcnt(DevStats* st, unsigned int primary, const ggen::GCnt& c, unsigned int ovf) {
    RayCounters rc;
    rc.shadow = c.shadow; rc.secondary = c.secondary; rc.perlin_range = c.perlin;
    rc.cnt.bih = c.bih; rc.cnt.prim = c.prim; rc.cnt.bvh = c.bvh; rc.cnt.tri = c.tri;
    flush_counters(st, primary, rc, ovf, c.inst, c.csg);
}

__device__ __forceinline__ void hit_out(const Hit& h, GlomeHit* o) {
    o->t = ridepth(h);
    o->hit = h.hit;
    o->prim = h.hit ? h.prim : -1;
    o->sub = h.hit ? h.sub : -1;
    o->flags = h.flags;
    o->ntex = h.hit ? h.tex.n : 0;
    o->ntag = h.hit ? h.tag.n : 0;
#pragma unroll
    for (int i = 0; i < 3; i++) { o->pos[i] = 0; o->norm[i] = 0; }
    if (h.hit) {
        o->pos[0] = h.pos.x; o->pos[1] = h.pos.y; o->pos[2] = h.pos.z;
        o->norm[0] = h.norm.x; o->norm[1] = h.norm.y; o->norm[2] = h.norm.z;
    }
#pragma unroll
    for (int i = 0; i < GLOME_MAX_STACK; i++) {
        o->tex[i] = (h.hit && i < h.tex.n) ? h.tex.v[i] : 0;
        o->tag[i] = (h.hit && i < h.tag.n) ? h.tag.v[i] : 0;
    }
}

__device__ __forceinline__ Ray ld_ray(const double* __restrict__ rays, long long i) {
    const double* p = rays + 6 * i;
    return mkray(vec(p[0], p[1], p[2]), vec(p[3], p[4], p[5]));
}

// ---------------------------------------------------------------------------------------------
// batch query kernels (parity surfaces)
// ---------------------------------------------------------------------------------------------
// GEN = false: flat-class scenes (three fixed levels, glome_device.cuh); GEN = true: any scene graph, evaluated by the
// iterative machine of glome_gen.cuh (its stacks are this thread's local memory; no device recursion anywhere).
template <bool GEN>
__global__ void __launch_bounds__(128) k_rayint_batch(DScene S, long long n, const double* __restrict__ rays,
                                                      const double* __restrict__ tmax, int stride, GlomeHit* __restrict__ out) {
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        if constexpr (GEN) {
            ggen::QVM vm;
            ggen::GCnt c;
            ggen::gcnt_clear(c);
            ggen::gq_query(S, vm, S.root, ld_ray(rays, i), tmax[stride ? i : 0], false, c);
            ggen::ghit_out(S, vm.slot[0], out + i);
        } else {
            Hit h;
            rayint_scene_flat(S, S.root, ld_ray(rays, i), tmax[stride ? i : 0], h);
            hit_out(h, out + i);
        }
    }
}
template <bool GEN>
__global__ void __launch_bounds__(128) k_shadow_batch(DScene S, long long n, const double* __restrict__ rays,
                                                      const double* __restrict__ tmax, int stride, uint8_t* __restrict__ out) {
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        if constexpr (GEN) {
            ggen::QVM vm;
            ggen::GCnt c;
            ggen::gcnt_clear(c);
            out[i] = ggen::gq_query(S, vm, S.root, ld_ray(rays, i), tmax[stride ? i : 0], true, c) ? 1 : 0;
        } else out[i] = shadow_scene_flat(S, S.root, ld_ray(rays, i), tmax[stride ? i : 0]) ? 1 : 0;
    }
}
__global__ void __launch_bounds__(128) k_inside_batch(DScene S, long long n, const double* __restrict__ pts,
                                                      uint8_t* __restrict__ out) {
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        int ovf = 0;
        out[i] = ggen::gq_inside(S, S.root, vec(pts[3 * i], pts[3 * i + 1], pts[3 * i + 2]), &ovf) ? 1 : 0;
    }
}
// the pick query's variant: also the TraceResult's tag list, n*(2+16) int32 = {count, overflow, tags...}
template <bool GEN>
__global__ void __launch_bounds__(64) k_trace_tags(DScene S, long long n, const double* __restrict__ rays,
                                                   const double* __restrict__ tmax, int stride, int recurs,
                                                   GlomeHit* __restrict__ hits, int* __restrict__ tags) {
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        int* o = tags + (2 + 16) * i;
        if constexpr (GEN) {
            ggen::SHM sh;
            ggen::GCnt c;
            ggen::gcnt_clear(c);
            ggen::TagArena ta;
            ColorA col;
            int fl = 0;
            ggen::gs_trace<true>(S, sh, 0, S.root, ld_ray(rays, i), tmax[stride ? i : 0], recurs, col, c, fl, &ta);
            if (hits) { ggen::ghit_out(S, sh.tf[0].ri, hits + i); hits[i].flags |= fl; }
            o[0] = ta.n; o[1] = ta.overflow || ta.n > 16;
            for (int k = 0; k < 16; k++) o[2 + k] = k < ta.n ? S.tagvals[ta.v[k]] : -1;
        } else {
            RayCounters rc = {0, 0, 0, {0, 0, 0, 0}};
            ColorA col;
            Hit h;
            trace_flat(S, 0, S.root, ld_ray(rays, i), tmax[stride ? i : 0], recurs, col, h, rc);
            if (hits) hit_out(h, hits + i);
            const int nt = h.hit ? h.tag.n : 0;  // ts is empty for Surface materials: `ts ++ tags` = the hit's tag stack
            o[0] = nt; o[1] = 0;
            for (int k = 0; k < 16; k++) o[2 + k] = k < nt ? h.tag.v[k] : -1;
        }
    }
}
template <bool GEN>
__global__ void __launch_bounds__(GEN ? 64 : 128) k_trace_batch(DScene S, long long n, const double* __restrict__ rays,
                                                                const double* __restrict__ tmax, int stride, int recurs,
                                                                double* __restrict__ rgba, double* __restrict__ depth,
                                                                GlomeHit* __restrict__ hits, DevStats* st) {
    RayCounters rc = {0, 0, 0, {0, 0, 0, 0}};
    ggen::GCnt gc;
    ggen::gcnt_clear(gc);
    unsigned int ovf = 0, prim = 0;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        ColorA c;
        if constexpr (GEN) {
            ggen::SHM sh;
            int fl = 0;
            ggen::gs_trace<false>(S, sh, 0, S.root, ld_ray(rays, i), tmax[stride ? i : 0], recurs, c, gc, fl, nullptr);
            depth[i] = ggen::ghit_depth(sh.tf[0].ri);
            if (hits) { ggen::ghit_out(S, sh.tf[0].ri, hits + i); hits[i].flags |= fl; }
            if (fl) ovf++;
        } else {
            Hit h;
            trace_flat(S, 0, S.root, ld_ray(rays, i), tmax[stride ? i : 0], recurs, c, h, rc);
            depth[i] = ridepth(h);
            if (hits) hit_out(h, hits + i);
            if (h.flags) ovf++;
        }
        rgba[4 * i] = c.r; rgba[4 * i + 1] = c.g; rgba[4 * i + 2] = c.b; rgba[4 * i + 3] = c.a;
        prim++;
    }
    __syncwarp();
    if constexpr (GEN) flush_gcnt(st, prim, gc, ovf);
    else flush_counters(st, prim, rc, ovf);
}

// ---------------------------------------------------------------------------------------------
// K2/K3: the persistent sample tracer.
//   MODE 0: one ray per pixel over the selected tiles (renderTile, Glome.hs:162-176); work items
//           are 8x4 micro-tiles of each tile so a warp's rays are coherent.
//   MODE 1: adaptive-AA passes 1-4: queue of pixel ids, sample at getCoords x y, write v.
//   MODE 5: adaptive-AA pass 5: queue of pixel ids, sample at (x+0.5, y+0.5), write v2 via
//           pass5_combine.
// ---------------------------------------------------------------------------------------------
struct TraceParams {
    TileGeom g;
    DCamera cam;
    int recurs, tint;
    int tile_first, tile_stride, n_sel;  // selected tiles: ti = tile_first + k*tile_stride
    const int* queue;                    // MODE 1/5
    const int* queue_count;
    unsigned int* work_counter;          // zeroed before launch
    int chunk;                           // samples per warp fetch (1..32)
    const double* v;                     // MODE 5 reads
    double* out;                         // MODE 0/1: v ; MODE 5: v2
    DevStats* st;
};

// flat-class scenes the wavefront pipeline cannot take (more than 32 lights, more than GW_MAX_SEGS segments): one thread per sample
template <int MODE>
__global__ void __launch_bounds__(128, 1) k_trace_samples(DScene S, TraceParams P) {
    const int lane = threadIdx.x & 31;
    RayCounters rc = {0, 0, 0, {0, 0, 0, 0}};
    unsigned int ovf = 0, nprim = 0;
    const long long total = (MODE == 0) ? (long long)P.n_sel * P.g.slots_per_tile : (long long)(*P.queue_count);
    // Rays per warp-chunk.  The general interpreter serialises divergent lanes, so when a wave has too few samples
    // to fill the machine a full 32-ray chunk only lengthens the critical path: hand out smaller chunks then.
    const unsigned int chunk = (unsigned int)P.chunk;
    for (;;) {
        unsigned int base = 0;
        if (lane == 0) base = atomicAdd(P.work_counter, chunk);
        base = __shfl_sync(0xffffffffu, base, 0);
        if ((long long)base >= total) break;
        long long w = (long long)base + lane;
        bool valid = w < total && lane < (int)chunk;
        int x = 0, y = 0;
        if (MODE == 0) {
            if (valid) {
                int k = (int)(w / P.g.slots_per_tile), local = (int)(w % P.g.slots_per_tile);
                int ti = P.tile_first + k * P.tile_stride;
                int xt, yt, tw, th;
                tile_rect(P.g, ti, xt, yt, tw, th);
                int blk = local >> 5, l = local & 31;
                int px = (blk % P.g.nbx) * 8 + (l & 7), py = (blk / P.g.nbx) * 4 + (l >> 3);
                valid = px < tw && py < th;
                x = xt + px; y = yt + py;
            }
        } else if (valid) {
            int pix = P.queue[w];
            x = pix % P.g.width; y = pix / P.g.width;
        }
        if (valid) {
            Flt xc, yc;
            if (MODE == 5) getCoordsf(P.g.width, P.g.height, (Flt)x + FL(0.5), (Flt)y + FL(0.5), xc, yc);
            else getCoordsf(P.g.width, P.g.height, (Flt)x, (Flt)y, xc, yc);
            ColorA c;
            Flt hd;
            {
                Hit h;
                trace_flat(S, 0, S.root, camera_ray(P.cam, xc, yc), GLM_INFINITY, P.recurs, c, h, rc);
                hd = ridepth(h);
                if (h.flags) ovf++;
            }
            nprim++;
            TC col;
            col.r = c.r; col.g = c.g; col.b = c.b; col.a = c.a; col.d = hd;
            size_t pix = (size_t)y * P.g.width + x;
            if (MODE == 0) {
                if (P.tint) col.r = col.r + (col.d / 400);  // Glome.hs:174
                st_tc(P.out, pix, col);
            } else if (MODE == 1) {
                st_tc(P.out, pix, col);
            } else {
                int tx = x / P.g.bs, ty = y / P.g.bs;
                int xt = tx * P.g.bs, yt = ty * P.g.bs;
                int tw = min(P.g.bs, P.g.width - xt), th = min(P.g.bs, P.g.height - yt);
                TC a = getc(P.v, P.g.width, xt, yt, tw, th, x, y), b = getc(P.v, P.g.width, xt, yt, tw, th, x, y + 1);
                TC cc = getc(P.v, P.g.width, xt, yt, tw, th, x + 1, y + 1), d = getc(P.v, P.g.width, xt, yt, tw, th, x + 1, y);
                st_tc(P.out, pix, pass5_combine(col, a, b, cc, d, x == xt + tw - 1, y == yt + th - 1));
            }
        }
    }
    __syncwarp();
    flush_counters(P.st, nprim, rc, ovf);
}

// ---------------------------------------------------------------------------------------------
// General scenes: the persistent tracer.  A warp takes an 8x4 micro-tile of samples; every lane runs the scene-graph
// machine of glome_gen.cuh (QVM for rayint / shadow, SHM for trace + materialShader) on its sample.  Stacks (control
// words, hit slots, trace / material frames) are the thread's local memory: no device recursion.
// Measured on B200 (config 1 / config 4, ms per frame): a lane that pulls a new sample as soon as its own is done
// (per-lane refill, as the flat kernels do) is 2.7x SLOWER here (41 / 28 vs 15 / 8.4): the machine's rounds cost the sum
// of the states present in the warp, and coherent neighbours stay in the same state while strangers do not.
// ---------------------------------------------------------------------------------------------
#ifndef GEN_BLOCK_SYNC
#define GEN_BLOCK_SYNC 1
#endif
#ifndef GEN_PHASE_LOCK
#define GEN_PHASE_LOCK 1  /* needs GEN_BLOCK_SYNC 1 (block-uniform loop trip counts) */
#endif
#ifndef GEN_SYNC_ROUNDS
#define GEN_SYNC_ROUNDS 1  /* 1: the lanes of a warp start their queries together (warp-synchronous rounds) */
#endif
template <int MODE>
__global__ void __launch_bounds__(GEN_THREADS, GEN_MINBLOCKS) k_gen_trace(DScene S, TraceParams P) {
    const unsigned int FULL = 0xffffffffu;
    const int lane = threadIdx.x & 31;
    ggen::SHM sh;
    ggen::QRegs q;
    ggen::SRegs sr;
    ggen::GCnt gc;
    ggen::gcnt_clear(gc);
    unsigned int ovf = 0, nprim = 0;
    const long long total = (MODE == 0) ? (long long)P.n_sel * P.g.slots_per_tile : (long long)(*P.queue_count);
    const unsigned int chunk = (unsigned int)P.chunk;
    q.st = ggen::GS_DONE;
#if GEN_BLOCK_SYNC
    __shared__ unsigned int sbase;
#endif
    for (;;) {
        unsigned int base = 0;
#if GEN_BLOCK_SYNC
        // the warps of a block take neighbouring micro-tiles at the same moment, so that they run the same parts of the
        // machine's code at about the same time (instruction-cache locality across warps)
        __syncthreads();
        if (threadIdx.x == 0) sbase = atomicAdd(P.work_counter, chunk * (GEN_THREADS / 32));
        __syncthreads();
        base = sbase + (threadIdx.x >> 5) * chunk;
        if ((long long)sbase >= total) break;
#else
        if (lane == 0) base = atomicAdd(P.work_counter, chunk);
        base = __shfl_sync(FULL, base, 0);
        if ((long long)base >= total) break;
#endif
        const long long w = (long long)base + lane;
        bool valid = w < total && lane < (int)chunk;
        int px = 0, py = 0;
        if (MODE == 0) {
            if (valid) {
                int k = (int)(w / P.g.slots_per_tile), local = (int)(w % P.g.slots_per_tile);
                int ti = P.tile_first + k * P.tile_stride;
                int xt, yt, tw, th;
                tile_rect(P.g, ti, xt, yt, tw, th);
                int blk = local >> 5, l = local & 31;
                int qx = (blk % P.g.nbx) * 8 + (l & 7), qy = (blk / P.g.nbx) * 4 + (l >> 3);
                valid = qx < tw && qy < th;
                px = xt + qx; py = yt + qy;
            }
        } else if (valid) {
            int pix = P.queue[w];
            px = pix % P.g.width; py = pix / P.g.width;
        }
        sr.st = ggen::SS_FINISHED;
        if (valid) {
            Flt xc, yc;
            if (MODE == 5) getCoordsf(P.g.width, P.g.height, (Flt)px + FL(0.5), (Flt)py + FL(0.5), xc, yc);
            else getCoordsf(P.g.width, P.g.height, (Flt)px, (Flt)py, xc, yc);
            ggen::shm_start(sr, sh, 0, S.root, camera_ray(P.cam, xc, yc), GLM_INFINITY, P.recurs, nullptr);
        }
        // rounds: every unfinished lane advances its shading machine to its next query (or to the end of the sample), then
        // all those queries run side by side from their first step: the shadow rays of neighbouring pixels towards one
        // light, or their reflected rays, walk the scene graph as coherently as the primary rays did
        for (;;) {
            bool need = false;
            if (sr.st != ggen::SS_FINISHED) need = ggen::shm_step<false>(S, sr, sh, q, gc, nullptr);
#if GEN_PHASE_LOCK
            // all warps of the block advance their machines in lockstep rounds, so that at any moment the SM fetches the
            // code of one part of the machine (the kernel is bound by instruction fetch)
            if (__syncthreads_or(need) == 0) break;
            while (__syncthreads_or(need && q.st != ggen::GS_DONE)) {
#if GEN_PHASE_LOCK == 2
                if (need && q.st != ggen::GS_DONE) ggen::qvm_step_part<1>(S, q, sh.q, gc);
                __syncthreads();
                if (need && q.st != ggen::GS_DONE) ggen::qvm_step_part<2>(S, q, sh.q, gc);
#else
                if (need && q.st != ggen::GS_DONE) ggen::qvm_step(S, q, sh.q, gc);
#endif
            }
#elif GEN_SYNC_ROUNDS
            if (__ballot_sync(FULL, need) == 0) break;
            while (__ballot_sync(FULL, need && q.st != ggen::GS_DONE)) {
                if (need && q.st != ggen::GS_DONE) ggen::qvm_step(S, q, sh.q, gc);
            }
#else
            if (!need) break;
            while (q.st != ggen::GS_DONE) ggen::qvm_step(S, q, sh.q, gc);
#endif
        }
        if (valid) {  // get_color's result for this sample
            nprim++;
            if (sr.fl) ovf++;
            TC col;
            col.r = sr.rv_c.r; col.g = sr.rv_c.g; col.b = sr.rv_c.b; col.a = sr.rv_c.a; col.d = ggen::ghit_depth(sh.tf[0].ri);
            const size_t pix = (size_t)py * P.g.width + px;
            if (MODE == 0) {
                if (P.tint) col.r = col.r + (col.d / 400);  // Glome.hs:174
                st_tc(P.out, pix, col);
            } else if (MODE == 1) {
                st_tc(P.out, pix, col);
            } else {
                int tx = px / P.g.bs, ty = py / P.g.bs;
                int xt = tx * P.g.bs, yt = ty * P.g.bs;
                int tw = min(P.g.bs, P.g.width - xt), th = min(P.g.bs, P.g.height - yt);
                TC a = getc(P.v, P.g.width, xt, yt, tw, th, px, py), b = getc(P.v, P.g.width, xt, yt, tw, th, px, py + 1);
                TC cc = getc(P.v, P.g.width, xt, yt, tw, th, px + 1, py + 1), d = getc(P.v, P.g.width, xt, yt, tw, th, px + 1, py);
                st_tc(P.out, pix, pass5_combine(col, a, b, cc, d, px == xt + tw - 1, py == yt + th - 1));
            }
        }
    }
    __syncwarp();
    flush_gcnt(P.st, nprim, gc, ovf);
}


// ---------------------------------------------------------------------------------------------
// K3: adaptive-AA decide kernels (Glome.hs:226-323).  One block per selected tile.  Each
// candidate pixel of the pass either gets the average of its neighbours or is appended to the
// ray queue; queue slots are handed out per warp with ballot + popc prefix and one atomicAdd.
// ---------------------------------------------------------------------------------------------
struct DecideParams {
    TileGeom g;
    int tile_first, tile_stride;
    int pass;            // 1..5
    Flt threshold;
    double* v;           // pass 1 initialises it; passes 2-4 write averages into it
    double* v2;          // pass 5 writes averages into it
    int* queue;
    int* queue_count;    // zeroed before launch
    const double* spec;  // non-null (passes 1-4): samples at every pixel centre were traced up front, so a
                         // "trace this pixel" decision is a copy instead of a queue entry
};

#ifndef GK_DECIDE_MINBLOCKS
#define GK_DECIDE_MINBLOCKS 4  /* 64 regs, 32 warps per SM (79 regs / 24 warps: configs[4] +1 %) */
#endif
__global__ void __launch_bounds__(256, GK_DECIDE_MINBLOCKS) k_aa_decide(DecideParams P) {
    int ti = P.tile_first + blockIdx.x * P.tile_stride;
    int xt, yt, tw, th;
    tile_rect(P.g, ti, xt, yt, tw, th);
    const int width = P.g.width;
    const int lane = threadIdx.x & 31;
    // Only the pass's own pixels are enumerated (1/8, 1/8, 1/4, 1/2 and all of the tile): rows x slots-per-row, a slot
    // past the tile's edge is skipped.  The reference's initial fill of the tile buffer (MUV.replicate, Glome.hs:231) is
    // never read -- every pixel a pass looks at belongs to an earlier pass (Appendix C) and out-of-tile reads are
    // answered by getc -- so it is not written.
    int rows, per_row;
    switch (P.pass) {
        case 1: case 2: rows = (th + 1) >> 1; per_row = (((tw + 1) >> 1) + 1) >> 1; break;  // dy even; every other even dx
        case 3: rows = th >> 1; per_row = tw >> 1; break;                                      // dx, dy odd
        case 4: rows = th; per_row = (tw + 1) >> 1; break;                                     // dx + dy odd
        default: rows = th; per_row = tw; break;
    }
    const int nslots = rows * per_row;
    for (int base = 0; base < nslots; base += blockDim.x) {
        int i = base + threadIdx.x;
        bool need = false;
        int x = 0, y = 0;
        bool member = false;
        int dx = 0, dy = 0;
        if (i < nslots) {
            const int ry = i / per_row, k = i % per_row;
            switch (P.pass) {
                case 1: dy = 2 * ry; dx = 2 * (2 * k + (ry & 1)); break;          // (dx + dy) mod 4 == 0  (Glome.hs:241-250)
                case 2: dy = 2 * ry; dx = 2 * (2 * k + ((ry + 1) & 1)); break;    // (dx + dy) mod 4 == 2  (:251-262)
                case 3: dy = 2 * ry + 1; dx = 2 * k + 1; break;                    // (:272-281)
                case 4: dy = ry; dx = 2 * k + ((ry + 1) & 1); break;               // (:283-295)
                default: dy = ry; dx = k; break;                                   // (:301-319)
            }
            member = dx < tw && dy < th;
        }
        if (member) {
            x = xt + dx; y = yt + dy;
            size_t pix = (size_t)y * width + x;
            if (P.pass == 1) {
                need = true;
            } else {
                TC a, b, c, d;
                if (P.pass == 2) {
                    a = getc(P.v, width, xt, yt, tw, th, x - 2, y); b = getc(P.v, width, xt, yt, tw, th, x, y + 2);
                    c = getc(P.v, width, xt, yt, tw, th, x + 2, y); d = getc(P.v, width, xt, yt, tw, th, x, y - 2);
                } else if (P.pass == 3) {
                    a = getc(P.v, width, xt, yt, tw, th, x - 1, y - 1); b = getc(P.v, width, xt, yt, tw, th, x + 1, y - 1);
                    c = getc(P.v, width, xt, yt, tw, th, x + 1, y + 1); d = getc(P.v, width, xt, yt, tw, th, x - 1, y + 1);
                } else if (P.pass == 4) {
                    a = getc(P.v, width, xt, yt, tw, th, x - 1, y); b = getc(P.v, width, xt, yt, tw, th, x, y + 1);
                    c = getc(P.v, width, xt, yt, tw, th, x + 1, y); d = getc(P.v, width, xt, yt, tw, th, x, y - 1);
                } else {
                    a = getc(P.v, width, xt, yt, tw, th, x, y); b = getc(P.v, width, xt, yt, tw, th, x, y + 1);
                    c = getc(P.v, width, xt, yt, tw, th, x + 1, y + 1); d = getc(P.v, width, xt, yt, tw, th, x + 1, y);
                }
                Flt variance = fmax_(cCmp(a, c), cCmp(b, d));  // decide (Glome.hs:213-219)
                if (variance > P.threshold) need = true;
                else {
                    TC avg = cAvg(a, b, c, d);
                    if (P.pass == 5) st_tc(P.v2, pix, pass5_combine(avg, a, b, c, d, x == xt + tw - 1, y == yt + th - 1));
                    else st_tc(P.v, pix, avg);
                }
            }
        }
        if (P.spec && P.pass < 5) {
            if (need) st_tc(P.v, (size_t)y * width + x, ld_tc(P.spec, (size_t)y * width + x));
            // keep counting what the adaptive schedule would have traced (the next frame's schedule looks at it)
            unsigned int mm = __ballot_sync(0xffffffffu, need);
            if (mm && lane == __ffs(mm) - 1) atomicAdd(P.queue_count, __popc(mm));
            continue;
        }
        // warp-aggregated compaction
        unsigned int m = __ballot_sync(0xffffffffu, need);
        if (m) {
            int leader = __ffs(m) - 1;
            int slot = 0;
            if (lane == leader) slot = atomicAdd(P.queue_count, __popc(m));
            slot = __shfl_sync(0xffffffffu, slot, leader);
            if (need) P.queue[slot + __popc(m & ((1u << lane) - 1))] = y * width + x;
        }
    }
}

// K4: blitTile / rgbf (Glome.hs:353-358, 107-110)
__global__ void __launch_bounds__(256) k_pack_rgb8(TileGeom g, int tile_first, int tile_stride, const double* __restrict__ tc,
                                                   uint32_t* __restrict__ rgb8) {
    int ti = tile_first + blockIdx.x * tile_stride;
    int xt, yt, tw, th;
    tile_rect(g, ti, xt, yt, tw, th);
    for (int i = threadIdx.x; i < tw * th; i += blockDim.x) {
        size_t pix = (size_t)(yt + i / tw) * g.width + (xt + i % tw);
        TC c = ld_tc(tc, pix);
        rgb8[pix] = rgbf(c.r * c.a, c.g * c.a, c.b * c.a);
    }
}

}  // namespace GLOME_PREC_NS
using namespace GLOME_PREC_NS;

// ---------------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------------
struct GlomeScene {  // (global scope: the C-ABI's opaque handle)
    int precision;   // 64 or 32: FIRST member of both precision twins' handles
    int device;
    int scene_class;
    int sm_count;
    DScene d;
    std::vector<void*> bufs;
    // render workspace (grown on demand; one capacity per buffer: glome_render and glome_render_dev size them separately)
    double* v; size_t v_pix;
    double* v2; size_t v2_pix;
    uint32_t* rgb8; size_t rgb8_pix;
    double* spec; size_t spec_pix;
    int* queue; size_t queue_cap;
    int* queue_count; unsigned int* work_counter; DevStats* stats;
    // batch workspace
    void* bw[4]; size_t bw_cap[4];
    cudaEvent_t ev0, ev1;
    int launches;
    // launch geometry of the persistent kernels on this device, computed once at creation
    int g_bih[4], g_bvh, g_gen[3], g_flat[3];
    // environment switches (A/B runs), read once at creation: never in the frame path
    int env_no_speculate, env_aa_speculate, env_gen_chunk;
    // wavefront pipeline (flat scenes)
    bool use_wave;
    std::vector<gwave::Seg> segs;
    std::vector<int> segs_linear;
    std::vector<int> segs_reps;   // per segment: k_bih_traverse's branch-phase bound
    int env_bih_reps;
    gwave::Seg* segs_dev;
    size_t wave_cap;           // samples
    Flt* w_hit_t; int* w_hit_seg; int* w_hit_item; int* w_hit_sub; int* w_hit_flags;
    Flt* w_surf; unsigned int* w_occl; int2* w_squeue; int* w_squeue_count;
    unsigned int* w_counters;  // one work counter per persistent launch of a frame
    // two closest-hit segments side by side (a Bih next to a Mesh): the second one's own result set, stream and events
    Flt* w2_hit_t; int* w2_hit_seg; int* w2_hit_item; int* w2_hit_sub; int* w2_hit_flags;
    cudaStream_t st2; cudaEvent_t ev_fork, ev_join;
    int env_seg_concurrent, env_group_accel;
    int w_counter_next;
    std::vector<cudaEvent_t> tev;  // start/stop pairs around the traversal kernels of the last timed frame
    std::vector<int> tev_family;   // kernel family of each pair (GlomeRenderStats.family_ms)
    int tev_used;
    bool time_traversal;
    int n_scene_lights;
    // adaptive-AA schedule memory (flat scenes): how many pixel centres the last AA frame traced (or, when it
    // speculated, would have traced) in passes 1-4, read back asynchronously into pinned memory
    int* aa_counts_host;        // [0] unused, [1..4] per pass
    cudaEvent_t aa_ev;
    bool aa_valid;
    int aa_w, aa_h, aa_first, aa_stride, aa_bs;
    // ... and which of the two schedules (0 = adaptive passes, 1 = speculated centres) is faster for this geometry on this
    // device: each is timed with its own event pair, harvested without a host sync when a later frame starts
    cudaEvent_t aa_t0[2], aa_t1[2];
    bool aa_pending[2], aa_ms_valid[2];
    int aa_nsamp[2];  // a schedule's first sample (buffers allocated, caches cold) is discarded
    float aa_ms[2];
    unsigned int aa_frames;
};

#ifndef GLOME_F32
// the FP32 twin's entries (glome_cuda_f32.cu); a handle created by glome_scene_create_f32 is forwarded to them
extern "C" {
int glome_scene_destroy_f32(GlomeScene* s);
int glome_rayint_batch_f32(GlomeScene* s, int64_t n, const double* rays, const double* tmax, int tmax_stride, GlomeHit* out);
int glome_shadow_batch_f32(GlomeScene* s, int64_t n, const double* rays, const double* tmax, int tmax_stride, uint8_t* out);
int glome_inside_batch_f32(GlomeScene* s, int64_t n, const double* pts, uint8_t* inside);
int glome_trace_batch_f32(GlomeScene* s, int64_t n, const double* rays, const double* tmax, int tmax_stride, int recurs,
                          double* rgba, double* depth, GlomeHit* hits);
int glome_get_tags_f32(GlomeScene* s, const GlomeCamera* cam, int width, int height, int px, int py, int recurs, int32_t* tags,
                       int max_tags, int* ntags, int* truncated, GlomeHit* hit_out);
int glome_debug_count_batch_f32(GlomeScene* s, int64_t n, const double* rays, const double* tmax, int tmax_stride, int32_t* counts);
int glome_render_dev_f32(GlomeScene* s, const GlomeCamera* cam, int width, int height, const GlomeRenderOpts* o, double* tcolor_dev,
                         uint32_t* rgb8_dev, GlomeRenderStats* stats, void* stream);
int glome_render_f32(GlomeScene* s, const GlomeCamera* cam, int width, int height, const GlomeRenderOpts* o, double* tcolor,
                     uint32_t* rgb8, GlomeRenderStats* stats);
int64_t glome_scene_launches_f32(GlomeScene* s);
int glome_scene_set_option_f32(GlomeScene* s, int option, int value);
}
#define TWIN(call) do { if (s && s->precision == 32) return call; } while (0)
#else
#define TWIN(call) do { } while (0)
#endif

namespace GLOME_PREC_NS {

#ifndef GLOME_F32
extern "C" int glome_device_count(void) {
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) return 0;
    return n;
}
#endif

template <typename T, typename D>
static int upload(GlomeScene* s, const T* src, size_t n, D* dst) {
    void* p = nullptr;
    size_t bytes = (n ? n : 1) * sizeof(T);
    CK(cudaMalloc(&p, bytes));
    s->bufs.push_back(p);
    if (n) CK(cudaMemcpy(p, src, n * sizeof(T), cudaMemcpyHostToDevice));
    *dst = (const T*)p;
    return GLOME_OK;
}

// Top-level segments of a GLOME_CLASS_FLAT scene, in the reference's list-fold order (Solid.hs:327).
static bool stk_push(gwave::Stk& st, int x) {
    if (st.n >= GLOME_MAX_STACK) return false;
    for (int i = st.n; i >= 1; i--) st.v[i] = st.v[i - 1];
    st.v[0] = x;
    st.n++;
    return true;
}
static bool build_segments(const GlomeFlatScene* d, std::vector<gwave::Seg>& segs) {
    segs.clear();
    gwave::Stk rtex, rtag;
    memset(&rtex, 0, sizeof(rtex));
    memset(&rtag, 0, sizeof(rtag));
    int ni = d->root;
    GlomeNode nd = d->nodes[ni];
    gwave::Stk t0 = rtex, g0 = rtag;
    while (nd.type == GLOME_TEX || nd.type == GLOME_TAG) {
        if (!(nd.type == GLOME_TEX ? stk_push(t0, nd.b) : stk_push(g0, nd.b))) return false;
        ni = nd.a;
        nd = d->nodes[ni];
    }
    int first, count;
    if (nd.type == GLOME_GROUP) { rtex = t0; rtag = g0; first = nd.a; count = nd.b; }
    else { first = d->root; count = 1; }  // single object: its wrappers are handled as item wrappers
    for (int i = 0; i < count; i++) {
        int ci = first + i;
        GlomeNode c = d->nodes[ci];
        gwave::Stk ct = rtex, cg = rtag;
        int cj = ci;
        while (c.type == GLOME_TEX || c.type == GLOME_TAG) {
            if (!(c.type == GLOME_TEX ? stk_push(ct, c.b) : stk_push(cg, c.b))) return false;
            cj = c.a;
            c = d->nodes[cj];
        }
        gwave::Seg sg;
        memset(&sg, 0, sizeof(sg));
        if (c.type == GLOME_BIH) { sg.kind = gwave::SEG_BIH; sg.node = cj; sg.count = 1; sg.tex = ct; sg.tag = cg; segs.push_back(sg); }
        else if (c.type == GLOME_MESH) { sg.kind = gwave::SEG_MESH; sg.node = cj; sg.count = 1; sg.tex = ct; sg.tag = cg; segs.push_back(sg); }
        else if (c.type == GLOME_VOID || (c.type >= GLOME_SPHERE && c.type <= GLOME_CONE)) {
            if (!segs.empty() && segs.back().kind == gwave::SEG_PRIMS && segs.back().node + segs.back().count == ci) segs.back().count++;
            else { sg.kind = gwave::SEG_PRIMS; sg.node = ci; sg.count = 1; sg.tex = rtex; sg.tag = rtag; segs.push_back(sg); }
        } else return false;
    }
    return !segs.empty() && segs.size() <= GW_MAX_SEGS;
}

static int validate(const GlomeFlatScene* d) {
    if (!d || d->version != GLOME_FLAT_VERSION) { g_err = "FlatScene: bad version"; return GLOME_EINVAL; }
    if (d->n_nodes <= 0 || d->root < 0 || d->root >= d->n_nodes) { g_err = "FlatScene: bad root"; return GLOME_EINVAL; }
    for (int i = 0; i < d->n_nodes; i++) {
        const GlomeNode& n = d->nodes[i];
        if (n.type < 0 || n.type >= GLOME_NODE_TYPE_COUNT) { g_err = "FlatScene: bad node type"; return GLOME_EINVAL; }
        bool bad = false;
        switch (n.type) {
            case GLOME_GROUP: case GLOME_INTERSECTION: bad = n.b < 0 || n.a < 0 || (long long)n.a + n.b > d->n_nodes; break;
            case GLOME_INSTANCE: bad = n.a < 0 || n.a >= d->n_nodes || n.b < 0 || (long long)n.b + 24 > d->n_dpool; break;
            case GLOME_DIFFERENCE: case GLOME_BOUND: case GLOME_INNERBOUND:
                bad = n.a < 0 || n.a >= d->n_nodes || n.b < 0 || n.b >= d->n_nodes; break;
            case GLOME_TEX: bad = n.a < 0 || n.a >= d->n_nodes || n.b < 0 || n.b >= d->n_textures; break;
            case GLOME_TAG: case GLOME_NOSHADOW: case GLOME_ONLYSHADOW: bad = n.a < 0 || n.a >= d->n_nodes; break;
            case GLOME_BIH: bad = n.b < 0 || (long long)n.b + 6 > d->n_dpool || (n.b & 1) || (n.a >= 0 && n.a >= d->n_bihnodes); break;
            case GLOME_MESH: {
                bad = n.a < 0 || (long long)n.a + 12 > d->n_ipool;
                if (!bad) {
                    const GlomeMeshHeader* h = reinterpret_cast<const GlomeMeshHeader*>(d->ipool + n.a);
                    bad = h->bb_off < 0 || (long long)h->bb_off + 6 > d->n_dpool || (h->bb_off & 1) || h->ntris < 0 || h->nverts < 0 ||
                          h->verts_off < 0 || (long long)h->verts_off + 3LL * h->nverts > d->n_dpool ||
                          h->norms_off < 0 || (long long)h->norms_off + 3LL * h->nnorms > d->n_dpool ||
                          h->tris_off < 0 || (h->tris_off & 3) || (long long)h->tris_off + 8LL * h->ntris > d->n_ipool ||
                          h->texs_off < 0 || (long long)h->texs_off + h->ntexs > d->n_ipool ||
                          h->tags_off < 0 || (long long)h->tags_off + h->ntags > d->n_ipool ||
                          (h->root >= 0 && h->root >= d->n_bvhnodes) || (h->root < 0 && ~h->root >= d->n_ipool);
                }
                break;
            }
            case GLOME_VOID: break;
            case GLOME_SPHERE: bad = n.a < 0 || (long long)n.a + 4 > d->n_dpool || (n.a & 1); break;  // two 16-byte loads
            case GLOME_TRIANGLE: bad = n.a < 0 || (long long)n.a + 9 > d->n_dpool; break;
            case GLOME_TRIANGLENORM: bad = n.a < 0 || (long long)n.a + 18 > d->n_dpool; break;
            case GLOME_BOX: bad = n.a < 0 || (long long)n.a + 6 > d->n_dpool || (n.a & 1); break;
            case GLOME_PLANE: case GLOME_CONE: bad = n.a < 0 || (long long)n.a + 4 > d->n_dpool; break;
            case GLOME_DISC: bad = n.a < 0 || (long long)n.a + 7 > d->n_dpool; break;
            case GLOME_CYLINDER: bad = n.a < 0 || (long long)n.a + 3 > d->n_dpool; break;
            default: bad = n.a < 0 || n.a >= d->n_dpool; break;
        }
        if (bad) { g_err = "FlatScene: node " + std::to_string(i) + " has out-of-range payload"; return GLOME_EINVAL; }
    }
    // BIH / BVH child refs, leaf ranges
    for (int i = 0; i < d->n_bihnodes; i++) {
        const GlomeBihNode& b = d->bihnodes[i];
        if (b.axis < 0 || b.axis > 2) { g_err = "FlatScene: BIH node " + std::to_string(i) + " has a bad axis"; return GLOME_EINVAL; }
        for (int c = 0; c < 2; c++) {
            int32_t ref = c ? b.right : b.left;
            bool bad;
            if (ref >= 0) bad = ref >= d->n_bihnodes;
            else {
                int32_t k = ~ref;
                if ((k & 7) != 7) bad = (long long)(k >> 3) + (k & 7) > d->n_nodes;
                else bad = (long long)(k >> 3) + 2 > d->n_ipool || d->ipool[k >> 3] < 0 || d->ipool[(k >> 3) + 1] < 0 ||
                           (long long)d->ipool[k >> 3] + d->ipool[(k >> 3) + 1] > d->n_nodes;
            }
            if (bad) { g_err = "FlatScene: BIH node " + std::to_string(i) + " has an out-of-range child"; return GLOME_EINVAL; }
        }
    }
    for (int i = 0; i < d->n_bvhnodes; i++) {
        const GlomeBvhNode& b = d->bvhnodes[i];
        for (int c = 0; c < 2; c++) {
            int32_t ref = c ? b.right : b.left;
            bool bad = ref >= 0 ? ref >= d->n_bvhnodes : (~ref >= d->n_ipool || (long long)~ref + 1 + d->ipool[~ref] > d->n_ipool);
            if (bad) { g_err = "FlatScene: BVH node " + std::to_string(i) + " has an out-of-range child"; return GLOME_EINVAL; }
        }
    }
    for (int i = 0; i < d->n_textures; i++) {
        const GlomeTexture& t = d->textures[i];
        bool bad = t.kind < GLOME_TEX_UNIFORM || t.kind > GLOME_TEX_PERLIN_BLEND || t.a < 0 || t.a >= d->n_materials ||
                   (t.kind != GLOME_TEX_UNIFORM && (t.b < 0 || t.b >= d->n_materials));
        if (bad) { g_err = "FlatScene: texture " + std::to_string(i) + " names a bad material"; return GLOME_EINVAL; }
    }
    for (int i = 0; i < d->n_materials; i++) {
        const GlomeMaterial& m = d->materials[i];
        bool bad = false;
        switch (m.kind) {
            case GLOME_MAT_SURFACE: case GLOME_MAT_REFLECT: case GLOME_MAT_REFRACT: break;
            case GLOME_MAT_WARP:
                bad = m.a < 0 || m.a >= d->n_nodes || m.b < 0 || m.b >= d->n_nodes || m.c < 0 ||
                      m.c >= (d->n_lightsets > 0 ? d->n_lightsets : 1) || m.d < 0 || (long long)m.d + 24 > d->n_dpool;
                break;
            case GLOME_MAT_ADDITIVE:
                bad = m.b < 0 || m.a < 0 || (long long)m.a + m.b > d->n_ipool;
                for (int k = 0; !bad && k < m.b; k++) bad = d->ipool[m.a + k] < 0 || d->ipool[m.a + k] >= d->n_materials;
                break;
            case GLOME_MAT_BLEND: bad = m.a < 0 || m.a >= d->n_materials || m.b < 0 || m.b >= d->n_materials; break;
            default: bad = true;
        }
        if (bad) { g_err = "FlatScene: material " + std::to_string(i) + " is malformed"; return GLOME_EINVAL; }
    }
    for (int i = 0; i < d->n_lights; i++)
        if (d->lights[i].falloff != 0) { g_err = "FlatScene: unknown light falloff (only 1/(x*x), Shader.hs:23)"; return GLOME_EINVAL; }
    for (int i = 0; i < d->n_lightsets; i++) {
        if (d->lightsets[2 * i] < 0 || d->lightsets[2 * i + 1] < 0 || d->lightsets[2 * i] + d->lightsets[2 * i + 1] > d->n_lights) {
            g_err = "FlatScene: bad light set"; return GLOME_EINVAL;
        }
        if (d->lightsets[2 * i + 1] > GDEV_MAX_LIGHTS) { g_err = "FlatScene: more than 64 lights in one light set"; return GLOME_ELIMIT; }
    }
    if (d->n_lightsets <= 0 && d->n_lights > GDEV_MAX_LIGHTS) { g_err = "FlatScene: more than 64 lights"; return GLOME_ELIMIT; }
    return GLOME_OK;
}

template <typename K>
static int persistent_grid(GlomeScene* s, K kernel, int threads) {
    int b = 0;
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&b, kernel, threads, 0) != cudaSuccess || b < 1) b = 1;
    return s->sm_count * b;
}
static int env_int(const char* name, int dflt) {
    const char* e = getenv(name);
    return e ? atoi(e) : dflt;
}

static int scene_create_impl(GlomeScene* s, const GlomeFlatScene* desc, int device);
extern "C" int GLOME_API(glome_scene_destroy)(GlomeScene* s);

extern "C" int GLOME_API(glome_scene_create)(const GlomeFlatScene* desc, int device, GlomeScene** out) {
    if (!out) { g_err = "null out"; return GLOME_EINVAL; }
    *out = nullptr;
    int rc = validate(desc);
    if (rc) return rc;
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) {
        g_err = "no CUDA device: libglomecuda has no CPU fallback";
        return GLOME_ENODEV;
    }
    if (device < 0 || device >= ndev) { g_err = "bad device index"; return GLOME_EINVAL; }
    CK(cudaSetDevice(device));
    GlomeScene* s = nullptr;
    try {
        s = new GlomeScene();  // value-initialised: every pointer starts null, so a partial scene can be destroyed
        rc = scene_create_impl(s, desc, device);
    } catch (const std::exception& e) {
        g_err = std::string("glome_scene_create: ") + e.what();
        rc = GLOME_ECUDA;
    } catch (...) {
        g_err = "glome_scene_create: unknown exception";
        rc = GLOME_ECUDA;
    }
    if (rc) {
        std::string keep = g_err;
        if (s) GLOME_API(glome_scene_destroy)(s);
        g_err = keep;
        return rc;
    }
    *out = s;
    return GLOME_OK;
}

static int scene_create_impl(GlomeScene* s, const GlomeFlatScene* desc, int device) {
    int rc;
    s->precision = GLOME_PRECISION;
    s->device = device;
    s->scene_class = desc->scene_class;
    cudaDeviceProp prop;
    CK(cudaGetDeviceProperties(&prop, device));
    s->sm_count = prop.multiProcessorCount;
    memset(&s->d, 0, sizeof(s->d));
    s->d.root = desc->root;
    s->d.n_lights = desc->n_lights;
    std::vector<int32_t> ls;
    if (desc->n_lightsets > 0) ls.assign(desc->lightsets, desc->lightsets + 2 * desc->n_lightsets);
    else { ls.push_back(0); ls.push_back(desc->n_lights); }
    std::vector<GlomeBihNode> bihv;
    std::vector<double> dpv;
    const GlomeBihNode* bih_src = desc->bihnodes;
    const double* dp_src = desc->dpool;
    size_t bih_n = (size_t)desc->n_bihnodes, dp_n = (size_t)desc->n_dpool;
    s->env_group_accel = env_int("GLOME_GROUP_ACCEL", 1);
    if (s->scene_class == GLOME_CLASS_GENERAL) {
        // the scene-graph machine packs tag ids into 16 bits: upload dense ids and the table that maps them back
        std::vector<GlomeNode> nodes;
        std::vector<int32_t> ipool, tagvals;
        std::string e = glome_tagmap::remap_tags(desc, nodes, ipool, tagvals);
        if (!e.empty()) { g_err = e; return GLOME_ELIMIT; }
        std::vector<int32_t> items;
        glome_tagmap::build_items(nodes, items);
        if (s->env_group_accel) {  // implicit BIHs over large plain groups: more nodes, items, BIH nodes and two pool records
            bihv.assign(desc->bihnodes, desc->bihnodes + desc->n_bihnodes);
            dpv.assign(desc->dpool, desc->dpool + desc->n_dpool);
            glome_tagmap::build_group_accels(nodes, items, ipool, bihv, dpv);
            bih_src = bihv.data(); bih_n = bihv.size(); dp_src = dpv.data(); dp_n = dpv.size();
        }
        if ((rc = upload(s, nodes.data(), nodes.size(), &s->d.nodes))) return rc;
        if ((rc = upload(s, ipool.data(), ipool.size(), &s->d.ipool))) return rc;
        if ((rc = upload(s, tagvals.data(), tagvals.size(), &s->d.tagvals))) return rc;
        const int32_t* items_dev = nullptr;
        if ((rc = upload(s, items.data(), items.size(), &items_dev))) return rc;
        s->d.items = reinterpret_cast<const int4*>(items_dev);
    } else {
        if ((rc = upload(s, desc->nodes, (size_t)desc->n_nodes, &s->d.nodes))) return rc;
        if ((rc = upload(s, desc->ipool, (size_t)desc->n_ipool, &s->d.ipool))) return rc;
    }
#ifdef GLOME_F32
    {   // FP32 payloads: round every double once, here; refs and indices are unchanged
        std::vector<DBihNode> bih(bih_n);
        for (size_t i = 0; i < bih_n; i++) {
            const GlomeBihNode& b = bih_src[i];
            if (b.right > 0x1fffffff || b.right < -0x20000000) { g_err = "FP32 mode: a BIH child ref needs more than 30 bits"; return GLOME_ELIMIT; }
            bih[i].lsplit = (float)b.lsplit; bih[i].rsplit = (float)b.rsplit; bih[i].left = b.left;
            bih[i].right_axis = (int32_t)(((uint32_t)b.right << 2) | (uint32_t)b.axis);
        }
        std::vector<DBvhNode> bvh((size_t)desc->n_bvhnodes);
        for (int i = 0; i < desc->n_bvhnodes; i++) {
            const GlomeBvhNode& b = desc->bvhnodes[i];
            for (int k = 0; k < 6; k++) { bvh[i].lbb[k] = (float)b.lbb[k]; bvh[i].rbb[k] = (float)b.rbb[k]; }
            bvh[i].left = b.left; bvh[i].right = b.right; bvh[i].pad[0] = bvh[i].pad[1] = 0;
        }
        std::vector<float> dp(dp_n);
        for (size_t i = 0; i < dp_n; i++) dp[i] = (float)dp_src[i];
        if ((rc = upload(s, bih.data(), bih.size(), &s->d.bih))) return rc;
        if ((rc = upload(s, bvh.data(), bvh.size(), &s->d.bvh))) return rc;
        if ((rc = upload(s, dp.data(), dp.size(), &s->d.dpool))) return rc;
    }
#else
    if ((rc = upload(s, bih_src, bih_n, &s->d.bih))) return rc;
    if ((rc = upload(s, desc->bvhnodes, (size_t)desc->n_bvhnodes, &s->d.bvh))) return rc;
    if ((rc = upload(s, dp_src, dp_n, &s->d.dpool))) return rc;
#endif
    if (env_int("GLOME_LEAF_INLINE", 1)) {
        // inline leaf vertices of every Mesh (DScene::leafv).  The leaf pools are walked from each Mesh's BVH; the array spans
        // the ipool range that holds them (one Mesh: exactly its leaf pool).
        long long lo = -1, hi = -1;
        std::vector<std::pair<int32_t, int32_t>> meshes;  // (header offset, root)
        for (int i = 0; i < desc->n_nodes; i++)
            if (desc->nodes[i].type == GLOME_MESH) {
                const GlomeMeshHeader* h = reinterpret_cast<const GlomeMeshHeader*>(desc->ipool + desc->nodes[i].a);
                meshes.push_back(std::make_pair(desc->nodes[i].a, h->root));
            }
        std::sort(meshes.begin(), meshes.end());
        meshes.erase(std::unique(meshes.begin(), meshes.end()), meshes.end());
        std::vector<std::pair<int32_t, int32_t>> leaves;  // (ipool offset k of {count, tri...}, header offset)
        for (size_t m = 0; m < meshes.size(); m++) {
            std::vector<int32_t> stk(1, meshes[m].second);
            while (!stk.empty()) {
                const int32_t r = stk.back(); stk.pop_back();
                if (r < 0) { leaves.push_back(std::make_pair(~r, meshes[m].first)); continue; }
                stk.push_back(desc->bvhnodes[r].left);
                stk.push_back(desc->bvhnodes[r].right);
            }
        }
        for (size_t i = 0; i < leaves.size(); i++) {
            const long long k = leaves[i].first, n = desc->ipool[k];
            if (n <= 0) continue;
            if (lo < 0 || k + 1 < lo) lo = k + 1;
            if (k + n > hi) hi = k + n;
        }
        if (lo >= 0 && (hi - lo + 1) * 9LL * (long long)sizeof(Flt) <= (4LL << 30)) {
            std::vector<Flt> lv((size_t)(hi - lo + 1) * 9, (Flt)0);
            for (size_t i = 0; i < leaves.size(); i++) {
                const long long k = leaves[i].first, n = desc->ipool[k];
                const GlomeMeshHeader* h = reinterpret_cast<const GlomeMeshHeader*>(desc->ipool + leaves[i].second);
                for (long long j = 0; j < n; j++) {
                    const int32_t ti = desc->ipool[k + 1 + j];
                    const int32_t* t = desc->ipool + h->tris_off + 8 * (long long)ti;
                    Flt* o = lv.data() + (size_t)(k + 1 + j - lo) * 9;
                    for (int v = 0; v < 3; v++)
                        for (int c = 0; c < 3; c++) o[3 * v + c] = (Flt)desc->dpool[h->verts_off + 3LL * t[v] + c];
                }
            }
            if ((rc = upload(s, lv.data(), lv.size(), &s->d.leafv))) return rc;
            s->d.leafv_base = (int)lo;
        }
    }
    if ((rc = upload(s, desc->textures, (size_t)desc->n_textures, &s->d.textures))) return rc;
    if ((rc = upload(s, desc->materials, (size_t)desc->n_materials, &s->d.materials))) return rc;
    if ((rc = upload(s, desc->lights, (size_t)desc->n_lights, &s->d.lights))) return rc;
    if ((rc = upload(s, ls.data(), ls.size(), &s->d.lightsets))) return rc;
    CK(cudaMalloc((void**)&s->queue_count, sizeof(int)));
    CK(cudaMalloc((void**)&s->work_counter, sizeof(unsigned int)));
    CK(cudaMalloc((void**)&s->stats, sizeof(DevStats)));
    CK(cudaEventCreate(&s->ev0));
    CK(cudaEventCreate(&s->ev1));
    s->n_scene_lights = ls[1];
    s->env_no_speculate = env_int("GLOME_NO_SPECULATE", 0);
    s->env_aa_speculate = env_int("GLOME_AA_SPECULATE", -1);
    s->env_gen_chunk = env_int("GLOME_GEN_CHUNK", 32);
    s->env_seg_concurrent = env_int("GLOME_SEG_CONCURRENT", 1);
    if (s->env_gen_chunk < 1) s->env_gen_chunk = 1;
    if (s->env_gen_chunk > 32) s->env_gen_chunk = 32;
    // launch geometry of the persistent kernels: once per scene, so that frames issued from several host threads
    // (glome_multi_render) never race on a lazily initialised cache
    s->g_bih[0] = persistent_grid(s, gwave::k_bih_traverse<false, false, false>, GW_THREADS);
    s->g_bih[1] = persistent_grid(s, gwave::k_bih_traverse<false, true, false>, GW_THREADS);
    s->g_bih[2] = persistent_grid(s, gwave::k_bih_traverse<true, false, false>, GW_THREADS);
    s->g_bih[3] = persistent_grid(s, gwave::k_bih_traverse<true, true, false>, GW_THREADS);
    s->g_bvh = persistent_grid(s, gwave::k_bvh_closest, 128);
    s->g_gen[0] = persistent_grid(s, k_gen_trace<0>, GEN_THREADS);
    s->g_gen[1] = persistent_grid(s, k_gen_trace<1>, GEN_THREADS);
    s->g_gen[2] = persistent_grid(s, k_gen_trace<5>, GEN_THREADS);
    {   // A/B: GLOME_GEN_GRID_DIV=2 leaves one block of the general tracer per SM
        const int gd = env_int("GLOME_GEN_GRID_DIV", 1);
        if (gd > 1) for (int k = 0; k < 3; k++) s->g_gen[k] = std::max(s->sm_count, s->g_gen[k] / gd);
    }
    s->g_flat[0] = persistent_grid(s, k_trace_samples<0>, 128);
    s->g_flat[1] = persistent_grid(s, k_trace_samples<1>, 128);
    s->g_flat[2] = persistent_grid(s, k_trace_samples<5>, 128);
    if (s->scene_class == GLOME_CLASS_FLAT && !getenv("GLOME_FLAT_MEGAKERNEL") && build_segments(desc, s->segs) &&
        ls[0] + ls[1] <= 32) {
        CK(cudaMalloc((void**)&s->segs_dev, sizeof(gwave::Seg) * s->segs.size()));
        CK(cudaMemcpy(s->segs_dev, s->segs.data(), sizeof(gwave::Seg) * s->segs.size(), cudaMemcpyHostToDevice));
        CK(cudaMalloc((void**)&s->w_squeue_count, sizeof(int)));
        CK(cudaMalloc((void**)&s->w_counters, sizeof(unsigned int) * 1024));
        for (size_t i = 0; i < s->segs.size(); i++)
            s->segs_linear.push_back(s->segs[i].kind == gwave::SEG_BIH &&
                                     (desc->nodes[s->segs[i].node].c & GLOME_BIH_LINEAR_SPHERES) ? 1 : 0);
        s->env_bih_reps = env_int("GLOME_BIH_BOUNDED", -1);
        for (size_t i = 0; i < s->segs.size(); i++) {
            // the branch-phase bound pays in deep trees: count the Bih's nodes (walk from its root)
            int reps = 0;  // 0: unbounded branch phase
            if (s->segs[i].kind == gwave::SEG_BIH) {
                long long cnt = 0;
                std::vector<int32_t> stk;
                const int32_t root = desc->nodes[s->segs[i].node].a;
                if (root >= 0) stk.push_back(root);
                while (!stk.empty() && cnt <= 200000) {
                    const int32_t r = stk.back(); stk.pop_back(); cnt++;
                    if (desc->bihnodes[r].left >= 0) stk.push_back(desc->bihnodes[r].left);
                    if (desc->bihnodes[r].right >= 0) stk.push_back(desc->bihnodes[r].right);
                }
                if (cnt > 200000) reps = 1;
                if (s->env_bih_reps >= 0) reps = s->env_bih_reps;  // A/B: 0 = never bounded, 1 = always
            }
            s->segs_reps.push_back(reps);
        }
        s->use_wave = true;
    }
    return GLOME_OK;
}

extern "C" int GLOME_API(glome_scene_destroy)(GlomeScene* s) {
    if (!s) return GLOME_OK;
    TWIN(glome_scene_destroy_f32(s));
    cudaSetDevice(s->device);
    for (void* p : s->bufs) cudaFree(p);
    cudaFree(s->v); cudaFree(s->v2); cudaFree(s->rgb8); cudaFree(s->queue); cudaFree(s->spec);
    cudaFree(s->queue_count); cudaFree(s->work_counter); cudaFree(s->stats);
    for (int i = 0; i < 4; i++) cudaFree(s->bw[i]);
    cudaFree(s->segs_dev); cudaFree(s->w_hit_t); cudaFree(s->w_hit_seg); cudaFree(s->w_hit_item); cudaFree(s->w_hit_sub);
    cudaFree(s->w_hit_flags); cudaFree(s->w_surf); cudaFree(s->w_occl); cudaFree(s->w_squeue); cudaFree(s->w_squeue_count);
    cudaFree(s->w_counters);
    cudaFree(s->w2_hit_t); cudaFree(s->w2_hit_seg); cudaFree(s->w2_hit_item); cudaFree(s->w2_hit_sub); cudaFree(s->w2_hit_flags);
    if (s->st2) cudaStreamDestroy(s->st2);
    if (s->ev_fork) cudaEventDestroy(s->ev_fork);
    if (s->ev_join) cudaEventDestroy(s->ev_join);
    if (s->aa_counts_host) cudaFreeHost(s->aa_counts_host);
    if (s->aa_ev) cudaEventDestroy(s->aa_ev);
    for (int k = 0; k < 2; k++) { if (s->aa_t0[k]) cudaEventDestroy(s->aa_t0[k]); if (s->aa_t1[k]) cudaEventDestroy(s->aa_t1[k]); }
    if (s->ev0) cudaEventDestroy(s->ev0);
    if (s->ev1) cudaEventDestroy(s->ev1);
    for (cudaEvent_t e : s->tev) cudaEventDestroy(e);
    delete s;
    return GLOME_OK;
}

static int grow(GlomeScene* s, int slot, size_t bytes) {
    if (s->bw_cap[slot] >= bytes) return GLOME_OK;
    if (s->bw[slot]) cudaFree(s->bw[slot]);
    s->bw[slot] = nullptr; s->bw_cap[slot] = 0;
    CK(cudaMalloc(&s->bw[slot], bytes));
    s->bw_cap[slot] = bytes;
    return GLOME_OK;
}
static int batch_grid(GlomeScene* s, long long n) {
    long long blocks = (n + 127) / 128;
    long long cap = (long long)s->sm_count * 16;
    if (blocks > cap) blocks = cap;
    if (blocks < 1) blocks = 1;
    return (int)blocks;
}

extern "C" int GLOME_API(glome_rayint_batch)(GlomeScene* s, int64_t n, const double* rays, const double* tmax, int tmax_stride,
                                  GlomeHit* out) {
    TWIN(glome_rayint_batch_f32(s, n, rays, tmax, tmax_stride, out));
    if (!s || n < 0 || !rays || !tmax || !out) { g_err = "bad argument"; return GLOME_EINVAL; }
    if (n == 0) return GLOME_OK;
    CK(cudaSetDevice(s->device));
    int rc;
    size_t nt = tmax_stride ? (size_t)n : 1;
    if ((rc = grow(s, 0, (size_t)n * 48))) return rc;
    if ((rc = grow(s, 1, nt * 8))) return rc;
    if ((rc = grow(s, 2, (size_t)n * sizeof(GlomeHit)))) return rc;
    CK(cudaMemcpy(s->bw[0], rays, (size_t)n * 48, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(s->bw[1], tmax, nt * 8, cudaMemcpyHostToDevice));
    int grid = batch_grid(s, n);
    if (s->scene_class == GLOME_CLASS_FLAT)
        k_rayint_batch<false><<<grid, 128>>>(s->d, n, (const double*)s->bw[0], (const double*)s->bw[1], tmax_stride, (GlomeHit*)s->bw[2]);
    else
        k_rayint_batch<true><<<grid, 128>>>(s->d, n, (const double*)s->bw[0], (const double*)s->bw[1], tmax_stride, (GlomeHit*)s->bw[2]);
    s->launches++;
    CK(cudaGetLastError());
    CK(cudaMemcpy(out, s->bw[2], (size_t)n * sizeof(GlomeHit), cudaMemcpyDeviceToHost));
    return GLOME_OK;
}

extern "C" int GLOME_API(glome_shadow_batch)(GlomeScene* s, int64_t n, const double* rays, const double* tmax, int tmax_stride,
                                  uint8_t* occluded) {
    TWIN(glome_shadow_batch_f32(s, n, rays, tmax, tmax_stride, occluded));
    if (!s || n < 0 || !rays || !tmax || !occluded) { g_err = "bad argument"; return GLOME_EINVAL; }
    if (n == 0) return GLOME_OK;
    CK(cudaSetDevice(s->device));
    int rc;
    size_t nt = tmax_stride ? (size_t)n : 1;
    if ((rc = grow(s, 0, (size_t)n * 48))) return rc;
    if ((rc = grow(s, 1, nt * 8))) return rc;
    if ((rc = grow(s, 2, (size_t)n))) return rc;
    CK(cudaMemcpy(s->bw[0], rays, (size_t)n * 48, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(s->bw[1], tmax, nt * 8, cudaMemcpyHostToDevice));
    int grid = batch_grid(s, n);
    if (s->scene_class == GLOME_CLASS_FLAT)
        k_shadow_batch<false><<<grid, 128>>>(s->d, n, (const double*)s->bw[0], (const double*)s->bw[1], tmax_stride, (uint8_t*)s->bw[2]);
    else
        k_shadow_batch<true><<<grid, 128>>>(s->d, n, (const double*)s->bw[0], (const double*)s->bw[1], tmax_stride, (uint8_t*)s->bw[2]);
    s->launches++;
    CK(cudaGetLastError());
    CK(cudaMemcpy(occluded, s->bw[2], (size_t)n, cudaMemcpyDeviceToHost));
    return GLOME_OK;
}

extern "C" int GLOME_API(glome_inside_batch)(GlomeScene* s, int64_t n, const double* pts, uint8_t* inside) {
    TWIN(glome_inside_batch_f32(s, n, pts, inside));
    if (!s || n < 0 || !pts || !inside) { g_err = "bad argument"; return GLOME_EINVAL; }
    if (n == 0) return GLOME_OK;
    CK(cudaSetDevice(s->device));
    int rc;
    if ((rc = grow(s, 0, (size_t)n * 24))) return rc;
    if ((rc = grow(s, 2, (size_t)n))) return rc;
    CK(cudaMemcpy(s->bw[0], pts, (size_t)n * 24, cudaMemcpyHostToDevice));
    k_inside_batch<<<batch_grid(s, n), 128>>>(s->d, n, (const double*)s->bw[0], (uint8_t*)s->bw[2]);
    s->launches++;
    CK(cudaGetLastError());
    CK(cudaMemcpy(inside, s->bw[2], (size_t)n, cudaMemcpyDeviceToHost));
    return GLOME_OK;
}

static void read_stats(GlomeScene* s, GlomeRenderStats* out, float ms, int launches) {
    DevStats h;
    memset(&h, 0, sizeof(h));
    cudaMemcpy(&h, s->stats, sizeof(h), cudaMemcpyDeviceToHost);
    if (!out) return;
    out->rays_primary = (int64_t)h.primary;
    out->rays_shadow = (int64_t)h.shadow;
    out->rays_secondary = (int64_t)h.secondary;
    out->overflow_rays = (int64_t)h.overflow;
    out->perlin_range = (int64_t)h.perlin_range;
    out->kernel_ms = ms;
    out->launches = launches;
    out->reserved = 0;
    out->traverse_ms = 0;
    out->traverse_launches = 0;
    out->visits_bih = (int64_t)h.bih;
    out->tests_prim = (int64_t)h.prim;
    out->visits_bvh = (int64_t)h.bvh;
    out->tests_tri = (int64_t)h.tri;
    for (int k = 0; k < 4; k++) { out->family_ms[k] = 0; out->family_launches[k] = 0; }
    out->visits_instance = (int64_t)h.inst;
    out->csg_steps = (int64_t)h.csg;
}

extern "C" int GLOME_API(glome_trace_batch)(GlomeScene* s, int64_t n, const double* rays, const double* tmax, int tmax_stride, int recurs,
                                 double* rgba, double* depth, GlomeHit* hits) {
    TWIN(glome_trace_batch_f32(s, n, rays, tmax, tmax_stride, recurs, rgba, depth, hits));
    if (!s || n < 0 || !rays || !tmax || !rgba || !depth) { g_err = "bad argument"; return GLOME_EINVAL; }
    if (n == 0) return GLOME_OK;
    CK(cudaSetDevice(s->device));
    int rc;
    size_t nt = tmax_stride ? (size_t)n : 1;
    if ((rc = grow(s, 0, (size_t)n * 48))) return rc;
    if ((rc = grow(s, 1, nt * 8))) return rc;
    if ((rc = grow(s, 2, (size_t)n * 40))) return rc;
    if (hits && (rc = grow(s, 3, (size_t)n * sizeof(GlomeHit)))) return rc;
    CK(cudaMemcpy(s->bw[0], rays, (size_t)n * 48, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(s->bw[1], tmax, nt * 8, cudaMemcpyHostToDevice));
    CK(cudaMemset(s->stats, 0, sizeof(DevStats)));
    double* d_rgba = (double*)s->bw[2];
    double* d_depth = d_rgba + 4 * n;
    int grid = batch_grid(s, n);
    if (s->scene_class == GLOME_CLASS_FLAT)
        k_trace_batch<false><<<grid, 128>>>(s->d, n, (const double*)s->bw[0], (const double*)s->bw[1], tmax_stride, recurs, d_rgba,
                                            d_depth, hits ? (GlomeHit*)s->bw[3] : nullptr, s->stats);
    else
        k_trace_batch<true><<<2 * grid, 64>>>(s->d, n, (const double*)s->bw[0], (const double*)s->bw[1], tmax_stride, recurs, d_rgba,
                                           d_depth, hits ? (GlomeHit*)s->bw[3] : nullptr, s->stats);
    s->launches++;
    CK(cudaGetLastError());
    CK(cudaMemcpy(rgba, d_rgba, (size_t)n * 32, cudaMemcpyDeviceToHost));
    CK(cudaMemcpy(depth, d_depth, (size_t)n * 8, cudaMemcpyDeviceToHost));
    if (hits) CK(cudaMemcpy(hits, s->bw[3], (size_t)n * sizeof(GlomeHit), cudaMemcpyDeviceToHost));
    return GLOME_OK;
}

// getTags' / get_tags (Glome.hs:69-72, 410-414): the tag list of whatever is under pixel (px, py) -- what
// GlomeView prints on a mouse click.  One camera ray (getCoords, Glome.hs:119-128; get_rayint, :27-33) through
// trace instantiated with a tag list: `ts ++ tags`, the tags gathered by Reflect / Refract / Warp recursion
// followed by the hit's own tag stack (Trace.hs:82).
extern "C" int GLOME_API(glome_get_tags)(GlomeScene* s, const GlomeCamera* cam, int width, int height, int px, int py, int recurs,
                              int32_t* tags, int max_tags, int* ntags, int* truncated, GlomeHit* hit_out) {
    TWIN(glome_get_tags_f32(s, cam, width, height, px, py, recurs, tags, max_tags, ntags, truncated, hit_out));
    if (!s || !cam || !tags || !ntags || width <= 0 || height <= 0 || max_tags < 0) { g_err = "bad argument"; return GLOME_EINVAL; }
    const Flt xf = (Flt)px, yf = (Flt)py, widthf = (Flt)width, heightf = (Flt)height;
    const Flt xc = (((xf / widthf) * 2) - 1) * (widthf / heightf);
    const Flt yc = -(((yf / heightf) * 2) - 1);
    const Vec fwd = vec(cam->fwd[0], cam->fwd[1], cam->fwd[2]), up = vec(cam->up[0], cam->up[1], cam->up[2]);
    const Vec right = vec(cam->right[0], cam->right[1], cam->right[2]);
    const Vec dir = vnorm(vadd3(fwd, vscale(right, -xc), vscale(up, yc)));
    double ray[6] = {cam->pos[0], cam->pos[1], cam->pos[2], dir.x, dir.y, dir.z};
    double tmax = GLM_INFINITY;
    CK(cudaSetDevice(s->device));
    int rc;
    if ((rc = grow(s, 0, 48))) return rc;
    if ((rc = grow(s, 1, 8))) return rc;
    if ((rc = grow(s, 2, (2 + 16) * sizeof(int)))) return rc;
    if ((rc = grow(s, 3, sizeof(GlomeHit)))) return rc;
    CK(cudaMemcpy(s->bw[0], ray, 48, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(s->bw[1], &tmax, 8, cudaMemcpyHostToDevice));
    if (s->scene_class == GLOME_CLASS_FLAT)
        k_trace_tags<false><<<1, 64>>>(s->d, 1, (const double*)s->bw[0], (const double*)s->bw[1], 0, recurs, (GlomeHit*)s->bw[3], (int*)s->bw[2]);
    else
        k_trace_tags<true><<<1, 64>>>(s->d, 1, (const double*)s->bw[0], (const double*)s->bw[1], 0, recurs, (GlomeHit*)s->bw[3], (int*)s->bw[2]);
    s->launches++;
    CK(cudaGetLastError());
    int out[2 + 16];
    GlomeHit h;
    CK(cudaMemcpy(out, s->bw[2], sizeof(out), cudaMemcpyDeviceToHost));
    CK(cudaMemcpy(&h, s->bw[3], sizeof(h), cudaMemcpyDeviceToHost));
    *ntags = out[0];
    for (int i = 0; i < out[0] && i < max_tags; i++) tags[i] = out[2 + i];
    if (truncated) *truncated = (out[1] || out[0] > max_tags) ? 1 : 0;
    if (hit_out) *hit_out = h;
    return GLOME_OK;
}

// rayint_debug's count per ray (Solid.hs:155; Bih.hs:378-412) and get_color_debug's tint (Glome.hs:57-60)
__global__ void k_debug_count_batch(DScene S, long long n, const double* __restrict__ rays, const double* __restrict__ tmax,
                                    int stride, int* __restrict__ out) {
    long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (i >= n) return;
    Ray r = mkray(vec(rays[6 * i], rays[6 * i + 1], rays[6 * i + 2]), vec(rays[6 * i + 3], rays[6 * i + 4], rays[6 * i + 5]));
    ggen::QVM vm;
    ggen::GCnt c;
    ggen::gcnt_clear(c);
    int ovf = 0;
    out[i] = ggen::gq_debug_count(S, vm, S.root, r, tmax[stride ? i : 0], c, &ovf);
}
__global__ void k_debug_tint(DScene S, TileGeom g, DCamera cam, int tile_first, int tile_stride, double* __restrict__ out) {
    long long pix = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (pix >= (long long)g.width * g.height) return;
    int x = (int)(pix % g.width), y = (int)(pix / g.width);
    int ti = (x / g.bs) * g.nty + (y / g.bs);  // renderTiles enumerates tiles x-major (Glome.hs:371-377)
    if (ti % tile_stride != tile_first) return;
    Flt xc, yc;
    getCoordsf(g.width, g.height, (Flt)x, (Flt)y, xc, yc);
    ggen::QVM vm;
    ggen::GCnt c;
    ggen::gcnt_clear(c);
    int ovf = 0;
    int dbg = ggen::gq_debug_count(S, vm, S.root, camera_ray(cam, xc, yc), GLM_INFINITY, c, &ovf);
    out[5 * pix] = ((Flt)(dbg % 30) / 60) + out[5 * pix];
    out[5 * pix + 1] = out[5 * pix + 1] + ((Flt)dbg / 1000);
}
extern "C" int GLOME_API(glome_debug_count_batch)(GlomeScene* s, int64_t n, const double* rays, const double* tmax, int tmax_stride,
                                       int32_t* counts) {
    TWIN(glome_debug_count_batch_f32(s, n, rays, tmax, tmax_stride, counts));
    if (!s || n < 0 || !rays || !tmax || !counts) { g_err = "bad argument"; return GLOME_EINVAL; }
    if (n == 0) return GLOME_OK;
    CK(cudaSetDevice(s->device));
    int rc;
    size_t nt = tmax_stride ? (size_t)n : 1;
    if ((rc = grow(s, 0, (size_t)n * 48))) return rc;
    if ((rc = grow(s, 1, nt * 8))) return rc;
    if ((rc = grow(s, 2, (size_t)n * 4))) return rc;
    CK(cudaMemcpy(s->bw[0], rays, (size_t)n * 48, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(s->bw[1], tmax, nt * 8, cudaMemcpyHostToDevice));
    k_debug_count_batch<<<(unsigned)((n + 63) / 64), 64>>>(s->d, n, (const double*)s->bw[0], (const double*)s->bw[1], tmax_stride,
                                                           (int*)s->bw[2]);
    s->launches++;
    CK(cudaGetLastError());
    CK(cudaMemcpy(counts, s->bw[2], (size_t)n * 4, cudaMemcpyDeviceToHost));
    return GLOME_OK;
}

static void trav_mark(GlomeScene* s, cudaStream_t st, int family);

template <bool GEN, int MODE>
static int launch_trace(GlomeScene* s, const TraceParams& P, cudaStream_t st) {
    const int threads = GEN ? GEN_THREADS : 128;
    const int mi = MODE == 0 ? 0 : (MODE == 1 ? 1 : 2);
    const int grid = GEN ? s->g_gen[mi] : s->g_flat[mi];
    CK(cudaMemsetAsync(P.work_counter, 0, sizeof(unsigned int), st));
    TraceParams Q = P;
    Q.chunk = GEN ? s->env_gen_chunk : 32;
    if (GEN) {
        trav_mark(s, st, 3);
        k_gen_trace<MODE><<<grid, threads, 0, st>>>(s->d, Q);
        trav_mark(s, st, 3);
    } else k_trace_samples<MODE><<<grid, threads, 0, st>>>(s->d, Q);
    s->launches++;
    CK(cudaGetLastError());
    return GLOME_OK;
}
template <int MODE>
static int launch_trace_c(GlomeScene* s, const TraceParams& P, cudaStream_t st) {
    if (s->scene_class == GLOME_CLASS_FLAT) return launch_trace<false, MODE>(s, P, st);
    return launch_trace<true, MODE>(s, P, st);
}

static int wave_reserve(GlomeScene* s, size_t samples) {
    if (s->wave_cap >= samples) return GLOME_OK;
    cudaFree(s->w_hit_t); cudaFree(s->w_hit_seg); cudaFree(s->w_hit_item); cudaFree(s->w_hit_sub); cudaFree(s->w_hit_flags);
    cudaFree(s->w_surf); cudaFree(s->w_occl); cudaFree(s->w_squeue);
    cudaFree(s->w2_hit_t); cudaFree(s->w2_hit_seg); cudaFree(s->w2_hit_item); cudaFree(s->w2_hit_sub); cudaFree(s->w2_hit_flags);
    s->w2_hit_t = nullptr; s->w2_hit_seg = s->w2_hit_item = s->w2_hit_sub = s->w2_hit_flags = nullptr;
    s->wave_cap = 0;
    int nl = s->n_scene_lights > 0 ? s->n_scene_lights : 1;
    CK(cudaMalloc((void**)&s->w_hit_t, samples * sizeof(Flt)));
    CK(cudaMalloc((void**)&s->w_hit_seg, samples * sizeof(int)));
    CK(cudaMalloc((void**)&s->w_hit_item, samples * sizeof(int)));
    CK(cudaMalloc((void**)&s->w_hit_sub, samples * sizeof(int)));
    CK(cudaMalloc((void**)&s->w_hit_flags, samples * sizeof(int)));
    CK(cudaMalloc((void**)&s->w_surf, samples * 6 * sizeof(Flt)));
    CK(cudaMalloc((void**)&s->w_occl, samples * sizeof(unsigned int)));
    CK(cudaMalloc((void**)&s->w_squeue, samples * nl * sizeof(int2)));
    if (s->segs.size() == 2) {
        CK(cudaMalloc((void**)&s->w2_hit_t, samples * sizeof(Flt)));
        CK(cudaMalloc((void**)&s->w2_hit_seg, samples * sizeof(int)));
        CK(cudaMalloc((void**)&s->w2_hit_item, samples * sizeof(int)));
        CK(cudaMalloc((void**)&s->w2_hit_sub, samples * sizeof(int)));
        CK(cudaMalloc((void**)&s->w2_hit_flags, samples * sizeof(int)));
        if (!s->st2) {
            CK(cudaStreamCreateWithFlags(&s->st2, cudaStreamNonBlocking));
            CK(cudaEventCreateWithFlags(&s->ev_fork, cudaEventDisableTiming));
            CK(cudaEventCreateWithFlags(&s->ev_join, cudaEventDisableTiming));
        }
    }
    s->wave_cap = samples;
    return GLOME_OK;
}

static bool event_done(cudaEvent_t e) {
    if (cudaEventQuery(e) == cudaSuccess) return true;
    cudaGetLastError();  // cudaErrorNotReady is not an error here: do not leave it for the next CK()
    return false;
}
static void trav_mark(GlomeScene* s, cudaStream_t st, int family) {
    if (!s->time_traversal) return;
    if (s->tev_used >= (int)s->tev.size()) {
        cudaEvent_t e;
        if (cudaEventCreate(&e) != cudaSuccess) return;
        s->tev.push_back(e);
        s->tev_family.push_back(0);
    }
    s->tev_family[s->tev_used] = family;
    cudaEventRecord(s->tev[s->tev_used++], st);
}

// One trace wave over the sample list described by W (mode / queue): K1 per segment, K2a, K1' per segment, K2b.
static int launch_wave(GlomeScene* s, gwave::WaveParams W, long long max_samples, cudaStream_t st) {
    using namespace gwave;
    const int* g_bih = s->g_bih;
    const int g_bvh = s->g_bvh;
    long long blocks = (max_samples + 127) / 128;
    long long cap = (long long)s->sm_count * 16;
    int sgrid = (int)(blocks < cap ? (blocks < 1 ? 1 : blocks) : cap);
    W.hit_t = s->w_hit_t; W.hit_seg = s->w_hit_seg; W.hit_item = s->w_hit_item; W.hit_sub = s->w_hit_sub;
    W.hit_flags = s->w_hit_flags; W.surf = s->w_surf; W.occl = s->w_occl; W.squeue = s->w_squeue;
    W.squeue_count = s->w_squeue_count; W.stats = (unsigned long long*)s->stats;
    CK(cudaMemsetAsync(s->w_squeue_count, 0, sizeof(int), st));
    // Two traversal segments (a Bih next to a Mesh -- configs[2] / configs[4]): their closest-hit walks are independent
    // apart from the fold that joins them (Solid.hs:327-328: nearest, ties to the later element), so they run side by side
    // on two streams, each into its own result set, and k_merge_hits applies the fold.  A small wave sits on the latency
    // floor of its slowest ray; side by side the two floors overlap instead of adding up.
    const bool side_by_side = s->w2_hit_t && s->env_seg_concurrent && s->segs.size() == 2 && s->segs[0].kind != SEG_PRIMS && s->segs[1].kind != SEG_PRIMS;
    for (size_t i = 0; i < s->segs.size(); i++) {
        const Seg& sg = s->segs[i];
        cudaStream_t st_outer = st;
        gwave::WaveParams W_outer = W;
        const int segidx_outer = (int)i;
        if (side_by_side && i == 0) {  // fork before the first segment is enqueued
            CK(cudaEventRecord(s->ev_fork, st_outer));
            CK(cudaStreamWaitEvent(s->st2, s->ev_fork, 0));
        }
        // (inside this iteration `st`, `W` and the segment index are those of the lane the segment runs in)
        cudaStream_t st = (side_by_side && i == 1) ? s->st2 : st_outer;
        gwave::WaveParams W = W_outer;
        int segidx = segidx_outer;
        if (side_by_side && i == 1) {
            W.hit_t = s->w2_hit_t; W.hit_seg = s->w2_hit_seg; W.hit_item = s->w2_hit_item; W.hit_sub = s->w2_hit_sub; W.hit_flags = s->w2_hit_flags;
            segidx = 0;  // a fold of its own: every sample gets a record, hit_seg is 0 / -1 (k_merge_hits renames it)
        }
        if (sg.kind == SEG_BIH) {
            bool linear = (s->segs_linear[i] != 0);
            unsigned int* ctr = s->w_counters + (s->w_counter_next++ & 1023);
            trav_mark(s, st, 0);
            const bool bounded = s->segs_reps[i] != 0;
            if (linear && bounded) k_bih_traverse<false, true, true><<<g_bih[1], GW_THREADS, 0, st>>>(s->d, W, segidx, sg, ctr);
            else if (linear) k_bih_traverse<false, true, false><<<g_bih[1], GW_THREADS, 0, st>>>(s->d, W, segidx, sg, ctr);
            else if (bounded) k_bih_traverse<false, false, true><<<g_bih[0], GW_THREADS, 0, st>>>(s->d, W, segidx, sg, ctr);
            else k_bih_traverse<false, false, false><<<g_bih[0], GW_THREADS, 0, st>>>(s->d, W, segidx, sg, ctr);
            trav_mark(s, st, 0);
        } else if (sg.kind == SEG_MESH) {
            unsigned int* ctr = s->w_counters + (s->w_counter_next++ & 1023);
            trav_mark(s, st, 2);
            k_bvh_closest<<<g_bvh, 128, 0, st>>>(s->d, W, segidx, sg, ctr);
            trav_mark(s, st, 2);
        } else {
            k_prims_closest<<<sgrid, 128, 0, st>>>(s->d, W, segidx, sg);
        }
        s->launches++;
        CK(cudaGetLastError());
    }
    if (side_by_side) {
        CK(cudaEventRecord(s->ev_join, s->st2));
        CK(cudaStreamWaitEvent(st, s->ev_join, 0));
        k_merge_hits<<<sgrid, 128, 0, st>>>(W, s->w2_hit_t, s->w2_hit_seg, s->w2_hit_item, s->w2_hit_sub, s->w2_hit_flags, 1);
        s->launches++;
        CK(cudaGetLastError());
    }
    k_surface<<<sgrid, 128, 0, st>>>(s->d, W, s->segs_dev);
    s->launches++;
    CK(cudaGetLastError());
    if (s->n_scene_lights > 0) {
        for (size_t i = 0; i < s->segs.size(); i++) {
            const Seg& sg = s->segs[i];
            if (sg.kind == SEG_BIH) {
                bool linear = (s->segs_linear[i] != 0);
                unsigned int* ctr = s->w_counters + (s->w_counter_next++ & 1023);
                trav_mark(s, st, 1);
                const bool bounded = s->segs_reps[i] != 0;
                if (linear && bounded) k_bih_traverse<true, true, true><<<g_bih[3], GW_THREADS, 0, st>>>(s->d, W, (int)i, sg, ctr);
                else if (linear) k_bih_traverse<true, true, false><<<g_bih[3], GW_THREADS, 0, st>>>(s->d, W, (int)i, sg, ctr);
                else if (bounded) k_bih_traverse<true, false, true><<<g_bih[2], GW_THREADS, 0, st>>>(s->d, W, (int)i, sg, ctr);
                else k_bih_traverse<true, false, false><<<g_bih[2], GW_THREADS, 0, st>>>(s->d, W, (int)i, sg, ctr);
                trav_mark(s, st, 1);
            } else if (sg.kind == SEG_PRIMS) {
                k_prims_any<<<sgrid, 128, 0, st>>>(s->d, W, sg);
            } else continue;  // a Mesh casts no shadows (Mesh.hs:210)
            s->launches++;
            CK(cudaGetLastError());
        }
    }
    k_shade<<<sgrid, 128, 0, st>>>(s->d, W, s->segs_dev);
    s->launches++;
    CK(cudaGetLastError());
    return GLOME_OK;
}

extern "C" int GLOME_API(glome_render_dev)(GlomeScene* s, const GlomeCamera* cam, int width, int height, const GlomeRenderOpts* o,
                                double* tcolor_dev, uint32_t* rgb8_dev, GlomeRenderStats* stats, void* stream) {
    TWIN(glome_render_dev_f32(s, cam, width, height, o, tcolor_dev, rgb8_dev, stats, stream));
    if (!s || !cam || !o || !tcolor_dev || width <= 0 || height <= 0 || o->blocksize <= 0 || o->tile_stride <= 0 ||
        o->tile_first < 0 || o->tile_first >= o->tile_stride) { g_err = "bad argument"; return GLOME_EINVAL; }
    if (o->mode < GLOME_MODE_ONE_RAY || o->mode > GLOME_MODE_ADAPTIVE_AA_STRICT) { g_err = "bad render mode"; return GLOME_EINVAL; }
    if (o->debug_heatmap && (o->mode != GLOME_MODE_ONE_RAY || o->tint_depth)) {
        g_err = "debug_heatmap is get_color_debug per pixel (Glome.hs:57-60): GLOME_MODE_ONE_RAY without tint_depth only";
        return GLOME_EINVAL;
    }
    if (s->scene_class == GLOME_CLASS_GENERAL && o->recurs > GS_MAX_RECURS) {
        g_err = "recurs above the general tracer's limit of " + std::to_string(GS_MAX_RECURS) + " generations";
        return GLOME_ELIMIT;
    }
    CK(cudaSetDevice(s->device));
    cudaStream_t st = (cudaStream_t)stream;
    TileGeom g = make_geom(width, height, o->blocksize);
    int ntiles = g.ntx * g.nty;
    int n_sel = (ntiles - o->tile_first + o->tile_stride - 1) / o->tile_stride;
    if (n_sel < 0) n_sel = 0;
    int launches0 = s->launches;
    size_t npix = (size_t)width * height;
    s->time_traversal = (stats != nullptr);
    s->tev_used = 0;
    CK(cudaMemsetAsync(s->stats, 0, sizeof(DevStats), st));
    CK(cudaEventRecord(s->ev0, st));
    TraceParams P;
    memset(&P, 0, sizeof(P));
    P.g = g;
    P.cam.pos = vec(cam->pos[0], cam->pos[1], cam->pos[2]);
    P.cam.fwd = vec(cam->fwd[0], cam->fwd[1], cam->fwd[2]);
    P.cam.up = vec(cam->up[0], cam->up[1], cam->up[2]);
    P.cam.right = vec(cam->right[0], cam->right[1], cam->right[2]);
    P.recurs = o->recurs; P.tint = o->tint_depth;
    P.tile_first = o->tile_first; P.tile_stride = o->tile_stride; P.n_sel = n_sel;
    P.work_counter = s->work_counter; P.st = s->stats;
    int rc;
    gwave::WaveParams W;
    memset(&W, 0, sizeof(W));
    if (s->use_wave && n_sel > 0) {
        W.g.width = g.width; W.g.height = g.height; W.g.bs = g.bs; W.g.ntx = g.ntx; W.g.nty = g.nty; W.g.nbx = g.nbx;
        W.g.nby = g.nby; W.g.slots_per_tile = g.slots_per_tile;
        W.cam = P.cam; W.recurs = o->recurs; W.tint = o->tint_depth;
        W.tile_first = o->tile_first; W.tile_stride = o->tile_stride; W.n_sel = n_sel;
        size_t need = (size_t)n_sel * g.slots_per_tile;  // padded 8x4 micro-tiles of the selected tiles
        if (o->mode != GLOME_MODE_ONE_RAY && npix > need) need = npix;
        if ((rc = wave_reserve(s, need))) return rc;
        s->w_counter_next = 0;
        CK(cudaMemsetAsync(s->w_counters, 0, sizeof(unsigned int) * 1024, st));
    }
    if (n_sel > 0) {
        if (o->mode == GLOME_MODE_ONE_RAY) {
            P.out = tcolor_dev;
            if (s->use_wave) {
                W.mode = 0; W.out = tcolor_dev;
                if ((rc = launch_wave(s, W, (long long)n_sel * g.slots_per_tile, st))) return rc;
            } else if ((rc = launch_trace_c<0>(s, P, st))) return rc;
            if (o->debug_heatmap) {
                k_debug_tint<<<(unsigned)((npix + 63) / 64), 64, 0, st>>>(s->d, g, P.cam, o->tile_first, o->tile_stride, tcolor_dev);
                s->launches++;
                CK(cudaGetLastError());
            }
        } else {
            // workspace: v (pass 1-4 samples), ray queue
            if (s->v_pix < npix) {
                cudaFree(s->v); s->v = nullptr; s->v_pix = 0;
                CK(cudaMalloc((void**)&s->v, npix * 5 * sizeof(double)));
                s->v_pix = npix;
            }
            if (s->queue_cap < npix) {
                cudaFree(s->queue); s->queue = nullptr; s->queue_cap = 0;
                CK(cudaMalloc((void**)&s->queue, npix * sizeof(int)));
                s->queue_cap = npix;
            }
            DecideParams D;
            memset(&D, 0, sizeof(D));
            D.g = g; D.tile_first = o->tile_first; D.tile_stride = o->tile_stride;
            D.v = s->v; D.v2 = tcolor_dev; D.queue = s->queue; D.queue_count = s->queue_count;
            P.queue = s->queue; P.queue_count = s->queue_count; P.v = s->v;
            // General scenes: one wave costs several ms of latency however few rays it holds (the recursive
            // interpreter is latency-bound per warp), so passes 1-4 are speculated: every pixel centre is traced
            // once, up front, and the per-pass decisions copy from that buffer.  get_color is a pure function of the
            // sample position, so the frame is bit-identical to the adaptive schedule; only the ray count differs.
            bool speculate = !s->use_wave && !s->env_no_speculate;
            int tune = -1;  // flat scenes: the schedule this frame is a timing sample of (-1: none)
            if (o->mode == GLOME_MODE_ADAPTIVE_AA_STRICT) speculate = false;
            if (s->use_wave) {
                // Flat scenes: a wave is cheap per ray but five dependent waves are not (each sits on its latency
                // floor), so the same speculation pays when the adaptive schedule ends up tracing most pixel centres
                // anyway (a cloud of small spheres: 95 %) and costs rays when it does not (a smooth mesh: 30 %).
                // The schedule of this frame follows what the previous frame of the same geometry did; the frame
                // itself is bit-identical either way.
                // That ratio is only the first guess.  Whether four more dependent waves or 2-3x the rays cost more depends on
                // how many rays this device owns (1/N of the tiles when the frame is sharded: the waves of a small share
                // sit on their latency floors, and tracing every centre is then the cheaper schedule), so both schedules
                // are timed on the device and the faster one is kept; the other is re-probed now and then.
                const bool same_geom = s->aa_valid && s->aa_w == width && s->aa_h == height && s->aa_first == o->tile_first &&
                                       s->aa_stride == o->tile_stride && s->aa_bs == o->blocksize;
                if (!same_geom) { s->aa_ms_valid[0] = s->aa_ms_valid[1] = false; s->aa_pending[0] = s->aa_pending[1] = false; s->aa_frames = 0; s->aa_nsamp[0] = s->aa_nsamp[1] = 0; }
                if (s->env_aa_speculate >= 0) speculate = s->env_aa_speculate != 0;
                else if (o->mode == GLOME_MODE_ADAPTIVE_AA_STRICT) speculate = false;
                else {
                    for (int k = 0; k < 2; k++)
                        if (s->aa_pending[k] && event_done(s->aa_t1[k])) {
                            float ms = 0;
                            if (cudaEventElapsedTime(&ms, s->aa_t0[k], s->aa_t1[k]) == cudaSuccess) { s->aa_ms[k] = ms; s->aa_ms_valid[k] = ++s->aa_nsamp[k] >= 2; }
                            else cudaGetLastError();
                            s->aa_pending[k] = false;
                        }
                    if (s->aa_ms_valid[0] && s->aa_ms_valid[1]) {
                        const int best = s->aa_ms[1] < s->aa_ms[0] ? 1 : 0;
                        tune = (++s->aa_frames % 128 == 0) ? 1 - best : best;  // re-probe the other schedule now and then
                    } else if (s->aa_ms_valid[0] || s->aa_ms_valid[1]) {
                        const int known = s->aa_ms_valid[0] ? 0 : 1;
                        tune = s->aa_pending[1 - known] ? known : 1 - known;
                    } else if (same_geom && event_done(s->aa_ev)) {
                        long long centres = 0, traced = 0;
                        for (int k = 0; k < n_sel; k++) {
                            int ti = o->tile_first + k * o->tile_stride, tx = ti / g.nty, ty = ti % g.nty;
                            centres += (long long)std::min(g.bs, width - tx * g.bs) * std::min(g.bs, height - ty * g.bs);
                        }
                        for (int p = 1; p <= 4; p++) traced += s->aa_counts_host[p];
                        tune = (double)traced >= GLOME_AA_SPEC_RATIO * (double)centres ? 1 : 0;
                    } else tune = 0;
                    speculate = tune == 1;
                    if (s->aa_pending[tune] || !same_geom) tune = -1;  // its last measurement is still in flight; a geometry's first frame (allocations, cold caches) is no sample
                }
                if (!s->aa_counts_host) {
                    CK(cudaMallocHost((void**)&s->aa_counts_host, 8 * sizeof(int)));
                    CK(cudaEventCreateWithFlags(&s->aa_ev, cudaEventDisableTiming));
                    for (int k = 0; k < 2; k++) { CK(cudaEventCreate(&s->aa_t0[k])); CK(cudaEventCreate(&s->aa_t1[k])); }
                }
                if (tune >= 0) CK(cudaEventRecord(s->aa_t0[tune], st));
            }
            D.spec = nullptr;
            if (speculate) {
                if (s->spec_pix < npix) {
                    cudaFree(s->spec); s->spec = nullptr; s->spec_pix = 0;
                    CK(cudaMalloc((void**)&s->spec, npix * 5 * sizeof(double)));
                    s->spec_pix = npix;
                }
                if (s->use_wave) {
                    W.mode = 0; W.out = s->spec; W.tint = 0;
                    if ((rc = launch_wave(s, W, (long long)n_sel * g.slots_per_tile, st))) return rc;
                    W.tint = o->tint_depth;
                } else {
                    TraceParams P0 = P;
                    P0.out = s->spec; P0.tint = 0;
                    if ((rc = launch_trace_c<0>(s, P0, st))) return rc;
                }
                D.spec = s->spec;
            }
            for (int pass = 1; pass <= 5; pass++) {
                D.pass = pass;
                D.threshold = pass >= 2 ? o->thresholds[pass - 2] : 0;
                CK(cudaMemsetAsync(s->queue_count, 0, sizeof(int), st));
                k_aa_decide<<<n_sel, 256, 0, st>>>(D);
                s->launches++;
                CK(cudaGetLastError());
                if (s->use_wave && pass < 5) CK(cudaMemcpyAsync(s->aa_counts_host + pass, s->queue_count, sizeof(int), cudaMemcpyDeviceToHost, st));
                if (speculate && pass < 5) continue;
                if (s->use_wave) {
                    W.mode = pass < 5 ? 1 : 5; W.queue = s->queue; W.queue_count = s->queue_count; W.v = s->v;
                    W.out = pass < 5 ? s->v : tcolor_dev;
                    if ((rc = launch_wave(s, W, (long long)npix, st))) return rc;
                } else if (pass < 5) { P.out = s->v; if ((rc = launch_trace_c<1>(s, P, st))) return rc; }
                else { P.out = tcolor_dev; if ((rc = launch_trace_c<5>(s, P, st))) return rc; }
            }
            if (s->use_wave) {
                if (tune >= 0) { CK(cudaEventRecord(s->aa_t1[tune], st)); s->aa_pending[tune] = true; }
                CK(cudaEventRecord(s->aa_ev, st));
                s->aa_valid = true;
                s->aa_w = width; s->aa_h = height; s->aa_first = o->tile_first; s->aa_stride = o->tile_stride; s->aa_bs = o->blocksize;
            }
        }
        if (rgb8_dev) {
            k_pack_rgb8<<<n_sel, 256, 0, st>>>(g, o->tile_first, o->tile_stride, tcolor_dev, rgb8_dev);
            s->launches++;
            CK(cudaGetLastError());
        }
    }
    CK(cudaEventRecord(s->ev1, st));
    if (stats) {
        CK(cudaEventSynchronize(s->ev1));
        float ms = 0;
        CK(cudaEventElapsedTime(&ms, s->ev0, s->ev1));
        read_stats(s, stats, ms, s->launches - launches0);
        double tms = 0;
        for (int k = 0; k + 1 < s->tev_used; k += 2) {
            float t = 0;
            if (cudaEventElapsedTime(&t, s->tev[k], s->tev[k + 1]) == cudaSuccess) {
                tms += t;
                stats->family_ms[s->tev_family[k] & 3] += t;
                stats->family_launches[s->tev_family[k] & 3] += 1;
            }
        }
        stats->traverse_ms = tms;
        stats->traverse_launches = s->tev_used / 2;
    }
    return GLOME_OK;
}

extern "C" int GLOME_API(glome_render)(GlomeScene* s, const GlomeCamera* cam, int width, int height, const GlomeRenderOpts* o,
                            double* tcolor, uint32_t* rgb8, GlomeRenderStats* stats) {
    TWIN(glome_render_f32(s, cam, width, height, o, tcolor, rgb8, stats));
    if (!s || (!tcolor && !rgb8) || width <= 0 || height <= 0) { g_err = "bad argument"; return GLOME_EINVAL; }
    CK(cudaSetDevice(s->device));
    size_t npix = (size_t)width * height;
    if (s->v2_pix < npix) {
        cudaFree(s->v2); s->v2 = nullptr; s->v2_pix = 0;
        CK(cudaMalloc((void**)&s->v2, npix * 5 * sizeof(double)));
        s->v2_pix = npix;
    }
    if (rgb8 && s->rgb8_pix < npix) {
        cudaFree(s->rgb8); s->rgb8 = nullptr; s->rgb8_pix = 0;
        CK(cudaMalloc((void**)&s->rgb8, npix * sizeof(uint32_t)));
        s->rgb8_pix = npix;
    }
    // pixels of unselected tiles must be left untouched: start from the caller's buffer
    if (o && o->tile_stride > 1) {
        if (tcolor) CK(cudaMemcpy(s->v2, tcolor, npix * 5 * sizeof(double), cudaMemcpyHostToDevice));
        if (rgb8) CK(cudaMemcpy(s->rgb8, rgb8, npix * sizeof(uint32_t), cudaMemcpyHostToDevice));
    }
    GlomeRenderStats local;
    int rc = GLOME_API(glome_render_dev)(s, cam, width, height, o, s->v2, rgb8 ? s->rgb8 : nullptr, &local, nullptr);
    if (rc) return rc;
    if (tcolor) CK(cudaMemcpy(tcolor, s->v2, npix * 5 * sizeof(double), cudaMemcpyDeviceToHost));
    if (rgb8) CK(cudaMemcpy(rgb8, s->rgb8, npix * sizeof(uint32_t), cudaMemcpyDeviceToHost));
    if (stats) *stats = local;
    return GLOME_OK;
}

// Run-time switches of one scene (measurement and A/B runs; the frame never depends on them).
extern "C" int GLOME_API(glome_scene_set_option)(GlomeScene* s, int option, int value) {
    TWIN(glome_scene_set_option_f32(s, option, value));
    if (!s) { g_err = "bad argument"; return GLOME_EINVAL; }
    switch (option) {
        case GLOME_OPT_SEG_CONCURRENT: s->env_seg_concurrent = value != 0; return GLOME_OK;
        case GLOME_OPT_AA_SPECULATE: s->env_aa_speculate = value < 0 ? -1 : (value != 0); return GLOME_OK;
    }
    g_err = "unknown scene option";
    return GLOME_EINVAL;
}

extern "C" int64_t GLOME_API(glome_scene_launches)(GlomeScene* s) {
    TWIN(glome_scene_launches_f32(s));
    return s ? (int64_t)s->launches : 0;
}

#ifndef GLOME_F32  // device memory helpers, tile plumbing and the several-GPUs-in-one-process driver exist once (they hold FP64 scenes)
extern "C" int glome_dev_alloc(int device, int64_t bytes, void** out) {
    if (!out || bytes < 0) { g_err = "bad argument"; return GLOME_EINVAL; }
    CK(cudaSetDevice(device));
    CK(cudaMalloc(out, (size_t)(bytes ? bytes : 1)));
    return GLOME_OK;
}
extern "C" int glome_dev_free(int device, void* p) {
    CK(cudaSetDevice(device));
    CK(cudaFree(p));
    return GLOME_OK;
}

// ---------------------------------------------------------------------------------------------
// multi-GPU plumbing: a rank's tiles <-> a contiguous slot buffer (equal-sized all-gather payload)
// ---------------------------------------------------------------------------------------------
// blockIdx.x = slot, blockIdx.y = rank offset (unpack of a whole gathered buffer: rank r's block holds the tiles
// r, r+stride, ...; skip_rank's own tiles are already in place)
template <bool PACK>
__global__ void __launch_bounds__(256) k_tiles_copy(TileGeom g, int tile_first, int tile_stride, int words,
                                                    uint32_t* __restrict__ frame, uint32_t* __restrict__ packed,
                                                    int slots, int skip_rank) {
    int rank = tile_first + blockIdx.y;
    if (rank == skip_rank) return;
    int ti = rank + blockIdx.x * tile_stride;
    if (ti >= g.ntx * g.nty) return;
    int xt, yt, tw, th;
    tile_rect(g, ti, xt, yt, tw, th);
    size_t slot = ((size_t)blockIdx.y * slots + blockIdx.x) * g.bs * g.bs * words;
    int n = tw * th * words;
    for (int i = threadIdx.x; i < n; i += blockDim.x) {
        int p = i / words, wd = i % words;
        size_t fi = ((size_t)(yt + p / tw) * g.width + (xt + p % tw)) * words + wd;
        size_t pi = slot + (size_t)p * words + wd;
        if (PACK) packed[pi] = frame[fi];
        else frame[fi] = packed[pi];
    }
}
static int tiles_copy(bool pack, int width, int height, int blocksize, int tile_first, int tile_stride, int elem_bytes,
                      void* frame, void* packed, void* stream) {
    if (width <= 0 || height <= 0 || blocksize <= 0 || tile_stride <= 0 || tile_first < 0 || elem_bytes <= 0 ||
        (elem_bytes & 3) || !frame || !packed) { g_err = "bad argument"; return GLOME_EINVAL; }
    TileGeom g = make_geom(width, height, blocksize);
    int slots = glome_tile_slots(width, height, blocksize, tile_stride);
    if (slots == 0) return GLOME_OK;
    cudaStream_t st = (cudaStream_t)stream;
    if (pack) k_tiles_copy<true><<<slots, 256, 0, st>>>(g, tile_first, tile_stride, elem_bytes / 4, (uint32_t*)frame, (uint32_t*)packed, slots, -1);
    else k_tiles_copy<false><<<slots, 256, 0, st>>>(g, tile_first, tile_stride, elem_bytes / 4, (uint32_t*)frame, (uint32_t*)packed, slots, -1);
    CK(cudaGetLastError());
    return GLOME_OK;
}
extern "C" int glome_tiles_pack_dev(int width, int height, int blocksize, int tile_first, int tile_stride, int elem_bytes,
                                    const void* frame_dev, void* packed_dev, void* stream) {
    return tiles_copy(true, width, height, blocksize, tile_first, tile_stride, elem_bytes, (void*)frame_dev, packed_dev, stream);
}
extern "C" int glome_tiles_unpack_dev(int width, int height, int blocksize, int tile_first, int tile_stride, int elem_bytes,
                                      const void* packed_dev, void* frame_dev, void* stream) {
    return tiles_copy(false, width, height, blocksize, tile_first, tile_stride, elem_bytes, frame_dev, (void*)packed_dev, stream);
}

// Scatter a whole all-gathered buffer (tile_stride rank blocks of `slots` slots each) into the frame in one launch.
extern "C" int glome_tiles_unpack_all_dev(int width, int height, int blocksize, int tile_stride, int skip_rank, int elem_bytes,
                                          const void* gathered_dev, void* frame_dev, void* stream) {
    if (width <= 0 || height <= 0 || blocksize <= 0 || tile_stride <= 0 || elem_bytes <= 0 || (elem_bytes & 3) || !gathered_dev ||
        !frame_dev) { g_err = "bad argument"; return GLOME_EINVAL; }
    TileGeom g = make_geom(width, height, blocksize);
    int slots = glome_tile_slots(width, height, blocksize, tile_stride);
    if (slots == 0) return GLOME_OK;
    dim3 grid(slots, tile_stride);
    k_tiles_copy<false><<<grid, 256, 0, (cudaStream_t)stream>>>(g, 0, tile_stride, elem_bytes / 4, (uint32_t*)frame_dev,
                                                               (uint32_t*)gathered_dev, slots, skip_rank);
    CK(cudaGetLastError());
    return GLOME_OK;
}

// ---------------------------------------------------------------------------------------------
// One process driving several GPUs: what a Haskell caller links, with no torch and no
// NCCL in the picture.  The scene is replicated; device k renders tiles i with i % N == k on its own stream; each
// device packs its tiles and copies them to the first device with cudaMemcpyPeerAsync (NVLink P2P); the first device
// unpacks all of them into the frame.  Same kernels, same tile order, so the frame equals the 1-GPU frame bit for bit.
// ---------------------------------------------------------------------------------------------
struct GlomeMulti {
    int n;
    std::vector<int> dev;
    std::vector<GlomeScene*> scene;
    std::vector<cudaStream_t> stream;
    std::vector<cudaEvent_t> done;
    std::vector<double*> tc;         // per device: full TColor frame (its own tiles are valid)
    std::vector<uint32_t*> rgb;      // per device: packed image
    std::vector<void*> pk;           // per device: its packed slots (largest element size)
    void* gather;                    // device 0: N blocks of slots
    size_t pix, pk_bytes, gather_bytes;
    cudaEvent_t e0, e1;
};

static void multi_free_frames(GlomeMulti* m) {
    for (int k = 0; k < m->n; k++) {
        cudaSetDevice(m->dev[k]);
        cudaFree(m->tc[k]); cudaFree(m->rgb[k]); cudaFree(m->pk[k]);
        m->tc[k] = nullptr; m->rgb[k] = nullptr; m->pk[k] = nullptr;
    }
    cudaSetDevice(m->dev[0]);
    cudaFree(m->gather);
    m->gather = nullptr;
    m->pix = 0; m->pk_bytes = 0; m->gather_bytes = 0;
}

extern "C" int glome_multi_destroy(GlomeMulti* m) {
    if (!m) return GLOME_OK;
    multi_free_frames(m);
    for (int k = 0; k < m->n; k++) {
        cudaSetDevice(m->dev[k]);
        if (m->stream[k]) cudaStreamDestroy(m->stream[k]);
        if (m->done[k]) cudaEventDestroy(m->done[k]);
        if (m->scene[k]) glome_scene_destroy(m->scene[k]);
    }
    cudaSetDevice(m->dev[0]);
    if (m->e0) cudaEventDestroy(m->e0);
    if (m->e1) cudaEventDestroy(m->e1);
    delete m;
    return GLOME_OK;
}

extern "C" int glome_multi_create(const GlomeFlatScene* desc, int ndev, const int* devices, GlomeMulti** out) {
    if (!desc || !out || ndev < 1 || ndev > 64 || !devices) { g_err = "bad argument"; return GLOME_EINVAL; }
    GlomeMulti* m = new GlomeMulti();
    m->n = ndev;
    m->dev.assign(devices, devices + ndev);
    m->scene.assign(ndev, nullptr); m->stream.assign(ndev, nullptr); m->done.assign(ndev, nullptr);
    m->tc.assign(ndev, nullptr); m->rgb.assign(ndev, nullptr); m->pk.assign(ndev, nullptr);
    m->gather = nullptr; m->pix = 0; m->pk_bytes = 0; m->gather_bytes = 0; m->e0 = nullptr; m->e1 = nullptr;
    for (int k = 0; k < ndev; k++) {
        int rc = glome_scene_create(desc, devices[k], &m->scene[k]);
        if (rc) { glome_multi_destroy(m); return rc; }
        if (cudaSetDevice(devices[k]) != cudaSuccess || cudaStreamCreateWithFlags(&m->stream[k], cudaStreamNonBlocking) != cudaSuccess ||
            cudaEventCreateWithFlags(&m->done[k], cudaEventDisableTiming) != cudaSuccess) {
            g_err = "glome_multi_create: stream / event creation failed"; glome_multi_destroy(m); return GLOME_ECUDA;
        }
        if (k > 0) {  // peer access both ways with the gathering device (already-enabled is fine)
            int can = 0;
            cudaDeviceCanAccessPeer(&can, devices[k], devices[0]);
            if (can) { cudaDeviceEnablePeerAccess(devices[0], 0); cudaGetLastError(); }
            cudaSetDevice(devices[0]);
            cudaDeviceCanAccessPeer(&can, devices[0], devices[k]);
            if (can) { cudaDeviceEnablePeerAccess(devices[k], 0); cudaGetLastError(); }
        }
    }
    cudaSetDevice(devices[0]);
    if (cudaEventCreate(&m->e0) != cudaSuccess || cudaEventCreate(&m->e1) != cudaSuccess) {
        g_err = "glome_multi_create: event creation failed"; glome_multi_destroy(m); return GLOME_ECUDA;
    }
    *out = m;
    return GLOME_OK;
}

// renderTiles over the GPUs of this process.  tcolor (w*h*5 doubles) and / or rgb8 (w*h uint32) are host buffers.
// stats: ray counts summed over the devices, kernel_ms = device time of the whole frame on the gathering device's
// clock (render on all devices + peer copies + unpack).
extern "C" int glome_multi_render(GlomeMulti* m, const GlomeCamera* cam, int width, int height, const GlomeRenderOpts* o,
                                  double* tcolor, uint32_t* rgb8, GlomeRenderStats* stats) {
    if (!m || !cam || !o || (!tcolor && !rgb8) || width <= 0 || height <= 0 || o->blocksize <= 0) { g_err = "bad argument"; return GLOME_EINVAL; }
    if (o->tile_stride != 1 || o->tile_first != 0) { g_err = "glome_multi_render shards the tiles itself: tile_first = 0, tile_stride = 1"; return GLOME_EINVAL; }
    const int N = m->n;
    const size_t npix = (size_t)width * height;
    const int slots = glome_tile_slots(width, height, o->blocksize, N);
    const size_t slot_elems = (size_t)slots * o->blocksize * o->blocksize;
    const size_t pk_bytes = slot_elems * 40;
    if (m->pix < npix || m->pk_bytes < pk_bytes) {
        multi_free_frames(m);
        for (int k = 0; k < N; k++) {
            CK(cudaSetDevice(m->dev[k]));
            CK(cudaMalloc((void**)&m->tc[k], npix * 5 * sizeof(double)));
            CK(cudaMalloc((void**)&m->rgb[k], npix * sizeof(uint32_t)));
            CK(cudaMalloc(&m->pk[k], pk_bytes));
        }
        CK(cudaSetDevice(m->dev[0]));
        CK(cudaMalloc(&m->gather, pk_bytes * N));
        m->pix = npix; m->pk_bytes = pk_bytes; m->gather_bytes = pk_bytes * N;
    }
    CK(cudaSetDevice(m->dev[0]));
    CK(cudaEventRecord(m->e0, m->stream[0]));
    // the other devices start after e0 so that kernel_ms covers them
    int rc;
    {   // issue every device's frame from its own host thread: a frame is ~10 launches, and issuing eight of them one
        // after the other from one thread would delay the last device by as much as it then computes
        std::vector<int> rcs(N, 0);
        std::vector<std::string> errs(N);
        auto issue = [&](int k) {
            auto body = [&]() -> int {
                CK(cudaSetDevice(m->dev[k]));
                if (k > 0) CK(cudaStreamWaitEvent(m->stream[k], m->e0, 0));
                GlomeRenderOpts ok = *o;
                ok.tile_first = k; ok.tile_stride = N;
                return glome_render_dev(m->scene[k], cam, width, height, &ok, m->tc[k], rgb8 ? m->rgb[k] : nullptr, nullptr, m->stream[k]);
            };
            rcs[k] = body();
            if (rcs[k]) errs[k] = g_err;  // g_err is thread-local: hand the message to the calling thread
        };
        std::vector<std::thread> th;
        for (int k = 1; k < N; k++) th.emplace_back(issue, k);
        issue(0);
        for (auto& t : th) t.join();
        for (int k = 0; k < N; k++)
            if (rcs[k]) { g_err = errs[k]; return rcs[k]; }
        CK(cudaSetDevice(m->dev[0]));
    }
    // gather: TColor first (if wanted), then the packed image, through the same buffers
    for (int what = 0; what < 2; what++) {
        const bool want = what == 0 ? (tcolor != nullptr) : (rgb8 != nullptr);
        if (!want) continue;
        const int eb = what == 0 ? 40 : 4;
        const size_t bytes = slot_elems * eb;
        for (int k = 1; k < N; k++) {
            CK(cudaSetDevice(m->dev[k]));
            const void* frame = what == 0 ? (const void*)m->tc[k] : (const void*)m->rgb[k];
            if ((rc = glome_tiles_pack_dev(width, height, o->blocksize, k, N, eb, frame, m->pk[k], m->stream[k]))) return rc;
            CK(cudaMemcpyPeerAsync((char*)m->gather + (size_t)k * bytes, m->dev[0], m->pk[k], m->dev[k], bytes, m->stream[k]));
            CK(cudaEventRecord(m->done[k], m->stream[k]));
        }
        CK(cudaSetDevice(m->dev[0]));
        for (int k = 1; k < N; k++) CK(cudaStreamWaitEvent(m->stream[0], m->done[k], 0));
        void* frame0 = what == 0 ? (void*)m->tc[0] : (void*)m->rgb[0];
        if (N > 1 && (rc = glome_tiles_unpack_all_dev(width, height, o->blocksize, N, 0, eb, m->gather, frame0, m->stream[0]))) return rc;
        if (what == 0 && rgb8 && N > 1) {
            // the gather buffer is reused for the packed image: the other devices may overwrite it only after this unpack
            CK(cudaEventRecord(m->done[0], m->stream[0]));
            for (int k = 1; k < N; k++) { CK(cudaSetDevice(m->dev[k])); CK(cudaStreamWaitEvent(m->stream[k], m->done[0], 0)); }
            CK(cudaSetDevice(m->dev[0]));
        }
    }
    CK(cudaEventRecord(m->e1, m->stream[0]));
    if (tcolor) CK(cudaMemcpyAsync(tcolor, m->tc[0], npix * 5 * sizeof(double), cudaMemcpyDeviceToHost, m->stream[0]));
    if (rgb8) CK(cudaMemcpyAsync(rgb8, m->rgb[0], npix * sizeof(uint32_t), cudaMemcpyDeviceToHost, m->stream[0]));
    CK(cudaStreamSynchronize(m->stream[0]));
    if (stats) {
        float ms = 0;
        cudaEventElapsedTime(&ms, m->e0, m->e1);
        GlomeRenderStats tot;
        memset(&tot, 0, sizeof(tot));
        for (int k = 0; k < N; k++) {
            CK(cudaSetDevice(m->dev[k]));
            CK(cudaStreamSynchronize(m->stream[k]));
            GlomeRenderStats sk;
            memset(&sk, 0, sizeof(sk));
            read_stats(m->scene[k], &sk, 0, 0);
            tot.rays_primary += sk.rays_primary; tot.rays_shadow += sk.rays_shadow; tot.rays_secondary += sk.rays_secondary;
            tot.overflow_rays += sk.overflow_rays; tot.perlin_range += sk.perlin_range;
            tot.visits_bih += sk.visits_bih; tot.tests_prim += sk.tests_prim; tot.visits_bvh += sk.visits_bvh; tot.tests_tri += sk.tests_tri;
        }
        tot.kernel_ms = ms;
        *stats = tot;
    }
    return GLOME_OK;
}
#endif  // !GLOME_F32

}  // namespace GLOME_PREC_NS
