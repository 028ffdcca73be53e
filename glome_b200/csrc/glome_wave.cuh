// glome_wave.cuh -- wavefront pipeline for GLOME_CLASS_FLAT scenes.
//
// ncu on the first (megakernel) version showed the limiter of this path is SIMT divergence, not
// bandwidth: 5.6 of 32 lanes active on average, because a thread walked primary ray -> shadow ray
// 1 -> shadow ray 2 -> shade in sequence and the lanes of a warp finish their walks at very
// different times.  Here the frame is a sequence of waves over a sample list:
//
//   K1  k_bih_traverse<closest> / k_bvh_closest / k_prims_closest   one launch per top-level
//       segment of the scene, in the reference's fold order; persistent; A LANE WHOSE RAY HAS
//       FINISHED IMMEDIATELY PULLS THE NEXT SAMPLE, so warps stay full; while-while loop shape
//       (all lanes walk branches until every lane sits on a leaf, then all do leaves)
//   K2a k_surface   position/normal of the winning primitive, light facing tests, shadow-ray
//       queue built with ballot + popc prefix
//   K1' k_bih_traverse<any> / k_prims_any over the shadow queue (any-hit, early exit)
//   K2b k_shade     materialShader on the unoccluded lights, TColor out (+ pass-5 combine)
//
// Arithmetic per ray is exactly the megakernel's (same device functions), so results are
// bit-identical; only the schedule differs.
#pragma once
#include "glome_device.cuh"

namespace gwave {

using namespace gdev;

#define GW_MAX_SEGS 32
enum { SEG_BIH = 0, SEG_MESH = 1, SEG_PRIMS = 2 };

// A top-level segment of a flat scene: a Bih, a Mesh, or a run of loose (wrapped) primitives.
struct Seg {
    int kind;
    int node;        // BIH / MESH node index (after unwrapping); PRIMS: first group child node
    int count;       // PRIMS: number of consecutive group children
    int pad;
    Stk tex, tag;    // stacks at the segment (root + child wrappers), head first
};

struct TileGeomW {
    int width, height, bs, ntx, nty, nbx, nby, slots_per_tile;
};

struct WaveParams {
    TileGeomW g;
    DCamera cam;
    int mode;                // 0: implicit one-ray-per-pixel samples; 1: queue at pixel centre grid (getCoords);
                             // 5: queue at (x+0.5, y+0.5)
    int recurs, tint;
    int tile_first, tile_stride, n_sel;
    const int* queue;
    const int* queue_count;
    // per-sample state
    Flt* hit_t;
    int* hit_seg;
    int* hit_item;
    int* hit_sub;
    int* hit_flags;
    Flt* surf;               // pos(3), norm(3) per sample
    unsigned int* occl;      // bit li = light li is occluded
    int2* squeue;            // shadow queue {sample, light}
    int* squeue_count;
    const double* v;         // AA pass 5 reads
    double* out;             // TColor frame written by k_shade
    unsigned long long* stats;  // DevStats as 9 counters
};

__device__ __forceinline__ long long wave_total(const WaveParams& P) {
    return (P.mode == 0) ? (long long)P.n_sel * P.g.slots_per_tile : (long long)(*P.queue_count);
}
// sample index -> pixel; false for the padding slots of partial tiles
__device__ __forceinline__ bool sample_pixel(const WaveParams& P, long long s, int& x, int& y) {
    if (P.mode == 0) {
        int k = (int)(s / P.g.slots_per_tile), local = (int)(s % P.g.slots_per_tile);
        int ti = P.tile_first + k * P.tile_stride;
        int tx = ti / P.g.nty, ty = ti % P.g.nty;
        int xt = tx * P.g.bs, yt = ty * P.g.bs;
        int tw = min(P.g.bs, P.g.width - xt), th = min(P.g.bs, P.g.height - yt);
        int blk = local >> 5, l = local & 31;
        int px = (blk % P.g.nbx) * 8 + (l & 7), py = (blk / P.g.nbx) * 4 + (l >> 3);
        x = xt + px; y = yt + py;
        return px < tw && py < th;
    }
    int pix = P.queue[s];
    x = pix % P.g.width; y = pix / P.g.width;
    return true;
}
__device__ __forceinline__ Ray sample_ray(const WaveParams& P, int x, int y) {
    Flt xc, yc;
    if (P.mode == 5) getCoordsf(P.g.width, P.g.height, (Flt)x + FL(0.5), (Flt)y + FL(0.5), xc, yc);
    else getCoordsf(P.g.width, P.g.height, (Flt)x, (Flt)y, xc, yc);
    return camera_ray(P.cam, xc, yc);
}

__device__ __forceinline__ void wave_flush(unsigned long long* st, const unsigned int (&vals_in)[9]) {
    unsigned int vals[9];
#pragma unroll
    for (int k = 0; k < 9; k++) {
        unsigned int v = vals_in[k];
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
        vals[k] = v;
    }
    if ((threadIdx.x & 31) == 0) {
#pragma unroll
        for (int k = 0; k < 9; k++)
            if (vals[k]) atomicAdd(st + k, (unsigned long long)vals[k]);
    }
}
// stats slots: 0 primary 1 shadow 2 secondary 3 overflow 4 perlin_range 5 bih 6 prim 7 bvh 8 tri

// shadow-ray of light li for the surface point of sample s (Shader.hs:70-78); same arithmetic as mpreshade
__device__ __forceinline__ bool light_ray(const DScene& S, const Flt* surf, int li, Ray& r, Flt& d, Vec& ldir, Flt& llen,
                                          bool& need_shadow) {
    const GlomeLight* Lp = S.lights + li;
    Vec pos = vec(surf[0], surf[1], surf[2]), norm = vec(surf[3], surf[4], surf[5]);
    Vec lvec = vsub(vec(Lp->pos[0], Lp->pos[1], Lp->pos[2]), pos);
    if (vdot(lvec, norm) < 0) return false;
    llen = vlen(lvec);
    ldir = vscale(lvec, 1 / llen);
    if (llen > Lp->rad) return false;
    need_shadow = Lp->do_shadow != 0;
    r = mkray(vscaleadd(pos, norm, GLM_DELTA), ldir);
    d = llen - (2 * GLM_DELTA);
    return true;
}

// unwrap {Tex,Tag,(NoShadow|OnlyShadow)}* down to the primitive; false if the chain blocks this query
template <bool ANY>
__device__ __forceinline__ bool unwrap_prim(const DScene& S, int ni, GlomeNode& nd, int& prim) {
    nd = S.nodes[ni];
    for (;;) {
        if (nd.type == GLOME_TEX || nd.type == GLOME_TAG || nd.type == (ANY ? GLOME_ONLYSHADOW : GLOME_NOSHADOW)) {
            ni = nd.a; nd = S.nodes[ni];
            continue;
        }
        break;
    }
    prim = ni;
    return is_prim(nd.type);
}

// ---------------------------------------------------------------------------------------------
// K1 / K1': persistent BIH traversal with per-lane refill.
//   ANY = false: closest hit over the sample list (rayint_bih, Bih.hs:332-368)
//   ANY = true : any hit over the shadow queue   (shadow_bih, Bih.hs:510-544)
//   LINEAR: every leaf item is a bare sphere with linear payload addressing
// ---------------------------------------------------------------------------------------------
#define GW_STACK 48
#ifndef GW_POLICY
#define GW_POLICY 0
#endif
#ifndef GW_BRANCH_BOUND
#define GW_BRANCH_BOUND 6
#endif
#ifndef GW_BRANCH_REPS
#define GW_BRANCH_REPS 2
#endif
#ifndef GW_THREADS
#define GW_THREADS 128
#endif
#ifndef GW_MINBLOCKS
#define GW_MINBLOCKS 8
#endif
#ifndef GW_REFILL_MIN
#define GW_REFILL_MIN 16  /* 8 -> 16: +2 % on small waves and on the mesh AA frame, same at full load (r1g A/B) */
#endif
#ifndef GW_REFILL_MIN_BOUNDED
#define GW_REFILL_MIN_BOUNDED 12  /* the bounded-branch-phase variant (deep trees): bound 6 + refill at 12 idle lanes, configs[1] 3.72 -> 3.58 ms */
#endif
#ifndef GW_BVH_MINBLOCKS
#define GW_BVH_MINBLOCKS 6  /* 80 regs, 24 warps/SM: 3 % over 3 blocks (106 regs); the kernel is bound by its ~200 instructions per two-box node, not by occupancy */
#endif
#ifndef GW_BVH_REFILL_MIN
#define GW_BVH_REFILL_MIN 16
#endif
#ifndef GW_GUIDED
#define GW_GUIDED 1  /* guided self-scheduling of the sample list */
#endif
#ifndef GW_GUIDED_MIN
#define GW_GUIDED_MIN 4
#endif
#ifndef GW_MAXBATCH
#define GW_MAXBATCH 32
#endif
#ifndef GW_STEAL_AFTER
#define GW_STEAL_AFTER 0
#endif
#ifndef GW_BVH_BOUND
#define GW_BVH_BOUND 12
#endif
#ifndef GW_BVH_GUIDED_DIV
#define GW_BVH_GUIDED_DIV 2
#endif
#ifndef GW_BVH_GUIDED_MIN
#define GW_BVH_GUIDED_MIN GW_GUIDED_MIN
#endif
#ifndef GW_BVH_PREFETCH
#define GW_BVH_PREFETCH 0
#endif
#ifndef GW_BVH_STEAL
#define GW_BVH_STEAL 1  /* the same for k_bvh_closest */
#endif
#ifndef GW_STEAL
#define GW_STEAL 1   /* drain-phase subtree donation between the lanes of a warp */
#endif

template <bool ANY, bool LINEAR, bool BOUNDED>
__global__ void __launch_bounds__(GW_THREADS, GW_MINBLOCKS) k_bih_traverse(DScene S, WaveParams P, int segidx, Seg seg, unsigned int* counter) {
    const unsigned int FULL = 0xffffffffu;
    const int lane = threadIdx.x & 31;
    const GlomeNode bn = S.nodes[seg.node];
    const Bbox bb = ldbb(S.dpool + bn.b);
    const int j0 = bn.c >> 4;
    const int a0 = LINEAR ? S.nodes[j0].a : 0;
    const long long total = ANY ? (long long)(*P.squeue_count) : wave_total(P);

    TravEnt stack[GW_STACK];
    int sp = 0, sb = 0, ref = 0;     // live stack entries are [sb, sp): the drain phase donates from the bottom
    bool active = false, nomore = false;
    long long s = 0;
    int light = 0;                   // ANY: light index.  closest: drain-phase group id (root lane), -1 = none
    Ray r = mkray(vec(0, 0, 0), vec(0, 0, 1));
    Flt drx = 0, dry = 0, drz = 0, near_ = 0, far_ = 0, dmax = 0;
    Flt best_t = GLM_INFINITY;
    int best_item = -1, best_seg = -1;
    bool has = false;
    unsigned int n_bih = 0, n_prim = 0, n_ovf = 0;
    unsigned int n_start = 0;        // n_bih when the lane's current ray started
#if GW_GUIDED
    unsigned int next_base = 0;
    const long long nwarps = (long long)gridDim.x * (GW_THREADS / 32);
#endif
#if GW_STEAL
    // drain-phase groups (closest hit only): the lanes of one warp that work on the same ray fold
    // their partial results here; the lane that brings g_pend to 0 writes the sample's result
    __shared__ Flt g_t[ANY ? 1 : GW_THREADS];
    __shared__ int g_item[ANY ? 1 : GW_THREADS], g_pend[ANY ? 1 : GW_THREADS], g_ovf[ANY ? 1 : GW_THREADS];
    __shared__ long long g_s[ANY ? 1 : GW_THREADS];
    __shared__ unsigned long long g_cull[ANY ? 1 : GW_THREADS];  // bits of the group's nearest depth so far (culling only)
    __shared__ int g_tie[ANY ? 1 : GW_THREADS];  // two members brought different items at exactly the same depth
    const int gb = threadIdx.x & ~31;
    if (!ANY) light = -1;
    bool solo = false;               // closest: this lane re-walks a ray alone (after a tie) and must not donate
#endif

    for (;;) {
        // ---- refill idle lanes ----
        unsigned int idle = __ballot_sync(FULL, !active);
        // refill in batches: ray set-up (camera ray, root clip, three divisions) is ~200 instructions, so it
        // should run with many lanes at once rather than one lane at a time
        if (__popc(idle) >= (BOUNDED ? GW_REFILL_MIN_BOUNDED : GW_REFILL_MIN) && !nomore) {
            unsigned int base = 0;
            int cnt = __popc(idle);
#if GW_GUIDED
            // guided self-scheduling: batches shrink as the list drains (and are small from the start when the
            // wave has fewer samples than lanes), so the expensive pixels of one screen tile end up on many
            // warps, each with idle lanes to share them, instead of 32 on one
            cnt = min(cnt, (int)min((long long)GW_MAXBATCH, max((long long)GW_GUIDED_MIN, (total - (long long)next_base) / (2LL * nwarps) + 1)));
#endif
            int leader = __ffs(idle) - 1;
            if (lane == leader) base = atomicAdd(counter, (unsigned int)cnt);
            base = __shfl_sync(FULL, base, leader);
            if ((long long)base + cnt >= total) nomore = true;
#if GW_GUIDED
            next_base = base + (unsigned int)cnt;
#endif
            const int rank = __popc(idle & ((1u << lane) - 1));
            if (!active && rank < cnt) {
                long long w = (long long)base + rank;
                if (w < total) {
                    bool ok;
                    if (ANY) {
                        int2 e = P.squeue[w];
                        s = e.x; light = e.y;
                        ok = ((P.occl[s] >> light) & 1u) == 0;  // an earlier segment already occludes it
                        if (ok) {
                            Vec ldir; Flt llen; bool ns;
                            ok = light_ray(S, P.surf + 6 * s, light, r, dmax, ldir, llen, ns);
                        }
                    } else {
                        s = w;
                        int x, y;
                        ok = sample_pixel(P, s, x, y);
                        if (ok) {
                            r = sample_ray(P, x, y);
                            dmax = GLM_INFINITY;  // trace ... ray infinity (Glome.hs:33)
                            if (segidx > 0) {     // continue the fold: best so far from earlier segments
                                best_seg = P.hit_seg[s];
                                has = best_seg >= 0;
                                best_t = has ? P.hit_t[s] : (Flt)GLM_INFINITY;
                                best_item = -1;
                            } else { has = false; best_t = GLM_INFINITY; best_item = -1; best_seg = -1; }
                        }
                    }
                    if (ok) {
                        bbclip_ub(r, bb, near_, far_);
                        far_ = fmin_(dmax, far_);  // traverse root near (fmin d far)
                        if (near_ < 0) near_ = 0;  // origin clamp: nothing behind the origin can be hit
                        drx = 1 / r.d.x; dry = 1 / r.d.y; drz = 1 / r.d.z;
                        ref = bn.a; sp = 0; sb = 0;
                        n_start = n_bih;
                        active = true;
                        if (ref >= 0 && near_ > far_) {  // Bih.hs:347 at the root: miss
                            if (!ANY && segidx == 0) {
                                P.hit_t[s] = GLM_INFINITY; P.hit_seg[s] = -1; P.hit_item[s] = -1; P.hit_sub[s] = -1; P.hit_flags[s] = 0;
                            }
                            active = false;
                        }
                    }
                }
            }
        }
#if GW_STEAL
        // ---- drain phase: the sample list is exhausted, so an idle lane can only help a neighbour.
        // A few rays cross the whole cloud (p50 = 64 node visits, max > 1000) and at ~0.4 us of dependent
        // latency per visit one such ray alone costs 0.3-0.5 ms, which bounded small waves (AA passes,
        // 1/N of a frame per GPU).  A ray's pushed subtrees are independent given (ray, near, far), so a
        // busy lane donates the oldest entry of its stack (the subtree nearest the root) to an idle lane
        // of its warp, which traverses it with the donor's ray.  Every leaf still sees the (ray, far) of
        // the sequential walk, so the union of the partial results is the sequential result -- except for the
        // reference's tie rule (`nearest`: of two hits at exactly the same depth the one met LATER by the sequential
        // walk wins, Solid.hs:37-44, Bih.hs:350-366): members finish in any order.  So the fold only NOTES a tie
        // between different items; the lane that completes such a group walks the ray once more, alone and in
        // order (solo), and that result is the one written.
        // members of a group share their nearest depth every round, so that a subtree behind a neighbour's hit is culled
        if (!ANY && nomore && active && light >= 0) {
            if (has) atomicMin(&g_cull[gb + light], flt_key(best_t));
            const Flt c = key_flt(g_cull[gb + light]);
            if (c < best_t) { best_t = c; has = true; best_item = -1; }  // own hit is dominated: drop it
        }
        if (nomore && idle) {
            if (!ANY) {
                unsigned int fin = __ballot_sync(FULL, !active && light >= 0);  // finished members: fold in
                while (fin) {
                    const int l = __ffs(fin) - 1;
                    fin &= fin - 1;
                    if (lane == l) {
                        const int g = gb + light;
                        if (best_item >= 0 && g_item[g] >= 0 && g_item[g] != best_item && g_t[g] == best_t) g_tie[g] = 1;
                        if (best_item >= 0 && (g_item[g] < 0 || !(g_t[g] < best_t))) { g_t[g] = best_t; g_item[g] = best_item; }
                        g_ovf[g] |= (int)n_ovf;
                        n_ovf = 0;
                        const int left = g_pend[g] - 1;
                        g_pend[g] = left;
                        if (left == 0 && g_tie[g]) {
                            // the order of arrival decided between equal depths: redo this ray sequentially
                            // (every member carries the group's ray, dmax and sample)
                            s = g_s[g];
                            if (segidx > 0) {
                                best_seg = P.hit_seg[s];
                                has = best_seg >= 0;
                                best_t = has ? P.hit_t[s] : (Flt)GLM_INFINITY;
                            } else { has = false; best_t = GLM_INFINITY; best_seg = -1; }
                            best_item = -1;
                            bbclip_ub(r, bb, near_, far_);
                            far_ = fmin_(dmax, far_);
                            if (near_ < 0) near_ = 0;
                            ref = bn.a; sp = 0; sb = 0;
                            n_ovf = (unsigned int)(g_ovf[g] != 0);
                            solo = true;
                            active = true;
                        } else if (left == 0) {
                            const long long ss = g_s[g];
                            const bool found = g_item[g] >= 0;
                            if (found) { P.hit_t[ss] = g_t[g]; P.hit_seg[ss] = segidx; P.hit_item[ss] = g_item[g]; P.hit_sub[ss] = -1; }
                            else if (segidx == 0) { P.hit_t[ss] = GLM_INFINITY; P.hit_seg[ss] = -1; P.hit_item[ss] = -1; P.hit_sub[ss] = -1; }
                            if (segidx == 0) P.hit_flags[ss] = g_ovf[g] ? GLOME_HITFLAG_STACK_OVERFLOW : 0;
                            else if (g_ovf[g]) P.hit_flags[ss] |= GLOME_HITFLAG_STACK_OVERFLOW;
                        }
                        light = -1;
                    }
                    __syncwarp();
                }
            } else if (active && ((__ldcg(P.occl + s) >> light) & 1u)) {
                active = false;  // another lane working on this ray already found an occluder
                has = false;
            }
            const unsigned int free_ = __ballot_sync(FULL, !active);
            // GW_STEAL_AFTER > 0 lets only rays that already walked that many nodes donate (a typical ray ends at a near
            // hit that culls its pending subtrees, so walking those in parallel is wasted work).  Measured on B200: the
            // greedy setting (0) does up to 2x the node visits on small waves and is still the fastest (720x480 AA:
            // 4.26 ms vs 4.70 ms at 128 and 5.26 ms at 256), because those waves are latency-bound, not issue-bound.
            const bool can_give = active && !solo && sp > sb && (GW_STEAL_AFTER == 0 || (n_bih - n_start) >= GW_STEAL_AFTER);
            const unsigned int donors = __ballot_sync(FULL, can_give);
            const int np = min(__popc(free_), __popc(donors));
            if (np > 0) {
                const unsigned int lt = (1u << lane) - 1;
                const bool is_thief = !active && __popc(free_ & lt) < np;
                const bool is_donor = can_give && __popc(donors & lt) < np;
                int eref = 0;
                Flt en = 0, ef = 0;
                if (is_donor) {
                    eref = stack[sb].ref; en = stack[sb].near_; ef = stack[sb].far_;
                    sb++;
                    if (!ANY) {
                        if (light < 0) {  // first donation: this lane becomes the root of a group
                            light = lane;
                            g_t[gb + lane] = GLM_INFINITY; g_item[gb + lane] = -1; g_ovf[gb + lane] = 0; g_s[gb + lane] = s;
                            g_tie[gb + lane] = 0;
                            g_pend[gb + lane] = 1;
                            g_cull[gb + lane] = flt_key(has ? best_t : (Flt)GLM_INFINITY);
                        }
                    }
                }
                __syncwarp();
                if (!ANY && is_donor) atomicAdd(&g_pend[gb + light], 1);  // the thief's membership
                const int src = is_thief ? (int)__fns(donors, 0, __popc(free_ & lt) + 1) : lane;
                eref = __shfl_sync(FULL, eref, src);
                en = __shfl_sync(FULL, en, src);
                ef = __shfl_sync(FULL, ef, src);
                const Flt ox = __shfl_sync(FULL, r.o.x, src), oy = __shfl_sync(FULL, r.o.y, src), oz = __shfl_sync(FULL, r.o.z, src);
                const Flt dx = __shfl_sync(FULL, r.d.x, src), dy = __shfl_sync(FULL, r.d.y, src), dz = __shfl_sync(FULL, r.d.z, src);
                const Flt dm = __shfl_sync(FULL, dmax, src);
                const long long ss = __shfl_sync(FULL, s, src);
                const int lg = __shfl_sync(FULL, light, src);
                const Flt bt = __shfl_sync(FULL, best_t, src);
                const int hs = __shfl_sync(FULL, (int)has, src);
                const int bs = __shfl_sync(FULL, best_seg, src);
                if (is_thief) {
                    r = mkray(vec(ox, oy, oz), vec(dx, dy, dz));
                    dmax = dm; s = ss; light = lg;
                    drx = 1 / dx; dry = 1 / dy; drz = 1 / dz;
                    ref = eref; near_ = en; far_ = ef;
                    sp = 0; sb = 0;
                    n_start = n_bih - GW_STEAL_AFTER;  // a piece of a long ray may be split further at once
                    if (ANY) has = false;
                    else { has = hs != 0; best_t = bt; best_seg = bs; best_item = -1; }  // the donor's best only culls
                    active = true;
                }
                __syncwarp();
            }
        }
#endif
        if (__ballot_sync(FULL, active) == 0) {
            if (nomore) break;
            continue;
        }
        // ---- one scheduling round.  GW_POLICY 0: while-while (all lanes walk branches until every
        // lane sits on a leaf, then all do leaves).  GW_POLICY 1: majority vote (the warp executes the
        // step kind -- branch or leaf -- that more of its lanes are waiting for). ----
        bool done = false;
#if GW_POLICY == 1
        const bool atb = active && ref >= 0;
        const unsigned int mb = __ballot_sync(FULL, atb);
        const unsigned int ml = __ballot_sync(FULL, active && ref < 0);
        const bool do_branch = __popc(mb) >= __popc(ml);
        for (int rep = 0; rep < GW_BRANCH_REPS && do_branch && active && !done && ref >= 0; rep++) {
#else
        // while-while with a BOUNDED branch phase: at most P.branch_reps steps, then the lanes that sit on a leaf do it while
        // the others keep their place -- in a deep tree (10^6 small spheres) no lane should wait for the deepest descent of
        // its warp (8 steps: config 1 4.09 -> 3.79 ms; 6 steps with the refill threshold at 12 lanes: 3.72 -> 3.58 ms); a shallow tree of large spheres is better off unbounded
        // (configs[4]: 6.53 vs 6.68 ms), so the host picks the variant per Bih (glome_cuda.cu: launch_wave).  The bound is a
        // compile-time constant: as a kernel argument it gives half of the gain.
        const bool do_branch = false;
        for (int rep = 0; (!BOUNDED || rep < GW_BRANCH_BOUND) && active && !done && ref >= 0; rep++) {
#endif
#ifndef GW_NOCOUNT
            n_bih++;
#endif
            const BihStep bs_ = ld_bih(S.bih, ref);
            const Flt2 sp2 = {bs_.ls, bs_.rs};
            int4 ii; ii.x = bs_.axis; ii.y = bs_.left; ii.z = bs_.right; ii.w = 0;
            Flt dr_ = (ii.x == 0) ? drx : ((ii.x == 1) ? dry : drz);
            Flt o = (ii.x == 0) ? r.o.x : ((ii.x == 1) ? r.o.y : r.o.z);
            Flt dl = (sp2.x - o) * dr_;
            Flt dr = (sp2.y - o) * dr_;
            // (the reference's `near > far -> miss` test, Bih.hs:347, can only fire at the root: children are
            //  entered with near <= far by construction; the root case is handled when the ray is set up)
            // branch-free form of Bih.hs:350-366: the near child is the left one iff dirr > 0
            const bool fwd = dr_ > 0;
            const Flt dn = fwd ? dl : dr;   // plane bounding the near child
            const Flt df = fwd ? dr : dl;   // plane bounding the far child
            const int c1 = fwd ? ii.y : ii.z;
            const int c2 = fwd ? ii.z : ii.y;
            const bool v1 = near_ < dn;
            const Flt f1 = fmin_(dn, far_);
            bool v2 = df < far_;
            const Flt n2 = fmax_(df, near_);
            if (!ANY && v2 && has && n2 > best_t) v2 = false;  // best-hit culling
            if (v1 && v2) {
                if (sp < GW_STACK) { stack[sp].ref = c2; stack[sp].near_ = n2; stack[sp].far_ = far_; sp++; }
                else n_ovf = 1;
            }
            if (v1) { ref = c1; far_ = f1; }
            else if (v2) { ref = c2; near_ = n2; }
            else {
                for (;;) {
                    if (sp == sb) { done = true; break; }
                    sp--;
                    ref = stack[sp].ref; near_ = stack[sp].near_; far_ = stack[sp].far_;
                    if (ANY || !(has && near_ > best_t)) break;
                }
            }
        }
        // ---- leaf ----
        if (!do_branch && active && !done && ref < 0) {
            int2 lf;
            glome_bih_leaf(ref, S.ipool, &lf.x, &lf.y);
            Flt dd = ANY ? fmin_(dmax, far_) : far_;  // Bih.hs:515 / :339
            for (int i = 0; i < lf.y; i++) {
                int item = lf.x + i;
                if (LINEAR) {
#ifndef GW_NOCOUNT
                    n_prim++;
#endif
                    const Flt* sph = S.dpool + a0 + 4 * (item - j0);
                    if (ANY) {
                        if (shadow_sphere(sph, r, dd)) { has = true; break; }
                    } else {
                        Flt t; Vec pp, nn;
                        if (prim_sphere<false>(sph, r, dd, t, pp, nn) && (!has || !(best_t < t))) {
                            has = true; best_t = t; best_item = item; best_seg = segidx;
                        }
                    }
                } else {
                    // {Tex,Tag}* prim leaf item
                    GlomeNode nd;
                    int prim;
                    if (!unwrap_prim<ANY>(S, item, nd, prim)) continue;
                    n_prim++;
                    if (ANY) {
                        if (prim_shadow(S, nd, r, dd)) { has = true; break; }
                    } else {
                        Flt t; Vec pp, nn;
                        if (prim_rayint<false>(S, nd, r, dd, t, pp, nn) && (!has || !(best_t < t))) {
                            has = true; best_t = t; best_item = item; best_seg = segidx;
                        }
                    }
                }
            }
            if (ANY && has) done = true;
            else {
                for (;;) {
                    if (sp == sb) { done = true; break; }
                    sp--;
                    ref = stack[sp].ref; near_ = stack[sp].near_; far_ = stack[sp].far_;
                    if (ANY || !(has && near_ > best_t)) break;
                }
            }
        }
        if (active && done) {
            if (ANY) {
                if (has) atomicOr(P.occl + s, 1u << light);
                has = false;
            } else if (!GW_STEAL || light < 0) {  // (a member of a drain-phase group is folded in at the top of the loop)
                if (segidx == 0 || best_seg == segidx) {
                    P.hit_t[s] = has ? best_t : (Flt)GLM_INFINITY;
                    P.hit_seg[s] = has ? best_seg : -1;
                    P.hit_item[s] = best_item;
                    P.hit_sub[s] = -1;
                }
                if (segidx == 0) P.hit_flags[s] = n_ovf ? GLOME_HITFLAG_STACK_OVERFLOW : 0;
                else if (n_ovf) P.hit_flags[s] |= GLOME_HITFLAG_STACK_OVERFLOW;
                n_ovf = 0;
#if GW_STEAL
                solo = false;
#endif
            }
            active = false;
        }
    }
    __syncwarp();
    unsigned int vals[9] = {0, 0, 0, 0, 0, n_bih, n_prim, 0, 0};
    wave_flush(P.stats, vals);
}

// bbclip_ub_rcp (Vec.hs:725-741) with the sign tests of the reciprocal known at compile time (OCT bit k: rcp.k > 0);
// OCT == 8: tested per lane
template <int OCT>
__device__ __forceinline__ void bvh_clip(const Vec& o, const Vec& rcp, const Bbox& b, Flt& near_, Flt& far_) {
    Flt inx, outx, iny, outy, inz, outz;
    slab(OCT == 8 ? rcp.x > 0 : (OCT & 1) != 0, b.p1.x, b.p2.x, o.x, rcp.x, inx, outx);
    slab(OCT == 8 ? rcp.y > 0 : (OCT & 2) != 0, b.p1.y, b.p2.y, o.y, rcp.y, iny, outy);
    slab(OCT == 8 ? rcp.z > 0 : (OCT & 4) != 0, b.p1.z, b.p2.z, o.z, rcp.z, inz, outz);
    near_ = fmax3(inx, iny, inz);
    far_ = fmin3(outx, outy, outz);
}

// The branch phase of rayint_mesh's walk (Mesh.hs:166-198): descend until this lane sits on a leaf or its walk is over.
// Both children pass the same entry test.  The reference writes it as `near > far || near > depth || far < 0` for the
// child met first and, for the other, the same with far' = min far (ridepth firstresult) (Mesh.hs:178-182); with the walk's
// best hit so far as the cull depth (best = 10^6 = depth when there is none; a hit's depth is never negative)
//   n > min f best  <=>  n > f || n > best      and      min f best < 0  <=>  f < 0,
// so one symmetric test per child serves both, and only the choice of the child to descend into needs their order.
template <int OCT>
__device__ __forceinline__ void bvh_branch_phase(const DScene& S, const Ray& r, const Vec& rcp, bool walks, bool has, Flt best_t,
                                                 Flt depth, TravEnt* stack, int& sp, int sb, int& ref, Flt& near_, Flt& far_,
                                                 bool& done, unsigned int& n_bvh, unsigned int& n_ovf) {
    const Flt best = has ? best_t : (Flt)GLM_INFINITY;
    // (GW_BVH_BOUND > 0: at most that many steps, then the lanes that sit on a leaf do it while the others keep their place)
    for (int rep = 0; (GW_BVH_BOUND == 0 || rep < GW_BVH_BOUND) && walks && !done && ref >= 0; rep++) {
        n_bvh++;
        Bbox lbb_, rbb_;
        int2 kids;
        ld_bvh(S.bvh, ref, lbb_, rbb_, kids.x, kids.y);
#if GW_BVH_PREFETCH == 1
        if (kids.x >= 0) asm volatile("prefetch.global.L2 [%0];" ::"l"(S.bvh + kids.x));
        if (kids.y >= 0) asm volatile("prefetch.global.L2 [%0];" ::"l"(S.bvh + kids.y));
#elif GW_BVH_PREFETCH == 2
        // the node of whichever child the two box tests choose is one L2 round trip away: start both lines towards L1 now
        if (kids.x >= 0) asm volatile("prefetch.global.L1 [%0];" ::"l"(S.bvh + kids.x));
        if (kids.y >= 0) asm volatile("prefetch.global.L1 [%0];" ::"l"(S.bvh + kids.y));
#endif
        Flt lnearp, lfarp, rnearp, rfarp;
        bvh_clip<OCT>(r.o, rcp, lbb_, lnearp, lfarp);
        bvh_clip<OCT>(r.o, rcp, rbb_, rnearp, rfarp);
        const Flt lnear = hmax(near_, lnearp), lfar = hmin(far_, lfarp);
        const Flt rnear = hmax(near_, rnearp), rfar = hmin(far_, rfarp);
        const bool vl = !(lnear > lfar || lnear > depth || lfar < 0 || lnear > best);
        const bool vr = !(rnear > rfar || rnear > depth || rfar < 0 || rnear > best);
        const bool lfirst = lnear < rnear;
        if (vl && vr) {
            if (sp < GW_STACK) {
                stack[sp].ref = lfirst ? kids.y : kids.x; stack[sp].near_ = lfirst ? rnear : lnear; stack[sp].far_ = lfirst ? rfar : lfar;
                sp++;
            } else n_ovf = 1;
        }
        if (vl || vr) {
            const bool gol = vl && (!vr || lfirst);
            ref = gol ? kids.x : kids.y; near_ = gol ? lnear : rnear; far_ = gol ? lfar : rfar;
        } else {
            for (;;) {
                if (sp == sb) { done = true; break; }
                sp--;
                ref = stack[sp].ref; near_ = stack[sp].near_; far_ = stack[sp].far_;
                const Flt fc = hmin(far_, best);
                if (!(near_ > fc || fc < 0)) break;
            }
        }
    }
}

// ---------------------------------------------------------------------------------------------
// K1 for a Mesh segment (rayint_mesh, Mesh.hs:136-198): persistent, per-lane refill.
// A Mesh casts no shadows (Mesh.hs:210), so there is no any-hit variant.
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128, GW_BVH_MINBLOCKS) k_bvh_closest(DScene S, WaveParams P, int segidx, Seg seg, unsigned int* counter) {
    const unsigned int FULL = 0xffffffffu;
    const int lane = threadIdx.x & 31;
    const GlomeNode mn = S.nodes[seg.node];
    const GlomeMeshHeader* h = reinterpret_cast<const GlomeMeshHeader*>(S.ipool + mn.a);
    const int verts_off = h->verts_off, tris_off = h->tris_off, root = h->root;
    const Bbox bb = ldbb(S.dpool + h->bb_off);
    const long long total = wave_total(P);

    TravEnt stack[GW_STACK];
    int sp = 0, ref = 0;
    bool active = false, nomore = false;
    long long s = 0;
    Ray r = mkray(vec(0, 0, 0), vec(0, 0, 1));
    Vec rcp = vec(0, 0, 0);
    int oct = 0;  // bit k: rcp.k > 0 (the sign tests of bbclip_ub_rcp)
    Flt near_ = 0, far_ = 0;
    const Flt depth = GLM_INFINITY;
    Flt best_t = GLM_INFINITY;
    int best_sub = -1, best_seg = -1;
    bool has = false;
    unsigned int n_bvh = 0, n_tri = 0, n_ovf = 0;
#if GW_GUIDED
    unsigned int next_base = 0;
    const long long nwarps = (long long)gridDim.x * 4;
#endif
#if GW_BVH_STEAL
    // drain-phase groups, as in k_bih_traverse: the lanes of one warp that work on the same ray fold their partial
    // results here; the lane that brings g_pend to 0 writes the sample's result
    __shared__ Flt g_t[128];
    __shared__ int g_sub[128], g_pend[128], g_ovf[128], g_tie[128];
    __shared__ long long g_s[128];
    __shared__ unsigned long long g_cull[128];
    const int gb = threadIdx.x & ~31;
    int grp = -1;       // root lane of the group this lane is a member of
    int sb = 0;         // live stack entries are [sb, sp): the drain phase donates from the bottom
    bool solo = false;  // this lane re-walks a ray alone (after a tie) and must not donate
#else
    const int sb = 0;
#endif

    for (;;) {
        unsigned int idle = __ballot_sync(FULL, !active);
        if (__popc(idle) >= GW_BVH_REFILL_MIN && !nomore) {
            unsigned int base = 0;
            int cnt = __popc(idle);
#if GW_GUIDED
            cnt = min(cnt, (int)min((long long)GW_MAXBATCH, max((long long)GW_BVH_GUIDED_MIN, (total - (long long)next_base) / ((long long)GW_BVH_GUIDED_DIV * nwarps) + 1)));
#endif
            int leader = __ffs(idle) - 1;
            if (lane == leader) base = atomicAdd(counter, (unsigned int)cnt);
            base = __shfl_sync(FULL, base, leader);
            if ((long long)base + cnt >= total) nomore = true;
#if GW_GUIDED
            next_base = base + (unsigned int)cnt;
#endif
            const int rank = __popc(idle & ((1u << lane) - 1));
            if (!active && rank < cnt) {
                long long w = (long long)base + rank;
                if (w < total) {
                    s = w;
                    int x, y;
                    if (sample_pixel(P, s, x, y)) {
                        r = sample_ray(P, x, y);
                        if (segidx > 0) {
                            best_seg = P.hit_seg[s];
                            has = best_seg >= 0;
                            best_t = has ? P.hit_t[s] : (Flt)GLM_INFINITY;
                        } else { has = false; best_t = GLM_INFINITY; best_seg = -1; }
                        best_sub = -1;
                        rcp = vrcp(r.d);
                        oct = (rcp.x > 0 ? 1 : 0) | (rcp.y > 0 ? 2 : 0) | (rcp.z > 0 ? 4 : 0);
                        bbclip_ub_rcp(r.o, rcp, bb, near_, far_);
                        ref = root; sp = 0;
#if GW_BVH_STEAL
                        sb = 0;
#endif
                        active = true;
                        // Mesh.hs:140: root reject (the fold's earlier results stay as they are)
                        if (near_ > far_ || near_ > depth || far_ < 0) {
                            if (segidx == 0) { P.hit_t[s] = GLM_INFINITY; P.hit_seg[s] = -1; P.hit_item[s] = -1; P.hit_sub[s] = -1; P.hit_flags[s] = 0; }
                            active = false;
                        }
                    }
                }
            }
        }
#if GW_BVH_STEAL
        // ---- drain phase (see k_bih_traverse): the sample list is exhausted, so an idle lane can only help a neighbour.
        // A ray that grazes the surface walks hundreds of nodes at ~0.4 us of dependent latency each, and a small wave (an AA
        // pass, 1/N of a frame) sat at 0.2-0.3 ms = its longest ray.  A pushed subtree is independent given (ray, near, far):
        // the walk's best hit only culls.  So a busy lane donates the oldest entry of its stack to an idle lane of its warp.
        // The walk order only matters between two different triangles at exactly the same depth (the later one wins,
        // Mesh.hs:172-176 / Solid.hs:37-44): the fold notes such a tie and the lane that completes the group walks the ray
        // once more, alone and in order.
        if (nomore && active && grp >= 0) {
            if (has) atomicMin(&g_cull[gb + grp], flt_key(best_t));
            const Flt c = key_flt(g_cull[gb + grp]);
            if (c < best_t) { best_t = c; has = true; best_sub = -1; }  // own hit is dominated: drop it
        }
        if (nomore && idle) {
            unsigned int fin = __ballot_sync(FULL, !active && grp >= 0);  // finished members: fold in
            while (fin) {
                const int l = __ffs(fin) - 1;
                fin &= fin - 1;
                if (lane == l) {
                    const int g = gb + grp;
                    if (best_sub >= 0 && g_sub[g] >= 0 && g_sub[g] != best_sub && g_t[g] == best_t) g_tie[g] = 1;
                    if (best_sub >= 0 && (g_sub[g] < 0 || !(g_t[g] < best_t))) { g_t[g] = best_t; g_sub[g] = best_sub; }
                    g_ovf[g] |= (int)n_ovf;
                    n_ovf = 0;
                    const int left = g_pend[g] - 1;
                    g_pend[g] = left;
                    if (left == 0 && g_tie[g]) {
                        // the order of arrival decided between equal depths: redo this ray sequentially
                        s = g_s[g];
                        if (segidx > 0) {
                            best_seg = P.hit_seg[s];
                            has = best_seg >= 0;
                            best_t = has ? P.hit_t[s] : (Flt)GLM_INFINITY;
                        } else { has = false; best_t = GLM_INFINITY; best_seg = -1; }
                        best_sub = -1;
                        bbclip_ub_rcp(r.o, rcp, bb, near_, far_);
                        ref = root; sp = 0; sb = 0;
                        n_ovf = (unsigned int)(g_ovf[g] != 0);
                        solo = true;
                        active = true;
                    } else if (left == 0) {
                        const long long ss = g_s[g];
                        if (g_sub[g] >= 0) { P.hit_t[ss] = g_t[g]; P.hit_seg[ss] = segidx; P.hit_item[ss] = seg.node; P.hit_sub[ss] = g_sub[g]; }
                        else if (segidx == 0) { P.hit_t[ss] = GLM_INFINITY; P.hit_seg[ss] = -1; P.hit_item[ss] = seg.node; P.hit_sub[ss] = -1; }
                        if (segidx == 0) P.hit_flags[ss] = g_ovf[g] ? GLOME_HITFLAG_STACK_OVERFLOW : 0;
                        else if (g_ovf[g]) P.hit_flags[ss] |= GLOME_HITFLAG_STACK_OVERFLOW;
                    }
                    grp = -1;
                }
                __syncwarp();
            }
            const unsigned int free_ = __ballot_sync(FULL, !active);
            const bool can_give = active && !solo && sp > sb;
            const unsigned int donors = __ballot_sync(FULL, can_give);
            const int np = min(__popc(free_), __popc(donors));
            if (np > 0) {
                const unsigned int lt = (1u << lane) - 1;
                const bool is_thief = !active && __popc(free_ & lt) < np;
                const bool is_donor = can_give && __popc(donors & lt) < np;
                int eref = 0;
                Flt en = 0, ef = 0;
                if (is_donor) {
                    eref = stack[sb].ref; en = stack[sb].near_; ef = stack[sb].far_;
                    sb++;
                    if (grp < 0) {  // first donation: this lane becomes the root of a group
                        grp = lane;
                        g_t[gb + lane] = GLM_INFINITY; g_sub[gb + lane] = -1; g_ovf[gb + lane] = 0; g_s[gb + lane] = s;
                        g_tie[gb + lane] = 0;
                        g_pend[gb + lane] = 1;
                        g_cull[gb + lane] = flt_key(has ? best_t : (Flt)GLM_INFINITY);
                    }
                }
                __syncwarp();
                if (is_donor) atomicAdd(&g_pend[gb + grp], 1);  // the thief's membership
                const int src = is_thief ? (int)__fns(donors, 0, __popc(free_ & lt) + 1) : lane;
                eref = __shfl_sync(FULL, eref, src);
                en = __shfl_sync(FULL, en, src);
                ef = __shfl_sync(FULL, ef, src);
                const Flt ox = __shfl_sync(FULL, r.o.x, src), oy = __shfl_sync(FULL, r.o.y, src), oz = __shfl_sync(FULL, r.o.z, src);
                const Flt dx = __shfl_sync(FULL, r.d.x, src), dy = __shfl_sync(FULL, r.d.y, src), dz = __shfl_sync(FULL, r.d.z, src);
                const long long ss = __shfl_sync(FULL, s, src);
                const int lg = __shfl_sync(FULL, grp, src);
                const Flt bt = __shfl_sync(FULL, best_t, src);
                const int hs = __shfl_sync(FULL, (int)has, src);
                const int bs = __shfl_sync(FULL, best_seg, src);
                if (is_thief) {
                    r = mkray(vec(ox, oy, oz), vec(dx, dy, dz));
                    rcp = vrcp(r.d);
                    oct = (rcp.x > 0 ? 1 : 0) | (rcp.y > 0 ? 2 : 0) | (rcp.z > 0 ? 4 : 0);
                    s = ss; grp = lg;
                    ref = eref; near_ = en; far_ = ef;
                    sp = 0; sb = 0;
                    has = hs != 0; best_t = bt; best_seg = bs; best_sub = -1;  // the donor's best only culls
                    active = true;
                }
                __syncwarp();
            }
        }
#endif
        if (__ballot_sync(FULL, active) == 0) {
            if (nomore) break;
            continue;
        }
        bool done = false;
        {
            // the branch phase, compiled once per sign pattern of the ray direction: when every lane that walks has the same
            // one (a camera's rays nearly always do), the slab tests know at compile time which plane of a box is the near one
            const bool walks = active && ref >= 0;
            const unsigned int wm = __ballot_sync(FULL, walks);
            if (wm) {
                const int o0 = __shfl_sync(FULL, oct, __ffs(wm) - 1);
                const bool uni = __all_sync(FULL, !walks || oct == o0);
                switch (uni ? o0 : 8) {
                    case 0: bvh_branch_phase<0>(S, r, rcp, walks, has, best_t, depth, stack, sp, sb, ref, near_, far_, done, n_bvh, n_ovf); break;
                    case 1: bvh_branch_phase<1>(S, r, rcp, walks, has, best_t, depth, stack, sp, sb, ref, near_, far_, done, n_bvh, n_ovf); break;
                    case 2: bvh_branch_phase<2>(S, r, rcp, walks, has, best_t, depth, stack, sp, sb, ref, near_, far_, done, n_bvh, n_ovf); break;
                    case 3: bvh_branch_phase<3>(S, r, rcp, walks, has, best_t, depth, stack, sp, sb, ref, near_, far_, done, n_bvh, n_ovf); break;
                    case 4: bvh_branch_phase<4>(S, r, rcp, walks, has, best_t, depth, stack, sp, sb, ref, near_, far_, done, n_bvh, n_ovf); break;
                    case 5: bvh_branch_phase<5>(S, r, rcp, walks, has, best_t, depth, stack, sp, sb, ref, near_, far_, done, n_bvh, n_ovf); break;
                    case 6: bvh_branch_phase<6>(S, r, rcp, walks, has, best_t, depth, stack, sp, sb, ref, near_, far_, done, n_bvh, n_ovf); break;
                    case 7: bvh_branch_phase<7>(S, r, rcp, walks, has, best_t, depth, stack, sp, sb, ref, near_, far_, done, n_bvh, n_ovf); break;
                    default: bvh_branch_phase<8>(S, r, rcp, walks, has, best_t, depth, stack, sp, sb, ref, near_, far_, done, n_bvh, n_ovf); break;
                }
            }
        }
        if (active && !done && ref < 0) {
            int k = ~ref;
            int ntri = __ldg(S.ipool + k);
            if (S.leafv) {  // the leaf's triangles with their vertices inline: no dependent index loads (DScene::leafv)
                const Flt* __restrict__ lv = S.leafv + 9 * (size_t)(k + 1 - S.leafv_base);
                for (int j = 0; j < ntri; j++, lv += 9) {
                    n_tri++;
                    const Vec a = vec(__ldg(lv), __ldg(lv + 1), __ldg(lv + 2));
                    const Vec b = vec(__ldg(lv + 3), __ldg(lv + 4), __ldg(lv + 5));
                    const Vec c = vec(__ldg(lv + 6), __ldg(lv + 7), __ldg(lv + 8));
                    Flt t; Vec pp, nn;
                    Vec z = vec(0, 0, 0);
                    if (prim_triangle<false>(a, b, c, false, z, z, z, r, far_, t, pp, nn) && (!has || !(best_t < t))) {
                        has = true; best_t = t; best_sub = __ldg(S.ipool + k + 1 + j); best_seg = segidx;
                    }
                }
            } else
            for (int j = 0; j < ntri; j++) {
                int ti = __ldg(S.ipool + k + 1 + j);
                n_tri++;
                int4 t0 = __ldg(reinterpret_cast<const int4*>(S.ipool + tris_off + 8 * ti));
                Vec a = ldv(S.dpool + verts_off + 3 * t0.x);
                Vec b = ldv(S.dpool + verts_off + 3 * t0.y);
                Vec c = ldv(S.dpool + verts_off + 3 * t0.z);
                Flt t; Vec pp, nn;
                Vec z = vec(0, 0, 0);
                if (prim_triangle<false>(a, b, c, false, z, z, z, r, far_, t, pp, nn) && (!has || !(best_t < t))) {
                    has = true; best_t = t; best_sub = ti; best_seg = segidx;
                }
            }
            for (;;) {
                if (sp == sb) { done = true; break; }
                sp--;
                ref = stack[sp].ref; near_ = stack[sp].near_; far_ = stack[sp].far_;
                Flt fc = hmin(far_, has ? best_t : (Flt)GLM_INFINITY);
                if (!(near_ > fc || fc < 0)) break;
            }
        }
#if GW_BVH_STEAL
        if (active && done && grp < 0) {  // (a member of a drain-phase group is folded in at the top of the loop)
            solo = false;
#else
        if (active && done) {
#endif
            if (segidx == 0 || (best_seg == segidx && best_sub >= 0)) {
                P.hit_t[s] = has ? best_t : (Flt)GLM_INFINITY;
                P.hit_seg[s] = has ? best_seg : -1;
                P.hit_item[s] = seg.node;
                P.hit_sub[s] = best_sub;
            }
            if (segidx == 0) P.hit_flags[s] = n_ovf ? GLOME_HITFLAG_STACK_OVERFLOW : 0;
            else if (n_ovf) P.hit_flags[s] |= GLOME_HITFLAG_STACK_OVERFLOW;
            n_ovf = 0;
        }
        if (active && done) active = false;
    }
    __syncwarp();
    unsigned int vals[9] = {0, 0, 0, 0, 0, 0, 0, n_bvh, n_tri};
    wave_flush(P.stats, vals);
}

// The list fold between two segments that were walked side by side (Solid.hs:327-328: foldl' nearest; of two hits at the
// same depth the later element's wins, Solid.hs:37-44): segment `seg_b`'s own result set is folded into the wave's.
__global__ void __launch_bounds__(128) k_merge_hits(WaveParams P, const Flt* __restrict__ bt, const int* __restrict__ bseg,
                                                    const int* __restrict__ bitem, const int* __restrict__ bsub,
                                                    const int* __restrict__ bflags, int seg_b) {
    const long long total = wave_total(P);
    for (long long s = blockIdx.x * (long long)blockDim.x + threadIdx.x; s < total; s += (long long)gridDim.x * blockDim.x) {
        int x, y;
        if (!sample_pixel(P, s, x, y)) continue;
        const int fl = bflags[s];
        if (fl) P.hit_flags[s] |= fl;
        if (bseg[s] < 0) continue;
        const Flt t = bt[s];
        if (P.hit_seg[s] >= 0 && P.hit_t[s] < t) continue;  // the earlier segment's hit is strictly nearer
        P.hit_t[s] = t; P.hit_seg[s] = seg_b; P.hit_item[s] = bitem[s]; P.hit_sub[s] = bsub[s];
    }
}

// K1 for a run of loose primitives (group children that are wrapped primitives), list fold order
__global__ void __launch_bounds__(128) k_prims_closest(DScene S, WaveParams P, int segidx, Seg seg) {
    const long long total = wave_total(P);
    unsigned int n_prim = 0;
    for (long long s = blockIdx.x * (long long)blockDim.x + threadIdx.x; s < total; s += (long long)gridDim.x * blockDim.x) {
        int x, y;
        if (!sample_pixel(P, s, x, y)) continue;
        Ray r = sample_ray(P, x, y);
        bool has = false;
        Flt best_t = GLM_INFINITY;
        int best_item = -1, best_seg = -1;
        if (segidx > 0) { best_seg = P.hit_seg[s]; has = best_seg >= 0; best_t = has ? P.hit_t[s] : (Flt)GLM_INFINITY; }
        for (int i = 0; i < seg.count; i++) {
            GlomeNode nd;
            int prim;
            if (!unwrap_prim<false>(S, seg.node + i, nd, prim)) continue;
            n_prim++;
            Flt t; Vec pp, nn;
            if (prim_rayint<false>(S, nd, r, GLM_INFINITY, t, pp, nn) && (!has || !(best_t < t))) {
                has = true; best_t = t; best_item = seg.node + i; best_seg = segidx;
            }
        }
        if (segidx == 0 || best_seg == segidx) {
            P.hit_t[s] = has ? best_t : (Flt)GLM_INFINITY;
            P.hit_seg[s] = has ? best_seg : -1;
            P.hit_item[s] = best_item;
            P.hit_sub[s] = -1;
        }
        if (segidx == 0) P.hit_flags[s] = 0;
    }
    __syncwarp();
    unsigned int vals[9] = {0, 0, 0, 0, 0, 0, n_prim, 0, 0};
    wave_flush(P.stats, vals);
}
__global__ void __launch_bounds__(128) k_prims_any(DScene S, WaveParams P, Seg seg) {
    const long long total = *P.squeue_count;
    unsigned int n_prim = 0;
    for (long long w = blockIdx.x * (long long)blockDim.x + threadIdx.x; w < total; w += (long long)gridDim.x * blockDim.x) {
        int2 e = P.squeue[w];
        if ((P.occl[e.x] >> e.y) & 1u) continue;
        Ray r; Flt d, llen; Vec ldir; bool ns;
        if (!light_ray(S, P.surf + 6 * (long long)e.x, e.y, r, d, ldir, llen, ns)) continue;
        for (int i = 0; i < seg.count; i++) {
            GlomeNode nd;
            int prim;
            if (!unwrap_prim<true>(S, seg.node + i, nd, prim)) continue;
            n_prim++;
            if (prim_shadow(S, nd, r, d)) { atomicOr(P.occl + e.x, 1u << e.y); break; }
        }
    }
    __syncwarp();
    unsigned int vals[9] = {0, 0, 0, 0, 0, 0, n_prim, 0, 0};
    wave_flush(P.stats, vals);
}

// Rebuild the hit's texture / tag stacks: item wrappers (innermost first), then the segment's.
__device__ __forceinline__ void rebuild_stacks(const DScene& S, const Seg* segs, int seg, int item, int sub, Stk& tex, Stk& tag,
                                               int& prim, int& flags) {
    tex = segs[seg].tex;
    tag = segs[seg].tag;
    prim = item;
    if (segs[seg].kind == SEG_MESH) {
        const GlomeNode mn = S.nodes[item];
        const GlomeMeshHeader* h = reinterpret_cast<const GlomeMeshHeader*>(S.ipool + mn.a);
        const int32_t* T = S.ipool + h->tris_off + 8 * sub;
        if (T[6] != -1) { Stk t2; if (stk_cons(t2, S.ipool[h->texs_off + T[6]], tex)) flags |= GLOME_HITFLAG_STACK_OVERFLOW; tex = t2; }
        if (T[7] != -1) { Stk t2; if (stk_cons(t2, S.ipool[h->tags_off + T[7]], tag)) flags |= GLOME_HITFLAG_STACK_OVERFLOW; tag = t2; }
        return;
    }
    int ni = item;
    GlomeNode nd = S.nodes[ni];
    while (nd.type == GLOME_TEX || nd.type == GLOME_TAG || nd.type == GLOME_NOSHADOW) {
        if (nd.type == GLOME_TEX) { Stk t2; if (stk_cons(t2, nd.b, tex)) flags |= GLOME_HITFLAG_STACK_OVERFLOW; tex = t2; }
        else if (nd.type == GLOME_TAG) { Stk t2; if (stk_cons(t2, nd.b, tag)) flags |= GLOME_HITFLAG_STACK_OVERFLOW; tag = t2; }
        ni = nd.a;
        nd = S.nodes[ni];
    }
    prim = ni;
}

// materialise the full Hit of sample s from the wave buffers (same arithmetic as finalize_flat)
__device__ __forceinline__ void load_hit(const DScene& S, const WaveParams& P, const Seg* segs, long long s, const Ray& r, Hit& h) {
    hit_clear(h);
    h.flags = P.hit_flags[s];
    int seg = P.hit_seg[s];
    if (seg < 0) return;
    h.hit = 1;
    h.t = P.hit_t[s];
    h.sub = P.hit_sub[s];
    int fl = 0;
    rebuild_stacks(S, segs, seg, P.hit_item[s], h.sub, h.tex, h.tag, h.prim, fl);
    h.flags |= fl;
    h.ray = r;
    const Flt* sf = P.surf + 6 * s;
    h.pos = vec(sf[0], sf[1], sf[2]);
    h.norm = vec(sf[3], sf[4], sf[5]);
}

// K2a: surface point of every hit + shadow-ray queue (mpreshade's per-light tests, Shader.hs:65-80)
#ifndef GW_SURFACE_MINBLOCKS
#define GW_SURFACE_MINBLOCKS 8  /* 64 regs, 32 warps per SM (88 regs / 20 warps) */
#endif
__global__ void __launch_bounds__(128, GW_SURFACE_MINBLOCKS) k_surface(DScene S, WaveParams P, const Seg* __restrict__ segs) {
    const unsigned int FULL = 0xffffffffu;
    const int lane = threadIdx.x & 31;
    const long long total = wave_total(P);
    const int lfirst = S.lightsets[0], lcnt = S.lightsets[1];
    unsigned int n_shadow = 0;
    const long long stride = (long long)gridDim.x * blockDim.x;
    const long long iters = (total + stride - 1) / stride;
    for (long long it = 0; it < iters; it++) {
        long long s = it * stride + blockIdx.x * (long long)blockDim.x + threadIdx.x;
        bool live = false;
        if (s < total) {
            int x, y;
            if (sample_pixel(P, s, x, y)) {
                int seg = P.hit_seg[s];
                P.occl[s] = 0;
                if (seg >= 0) {
                    Ray r = sample_ray(P, x, y);
                    Hit h;
                    hit_clear(h);
                    h.hit = 1; h.t = P.hit_t[s]; h.sub = P.hit_sub[s];
                    int fl = 0;
                    rebuild_stacks(S, segs, seg, P.hit_item[s], h.sub, h.tex, h.tag, h.prim, fl);
                    if (fl) P.hit_flags[s] |= fl;
                    finalize_flat(S, r, h);
                    Flt* sf = P.surf + 6 * s;
                    sf[0] = h.pos.x; sf[1] = h.pos.y; sf[2] = h.pos.z;
                    sf[3] = h.norm.x; sf[4] = h.norm.y; sf[5] = h.norm.z;
                    live = h.tex.n > 0;  // ctxb is only forced when a texture is shaded (Trace.hs:63)
                }
            }
        }
        for (int li = 0; li < lcnt; li++) {
            bool emit = false;
            if (live) {
                Ray r; Flt d, llen; Vec ldir; bool ns = false;
                emit = light_ray(S, P.surf + 6 * s, lfirst + li, r, d, ldir, llen, ns) && ns;
            }
            unsigned int m = __ballot_sync(FULL, emit);
            if (m) {
                int leader = __ffs(m) - 1;
                int slot = 0;
                if (lane == leader) slot = atomicAdd(P.squeue_count, __popc(m));
                slot = __shfl_sync(FULL, slot, leader);
                if (emit) { P.squeue[slot + __popc(m & ((1u << lane) - 1))] = make_int2((int)s, lfirst + li); n_shadow++; }
            }
        }
    }
    __syncwarp();
    unsigned int vals[9] = {0, n_shadow, 0, 0, 0, 0, 0, 0, 0};
    wave_flush(P.stats, vals);
}

struct TCw { Flt r, g, b, a, d; };
__device__ __forceinline__ TCw ldtc(const double* __restrict__ buf, size_t pix) {
    const double* p = buf + pix * 5;
    TCw c; c.r = p[0]; c.g = p[1]; c.b = p[2]; c.a = p[3]; c.d = p[4];
    return c;
}
__device__ __forceinline__ TCw getcw(const double* __restrict__ v, int width, int xt, int yt, int tw, int th, int x, int y) {
    if ((x >= xt) && (x < xt + tw) && (y >= yt) && (y < yt + th)) return ldtc(v, (size_t)y * width + x);
    TCw c; c.r = 0; c.g = 0; c.b = 0; c.a = 0; c.d = GLM_INFINITY;
    return c;
}

// K2b: trace's texture fold + materialShader for Surface materials (Trace.hs:59-82, Shader.hs:82-105)
#ifndef GW_SHADE_MINBLOCKS
#define GW_SHADE_MINBLOCKS 4    /* 128 regs, 16 warps per SM (160 regs / 12 warps: configs[2] +2 %, configs[4] +1 %) */
#endif
__global__ void __launch_bounds__(128, GW_SHADE_MINBLOCKS) k_shade(DScene S, WaveParams P, const Seg* __restrict__ segs) {
    const long long total = wave_total(P);
    const int lfirst = S.lightsets[0], lcnt = S.lightsets[1];
    unsigned int n_primary = 0, n_ovf = 0, n_perlin = 0;
    for (long long s = blockIdx.x * (long long)blockDim.x + threadIdx.x; s < total; s += (long long)gridDim.x * blockDim.x) {
        int x, y;
        if (!sample_pixel(P, s, x, y)) continue;
        n_primary++;
        Ray ray = sample_ray(P, x, y);
        Hit ri;
        load_hit(S, P, segs, s, ray, ri);
        if (ri.flags) n_ovf++;
        ColorA colora = mkca(0, 0, 0, 0);
        if (ri.hit && P.recurs != 0) {
            LightSel ctx;
            ctx.done = 0; ctx.first = lfirst; ctx.count = lcnt; ctx.mask = 0;
            Vec eyedir = vinvert(ray.d);
            for (int i = 0; i < ri.tex.n; i++) {
                if (colora.a + GLM_DELTA >= 1) continue;  // opaque (Trace.hs:50)
                if (!ctx.done) {  // mpreshade with the shadow results of K1'
                    ctx.done = 1;
                    unsigned int oc = P.occl[s];
                    for (int li = 0; li < lcnt; li++) {
                        Ray sr; Flt d, llen; Vec ldir; bool ns;
                        if (!light_ray(S, P.surf + 6 * s, lfirst + li, sr, d, ldir, llen, ns)) continue;
                        if (ns && ((oc >> (lfirst + li)) & 1u)) continue;
                        ctx.mask |= 1ull << li;
                    }
                }
                MatVal m;
                eval_texture(S, ri.tex.v[i], ri.pos, m, n_perlin);
                ColorA colorb;
                mshade_flat(S, ctx, m, ri, eyedir, colorb);
                colora = cafold(colora, colorb);
            }
        }
        TCw col;
        col.r = colora.r; col.g = colora.g; col.b = colora.b; col.a = colora.a;
        col.d = (P.recurs != 0) ? ridepth(ri) : (Flt)GLM_INFINITY;
        size_t pix = (size_t)y * P.g.width + x;
        double* o = P.out + pix * 5;
        if (P.mode == 5) {
            int tx = x / P.g.bs, ty = y / P.g.bs;
            int xt = tx * P.g.bs, yt = ty * P.g.bs;
            int tw = min(P.g.bs, P.g.width - xt), th = min(P.g.bs, P.g.height - yt);
            TCw a = getcw(P.v, P.g.width, xt, yt, tw, th, x, y), b = getcw(P.v, P.g.width, xt, yt, tw, th, x, y + 1);
            TCw c = getcw(P.v, P.g.width, xt, yt, tw, th, x + 1, y + 1), d = getcw(P.v, P.g.width, xt, yt, tw, th, x + 1, y);
            bool lastx = x == xt + tw - 1, lasty = y == yt + th - 1;
            // pass5_combine (Glome.hs:309-316)
            TCw q;
            if (lastx && lasty) q = col;
            else {
                TCw m;
                if (lastx) { m.r = (a.r + b.r) * FL(0.5); m.g = (a.g + b.g) * FL(0.5); m.b = (a.b + b.b) * FL(0.5); m.a = (a.a + b.a) * FL(0.5); m.d = (a.d + b.d) * FL(0.5); }
                else if (lasty) { m.r = (a.r + d.r) * FL(0.5); m.g = (a.g + d.g) * FL(0.5); m.b = (a.b + d.b) * FL(0.5); m.a = (a.a + d.a) * FL(0.5); m.d = (a.d + d.d) * FL(0.5); }
                else { m.r = (a.r + b.r + c.r + d.r) * FL(0.25); m.g = (a.g + b.g + c.g + d.g) * FL(0.25); m.b = (a.b + b.b + c.b + d.b) * FL(0.25);
                       m.a = (a.a + b.a + c.a + d.a) * FL(0.25); m.d = (a.d + b.d + c.d + d.d) * FL(0.25); }
                q.r = (col.r + m.r) * FL(0.5); q.g = (col.g + m.g) * FL(0.5); q.b = (col.b + m.b) * FL(0.5); q.a = (col.a + m.a) * FL(0.5); q.d = (col.d + m.d) * FL(0.5);
            }
            col = q;
        } else if (P.mode == 0 && P.tint) {
            col.r = col.r + (col.d / 400);  // Glome.hs:174
        }
        o[0] = col.r; o[1] = col.g; o[2] = col.b; o[3] = col.a; o[4] = col.d;
    }
    __syncwarp();
    unsigned int vals[9] = {n_primary, 0, 0, n_ovf, n_perlin, 0, 0, 0, 0};
    wave_flush(P.stats, vals);
}

}  // namespace gwave
