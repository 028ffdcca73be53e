// glome_build.cu -- the reference's BIH builder (`bih` / `build_rec`, Bih.hs:211-324) on the GPU.
//
// The reference builds the tree by recursive list partition; host_builder.cpp does the same with index
// ranges.  Here the recursion is turned into a level-synchronous sweep over the item array:
//
//   per level   k_accum    every item of a pending node evaluates the four candidate partitions (x, y, z mid
//                          split, big/small, Bih.hs:218-232) and folds count / lmax / rmin into its node's
//                          accumulators (warp- and block-aggregated while a block sits inside one node)
//               k_decide   one thread per node: costs, leaf rule and the (sic) selection chain (Bih.hs:252-285),
//                          child boxes; allocates the two children
//               k_scan_*   exclusive scan of the "goes left" flag over the whole array
//               k_scatter  stable partition of every splitting node at once: left rank = S[p] - S[lo]
//   afterwards  k_sizes / k_number / k_emit: subtree sizes bottom-up, pre-order numbers top-down, and the
//               node / leaf arrays in exactly the layout bih_build() produces.
//
// Arithmetic is glome_math.h compiled with -fmad=false, and the min / max folds are order-independent for the
// finite values involved, so the tree is the host builder's tree bit for bit (tests/test_gpu_build.py).
#include <cuda_runtime.h>

#include <cstring>
#include <map>
#include <mutex>
#include <string>
#include <vector>

#include "glome_build.h"

namespace glome_host {
namespace {

using namespace glm;

#define BK(x)                                                                                              \
    do {                                                                                                   \
        cudaError_t e_ = (x);                                                                              \
        if (e_ != cudaSuccess) throw BuildError(std::string("bih_build_gpu: ") + #x + ": " + cudaGetErrorString(e_)); \
    } while (0)

enum { ST_PENDING = 0, ST_BRANCH = 1, ST_LEAF = 2 };
enum { ERR_DEPTH = 1, ERR_NODES = 2 };
#define BUILD_MAX_DEPTH 512 /* host_builder.cpp's guard */

struct BNode {
    int lo, hi;         // item positions [lo, hi)
    int state, k;       // k: winning partition 0..2 = x,y,z, 3 = big/small
    int left, right;    // children (node ids)
    int split, depth;   // first position of the right child
    int acc, pad;       // accumulator slot while pending
    double bb[6];
    double mid[3];
    double thresh;      // 0.4 * bbsa'(bb)   (Bih.hs:223)
    double lmax, rmin;  // of the winning partition
};

struct LevelAcc {       // accumulators of one node of the current level
    int lc[4];
    unsigned long long lmax[4], rmin[4];  // order-preserving keys
};

struct Counters { int n_nodes, n_pending, err, pad; };

// order-preserving map double -> uint64 (for atomicMax / atomicMin)
__device__ __forceinline__ unsigned long long f2key(double x) {
    unsigned long long b = (unsigned long long)__double_as_longlong(x);
    return (b >> 63) ? ~b : (b | 0x8000000000000000ull);
}
__device__ __forceinline__ double key2f(unsigned long long k) {
    unsigned long long b = (k >> 63) ? (k & 0x7fffffffffffffffull) : ~k;
    return __longlong_as_double((long long)b);
}
__device__ __forceinline__ Flt bbsa_p(const Bbox& b) { return hmax(0, bbsa(b)); }  // Bih.hs:208
__device__ __forceinline__ Bbox ldbb6(const double* b) { return mkbb(vec(b[0], b[1], b[2]), vec(b[3], b[4], b[5])); }

__device__ __forceinline__ unsigned long long warp_max(unsigned long long v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) { unsigned long long w = __shfl_xor_sync(0xffffffffu, v, o); v = w > v ? w : v; }
    return v;
}
__device__ __forceinline__ unsigned long long warp_min(unsigned long long v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) { unsigned long long w = __shfl_xor_sync(0xffffffffu, v, o); v = w < v ? w : v; }
    return v;
}

// per item: bbmid and bbsa' of its box; the BIH's box = foldl' bbjoin empty_bbox (Bih.hs:315)
__global__ void k_prep(int n, const double* __restrict__ bb, double* __restrict__ mid, double* __restrict__ sa,
                       int* __restrict__ idx, int* __restrict__ node_of, unsigned long long* bbkeys) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    unsigned long long lo[3], hi[3];
    const unsigned long long KMAX = f2key(GLM_INFINITY), KMIN = f2key(-GLM_INFINITY);
    for (int a = 0; a < 3; a++) { lo[a] = KMAX; hi[a] = KMIN; }
    if (i < n) {
        Bbox ob = ldbb6(bb + 6 * (size_t)i);
        Vec m = bbmid(ob);
        mid[3 * (size_t)i] = m.x; mid[3 * (size_t)i + 1] = m.y; mid[3 * (size_t)i + 2] = m.z;
        sa[i] = bbsa_p(ob);
        idx[i] = i;
        node_of[i] = 0;
        lo[0] = f2key(ob.p1.x); lo[1] = f2key(ob.p1.y); lo[2] = f2key(ob.p1.z);
        hi[0] = f2key(ob.p2.x); hi[1] = f2key(ob.p2.y); hi[2] = f2key(ob.p2.z);
    }
    for (int a = 0; a < 3; a++) { lo[a] = warp_min(lo[a]); hi[a] = warp_max(hi[a]); }
    if ((threadIdx.x & 31) == 0)
        for (int a = 0; a < 3; a++) { atomicMin(bbkeys + a, lo[a]); atomicMax(bbkeys + 3 + a, hi[a]); }
}

__device__ __forceinline__ void node_set_box(BNode& nd, const Bbox& b) {
    nd.bb[0] = b.p1.x; nd.bb[1] = b.p1.y; nd.bb[2] = b.p1.z; nd.bb[3] = b.p2.x; nd.bb[4] = b.p2.y; nd.bb[5] = b.p2.z;
    Vec m = bbmid(b);  // build_rec ... (bbmid childbb)  (Bih.hs:260-274)
    nd.mid[0] = m.x; nd.mid[1] = m.y; nd.mid[2] = m.z;
    nd.thresh = bbsa_p(b) * 0.4;
}

__global__ void k_init_root(int n, const unsigned long long* bbkeys, BNode* nodes, Counters* ctr, double* bb_out) {
    Bbox b = mkbb(vec(key2f(bbkeys[0]), key2f(bbkeys[1]), key2f(bbkeys[2])), vec(key2f(bbkeys[3]), key2f(bbkeys[4]), key2f(bbkeys[5])));
    BNode nd;
    memset(&nd, 0, sizeof(nd));
    nd.lo = 0; nd.hi = n; nd.depth = 0;
    nd.state = (n <= 3) ? ST_LEAF : ST_PENDING;  // Bih.hs:214
    nd.left = nd.right = -1;
    node_set_box(nd, b);
    nodes[0] = nd;
    ctr->n_nodes = 1; ctr->n_pending = (n <= 3) ? 0 : 1; ctr->err = 0;
    bb_out[0] = b.p1.x; bb_out[1] = b.p1.y; bb_out[2] = b.p1.z; bb_out[3] = b.p2.x; bb_out[4] = b.p2.y; bb_out[5] = b.p2.z;
}

__global__ void k_acc_reset(int count, LevelAcc* acc) {
    int a = blockIdx.x * blockDim.x + threadIdx.x;
    if (a >= count) return;
    const unsigned long long KMAX = f2key(GLM_INFINITY), KMIN = f2key(-GLM_INFINITY);
    for (int k = 0; k < 4; k++) { acc[a].lc[k] = 0; acc[a].lmax[k] = KMIN; acc[a].rmin[k] = KMAX; }  // seeds: Bih.hs:225-232
}

// the four candidate partitions of Bih.hs:218-232, evaluated per item
__global__ void __launch_bounds__(256) k_accum(int n, int lvl_start, const BNode* __restrict__ nodes, const int* __restrict__ idx,
                                               const int* __restrict__ node_of, const double* __restrict__ bb,
                                               const double* __restrict__ mid, const double* __restrict__ sa, LevelAcc* acc) {
    __shared__ int s_lc[4];
    __shared__ unsigned long long s_lmax[4], s_rmin[4];
    const int p = blockIdx.x * blockDim.x + threadIdx.x;
    const unsigned long long KMAX = f2key(GLM_INFINITY), KMIN = f2key(-GLM_INFINITY);
    int v = -1;
    bool left[4] = {false, false, false, false};
    unsigned long long kmaxv[4] = {KMIN, KMIN, KMIN, KMIN}, kminv[4] = {KMAX, KMAX, KMAX, KMAX};
    if (p < n) {
        int vv = node_of[p];
        if (vv >= lvl_start && nodes[vv].state == ST_PENDING) {
            v = vv;
            const int i = idx[p];
            const double* b = bb + 6 * (size_t)i;
            const double mx = mid[3 * (size_t)i], my = mid[3 * (size_t)i + 1], mz = mid[3 * (size_t)i + 2];
            const BNode& nd = nodes[v];
            left[0] = mx < nd.mid[0]; left[1] = my < nd.mid[1]; left[2] = mz < nd.mid[2];
            left[3] = sa[i] > nd.thresh;
            // lmax = max p2.k over the left, rmin = min p1.k over the right; big/small uses the x planes (Bih.hs:231-232)
            if (left[0]) kmaxv[0] = f2key(b[3]); else kminv[0] = f2key(b[0]);
            if (left[1]) kmaxv[1] = f2key(b[4]); else kminv[1] = f2key(b[1]);
            if (left[2]) kmaxv[2] = f2key(b[5]); else kminv[2] = f2key(b[2]);
            if (left[3]) kmaxv[3] = f2key(b[3]); else kminv[3] = f2key(b[0]);
        }
    }
    const unsigned int FULL = 0xffffffffu;
    const int v0 = __shfl_sync(FULL, v, 0);
    const bool warp_uniform = __all_sync(FULL, v == v0);
    // block-uniform: every thread of the block is in the same pending node
    if (threadIdx.x < 4) { s_lc[threadIdx.x] = 0; s_lmax[threadIdx.x] = KMIN; s_rmin[threadIdx.x] = KMAX; }
    __shared__ int s_v0;
    if (threadIdx.x == 0) s_v0 = v;
    __syncthreads();
    const bool block_uniform = __syncthreads_and(v == s_v0 && v >= 0) != 0;
    if (warp_uniform) {
        if (v0 < 0) return;  // (block_uniform is false then, no barrier follows)
        int cnt[4];
        unsigned long long mxk[4], mnk[4];
        for (int k = 0; k < 4; k++) {
            cnt[k] = __popc(__ballot_sync(FULL, left[k]));
            mxk[k] = warp_max(kmaxv[k]);
            mnk[k] = warp_min(kminv[k]);
        }
        if ((threadIdx.x & 31) == 0) {
            if (block_uniform) {
                for (int k = 0; k < 4; k++) { atomicAdd(&s_lc[k], cnt[k]); atomicMax(&s_lmax[k], mxk[k]); atomicMin(&s_rmin[k], mnk[k]); }
            } else {
                LevelAcc* A = acc + nodes[v0].acc;
                for (int k = 0; k < 4; k++) {
                    if (cnt[k]) atomicAdd(&A->lc[k], cnt[k]);
                    if (mxk[k] != KMIN) atomicMax(&A->lmax[k], mxk[k]);
                    if (mnk[k] != KMAX) atomicMin(&A->rmin[k], mnk[k]);
                }
            }
        }
    } else if (v >= 0) {
        LevelAcc* A = acc + nodes[v].acc;
        for (int k = 0; k < 4; k++) {
            if (left[k]) { atomicAdd(&A->lc[k], 1); atomicMax(&A->lmax[k], kmaxv[k]); }
            else atomicMin(&A->rmin[k], kminv[k]);
        }
    }
    if (block_uniform) {
        __syncthreads();
        if (threadIdx.x < 4) {
            LevelAcc* A = acc + nodes[s_v0].acc;
            const int k = threadIdx.x;
            if (s_lc[k]) atomicAdd(&A->lc[k], s_lc[k]);
            if (s_lmax[k] != KMIN) atomicMax(&A->lmax[k], s_lmax[k]);
            if (s_rmin[k] != KMAX) atomicMin(&A->rmin[k], s_rmin[k]);
        }
    }
}

__device__ __forceinline__ Bbox set_p2(Bbox b, int ax, Flt f) { if (ax == 0) b.p2.x = f; else if (ax == 1) b.p2.y = f; else b.p2.z = f; return b; }
__device__ __forceinline__ Bbox set_p1(Bbox b, int ax, Flt f) { if (ax == 0) b.p1.x = f; else if (ax == 1) b.p1.y = f; else b.p1.z = f; return b; }

// build_rec's decision for every pending node of the level (Bih.hs:243-285)
__global__ void k_decide(int lvl_start, int lvl_end, BNode* nodes, const LevelAcc* __restrict__ acc, Counters* ctr, int max_nodes) {
    int v = lvl_start + blockIdx.x * blockDim.x + threadIdx.x;
    if (v >= lvl_end) return;
    BNode nd = nodes[v];
    if (nd.state != ST_PENDING) return;
    if (nd.depth > BUILD_MAX_DEPTH) { atomicOr(&ctr->err, ERR_DEPTH); nodes[v].state = ST_LEAF; return; }
    const LevelAcc A = acc[nd.acc];
    const int n = nd.hi - nd.lo;
    const Bbox bb = ldbb6(nd.bb);
    const Flt sa = bbsa_p(bb);
    Flt lmax[4], rmin[4], cost[4];
    Bbox lbb[4], rbb[4];
    for (int k = 0; k < 4; k++) {
        lmax[k] = key2f(A.lmax[k]); rmin[k] = key2f(A.rmin[k]);
        const int ax = (k == 3) ? 0 : k;
        lbb[k] = set_p2(bb, ax, lmax[k]);
        rbb[k] = set_p1(bb, ax, rmin[k]);
        const Flt fac = (k == 3) ? 1.2 : 1.1;  // Bih.hs:252-255
        cost[k] = ((bbsa_p(lbb[k]) * (Flt)A.lc[k]) + (bbsa_p(rbb[k]) * (Flt)(n - A.lc[k]))) * fac;
    }
    const Flt costorig = sa * (Flt)n;
    if (costorig < cost[0] && costorig < cost[1] && costorig < cost[2] && costorig < cost[3]) {  // Bih.hs:276
        nodes[v].state = ST_LEAF;
        return;
    }
    int k;
    if (cost[0] < cost[1] && cost[0] < cost[2] && cost[0] < cost[3]) k = 0;
    else if (cost[1] < cost[2] && cost[1] < cost[3]) k = 1;
    else if (cost[1] < cost[3]) k = 2;  // sic (Bih.hs:283 tests costy)
    else k = 3;
    const int base = atomicAdd(&ctr->n_nodes, 2);
    if (base + 2 > max_nodes) { atomicOr(&ctr->err, ERR_NODES); nodes[v].state = ST_LEAF; return; }
    const int split = nd.lo + A.lc[k];
    BNode l, r;
    memset(&l, 0, sizeof(l)); memset(&r, 0, sizeof(r));
    l.lo = nd.lo; l.hi = split; r.lo = split; r.hi = nd.hi;
    l.depth = r.depth = nd.depth + 1;
    l.left = l.right = r.left = r.right = -1;
    l.state = (l.hi - l.lo <= 3) ? ST_LEAF : ST_PENDING;
    r.state = (r.hi - r.lo <= 3) ? ST_LEAF : ST_PENDING;
    node_set_box(l, lbb[k]);
    node_set_box(r, rbb[k]);
    const int np = (l.state == ST_PENDING) + (r.state == ST_PENDING);
    if (np) {  // the pending children's accumulator slots for the next level
        const int slot = atomicAdd(&ctr->n_pending, np);
        if (l.state == ST_PENDING) { l.acc = slot; r.acc = slot + 1; } else r.acc = slot;
    }
    nodes[base] = l;
    nodes[base + 1] = r;
    nodes[v].state = ST_BRANCH; nodes[v].k = k; nodes[v].left = base; nodes[v].right = base + 1; nodes[v].split = split;
    nodes[v].lmax = lmax[k]; nodes[v].rmin = rmin[k];
}

// "goes left" under its node's winning partition; 0 for items whose node did not split at this level
template <typename NodeT>
__device__ __forceinline__ int goes_left(int p, int lvl_start, int lvl_end, const NodeT* __restrict__ nodes, const int* __restrict__ idx,
                                         const int* __restrict__ node_of, const double* __restrict__ mid,
                                         const double* __restrict__ sa, int& v_out) {
    const int v = node_of[p];
    v_out = v;
    if (v < lvl_start || v >= lvl_end) return -1;
    const NodeT& nd = nodes[v];
    if (nd.state != ST_BRANCH) return -1;
    const int i = idx[p];
    const int k = nd.k;
    if (k == 3) return sa[i] > nd.thresh ? 1 : 0;
    return mid[3 * (size_t)i + k] < nd.mid[k] ? 1 : 0;
}

#define SCAN_ITEMS 4
#define SCAN_THREADS 256
#define SCAN_TILE (SCAN_ITEMS * SCAN_THREADS)
// exclusive scan of the flags inside each 1024-item tile + the tile totals
template <typename NodeT>
__global__ void __launch_bounds__(SCAN_THREADS) k_scan_local(int n, int lvl_start, int lvl_end, const NodeT* __restrict__ nodes,
                                                              const int* __restrict__ idx, const int* __restrict__ node_of,
                                                              const double* __restrict__ mid, const double* __restrict__ sa,
                                                              int* __restrict__ S, int* __restrict__ tile_sum) {
    __shared__ int wsum[SCAN_THREADS / 32];
    const int base = blockIdx.x * SCAN_TILE + threadIdx.x * SCAN_ITEMS;
    int f[SCAN_ITEMS], tot = 0;
    for (int j = 0; j < SCAN_ITEMS; j++) {
        const int p = base + j;
        int v;
        f[j] = (p < n && goes_left(p, lvl_start, lvl_end, nodes, idx, node_of, mid, sa, v) == 1) ? 1 : 0;
        tot += f[j];
    }
    int incl = tot;
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    for (int o = 1; o < 32; o <<= 1) { int t = __shfl_up_sync(0xffffffffu, incl, o); if (lane >= o) incl += t; }
    if (lane == 31) wsum[w] = incl;
    __syncthreads();
    int woff = 0, all = 0;
    for (int j = 0; j < SCAN_THREADS / 32; j++) { if (j < w) woff += wsum[j]; all += wsum[j]; }
    int run = woff + incl - tot;
    for (int j = 0; j < SCAN_ITEMS; j++) {
        const int p = base + j;
        if (p < n) S[p] = run;
        run += f[j];
    }
    if (threadIdx.x == 0) tile_sum[blockIdx.x] = all;
}
// exclusive scan of the tile totals (one block)
__global__ void __launch_bounds__(1024) k_scan_tiles(int ntiles, int* tile_sum) {
    __shared__ int wsum[32];
    __shared__ int carry;
    if (threadIdx.x == 0) carry = 0;
    __syncthreads();
    for (int base = 0; base < ntiles; base += 1024) {
        const int t = base + threadIdx.x;
        const int x = (t < ntiles) ? tile_sum[t] : 0;
        int incl = x;
        const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
        for (int o = 1; o < 32; o <<= 1) { int y = __shfl_up_sync(0xffffffffu, incl, o); if (lane >= o) incl += y; }
        if (lane == 31) wsum[w] = incl;
        __syncthreads();
        int woff = 0, all = 0;
        for (int j = 0; j < 32; j++) { if (j < w) woff += wsum[j]; all += wsum[j]; }
        const int c = carry;
        if (t < ntiles) tile_sum[t] = c + woff + incl - x;
        __syncthreads();
        if (threadIdx.x == 0) carry = c + all;
        __syncthreads();
    }
}
// stable partition of every node that split at this level
template <typename NodeT>
__global__ void __launch_bounds__(256) k_scatter(int n, int lvl_start, int lvl_end, const NodeT* __restrict__ nodes,
                                                 const int* __restrict__ idx, const int* __restrict__ node_of,
                                                 const double* __restrict__ mid, const double* __restrict__ sa,
                                                 const int* __restrict__ S, const int* __restrict__ tile_off,
                                                 int* __restrict__ idx2, int* __restrict__ node_of2) {
    const int p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= n) return;
    int v;
    const int gl = goes_left(p, lvl_start, lvl_end, nodes, idx, node_of, mid, sa, v);
    if (gl < 0) { idx2[p] = idx[p]; node_of2[p] = v; return; }
    const NodeT& nd = nodes[v];
    const int lo = nd.lo;
    const int before = (S[p] + tile_off[p / SCAN_TILE]) - (S[lo] + tile_off[lo / SCAN_TILE]);  // lefts in [lo, p)
    int q, child;
    if (gl) { q = lo + before; child = nd.left; }
    else { q = nd.split + ((p - lo) - before); child = nd.right; }
    idx2[q] = idx[p];
    node_of2[q] = child;
}

// ---- numbering: bih_build() emits nodes in pre-order and leaves in traversal order ----
template <typename NodeT>
__global__ void k_sizes(int lvl_start, int lvl_end, const NodeT* __restrict__ nodes, int* nb, int* nl) {
    int v = lvl_start + blockIdx.x * blockDim.x + threadIdx.x;
    if (v >= lvl_end) return;
    if (nodes[v].state == ST_BRANCH) {
        nb[v] = 1 + nb[nodes[v].left] + nb[nodes[v].right];
        nl[v] = nl[nodes[v].left] + nl[nodes[v].right];
    } else { nb[v] = 0; nl[v] = 1; }
}
template <typename NodeT>
__global__ void k_number(int lvl_start, int lvl_end, const NodeT* __restrict__ nodes, const int* __restrict__ nb,
                         const int* __restrict__ nl, int* pre, int* lbase) {
    int v = lvl_start + blockIdx.x * blockDim.x + threadIdx.x;
    if (v >= lvl_end) return;
    if (nodes[v].state != ST_BRANCH) return;
    const int l = nodes[v].left, r = nodes[v].right;
    pre[l] = pre[v] + 1; lbase[l] = lbase[v];
    pre[r] = pre[v] + 1 + nb[l]; lbase[r] = lbase[v] + nl[l];
}
__global__ void k_emit(int n_nodes, const BNode* __restrict__ nodes, const int* __restrict__ pre, const int* __restrict__ lbase,
                       GlomeBihNode* out_nodes, int* out_leaves) {
    int v = blockIdx.x * blockDim.x + threadIdx.x;
    if (v >= n_nodes) return;
    const BNode& nd = nodes[v];
    if (nd.state == ST_BRANCH) {
        GlomeBihNode o;
        o.lsplit = nd.lmax + GLM_DELTA;  // Bih.hs:257-258
        o.rsplit = nd.rmin - GLM_DELTA;
        o.axis = (nd.k == 3) ? 0 : nd.k;
        const int l = nd.left, r = nd.right;
        o.left = (nodes[l].state == ST_BRANCH) ? pre[l] : ~lbase[l];
        o.right = (nodes[r].state == ST_BRANCH) ? pre[r] : ~lbase[r];
        o.pad = 0;
        out_nodes[pre[v]] = o;
    } else {
        out_leaves[2 * lbase[v]] = nd.lo;
        out_leaves[2 * lbase[v] + 1] = nd.hi - nd.lo;
    }
}


// =====================================================================================================
// mesh BVH (`mesh` / `build_tree`, Mesh.hs:50-134): same sweep; a node keeps the joined triangle boxes of
// both halves of each candidate partition (trisbb, Mesh.hs:124-125), all costs x 1.1, leaf below 3 triangles.
// =====================================================================================================
struct MNodeD {
    int lo, hi;
    int state, k;
    int left, right;
    int split, depth;
    int acc, pad;
    double bb[6];
    double mid[3];
    double thresh;          // 0.4 * bbsa bb  (Mesh.hs:87)
    double lbb[6], rbb[6];  // of the winning partition
};
struct MeshAcc {
    int lc[4];
    unsigned long long lb[4][6], rb[4][6];  // box keys: [0..2] p1 (min-folded), [3..5] p2 (max-folded)
};

__device__ __forceinline__ void mnode_set_box(MNodeD& nd, const Bbox& b) {
    nd.bb[0] = b.p1.x; nd.bb[1] = b.p1.y; nd.bb[2] = b.p1.z; nd.bb[3] = b.p2.x; nd.bb[4] = b.p2.y; nd.bb[5] = b.p2.z;
    Vec m = bbmid(b);  // Mesh.hs:84
    nd.mid[0] = m.x; nd.mid[1] = m.y; nd.mid[2] = m.z;
    nd.thresh = bbsa(b) * 0.4;
}

// bbpts of all vertices (Mesh.hs:55; Vec.hs:676-690): every point inflated by delta
__global__ void k_mesh_vbox(int nverts, const double* __restrict__ verts, unsigned long long* bbkeys) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    unsigned long long lo[3], hi[3];
    const unsigned long long KMAX = f2key(GLM_INFINITY), KMIN = f2key(-GLM_INFINITY);
    for (int a = 0; a < 3; a++) { lo[a] = KMAX; hi[a] = KMIN; }
    if (i < nverts)
        for (int a = 0; a < 3; a++) { const double x = verts[3 * (size_t)i + a]; lo[a] = f2key(x - GLM_DELTA); hi[a] = f2key(x + GLM_DELTA); }
    for (int a = 0; a < 3; a++) { lo[a] = warp_min(lo[a]); hi[a] = warp_max(hi[a]); }
    if ((threadIdx.x & 31) == 0)
        for (int a = 0; a < 3; a++) { atomicMin(bbkeys + a, lo[a]); atomicMax(bbkeys + 3 + a, hi[a]); }
}
// alltribbs (Mesh.hs:119-121): bbpts [a, b, c]
__global__ void k_mesh_prep(int ntris, const double* __restrict__ verts, const int* __restrict__ tris, double* __restrict__ tbb,
                            double* __restrict__ mid, double* __restrict__ sa, int* __restrict__ idx, int* __restrict__ node_of) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= ntris) return;
    const int* T = tris + 8 * (size_t)i;
    double p1[3], p2[3];
    for (int a = 0; a < 3; a++) {
        const double xa = verts[3 * (size_t)T[0] + a], xb = verts[3 * (size_t)T[1] + a], xc = verts[3 * (size_t)T[2] + a];
        // fold from the last point (host_builder.cpp mesh_build): r = box(c); r = join(box(b), r); r = join(box(a), r)
        double lo = xc - GLM_DELTA, hi = xc + GLM_DELTA;
        lo = fmin_(xb - GLM_DELTA, lo); hi = fmax_(xb + GLM_DELTA, hi);
        lo = fmin_(xa - GLM_DELTA, lo); hi = fmax_(xa + GLM_DELTA, hi);
        p1[a] = lo; p2[a] = hi;
    }
    Bbox r = mkbb(vec(p1[0], p1[1], p1[2]), vec(p2[0], p2[1], p2[2]));
    double* o = tbb + 6 * (size_t)i;
    o[0] = p1[0]; o[1] = p1[1]; o[2] = p1[2]; o[3] = p2[0]; o[4] = p2[1]; o[5] = p2[2];
    Vec m = bbmid(r);
    mid[3 * (size_t)i] = m.x; mid[3 * (size_t)i + 1] = m.y; mid[3 * (size_t)i + 2] = m.z;
    sa[i] = bbsa(r);
    idx[i] = i;
    node_of[i] = 0;
}
__global__ void k_mesh_init_root(int n, const unsigned long long* bbkeys, MNodeD* nodes, Counters* ctr, double* bb_out) {
    Bbox b = mkbb(vec(key2f(bbkeys[0]), key2f(bbkeys[1]), key2f(bbkeys[2])), vec(key2f(bbkeys[3]), key2f(bbkeys[4]), key2f(bbkeys[5])));
    MNodeD nd;
    memset(&nd, 0, sizeof(nd));
    nd.lo = 0; nd.hi = n;
    nd.state = (n < 3) ? ST_LEAF : ST_PENDING;  // Mesh.hs:72
    nd.left = nd.right = -1;
    mnode_set_box(nd, b);
    nodes[0] = nd;
    ctr->n_nodes = 1; ctr->n_pending = (n < 3) ? 0 : 1; ctr->err = 0;
    bb_out[0] = b.p1.x; bb_out[1] = b.p1.y; bb_out[2] = b.p1.z; bb_out[3] = b.p2.x; bb_out[4] = b.p2.y; bb_out[5] = b.p2.z;
}
__global__ void k_mesh_acc_reset(int count, MeshAcc* acc) {
    int a = blockIdx.x * blockDim.x + threadIdx.x;
    if (a >= count) return;
    const unsigned long long KMAX = f2key(GLM_INFINITY), KMIN = f2key(-GLM_INFINITY);
    for (int k = 0; k < 4; k++) {
        acc[a].lc[k] = 0;
        for (int j = 0; j < 3; j++) { acc[a].lb[k][j] = KMAX; acc[a].rb[k][j] = KMAX; acc[a].lb[k][3 + j] = KMIN; acc[a].rb[k][3 + j] = KMIN; }  // empty_bbox
    }
}
// the four candidate partitions of Mesh.hs:84-96: count and joined boxes of both halves
__global__ void __launch_bounds__(256) k_mesh_accum(int n, int lvl_start, const MNodeD* __restrict__ nodes, const int* __restrict__ idx,
                                                    const int* __restrict__ node_of, const double* __restrict__ tbb,
                                                    const double* __restrict__ mid, const double* __restrict__ sa, MeshAcc* acc) {
    __shared__ int s_lc[4];
    __shared__ unsigned long long s_lb[4][6], s_rb[4][6];
    __shared__ int s_v0;
    const int p = blockIdx.x * blockDim.x + threadIdx.x;
    const unsigned long long KMAX = f2key(GLM_INFINITY), KMIN = f2key(-GLM_INFINITY);
    int v = -1;
    bool left[4] = {false, false, false, false};
    unsigned long long key[6] = {KMAX, KMAX, KMAX, KMIN, KMIN, KMIN};
    if (p < n) {
        int vv = node_of[p];
        if (vv >= lvl_start && nodes[vv].state == ST_PENDING) {
            v = vv;
            const int i = idx[p];
            const double* b = tbb + 6 * (size_t)i;
            const MNodeD& nd = nodes[v];
            left[0] = mid[3 * (size_t)i] < nd.mid[0]; left[1] = mid[3 * (size_t)i + 1] < nd.mid[1]; left[2] = mid[3 * (size_t)i + 2] < nd.mid[2];
            left[3] = sa[i] > nd.thresh;
            for (int j = 0; j < 6; j++) key[j] = f2key(b[j]);
        }
    }
    const unsigned int FULL = 0xffffffffu;
    const int v0 = __shfl_sync(FULL, v, 0);
    const bool warp_uniform = __all_sync(FULL, v == v0);
    if (threadIdx.x < 4) {
        s_lc[threadIdx.x] = 0;
        for (int j = 0; j < 3; j++) { s_lb[threadIdx.x][j] = KMAX; s_rb[threadIdx.x][j] = KMAX; s_lb[threadIdx.x][3 + j] = KMIN; s_rb[threadIdx.x][3 + j] = KMIN; }
    }
    if (threadIdx.x == 0) s_v0 = v;
    __syncthreads();
    const bool block_uniform = __syncthreads_and(v == s_v0 && v >= 0) != 0;
    if (warp_uniform) {
        if (v0 < 0) return;
        MeshAcc* A = acc + nodes[v0].acc;
        const bool lead = (threadIdx.x & 31) == 0;
        for (int k = 0; k < 4; k++) {
            const int cnt = __popc(__ballot_sync(FULL, left[k]));
            if (lead) { if (block_uniform) atomicAdd(&s_lc[k], cnt); else if (cnt) atomicAdd(&A->lc[k], cnt); }
            for (int j = 0; j < 6; j++) {
                const unsigned long long seed = (j < 3) ? KMAX : KMIN;
                unsigned long long lv = left[k] ? key[j] : seed, rv = left[k] ? seed : key[j];
                if (j < 3) { lv = warp_min(lv); rv = warp_min(rv); } else { lv = warp_max(lv); rv = warp_max(rv); }
                if (lead) {
                    if (block_uniform) {
                        if (j < 3) { atomicMin(&s_lb[k][j], lv); atomicMin(&s_rb[k][j], rv); }
                        else { atomicMax(&s_lb[k][j], lv); atomicMax(&s_rb[k][j], rv); }
                    } else {
                        if (j < 3) { if (lv != seed) atomicMin(&A->lb[k][j], lv); if (rv != seed) atomicMin(&A->rb[k][j], rv); }
                        else { if (lv != seed) atomicMax(&A->lb[k][j], lv); if (rv != seed) atomicMax(&A->rb[k][j], rv); }
                    }
                }
            }
        }
    } else if (v >= 0) {
        MeshAcc* A = acc + nodes[v].acc;
        for (int k = 0; k < 4; k++) {
            if (left[k]) atomicAdd(&A->lc[k], 1);
            unsigned long long* dst = left[k] ? A->lb[k] : A->rb[k];
            for (int j = 0; j < 3; j++) { atomicMin(dst + j, key[j]); atomicMax(dst + 3 + j, key[3 + j]); }
        }
    }
    if (block_uniform) {
        __syncthreads();
        if (threadIdx.x < 4) {
            MeshAcc* A = acc + nodes[s_v0].acc;
            const int k = threadIdx.x;
            if (s_lc[k]) atomicAdd(&A->lc[k], s_lc[k]);
            for (int j = 0; j < 3; j++) {
                if (s_lb[k][j] != KMAX) atomicMin(&A->lb[k][j], s_lb[k][j]);
                if (s_rb[k][j] != KMAX) atomicMin(&A->rb[k][j], s_rb[k][j]);
                if (s_lb[k][3 + j] != KMIN) atomicMax(&A->lb[k][3 + j], s_lb[k][3 + j]);
                if (s_rb[k][3 + j] != KMIN) atomicMax(&A->rb[k][3 + j], s_rb[k][3 + j]);
            }
        }
    }
}
__device__ __forceinline__ Bbox keys2box(const unsigned long long* k) {
    return mkbb(vec(key2f(k[0]), key2f(k[1]), key2f(k[2])), vec(key2f(k[3]), key2f(k[4]), key2f(k[5])));
}
// build_tree's decision (Mesh.hs:98-113)
__global__ void k_mesh_decide(int lvl_start, int lvl_end, MNodeD* nodes, const MeshAcc* __restrict__ acc, Counters* ctr, int max_nodes) {
    int v = lvl_start + blockIdx.x * blockDim.x + threadIdx.x;
    if (v >= lvl_end) return;
    if (nodes[v].state != ST_PENDING) return;
    const int lo = nodes[v].lo, hi = nodes[v].hi, depth = nodes[v].depth;
    if (depth > BUILD_MAX_DEPTH) { atomicOr(&ctr->err, ERR_DEPTH); nodes[v].state = ST_LEAF; return; }
    const MeshAcc& A = acc[nodes[v].acc];
    const int n = hi - lo;
    const Bbox bb = ldbb6(nodes[v].bb);
    Flt cost[4];
    for (int k = 0; k < 4; k++)
        cost[k] = ((bbsa(keys2box(A.lb[k])) * (Flt)A.lc[k]) + (bbsa(keys2box(A.rb[k])) * (Flt)(n - A.lc[k]))) * 1.1;  // Mesh.hs:98-101
    const Flt lcost = bbsa(bb) * (Flt)n;
    if (lcost < cost[0] && lcost < cost[1] && lcost < cost[2] && lcost < cost[3]) { nodes[v].state = ST_LEAF; return; }  // Mesh.hs:104
    int k;
    if (cost[0] < cost[1] && cost[0] < cost[2] && cost[0] < cost[3]) k = 0;
    else if (cost[1] < cost[2] && cost[1] < cost[3]) k = 1;
    else if (cost[2] < cost[3]) k = 2;
    else k = 3;
    const int base = atomicAdd(&ctr->n_nodes, 2);
    if (base + 2 > max_nodes) { atomicOr(&ctr->err, ERR_NODES); nodes[v].state = ST_LEAF; return; }
    const int split = lo + A.lc[k];
    const Bbox L = keys2box(A.lb[k]), R = keys2box(A.rb[k]);
    MNodeD l, r;
    memset(&l, 0, sizeof(l)); memset(&r, 0, sizeof(r));
    l.lo = lo; l.hi = split; r.lo = split; r.hi = hi;
    l.depth = r.depth = depth + 1;
    l.left = l.right = r.left = r.right = -1;
    l.state = (l.hi - l.lo < 3) ? ST_LEAF : ST_PENDING;
    r.state = (r.hi - r.lo < 3) ? ST_LEAF : ST_PENDING;
    mnode_set_box(l, L);
    mnode_set_box(r, R);
    const int np = (l.state == ST_PENDING) + (r.state == ST_PENDING);
    if (np) {
        const int slot = atomicAdd(&ctr->n_pending, np);
        if (l.state == ST_PENDING) { l.acc = slot; r.acc = slot + 1; } else r.acc = slot;
    }
    nodes[base] = l;
    nodes[base + 1] = r;
    MNodeD& me = nodes[v];
    me.state = ST_BRANCH; me.k = k; me.left = base; me.right = base + 1; me.split = split;
    me.lbb[0] = L.p1.x; me.lbb[1] = L.p1.y; me.lbb[2] = L.p1.z; me.lbb[3] = L.p2.x; me.lbb[4] = L.p2.y; me.lbb[5] = L.p2.z;
    me.rbb[0] = R.p1.x; me.rbb[1] = R.p1.y; me.rbb[2] = R.p1.z; me.rbb[3] = R.p2.x; me.rbb[4] = R.p2.y; me.rbb[5] = R.p2.z;
}
// leaf record li lives at leafpool[li + lo]: {count, tri ...}, because the leaves tile the permutation in order
__global__ void k_mesh_emit(int n_nodes, const MNodeD* __restrict__ nodes, const int* __restrict__ pre, const int* __restrict__ lbase,
                            const int* __restrict__ idx, GlomeBvhNode* out_nodes, int* leafpool, int* leafoff) {
    int v = blockIdx.x * blockDim.x + threadIdx.x;
    if (v >= n_nodes) return;
    const MNodeD& nd = nodes[v];
    if (nd.state == ST_BRANCH) {
        GlomeBvhNode o;
        memset(&o, 0, sizeof(o));
        for (int j = 0; j < 6; j++) { o.lbb[j] = nd.lbb[j]; o.rbb[j] = nd.rbb[j]; }
        const int l = nd.left, r = nd.right;
        o.left = (nodes[l].state == ST_BRANCH) ? pre[l] : ~lbase[l];
        o.right = (nodes[r].state == ST_BRANCH) ? pre[r] : ~lbase[r];
        out_nodes[pre[v]] = o;
    } else {
        const int li = lbase[v], off = li + nd.lo, cnt = nd.hi - nd.lo;
        leafoff[li] = off;
        leafpool[off] = cnt;
        for (int j = 0; j < cnt; j++) leafpool[off + 1 + j] = idx[nd.lo + j];
    }
}

// Work space of a build.  cudaMalloc / cudaFree of the ~0.5 GB a 10^6-item build touches cost several times the build
// itself (BENCH_r01: 6 ms on the device, 67 ms wall), so freed blocks go to a per-device cache (power-of-two size
// classes) and the next build of a similar size allocates nothing.  glome_build_release_cache() gives the memory back.
struct BlockCache {
    std::mutex mu;
    std::multimap<size_t, void*> free_blocks[64];  // per device: size class -> block
    size_t cached_bytes[64] = {0};
};
static BlockCache& block_cache() { static BlockCache c; return c; }
static const size_t BUILD_CACHE_LIMIT = (size_t)8 << 30;  // per device

struct DevMem {
    int device;
    std::vector<std::pair<size_t, void*>> blocks;
    explicit DevMem(int dev) : device(dev & 63) {}
    ~DevMem() {
        BlockCache& C = block_cache();
        std::lock_guard<std::mutex> g(C.mu);
        for (auto& b : blocks) {
            if (C.cached_bytes[device] + b.first > BUILD_CACHE_LIMIT) { cudaFree(b.second); continue; }
            C.free_blocks[device].insert(b);
            C.cached_bytes[device] += b.first;
        }
    }
    template <typename T> T* alloc(size_t count) {
        size_t want = (count ? count : 1) * sizeof(T), cls = 256;
        while (cls < want) cls <<= 1;
        {
            BlockCache& C = block_cache();
            std::lock_guard<std::mutex> g(C.mu);
            auto it = C.free_blocks[device].find(cls);
            if (it != C.free_blocks[device].end()) {
                void* p = it->second;
                C.free_blocks[device].erase(it);
                C.cached_bytes[device] -= cls;
                blocks.push_back({cls, p});
                return (T*)p;
            }
        }
        void* p = nullptr;
        cudaError_t e = cudaMalloc(&p, cls);
        if (e != cudaSuccess) {  // out of memory with blocks parked in the cache: give them back and try once more
            cudaGetLastError();
            release_cache(device);
            e = cudaMalloc(&p, cls);
        }
        if (e != cudaSuccess) throw BuildError(std::string("bih_build_gpu: cudaMalloc: ") + cudaGetErrorString(e));
        blocks.push_back({cls, p});
        return (T*)p;
    }
    static void release_cache(int dev) {
        BlockCache& C = block_cache();
        std::lock_guard<std::mutex> g(C.mu);
        for (auto& b : C.free_blocks[dev & 63]) cudaFree(b.second);
        C.free_blocks[dev & 63].clear();
        C.cached_bytes[dev & 63] = 0;
    }
};

static inline int cdiv(long long a, int b) { return (int)((a + b - 1) / b); }

}  // namespace

static void bih_build_gpu_impl(int64_t n64, const double* bboxes, int device, BihTree& out, double* timings_ms) {
    if (n64 < 0 || n64 > 0x3fffffff) throw BuildError("bih_build_gpu: item count out of range");
    const int n = (int)n64;
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || device < 0 || device >= ndev) throw BuildError("bih_build_gpu: no such CUDA device");
    BK(cudaSetDevice(device));
    out.nodes.clear(); out.leaves.clear(); out.order.clear();
    if (n == 0) {  // bih [] : an empty leaf over the empty box (bih_build does the same)
        out.bb = glm::empty_bbox();
        out.leaves.push_back(0); out.leaves.push_back(0);
        out.root = ~0;
        if (timings_ms) timings_ms[0] = timings_ms[1] = timings_ms[2] = 0;
        return;
    }
    DevMem M(device);
    const int max_nodes = 3 * n + 4096;
    double* d_bb = M.alloc<double>(6 * (size_t)n);
    double* d_mid = M.alloc<double>(3 * (size_t)n);
    double* d_sa = M.alloc<double>(n);
    int* d_idx[2] = {M.alloc<int>(n), M.alloc<int>(n)};
    int* d_nof[2] = {M.alloc<int>(n), M.alloc<int>(n)};
    int* d_S = M.alloc<int>(n);
    const int ntiles = cdiv(n, SCAN_TILE);
    int* d_tiles = M.alloc<int>(ntiles);
    BNode* d_nodes = M.alloc<BNode>(max_nodes);
    LevelAcc* d_acc = M.alloc<LevelAcc>((size_t)n / 4 + 8);   // a pending node holds more than 3 items
    Counters* d_ctr = M.alloc<Counters>(1);
    unsigned long long* d_bbkeys = M.alloc<unsigned long long>(6);
    double* d_bbout = M.alloc<double>(6);
    cudaEvent_t ev[4];
    for (auto& e : ev) BK(cudaEventCreate(&e));
    struct EvGuard { cudaEvent_t* e; ~EvGuard() { for (int i = 0; i < 4; i++) cudaEventDestroy(e[i]); } } evg{ev};

    BK(cudaEventRecord(ev[0]));
    BK(cudaMemcpy(d_bb, bboxes, sizeof(double) * 6 * (size_t)n, cudaMemcpyHostToDevice));
    BK(cudaEventRecord(ev[1]));
    {   // bbox key seeds = empty_bbox (Vec.hs:706)
        unsigned long long seeds[6];
        // f2key on the host: same bit trick
        auto hk = [](double x) { unsigned long long b; memcpy(&b, &x, 8); return (b >> 63) ? ~b : (b | 0x8000000000000000ull); };
        for (int a = 0; a < 3; a++) { seeds[a] = hk(GLM_INFINITY); seeds[3 + a] = hk(-GLM_INFINITY); }
        BK(cudaMemcpy(d_bbkeys, seeds, sizeof(seeds), cudaMemcpyHostToDevice));
    }
    k_prep<<<cdiv(n, 256), 256>>>(n, d_bb, d_mid, d_sa, d_idx[0], d_nof[0], d_bbkeys);
    k_init_root<<<1, 1>>>(n, d_bbkeys, d_nodes, d_ctr, d_bbout);
    BK(cudaGetLastError());

    std::vector<std::pair<int, int>> levels;  // node id ranges
    int lvl_start = 0, lvl_end = 1, cur = 0;
    Counters hc;
    BK(cudaMemcpy(&hc, d_ctr, sizeof(hc), cudaMemcpyDeviceToHost));
    double hbb[6];
    BK(cudaMemcpy(hbb, d_bbout, sizeof(hbb), cudaMemcpyDeviceToHost));
    for (int a = 0; a < 3; a++)
        if (hbb[a] == -GLM_INFINITY || hbb[3 + a] == GLM_INFINITY) throw BuildError("bih: infinite bounding box");  // Bih.hs:319-322
    levels.push_back({0, 1});
    while (hc.n_pending > 0) {
        const int cnt = lvl_end - lvl_start;
        BK(cudaMemsetAsync(&d_ctr->n_pending, 0, sizeof(int)));
        k_acc_reset<<<cdiv(hc.n_pending, 256), 256>>>(hc.n_pending, d_acc);
        k_accum<<<cdiv(n, 256), 256>>>(n, lvl_start, d_nodes, d_idx[cur], d_nof[cur], d_bb, d_mid, d_sa, d_acc);
        k_decide<<<cdiv(cnt, 128), 128>>>(lvl_start, lvl_end, d_nodes, d_acc, d_ctr, max_nodes);
        k_scan_local<<<ntiles, SCAN_THREADS>>>(n, lvl_start, lvl_end, d_nodes, d_idx[cur], d_nof[cur], d_mid, d_sa, d_S, d_tiles);
        k_scan_tiles<<<1, 1024>>>(ntiles, d_tiles);
        k_scatter<<<cdiv(n, 256), 256>>>(n, lvl_start, lvl_end, d_nodes, d_idx[cur], d_nof[cur], d_mid, d_sa, d_S, d_tiles,
                                         d_idx[cur ^ 1], d_nof[cur ^ 1]);
        BK(cudaGetLastError());
        cur ^= 1;
        BK(cudaMemcpy(&hc, d_ctr, sizeof(hc), cudaMemcpyDeviceToHost));
        if (hc.err & ERR_DEPTH) throw BuildError("bih: recursion too deep (degenerate input)");
        if (hc.err & ERR_NODES) throw BuildError("bih_build_gpu: node table overflow");
        lvl_start = lvl_end;
        lvl_end = hc.n_nodes;
        if (lvl_end > lvl_start) levels.push_back({lvl_start, lvl_end});
        if ((int)levels.size() > BUILD_MAX_DEPTH + 8) throw BuildError("bih: recursion too deep (degenerate input)");
    }
    const int n_nodes = hc.n_nodes;
    int* d_nb = M.alloc<int>(n_nodes);
    int* d_nl = M.alloc<int>(n_nodes);
    int* d_pre = M.alloc<int>(n_nodes);
    int* d_lbase = M.alloc<int>(n_nodes);
    for (int L = (int)levels.size() - 1; L >= 0; L--) {
        const int c = levels[L].second - levels[L].first;
        k_sizes<<<cdiv(c, 256), 256>>>(levels[L].first, levels[L].second, d_nodes, d_nb, d_nl);
    }
    BK(cudaMemsetAsync(d_pre, 0, sizeof(int)));
    BK(cudaMemsetAsync(d_lbase, 0, sizeof(int)));
    for (size_t L = 0; L < levels.size(); L++) {
        const int c = levels[L].second - levels[L].first;
        k_number<<<cdiv(c, 256), 256>>>(levels[L].first, levels[L].second, d_nodes, d_nb, d_nl, d_pre, d_lbase);
    }
    int root_nb = 0, root_nl = 0;
    BK(cudaMemcpy(&root_nb, d_nb, sizeof(int), cudaMemcpyDeviceToHost));
    BK(cudaMemcpy(&root_nl, d_nl, sizeof(int), cudaMemcpyDeviceToHost));
    GlomeBihNode* d_out_nodes = M.alloc<GlomeBihNode>(root_nb);
    int* d_out_leaves = M.alloc<int>(2 * (size_t)root_nl);
    k_emit<<<cdiv(n_nodes, 256), 256>>>(n_nodes, d_nodes, d_pre, d_lbase, d_out_nodes, d_out_leaves);
    BK(cudaGetLastError());
    BK(cudaEventRecord(ev[2]));
    out.nodes.resize(root_nb);
    out.leaves.resize(2 * (size_t)root_nl);
    out.order.resize(n);
    if (root_nb) BK(cudaMemcpy(out.nodes.data(), d_out_nodes, sizeof(GlomeBihNode) * (size_t)root_nb, cudaMemcpyDeviceToHost));
    BK(cudaMemcpy(out.leaves.data(), d_out_leaves, sizeof(int) * 2 * (size_t)root_nl, cudaMemcpyDeviceToHost));
    BK(cudaMemcpy(out.order.data(), d_idx[cur], sizeof(int) * (size_t)n, cudaMemcpyDeviceToHost));
    BK(cudaEventRecord(ev[3]));
    BK(cudaEventSynchronize(ev[3]));
    out.bb = mkbb(vec(hbb[0], hbb[1], hbb[2]), vec(hbb[3], hbb[4], hbb[5]));
    out.root = (root_nb > 0) ? 0 : ~0;
    if (timings_ms) {
        float a = 0, b = 0, c = 0;
        cudaEventElapsedTime(&a, ev[0], ev[1]);
        cudaEventElapsedTime(&b, ev[1], ev[2]);
        cudaEventElapsedTime(&c, ev[2], ev[3]);
        timings_ms[0] = a; timings_ms[1] = b; timings_ms[2] = c;
    }
}

static void mesh_build_gpu_impl(int64_t nverts64, const double* verts, int64_t ntris64, const int32_t* tris, int device, MeshTree& out,
                    double* timings_ms) {
    if (ntris64 < 0 || ntris64 > 0x3fffffff || nverts64 < 0 || nverts64 > 0x3fffffff) throw BuildError("mesh_build_gpu: size out of range");
    if (ntris64 == 0 || nverts64 == 0) {  // nothing to sweep
        mesh_build(nverts64, verts, ntris64, tris, out);
        if (timings_ms) timings_ms[0] = timings_ms[1] = timings_ms[2] = 0;
        return;
    }
    const int n = (int)ntris64, nverts = (int)nverts64;
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || device < 0 || device >= ndev) throw BuildError("mesh_build_gpu: no such CUDA device");
    BK(cudaSetDevice(device));
    DevMem M(device);
    const int max_nodes = 3 * n + 4096;
    double* d_verts = M.alloc<double>(3 * (size_t)nverts);
    int* d_tris = M.alloc<int>(8 * (size_t)n);
    double* d_tbb = M.alloc<double>(6 * (size_t)n);
    double* d_mid = M.alloc<double>(3 * (size_t)n);
    double* d_sa = M.alloc<double>(n);
    int* d_idx[2] = {M.alloc<int>(n), M.alloc<int>(n)};
    int* d_nof[2] = {M.alloc<int>(n), M.alloc<int>(n)};
    int* d_S = M.alloc<int>(n);
    const int ntiles = cdiv(n, SCAN_TILE);
    int* d_tiles = M.alloc<int>(ntiles);
    MNodeD* d_nodes = M.alloc<MNodeD>(max_nodes);
    MeshAcc* d_acc = M.alloc<MeshAcc>((size_t)n / 3 + 8);     // a pending node holds at least 3 triangles
    Counters* d_ctr = M.alloc<Counters>(1);
    unsigned long long* d_bbkeys = M.alloc<unsigned long long>(6);
    double* d_bbout = M.alloc<double>(6);
    cudaEvent_t ev[4];
    for (auto& e : ev) BK(cudaEventCreate(&e));
    struct EvGuard { cudaEvent_t* e; ~EvGuard() { for (int i = 0; i < 4; i++) cudaEventDestroy(e[i]); } } evg{ev};

    BK(cudaEventRecord(ev[0]));
    BK(cudaMemcpy(d_verts, verts, sizeof(double) * 3 * (size_t)nverts, cudaMemcpyHostToDevice));
    BK(cudaMemcpy(d_tris, tris, sizeof(int) * 8 * (size_t)n, cudaMemcpyHostToDevice));
    BK(cudaEventRecord(ev[1]));
    {
        unsigned long long seeds[6];
        auto hk = [](double x) { unsigned long long b; memcpy(&b, &x, 8); return (b >> 63) ? ~b : (b | 0x8000000000000000ull); };
        for (int a = 0; a < 3; a++) { seeds[a] = hk(GLM_INFINITY); seeds[3 + a] = hk(-GLM_INFINITY); }
        BK(cudaMemcpy(d_bbkeys, seeds, sizeof(seeds), cudaMemcpyHostToDevice));
    }
    k_mesh_vbox<<<cdiv(nverts, 256), 256>>>(nverts, d_verts, d_bbkeys);
    k_mesh_prep<<<cdiv(n, 256), 256>>>(n, d_verts, d_tris, d_tbb, d_mid, d_sa, d_idx[0], d_nof[0]);
    k_mesh_init_root<<<1, 1>>>(n, d_bbkeys, d_nodes, d_ctr, d_bbout);
    BK(cudaGetLastError());
    std::vector<std::pair<int, int>> levels;
    int lvl_start = 0, lvl_end = 1, cur = 0;
    Counters hc;
    BK(cudaMemcpy(&hc, d_ctr, sizeof(hc), cudaMemcpyDeviceToHost));
    double hbb[6];
    BK(cudaMemcpy(hbb, d_bbout, sizeof(hbb), cudaMemcpyDeviceToHost));
    levels.push_back({0, 1});
    while (hc.n_pending > 0) {
        const int cnt = lvl_end - lvl_start;
        BK(cudaMemsetAsync(&d_ctr->n_pending, 0, sizeof(int)));
        k_mesh_acc_reset<<<cdiv(hc.n_pending, 256), 256>>>(hc.n_pending, d_acc);
        k_mesh_accum<<<cdiv(n, 256), 256>>>(n, lvl_start, d_nodes, d_idx[cur], d_nof[cur], d_tbb, d_mid, d_sa, d_acc);
        k_mesh_decide<<<cdiv(cnt, 128), 128>>>(lvl_start, lvl_end, d_nodes, d_acc, d_ctr, max_nodes);
        k_scan_local<<<ntiles, SCAN_THREADS>>>(n, lvl_start, lvl_end, d_nodes, d_idx[cur], d_nof[cur], d_mid, d_sa, d_S, d_tiles);
        k_scan_tiles<<<1, 1024>>>(ntiles, d_tiles);
        k_scatter<<<cdiv(n, 256), 256>>>(n, lvl_start, lvl_end, d_nodes, d_idx[cur], d_nof[cur], d_mid, d_sa, d_S, d_tiles,
                                         d_idx[cur ^ 1], d_nof[cur ^ 1]);
        BK(cudaGetLastError());
        cur ^= 1;
        BK(cudaMemcpy(&hc, d_ctr, sizeof(hc), cudaMemcpyDeviceToHost));
        if (hc.err & ERR_DEPTH) throw BuildError("mesh: recursion too deep (degenerate input)");
        if (hc.err & ERR_NODES) throw BuildError("mesh_build_gpu: node table overflow");
        lvl_start = lvl_end;
        lvl_end = hc.n_nodes;
        if (lvl_end > lvl_start) levels.push_back({lvl_start, lvl_end});
        if ((int)levels.size() > BUILD_MAX_DEPTH + 8) throw BuildError("mesh: recursion too deep (degenerate input)");
    }
    const int n_nodes = hc.n_nodes;
    int* d_nb = M.alloc<int>(n_nodes);
    int* d_nl = M.alloc<int>(n_nodes);
    int* d_pre = M.alloc<int>(n_nodes);
    int* d_lbase = M.alloc<int>(n_nodes);
    for (int L = (int)levels.size() - 1; L >= 0; L--) {
        const int c = levels[L].second - levels[L].first;
        k_sizes<<<cdiv(c, 256), 256>>>(levels[L].first, levels[L].second, d_nodes, d_nb, d_nl);
    }
    BK(cudaMemsetAsync(d_pre, 0, sizeof(int)));
    BK(cudaMemsetAsync(d_lbase, 0, sizeof(int)));
    for (size_t L = 0; L < levels.size(); L++) {
        const int c = levels[L].second - levels[L].first;
        k_number<<<cdiv(c, 256), 256>>>(levels[L].first, levels[L].second, d_nodes, d_nb, d_nl, d_pre, d_lbase);
    }
    int root_nb = 0, root_nl = 0;
    BK(cudaMemcpy(&root_nb, d_nb, sizeof(int), cudaMemcpyDeviceToHost));
    BK(cudaMemcpy(&root_nl, d_nl, sizeof(int), cudaMemcpyDeviceToHost));
    GlomeBvhNode* d_out_nodes = M.alloc<GlomeBvhNode>(root_nb);
    int* d_leafpool = M.alloc<int>((size_t)root_nl + n);
    int* d_leafoff = M.alloc<int>(root_nl);
    k_mesh_emit<<<cdiv(n_nodes, 256), 256>>>(n_nodes, d_nodes, d_pre, d_lbase, d_idx[cur], d_out_nodes, d_leafpool, d_leafoff);
    BK(cudaGetLastError());
    BK(cudaEventRecord(ev[2]));
    out.nodes.resize(root_nb);
    out.leafpool.resize((size_t)root_nl + n);
    out.leafoff.resize(root_nl);
    if (root_nb) BK(cudaMemcpy(out.nodes.data(), d_out_nodes, sizeof(GlomeBvhNode) * (size_t)root_nb, cudaMemcpyDeviceToHost));
    BK(cudaMemcpy(out.leafpool.data(), d_leafpool, sizeof(int) * ((size_t)root_nl + n), cudaMemcpyDeviceToHost));
    BK(cudaMemcpy(out.leafoff.data(), d_leafoff, sizeof(int) * (size_t)root_nl, cudaMemcpyDeviceToHost));
    BK(cudaEventRecord(ev[3]));
    BK(cudaEventSynchronize(ev[3]));
    out.bb = mkbb(vec(hbb[0], hbb[1], hbb[2]), vec(hbb[3], hbb[4], hbb[5]));
    out.root = (root_nb > 0) ? 0 : ~0;
    if (timings_ms) {
        float a = 0, b = 0, c = 0;
        cudaEventElapsedTime(&a, ev[0], ev[1]);
        cudaEventElapsedTime(&b, ev[1], ev[2]);
        cudaEventElapsedTime(&c, ev[2], ev[3]);
        timings_ms[0] = a; timings_ms[1] = b; timings_ms[2] = c;
    }
}

}  // namespace glome_host
// give the builders' cached device work space of `device` back to the driver (it is otherwise kept for the next build)
extern "C" int glome_build_release_cache(int device) {
    if (device < 0 || device > 63) return GLOME_EINVAL;
    if (cudaSetDevice(device) != cudaSuccess) { cudaGetLastError(); return GLOME_ENODEV; }
    glome_host::DevMem::release_cache(device);
    return GLOME_OK;
}
namespace glome_host {

// libglomecuda.so hands its builders to libglomehost.so when it is loaded (host_base.cpp)
extern void (*hook_bih_build_gpu)(int64_t, const double*, int, BihTree&, double*);
extern void (*hook_mesh_build_gpu)(int64_t, const double*, int64_t, const int32_t*, int, MeshTree&, double*);
namespace {
struct RegisterGpuBuilders {
    RegisterGpuBuilders() { hook_bih_build_gpu = &bih_build_gpu_impl; hook_mesh_build_gpu = &mesh_build_gpu_impl; }
} g_register_gpu_builders;
}  // namespace

}  // namespace glome_host
