// glome_tagmap.h -- host-side preparation of a general-class FlatScene for the scene-graph machine (glome_gen.cuh).
//
// The machine keeps texture / tag stacks as 8 x 16-bit packed ids.  Texture ids are scene indices already; tag VALUES are
// arbitrary int32 in the reference's API (`tag :: SolidItem t m -> t -> SolidItem t m`, Tex.hs:38), so the device copy of a
// general-class scene carries dense tag ids (in Tag nodes and in the per-Mesh tag tables) plus the table that maps them
// back when a hit leaves the device.
#pragma once
#include <stdint.h>
#include <string.h>

#include <algorithm>
#include <cmath>
#include <set>
#include <string>
#include <unordered_map>
#include <vector>

#include "../../include/glome_cuda.h"

namespace glome_tagmap {

#define GLOME_GEN_MAX_ID 65534
#define GLOME_GEN_MAX_LIGHTS 64

// nodes / ipool: copies of the scene's arrays, rewritten in place.  Returns "" or the reason the scene is over a limit.
inline std::string remap_tags(const GlomeFlatScene* d, std::vector<GlomeNode>& nodes, std::vector<int32_t>& ipool,
                              std::vector<int32_t>& tagvals) {
    nodes.assign(d->nodes, d->nodes + d->n_nodes);
    ipool.assign(d->ipool, d->ipool + d->n_ipool);
    tagvals.clear();
    if (d->n_textures > GLOME_GEN_MAX_ID) return "general scene: more than 65534 textures";
    if (d->n_nodes >= (1 << 27)) return "general scene: more than 2^27 nodes";
    std::unordered_map<int32_t, int32_t> ids;
    auto dense = [&](int32_t v) {
        auto it = ids.find(v);
        if (it != ids.end()) return it->second;
        int32_t k = (int32_t)tagvals.size();
        ids[v] = k;
        tagvals.push_back(v);
        return k;
    };
    std::set<int32_t> done;
    for (int i = 0; i < d->n_nodes; i++) {
        GlomeNode& n = nodes[i];
        if (n.type == GLOME_TAG) n.b = dense(n.b);
        else if (n.type == GLOME_MESH) {
            if (!done.insert(n.a).second) continue;
            const GlomeMeshHeader* h = reinterpret_cast<const GlomeMeshHeader*>(d->ipool + n.a);
            for (int k = 0; k < h->ntags; k++) ipool[h->tags_off + k] = dense(d->ipool[h->tags_off + k]);
        }
    }
    if ((int)tagvals.size() > GLOME_GEN_MAX_ID) return "general scene: more than 65534 distinct tags";
    if (tagvals.empty()) tagvals.push_back(0);
    return "";
}

// Item descriptors: one 16-byte record per node, so that a list element (group child, BIH leaf item, CSG operand ...) is
// classified with ONE load instead of a walk down its Tex / Tag / NoShadow / OnlyShadow wrappers:
//   x = class | prim type << 4 | flags        y = core node (the primitive, or the first non-wrapper node)
//   z = dpool offset of the primitive record   w = dpool offset of the Instance's Xfm (GI_INST_PRIM)
enum { GI_DEAD = 0, GI_PRIM = 1, GI_INST_PRIM = 2, GI_COMPLEX = 3 };
#define GI_VIS_R 0x100   /* visible to rayint: no OnlyShadow in the chain (Tex.hs:89)                  */
#define GI_VIS_S 0x200   /* visible to shadow: no NoShadow in the chain, not a Mesh (Tex.hs:81, Mesh.hs:210) */
#define GI_WRAPPED 0x400 /* a Tex or Tag wrapper in the chain: the winner's stacks need the walk        */
#define GI_HAS_META 0x800 /* a Tex or Tag node somewhere in the subtree: get_metainfo (Solid.hs:200) can return something;
                             without it the answer is ([],[]) and the walk is skipped                    */

// does the subtree under node i hold a Tex or Tag?  (memoised; a Bih is answered conservatively unless it is a block of
// bare spheres)
inline bool has_meta(const std::vector<GlomeNode>& nodes, std::vector<signed char>& memo, int i) {
    if (memo[i] >= 0) return memo[i] != 0;
    memo[i] = 1;  // (cycles cannot occur in a scene graph; be conservative if one does)
    const GlomeNode& n = nodes[i];
    bool r = false;
    switch (n.type) {
        case GLOME_TEX: case GLOME_TAG: r = true; break;
        case GLOME_NOSHADOW: case GLOME_ONLYSHADOW: case GLOME_INSTANCE: r = has_meta(nodes, memo, n.a); break;
        case GLOME_GROUP: case GLOME_INTERSECTION:
            for (int k = 0; k < n.b && !r; k++) r = has_meta(nodes, memo, n.a + k);
            break;
        case GLOME_DIFFERENCE: case GLOME_BOUND: case GLOME_INNERBOUND: r = has_meta(nodes, memo, n.a) || has_meta(nodes, memo, n.b); break;
        case GLOME_BIH: r = !(n.c & GLOME_BIH_LINEAR_SPHERES); break;
        default: r = false; break;  // primitives, Void, Mesh: ([],[]) (Solid.hs:254)
    }
    memo[i] = r ? 1 : 0;
    return r;
}

inline void build_items(const std::vector<GlomeNode>& nodes, std::vector<int32_t>& items) {
    const int n = (int)nodes.size();
    items.assign((size_t)n * 4, 0);
    std::vector<signed char> memo((size_t)n, -1);
    auto is_prim = [](int t) { return t >= GLOME_SPHERE && t <= GLOME_CONE; };
    for (int i = 0; i < n; i++) {
        int vis = GI_VIS_R | GI_VIS_S, wrapped = 0;
        int cj = i;
        GlomeNode c = nodes[cj];
        auto peel = [&]() {
            while (c.type == GLOME_TEX || c.type == GLOME_TAG || c.type == GLOME_NOSHADOW || c.type == GLOME_ONLYSHADOW) {
                if (c.type == GLOME_TEX || c.type == GLOME_TAG) wrapped = GI_WRAPPED;
                if (c.type == GLOME_NOSHADOW) vis &= ~GI_VIS_S;
                if (c.type == GLOME_ONLYSHADOW) vis &= ~GI_VIS_R;
                cj = c.a;
                c = nodes[cj];
            }
        };
        peel();
        int32_t* o = &items[(size_t)i * 4];
        const int meta = has_meta(nodes, memo, i) ? GI_HAS_META : 0;
        struct AddMeta { int32_t* o; int meta; ~AddMeta() { o[0] |= meta; } } addmeta{o, meta};
        if (is_prim(c.type)) { o[0] = GI_PRIM | (c.type << 4) | vis | wrapped; o[1] = cj; o[2] = c.a; o[3] = 0; continue; }
        if (c.type == GLOME_VOID) { o[0] = GI_DEAD; o[1] = cj; continue; }
        if (c.type == GLOME_INSTANCE) {
            const int inst = cj, vis0 = vis, wr0 = wrapped;
            const int xfm = c.b;
            cj = c.a;
            c = nodes[cj];
            peel();
            if (is_prim(c.type)) { o[0] = GI_INST_PRIM | (c.type << 4) | vis | wrapped; o[1] = cj; o[2] = c.a; o[3] = xfm; continue; }
            if (c.type == GLOME_VOID) { o[0] = GI_DEAD; o[1] = cj; continue; }
            vis = vis0; wrapped = wr0;  // a complex child: the Instance itself is the core
            o[0] = GI_COMPLEX | vis | wrapped; o[1] = inst; o[2] = 0; o[3] = 0;
            continue;
        }
        if (c.type == GLOME_MESH) vis &= ~GI_VIS_S;
        o[0] = GI_COMPLEX | vis | wrapped; o[1] = cj; o[2] = 0; o[3] = 0;
    }
}

// ---------------------------------------------------------------------------------------------
// Implicit BIH over large plain groups.
//
// `group xs` (Solid.hs:293-302, 326-331) tests every element for every ray: TestScene's chessboard is a group of 64 boxes
// under a Difference, and about 60 % of that scene's primitive tests are spent there.  The result of the fold is
// min-depth with ties going to the LATER list element (Solid.hs:37-44), whatever order the elements are visited in, so
// a group of simple items (`{Tex,Tag}* prim`, `Instance` of one) gets a small BIH of our own over (padded) bounding
// boxes; the machine walks it like any Bih and keeps the list position as the tie key (glome_gen.cuh).  Rays the BIH
// arithmetic does not serve exactly like the list would (a zero direction component -- the reference's own Bih misses
// those, SURVEY A3 -- or a non-unit direction) take the plain list.  rayint_debug, inside and get_metainfo keep reading
// the group as the list it is.
//
// Appends to: nodes / items (copies of the children in leaf order), bih (the tree), dpool (the group's box), ipool
// (per group {root ref, dpool offset of the box, first copy, offset of the list positions, count} + the positions).
// The group's node gets c = (ipool offset << 4) | 2.
// ---------------------------------------------------------------------------------------------
#ifndef GLOME_GROUP_ACCEL
#define GLOME_GROUP_ACCEL 2
#endif
struct AccelBox { double lo[3], hi[3]; };

inline bool simple_item_box(const std::vector<GlomeNode>& nodes, const int32_t* it, const std::vector<double>& dpool, AccelBox& out) {
    const int cls = it[0] & 15, type = (it[0] >> 4) & 15;
    if (cls != GI_PRIM && cls != GI_INST_PRIM) return false;
    const double* p = dpool.data() + it[2];
    double lo[3], hi[3];
    switch (type) {
        case GLOME_SPHERE: for (int a = 0; a < 3; a++) { lo[a] = p[a] - p[3]; hi[a] = p[a] + p[3]; } break;
        case GLOME_TRIANGLE: case GLOME_TRIANGLENORM:
            for (int a = 0; a < 3; a++) { lo[a] = std::min(p[a], std::min(p[3 + a], p[6 + a])); hi[a] = std::max(p[a], std::max(p[3 + a], p[6 + a])); }
            break;
        case GLOME_BOX: for (int a = 0; a < 3; a++) { lo[a] = std::min(p[a], p[3 + a]); hi[a] = std::max(p[a], p[3 + a]); } break;
        case GLOME_DISC: { const double r = std::sqrt(std::fabs(p[6])); for (int a = 0; a < 3; a++) { lo[a] = p[a] - r; hi[a] = p[a] + r; } break; }
        case GLOME_CYLINDER: lo[0] = lo[1] = -std::fabs(p[0]); hi[0] = hi[1] = std::fabs(p[0]); lo[2] = std::min(p[1], p[2]); hi[2] = std::max(p[1], p[2]); break;
        case GLOME_CONE: {  // radius r at z = 0 shrinking to 0 at z = height, clipped to [c1, c2]
            const double h = p[3], r = std::fabs(p[0]);
            double f = 1;
            if (h != 0) f = std::max(std::fabs(1 - p[1] / h), std::fabs(1 - p[2] / h));
            const double rm = r * std::max(1.0, f);
            lo[0] = lo[1] = -rm; hi[0] = hi[1] = rm; lo[2] = std::min(p[1], p[2]); hi[2] = std::max(p[1], p[2]);
            break;
        }
        default: return false;  // Plane: unbounded
    }
    if (cls == GI_INST_PRIM) {  // the eight corners through the forward matrix
        const double* m = dpool.data() + it[3];
        double nlo[3] = {1e300, 1e300, 1e300}, nhi[3] = {-1e300, -1e300, -1e300};
        for (int k = 0; k < 8; k++) {
            const double x = (k & 1) ? hi[0] : lo[0], y = (k & 2) ? hi[1] : lo[1], z = (k & 4) ? hi[2] : lo[2];
            for (int a = 0; a < 3; a++) {
                const double v = m[4 * a] * x + m[4 * a + 1] * y + m[4 * a + 2] * z + m[4 * a + 3];
                nlo[a] = std::min(nlo[a], v); nhi[a] = std::max(nhi[a], v);
            }
        }
        for (int a = 0; a < 3; a++) { lo[a] = nlo[a]; hi[a] = nhi[a]; }
    }
    for (int a = 0; a < 3; a++) {
        if (!(lo[a] > -1e30 && hi[a] < 1e30) || lo[a] != lo[a] || hi[a] != hi[a]) return false;
        // padding far above the rounding of any intersection test and of the FP32 twin's payloads
        const double pad = 1e-3 + 1e-5 * std::max(std::fabs(lo[a]), std::fabs(hi[a]));
        out.lo[a] = lo[a] - pad; out.hi[a] = hi[a] + pad;
    }
    return true;
}

struct AccelBuild {
    std::vector<AccelBox> box;
    std::vector<int> order;          // positions in the list, permuted into leaf order
    std::vector<GlomeBihNode>* bih;
    int copy_base;                   // node index of the first copy
    // returns a child ref: >= 0 node index (global), < 0 inline leaf over copies [lo, hi)
    int32_t rec(int lo, int hi) {
        const int n = hi - lo;
        double clo[3] = {1e300, 1e300, 1e300}, chi[3] = {-1e300, -1e300, -1e300};
        for (int i = lo; i < hi; i++)
            for (int a = 0; a < 3; a++) {
                const double c = 0.5 * (box[order[i]].lo[a] + box[order[i]].hi[a]);
                clo[a] = std::min(clo[a], c); chi[a] = std::max(chi[a], c);
            }
        int ax = 0;
        for (int a = 1; a < 3; a++) if (chi[a] - clo[a] > chi[ax] - clo[ax]) ax = a;
        if (n <= 4 || !(chi[ax] - clo[ax] > 0)) {
            if (n <= 6) return glome_bih_leaf_ref_inline(copy_base + lo, n);
            // more than six coincident centres: split by position (any split is correct, the planes come from the boxes)
        }
        const int mid = lo + n / 2;
        if (chi[ax] - clo[ax] > 0)
            std::nth_element(order.begin() + lo, order.begin() + mid, order.begin() + hi, [&](int x, int y) {
                const double cx = box[x].lo[ax] + box[x].hi[ax], cy = box[y].lo[ax] + box[y].hi[ax];
                return cx < cy || (cx == cy && x < y);
            });
        GlomeBihNode nd;
        memset(&nd, 0, sizeof(nd));
        nd.axis = ax;
        double lmax = -1e300, rmin = 1e300;
        for (int i = lo; i < mid; i++) lmax = std::max(lmax, box[order[i]].hi[ax]);
        for (int i = mid; i < hi; i++) rmin = std::min(rmin, box[order[i]].lo[ax]);
        nd.lsplit = lmax; nd.rsplit = rmin;
        const int me = (int)bih->size();
        bih->push_back(nd);
        const int32_t l = rec(lo, mid), r = rec(mid, hi);
        (*bih)[me].left = l; (*bih)[me].right = r;
        return me;
    }
};

inline void build_group_accels(std::vector<GlomeNode>& nodes, std::vector<int32_t>& items, std::vector<int32_t>& ipool,
                               std::vector<GlomeBihNode>& bih, std::vector<double>& dpool, int min_items = 12) {
    const int n0 = (int)nodes.size();
    for (int g = 0; g < n0; g++) {
        if (nodes[g].type != GLOME_GROUP || nodes[g].b < min_items || nodes[g].c != 0) continue;
        const int first = nodes[g].a, count = nodes[g].b;
        AccelBuild B;
        std::vector<int> live;  // list positions that can be hit at all (Void elements never are)
        bool ok = true;
        B.box.resize(count);
        for (int k = 0; k < count && ok; k++) {
            const int32_t* it = &items[(size_t)(first + k) * 4];
            if ((it[0] & 15) == GI_DEAD) continue;
            ok = simple_item_box(nodes, it, dpool, B.box[k]);
            live.push_back(k);
        }
        if (!ok || (int)live.size() < min_items) continue;
        B.order = live;
        B.bih = &bih;
        B.copy_base = (int)nodes.size();
        const int32_t root = B.rec(0, (int)live.size());
        // copies of the children in leaf order (a copy's wrapper chain continues into the original nodes)
        for (size_t i = 0; i < B.order.size(); i++) {
            const int src = first + B.order[i];
            nodes.push_back(nodes[src]);
            for (int q = 0; q < 4; q++) items.push_back(items[(size_t)src * 4 + q]);
        }
        if (dpool.size() & 1) dpool.push_back(0);  // 16-byte alignment of the box record
        const int bb_off = (int)dpool.size();
        double lo[3] = {1e300, 1e300, 1e300}, hi[3] = {-1e300, -1e300, -1e300};
        for (int k : live) for (int a = 0; a < 3; a++) { lo[a] = std::min(lo[a], B.box[k].lo[a]); hi[a] = std::max(hi[a], B.box[k].hi[a]); }
        for (int a = 0; a < 3; a++) dpool.push_back(lo[a]);
        for (int a = 0; a < 3; a++) dpool.push_back(hi[a]);
        const int ao = (int)ipool.size();
        ipool.push_back(root); ipool.push_back(bb_off); ipool.push_back(B.copy_base); ipool.push_back(ao + 5); ipool.push_back((int)live.size());
        for (size_t i = 0; i < B.order.size(); i++) ipool.push_back(B.order[i]);
        nodes[g].c = (ao << 4) | GLOME_GROUP_ACCEL;
    }
}

}  // namespace glome_tagmap
