// glome_tagmap.h -- host-side preparation of a general-class FlatScene for the scene-graph machine (glome_gen.cuh).
//
// The machine keeps texture / tag stacks as 8 x 16-bit packed ids.  Texture ids are scene indices already; tag VALUES are
// arbitrary int32 in the reference's API (`tag :: SolidItem t m -> t -> SolidItem t m`, Tex.hs:38), so the device copy of a
// general-class scene carries dense tag ids (in Tag nodes and in the per-Mesh tag tables) plus the table that maps them
// back when a hit leaves the device.
#pragma once
#include <stdint.h>

#include <algorithm>
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

}  // namespace glome_tagmap
