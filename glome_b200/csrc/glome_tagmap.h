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

}  // namespace glome_tagmap
