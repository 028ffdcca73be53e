// gen_host.cpp -- TEST INFRASTRUCTURE ONLY: the scene-graph machine of glome_b200/csrc/glome_gen.cuh compiled by g++ so
// that its control flow (the explicit stacks that replaced the reference's recursion) can be checked against the oracle
// on a machine without a GPU.  Built by tests/ into tests/tools/_build/libgenhost.so; nothing under glome_b200/ loads,
// links or calls it, and no GPU test uses it: the product runs the same header inside its sm_100a kernels only.
#include <stdlib.h>
#include <string.h>

#include <atomic>
#include <string>
#include <thread>
#include <vector>

#include "../../glome_b200/csrc/glome_gen.cuh"
#include "../../glome_b200/csrc/glome_tagmap.h"

using namespace ggen;

struct HostScene {
    std::vector<GlomeNode> nodes;
    std::vector<int32_t> ipool, tagvals, lightsets, items;
    std::vector<GlomeBihNode> bihv;
    std::vector<double> dpv;
    DScene d;
};

template <typename F>
static void par_for(int64_t n, int threads, F f) {
    if (threads < 1) threads = 1;
    std::atomic<int64_t> next(0);
    std::vector<std::thread> th;
    auto body = [&]() {
        SHM* sh = new SHM();
        for (;;) {
            int64_t i0 = next.fetch_add(64);
            if (i0 >= n) break;
            for (int64_t i = i0; i < n && i < i0 + 64; i++) f(*sh, i);
        }
        delete sh;
    };
    for (int t = 1; t < threads; t++) th.emplace_back(body);
    body();
    for (auto& t : th) t.join();
}
static inline Ray ldray(const double* rays, int64_t i) {
    const double* p = rays + 6 * i;
    return mkray(vec(p[0], p[1], p[2]), vec(p[3], p[4], p[5]));
}

extern "C" {

void* genh_create(const GlomeFlatScene* fs) {
    HostScene* h = new HostScene();
    std::string e = glome_tagmap::remap_tags(fs, h->nodes, h->ipool, h->tagvals);
    if (!e.empty()) { delete h; return nullptr; }
    if (fs->n_lightsets > 0) h->lightsets.assign(fs->lightsets, fs->lightsets + 2 * fs->n_lightsets);
    else { h->lightsets.push_back(0); h->lightsets.push_back(fs->n_lights); }
    glome_tagmap::build_items(h->nodes, h->items);
    h->bihv.assign(fs->bihnodes, fs->bihnodes + fs->n_bihnodes);
    h->dpv.assign(fs->dpool, fs->dpool + fs->n_dpool);
    {   // the implicit BIHs over large plain groups, as glome_scene_create builds them (GLOME_GROUP_ACCEL=0: plain lists)
        const char* e = getenv("GLOME_GROUP_ACCEL");
        if (!e || atoi(e) != 0) glome_tagmap::build_group_accels(h->nodes, h->items, h->ipool, h->bihv, h->dpv);
    }
    memset(&h->d, 0, sizeof(h->d));
    h->d.items = reinterpret_cast<const int4*>(h->items.data());
    h->d.nodes = h->nodes.data(); h->d.bih = h->bihv.data(); h->d.bvh = fs->bvhnodes; h->d.ipool = h->ipool.data();
    h->d.dpool = h->dpv.data(); h->d.textures = fs->textures; h->d.materials = fs->materials; h->d.lights = fs->lights;
    h->d.lightsets = h->lightsets.data(); h->d.tagvals = h->tagvals.data(); h->d.root = fs->root; h->d.n_lights = fs->n_lights;
    return h;
}
void genh_destroy(void* p) { delete (HostScene*)p; }

void genh_rayint_batch(void* hp, int64_t n, const double* rays, const double* tmax, int stride, GlomeHit* out, int threads) {
    HostScene* h = (HostScene*)hp;
    par_for(n, threads, [&](SHM& sh, int64_t i) {
        GCnt c; gcnt_clear(c);
        gq_query(h->d, sh.q, h->d.root, ldray(rays, i), tmax[stride ? i : 0], false, c);
        ghit_out(h->d, sh.q.slot[0], out + i);
    });
}
void genh_shadow_batch(void* hp, int64_t n, const double* rays, const double* tmax, int stride, uint8_t* out, int threads) {
    HostScene* h = (HostScene*)hp;
    par_for(n, threads, [&](SHM& sh, int64_t i) {
        GCnt c; gcnt_clear(c);
        out[i] = gq_query(h->d, sh.q, h->d.root, ldray(rays, i), tmax[stride ? i : 0], true, c) ? 1 : 0;
    });
}
void genh_inside_batch(void* hp, int64_t n, const double* pts, uint8_t* out, int threads) {
    HostScene* h = (HostScene*)hp;
    par_for(n, threads, [&](SHM&, int64_t i) {
        int ovf = 0;
        out[i] = gq_inside(h->d, h->d.root, vec(pts[3 * i], pts[3 * i + 1], pts[3 * i + 2]), &ovf) ? 1 : 0;
    });
}
void genh_debug_count_batch(void* hp, int64_t n, const double* rays, const double* tmax, int stride, int32_t* out, int threads) {
    HostScene* h = (HostScene*)hp;
    par_for(n, threads, [&](SHM& sh, int64_t i) {
        GCnt c; gcnt_clear(c);
        int ovf = 0;
        out[i] = gq_debug_count(h->d, sh.q, h->d.root, ldray(rays, i), tmax[stride ? i : 0], c, &ovf);
    });
}
// tags: 17 int32 per ray {count, tags...} or null
void genh_trace_batch(void* hp, int64_t n, const double* rays, const double* tmax, int stride, int recurs, double* rgba,
                      double* depth, GlomeHit* hits, int32_t* tags, int64_t* counters, int threads) {
    HostScene* h = (HostScene*)hp;
    std::atomic<int64_t> nshadow(0), nsec(0), ncsg(0), nbih(0), nprim(0), ninst(0);
    par_for(n, threads, [&](SHM& sh, int64_t i) {
        GCnt c; gcnt_clear(c);
        ColorA col;
        int fl = 0;
        TagArena ta;
        if (tags) gs_trace<true>(h->d, sh, 0, h->d.root, ldray(rays, i), tmax[stride ? i : 0], recurs, col, c, fl, &ta);
        else gs_trace<false>(h->d, sh, 0, h->d.root, ldray(rays, i), tmax[stride ? i : 0], recurs, col, c, fl, nullptr);
        rgba[4 * i] = col.r; rgba[4 * i + 1] = col.g; rgba[4 * i + 2] = col.b; rgba[4 * i + 3] = col.a;
        depth[i] = ghit_depth(sh.tf[0].ri);
        if (hits) { ghit_out(h->d, sh.tf[0].ri, hits + i); hits[i].flags |= fl; }
        if (tags) {
            int32_t* o = tags + 17 * i;
            o[0] = ta.n < 16 ? ta.n : 16;  // (the oracle's list is capped at 16 entries)
            for (int k = 0; k < 16; k++) o[1 + k] = k < ta.n ? h->d.tagvals[ta.v[k]] : -1;
        }
        nshadow += c.shadow; nsec += c.secondary; ncsg += c.csg; nbih += c.bih; nprim += c.prim; ninst += c.inst;
    });
    if (counters) { counters[0] = nshadow; counters[1] = nsec; counters[2] = ncsg; counters[3] = nbih; counters[4] = nprim; counters[5] = ninst; }
}

}  // extern "C"
