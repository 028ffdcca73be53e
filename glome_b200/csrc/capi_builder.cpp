// capi_builder.cpp -- extern "C" surface of the host scene-construction mirror (include/glome_cuda.h).
// Haskell `error` calls in the mirrored constructors surface as GLOME_EBUILD + glome_last_error().
#include <cstdlib>
#include <cstring>
#include <string>

#include "host_builder.h"
#include "glome_build.h"

using namespace glome_host;
using glm::Vec;
using glm::vec;

void glome_set_error(const std::string& s);  // glome_cuda.cu

struct GlomeBuilder { Builder b; };

namespace glome_host {  // scenes.cpp: the StdGen of random-1.2 (TestScene's oak)
struct StdGen { uint64_t seed, gamma; };
StdGen mkStdGen(int64_t n);
uint64_t stdgen_next_word64(StdGen& g);
void stdgen_split(const StdGen& g, StdGen& a, StdGen& b);
double stdgen_randomR_double(double l, double h, StdGen& g);
uint64_t splitmix64_vigna(uint64_t& x);
}

// no exception crosses the C boundary (a Haskell or ctypes caller would be terminated by it)
#define GUARD_CATCH                                                                          \
    catch (const BuildError& e) { glome_set_error(e.msg); return GLOME_EBUILD; }              \
    catch (const std::exception& e) { glome_set_error(e.what()); return GLOME_EBUILD; }       \
    catch (...) { glome_set_error("unknown exception"); return GLOME_EBUILD; }
#define GUARD(expr)        \
    try { return (expr); } \
    GUARD_CATCH

static inline Vec V(const double* p) { return vec(p[0], p[1], p[2]); }
static inline Xfm X(const double* p) { Xfm x; memcpy(x.m, p, sizeof(x.m)); return x; }
static std::vector<int32_t> IV(int n, const int32_t* p) { return std::vector<int32_t>(p, p + (n > 0 ? n : 0)); }

extern "C" {

// System.Random (random-1.2) as restated for TestScene's oak: out = {seed, gamma of mkStdGen n; the next three Word64 of
// that generator; seed, gamma of both halves of `split`}; dout = three successive `randomR (0, 0.5)` on mkStdGen n;
// vigna = the first output of Vigna's splitmix64 started at x = n (a published known answer for the mixing function).
int glome_stdgen_probe(int64_t n, uint64_t out[9], double dout[3], uint64_t* vigna) {
    if (!out || !dout || !vigna) return GLOME_EINVAL;
    StdGen g = mkStdGen(n);
    out[0] = g.seed; out[1] = g.gamma;
    StdGen h = g;
    for (int i = 0; i < 3; i++) out[2 + i] = stdgen_next_word64(h);
    StdGen a, b;
    stdgen_split(g, a, b);
    out[5] = a.seed; out[6] = a.gamma; out[7] = b.seed; out[8] = b.gamma;
    StdGen d = g;
    for (int i = 0; i < 3; i++) dout[i] = stdgen_randomR_double(0, 0.5, d);
    uint64_t x = (uint64_t)n;
    *vigna = splitmix64_vigna(x);
    return GLOME_OK;
}

int glome_builder_create(GlomeBuilder** out) {
    if (!out) return GLOME_EINVAL;
    *out = new GlomeBuilder();
    return GLOME_OK;
}
int glome_builder_destroy(GlomeBuilder* b) { delete b; return GLOME_OK; }
int glome_builder_set_build_device(GlomeBuilder* b, int device) {
    if (!b) return GLOME_EINVAL;
    b->b.build_device = device;
    return GLOME_OK;
}
int glome_builder_last_build_ms(GlomeBuilder* b, double out[4]) {
    if (!b || !out) return GLOME_EINVAL;
    memcpy(out, b->b.build_ms, sizeof(double) * 4);
    return GLOME_OK;
}

int glome_sb_void(GlomeBuilder* b) { GUARD(b->b.void_()); }
int glome_sb_sphere(GlomeBuilder* b, const double c[3], double r) { GUARD(b->b.sphere(V(c), r)); }
int glome_sb_spheres(GlomeBuilder* b, int64_t n, const double* centers, const double* radii, int32_t* ids_out) {
    try {
        b->b.items.reserve(b->b.items.size() + (size_t)n);
        for (int64_t i = 0; i < n; i++) ids_out[i] = b->b.sphere(V(centers + 3 * i), radii[i]);
        return GLOME_OK;
    } GUARD_CATCH
}
int glome_sb_triangle(GlomeBuilder* b, const double p[9]) { GUARD(b->b.triangle(V(p), V(p + 3), V(p + 6))); }
int glome_sb_trianglenorm(GlomeBuilder* b, const double p[18]) {
    GUARD(b->b.trianglenorm(V(p), V(p + 3), V(p + 6), V(p + 9), V(p + 12), V(p + 15)));
}
int glome_sb_box(GlomeBuilder* b, const double p1[3], const double p2[3]) { GUARD(b->b.box(V(p1), V(p2))); }
int glome_sb_plane(GlomeBuilder* b, const double orig[3], const double norm[3]) { GUARD(b->b.plane(V(orig), V(norm))); }
int glome_sb_plane_offset(GlomeBuilder* b, const double norm[3], double off) { GUARD(b->b.plane_offset(V(norm), off)); }
int glome_sb_disc(GlomeBuilder* b, const double pos[3], const double norm[3], double r) { GUARD(b->b.disc(V(pos), V(norm), r)); }
int glome_sb_cylinder(GlomeBuilder* b, const double p1[3], const double p2[3], double r) { GUARD(b->b.cylinder(V(p1), V(p2), r)); }
int glome_sb_cone(GlomeBuilder* b, const double p1[3], double r1, const double p2[3], double r2) {
    GUARD(b->b.cone(V(p1), r1, V(p2), r2));
}
int glome_sb_cylinder_z(GlomeBuilder* b, double r, double h1, double h2) { GUARD(b->b.cylinder_z(r, h1, h2)); }
int glome_sb_cone_z(GlomeBuilder* b, double r, double h1, double h2, double height) { GUARD(b->b.cone_z(r, h1, h2, height)); }
int glome_sb_group(GlomeBuilder* b, int n, const int32_t* items) { GUARD(b->b.group(IV(n, items))); }
// raw forms: the data an already-constructed Haskell value holds, taken as is (no constructor logic re-run)
int glome_sb_list(GlomeBuilder* b, int n, const int32_t* items) { GUARD(b->b.list_raw(IV(n, items))); }
int glome_sb_instance(GlomeBuilder* b, int item, const double xfm[24]) { GUARD(b->b.instance_raw(item, X(xfm))); }
int glome_sb_disc_raw(GlomeBuilder* b, const double pos[3], const double norm[3], double rsqr) {
    GUARD(b->b.disc_raw(V(pos), V(norm), rsqr));
}
int glome_sb_difference_ex(GlomeBuilder* b, int sa, int sb, int useatex) { GUARD(b->b.difference_ex(sa, sb, useatex != 0)); }
int glome_sb_bih(GlomeBuilder* b, int64_t n, const int32_t* items) { GUARD(b->b.bih(std::vector<int32_t>(items, items + n))); }
int glome_sb_mesh(GlomeBuilder* b, int64_t nverts, const double* verts, int64_t nnorms, const double* norms, int64_t ntris,
                  const int32_t* tris, int ntexs, const int32_t* texs, int ntags, const int32_t* tags) {
    GUARD(b->b.mesh(nverts, verts, nnorms, norms, ntris, tris, ntexs, texs, ntags, tags));
}
int glome_sb_bih_prebuilt(GlomeBuilder* b, int64_t n, const int32_t* items, int64_t n_nodes, const int32_t* kinds,
                          const double* splits, const double bb[6]) {
    if (!b || n < 0 || (n > 0 && !items) || !kinds || !splits || !bb) return GLOME_EINVAL;
    GUARD(b->b.bih_prebuilt(std::vector<int32_t>(items, items + n), n_nodes, kinds, splits, bb));
}
int glome_sb_mesh_prebuilt(GlomeBuilder* b, int64_t nverts, const double* verts, int64_t nnorms, const double* norms,
                           int64_t ntris, const int32_t* tris, int ntexs, const int32_t* texs, int ntags, const int32_t* tags,
                           int64_t n_nodes, const int32_t* kinds, const double* boxes, int64_t n_leaf_tris,
                           const int32_t* leaf_tris, const double bb[6]) {
    if (!b || !verts || !tris || !kinds || !boxes || (n_leaf_tris > 0 && !leaf_tris) || !bb) return GLOME_EINVAL;
    GUARD(b->b.mesh_prebuilt(nverts, verts, nnorms, norms, ntris, tris, ntexs, texs, ntags, tags, n_nodes, kinds, boxes,
                             n_leaf_tris, leaf_tris, bb));
}
int glome_sb_difference(GlomeBuilder* b, int sa, int sb) { GUARD(b->b.difference(sa, sb)); }
int glome_sb_intersection(GlomeBuilder* b, int n, const int32_t* items) { GUARD(b->b.intersection(IV(n, items))); }
int glome_sb_tex(GlomeBuilder* b, int item, int texture) { GUARD(b->b.tex(item, texture)); }
int glome_sb_tag(GlomeBuilder* b, int item, int tag) { GUARD(b->b.tag(item, tag)); }
int glome_sb_noshadow(GlomeBuilder* b, int item) { GUARD(b->b.noshadow(item)); }
int glome_sb_onlyshadow(GlomeBuilder* b, int item) { GUARD(b->b.onlyshadow(item)); }
int glome_sb_bound_object(GlomeBuilder* b, int sa, int sb) { GUARD(b->b.bound_object(sa, sb)); }
int glome_sb_innerbound(GlomeBuilder* b, int sa, int sb) { GUARD(b->b.innerbound(sa, sb)); }
int glome_sb_transform(GlomeBuilder* b, int item, int nx, const double* xfms) {
    try {
        std::vector<Xfm> xs;
        for (int i = 0; i < nx; i++) xs.push_back(X(xfms + 24 * i));
        return b->b.transform(item, xs);
    } GUARD_CATCH
}
int glome_xfm_translate(const double v[3], double out[24]) {
    try { Xfm x = translate(V(v)); memcpy(out, x.m, sizeof(x.m)); return GLOME_OK; }
    GUARD_CATCH
}
int glome_xfm_scale(const double v[3], double out[24]) {
    try { Xfm x = scale(V(v)); memcpy(out, x.m, sizeof(x.m)); return GLOME_OK; }
    GUARD_CATCH
}
int glome_xfm_rotate(const double axis[3], double angle, double out[24]) {
    try { Xfm x = rotate(V(axis), angle); memcpy(out, x.m, sizeof(x.m)); return GLOME_OK; }
    GUARD_CATCH
}
int glome_xfm_compose(int n, const double* xfms, double out[24]) {
    try {
        std::vector<Xfm> xs;
        for (int i = 0; i < n; i++) xs.push_back(X(xfms + 24 * i));
        Xfm x = compose(xs);
        memcpy(out, x.m, sizeof(x.m));
        return GLOME_OK;
    } GUARD_CATCH
}
int glome_sb_flatten_transform_bih(GlomeBuilder* b, int item) {
    try {
        std::vector<int32_t> leaves = b->b.tolist(b->b.list_raw(b->b.flatten_transform(item)));
        return b->b.bih(leaves);
    } GUARD_CATCH
}
int glome_sb_bound(GlomeBuilder* b, int item, double out[6]) {
    try {
        glm::Bbox bb = b->b.bound(item);
        out[0] = bb.p1.x; out[1] = bb.p1.y; out[2] = bb.p1.z; out[3] = bb.p2.x; out[4] = bb.p2.y; out[5] = bb.p2.z;
        return GLOME_OK;
    } GUARD_CATCH
}

int glome_sb_mat_surface(GlomeBuilder* b, const double rgb[3], double alpha, double amb, double kd, double ks, double shine) {
    GUARD(b->b.mat_surface(rgb[0], rgb[1], rgb[2], alpha, amb, kd, ks, shine));
}
int glome_sb_mat_reflect(GlomeBuilder* b, double refl) { GUARD(b->b.mat_reflect(refl)); }
int glome_sb_mat_refract(GlomeBuilder* b, double refl, double refr, double ior) { GUARD(b->b.mat_refract(refl, refr, ior)); }
int glome_sb_mat_warp(GlomeBuilder* b, int frame_item, int scene_item, int lightset, const double xfm[24]) {
    GUARD(b->b.mat_warp(frame_item, scene_item, lightset, X(xfm)));
}
int glome_sb_mat_additive(GlomeBuilder* b, int n, const int32_t* mats) { GUARD(b->b.mat_additive(IV(n, mats))); }
int glome_sb_mat_blend(GlomeBuilder* b, int ma, int mb, double weight) { GUARD(b->b.mat_blend(ma, mb, weight)); }
int glome_sb_tex_uniform(GlomeBuilder* b, int mat) { GUARD(b->b.tex_uniform(mat)); }
int glome_sb_tex_stripe_blend(GlomeBuilder* b, int ma, int mb, const double axis[3]) { GUARD(b->b.tex_stripe_blend(ma, mb, V(axis))); }
int glome_sb_tex_perlin_blend(GlomeBuilder* b, int ma, int mb, double scale) { GUARD(b->b.tex_perlin_blend(ma, mb, scale)); }
int glome_sb_light(GlomeBuilder* b, const double pos[3], const double color[3]) {
    GUARD(b->b.light(V(pos), color[0], color[1], color[2]));
}
int glome_sb_lightset(GlomeBuilder* b, int n, const int32_t* lights) { GUARD(b->b.lightset(IV(n, lights))); }
int glome_sb_mat_warp_set_scene(GlomeBuilder* b, int mat, int scene_item) {
    if (mat < 0 || mat >= (int)b->b.materials.size() || b->b.materials[mat].kind != GLOME_MAT_WARP) {
        glome_set_error("not a Warp material");
        return GLOME_EINVAL;
    }
    b->b.materials[mat].b = scene_item;
    return GLOME_OK;
}

int glome_camera(const double pos[3], const double at[3], const double up[3], double angle_deg, GlomeCamera* out) {
    make_camera(V(pos), V(at), V(up), angle_deg, out);
    return GLOME_OK;
}

int glome_sb_flatten(GlomeBuilder* b, int root, GlomeFlatScene* out) {
    try { b->b.flatten(root, out); return GLOME_OK; }
    GUARD_CATCH
}

int glome_sb_config_scene(GlomeBuilder* b, int config, int64_t n, uint64_t seed, GlomeCamera* cam, int* recurs_out) {
    GUARD(config_scene(b->b, config, n, seed, cam, recurs_out));
}

static int bih_build_any(int64_t n, const double* bboxes, int device, double* timings_ms, GlomeBihNode** nodes_out,
                         int32_t* n_nodes_out, int32_t** leaves_out, int32_t* n_leaves_out, int32_t** item_order_out,
                         int32_t* root_ref_out, double bb_out[6]);
int glome_sb_load_nff(GlomeBuilder* b, const char* text, int64_t len, GlomeCamera* cam, double bg[3], int64_t* consumed) {
    if (!b || !text) return GLOME_EINVAL;
    GUARD(nff_load(b->b, text, len, cam, bg, consumed));
}
int glome_bih_build(int64_t n, const double* bboxes, GlomeBihNode** nodes_out, int32_t* n_nodes_out, int32_t** leaves_out,
                    int32_t* n_leaves_out, int32_t** item_order_out, int32_t* root_ref_out, double bb_out[6]) {
    return bih_build_any(n, bboxes, -1, nullptr, nodes_out, n_nodes_out, leaves_out, n_leaves_out, item_order_out, root_ref_out, bb_out);
}
int glome_bih_build_gpu(int64_t n, const double* bboxes, int device, GlomeBihNode** nodes_out, int32_t* n_nodes_out,
                        int32_t** leaves_out, int32_t* n_leaves_out, int32_t** item_order_out, int32_t* root_ref_out,
                        double bb_out[6], double timings_ms[3]) {
    if (device < 0) { glome_set_error("glome_bih_build_gpu needs a device index"); return GLOME_ENODEV; }
    return bih_build_any(n, bboxes, device, timings_ms, nodes_out, n_nodes_out, leaves_out, n_leaves_out, item_order_out, root_ref_out, bb_out);
}
static int bih_build_any(int64_t n, const double* bboxes, int device, double* timings_ms, GlomeBihNode** nodes_out,
                         int32_t* n_nodes_out, int32_t** leaves_out, int32_t* n_leaves_out, int32_t** item_order_out,
                         int32_t* root_ref_out, double bb_out[6]) {
    try {
        BihTree t;
        if (device >= 0) bih_build_gpu(n, bboxes, device, t, timings_ms);
        else bih_build(n, bboxes, t);
        *n_nodes_out = (int32_t)t.nodes.size();
        *n_leaves_out = (int32_t)(t.leaves.size() / 2);
        *nodes_out = (GlomeBihNode*)malloc(sizeof(GlomeBihNode) * (t.nodes.size() + 1));
        memcpy(*nodes_out, t.nodes.data(), sizeof(GlomeBihNode) * t.nodes.size());
        *leaves_out = (int32_t*)malloc(sizeof(int32_t) * (t.leaves.size() + 2));
        memcpy(*leaves_out, t.leaves.data(), sizeof(int32_t) * t.leaves.size());
        *item_order_out = (int32_t*)malloc(sizeof(int32_t) * (t.order.size() + 1));
        memcpy(*item_order_out, t.order.data(), sizeof(int32_t) * t.order.size());
        *root_ref_out = t.root;
        bb_out[0] = t.bb.p1.x; bb_out[1] = t.bb.p1.y; bb_out[2] = t.bb.p1.z;
        bb_out[3] = t.bb.p2.x; bb_out[4] = t.bb.p2.y; bb_out[5] = t.bb.p2.z;
        return GLOME_OK;
    } GUARD_CATCH
}
static int mesh_build_any(int64_t nverts, const double* verts, int64_t ntris, const int32_t* tris, int device, double* timings_ms,
                          GlomeBvhNode** nodes_out, int32_t* n_nodes_out, int32_t** leafpool_out, int32_t* n_leafpool_out,
                          int32_t** leafoff_out, int32_t* n_leaves_out, int32_t* root_ref_out, double bb_out[6]);
int glome_mesh_build(int64_t nverts, const double* verts, int64_t ntris, const int32_t* tris, GlomeBvhNode** nodes_out,
                     int32_t* n_nodes_out, int32_t** leafpool_out, int32_t* n_leafpool_out, int32_t** leafoff_out,
                     int32_t* n_leaves_out, int32_t* root_ref_out, double bb_out[6]) {
    return mesh_build_any(nverts, verts, ntris, tris, -1, nullptr, nodes_out, n_nodes_out, leafpool_out, n_leafpool_out, leafoff_out,
                          n_leaves_out, root_ref_out, bb_out);
}
int glome_mesh_build_gpu(int64_t nverts, const double* verts, int64_t ntris, const int32_t* tris, int device,
                         GlomeBvhNode** nodes_out, int32_t* n_nodes_out, int32_t** leafpool_out, int32_t* n_leafpool_out,
                         int32_t** leafoff_out, int32_t* n_leaves_out, int32_t* root_ref_out, double bb_out[6],
                         double timings_ms[3]) {
    if (device < 0) { glome_set_error("glome_mesh_build_gpu needs a device index"); return GLOME_ENODEV; }
    return mesh_build_any(nverts, verts, ntris, tris, device, timings_ms, nodes_out, n_nodes_out, leafpool_out, n_leafpool_out,
                          leafoff_out, n_leaves_out, root_ref_out, bb_out);
}
static int mesh_build_any(int64_t nverts, const double* verts, int64_t ntris, const int32_t* tris, int device, double* timings_ms,
                          GlomeBvhNode** nodes_out, int32_t* n_nodes_out, int32_t** leafpool_out, int32_t* n_leafpool_out,
                          int32_t** leafoff_out, int32_t* n_leaves_out, int32_t* root_ref_out, double bb_out[6]) {
    try {
        MeshTree t;
        if (device >= 0) mesh_build_gpu(nverts, verts, ntris, tris, device, t, timings_ms);
        else mesh_build(nverts, verts, ntris, tris, t);
        *n_nodes_out = (int32_t)t.nodes.size();
        *nodes_out = (GlomeBvhNode*)malloc(sizeof(GlomeBvhNode) * (t.nodes.size() + 1));
        memcpy(*nodes_out, t.nodes.data(), sizeof(GlomeBvhNode) * t.nodes.size());
        *n_leafpool_out = (int32_t)t.leafpool.size();
        *leafpool_out = (int32_t*)malloc(sizeof(int32_t) * (t.leafpool.size() + 1));
        memcpy(*leafpool_out, t.leafpool.data(), sizeof(int32_t) * t.leafpool.size());
        *n_leaves_out = (int32_t)t.leafoff.size();
        *leafoff_out = (int32_t*)malloc(sizeof(int32_t) * (t.leafoff.size() + 1));
        memcpy(*leafoff_out, t.leafoff.data(), sizeof(int32_t) * t.leafoff.size());
        *root_ref_out = t.root;
        bb_out[0] = t.bb.p1.x; bb_out[1] = t.bb.p1.y; bb_out[2] = t.bb.p1.z;
        bb_out[3] = t.bb.p2.x; bb_out[4] = t.bb.p2.y; bb_out[5] = t.bb.p2.z;
        return GLOME_OK;
    } GUARD_CATCH
}
void glome_free(void* p) { free(p); }

}  // extern "C"
