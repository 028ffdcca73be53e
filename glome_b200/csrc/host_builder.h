// host_builder.h -- host-side mirror of the GlomeTrace scene-construction API (the part of the
// reference that "stays in Haskell"): constructors with the reference's names and semantics, the
// `bih` / `mesh` tree builders, and the flattener that emits the FlatScene the device consumes.
//
// This is what a GlomeTrace.CUDA `flatten` method produces on the Haskell side;
// here it lets C++/Python callers (tests, bench.py) build the same scenes without GHC.
#pragma once
#include <cstdint>
#include <string>
#include <vector>

#include "../../include/glome_cuda.h"
#include "glome_math.h"

namespace glome_host {

using glm::Bbox;
using glm::Flt;
using glm::Vec;

struct Xfm { Flt m[24]; };  // forward 12, inverse 12 (Vec.hs:414)

// Vec.hs transformation constructors (host only: they use sin/cos)
Xfm ident_xfm();
Xfm xfm_mult(const Xfm& a, const Xfm& b);                 // Vec.hs:447
bool check_xfm(const Xfm& x, std::string* err);           // Vec.hs:466
Xfm compose(const std::vector<Xfm>& xs);                  // Vec.hs:461 (throws BuildError on corrupt)
Xfm translate(const Vec& v);                              // Vec.hs:564
Xfm scale(const Vec& v);                                  // Vec.hs:571
Xfm rotate(const Vec& axis, Flt angle);                   // Vec.hs:577
Xfm xyz_to_uvw(const Vec& u, const Vec& v, const Vec& w); // Vec.hs:602
void orth(const Vec& v1, Vec& v2, Vec& v3);               // Vec.hs:366
Flt deg(Flt x);                                           // Vec.hs:17 (truncated pi)
bool about_equal(Flt a, Flt b);                           // Vec.hs:96

struct BuildError {
    std::string msg;
    explicit BuildError(const std::string& m) : msg(m) {}
};

struct BihTree {
    Bbox bb;
    std::vector<GlomeBihNode> nodes;  // pre-order; child refs local (>=0 node, <0 ~leaf index)
    std::vector<int32_t> leaves;      // {first, count} into order
    std::vector<int32_t> order;       // leaf-ordered permutation of 0..n-1
    int32_t root;
};
// Bih.hs:211-324.  Multi-threaded; output numbering is deterministic (pre-order).
void bih_build(int64_t n, const double* bboxes, BihTree& out);

struct MeshTree {
    Bbox bb;
    std::vector<GlomeBvhNode> nodes;
    std::vector<int32_t> leafpool;  // {count, tri...}
    std::vector<int32_t> leafoff;
    int32_t root;
};
// Mesh.hs:50-134
void mesh_build(int64_t nverts, const double* verts, int64_t ntris, const int32_t* tris, MeshTree& out);

struct MeshData {
    std::vector<double> verts, norms;
    std::vector<int32_t> tris, texs, tags;
    MeshTree tree;
};

struct Item {
    int type;
    int ia, ib;              // tex id / tag id / useatex / table index (bih, mesh)
    int nd;
    double d[24];
    std::vector<int32_t> kids;
};

class Builder {
  public:
    std::vector<Item> items;
    std::vector<BihTree> bihs;
    int build_device = -1;           // >= 0: `bih` builds its tree on that GPU (glome_build.cu), same result
    double build_ms[4] = {0, 0, 0, 0}; // last bih: H2D, device build, D2H (GPU) or 0,0,0 + [3] host build wall ms
    std::vector<MeshData> meshes;
    std::vector<GlomeMaterial> materials;  // WARP a/b hold ITEM ids until flatten
    std::vector<GlomeTexture> textures;
    std::vector<GlomeLight> lights;
    std::vector<int32_t> lightsets;
    std::vector<int32_t> mat_lists;        // ADDITIVE payloads

    // ---- constructors (names follow the reference) ----
    int void_();
    int sphere(const Vec& c, Flt r);
    int triangle(const Vec& a, const Vec& b, const Vec& c);
    int trianglenorm(const Vec& a, const Vec& b, const Vec& c, const Vec& na, const Vec& nb, const Vec& nc);
    int box(const Vec& a, const Vec& b);
    int plane(const Vec& orig, const Vec& norm);
    int plane_offset(const Vec& n, Flt off);
    int disc(const Vec& pos, const Vec& norm, Flt r);
    int cylinder_z(Flt r, Flt h1, Flt h2);
    int cone_z(Flt r, Flt h1, Flt h2, Flt height);
    int cylinder(const Vec& p1, const Vec& p2, Flt r);
    int cone(const Vec& p1, Flt r1, const Vec& p2, Flt r2);
    int group(const std::vector<int32_t>& xs);
    int list_raw(const std::vector<int32_t>& xs);  // SolidItem [s] without group's flattening
    int bih(const std::vector<int32_t>& xs);
    int mesh(int64_t nverts, const double* verts, int64_t nnorms, const double* norms, int64_t ntris,
             const int32_t* tris, int ntexs, const int32_t* texs, int ntags, const int32_t* tags);
    // trees the caller (the Haskell constructors) already built, as pre-order streams (host_builder.cpp)
    int bih_prebuilt(const std::vector<int32_t>& xs, int64_t n_nodes, const int32_t* kinds, const double* splits, const double bb[6]);
    int mesh_prebuilt(int64_t nverts, const double* verts, int64_t nnorms, const double* norms, int64_t ntris,
                      const int32_t* tris, int ntexs, const int32_t* texs, int ntags, const int32_t* tags, int64_t n_nodes,
                      const int32_t* kinds, const double* boxes, int64_t n_leaf_tris, const int32_t* leaf_tris, const double bb[6]);
    int difference(int sa, int sb);
    int difference_ex(int sa, int sb, bool useatex);        // Difference a b Bool as stored (Csg.hs:14, 26-30)
    int disc_raw(const Vec& pos, const Vec& norm, Flt rsqr); // Disc pos norm (r*r) as stored (Cone.hs:21)
    int intersection(const std::vector<int32_t>& xs);
    int tex(int s, int texture);
    int tag(int s, int tagid);
    int noshadow(int s);
    int onlyshadow(int s);
    int bound_object(int sa, int sb);
    int innerbound(int sa, int sb);
    int instance_raw(int s, const Xfm& x);
    int transform(int s, const std::vector<Xfm>& xs);       // Solid.hs:184
    int transform_leaf(int s, const std::vector<Xfm>& xs);  // Solid.hs:187
    std::vector<int32_t> tolist(int s);                     // Solid.hs:177
    std::vector<int32_t> flatten_transform(int s);          // Solid.hs:192
    Bbox bound(int s);                                      // Solid.hs:171

    int mat_surface(Flt r, Flt g, Flt b, Flt alpha, Flt amb, Flt kd, Flt ks, Flt shine);
    int mat_reflect(Flt refl);
    int mat_refract(Flt refl, Flt refr, Flt ior);
    int mat_warp(int frame, int scene, int lightset, const Xfm& x);
    int mat_additive(const std::vector<int32_t>& ms);
    int mat_blend(int ma, int mb, Flt w);
    int tex_uniform(int mat);
    int tex_stripe_blend(int ma, int mb, const Vec& axis);
    int tex_perlin_blend(int ma, int mb, Flt scale);
    int light(const Vec& pos, Flt r, Flt g, Flt b);
    int lightset(const std::vector<int32_t>& ls);

    // ---- flatten ----
    void flatten(int root, GlomeFlatScene* out);

  private:
    int add(const Item& it);
    int check(int s) const;
    // flatten state
    std::vector<GlomeNode> f_nodes;
    std::vector<GlomeBihNode> f_bih;
    std::vector<GlomeBvhNode> f_bvh;
    std::vector<int32_t> f_ipool;
    std::vector<double> f_dpool;
    std::vector<GlomeMaterial> f_mats;
    std::vector<int32_t> f_lightsets;
    std::vector<int32_t> f_bih_memo, f_mesh_memo;  // table index -> ipool offset of memo record, -1
    std::vector<Xfm> warp_xfms;
    int f_maxdepth;
    int alloc_nodes(int n);
    int alloc_d(int n, int align);
    void flatten_into(int item, int slot, int depth);
    bool flat_class(int item, int level) const;
};

// BASELINE.json config scenes (scenes.cpp)
int config_scene(Builder& b, int config, int64_t n, uint64_t seed, GlomeCamera* cam, int* recurs);
void make_camera(const Vec& pos, const Vec& at, const Vec& up, Flt angle, GlomeCamera* out);  // Scene.hs:48
// NFF / SPD scene text -> items, lights and camera (Spd.hs:1-261; nff.cpp).  Returns the scene's root item
// (bih of the fill groups); consumed = bytes parsed before the reader stopped.
int nff_load(Builder& b, const char* text, int64_t len, GlomeCamera* cam, double bg[3], int64_t* consumed);

}  // namespace glome_host
