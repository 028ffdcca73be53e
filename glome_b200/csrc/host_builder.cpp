// host_builder.cpp -- see host_builder.h.  Host code only (no CUDA); compiled -ffp-contract=off.
#include "host_builder.h"
#include "glome_build.h"
#include <chrono>

#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstring>
#include <future>
#include <thread>

namespace glome_host {

using namespace glm;

static const Flt kDelta = GLM_DELTA;
static const Flt kInf = GLM_INFINITY;

// ---------------------------------------------------------------------------------------------
// Vec.hs transformation constructors
// ---------------------------------------------------------------------------------------------
Flt deg(Flt x) { return (x * 3.1415926535897) / 180; }  // Vec.hs:17-18

bool about_equal(Flt a, Flt b) {  // Vec.hs:96-102
    if (a > 1) return fabs_(1 - (a / b)) < (kDelta * 10);
    return fabs_(a - b) < (kDelta * 10);
}

Xfm ident_xfm() {
    Xfm x;
    static const Flt id[12] = {1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1, 0};
    memcpy(x.m, id, sizeof(id));
    memcpy(x.m + 12, id, sizeof(id));
    return x;
}

static void mat_mult(const Flt* a, const Flt* b, Flt* o) {  // Vec.hs:426-443
    Flt r[12];
    for (int i = 0; i < 3; i++) {
        const Flt* ar = a + 4 * i;
        r[4 * i + 0] = ar[0] * b[0] + ar[1] * b[4] + ar[2] * b[8];
        r[4 * i + 1] = ar[0] * b[1] + ar[1] * b[5] + ar[2] * b[9];
        r[4 * i + 2] = ar[0] * b[2] + ar[1] * b[6] + ar[2] * b[10];
        r[4 * i + 3] = ar[0] * b[3] + ar[1] * b[7] + ar[2] * b[11] + ar[3];
    }
    memcpy(o, r, sizeof(r));
}

Xfm xfm_mult(const Xfm& a, const Xfm& b) {  // Vec.hs:447-449
    Xfm o;
    mat_mult(a.m, b.m, o.m);
    mat_mult(b.m + 12, a.m + 12, o.m + 12);
    return o;
}

bool check_xfm(const Xfm& x, std::string* err) {  // Vec.hs:466-477
    Flt p[12];
    mat_mult(x.m, x.m + 12, p);
    static const Flt id[12] = {1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1, 0};
    for (int i = 0; i < 12; i++)
        if (!about_equal(p[i], id[i])) {
            if (err) *err = "corrupt matrix";
            return false;
        }
    return true;
}
static Xfm checked(const Xfm& x) {
    std::string e;
    if (!check_xfm(x, &e)) throw BuildError(e);
    return x;
}

Xfm compose(const std::vector<Xfm>& xs) {  // Vec.hs:461-462: foldr xfm_mult ident_xfm (reverse xfms)
    Xfm acc = ident_xfm();
    for (size_t i = 0; i < xs.size(); i++) acc = xfm_mult(xs[i], acc);
    return checked(acc);
}

static Xfm mk(const Flt* f, const Flt* i) {
    Xfm x;
    memcpy(x.m, f, 12 * sizeof(Flt));
    memcpy(x.m + 12, i, 12 * sizeof(Flt));
    return x;
}

Xfm translate(const Vec& v) {  // Vec.hs:564-567
    Flt f[12] = {1, 0, 0, v.x, 0, 1, 0, v.y, 0, 0, 1, v.z};
    Flt i[12] = {1, 0, 0, -v.x, 0, 1, 0, -v.y, 0, 0, 1, -v.z};
    return checked(mk(f, i));
}
Xfm scale(const Vec& v) {  // Vec.hs:571-574
    Flt f[12] = {v.x, 0, 0, 0, 0, v.y, 0, 0, 0, 0, v.z, 0};
    Flt i[12] = {1 / v.x, 0, 0, 0, 0, 1 / v.y, 0, 0, 0, 0, 1 / v.z, 0};
    return checked(mk(f, i));
}
Xfm rotate(const Vec& v, Flt angle) {  // Vec.hs:577-598
    if (!about_equal(vlen(v), 1)) throw BuildError("please use a normalized vector for rotation");
    Flt x = v.x, y = v.y, z = v.z;
    Flt s = sin(angle), c = cos(angle);
    Flt m00 = ((x * x) + ((1 - (x * x)) * c));
    Flt m01 = (((x * y) * (1 - c)) - (z * s));
    Flt m02 = ((x * z * (1 - c)) + (y * s));
    Flt m10 = (((x * y) * (1 - c)) + (z * s));
    Flt m11 = ((y * y) + ((1 - (y * y)) * c));
    Flt m12 = ((y * z * (1 - c)) - (x * s));
    Flt m20 = ((x * z * (1 - c)) - (y * s));
    Flt m21 = ((y * z * (1 - c)) + (x * s));
    Flt m22 = ((z * z) + ((1 - (z * z)) * c));
    Flt f[12] = {m00, m01, m02, 0, m10, m11, m12, 0, m20, m21, m22, 0};
    Flt i[12] = {m00, m10, m20, 0, m01, m11, m21, 0, m02, m12, m22, 0};
    return checked(mk(f, i));
}
Xfm xyz_to_uvw(const Vec& u, const Vec& v, const Vec& w) {  // Vec.hs:602-622
    if (!about_equal(vdot(u, u), 1)) throw BuildError("unnormalized u");
    if (!about_equal(vdot(v, v), 1)) throw BuildError("unnormalized v");
    if (!about_equal(vdot(w, w), 1)) throw BuildError("unnormalized w");
    if (!(about_equal(vdot(u, v), 0) && about_equal(vdot(u, w), 0) && about_equal(vdot(v, w), 0)))
        throw BuildError("vectors aren't orthogonal");
    Flt f[12] = {u.x, v.x, w.x, 0, u.y, v.y, w.y, 0, u.z, v.z, w.z, 0};
    Flt i[12] = {u.x, u.y, u.z, 0, v.x, v.y, v.z, 0, w.x, w.y, w.z, 0};
    return checked(mk(f, i));
}
void orth(const Vec& v1, Vec& v2, Vec& v3) {  // Vec.hs:366-378
    if (!about_equal(vdot(v1, v1), 1)) throw BuildError("orth: unnormalized vector");
    Vec x = vec(1, 0, 0), y = vec(0, 1, 0);
    Flt dvx = vdot(v1, x);
    if (dvx < 0.8 && dvx > (-0.8)) v2 = vnorm(vcross(v1, x));
    else v2 = vnorm(vcross(v1, y));
    v3 = vcross(v1, v2);
}

void make_camera(const Vec& pos, const Vec& at, const Vec& up, Flt angle, GlomeCamera* out) {  // Scene.hs:48-57
    Vec fwd = vnorm(vsub(at, pos));
    Vec right = vnorm(vcross(up, fwd));
    Vec up_ = vnorm(vcross(fwd, right));
    Flt cam_scale = tan((M_PI / 180) * (angle / 2));
    Vec u = vscale(up_, cam_scale), r = vscale(right, cam_scale);
    out->pos[0] = pos.x; out->pos[1] = pos.y; out->pos[2] = pos.z;
    out->fwd[0] = fwd.x; out->fwd[1] = fwd.y; out->fwd[2] = fwd.z;
    out->up[0] = u.x; out->up[1] = u.y; out->up[2] = u.z;
    out->right[0] = r.x; out->right[1] = r.y; out->right[2] = r.z;
}

// ---------------------------------------------------------------------------------------------
// bih (Bih.hs:211-324): index-based builder.  One pass over a segment evaluates the four
// candidate partitions (x, y, z, big/small) without materialising them; the winner is applied
// as an in-place stable partition, so the final index array is the leaf-ordered permutation.
// ---------------------------------------------------------------------------------------------
namespace {

struct TNode {
    bool leaf;
    int axis;
    Flt lsplit, rsplit;
    int64_t lo, hi;
    TNode *l, *r;
};

struct BihBuildCtx {
    const double* bb;         // n*6
    std::vector<Vec> mid;     // bbmid of each object
    std::vector<Flt> sa;      // bbsa' of each object
    std::vector<int32_t> idx, scratch;
};

static inline Flt bbsa_p(const Bbox& b) { return hmax(0, bbsa(b)); }  // Bih.hs:208
static inline Bbox set_p2(Bbox b, int ax, Flt f) { if (ax == 0) b.p2.x = f; else if (ax == 1) b.p2.y = f; else b.p2.z = f; return b; }
static inline Bbox set_p1(Bbox b, int ax, Flt f) { if (ax == 0) b.p1.x = f; else if (ax == 1) b.p1.y = f; else b.p1.z = f; return b; }

static TNode* bih_rec(BihBuildCtx& c, int64_t lo, int64_t hi, const Bbox& bb, const Vec& mid, int depth) {
    TNode* t = new TNode();
    t->lo = lo; t->hi = hi; t->l = t->r = nullptr; t->leaf = true; t->axis = 0; t->lsplit = t->rsplit = 0;
    int64_t n = hi - lo;
    if (n <= 3) return t;  // Bih.hs:214
    if (depth > 512) throw BuildError("bih: recursion too deep (degenerate input)");
    Flt sa = bbsa_p(bb);
    Flt thresh = sa * 0.4;
    int64_t lc[4] = {0, 0, 0, 0};
    Flt lmax[4] = {-kInf, -kInf, -kInf, -kInf}, rmin[4] = {kInf, kInf, kInf, kInf};
    for (int64_t p = lo; p < hi; p++) {
        int32_t i = c.idx[p];
        const double* b = c.bb + 6 * (int64_t)i;
        const Vec& m = c.mid[i];
        if (m.x < mid.x) { lc[0]++; lmax[0] = fmax_(lmax[0], b[3]); } else rmin[0] = fmin_(rmin[0], b[0]);
        if (m.y < mid.y) { lc[1]++; lmax[1] = fmax_(lmax[1], b[4]); } else rmin[1] = fmin_(rmin[1], b[1]);
        if (m.z < mid.z) { lc[2]++; lmax[2] = fmax_(lmax[2], b[5]); } else rmin[2] = fmin_(rmin[2], b[2]);
        if (c.sa[i] > thresh) { lc[3]++; lmax[3] = fmax_(lmax[3], b[3]); } else rmin[3] = fmin_(rmin[3], b[0]);  // x planes (Bih.hs:231-232)
    }
    Bbox lbb[4], rbb[4];
    Flt cost[4];
    for (int k = 0; k < 4; k++) {
        int ax = (k == 3) ? 0 : k;
        lbb[k] = set_p2(bb, ax, lmax[k]);
        rbb[k] = set_p1(bb, ax, rmin[k]);
        Flt fac = (k == 3) ? 1.2 : 1.1;  // Bih.hs:252-255
        cost[k] = ((bbsa_p(lbb[k]) * (Flt)lc[k]) + (bbsa_p(rbb[k]) * (Flt)(n - lc[k]))) * fac;
    }
    Flt costorig = sa * (Flt)n;
    if (costorig < cost[0] && costorig < cost[1] && costorig < cost[2] && costorig < cost[3]) return t;  // Bih.hs:276
    int k;
    if (cost[0] < cost[1] && cost[0] < cost[2] && cost[0] < cost[3]) k = 0;
    else if (cost[1] < cost[2] && cost[1] < cost[3]) k = 1;
    else if (cost[1] < cost[3]) k = 2;  // sic (Bih.hs:283 tests costy)
    else k = 3;
    // stable partition of idx[lo,hi) by the winning predicate
    int64_t a = lo, s = 0;
    for (int64_t p = lo; p < hi; p++) {
        int32_t i = c.idx[p];
        bool left;
        if (k == 0) left = c.mid[i].x < mid.x;
        else if (k == 1) left = c.mid[i].y < mid.y;
        else if (k == 2) left = c.mid[i].z < mid.z;
        else left = c.sa[i] > thresh;
        if (left) c.idx[a++] = i;
        else c.scratch[lo + s++] = i;
    }
    memcpy(&c.idx[a], &c.scratch[lo], sizeof(int32_t) * (size_t)s);
    t->leaf = false;
    t->axis = (k == 3) ? 0 : k;
    t->lsplit = lmax[k] + kDelta;
    t->rsplit = rmin[k] - kDelta;
    Bbox lb = lbb[k], rb = rbb[k];
    int64_t midp = a;
    if (depth < 6 && n > 20000) {  // task-parallel like spawnP at depth <= 5 (Bih.hs:266-274)
        auto fut = std::async(std::launch::async, [&c, lo, midp, lb, depth]() { return bih_rec(c, lo, midp, lb, bbmid(lb), depth + 1); });
        t->r = bih_rec(c, midp, hi, rb, bbmid(rb), depth + 1);
        t->l = fut.get();
    } else {
        t->l = bih_rec(c, lo, midp, lb, bbmid(lb), depth + 1);
        t->r = bih_rec(c, midp, hi, rb, bbmid(rb), depth + 1);
    }
    return t;
}

static int32_t bih_number(TNode* t, BihTree& out) {
    if (t->leaf) {
        int32_t li = (int32_t)(out.leaves.size() / 2);
        out.leaves.push_back((int32_t)t->lo);
        out.leaves.push_back((int32_t)(t->hi - t->lo));
        delete t;
        return ~li;
    }
    int32_t me = (int32_t)out.nodes.size();
    GlomeBihNode nd;
    memset(&nd, 0, sizeof(nd));
    nd.lsplit = t->lsplit; nd.rsplit = t->rsplit; nd.axis = t->axis;
    out.nodes.push_back(nd);
    int32_t l = bih_number(t->l, out);
    int32_t r = bih_number(t->r, out);
    out.nodes[me].left = l;
    out.nodes[me].right = r;
    delete t;
    return me;
}

}  // namespace

void bih_build(int64_t n, const double* bboxes, BihTree& out) {
    BihBuildCtx c;
    c.bb = bboxes;
    c.mid.resize((size_t)n); c.sa.resize((size_t)n); c.idx.resize((size_t)n); c.scratch.resize((size_t)n);
    Bbox bb = empty_bbox();
    for (int64_t i = 0; i < n; i++) {
        const double* b = bboxes + 6 * i;
        Bbox ob = mkbb(vec(b[0], b[1], b[2]), vec(b[3], b[4], b[5]));
        bb = bbjoin(bb, ob);  // foldl' bbjoin empty_bbox (Bih.hs:315)
        c.mid[i] = bbmid(ob);
        c.sa[i] = bbsa_p(ob);
        c.idx[i] = (int32_t)i;
    }
    if (n > 0 && (bb.p1.x == -kInf || bb.p1.y == -kInf || bb.p1.z == -kInf || bb.p2.x == kInf || bb.p2.y == kInf ||
                  bb.p2.z == kInf))
        throw BuildError("bih: infinite bounding box");  // Bih.hs:319-322
    TNode* root = bih_rec(c, 0, n, bb, bbmid(bb), 0);
    out.bb = bb;
    out.nodes.clear(); out.leaves.clear();
    out.root = bih_number(root, out);
    out.order.swap(c.idx);
}

// ---------------------------------------------------------------------------------------------
// mesh (Mesh.hs:50-134)
// ---------------------------------------------------------------------------------------------
namespace {

struct MNode {
    bool leaf;
    Bbox lbb, rbb;
    int64_t lo, hi;
    MNode *l, *r;
};
struct MeshBuildCtx {
    std::vector<Bbox> tbb;
    std::vector<Vec> mid;
    std::vector<Flt> sa;
    std::vector<int32_t> idx, scratch;
};

static MNode* mesh_rec(MeshBuildCtx& c, int64_t lo, int64_t hi, const Bbox& bb, int depth) {
    MNode* t = new MNode();
    t->leaf = true; t->lo = lo; t->hi = hi; t->l = t->r = nullptr;
    int64_t n = hi - lo;
    if (n < 3) return t;  // Mesh.hs:72
    if (depth > 512) throw BuildError("mesh: recursion too deep (degenerate input)");
    Vec mid = bbmid(bb);
    Flt sa = bbsa(bb);
    Flt thresh = sa * 0.4;
    int64_t lc[4] = {0, 0, 0, 0};
    Bbox lb[4], rb[4];
    for (int k = 0; k < 4; k++) lb[k] = rb[k] = empty_bbox();
    for (int64_t p = lo; p < hi; p++) {
        int32_t i = c.idx[p];
        const Bbox& tb = c.tbb[i];
        const Vec& m = c.mid[i];
        if (m.x < mid.x) { lc[0]++; lb[0] = bbjoin(lb[0], tb); } else rb[0] = bbjoin(rb[0], tb);
        if (m.y < mid.y) { lc[1]++; lb[1] = bbjoin(lb[1], tb); } else rb[1] = bbjoin(rb[1], tb);
        if (m.z < mid.z) { lc[2]++; lb[2] = bbjoin(lb[2], tb); } else rb[2] = bbjoin(rb[2], tb);
        if (c.sa[i] > thresh) { lc[3]++; lb[3] = bbjoin(lb[3], tb); } else rb[3] = bbjoin(rb[3], tb);
    }
    Flt cost[4];
    for (int k = 0; k < 4; k++) cost[k] = ((bbsa(lb[k]) * (Flt)lc[k]) + (bbsa(rb[k]) * (Flt)(n - lc[k]))) * 1.1;  // Mesh.hs:98-101
    Flt lcost = bbsa(bb) * (Flt)n;
    if (lcost < cost[0] && lcost < cost[1] && lcost < cost[2] && lcost < cost[3]) return t;  // Mesh.hs:104
    int k;
    if (cost[0] < cost[1] && cost[0] < cost[2] && cost[0] < cost[3]) k = 0;
    else if (cost[1] < cost[2] && cost[1] < cost[3]) k = 1;
    else if (cost[2] < cost[3]) k = 2;
    else k = 3;
    int64_t a = lo, s = 0;
    for (int64_t p = lo; p < hi; p++) {
        int32_t i = c.idx[p];
        bool left;
        if (k == 0) left = c.mid[i].x < mid.x;
        else if (k == 1) left = c.mid[i].y < mid.y;
        else if (k == 2) left = c.mid[i].z < mid.z;
        else left = c.sa[i] > thresh;
        if (left) c.idx[a++] = i;
        else c.scratch[lo + s++] = i;
    }
    memcpy(&c.idx[a], &c.scratch[lo], sizeof(int32_t) * (size_t)s);
    t->leaf = false;
    t->lbb = lb[k]; t->rbb = rb[k];
    int64_t midp = a;
    Bbox L = lb[k], R = rb[k];
    if (depth < 6 && (midp - lo) > 1000 && (hi - midp) > 1000) {  // Mesh.hs:57-65
        auto fut = std::async(std::launch::async, [&c, lo, midp, L, depth]() { return mesh_rec(c, lo, midp, L, depth + 1); });
        t->r = mesh_rec(c, midp, hi, R, depth + 1);
        t->l = fut.get();
    } else {
        t->l = mesh_rec(c, lo, midp, L, depth + 1);
        t->r = mesh_rec(c, midp, hi, R, depth + 1);
    }
    return t;
}

static int32_t mesh_number(MNode* t, const MeshBuildCtx& c, MeshTree& out) {
    if (t->leaf) {
        int32_t li = (int32_t)out.leafoff.size();
        out.leafoff.push_back((int32_t)out.leafpool.size());
        out.leafpool.push_back((int32_t)(t->hi - t->lo));
        for (int64_t p = t->lo; p < t->hi; p++) out.leafpool.push_back(c.idx[p]);
        delete t;
        return ~li;
    }
    int32_t me = (int32_t)out.nodes.size();
    GlomeBvhNode nd;
    memset(&nd, 0, sizeof(nd));
    const Bbox& L = t->lbb; const Bbox& R = t->rbb;
    nd.lbb[0] = L.p1.x; nd.lbb[1] = L.p1.y; nd.lbb[2] = L.p1.z; nd.lbb[3] = L.p2.x; nd.lbb[4] = L.p2.y; nd.lbb[5] = L.p2.z;
    nd.rbb[0] = R.p1.x; nd.rbb[1] = R.p1.y; nd.rbb[2] = R.p1.z; nd.rbb[3] = R.p2.x; nd.rbb[4] = R.p2.y; nd.rbb[5] = R.p2.z;
    out.nodes.push_back(nd);
    int32_t l = mesh_number(t->l, c, out);
    int32_t r = mesh_number(t->r, c, out);
    out.nodes[me].left = l;
    out.nodes[me].right = r;
    delete t;
    return me;
}

static inline Bbox pt_box(const Vec& p) {
    return mkbb(vec(p.x - kDelta, p.y - kDelta, p.z - kDelta), vec(p.x + kDelta, p.y + kDelta, p.z + kDelta));
}

}  // namespace

void mesh_build(int64_t nverts, const double* verts, int64_t ntris, const int32_t* tris, MeshTree& out) {
    // bbox = bbpts (V.toList verts) (Mesh.hs:55; Vec.hs:676-690): every point inflated by delta.
    Bbox bb = empty_bbox();
    for (int64_t i = nverts - 1; i >= 0; i--) {
        Bbox pb = pt_box(vec(verts[3 * i], verts[3 * i + 1], verts[3 * i + 2]));
        if (i == nverts - 1) bb = pb;
        else bb = mkbb(vec(fmin_(pb.p1.x, bb.p1.x), fmin_(pb.p1.y, bb.p1.y), fmin_(pb.p1.z, bb.p1.z)),
                       vec(fmax_(pb.p2.x, bb.p2.x), fmax_(pb.p2.y, bb.p2.y), fmax_(pb.p2.z, bb.p2.z)));
    }
    MeshBuildCtx c;
    c.tbb.resize((size_t)ntris); c.mid.resize((size_t)ntris); c.sa.resize((size_t)ntris);
    c.idx.resize((size_t)ntris); c.scratch.resize((size_t)ntris);
    for (int64_t i = 0; i < ntris; i++) {
        const int32_t* T = tris + 8 * i;
        // alltribbs: bbpts [a, b, c] (Mesh.hs:119-121)
        Bbox r = pt_box(vec(verts[3 * T[2]], verts[3 * T[2] + 1], verts[3 * T[2] + 2]));
        for (int j = 1; j >= 0; j--) {
            Bbox pb = pt_box(vec(verts[3 * T[j]], verts[3 * T[j] + 1], verts[3 * T[j] + 2]));
            r = mkbb(vec(fmin_(pb.p1.x, r.p1.x), fmin_(pb.p1.y, r.p1.y), fmin_(pb.p1.z, r.p1.z)),
                     vec(fmax_(pb.p2.x, r.p2.x), fmax_(pb.p2.y, r.p2.y), fmax_(pb.p2.z, r.p2.z)));
        }
        c.tbb[i] = r;
        c.mid[i] = bbmid(r);
        c.sa[i] = bbsa(r);
        c.idx[i] = (int32_t)i;
    }
    MNode* root = mesh_rec(c, 0, ntris, bb, 0);
    out.bb = bb;
    out.nodes.clear(); out.leafpool.clear(); out.leafoff.clear();
    out.root = mesh_number(root, c, out);
}

// ---------------------------------------------------------------------------------------------
// constructors
// ---------------------------------------------------------------------------------------------
int Builder::add(const Item& it) {
    items.push_back(it);
    return (int)items.size() - 1;
}
int Builder::check(int s) const {
    if (s < 0 || s >= (int)items.size()) throw BuildError("bad item id");
    return s;
}
static Item mkitem(int type) {
    Item it;
    it.type = type; it.ia = it.ib = 0; it.nd = 0;
    memset(it.d, 0, sizeof(it.d));
    return it;
}
static void putv(Item& it, int off, const Vec& v) { it.d[off] = v.x; it.d[off + 1] = v.y; it.d[off + 2] = v.z; }
static Vec getv(const Item& it, int off) { return vec(it.d[off], it.d[off + 1], it.d[off + 2]); }

int Builder::void_() { return add(mkitem(GLOME_VOID)); }
int Builder::sphere(const Vec& c, Flt r) {  // Sphere.hs:15-17
    Item it = mkitem(GLOME_SPHERE);
    putv(it, 0, c); it.d[3] = r; it.nd = 4;
    return add(it);
}
int Builder::triangle(const Vec& a, const Vec& b, const Vec& c) {  // Triangle.hs:18
    Item it = mkitem(GLOME_TRIANGLE);
    putv(it, 0, a); putv(it, 3, b); putv(it, 6, c); it.nd = 9;
    return add(it);
}
int Builder::trianglenorm(const Vec& a, const Vec& b, const Vec& c, const Vec& na, const Vec& nb, const Vec& nc) {
    Item it = mkitem(GLOME_TRIANGLENORM);
    putv(it, 0, a); putv(it, 3, b); putv(it, 6, c); putv(it, 9, na); putv(it, 12, nb); putv(it, 15, nc); it.nd = 18;
    return add(it);
}
int Builder::box(const Vec& a, const Vec& b) {  // Box.hs:12-15
    Item it = mkitem(GLOME_BOX);
    putv(it, 0, vec(fmin_(a.x, b.x), fmin_(a.y, b.y), fmin_(a.z, b.z)));
    putv(it, 3, vec(fmax_(a.x, b.x), fmax_(a.y, b.y), fmax_(a.z, b.z)));
    it.nd = 6;
    return add(it);
}
int Builder::plane(const Vec& orig, const Vec& norm_) {  // Plane.hs:17-20
    Vec norm = vnorm(norm_);
    return plane_offset(norm, vdot(orig, norm));
}
int Builder::plane_offset(const Vec& n, Flt off) {  // Plane.hs:24-25
    Item it = mkitem(GLOME_PLANE);
    putv(it, 0, n); it.d[3] = off; it.nd = 4;
    return add(it);
}
int Builder::disc(const Vec& pos, const Vec& norm, Flt r) {  // Cone.hs:29-31
    Item it = mkitem(GLOME_DISC);
    putv(it, 0, pos); putv(it, 3, norm); it.d[6] = r * r; it.nd = 7;
    return add(it);
}
int Builder::disc_raw(const Vec& pos, const Vec& norm, Flt rsqr) {  // Cone.hs:21: the stored fields
    Item it = mkitem(GLOME_DISC);
    putv(it, 0, pos); putv(it, 3, norm); it.d[6] = rsqr; it.nd = 7;
    return add(it);
}
int Builder::cylinder_z(Flt r, Flt h1, Flt h2) {  // Cone.hs:33
    Item it = mkitem(GLOME_CYLINDER);
    it.d[0] = r; it.d[1] = h1; it.d[2] = h2; it.nd = 3;
    return add(it);
}
int Builder::cone_z(Flt r, Flt h1, Flt h2, Flt height) {  // Cone.hs:36
    Item it = mkitem(GLOME_CONE);
    it.d[0] = r; it.d[1] = h1; it.d[2] = h2; it.d[3] = height; it.nd = 4;
    return add(it);
}
int Builder::cylinder(const Vec& p1, const Vec& p2, Flt r) {  // Cone.hs:40-48
    Vec axis = vsub(p2, p1);
    Flt len = vlen(axis);
    Vec ax1 = vscale(axis, 1 / len);
    Vec ax2, ax3;
    orth(ax1, ax2, ax3);
    std::vector<Xfm> xs;
    xs.push_back(xyz_to_uvw(ax2, ax3, ax1));
    xs.push_back(translate(p1));
    return transform(cylinder_z(r, 0, len), xs);
}
int Builder::cone(const Vec& p1, Flt r1, const Vec& p2, Flt r2) {  // Cone.hs:52-67
    if (r1 < r2) return cone(p2, r2, p1, r1);
    if (r1 - r2 < kDelta) return cylinder(p1, p2, r2);
    Vec axis = vsub(p2, p1);
    Flt len = vlen(axis);
    Vec ax1 = vscale(axis, 1 / len);
    Vec ax2, ax3;
    orth(ax1, ax2, ax3);
    Flt height = (r1 * len) / (r1 - r2);
    std::vector<Xfm> xs;
    xs.push_back(xyz_to_uvw(ax2, ax3, ax1));
    xs.push_back(translate(p1));
    return transform(cone_z(r1, 0, len, height), xs);
}
int Builder::list_raw(const std::vector<int32_t>& xs) {
    Item it = mkitem(GLOME_GROUP);
    it.kids = xs;
    for (size_t i = 0; i < xs.size(); i++) check(xs[i]);
    return add(it);
}
int Builder::group(const std::vector<int32_t>& xs) {  // Solid.hs:293-302
    if (xs.empty()) return void_();
    if (xs.size() == 1) return check(xs[0]);
    std::vector<int32_t> flat;
    for (size_t i = 0; i < xs.size(); i++) {
        std::vector<int32_t> l = tolist(xs[i]);
        flat.insert(flat.end(), l.begin(), l.end());
    }
    return list_raw(flat);
}
int Builder::bih(const std::vector<int32_t>& xs) {  // Bih.hs:309-324
    if (xs.empty()) return void_();
    std::vector<double> bbs(xs.size() * 6);
    for (size_t i = 0; i < xs.size(); i++) {
        Bbox b = bound(xs[i]);
        double* o = &bbs[6 * i];
        o[0] = b.p1.x; o[1] = b.p1.y; o[2] = b.p1.z; o[3] = b.p2.x; o[4] = b.p2.y; o[5] = b.p2.z;
    }
    bihs.emplace_back();
    auto t0 = std::chrono::steady_clock::now();
    if (build_device >= 0) bih_build_gpu((int64_t)xs.size(), bbs.data(), build_device, bihs.back(), build_ms);
    else { bih_build((int64_t)xs.size(), bbs.data(), bihs.back()); build_ms[0] = build_ms[1] = build_ms[2] = 0; }
    build_ms[3] = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count();
    Item it = mkitem(GLOME_BIH);
    it.ia = (int)bihs.size() - 1;
    it.kids = xs;
    return add(it);
}
// A Bih whose tree the caller already built (the Haskell `bih` constructor: Bih.hs:309-324 returns `Bih bb root`).
// The tree arrives as a PRE-ORDER stream mirroring the constructors of Bih.hs:51-57:
//   kinds[i] >= 0 : BihBranch lsplit rsplit axis l r   with axis = kinds[i], lsplit = splits[2i], rsplit = splits[2i+1];
//                   the records of l follow, then those of r
//   kinds[i] <  0 : BihLeaf [s]  holding the next  -(kinds[i] + 1)  entries of `xs`
// so `xs` lists the items in the order the leaves hold them.  The result is the Item glome_sb_bih would have made had
// its own builder produced this tree (tests/test_host_builder.py imports a tree it built and compares the FlatScenes).
int Builder::bih_prebuilt(const std::vector<int32_t>& xs, int64_t n_nodes, const int32_t* kinds, const double* splits, const double bb[6]) {
    if (n_nodes <= 0) throw BuildError("bih_prebuilt: empty tree (an empty bih is Void, Bih.hs:312)");
    for (size_t i = 0; i < xs.size(); i++) check(xs[i]);
    bihs.emplace_back();
    BihTree& T = bihs.back();
    T.bb = mkbb(vec(bb[0], bb[1], bb[2]), vec(bb[3], bb[4], bb[5]));
    T.order.resize(xs.size());
    for (size_t i = 0; i < xs.size(); i++) T.order[i] = (int32_t)i;
    int64_t cur = 0, item = 0;
    // explicit stack: (node index whose child slot to fill, which slot); -1 = the root ref
    struct Pending { int32_t node; int slot; };
    std::vector<Pending> st;
    st.push_back({-1, 0});
    while (!st.empty()) {
        Pending p = st.back();
        st.pop_back();
        if (cur >= n_nodes) throw BuildError("bih_prebuilt: the pre-order stream ends inside a branch");
        const int32_t k = kinds[cur];
        int32_t ref;
        if (k < 0) {
            const int64_t cnt = -((int64_t)k + 1);
            if (item + cnt > (int64_t)xs.size()) throw BuildError("bih_prebuilt: leaves hold more items than were passed");
            ref = ~(int32_t)(T.leaves.size() / 2);
            T.leaves.push_back((int32_t)item);
            T.leaves.push_back((int32_t)cnt);
            item += cnt;
        } else {
            if (k > 2) throw BuildError("bih_prebuilt: axis out of range");
            ref = (int32_t)T.nodes.size();
            GlomeBihNode nd;
            memset(&nd, 0, sizeof(nd));
            nd.lsplit = splits[2 * cur]; nd.rsplit = splits[2 * cur + 1]; nd.axis = k;
            T.nodes.push_back(nd);
            st.push_back({ref, 1});  // right is parsed after the whole left subtree
            st.push_back({ref, 0});
            if (st.size() > 4096) throw BuildError("bih_prebuilt: tree too deep");
        }
        if (p.node < 0) T.root = ref;
        else if (p.slot == 0) T.nodes[p.node].left = ref;
        else T.nodes[p.node].right = ref;
        cur++;
    }
    if (cur != n_nodes) throw BuildError("bih_prebuilt: records left over after the tree");
    if (item != (int64_t)xs.size()) throw BuildError("bih_prebuilt: the leaves do not hold every item");
    build_ms[0] = build_ms[1] = build_ms[2] = build_ms[3] = 0;
    Item it = mkitem(GLOME_BIH);
    it.ia = (int)bihs.size() - 1;
    it.kids = xs;
    return add(it);
}

// A Mesh whose BVH the caller already built (Mesh.hs:36-42: `Branch lbb rbb l r` / `Leaf [Tri]`), same pre-order stream:
//   kinds[i] >= 0 : Branch with lbb = boxes[12i .. 12i+5], rbb = boxes[12i+6 .. 12i+11]
//   kinds[i] <  0 : Leaf holding the next -(kinds[i] + 1) entries of leaf_tris (mesh-local triangle indices)
int Builder::mesh_prebuilt(int64_t nverts, const double* verts, int64_t nnorms, const double* norms, int64_t ntris,
                           const int32_t* tris, int ntexs, const int32_t* texs, int ntags, const int32_t* tags, int64_t n_nodes,
                           const int32_t* kinds, const double* boxes, int64_t n_leaf_tris, const int32_t* leaf_tris, const double bb[6]) {
    if (n_nodes <= 0) throw BuildError("mesh_prebuilt: empty tree");
    for (int64_t i = 0; i < ntris; i++) {
        const int32_t* T = tris + 8 * i;
        for (int j = 0; j < 3; j++)
            if (T[j] < 0 || T[j] >= nverts) throw BuildError("mesh: vertex index out of range");
        if (T[3] != -1)
            for (int j = 3; j < 6; j++)
                if (T[j] < 0 || T[j] >= nnorms) throw BuildError("mesh: normal index out of range");
        if (T[6] < -1 || T[6] >= ntexs) throw BuildError("mesh: texture index out of range");
        if (T[7] < -1 || T[7] >= ntags) throw BuildError("mesh: tag index out of range");
    }
    meshes.emplace_back();
    MeshData& m = meshes.back();
    m.verts.assign(verts, verts + 3 * nverts);
    if (nnorms) m.norms.assign(norms, norms + 3 * nnorms);
    m.tris.assign(tris, tris + 8 * ntris);
    if (ntexs) m.texs.assign(texs, texs + ntexs);
    if (ntags) m.tags.assign(tags, tags + ntags);
    MeshTree& T = m.tree;
    T.bb = mkbb(vec(bb[0], bb[1], bb[2]), vec(bb[3], bb[4], bb[5]));
    int64_t cur = 0, tpos = 0;
    struct Pending { int32_t node; int slot; };
    std::vector<Pending> st;
    st.push_back({-1, 0});
    while (!st.empty()) {
        Pending p = st.back();
        st.pop_back();
        if (cur >= n_nodes) throw BuildError("mesh_prebuilt: the pre-order stream ends inside a branch");
        const int32_t k = kinds[cur];
        int32_t ref;
        if (k < 0) {
            const int64_t cnt = -((int64_t)k + 1);
            if (tpos + cnt > n_leaf_tris) throw BuildError("mesh_prebuilt: leaves hold more triangles than were passed");
            ref = ~(int32_t)T.leafoff.size();
            T.leafoff.push_back((int32_t)T.leafpool.size());
            T.leafpool.push_back((int32_t)cnt);
            for (int64_t q = 0; q < cnt; q++) {
                if (leaf_tris[tpos + q] < 0 || leaf_tris[tpos + q] >= ntris) throw BuildError("mesh_prebuilt: triangle index out of range");
                T.leafpool.push_back(leaf_tris[tpos + q]);
            }
            tpos += cnt;
        } else {
            ref = (int32_t)T.nodes.size();
            GlomeBvhNode nd;
            memset(&nd, 0, sizeof(nd));
            memcpy(nd.lbb, boxes + 12 * cur, 6 * sizeof(double));
            memcpy(nd.rbb, boxes + 12 * cur + 6, 6 * sizeof(double));
            T.nodes.push_back(nd);
            st.push_back({ref, 1});
            st.push_back({ref, 0});
            if (st.size() > 4096) throw BuildError("mesh_prebuilt: tree too deep");
        }
        if (p.node < 0) T.root = ref;
        else if (p.slot == 0) T.nodes[p.node].left = ref;
        else T.nodes[p.node].right = ref;
        cur++;
    }
    if (cur != n_nodes) throw BuildError("mesh_prebuilt: records left over after the tree");
    if (tpos != n_leaf_tris) throw BuildError("mesh_prebuilt: triangle indices left over after the tree");
    build_ms[0] = build_ms[1] = build_ms[2] = build_ms[3] = 0;
    Item it = mkitem(GLOME_MESH);
    it.ia = (int)meshes.size() - 1;
    return add(it);
}
int Builder::mesh(int64_t nverts, const double* verts, int64_t nnorms, const double* norms, int64_t ntris,
                  const int32_t* tris, int ntexs, const int32_t* texs, int ntags, const int32_t* tags) {
    for (int64_t i = 0; i < ntris; i++) {
        const int32_t* T = tris + 8 * i;
        for (int j = 0; j < 3; j++)
            if (T[j] < 0 || T[j] >= nverts) throw BuildError("mesh: vertex index out of range");
        if (T[3] != -1)
            for (int j = 3; j < 6; j++)
                if (T[j] < 0 || T[j] >= nnorms) throw BuildError("mesh: normal index out of range");
        if (T[6] < -1 || T[6] >= ntexs) throw BuildError("mesh: texture index out of range");
        if (T[7] < -1 || T[7] >= ntags) throw BuildError("mesh: tag index out of range");
    }
    meshes.emplace_back();
    MeshData& m = meshes.back();
    m.verts.assign(verts, verts + 3 * nverts);
    if (nnorms) m.norms.assign(norms, norms + 3 * nnorms);
    m.tris.assign(tris, tris + 8 * ntris);
    if (ntexs) m.texs.assign(texs, texs + ntexs);
    if (ntags) m.tags.assign(tags, tags + ntags);
    auto t0 = std::chrono::steady_clock::now();
    if (build_device >= 0) mesh_build_gpu(nverts, verts, ntris, tris, build_device, m.tree, build_ms);
    else { mesh_build(nverts, verts, ntris, tris, m.tree); build_ms[0] = build_ms[1] = build_ms[2] = 0; }
    build_ms[3] = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count();
    Item it = mkitem(GLOME_MESH);
    it.ia = (int)meshes.size() - 1;
    return add(it);
}
int Builder::difference(int sa, int sb) {  // Csg.hs:26-27
    Item it = mkitem(GLOME_DIFFERENCE);
    it.kids.push_back(check(sa)); it.kids.push_back(check(sb)); it.ia = 1;
    return add(it);
}
int Builder::difference_ex(int sa, int sb, bool useatex) {  // Csg.hs:26-30 (difference / difference_retexture)
    Item it = mkitem(GLOME_DIFFERENCE);
    it.kids.push_back(check(sa)); it.kids.push_back(check(sb)); it.ia = useatex ? 1 : 0;
    return add(it);
}
int Builder::intersection(const std::vector<int32_t>& xs) {  // Csg.hs:64-65
    Item it = mkitem(GLOME_INTERSECTION);
    it.kids = xs;
    for (size_t i = 0; i < xs.size(); i++) check(xs[i]);
    return add(it);
}
int Builder::tex(int s, int texture) {
    if (texture < 0 || texture >= (int)textures.size()) throw BuildError("bad texture id");
    Item it = mkitem(GLOME_TEX);
    it.kids.push_back(check(s)); it.ia = texture;
    return add(it);
}
int Builder::tag(int s, int tagid) {
    Item it = mkitem(GLOME_TAG);
    it.kids.push_back(check(s)); it.ia = tagid;
    return add(it);
}
int Builder::noshadow(int s) { Item it = mkitem(GLOME_NOSHADOW); it.kids.push_back(check(s)); return add(it); }
int Builder::onlyshadow(int s) { Item it = mkitem(GLOME_ONLYSHADOW); it.kids.push_back(check(s)); return add(it); }
int Builder::bound_object(int sa, int sb) {
    Item it = mkitem(GLOME_BOUND);
    it.kids.push_back(check(sa)); it.kids.push_back(check(sb));
    return add(it);
}
int Builder::innerbound(int sa, int sb) {
    Item it = mkitem(GLOME_INNERBOUND);
    it.kids.push_back(check(sa)); it.kids.push_back(check(sb));
    return add(it);
}
int Builder::instance_raw(int s, const Xfm& x) {
    Item it = mkitem(GLOME_INSTANCE);
    it.kids.push_back(check(s));
    memcpy(it.d, x.m, sizeof(x.m)); it.nd = 24;
    return add(it);
}
static Xfm item_xfm(const Item& it) { Xfm x; memcpy(x.m, it.d, sizeof(x.m)); return x; }

int Builder::transform(int s, const std::vector<Xfm>& xs) {
    const Item it = items[check(s)];
    switch (it.type) {
        case GLOME_VOID: return void_();  // Solid.hs:360
        case GLOME_TRIANGLE: {            // Triangle.hs:164-168
            Xfm c = compose(xs);
            return triangle(xfm_point(c.m, getv(it, 0)), xfm_point(c.m, getv(it, 3)), xfm_point(c.m, getv(it, 6)));
        }
        case GLOME_TRIANGLENORM: {        // Triangle.hs:170-177
            Xfm c = compose(xs);
            return trianglenorm(xfm_point(c.m, getv(it, 0)), xfm_point(c.m, getv(it, 3)), xfm_point(c.m, getv(it, 6)),
                                vnorm(xfm_vec(c.m, getv(it, 9))), vnorm(xfm_vec(c.m, getv(it, 12))),
                                vnorm(xfm_vec(c.m, getv(it, 15))));
        }
        case GLOME_INSTANCE: {            // Solid.hs:494-496
            std::vector<Xfm> ys;
            ys.push_back(item_xfm(it));
            ys.insert(ys.end(), xs.begin(), xs.end());
            std::vector<Xfm> one(1, compose(ys));
            return transform(it.kids[0], one);
        }
        default: return instance_raw(s, compose(xs));  // Solid.hs:235
    }
}
int Builder::transform_leaf(int s, const std::vector<Xfm>& xs) {
    const Item it = items[check(s)];
    switch (it.type) {
        case GLOME_GROUP: {  // Solid.hs:334
            std::vector<int32_t> l = tolist(s), o;
            for (size_t i = 0; i < l.size(); i++) o.push_back(transform_leaf(l[i], xs));
            return list_raw(o);
        }
        case GLOME_INSTANCE: {  // Solid.hs:498-500
            std::vector<Xfm> ys;
            ys.push_back(item_xfm(it));
            ys.insert(ys.end(), xs.begin(), xs.end());
            std::vector<Xfm> one(1, compose(ys));
            return transform_leaf(it.kids[0], one);
        }
        case GLOME_BOUND:
        case GLOME_INNERBOUND: return transform_leaf(it.kids[1], xs);  // Bound.hs:69-71, 114
        default: return transform(s, xs);                              // Solid.hs:240
    }
}
std::vector<int32_t> Builder::tolist(int s) {
    const Item& it = items[check(s)];
    std::vector<int32_t> o;
    if (it.type == GLOME_VOID) return o;  // Solid.hs:359
    if (it.type == GLOME_GROUP) {         // Solid.hs:333
        std::vector<int32_t> kids = it.kids;
        for (size_t i = 0; i < kids.size(); i++) {
            std::vector<int32_t> l = tolist(kids[i]);
            o.insert(o.end(), l.begin(), l.end());
        }
        return o;
    }
    o.push_back(s);  // Solid.hs:230
    return o;
}
std::vector<int32_t> Builder::flatten_transform(int s) {
    // flatten_transform (SolidItem s) = [SolidItem (flatten_transform s)]  (Solid.hs:273)
    const Item it = items[check(s)];
    std::vector<int32_t> inner;
    switch (it.type) {
        case GLOME_GROUP:  // Solid.hs:335
            for (size_t i = 0; i < it.kids.size(); i++) {
                std::vector<int32_t> l = flatten_transform(it.kids[i]);
                inner.insert(inner.end(), l.begin(), l.end());
            }
            break;
        case GLOME_INSTANCE: {  // Solid.hs:509-511
            std::vector<Xfm> one(1, item_xfm(it));
            inner.push_back(transform_leaf(it.kids[0], one));
            break;
        }
        case GLOME_BOUND:
        case GLOME_INNERBOUND: inner = flatten_transform(it.kids[1]); break;  // Bound.hs:73-74, 113
        default: inner = tolist(s); break;                                    // Solid.hs:246
    }
    std::vector<int32_t> o(1, list_raw(inner));
    return o;
}

Bbox Builder::bound(int s) {
    const Item& it = items[check(s)];
    switch (it.type) {
        case GLOME_VOID: return empty_bbox();  // Solid.hs:358
        case GLOME_SPHERE: {                   // Sphere.hs:78-81
            Vec c = getv(it, 0), off = vec(it.d[3], it.d[3], it.d[3]);
            return mkbb(vsub(c, off), vadd(c, off));
        }
        case GLOME_TRIANGLE:
        case GLOME_TRIANGLENORM: {  // Triangle.hs:147-162
            Vec a = getv(it, 0), b = getv(it, 3), c = getv(it, 6);
            return mkbb(vec(fmin_(fmin_(a.x, b.x), c.x) - kDelta, fmin_(fmin_(a.y, b.y), c.y) - kDelta,
                            fmin_(fmin_(a.z, b.z), c.z) - kDelta),
                        vec(fmax_(fmax_(a.x, b.x), c.x) + kDelta, fmax_(fmax_(a.y, b.y), c.y) + kDelta,
                            fmax_(fmax_(a.z, b.z), c.z) + kDelta));
        }
        case GLOME_BOX: return mkbb(getv(it, 0), getv(it, 3));  // Box.hs:70-71
        case GLOME_PLANE: return everything_bbox();             // Plane.hs:43
        case GLOME_DISC: {                                      // Cone.hs:93-95
            Vec c = getv(it, 0);
            Flt r = sqrt(it.d[6]);
            Vec off = vec(r, r, r);
            return mkbb(vsub(c, off), vadd(c, off));
        }
        case GLOME_CYLINDER: return mkbb(vec(-it.d[0], -it.d[0], it.d[1]), vec(it.d[0], it.d[0], it.d[2]));  // Cone.hs:145
        case GLOME_CONE: return mkbb(vec(-it.d[0], -it.d[0], it.d[1]), vec(it.d[0], it.d[0], it.d[2]));      // Cone.hs:253
        case GLOME_GROUP: {  // Solid.hs:332
            Bbox b = empty_bbox();
            std::vector<int32_t> kids = it.kids;
            for (size_t i = 0; i < kids.size(); i++) b = bbjoin(b, bound(kids[i]));
            return b;
        }
        case GLOME_INSTANCE: {  // Solid.hs:477-484: bbpts of the 8 transformed corners
            Xfm x = item_xfm(it);
            Bbox cb = bound(it.kids[0]);
            Flt xs[2] = {cb.p1.x, cb.p2.x}, ys[2] = {cb.p1.y, cb.p2.y}, zs[2] = {cb.p1.z, cb.p2.z};
            Vec pts[8];
            int k = 0;
            for (int a = 0; a < 2; a++)
                for (int b = 0; b < 2; b++)
                    for (int c = 0; c < 2; c++) pts[k++] = xfm_point(x.m, vec(xs[a], ys[b], zs[c]));
            Bbox r = mkbb(vec(pts[7].x - kDelta, pts[7].y - kDelta, pts[7].z - kDelta),
                          vec(pts[7].x + kDelta, pts[7].y + kDelta, pts[7].z + kDelta));
            for (int i = 6; i >= 0; i--) {
                const Vec& p = pts[i];
                r = mkbb(vec(fmin_(p.x - kDelta, r.p1.x), fmin_(p.y - kDelta, r.p1.y), fmin_(p.z - kDelta, r.p1.z)),
                         vec(fmax_(p.x + kDelta, r.p2.x), fmax_(p.y + kDelta, r.p2.y), fmax_(p.z + kDelta, r.p2.z)));
            }
            return r;
        }
        case GLOME_BIH: return bihs[it.ia].bb;           // Bih.hs:588-589
        case GLOME_MESH: return meshes[it.ia].tree.bb;   // Mesh.hs:212
        case GLOME_DIFFERENCE: return bound(it.kids[0]); // Csg.hs:113-114
        case GLOME_INTERSECTION: {                       // Csg.hs:116-120
            if (it.kids.empty()) return empty_bbox();
            Bbox b = everything_bbox();
            std::vector<int32_t> kids = it.kids;
            for (size_t i = 0; i < kids.size(); i++) b = bboverlap(b, bound(kids[i]));
            return b;
        }
        case GLOME_TEX:
        case GLOME_TAG:
        case GLOME_NOSHADOW:
        case GLOME_ONLYSHADOW: return bound(it.kids[0]);
        case GLOME_BOUND: { int a = it.kids[0], b = it.kids[1]; return bboverlap(bound(a), bound(b)); }  // Bound.hs:61-62
        case GLOME_INNERBOUND: return bound(it.kids[1]);                                                 // Bound.hs:110
    }
    return empty_bbox();
}

// ---- materials / textures / lights ----
static GlomeMaterial mkmat(int kind) {
    GlomeMaterial m;
    memset(&m, 0, sizeof(m));
    m.kind = kind;
    return m;
}
int Builder::mat_surface(Flt r, Flt g, Flt b, Flt alpha, Flt amb, Flt kd, Flt ks, Flt shine) {
    GlomeMaterial m = mkmat(GLOME_MAT_SURFACE);
    m.p[0] = r; m.p[1] = g; m.p[2] = b; m.p[3] = alpha; m.p[4] = amb; m.p[5] = kd; m.p[6] = ks; m.p[7] = shine;
    materials.push_back(m);
    return (int)materials.size() - 1;
}
int Builder::mat_reflect(Flt refl) {
    GlomeMaterial m = mkmat(GLOME_MAT_REFLECT);
    m.p[0] = refl;
    materials.push_back(m);
    return (int)materials.size() - 1;
}
int Builder::mat_refract(Flt refl, Flt refr, Flt ior) {
    GlomeMaterial m = mkmat(GLOME_MAT_REFRACT);
    m.p[0] = refl; m.p[1] = refr; m.p[2] = ior;
    materials.push_back(m);
    return (int)materials.size() - 1;
}
int Builder::mat_warp(int frame, int scene, int ls, const Xfm& x) {
    GlomeMaterial m = mkmat(GLOME_MAT_WARP);
    m.a = frame; m.b = scene; m.c = ls;
    m.d = (int)warp_xfms.size();
    warp_xfms.push_back(x);
    materials.push_back(m);
    return (int)materials.size() - 1;
}
int Builder::mat_additive(const std::vector<int32_t>& ms) {
    GlomeMaterial m = mkmat(GLOME_MAT_ADDITIVE);
    m.a = (int)mat_lists.size(); m.b = (int)ms.size();
    mat_lists.insert(mat_lists.end(), ms.begin(), ms.end());
    materials.push_back(m);
    return (int)materials.size() - 1;
}
int Builder::mat_blend(int ma, int mb, Flt w) {
    GlomeMaterial m = mkmat(GLOME_MAT_BLEND);
    m.a = ma; m.b = mb; m.p[0] = w;
    materials.push_back(m);
    return (int)materials.size() - 1;
}
static GlomeTexture mktex(int kind, int a, int b) {
    GlomeTexture t;
    memset(&t, 0, sizeof(t));
    t.kind = kind; t.a = a; t.b = b;
    return t;
}
int Builder::tex_uniform(int mat) { textures.push_back(mktex(GLOME_TEX_UNIFORM, mat, 0)); return (int)textures.size() - 1; }
int Builder::tex_stripe_blend(int ma, int mb, const Vec& axis) {
    GlomeTexture t = mktex(GLOME_TEX_STRIPE_BLEND, ma, mb);
    t.p[0] = axis.x; t.p[1] = axis.y; t.p[2] = axis.z;
    textures.push_back(t);
    return (int)textures.size() - 1;
}
int Builder::tex_perlin_blend(int ma, int mb, Flt scale) {
    GlomeTexture t = mktex(GLOME_TEX_PERLIN_BLEND, ma, mb);
    t.p[0] = scale;
    textures.push_back(t);
    return (int)textures.size() - 1;
}
int Builder::light(const Vec& pos, Flt r, Flt g, Flt b) {  // Shader.hs:22-23
    GlomeLight l;
    memset(&l, 0, sizeof(l));
    l.pos[0] = pos.x; l.pos[1] = pos.y; l.pos[2] = pos.z;
    l.color[0] = r; l.color[1] = g; l.color[2] = b;
    l.rad = kInf; l.falloff = 0; l.do_shadow = 1;
    lights.push_back(l);
    return (int)lights.size() - 1;
}
int Builder::lightset(const std::vector<int32_t>& ls) {
    // light sets are ranges: the lights of a set must be consecutive ids
    if (ls.empty()) { lightsets.push_back(0); lightsets.push_back(0); }
    else {
        for (size_t i = 1; i < ls.size(); i++)
            if (ls[i] != ls[0] + (int)i) throw BuildError("lightset: lights must be consecutive");
        lightsets.push_back(ls[0]); lightsets.push_back((int)ls.size());
    }
    return (int)(lightsets.size() / 2) - 1;
}

// ---------------------------------------------------------------------------------------------
// flatten
// ---------------------------------------------------------------------------------------------
int Builder::alloc_nodes(int n) {
    int first = (int)f_nodes.size();
    GlomeNode z;
    z.type = GLOME_VOID; z.a = z.b = z.c = 0;
    f_nodes.resize(f_nodes.size() + (size_t)n, z);
    return first;
}
int Builder::alloc_d(int n, int align) {
    while (f_dpool.size() % (size_t)align) f_dpool.push_back(0);
    int off = (int)f_dpool.size();
    f_dpool.resize(f_dpool.size() + (size_t)n, 0.0);
    return off;
}
static void put_bb(std::vector<double>& d, int off, const Bbox& b) {
    d[off] = b.p1.x; d[off + 1] = b.p1.y; d[off + 2] = b.p1.z; d[off + 3] = b.p2.x; d[off + 4] = b.p2.y; d[off + 5] = b.p2.z;
}

void Builder::flatten_into(int item, int slot, int depth) {
    if (depth > f_maxdepth) f_maxdepth = depth;
    if (depth > 256) throw BuildError("flatten: scene graph too deep (cycle?)");
    const Item it = items[check(item)];
    GlomeNode nd;
    nd.type = it.type; nd.a = nd.b = nd.c = 0;
    switch (it.type) {
        case GLOME_VOID: break;
        case GLOME_SPHERE: case GLOME_TRIANGLE: case GLOME_TRIANGLENORM: case GLOME_BOX: case GLOME_PLANE:
        case GLOME_DISC: case GLOME_CYLINDER: case GLOME_CONE: {
            int off = alloc_d(it.nd, it.type == GLOME_SPHERE ? 4 : 2);
            memcpy(&f_dpool[off], it.d, sizeof(double) * (size_t)it.nd);
            nd.a = off;
            break;
        }
        case GLOME_GROUP:
        case GLOME_INTERSECTION: {
            int first = alloc_nodes((int)it.kids.size());
            nd.a = first; nd.b = (int)it.kids.size();
            for (size_t i = 0; i < it.kids.size(); i++) flatten_into(it.kids[i], first + (int)i, depth + 1);
            break;
        }
        case GLOME_INSTANCE: {
            int off = alloc_d(24, 2);
            memcpy(&f_dpool[off], it.d, sizeof(double) * 24);
            int c = alloc_nodes(1);
            nd.a = c; nd.b = off;
            flatten_into(it.kids[0], c, depth + 1);
            break;
        }
        case GLOME_BIH: {
            if (f_bih_memo[it.ia] >= 0) {  // shared BIH: reuse the flattened tree
                const int32_t* m = &f_ipool[f_bih_memo[it.ia]];
                nd.a = m[0]; nd.b = m[1]; nd.c = m[2];
                break;
            }
            const BihTree& T = bihs[it.ia];
            int n = (int)it.kids.size();
            int first = alloc_nodes(n);
            bool all_spheres = true;
            for (int j = 0; j < n; j++)
                if (items[it.kids[T.order[j]]].type != GLOME_SPHERE) { all_spheres = false; break; }
            if (all_spheres) {  // linear sphere block: payloads contiguous in leaf order
                int off = alloc_d(4 * n, 4);
                for (int j = 0; j < n; j++) {
                    const Item& s = items[it.kids[T.order[j]]];
                    memcpy(&f_dpool[off + 4 * j], s.d, sizeof(double) * 4);
                    GlomeNode sn;
                    sn.type = GLOME_SPHERE; sn.a = off + 4 * j; sn.b = sn.c = 0;
                    f_nodes[first + j] = sn;
                }
                if (depth + 1 > f_maxdepth) f_maxdepth = depth + 1;
            } else {
                for (int j = 0; j < n; j++) flatten_into(it.kids[T.order[j]], first + j, depth + 1);
            }
            int nbase = (int)f_bih.size();
            // leaf refs: inline {first item node, count} when they fit, else a record in ipool
            std::vector<int32_t> leafoff(T.leaves.size() / 2);
            for (size_t l = 0; l < T.leaves.size() / 2; l++) {
                int32_t lfirst = first + T.leaves[2 * l], lcount = T.leaves[2 * l + 1];
                if (lcount <= 6 && lfirst < (1 << 27)) leafoff[l] = ~glome_bih_leaf_ref_inline(lfirst, lcount);
                else {
                    if ((int64_t)f_ipool.size() >= (1 << 27)) throw BuildError("flatten: ipool too large for a leaf record");
                    leafoff[l] = ~glome_bih_leaf_ref_escape((int32_t)f_ipool.size());
                    f_ipool.push_back(lfirst);
                    f_ipool.push_back(lcount);
                }
            }
            auto xlate = [&](int32_t ref) -> int32_t { return ref >= 0 ? ref + nbase : ~leafoff[~ref]; };
            for (size_t k = 0; k < T.nodes.size(); k++) {
                GlomeBihNode b = T.nodes[k];
                b.left = xlate(b.left);
                b.right = xlate(b.right);
                f_bih.push_back(b);
            }
            int bboff = alloc_d(6, 2);
            put_bb(f_dpool, bboff, T.bb);
            nd.a = xlate(T.root); nd.b = bboff; nd.c = all_spheres ? (GLOME_BIH_LINEAR_SPHERES | (first << 4)) : 0;
            f_bih_memo[it.ia] = (int32_t)f_ipool.size();
            f_ipool.push_back(nd.a); f_ipool.push_back(nd.b); f_ipool.push_back(nd.c);
            break;
        }
        case GLOME_MESH: {
            if (f_mesh_memo[it.ia] >= 0) { nd.a = f_mesh_memo[it.ia]; break; }
            const MeshData& M = meshes[it.ia];
            GlomeMeshHeader h;
            memset(&h, 0, sizeof(h));
            h.ntris = (int)(M.tris.size() / 8); h.nverts = (int)(M.verts.size() / 3); h.nnorms = (int)(M.norms.size() / 3);
            h.ntexs = (int)M.texs.size(); h.ntags = (int)M.tags.size();
            h.bb_off = alloc_d(6, 2);
            put_bb(f_dpool, h.bb_off, M.tree.bb);
            h.verts_off = alloc_d((int)M.verts.size(), 2);
            memcpy(&f_dpool[h.verts_off], M.verts.data(), sizeof(double) * M.verts.size());
            h.norms_off = alloc_d((int)M.norms.size(), 2);
            if (!M.norms.empty()) memcpy(&f_dpool[h.norms_off], M.norms.data(), sizeof(double) * M.norms.size());
            while (f_ipool.size() % 8) f_ipool.push_back(0);  // 32-byte aligned Tri records
            h.tris_off = (int)f_ipool.size();
            f_ipool.insert(f_ipool.end(), M.tris.begin(), M.tris.end());
            h.texs_off = (int)f_ipool.size();
            f_ipool.insert(f_ipool.end(), M.texs.begin(), M.texs.end());
            h.tags_off = (int)f_ipool.size();
            f_ipool.insert(f_ipool.end(), M.tags.begin(), M.tags.end());
            int lbase = (int)f_ipool.size();
            f_ipool.insert(f_ipool.end(), M.tree.leafpool.begin(), M.tree.leafpool.end());
            int nbase = (int)f_bvh.size();
            auto xlate = [&](int32_t ref) -> int32_t { return ref >= 0 ? ref + nbase : ~(lbase + M.tree.leafoff[~ref]); };
            for (size_t k = 0; k < M.tree.nodes.size(); k++) {
                GlomeBvhNode b = M.tree.nodes[k];
                b.left = xlate(b.left);
                b.right = xlate(b.right);
                f_bvh.push_back(b);
            }
            h.root = xlate(M.tree.root);
            while (f_ipool.size() % 4) f_ipool.push_back(0);
            int hoff = (int)f_ipool.size();
            const int32_t* hp = (const int32_t*)&h;
            f_ipool.insert(f_ipool.end(), hp, hp + sizeof(h) / 4);
            f_mesh_memo[it.ia] = hoff;
            nd.a = hoff;
            break;
        }
        case GLOME_DIFFERENCE:
        case GLOME_BOUND:
        case GLOME_INNERBOUND: {
            int c = alloc_nodes(2);
            nd.a = c; nd.b = c + 1; nd.c = it.ia;
            flatten_into(it.kids[0], c, depth + 1);
            flatten_into(it.kids[1], c + 1, depth + 1);
            break;
        }
        case GLOME_TEX:
        case GLOME_TAG:
        case GLOME_NOSHADOW:
        case GLOME_ONLYSHADOW: {
            int c = alloc_nodes(1);
            nd.a = c; nd.b = it.ia;
            flatten_into(it.kids[0], c, depth + 1);
            break;
        }
        default: throw BuildError("flatten: unknown item type");
    }
    f_nodes[slot] = nd;
}

// GLOME_CLASS_FLAT: {Tex,Tag}* over prim | Bih[{Tex,Tag}* prim] | Mesh | Group of those (one level)
bool Builder::flat_class(int item, int level) const {
    const Item* it = &items[item];
    while (it->type == GLOME_TEX || it->type == GLOME_TAG) it = &items[it->kids[0]];
    switch (it->type) {
        case GLOME_VOID: case GLOME_SPHERE: case GLOME_TRIANGLE: case GLOME_TRIANGLENORM: case GLOME_BOX:
        case GLOME_PLANE: case GLOME_DISC: case GLOME_CYLINDER: case GLOME_CONE: return true;
        case GLOME_MESH: return level <= 1;
        case GLOME_GROUP:
            if (level != 0) return false;
            for (size_t i = 0; i < it->kids.size(); i++)
                if (!flat_class(it->kids[i], 1)) return false;
            return true;
        case GLOME_BIH:
            if (level > 1) return false;
            for (size_t i = 0; i < it->kids.size(); i++)
                if (!flat_class(it->kids[i], 2)) return false;
            return true;
        default: return false;
    }
}

void Builder::flatten(int root, GlomeFlatScene* out) {
    check(root);
    f_nodes.clear(); f_bih.clear(); f_bvh.clear(); f_ipool.clear(); f_dpool.clear(); f_mats.clear(); f_lightsets.clear();
    f_bih_memo.assign(bihs.size(), -1);
    f_mesh_memo.assign(meshes.size(), -1);
    f_maxdepth = 0;
    int rslot = alloc_nodes(1);
    flatten_into(root, rslot, 1);
    // materials: translate Warp item ids to node indices, Additive lists into ipool
    f_mats = materials;
    bool all_surface = true;
    for (size_t i = 0; i < f_mats.size(); i++) {
        GlomeMaterial& m = f_mats[i];
        if (m.kind != GLOME_MAT_SURFACE && m.kind != GLOME_MAT_BLEND) all_surface = false;
        if (m.kind == GLOME_MAT_WARP) {
            int fs = alloc_nodes(1);
            flatten_into(m.a, fs, 1);
            int ss;
            if (m.b == root) ss = rslot;
            else { ss = alloc_nodes(1); flatten_into(m.b, ss, 1); }
            int xo = alloc_d(24, 2);
            memcpy(&f_dpool[xo], warp_xfms[m.d].m, sizeof(double) * 24);
            m.a = fs; m.b = ss; m.d = xo;
        } else if (m.kind == GLOME_MAT_ADDITIVE) {
            int off = (int)f_ipool.size();
            for (int k = 0; k < m.b; k++) f_ipool.push_back(mat_lists[m.a + k]);
            m.a = off;
        }
    }
    f_lightsets = lightsets;
    if (f_lightsets.empty()) { f_lightsets.push_back(0); f_lightsets.push_back((int)lights.size()); }
    if (f_ipool.empty()) f_ipool.push_back(0);
    if (f_dpool.empty()) f_dpool.push_back(0);
    memset(out, 0, sizeof(*out));
    out->version = GLOME_FLAT_VERSION;
    out->root = rslot;
    out->n_nodes = (int)f_nodes.size(); out->nodes = f_nodes.data();
    out->n_bihnodes = (int)f_bih.size(); out->bihnodes = f_bih.data();
    out->n_bvhnodes = (int)f_bvh.size(); out->bvhnodes = f_bvh.data();
    out->n_ipool = (int)f_ipool.size(); out->ipool = f_ipool.data();
    out->n_dpool = (int64_t)f_dpool.size(); out->dpool = f_dpool.data();
    out->n_textures = (int)textures.size(); out->textures = textures.data();
    out->n_materials = (int)f_mats.size(); out->materials = f_mats.data();
    out->n_lights = (int)lights.size(); out->lights = lights.data();
    out->n_lightsets = (int)(f_lightsets.size() / 2); out->lightsets = f_lightsets.data();
    out->max_depth = f_maxdepth;
    out->scene_class = (all_surface && flat_class(root, 0)) ? GLOME_CLASS_FLAT : GLOME_CLASS_GENERAL;
}

}  // namespace glome_host
