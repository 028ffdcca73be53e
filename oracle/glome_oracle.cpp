// glome_oracle.cpp -- CPU ORACLE.  TEST INFRASTRUCTURE ONLY.
//
// A literal C++ restatement of GlomeTrace's ray-cast path (jimsnow/glome, Haskell), operating on
// the FlatScene arrays of include/glome_cuda.h.  Only tests/, __graft_entry__.smoke() and
// bench.py's cpu_baseline / --impl reference legs may load this library; the product
// (glome_b200/, libglomecuda.so) never links, imports or calls it.
//
// PARITY PINNING: the reference ships no tests, golden vectors or fixtures and no
// Haskell toolchain exists in this image, so the reference itself cannot be run here.  This
// oracle is pinned by (i) line-by-line transcription, every function citing the file:line it
// follows, and (ii) the hand-derived known-answer tests of SURVEY.md Appendix B
// (tests/test_oracle_kat.py).  Status: "parity unpinned" against a running reference.
//
// Floating point: Flt = Double (Vec.hs:9); GHC emits no FMA, so this file must be compiled with
// -ffp-contract=off and without -ffast-math.  fmin/fmax are the reference's `>` chains, not IEEE
// min/max (Vec.hs:44-69); Haskell's Prelude max/min on Double are `if x <= y then y else x` /
// `if x <= y then x else y`.
#include <algorithm>
#include <atomic>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <thread>
#include <utility>
#include <vector>

#include "../include/glome_cuda.h"

namespace orc {

typedef double Flt;
static const Flt infinity_ = 1000000.0;  // Vec.hs:12-14 (NOT IEEE inf)
static const Flt delta = 0.0001;         // Vec.hs:40

// ---- Vec.hs:44-82 ---------------------------------------------------------------------------
static inline Flt fmin_(Flt a, Flt b) { return a > b ? b : a; }  // Vec.hs:44-45
static inline Flt fmax_(Flt a, Flt b) { return a > b ? a : b; }  // Vec.hs:48-49
static inline Flt fmin3(Flt a, Flt b, Flt c) {                   // Vec.hs:52-59
    if (a > b) { if (b > c) return c; else return b; }
    else { if (a > c) return c; else return a; }
}
static inline Flt fmax3(Flt a, Flt b, Flt c) {                   // Vec.hs:62-69
    if (a > b) { if (a > c) return a; else return c; }
    else { if (b > c) return b; else return c; }
}
static inline Flt fabs_(Flt a) { return a < 0 ? -a : a; }        // Vec.hs:80-82
static inline Flt hmax(Flt x, Flt y) { return x <= y ? y : x; }  // Prelude max (GHC.Classes default)
static inline Flt hmin(Flt x, Flt y) { return x <= y ? x : y; }  // Prelude min

struct Vec { Flt x, y, z; };
struct Ray { Vec o, d; };
struct Bbox { Vec p1, p2; };

static inline Vec vec(Flt x, Flt y, Flt z) { Vec v = {x, y, z}; return v; }
static inline Flt va(const Vec& v, int n) { return n == 0 ? v.x : (n == 1 ? v.y : v.z); }  // Vec.hs:167-172
static inline Flt vdot(const Vec& a, const Vec& b) { return (a.x * b.x) + (a.y * b.y) + (a.z * b.z); }  // Vec.hs:185-187
static inline Vec vcross(const Vec& a, const Vec& b) {  // Vec.hs:193-198
    return vec((a.y * b.z) - (a.z * b.y), (a.z * b.x) - (a.x * b.z), (a.x * b.y) - (a.y * b.x));
}
static inline Vec vinvert(const Vec& a) { return vec(-a.x, -a.y, -a.z); }                   // Vec.hs:213
static inline Flt vlen(const Vec& a) { return std::sqrt(vdot(a, a)); }                      // Vec.hs:222
static inline Vec vadd(const Vec& a, const Vec& b) { return vec(a.x + b.x, a.y + b.y, a.z + b.z); }  // Vec.hs:226
static inline Vec vadd3(const Vec& a, const Vec& b, const Vec& c) {                         // Vec.hs:233
    return vec(a.x + b.x + c.x, a.y + b.y + c.y, a.z + b.z + c.z);
}
static inline Vec vsub(const Vec& a, const Vec& b) { return vec(a.x - b.x, a.y - b.y, a.z - b.z); }  // Vec.hs:240
static inline Vec vscale(const Vec& a, Flt f) { return vec(a.x * f, a.y * f, a.z * f); }   // Vec.hs:294
static inline Vec vscaleadd(const Vec& a, const Vec& b, Flt f) {                            // Vec.hs:302
    return vec(a.x + (b.x * f), a.y + (b.y * f), a.z + (b.z * f));
}
static inline Vec vnorm(const Vec& a) {                                                     // Vec.hs:314-317
    Flt invlen = 1.0 / std::sqrt((a.x * a.x) + (a.y * a.y) + (a.z * a.z));
    return vec(a.x * invlen, a.y * invlen, a.z * invlen);
}
static inline Vec bisect(const Vec& a, const Vec& b) { return vnorm(vadd(a, b)); }          // Vec.hs:331
static inline Vec reflect(const Vec& v, const Vec& n) { return vscaleadd(v, n, (-2) * vdot(v, n)); }  // Vec.hs:340
static inline Vec vrcp(const Vec& a) { return vec(1 / a.x, 1 / a.y, 1 / a.z); }             // Vec.hs:345
static inline Ray ray_move(const Ray& r, Flt d) { Ray q = {vscaleadd(r.o, r.d, d), r.d}; return q; }  // Vec.hs:361
static inline Flt plane_int_dist(const Ray& r, const Vec& p, const Vec& norm) {             // Vec.hs:391-394
    Vec newo = vsub(r.o, p);
    return -(vdot(norm, newo)) / (vdot(norm, r.d));
}

// Xfm = 24 doubles: forward Matrix (12, row major 3x4) then inverse Matrix (12)  (Vec.hs:407-414)
static inline Vec xfm_point(const Flt* m, const Vec& v) {  // Vec.hs:502-509
    return vec(m[0] * v.x + m[1] * v.y + m[2] * v.z + m[3], m[4] * v.x + m[5] * v.y + m[6] * v.z + m[7],
               m[8] * v.x + m[9] * v.y + m[10] * v.z + m[11]);
}
static inline Vec invxfm_point(const Flt* m, const Vec& v) { return xfm_point(m + 12, v); }  // Vec.hs:512-519
static inline Vec xfm_vec(const Flt* m, const Vec& v) {  // Vec.hs:522-529
    return vec(m[0] * v.x + m[1] * v.y + m[2] * v.z, m[4] * v.x + m[5] * v.y + m[6] * v.z,
               m[8] * v.x + m[9] * v.y + m[10] * v.z);
}
static inline Vec invxfm_vec(const Flt* m, const Vec& v) { return xfm_vec(m + 12, v); }  // Vec.hs:532-539
static inline Vec invxfm_norm(const Flt* m, const Vec& v) {  // Vec.hs:543-550 (inverse transpose)
    const Flt* i = m + 12;
    return vec(i[0] * v.x + i[4] * v.y + i[8] * v.z, i[1] * v.x + i[5] * v.y + i[9] * v.z,
               i[2] * v.x + i[6] * v.y + i[10] * v.z);
}
static inline Ray xfm_ray(const Flt* m, const Ray& r) {  // Vec.hs:553-555
    Ray q = {xfm_point(m, r.o), vnorm(xfm_vec(m, r.d))};
    return q;
}

static inline Bbox bbjoin(const Bbox& a, const Bbox& b) {  // Vec.hs:652-654
    Bbox r = {vec(fmin_(a.p1.x, b.p1.x), fmin_(a.p1.y, b.p1.y), fmin_(a.p1.z, b.p1.z)),
              vec(fmax_(a.p2.x, b.p2.x), fmax_(a.p2.y, b.p2.y), fmax_(a.p2.z, b.p2.z))};
    return r;
}
static inline Flt bbsa(const Bbox& b) {  // Vec.hs:694-697
    Vec d = vsub(b.p2, b.p1);
    return hmax(0, 2 * (d.x * d.y + d.x * d.z + d.y * d.z));
}
static inline Bbox empty_bbox() {  // Vec.hs:706-709
    Bbox b = {vec(infinity_, infinity_, infinity_), vec(-infinity_, -infinity_, -infinity_)};
    return b;
}
static inline Vec bbmid(const Bbox& b) { return vscale(vadd(b.p1, b.p2), 0.5); }  // Bih.hs:162, Mesh.hs:48

// bbclip_ub_rcp (Vec.hs:725-741): r.d holds the reciprocal direction
static inline void bbclip_ub_rcp(const Vec& o, const Vec& rcp, const Bbox& b, Flt& near_, Flt& far_) {
    Flt inx, outx, iny, outy, inz, outz;
    if (rcp.x > 0) { inx = (b.p1.x - o.x) * rcp.x; outx = (b.p2.x - o.x) * rcp.x; }
    else { inx = (b.p2.x - o.x) * rcp.x; outx = (b.p1.x - o.x) * rcp.x; }
    if (rcp.y > 0) { iny = (b.p1.y - o.y) * rcp.y; outy = (b.p2.y - o.y) * rcp.y; }
    else { iny = (b.p2.y - o.y) * rcp.y; outy = (b.p1.y - o.y) * rcp.y; }
    if (rcp.z > 0) { inz = (b.p1.z - o.z) * rcp.z; outz = (b.p2.z - o.z) * rcp.z; }
    else { inz = (b.p2.z - o.z) * rcp.z; outz = (b.p1.z - o.z) * rcp.z; }
    near_ = fmax3(inx, iny, inz);
    far_ = fmin3(outx, outy, outz);
}
// bbclip_ub (Vec.hs:743-762): branches on the sign of d, not of 1/d (the +0.0 trap)
static inline void bbclip_ub(const Ray& r, const Bbox& b, Flt& near_, Flt& far_) {
    Flt dxrcp = 1 / r.d.x, dyrcp = 1 / r.d.y, dzrcp = 1 / r.d.z;
    Flt inx, outx, iny, outy, inz, outz;
    if (r.d.x > 0) { inx = (b.p1.x - r.o.x) * dxrcp; outx = (b.p2.x - r.o.x) * dxrcp; }
    else { inx = (b.p2.x - r.o.x) * dxrcp; outx = (b.p1.x - r.o.x) * dxrcp; }
    if (r.d.y > 0) { iny = (b.p1.y - r.o.y) * dyrcp; outy = (b.p2.y - r.o.y) * dyrcp; }
    else { iny = (b.p2.y - r.o.y) * dyrcp; outy = (b.p1.y - r.o.y) * dyrcp; }
    if (r.d.z > 0) { inz = (b.p1.z - r.o.z) * dzrcp; outz = (b.p2.z - r.o.z) * dzrcp; }
    else { inz = (b.p2.z - r.o.z) * dzrcp; outz = (b.p1.z - r.o.z) * dzrcp; }
    near_ = fmax3(inx, iny, inz);
    far_ = fmin3(outx, outy, outz);
}

// ---- lists (Haskell [a]) as small head-first arrays -----------------------------------------
struct List {
    int n;
    int v[16];
    List() : n(0) {}
};
static inline List cons(int x, const List& l) {  // x : l
    List r;
    r.n = l.n + 1;
    if (r.n > 16) r.n = 16;
    r.v[0] = x;
    for (int i = 1; i < r.n; i++) r.v[i] = l.v[i - 1];
    return r;
}
static inline List append(const List& a, const List& b) {  // a ++ b
    List r = a;
    for (int i = 0; i < b.n && r.n < 16; i++) r.v[r.n++] = b.v[i];
    return r;
}

// ---- Rayint (Solid.hs:20-34); riuvw is always vzero and is dropped ----------------------------
struct Rayint {
    bool hit;
    Flt depth;
    Vec pos, norm;
    Ray ray;
    List tex, tag;
    int prim, sub, flags;
};
static inline Rayint raymiss() {
    Rayint r;
    r.hit = false; r.depth = infinity_; r.pos = vec(0, 0, 0); r.norm = vec(0, 0, 0);
    r.ray.o = vec(0, 0, 0); r.ray.d = vec(0, 0, 0); r.prim = -1; r.sub = -1; r.flags = 0;
    return r;
}
static inline Flt ridepth(const Rayint& r) { return r.hit ? r.depth : infinity_; }  // Solid.hs:33-34
static inline Rayint rayhit(Flt d, const Vec& p, const Vec& n, const Ray& ray, const List& t, const List& tags,
                            int prim, int sub) {
    Rayint r;
    r.hit = true; r.depth = d; r.pos = p; r.norm = n; r.ray = ray; r.tex = t; r.tag = tags;
    r.prim = prim; r.sub = sub; r.flags = 0;
    return r;
}
// nearest (Solid.hs:37-44): ties go to the SECOND operand
static inline Rayint nearest(const Rayint& a, const Rayint& b) {
    if (!b.hit) { Rayint r = a; r.flags |= b.flags; return r; }
    if (!a.hit) { Rayint r = b; r.flags |= a.flags; return r; }
    if (a.depth < b.depth) { Rayint r = a; r.flags |= b.flags; return r; }
    Rayint r = b; r.flags |= a.flags; return r;
}

// ---- visit counters: +1 per BIH branch entered follows rayint_debug_bih (Bih.hs:389-410) ------
struct Stats {
    int64_t bih_branch, bvh_branch, bih_leaf_items;
    int64_t prim[GLOME_NODE_TYPE_COUNT];
    int64_t tri, trinorm, instance;
    int64_t rays_primary, rays_shadow, rays_secondary, overflow, perlin_range;
    void clear() { memset(this, 0, sizeof(*this)); }
    void add(const Stats& o) {
        int64_t* a = (int64_t*)this; const int64_t* b = (const int64_t*)&o;
        for (size_t i = 0; i < sizeof(Stats) / sizeof(int64_t); i++) a[i] += b[i];
    }
};
static thread_local Stats tl_stats;

struct Scene {
    int root;
    std::vector<GlomeNode> nodes;
    std::vector<GlomeBihNode> bih;
    std::vector<GlomeBvhNode> bvh;
    std::vector<int32_t> ipool;
    std::vector<double> dpool;
    std::vector<GlomeTexture> textures;
    std::vector<GlomeMaterial> materials;
    std::vector<GlomeLight> lights;
    std::vector<int32_t> lightsets;
    Stats total;
};

static const int CSG_CAP = 100000;  // the reference recursion is unbounded; the oracle flags instead of overflowing

static inline Vec ldv(const Scene& S, int off) { return vec(S.dpool[off], S.dpool[off + 1], S.dpool[off + 2]); }
static inline Bbox ldbb(const Scene& S, int off) { Bbox b = {ldv(S, off), ldv(S, off + 3)}; return b; }

static Rayint rayint(const Scene& S, int ni, const Ray& r, Flt d, const List& t, const List& tags, int csg);
static bool shadow(const Scene& S, int ni, const Ray& r, Flt d, int csg);
static bool inside(const Scene& S, int ni, const Vec& pt);
static void get_metainfo(const Scene& S, int ni, const Vec& v, List& texs, List& tags);

// ---- Sphere.hs --------------------------------------------------------------------------------
static Rayint rayint_sphere(const Scene& S, int ni, const Ray& ray, Flt dist, const List& t, const List& tags) {
    // Sphere.hs:20-41
    int a = S.nodes[ni].a;
    Vec center = ldv(S, a); Flt r = S.dpool[a + 3];
    Vec eo = vsub(center, ray.o);
    Flt v = vdot(eo, ray.d);
    Flt vsqr = v * v;
    Flt csqr = vdot(eo, eo);
    Flt rsqr = r * r;
    Flt disc = rsqr - (csqr - vsqr);
    if (disc < 0.0) return raymiss();
    Flt dd = std::sqrt(disc);
    Flt hitdist = ((v - dd) > 0) ? (v - dd) : (v + dd);
    if ((hitdist < 0) || (hitdist > dist)) return raymiss();
    Vec p = vscaleadd(ray.o, ray.d, hitdist);
    Vec n = vnorm(vsub(p, center));
    return rayhit(hitdist, p, n, ray, t, tags, ni, -1);
}
static bool shadow_sphere(const Scene& S, int ni, const Ray& ray, Flt dist) {
    // Sphere.hs:51-71
    int a = S.nodes[ni].a;
    Vec center = ldv(S, a); Flt r = S.dpool[a + 3];
    Vec eo = vsub(center, ray.o);
    Flt v = vdot(eo, ray.d);
    if ((dist >= (v - r)) && (v > 0.0)) {
        Flt vsqr = v * v;
        Flt csqr = vdot(eo, eo);
        Flt rsqr = r * r;
        Flt disc = rsqr - (csqr - vsqr);
        if (disc < 0.0) return false;
        Flt dd = std::sqrt(disc);
        Flt hitdist = ((v - dd) > 0) ? (v - dd) : (v + dd);
        if ((hitdist < 0) || (hitdist > dist)) return false;
        return true;
    }
    return false;
}
static bool inside_sphere(const Scene& S, int ni, const Vec& pt) {  // Sphere.hs:73-76
    int a = S.nodes[ni].a;
    Vec center = ldv(S, a); Flt r = S.dpool[a + 3];
    Vec offset = vsub(center, pt);
    return vdot(offset, offset) < r * r;
}

// ---- Triangle.hs ------------------------------------------------------------------------------
static Rayint rayint_triangle_v(const Vec& p1, const Vec& p2, const Vec& p3, const Ray& ray, Flt dist, const List& tex,
                                const List& tags, int prim, int sub) {
    // Triangle.hs:45-73
    Vec e1 = vsub(p2, p1);
    Vec e2 = vsub(p3, p1);
    Vec s1 = vcross(ray.d, e2);
    Flt divisor = vdot(s1, e1);
    if (divisor == 0) return raymiss();
    Flt invdivisor = 1.0 / divisor;
    Vec d = vsub(ray.o, p1);
    Flt b1 = vdot(d, s1) * invdivisor;
    if (b1 < 0 || b1 > 1) return raymiss();
    Vec s2 = vcross(d, e1);
    Flt b2 = vdot(ray.d, s2) * invdivisor;
    if (b2 < 0 || b1 + b2 > 1) return raymiss();
    Flt t = vdot(e2, s2) * invdivisor;
    if (t < 0 || t > dist) return raymiss();
    return rayhit(t, vscaleadd(ray.o, ray.d, t), vnorm(vcross(e1, e2)), ray, tex, tags, prim, sub);
}
static bool shadow_triangle_v(const Vec& p1, const Vec& p2, const Vec& p3, const Ray& ray, Flt dist) {
    // Triangle.hs:82-107
    Vec e1 = vsub(p2, p1);
    Vec e2 = vsub(p3, p1);
    Vec s1 = vcross(ray.d, e2);
    Flt divisor = vdot(s1, e1);
    if (divisor == 0) return false;
    Flt invdivisor = 1.0 / divisor;
    Vec d = vsub(ray.o, p1);
    Flt b1 = vdot(d, s1) * invdivisor;
    if ((b1 < 0) || (b1 > 1)) return false;
    Vec s2 = vcross(d, e1);
    Flt b2 = vdot(ray.d, s2) * invdivisor;
    if ((b2 < 0) || (b1 + b2 > 1)) return false;
    Flt t = vdot(e2, s2) * invdivisor;
    return (t >= 0) && (t <= dist);
}
static Rayint rayint_trianglenorm_v(const Vec& p1, const Vec& p2, const Vec& p3, const Vec& n1, const Vec& n2,
                                    const Vec& n3, const Ray& ray, Flt dist, const List& tex, const List& tags,
                                    int prim, int sub) {
    // Triangle.hs:109-141
    Vec e1 = vsub(p2, p1);
    Vec e2 = vsub(p3, p1);
    Vec s1 = vcross(ray.d, e2);
    Flt divisor = vdot(s1, e1);
    if (divisor == 0) return raymiss();
    Flt invdivisor = 1.0 / divisor;
    Vec d = vsub(ray.o, p1);
    Flt b1 = vdot(d, s1) * invdivisor;
    if ((b1 < 0) || (b1 > 1)) return raymiss();
    Vec s2 = vcross(d, e1);
    Flt b2 = vdot(ray.d, s2) * invdivisor;
    if ((b2 < 0) || (b1 + b2 > 1)) return raymiss();
    Flt t = vdot(e2, s2) * invdivisor;
    if ((t < 0) || (t > dist)) return raymiss();
    Vec n1s = vscale(n1, 1 - (b1 + b2));
    Vec n2s = vscale(n2, b1);
    Vec n3s = vscale(n3, b2);
    Vec norm = vnorm(vadd3(n1s, n2s, n3s));
    return rayhit(t, vscaleadd(ray.o, ray.d, t), norm, ray, tex, tags, prim, sub);
}

// ---- Box.hs -----------------------------------------------------------------------------------
static Rayint rayint_box(const Scene& S, int ni, const Ray& r, Flt d, const List& t, const List& tags) {
    // Box.hs:18-54
    Bbox b = ldbb(S, S.nodes[ni].a);
    Flt ox = r.o.x, oy = r.o.y, oz = r.o.z, dx = r.d.x, dy = r.d.y, dz = r.d.z;
    Flt dxrcp = 1 / dx, dyrcp = 1 / dy, dzrcp = 1 / dz;
    Flt inx, outx, iny, outy, inz, outz;
    if (dx > 0) { inx = (b.p1.x - ox) * dxrcp; outx = (b.p2.x - ox) * dxrcp; }
    else { inx = (b.p2.x - ox) * dxrcp; outx = (b.p1.x - ox) * dxrcp; }
    if (dy > 0) { iny = (b.p1.y - oy) * dyrcp; outy = (b.p2.y - oy) * dyrcp; }
    else { iny = (b.p2.y - oy) * dyrcp; outy = (b.p1.y - oy) * dyrcp; }
    if (dz > 0) { inz = (b.p1.z - oz) * dzrcp; outz = (b.p2.z - oz) * dzrcp; }
    else { inz = (b.p2.z - oz) * dzrcp; outz = (b.p1.z - oz) * dzrcp; }
    Flt lastin = fmax3(inx, iny, inz);
    Flt firstout = fmin3(outx, outy, outz);
    if (lastin > firstout || firstout < 0 || lastin > d) return raymiss();
    if (lastin < 0) {  // origin is inside
        Vec n;
        if (outx == firstout) n = (dx > 0) ? vec(1, 0, 0) : vec(-1, 0, 0);
        else if (outy == firstout) n = (dy > 0) ? vec(0, 1, 0) : vec(0, -1, 0);
        else n = (dz > 0) ? vec(0, 0, 1) : vec(0, 0, -1);
        return rayhit(firstout, vscaleadd(r.o, r.d, firstout), n, r, t, tags, ni, -1);
    } else {  // origin is outside
        Vec n;
        if (inx == lastin) n = (dx > 0) ? vec(-1, 0, 0) : vec(1, 0, 0);
        else if (iny == lastin) n = (dy > 0) ? vec(0, -1, 0) : vec(0, 1, 0);
        else n = (dz > 0) ? vec(0, 0, -1) : vec(0, 0, 1);
        return rayhit(lastin, vscaleadd(r.o, r.d, lastin), n, r, t, tags, ni, -1);
    }
}
static bool shadow_box(const Scene& S, int ni, const Ray& r, Flt d) {  // Box.hs:56-62
    Bbox b = ldbb(S, S.nodes[ni].a);
    Flt near_, far_;
    bbclip_ub(r, b, near_, far_);
    if ((near_ > far_) || far_ <= 0 || far_ > d) return false;
    return true;
}
static bool inside_box(const Scene& S, int ni, const Vec& p) {  // Box.hs:64-68
    Bbox b = ldbb(S, S.nodes[ni].a);
    return p.x > b.p1.x && p.x < b.p2.x && p.y > b.p1.y && p.y < b.p2.y && p.z > b.p1.z && p.z < b.p2.z;
}

// ---- Plane.hs ---------------------------------------------------------------------------------
static Rayint rayint_plane(const Scene& S, int ni, const Ray& ray, Flt d, const List& t, const List& tags) {
    // Plane.hs:27-32
    int a = S.nodes[ni].a;
    Vec norm = ldv(S, a); Flt offset = S.dpool[a + 3];
    Flt hit = -((vdot(norm, ray.o) - offset) / vdot(norm, ray.d));
    if (hit < 0 || hit > d) return raymiss();
    return rayhit(hit, vscaleadd(ray.o, ray.d, hit), norm, ray, t, tags, ni, -1);
}
static bool inside_plane(const Scene& S, int ni, const Vec& pt) {  // Plane.hs:34-38
    int a = S.nodes[ni].a;
    Vec norm = ldv(S, a); Flt offset = S.dpool[a + 3];
    Vec onplane = vscale(norm, offset);
    Vec newvec = vsub(onplane, pt);
    return vdot(newvec, norm) > 0;
}

// ---- Cone.hs ----------------------------------------------------------------------------------
static Rayint rayint_disc_v(const Vec& point, const Vec& norm, Flt radius_sqr, const Ray& r, Flt d, const List& t,
                            const List& tags, int prim) {
    // Cone.hs:69-79
    Flt dist = plane_int_dist(r, point, norm);
    if (dist < 0 || dist > d) return raymiss();
    Vec pos = vscaleadd(r.o, r.d, dist);
    Vec offset = vsub(pos, point);
    if (vdot(offset, offset) > radius_sqr) return raymiss();
    return rayhit(dist, pos, norm, r, t, tags, prim, -1);
}
static bool shadow_disc_v(const Vec& point, const Vec& norm, Flt radius_sqr, const Ray& r, Flt d) {
    // Cone.hs:81-91
    Flt dist = plane_int_dist(r, point, norm);
    if (dist < 0 || dist > d) return false;
    Vec pos = vscaleadd(r.o, r.d, dist);
    Vec offset = vsub(pos, point);
    if (vdot(offset, offset) > radius_sqr) return false;
    return true;
}
static Rayint rayint_cylinder(const Scene& S, int ni, const Ray& ray, Flt d, const List& t, const List& tags) {
    // Cone.hs:104-139
    int o = S.nodes[ni].a;
    Flt r = S.dpool[o], h1 = S.dpool[o + 1], h2 = S.dpool[o + 2];
    Flt ox = ray.o.x, oy = ray.o.y, oz = ray.o.z, dx = ray.d.x, dy = ray.d.y, dz = ray.d.z;
    Flt a = dx * dx + dy * dy;
    Flt b = 2 * (dx * ox + dy * oy);
    Flt c = ox * ox + oy * oy - r * r;
    Flt disc = b * b - 4 * a * c;
    if (disc < 0) return raymiss();
    Flt discsqrt = std::sqrt(disc);
    Flt q = (b < 0) ? (b - discsqrt) * (-0.5) : (b + discsqrt) * (-0.5);
    Flt t0p = q / a;
    Flt t1p = c / q;
    Flt t0 = fmin_(t0p, t1p);
    Flt t1 = fmax_(t0p, t1p);
    if (t1 < 0 || t0 > d) return raymiss();
    Flt dist = (t0 < 0) ? t1 : t0;
    if (dist < 0 || dist > d) return raymiss();
    Vec pos = vscaleadd(ray.o, ray.d, dist);
    if (pos.z > h1 && pos.z < h2) return rayhit(dist, pos, vec(pos.x / r, pos.y / r, 0), ray, t, tags, ni, -1);
    if (dz > 0) {
        if (oz < h1) return rayint_disc_v(vec(0, 0, h1), vec(0, 0, -1), r * r, ray, d, t, tags, ni);
        return raymiss();
    } else {
        if (oz > h2) return rayint_disc_v(vec(0, 0, h2), vec(0, 0, 1), r * r, ray, d, t, tags, ni);
        return raymiss();
    }
}
static bool inside_cylinder(const Scene& S, int ni, const Vec& p) {  // Cone.hs:141-143
    int o = S.nodes[ni].a;
    Flt r = S.dpool[o], h1 = S.dpool[o + 1], h2 = S.dpool[o + 2];
    return p.z > h1 && p.z < h2 && p.x * p.x + p.y * p.y < r * r;
}
// shared front half of rayint_cone / shadow_cone (Cone.hs:155-204, 206-245); mode 0 = rayint, 1 = shadow
static Rayint cone_common(const Scene& S, int ni, const Ray& ray, Flt d, const List& t, const List& tags, int mode,
                          bool& sh) {
    int o = S.nodes[ni].a;
    Flt r = S.dpool[o], clip1 = S.dpool[o + 1], clip2 = S.dpool[o + 2], height = S.dpool[o + 3];
    Flt ox = ray.o.x, oy = ray.o.y, oz = ray.o.z, dx = ray.d.x, dy = ray.d.y, dz = ray.d.z;
    sh = false;
    Flt kp = r / height;
    Flt k = kp * kp;
    Flt a = dx * dx + dy * dy - k * dz * dz;
    Flt b = 2 * (dx * ox + dy * oy - k * dz * (oz - height));
    Flt c = ox * ox + oy * oy - k * (oz - height) * (oz - height);
    Flt disc = b * b - 4 * a * c;
    if (disc < 0) return raymiss();
    Flt discsqrt = std::sqrt(disc);
    Flt q = (b < 0) ? (b - discsqrt) * (-0.5) : (b + discsqrt) * (-0.5);
    Flt t0p = q / a;
    Flt t1p = c / q;
    Flt t0 = fmin_(t0p, t1p);
    Flt t1 = fmax_(t0p, t1p);
    if (t1 < 0 || t0 > d) return raymiss();
    Flt dist = (t0 < 0) ? t1 : t0;
    if (dist < 0 || dist > d) return raymiss();
    Vec pos = vscaleadd(ray.o, ray.d, dist);
    if (pos.z > clip1 && pos.z < clip2) {
        if (mode == 1) { sh = true; return raymiss(); }
        Flt invhyp = 1 / std::sqrt(height * height + r * r);
        Flt up = r * invhyp;
        Flt out = height * invhyp;
        Flt r_ = std::sqrt(pos.x * pos.x + pos.y * pos.y);
        Flt correction = out / r_;
        return rayhit(dist, pos, vec(pos.x * correction, pos.y * correction, up), ray, t, tags, ni, -1);
    }
    if (dz > 0) {
        if (oz < clip1) {
            if (mode == 1) { sh = shadow_disc_v(vec(0, 0, clip1), vec(0, 0, -1), r * r, ray, d); return raymiss(); }
            return rayint_disc_v(vec(0, 0, clip1), vec(0, 0, -1), r * r, ray, d, t, tags, ni);
        }
        return raymiss();
    } else {
        if (oz > clip2) {
            Flt r2 = r * (1 - ((clip2 - clip1) / height));
            if (mode == 1) { sh = shadow_disc_v(vec(0, 0, clip2), vec(0, 0, 1), r2 * r2, ray, d); return raymiss(); }
            return rayint_disc_v(vec(0, 0, clip2), vec(0, 0, 1), r2 * r2, ray, d, t, tags, ni);
        }
        return raymiss();
    }
}
static bool inside_cone(const Scene& S, int ni, const Vec& p) {  // Cone.hs:248-251
    int o = S.nodes[ni].a;
    Flt rbase = S.dpool[o], h1 = S.dpool[o + 1], h2 = S.dpool[o + 2], height = S.dpool[o + 3];
    Flt r = rbase * (1 - ((p.z - h1) / height));
    return p.z > h1 && p.z < h2 && p.x * p.x + p.y * p.y < r * r;
}

// ---- list group (Solid.hs:326-339) over a contiguous block of nodes ---------------------------
static Rayint rayint_list(const Scene& S, int first, int count, const Ray& r, Flt d, const List& t, const List& tags,
                          int csg) {
    Rayint acc = raymiss();  // foldl' nearest RayMiss (map ...)
    for (int i = 0; i < count; i++) acc = nearest(acc, rayint(S, first + i, r, d, t, tags, csg));
    return acc;
}
static bool shadow_list(const Scene& S, int first, int count, const Ray& r, Flt d, int csg) {
    // foldl' (||) False (map ...): strict fold over a lazily mapped list; (||) short-circuits on its
    // LEFT operand only, so once acc is True later elements are not evaluated.
    bool acc = false;
    for (int i = 0; i < count; i++) acc = acc || shadow(S, first + i, r, d, csg);
    return acc;
}
static bool inside_list(const Scene& S, int first, int count, const Vec& pt) {
    bool acc = false;
    for (int i = 0; i < count; i++) acc = acc || inside(S, first + i, pt);
    return acc;
}

// ---- Bih.hs -----------------------------------------------------------------------------------
struct BihCtx {
    const Scene* S; const Ray* r; Flt d; const List* t; const List* tags; int csg;
    Flt dirr[3]; Flt o[3];
};
static Rayint bih_traverse(const BihCtx& c, int ref, Flt near_, Flt far_) {
    // Bih.hs:339-366
    const Scene& S = *c.S;
    if (ref < 0) {  // BihLeaf s -> rayint s r far t tags   (max distance is the CLIPPED far)
        int32_t lf, lc;
        glome_bih_leaf(ref, S.ipool.data(), &lf, &lc);
        tl_stats.bih_leaf_items += lc;
        return rayint_list(S, lf, lc, *c.r, far_, *c.t, *c.tags, c.csg);
    }
    const GlomeBihNode& n = S.bih[ref];
    tl_stats.bih_branch++;
    Flt dirr = c.dirr[n.axis], o = c.o[n.axis];
    Flt dl = (n.lsplit - o) * dirr;
    Flt dr = (n.rsplit - o) * dirr;
    if (near_ > far_) return raymiss();
    if (dirr > 0) {
        Rayint a = (near_ < dl) ? bih_traverse(c, n.left, near_, fmin_(dl, far_)) : raymiss();
        Rayint b = (dr < far_) ? bih_traverse(c, n.right, fmax_(dr, near_), far_) : raymiss();
        return nearest(a, b);
    } else {
        Rayint a = (near_ < dr) ? bih_traverse(c, n.right, near_, fmin_(dr, far_)) : raymiss();
        Rayint b = (dl < far_) ? bih_traverse(c, n.left, fmax_(dl, near_), far_) : raymiss();
        return nearest(a, b);
    }
}
static Rayint rayint_bih(const Scene& S, int ni, const Ray& r, Flt d, const List& t, const List& tags, int csg) {
    // Bih.hs:332-368
    const GlomeNode& nd = S.nodes[ni];
    Bbox bb = ldbb(S, nd.b);
    Flt near_, far_;
    bbclip_ub(r, bb, near_, far_);
    BihCtx c;
    c.S = &S; c.r = &r; c.d = d; c.t = &t; c.tags = &tags; c.csg = csg;
    c.dirr[0] = 1 / r.d.x; c.dirr[1] = 1 / r.d.y; c.dirr[2] = 1 / r.d.z;
    c.o[0] = r.o.x; c.o[1] = r.o.y; c.o[2] = r.o.z;
    return bih_traverse(c, nd.a, near_, fmin_(d, far_));
}
static bool bih_shadow_traverse(const BihCtx& c, int ref, Flt near_, Flt far_) {
    // Bih.hs:515-542
    const Scene& S = *c.S;
    if (ref < 0) {  // shadow s r (fmin d far)
        int32_t lf, lc;
        glome_bih_leaf(ref, S.ipool.data(), &lf, &lc);
        tl_stats.bih_leaf_items += lc;
        return shadow_list(S, lf, lc, *c.r, fmin_(c.d, far_), c.csg);
    }
    const GlomeBihNode& n = S.bih[ref];
    tl_stats.bih_branch++;
    Flt dirr = c.dirr[n.axis], o = c.o[n.axis];
    Flt dl = (n.lsplit - o) * dirr;
    Flt dr = (n.rsplit - o) * dirr;
    if (near_ > far_) return false;
    if (dirr > 0) {
        return ((near_ < dl) ? bih_shadow_traverse(c, n.left, near_, fmin_(dl, far_)) : false) ||
               ((dr < far_) ? bih_shadow_traverse(c, n.right, fmax_(dr, near_), far_) : false);
    } else {
        return ((near_ < dr) ? bih_shadow_traverse(c, n.right, near_, fmin_(dr, far_)) : false) ||
               ((dl < far_) ? bih_shadow_traverse(c, n.left, fmax_(dl, near_), far_) : false);
    }
}
static bool shadow_bih(const Scene& S, int ni, const Ray& r, Flt d, int csg) {
    // Bih.hs:510-544
    const GlomeNode& nd = S.nodes[ni];
    Bbox bb = ldbb(S, nd.b);
    Flt near_, farp;
    bbclip_ub(r, bb, near_, farp);
    Flt far_ = fmin_(d, farp);
    BihCtx c;
    c.S = &S; c.r = &r; c.d = d; c.t = 0; c.tags = 0; c.csg = csg;
    c.dirr[0] = 1 / r.d.x; c.dirr[1] = 1 / r.d.y; c.dirr[2] = 1 / r.d.z;
    c.o[0] = r.o.x; c.o[1] = r.o.y; c.o[2] = r.o.z;
    return bih_shadow_traverse(c, nd.a, near_, far_);
}
static bool bih_inside_traverse(const Scene& S, int ref, const Vec& pt) {
    // Bih.hs:552-561
    if (ref < 0) { int32_t lf, lc; glome_bih_leaf(ref, S.ipool.data(), &lf, &lc); return inside_list(S, lf, lc, pt); }
    const GlomeBihNode& n = S.bih[ref];
    Flt o = va(pt, n.axis);
    return ((o < n.lsplit) ? bih_inside_traverse(S, n.left, pt) : false) ||
           ((o > n.rsplit) ? bih_inside_traverse(S, n.right, pt) : false);
}
static bool inside_bih(const Scene& S, int ni, const Vec& pt) {
    // Bih.hs:550-565
    const GlomeNode& nd = S.nodes[ni];
    Bbox bb = ldbb(S, nd.b);
    return (pt.x > bb.p1.x) && (pt.x < bb.p2.x) && (pt.y > bb.p1.y) && (pt.y < bb.p2.y) && (pt.z > bb.p1.z) &&
           (pt.z < bb.p2.z) && bih_inside_traverse(S, nd.a, pt);
}
static void metainfo_list(const Scene& S, int first, int count, const Vec& v, List& texs, List& tags);
static void bih_metainfo_traverse(const Scene& S, int ref, const Vec& pt, List& texs, List& tags) {
    // Bih.hs:568-577
    if (ref < 0) { int32_t lf, lc; glome_bih_leaf(ref, S.ipool.data(), &lf, &lc); metainfo_list(S, lf, lc, pt, texs, tags); return; }
    const GlomeBihNode& n = S.bih[ref];
    Flt o = va(pt, n.axis);
    List lt, lg, rt, rg;
    if (o < n.lsplit) bih_metainfo_traverse(S, n.left, pt, lt, lg);
    if (o > n.rsplit) bih_metainfo_traverse(S, n.right, pt, rt, rg);
    texs = append(lt, rt);  // paircat
    tags = append(lg, rg);
}

// ---- Mesh.hs ----------------------------------------------------------------------------------
struct MeshCtx {
    const Scene* S; const GlomeMeshHeader* h; const Ray* ray; Vec rcp; Flt depth; const List* texs; const List* tags;
    int ni;
};
static Rayint mesh_rayint_tri(const MeshCtx& c, int i, Flt far_) {
    // Mesh.hs:143-161
    const Scene& S = *c.S;
    const int32_t* T = &S.ipool[c.h->tris_off + 8 * i];
    Vec a = ldv(S, c.h->verts_off + 3 * T[0]);
    Vec b = ldv(S, c.h->verts_off + 3 * T[1]);
    Vec cc = ldv(S, c.h->verts_off + 3 * T[2]);
    List tex = (T[6] == -1) ? *c.texs : cons(S.ipool[c.h->texs_off + T[6]], *c.texs);
    List tag = (T[7] == -1) ? *c.tags : cons(S.ipool[c.h->tags_off + T[7]], *c.tags);
    if (T[3] == -1) {
        tl_stats.tri++;
        return rayint_triangle_v(a, b, cc, *c.ray, far_, tex, tag, c.ni, i);
    }
    tl_stats.trinorm++;
    Vec an = ldv(S, c.h->norms_off + 3 * T[3]);
    Vec bn = ldv(S, c.h->norms_off + 3 * T[4]);
    Vec cn = ldv(S, c.h->norms_off + 3 * T[5]);
    return rayint_trianglenorm_v(a, b, cc, an, bn, cn, *c.ray, far_, tex, tag, c.ni, i);
}
static Rayint mesh_traverse(const MeshCtx& c, int ref, Flt near_, Flt far_) {
    // Mesh.hs:163-196
    const Scene& S = *c.S;
    if (ref < 0) {
        int k = ~ref;
        int cnt = S.ipool[k];
        Rayint acc = raymiss();
        for (int j = 0; j < cnt; j++) acc = nearest(acc, mesh_rayint_tri(c, S.ipool[k + 1 + j], far_));
        return acc;
    }
    const GlomeBvhNode& n = S.bvh[ref];
    tl_stats.bvh_branch++;
    Bbox lbb = {vec(n.lbb[0], n.lbb[1], n.lbb[2]), vec(n.lbb[3], n.lbb[4], n.lbb[5])};
    Bbox rbb = {vec(n.rbb[0], n.rbb[1], n.rbb[2]), vec(n.rbb[3], n.rbb[4], n.rbb[5])};
    Flt lnearp, lfarp, rnearp, rfarp;
    bbclip_ub_rcp(c.ray->o, c.rcp, lbb, lnearp, lfarp);
    bbclip_ub_rcp(c.ray->o, c.rcp, rbb, rnearp, rfarp);
    Flt lnear = hmax(near_, lnearp);
    Flt lfar = hmin(far_, lfarp);
    Flt rnear = hmax(near_, rnearp);
    Flt rfar = hmin(far_, rfarp);
    Flt depth = c.depth;
    if (lnear < rnear) {
        Rayint lresult = (lnear > lfar || lnear > depth || lfar < 0) ? raymiss() : mesh_traverse(c, n.left, lnear, lfar);
        Flt rfar2 = hmin(rfar, ridepth(lresult));
        Rayint second = (rnear > rfar2 || rnear > depth || rfar2 < 0) ? raymiss() : mesh_traverse(c, n.right, rnear, rfar);
        return nearest(lresult, second);
    } else {
        Rayint rresult = (rnear > rfar || rnear > depth || rfar < 0) ? raymiss() : mesh_traverse(c, n.right, rnear, rfar);
        Flt lfar2 = hmin(lfar, ridepth(rresult));
        Rayint second = (lnear > lfar2 || lnear > depth || lfar2 < 0) ? raymiss() : mesh_traverse(c, n.left, lnear, lfar);
        return nearest(rresult, second);
    }
}
static Rayint rayint_mesh(const Scene& S, int ni, const Ray& ray, Flt depth, const List& texs, const List& tags) {
    // Mesh.hs:136-198
    const GlomeMeshHeader* h = (const GlomeMeshHeader*)&S.ipool[S.nodes[ni].a];
    Bbox bb = ldbb(S, h->bb_off);
    Vec rcp = vrcp(ray.d);
    Flt near_, far_;
    bbclip_ub_rcp(ray.o, rcp, bb, near_, far_);
    if (near_ > far_ || near_ > depth || far_ < 0) return raymiss();
    MeshCtx c;
    c.S = &S; c.h = h; c.ray = &ray; c.rcp = rcp; c.depth = depth; c.texs = &texs; c.tags = &tags; c.ni = ni;
    return mesh_traverse(c, h->root, near_, far_);
}

// ---- Solid.hs: Instance -----------------------------------------------------------------------
static Rayint rayint_instance(const Scene& S, int ni, const Ray& r, Flt d, const List& t, const List& tags, int csg) {
    // Solid.hs:388-403
    const GlomeNode& nd = S.nodes[ni];
    const Flt* xfm = &S.dpool[nd.b];
    tl_stats.instance++;
    Vec newdir = invxfm_vec(xfm, r.d);
    Vec neworig = invxfm_point(xfm, r.o);
    Flt lenscale = vlen(newdir);
    Flt invlenscale = 1 / lenscale;
    Ray nr = {neworig, vscale(newdir, invlenscale)};
    Rayint ri = rayint(S, nd.a, nr, d * lenscale, t, tags, csg);
    if (!ri.hit) return ri;
    ri.depth = ri.depth * invlenscale;
    ri.pos = xfm_point(xfm, ri.pos);
    ri.norm = vnorm(invxfm_norm(xfm, ri.norm));
    return ri;  // riray stays the inner ray
}
static bool shadow_instance(const Scene& S, int ni, const Ray& r, Flt d, int csg) {
    // Solid.hs:464-471
    const GlomeNode& nd = S.nodes[ni];
    const Flt* xfm = &S.dpool[nd.b];
    tl_stats.instance++;
    Vec newdir = invxfm_vec(xfm, r.d);
    Vec neworig = invxfm_point(xfm, r.o);
    Flt lenscale = vlen(newdir);
    Flt invlenscale = 1 / lenscale;
    Ray nr = {neworig, vscale(newdir, invlenscale)};
    return shadow(S, nd.a, nr, d * lenscale, csg);
}

// ---- Solid.hs:85-91 rayint_advance ------------------------------------------------------------
static Rayint rayint_advance(const Scene& S, int ni, const Ray& r, Flt d, const List& t, const List& tags, Flt adv,
                             int csg) {
    if (csg >= CSG_CAP) { Rayint m = raymiss(); m.flags |= GLOME_HITFLAG_CSG_OVERFLOW; return m; }
    Flt a = adv + delta;
    Rayint ri = rayint(S, ni, ray_move(r, a), d - a, t, tags, csg + 1);
    if (!ri.hit) return ri;
    ri.depth = ri.depth + a;
    return ri;
}

// ---- Csg.hs -----------------------------------------------------------------------------------
static Rayint rayint_difference(const Scene& S, int ni, const Ray& r, Flt d, const List& t, const List& tags, int csg) {
    // Csg.hs:33-54
    const GlomeNode& nd = S.nodes[ni];
    int sa = nd.a, sb = nd.b;
    bool useatex = nd.c != 0;
    if (inside(S, sb, r.o)) {
        Rayint rib = rayint(S, sb, r, d, t, tags, csg);
        if (!rib.hit) return rib;
        if (inside(S, sa, rib.pos) && !inside(S, sb, vscaleadd(rib.pos, r.d, delta))) {
            if (useatex) {
                List atexs, atags;
                get_metainfo(S, sa, rib.pos, atexs, atags);
                Rayint out = rib;
                out.norm = vinvert(rib.norm);
                out.tex = atexs;
                out.tag = atags;
                return out;
            }
            Rayint out = rib;
            out.norm = vinvert(rib.norm);
            return out;
        }
        return rayint_advance(S, ni, r, d, t, tags, rib.depth, csg);
    }
    Rayint ria = rayint(S, sa, r, d, t, tags, csg);
    if (!ria.hit) return ria;
    Rayint rib = rayint(S, sb, r, d, t, tags, csg);
    if (rib.hit) {
        if (ria.depth < rib.depth) return ria;
        return rayint_advance(S, ni, r, d, t, tags, rib.depth, csg);
    }
    return ria;
}
// Intersection over the suffix [first, first+count) of the child block (Csg.hs:68-90)
static bool inside_isect(const Scene& S, int first, int count, const Vec& pt) {
    // Csg.hs:99-101: foldl' (&&) True (map ...): (&&) short-circuits on a False accumulator
    bool acc = true;
    for (int i = 0; i < count; i++) acc = acc && inside(S, first + i, pt);
    return acc;
}
static Rayint rayint_isect(const Scene& S, int first, int count, const Ray& r, Flt d, const List& t, const List& tags,
                           int csg);
static Rayint isect_advance(const Scene& S, int first, int count, const Ray& r, Flt d, const List& t, const List& tags,
                            Flt adv, int csg) {
    // rayint_advance (SolidItem (Intersection slds)) ...  (Solid.hs:85-91)
    if (csg >= CSG_CAP) { Rayint m = raymiss(); m.flags |= GLOME_HITFLAG_CSG_OVERFLOW; return m; }
    Flt a = adv + delta;
    Rayint ri = rayint_isect(S, first, count, ray_move(r, a), d - a, t, tags, csg + 1);
    if (!ri.hit) return ri;
    ri.depth = ri.depth + a;
    return ri;
}
static Rayint rayint_isect(const Scene& S, int first, int count, const Ray& r, Flt d, const List& t, const List& tags,
                           int csg) {
    if (count == 0 || d < 0) return raymiss();
    int s = first;
    if (count == 1) return rayint(S, s, r, d, t, tags, csg);
    if (inside(S, s, r.o)) {
        Rayint rs = rayint(S, s, r, d, t, tags, csg);
        if (!rs.hit) {
            Rayint q = rayint_isect(S, first + 1, count - 1, r, d, t, tags, csg);
            q.flags |= rs.flags;
            return q;
        }
        Rayint q = rayint_isect(S, first + 1, count - 1, r, rs.depth, t, tags, csg);
        if (!q.hit) return isect_advance(S, first, count, r, d, t, tags, rs.depth, csg);
        return q;
    } else {
        Rayint rs = rayint(S, s, r, d, t, tags, csg);
        if (!rs.hit) return rs;
        if (inside_isect(S, first + 1, count - 1, rs.pos)) {
            Rayint out = rs;
            out.ray = r;  // RayHit sd sp sn r vzero st stags
            return out;
        }
        return isect_advance(S, first, count, r, d, t, tags, rs.depth, csg);
    }
}

// ---- dispatch: class Solid (Solid.hs:138-254) ---------------------------------------------------
static Rayint rayint(const Scene& S, int ni, const Ray& r, Flt d, const List& t, const List& tags, int csg) {
    const GlomeNode& nd = S.nodes[ni];
    tl_stats.prim[nd.type]++;
    switch (nd.type) {
        case GLOME_VOID: return raymiss();  // Solid.hs:354
        case GLOME_SPHERE: return rayint_sphere(S, ni, r, d, t, tags);
        case GLOME_TRIANGLE: {
            int a = nd.a;
            return rayint_triangle_v(ldv(S, a), ldv(S, a + 3), ldv(S, a + 6), r, d, t, tags, ni, -1);
        }
        case GLOME_TRIANGLENORM: {
            int a = nd.a;
            return rayint_trianglenorm_v(ldv(S, a), ldv(S, a + 3), ldv(S, a + 6), ldv(S, a + 9), ldv(S, a + 12),
                                         ldv(S, a + 15), r, d, t, tags, ni, -1);
        }
        case GLOME_BOX: return rayint_box(S, ni, r, d, t, tags);
        case GLOME_PLANE: return rayint_plane(S, ni, r, d, t, tags);
        case GLOME_DISC: {
            int a = nd.a;
            return rayint_disc_v(ldv(S, a), ldv(S, a + 3), S.dpool[a + 6], r, d, t, tags, ni);
        }
        case GLOME_CYLINDER: return rayint_cylinder(S, ni, r, d, t, tags);
        case GLOME_CONE: { bool sh; return cone_common(S, ni, r, d, t, tags, 0, sh); }
        case GLOME_GROUP: return rayint_list(S, nd.a, nd.b, r, d, t, tags, csg);
        case GLOME_INSTANCE: return rayint_instance(S, ni, r, d, t, tags, csg);
        case GLOME_BIH: return rayint_bih(S, ni, r, d, t, tags, csg);
        case GLOME_MESH: return rayint_mesh(S, ni, r, d, t, tags);
        case GLOME_DIFFERENCE: return rayint_difference(S, ni, r, d, t, tags, csg);
        case GLOME_INTERSECTION: return rayint_isect(S, nd.a, nd.b, r, d, t, tags, csg);
        case GLOME_TEX: return rayint(S, nd.a, r, d, cons(nd.b, t), tags, csg);   // Tex.hs:66
        case GLOME_TAG: return rayint(S, nd.a, r, d, t, cons(nd.b, tags), csg);   // Tex.hs:54
        case GLOME_NOSHADOW: return rayint(S, nd.a, r, d, t, tags, csg);          // Tex.hs:78
        case GLOME_ONLYSHADOW: return raymiss();                                  // Tex.hs:89
        case GLOME_BOUND:                                                         // Bound.hs:30-35
            if (inside(S, nd.a, r.o) || shadow(S, nd.a, r, d, csg)) return rayint(S, nd.b, r, d, t, tags, csg);
            return raymiss();
        case GLOME_INNERBOUND: {                                                  // Bound.hs:98-99
            List e1, e2;
            Rayint ra = rayint(S, nd.a, r, d, e1, e2, csg);
            return rayint(S, nd.b, r, ridepth(ra), t, tags, csg);
        }
    }
    return raymiss();
}
static bool shadow_default(const Scene& S, int ni, const Ray& r, Flt d, int csg) {
    // Solid.hs:218-221: fall back on rayint
    List e1, e2;
    return rayint(S, ni, r, d, e1, e2, csg).hit;
}
static bool shadow(const Scene& S, int ni, const Ray& r, Flt d, int csg) {
    const GlomeNode& nd = S.nodes[ni];
    switch (nd.type) {
        case GLOME_VOID: return false;  // Solid.hs:356
        case GLOME_SPHERE: tl_stats.prim[nd.type]++; return shadow_sphere(S, ni, r, d);
        case GLOME_TRIANGLE:
        case GLOME_TRIANGLENORM: {  // Triangle.hs:82-107, 143-145
            tl_stats.prim[nd.type]++;
            int a = nd.a;
            return shadow_triangle_v(ldv(S, a), ldv(S, a + 3), ldv(S, a + 6), r, d);
        }
        case GLOME_BOX: tl_stats.prim[nd.type]++; return shadow_box(S, ni, r, d);
        case GLOME_PLANE: return shadow_default(S, ni, r, d, csg);     // Plane.hs has no shadow
        case GLOME_DISC: {
            tl_stats.prim[nd.type]++;
            int a = nd.a;
            return shadow_disc_v(ldv(S, a), ldv(S, a + 3), S.dpool[a + 6], r, d);
        }
        case GLOME_CYLINDER: return shadow_default(S, ni, r, d, csg);  // Cone.hs:149-152 no shadow
        case GLOME_CONE: {
            tl_stats.prim[nd.type]++;
            bool sh; List e1, e2;
            cone_common(S, ni, r, d, e1, e2, 1, sh);
            return sh;
        }
        case GLOME_GROUP: return shadow_list(S, nd.a, nd.b, r, d, csg);
        case GLOME_INSTANCE: return shadow_instance(S, ni, r, d, csg);
        case GLOME_BIH: return shadow_bih(S, ni, r, d, csg);
        case GLOME_MESH: return false;  // Mesh.hs:210
        case GLOME_DIFFERENCE:
        case GLOME_INTERSECTION: return shadow_default(S, ni, r, d, csg);  // Csg.hs: no shadow
        case GLOME_TEX:
        case GLOME_TAG: return shadow(S, nd.a, r, d, csg);  // Tex.hs:57,69
        case GLOME_NOSHADOW: return false;                  // Tex.hs:81
        case GLOME_ONLYSHADOW: return shadow(S, nd.a, r, d, csg);  // Tex.hs:92
        case GLOME_BOUND:                                   // Bound.hs:44-49
            if (inside(S, nd.a, r.o) || shadow(S, nd.a, r, d, csg)) return shadow(S, nd.b, r, d, csg);
            return false;
        case GLOME_INNERBOUND: return shadow(S, nd.a, r, d, csg) || shadow(S, nd.b, r, d, csg);  // Bound.hs:101-103
    }
    return false;
}
static bool inside(const Scene& S, int ni, const Vec& pt) {
    const GlomeNode& nd = S.nodes[ni];
    switch (nd.type) {
        case GLOME_VOID: return false;
        case GLOME_SPHERE: return inside_sphere(S, ni, pt);
        case GLOME_TRIANGLE:
        case GLOME_TRIANGLENORM: return false;  // Triangle.hs:183,190
        case GLOME_BOX: return inside_box(S, ni, pt);
        case GLOME_PLANE: return inside_plane(S, ni, pt);
        case GLOME_DISC: return false;  // Cone.hs:100
        case GLOME_CYLINDER: return inside_cylinder(S, ni, pt);
        case GLOME_CONE: return inside_cone(S, ni, pt);
        case GLOME_GROUP: return inside_list(S, nd.a, nd.b, pt);
        case GLOME_INSTANCE: return inside(S, nd.a, invxfm_point(&S.dpool[nd.b], pt));  // Solid.hs:473-475
        case GLOME_BIH: return inside_bih(S, ni, pt);
        case GLOME_MESH: return false;  // Mesh.hs:211
        case GLOME_DIFFERENCE: return inside(S, nd.a, pt) && !inside(S, nd.b, pt);  // Csg.hs:92-94
        case GLOME_INTERSECTION: return inside_isect(S, nd.a, nd.b, pt);
        case GLOME_TEX:
        case GLOME_TAG:
        case GLOME_NOSHADOW:
        case GLOME_ONLYSHADOW: return inside(S, nd.a, pt);
        case GLOME_BOUND: return inside(S, nd.a, pt) && inside(S, nd.b, pt);        // Bound.hs:51
        case GLOME_INNERBOUND: return inside(S, nd.a, pt) || inside(S, nd.b, pt);   // Bound.hs:109
    }
    return false;
}
static void metainfo_list(const Scene& S, int first, int count, const Vec& v, List& texs, List& tags) {
    // Solid.hs:337-339: foldl (\acc x -> if inside x v then paircat (get_metainfo x v) acc else acc)
    List at, ag;
    for (int i = 0; i < count; i++) {
        if (inside(S, first + i, v)) {
            List xt, xg;
            get_metainfo(S, first + i, v, xt, xg);
            at = append(xt, at);
            ag = append(xg, ag);
        }
    }
    texs = at;
    tags = ag;
}
static void get_metainfo(const Scene& S, int ni, const Vec& v, List& texs, List& tags) {
    const GlomeNode& nd = S.nodes[ni];
    texs = List(); tags = List();
    switch (nd.type) {
        case GLOME_GROUP: metainfo_list(S, nd.a, nd.b, v, texs, tags); return;
        case GLOME_INSTANCE: get_metainfo(S, nd.a, invxfm_point(&S.dpool[nd.b], v), texs, tags); return;  // Solid.hs:517
        case GLOME_BIH: {  // Bih.hs:567-585
            Bbox bb = ldbb(S, nd.b);
            if ((v.x > bb.p1.x) && (v.x < bb.p2.x) && (v.y > bb.p1.y) && (v.y < bb.p2.y) && (v.z > bb.p1.z) &&
                (v.z < bb.p2.z))
                bih_metainfo_traverse(S, nd.a, v, texs, tags);
            return;
        }
        case GLOME_DIFFERENCE:  // Csg.hs:103-106
            if (inside(S, nd.a, v) && !inside(S, nd.b, v)) get_metainfo(S, nd.a, v, texs, tags);
            return;
        case GLOME_INTERSECTION:  // Csg.hs:108-111
            if (inside_isect(S, nd.a, nd.b, v)) {
                for (int i = 0; i < nd.b; i++) {
                    List xt, xg;
                    get_metainfo(S, nd.a + i, v, xt, xg);
                    texs = append(texs, xt);
                    tags = append(tags, xg);
                }
            }
            return;
        case GLOME_TEX: {  // Tex.hs:73-74
            List xt, xg;
            get_metainfo(S, nd.a, v, xt, xg);
            texs = cons(nd.b, xt); tags = xg;
            return;
        }
        case GLOME_TAG: {  // Tex.hs:61-62
            List xt, xg;
            get_metainfo(S, nd.a, v, xt, xg);
            texs = xt; tags = cons(nd.b, xg);
            return;
        }
        case GLOME_NOSHADOW:
        case GLOME_ONLYSHADOW: get_metainfo(S, nd.a, v, texs, tags); return;
        case GLOME_BOUND:  // Bound.hs:54-58
            if (inside(S, nd.a, v)) get_metainfo(S, nd.b, v, texs, tags);
            return;
        case GLOME_INNERBOUND: get_metainfo(S, nd.b, v, texs, tags); return;  // Bound.hs:112
        default: return;  // Solid.hs:254: ([],[])
    }
}

// ---- Clr.hs -----------------------------------------------------------------------------------
struct Color { Flt r, g, b; };
struct ColorA { Flt r, g, b, a; };
static inline Color cadd(const Color& a, const Color& b) { Color c = {a.r + b.r, a.g + b.g, a.b + b.b}; return c; }  // Clr.hs:23
static inline Color cscale(const Color& a, Flt m) { Color c = {a.r * m, a.g * m, a.b * m}; return c; }               // Clr.hs:44
static inline Flt aclamp(Flt x) { if (x > 1) return 1; if (x < 0) return 0; return x; }                               // Clr.hs:75-79
static inline ColorA caweight(const ColorA& x, const ColorA& y, Flt w) {                                              // Clr.hs:87-91
    ColorA c = {(x.r * w) + (y.r * (1 - w)), (x.g * w) + (y.g * (1 - w)), (x.b * w) + (y.b * (1 - w)),
                (x.a * w) + (y.a * (1 - w))};
    return c;
}
static inline ColorA cafold(const ColorA& x, const ColorA& y) {  // Clr.hs:106-113
    Flt trans = 1 - x.a;
    ColorA c = {x.r + (y.r * trans * y.a), x.g + (y.g * trans * y.a), x.b + (y.b * trans * y.a), x.a + (y.a * trans)};
    return c;
}
static const ColorA ca_transparent = {0, 0, 0, 0};
static const ColorA ca_black = {0, 0, 0, 1};

// ---- Texture.hs -------------------------------------------------------------------------------
static inline Flt triangle_wave(Flt x) {  // Texture.hs:16-21
    Flt offset = x - std::floor(x);
    return (offset < 0.5) ? (offset * 2) : (2 - (offset * 2));
}
static inline Flt omega(Flt t_) {  // Texture.hs:49-54
    Flt t = fabs_(t_);
    Flt tsqr = t * t;
    Flt tcube = tsqr * t;
    return (-6) * tcube * tsqr + 15 * tcube * t - 10 * tcube + 1;
}
static const int phi_[12] = {3, 0, 2, 7, 4, 1, 5, 11, 8, 10, 9, 6};  // Texture.hs:57-58
// grad (Texture.hs:60-65): the 12 vectors of {-1,0,1}^3 with 1.1 < |v| < 1.5, in x-major enumeration order
static const int grad_[12][3] = {{-1, -1, 0}, {-1, 0, -1}, {-1, 0, 1}, {-1, 1, 0}, {0, -1, -1}, {0, -1, 1},
                                 {0, 1, -1},  {0, 1, 1},   {1, -1, 0}, {1, 0, -1}, {1, 0, 1},   {1, 1, 0}};
static inline int64_t iabs64(int64_t a) { return a < 0 ? -a : a; }
static inline Vec gamma_(int64_t i, int64_t j, int64_t k) {  // Texture.hs:67-72
    int a = phi_[iabs64(k) % 12];
    int b = phi_[iabs64(j + a) % 12];
    int c = phi_[iabs64(i + b) % 12];
    return vec(grad_[c][0], grad_[c][1], grad_[c][2]);
}
static inline Flt knot(int64_t i, int64_t j, int64_t k, const Vec& v) {  // Texture.hs:74-77
    return omega(v.x) * omega(v.y) * omega(v.z) * vdot(gamma_(i, j, k), v);
}
static Flt noise(const Vec& p) {  // Texture.hs:92-107
    Flt fx = std::floor(p.x), fy = std::floor(p.y), fz = std::floor(p.z);
    int64_t i = (int64_t)fx, j = (int64_t)fy, k = (int64_t)fz;
    Flt u = p.x - fx, v = p.y - fy, w = p.z - fz;
    return knot(i, j, k, vec(u, v, w)) + knot(i + 1, j, k, vec(u - 1, v, w)) + knot(i, j + 1, k, vec(u, v - 1, w)) +
           knot(i, j, k + 1, vec(u, v, w - 1)) + knot(i + 1, j + 1, k, vec(u - 1, v - 1, w)) +
           knot(i + 1, j, k + 1, vec(u - 1, v, w - 1)) + knot(i, j + 1, k + 1, vec(u, v - 1, w - 1)) +
           knot(i + 1, j + 1, k + 1, vec(u - 1, v - 1, w - 1));
}
static Flt perlin(const Vec& v) {  // Texture.hs:109-116 (`error` outside [0,1] -> counted, not trapped)
    Flt p = (noise(v) + 1) * 0.5;
    if (p > 1 || p < 0) tl_stats.perlin_range++;
    return p;
}

// ---- Trace.hs / Shader.hs ---------------------------------------------------------------------
struct Mat {  // a Material value (Shader.hs:43-52); Blend may carry a texture-computed weight
    int kind, a, b, c, d;
    Flt p[8];
};
static inline Mat mat_from_table(const Scene& S, int id) {
    const GlomeMaterial& m = S.materials[id];
    Mat r;
    r.kind = m.kind; r.a = m.a; r.b = m.b; r.c = m.c; r.d = m.d;
    for (int i = 0; i < 8; i++) r.p[i] = m.p[i];
    return r;
}
struct TraceResult { ColorA c; List tags; Rayint ri; };
struct LightCtx {  // ctxb = [(Color, Vec)], forced lazily (Trace.hs:63); a list: any number of lights
    bool done; int n; std::vector<Color> col; std::vector<Vec> dir;
};

static TraceResult trace(const Scene& S, int lightset, int sld, const Ray& ray, Flt depth, int recurs);

static void mpreshade(const Scene& S, int lightset, const Ray& ray, int scene, const Rayint& ri, LightCtx& ctx) {
    // Shader.hs:65-80
    (void)ray;
    ctx.done = true; ctx.n = 0;
    ctx.col.clear(); ctx.dir.clear();
    if (!ri.hit) return;
    int first = S.lightsets[2 * lightset], cnt = S.lightsets[2 * lightset + 1];
    for (int li = 0; li < cnt; li++) {
        const GlomeLight& L = S.lights[first + li];
        Vec lpos = vec(L.pos[0], L.pos[1], L.pos[2]);
        Vec lvec = vsub(lpos, ri.pos);
        if (vdot(lvec, ri.norm) < 0) continue;
        Flt llen = vlen(lvec);
        Vec ldir = vscale(lvec, 1 / llen);
        bool blocked = llen > L.rad;
        if (!blocked && L.do_shadow) {
            Ray sr = {vscaleadd(ri.pos, ri.norm, delta), ldir};
            tl_stats.rays_shadow++;
            blocked = shadow(S, scene, sr, llen - (2 * delta), 0);
        }
        if (blocked) continue;
        Flt fall = 1 / (llen * llen);  // Shader.hs:23
        Color lc = {L.color[0], L.color[1], L.color[2]};
        ctx.col.push_back(cscale(lc, fall)); ctx.dir.push_back(ldir); ctx.n++;
    }
}

static void mpostshade(const Scene& S, int ls, LightCtx& lights, const Mat& mat, const Ray& ray, int s, const Rayint& ri,
                       int recurs, ColorA& outc, List& outtags) {
    // Shader.hs:82-184
    outtags = List();
    if (!ri.hit) { outc = ca_transparent; return; }
    const Vec& dir = ray.d;
    const Vec& n = ri.norm;
    const Vec& p = ri.pos;
    Vec eyedir = vinvert(dir);
    switch (mat.kind) {
        case GLOME_MAT_SURFACE: {  // Shader.hs:90-105
            if (!lights.done) mpreshade(S, ls, ray, s, ri, lights);
            Color color = {mat.p[0], mat.p[1], mat.p[2]};
            Flt alpha = mat.p[3], amb = mat.p[4], kd = mat.p[5], ks = mat.p[6], shine = mat.p[7];
            Color ambient = cscale(color, amb);
            Color direct = {0, 0, 0};
            for (int i = 0; i < lights.n; i++) {
                Vec ldir = lights.dir[i];
                Vec halfangle = bisect(ldir, eyedir);
                Flt ldotn = fmax_(0, vdot(ldir, n));
                Flt blinn;
                if (ks <= delta) blinn = 0;
                else {
                    Flt b = fmax_(0, std::pow(vdot(halfangle, n), shine) * ldotn);
                    blinn = std::isnan(b) ? 0 : b;
                }
                Flt diffuse = vdot(ldir, n);
                direct = cadd(direct, cscale(lights.col[i], (blinn * ks) + (diffuse * kd)));
            }
            Color c = cadd(ambient, direct);
            outc.r = c.r; outc.g = c.g; outc.b = c.b; outc.a = alpha;
            return;
        }
        case GLOME_MAT_REFLECT: {  // Shader.hs:107-118
            Flt refl = mat.p[0];
            if ((refl > 0) && (recurs > 0)) {
                Vec outdir = reflect(dir, n);
                Ray nr = {vscaleadd(p, outdir, delta), outdir};
                tl_stats.rays_secondary++;
                TraceResult tr = trace(S, ls, s, nr, infinity_, recurs - 1);
                outc.r = tr.c.r; outc.g = tr.c.g; outc.b = tr.c.b; outc.a = tr.c.a * refl;
                outtags = tr.tags;
            } else outc = ca_black;
            return;
        }
        case GLOME_MAT_REFRACT: {  // Shader.hs:120-155
            Flt refl = mat.p[0], refr = mat.p[1], ior = mat.p[2];
            if ((refl > 0 || refr > 0) && (recurs > 0)) {
                Vec outdir = reflect(dir, n);
                Ray nr = {vscaleadd(p, outdir, delta), outdir};
                tl_stats.rays_secondary++;
                TraceResult a = trace(S, ls, s, nr, infinity_, recurs - 1);
                Flt eta = (vdot(n, eyedir) > 0) ? ior : 1 / ior;
                Flt c1 = vdot(dir, n);
                Flt cs2 = 1 - (eta * eta) * (1 - (c1 * c1));
                ColorA rc; List rtags;
                if (cs2 < 0) { rc = ca_black; }
                else {
                    Vec t = vadd(vscale(dir, eta), vscale(n, eta * c1 - std::sqrt(cs2)));
                    Ray tr = {vscaleadd(p, t, delta), t};
                    tl_stats.rays_secondary++;
                    TraceResult b = trace(S, ls, s, tr, infinity_, recurs - 1);
                    rc = b.c; rtags = b.tags;
                }
                outc.r = a.c.r * refl + rc.r * refr;
                outc.g = a.c.g * refl + rc.g * refr;
                outc.b = a.c.b * refl + rc.b * refr;
                outc.a = a.c.a * refl + rc.a * refr;
                outtags = append(a.tags, rtags);
            } else outc = ca_transparent;
            return;
        }
        case GLOME_MAT_WARP: {  // Shader.hs:157-175
            tl_stats.rays_secondary += 2;
            TraceResult f = trace(S, ls, mat.a, ri.ray, infinity_, recurs - 1);
            Ray hr = {ri.pos, vnorm(ray.d)};  // TestScene.hs:169-173
            Ray wr = xfm_ray(&S.dpool[mat.d], hr);
            TraceResult w = trace(S, mat.c, mat.b, wr, ridepth(f.ri), recurs - 1);
            if (ridepth(f.ri) < ridepth(w.ri)) { outc = f.c; outtags = f.tags; }
            else { outc = w.c; outtags = w.tags; }
            return;
        }
        case GLOME_MAT_ADDITIVE: {  // Shader.hs:177-179, casum Clr.hs:93-103
            Color acc = {0, 0, 0};
            Flt prod = 1;
            for (int i = 0; i < mat.b; i++) {
                ColorA c; List tg;
                Mat m = mat_from_table(S, S.ipool[mat.a + i]);
                mpostshade(S, ls, lights, m, ray, s, ri, recurs, c, tg);
                acc.r = acc.r + c.r * c.a; acc.g = acc.g + c.g * c.a; acc.b = acc.b + c.b * c.a;
                prod = prod * (1 - aclamp(c.a));  // alphas (Clr.hs:82-85): product folds from 1, left to right
                outtags = append(outtags, tg);
            }
            outc.r = acc.r; outc.g = acc.g; outc.b = acc.b; outc.a = 1 - prod;
            return;
        }
        case GLOME_MAT_BLEND: {  // Shader.hs:181-184
            ColorA ca, cb; List ta, tb;
            Mat ma = mat_from_table(S, mat.a), mb = mat_from_table(S, mat.b);
            mpostshade(S, ls, lights, ma, ray, s, ri, recurs, ca, ta);
            mpostshade(S, ls, lights, mb, ray, s, ri, recurs, cb, tb);
            outc = caweight(ca, cb, mat.p[0]);
            outtags = append(ta, tb);
            return;
        }
    }
    outc = ca_transparent;
}

static Mat eval_texture(const Scene& S, int tex, const Ray& ray, const Rayint& ri) {
    (void)ray;
    const GlomeTexture& T = S.textures[tex];
    switch (T.kind) {
        case GLOME_TEX_UNIFORM: return mat_from_table(S, T.a);
        case GLOME_TEX_STRIPE_BLEND: {  // TestScene.hs:225-231, Texture.hs:35-40
            Flt scale = triangle_wave(vdot(ri.pos, vec(T.p[0], T.p[1], T.p[2])));
            Mat m; memset(&m, 0, sizeof(m));
            m.kind = GLOME_MAT_BLEND; m.a = T.a; m.b = T.b; m.p[0] = scale;
            return m;
        }
        case GLOME_TEX_PERLIN_BLEND: {  // TestScene.hs:214-220
            Flt scale = perlin(vscale(ri.pos, T.p[0]));
            Mat m; memset(&m, 0, sizeof(m));
            m.kind = GLOME_MAT_BLEND; m.a = T.a; m.b = T.b; m.p[0] = scale;
            return m;
        }
    }
    return mat_from_table(S, T.a);
}

static inline bool opaque(const ColorA& c) { return c.a + delta >= 1; }  // Trace.hs:50-51

static TraceResult trace(const Scene& S, int lightset, int sld, const Ray& ray, Flt depth, int recurs) {
    // Trace.hs:59-82
    TraceResult out;
    if (recurs == 0) { out.c = ca_transparent; out.ri = raymiss(); return out; }
    List e1, e2;
    Rayint ri = rayint(S, sld, ray, depth, e1, e2, 0);
    out.ri = ri;
    if (!ri.hit) { out.c = ca_transparent; return out; }  // mmissshade (Shader.hs:186-187)
    LightCtx ctxb;
    ctxb.done = false; ctxb.n = 0;
    ColorA colora = ca_transparent;
    List tagsa;
    for (int i = 0; i < ri.tex.n; i++) {
        if (opaque(colora)) continue;
        ColorA colorb; List tagsb;
        Mat m = eval_texture(S, ri.tex.v[i], ray, ri);
        mpostshade(S, lightset, ctxb, m, ray, sld, ri, recurs, colorb, tagsb);
        colora = cafold(colora, colorb);
        tagsa = append(tagsb, tagsa);
    }
    out.c = colora;
    out.tags = append(tagsa, ri.tag);
    return out;
}

// ---- Glome.hs: camera rays, tiles, adaptive AA ------------------------------------------------
struct Camera { Vec pos, fwd, up, right; };
static inline void getCoordsf(int width, int height, Flt xf, Flt yf, Flt& xc, Flt& yc) {  // Glome.hs:119-140
    Flt widthf = (Flt)width, heightf = (Flt)height;
    xc = (((xf / widthf) * 2) - 1) * (widthf / heightf);
    yc = -(((yf / heightf) * 2) - 1);
}
struct TColor { Flt r, g, b, a, d; };
static TColor get_color(const Scene& S, const Camera& cam, Flt x, Flt y, int recurs) {
    // get_rayint (Glome.hs:27-33) + get_color_normal (:53-55)
    Vec dir = vnorm(vadd3(cam.fwd, vscale(cam.right, -x), vscale(cam.up, y)));
    Ray ray = {cam.pos, dir};
    tl_stats.rays_primary++;
    TraceResult tr = trace(S, 0, S.root, ray, infinity_, recurs);
    if (tr.ri.flags) tl_stats.overflow++;
    TColor c = {tr.c.r, tr.c.g, tr.c.b, tr.c.a, ridepth(tr.ri)};
    return c;
}
static inline Flt cCmp(const TColor& p, const TColor& q) {  // Glome.hs:179-189
    Flt md;
    if (p.d == 0 && q.d == 0) md = 0;
    else md = (p.d > q.d) ? (p.d / q.d) - 1 : (q.d / p.d) - 1;
    return fabs_(q.r - p.r) + fabs_(q.g - p.g) + fabs_(q.b - p.b) + fabs_(q.a - p.a) + md;
}
static inline TColor cAvg(const TColor& a, const TColor& b, const TColor& c, const TColor& d) {  // Glome.hs:191-197
    TColor r = {(a.r + b.r + c.r + d.r) * 0.25, (a.g + b.g + c.g + d.g) * 0.25, (a.b + b.b + c.b + d.b) * 0.25,
                (a.a + b.a + c.a + d.a) * 0.25, (a.d + b.d + c.d + d.d) * 0.25};
    return r;
}
static inline TColor cAvg2(const TColor& a, const TColor& b) {  // Glome.hs:199-205
    TColor r = {(a.r + b.r) * 0.5, (a.g + b.g) * 0.5, (a.b + b.b) * 0.5, (a.a + b.a) * 0.5, (a.d + b.d) * 0.5};
    return r;
}
static TColor decide(const Scene& S, const Camera& cam, int recurs, Flt threshold, Flt xf, Flt yf, const TColor& a,
                     const TColor& b, const TColor& c, const TColor& d) {  // Glome.hs:213-219
    Flt variance = fmax_(cCmp(a, c), cCmp(b, d));
    if (variance > threshold) return get_color(S, cam, xf, yf, recurs);
    return cAvg(a, b, c, d);
}
static void chunk(int size, int blocksize, std::vector<std::pair<int, int> >& out) {  // Glome.hs:371-377
    int pos = 0;
    for (;;) {
        if (pos + blocksize >= size) { out.push_back(std::make_pair(pos, size - pos)); return; }
        out.push_back(std::make_pair(pos, blocksize));
        pos += blocksize;
    }
}
static void tile_list(int width, int height, int blocksize, std::vector<int>& rects) {  // Glome.hs:382-384
    std::vector<std::pair<int, int> > xs, ys;
    chunk(width, blocksize, xs);
    chunk(height, blocksize, ys);
    for (size_t i = 0; i < xs.size(); i++)
        for (size_t j = 0; j < ys.size(); j++) {
            rects.push_back(xs[i].first); rects.push_back(ys[j].first);
            rects.push_back(xs[i].second); rects.push_back(ys[j].second);
        }
}
// ---- rayint_debug: the Int half of (Rayint, Int) -----------------------------------------------------
// The count does not depend on what is hit: Bih adds 1 per branch entered on the unculled walk and passes
// `fmin d far` to its leaves (Bih.hs:378-412); [s] sums (Solid.hs:312,329); Instance transforms the ray and scales
// d (Solid.hs:447-461); Bound adds 1 when its gate passes (Bound.hs:37-42); InnerBound defers to sb (Bound.hs:107);
// Tag / Tex / NoShadow pass through, OnlyShadow is (RayMiss,0) (Tex.hs:55,67,79,90); everything else is the class
// default ((rayint ...), 0) (Solid.hs:205).
static int64_t debug_count(const Scene& S, int ni, const Ray& r, Flt d);
static int64_t debug_count_bih(const Scene& S, const Ray& r, Flt d, const Flt* dirr, const Flt* o, int ref, Flt near_, Flt far_) {
    if (ref < 0) {
        int32_t lf, lc;
        glome_bih_leaf(ref, S.ipool.data(), &lf, &lc);
        int64_t c = 0;
        for (int i = 0; i < lc; i++) c += debug_count(S, lf + i, r, fmin_(d, far_));
        return c;
    }
    const GlomeBihNode& n = S.bih[ref];
    Flt dr_ = dirr[n.axis], oo = o[n.axis];
    Flt dl = (n.lsplit - oo) * dr_;
    Flt dr = (n.rsplit - oo) * dr_;
    int64_t c = 1;
    if (near_ > far_) return c;
    if (dr_ > 0) {
        if (near_ < dl) c += debug_count_bih(S, r, d, dirr, o, n.left, near_, fmin_(dl, far_));
        if (dr < far_) c += debug_count_bih(S, r, d, dirr, o, n.right, fmax_(dr, near_), far_);
    } else {
        if (near_ < dr) c += debug_count_bih(S, r, d, dirr, o, n.right, near_, fmin_(dr, far_));
        if (dl < far_) c += debug_count_bih(S, r, d, dirr, o, n.left, fmax_(dl, near_), far_);
    }
    return c;
}
static int64_t debug_count(const Scene& S, int ni, const Ray& r, Flt d) {
    const GlomeNode& nd = S.nodes[ni];
    switch (nd.type) {
        case GLOME_GROUP: {
            int64_t c = 0;
            for (int i = 0; i < nd.b; i++) c += debug_count(S, nd.a + i, r, d);
            return c;
        }
        case GLOME_INSTANCE: {
            const Flt* xfm = &S.dpool[nd.b];
            Vec newdir = invxfm_vec(xfm, r.d);
            Vec neworig = invxfm_point(xfm, r.o);
            Flt lenscale = vlen(newdir);
            Flt invlenscale = 1 / lenscale;
            Ray ir = {neworig, vscale(newdir, invlenscale)};
            return debug_count(S, nd.a, ir, d * lenscale);
        }
        case GLOME_BIH: {
            Bbox bb = ldbb(S, nd.b);
            Flt near_, far_;
            bbclip_ub(r, bb, near_, far_);  // Interval near far = bbclip r bb: no clip by d at the root (Bih.hs:381)
            Flt dirr[3] = {1 / r.d.x, 1 / r.d.y, 1 / r.d.z}, o[3] = {r.o.x, r.o.y, r.o.z};
            return debug_count_bih(S, r, d, dirr, o, nd.a, near_, far_);
        }
        case GLOME_TEX: case GLOME_TAG: case GLOME_NOSHADOW: return debug_count(S, nd.a, r, d);
        case GLOME_BOUND:
            if (inside(S, nd.a, r.o) || shadow(S, nd.a, r, d, 0)) return debug_count(S, nd.b, r, d) + 1;
            return 0;
        case GLOME_INNERBOUND: return debug_count(S, nd.b, r, d);
        default: return 0;
    }
}

static void renderTile(const Scene& S, const Camera& cam, int width, int height, const int* rect, int recurs, int tint,
                       TColor* img, int heatmap = 0) {
    // Glome.hs:162-176
    int xtmin = rect[0], ytmin = rect[1], tw = rect[2], th = rect[3];
    for (int i = 0; i < tw * th; i++) {
        int x = xtmin + (i % tw), y = ytmin + (i / tw);
        Flt xc, yc;
        getCoordsf(width, height, (Flt)x, (Flt)y, xc, yc);
        TColor c = get_color(S, cam, xc, yc, recurs);
        if (heatmap) {  // get_color_debug (Glome.hs:57-60)
            Vec dir = vnorm(vadd3(cam.fwd, vscale(cam.right, -xc), vscale(cam.up, yc)));
            Ray dray = {cam.pos, dir};
            int64_t dbg = debug_count(S, S.root, dray, infinity_);
            c.r = ((Flt)(dbg % 30) / 60) + c.r;
            c.g = c.g + ((Flt)dbg / 1000);
        }
        if (tint) c.r = c.r + (c.d / 400);
        img[(size_t)y * width + x] = c;
    }
}
static void renderTileSubsample(const Scene& S, const Camera& cam, int width, int height, const int* rect, int recurs,
                                const Flt* thr, TColor* img) {
    // Glome.hs:226-323
    int xtmin = rect[0], ytmin = rect[1], tw = rect[2], th = rect[3];
    const TColor init = {0, 0, 0, 0, infinity_};
    std::vector<TColor> v((size_t)tw * th, init), v2((size_t)tw * th, init);
    auto getc = [&](int x, int y) -> TColor {
        if ((x >= xtmin) && (x < xtmin + tw) && (y >= ytmin) && (y < ytmin + th)) return v[(x - xtmin) + ((y - ytmin) * tw)];
        return init;
    };
    auto putc = [&](std::vector<TColor>& buf, int x, int y, const TColor& c) { buf[(x - xtmin) + ((y - ytmin) * tw)] = c; };
    auto decide_int = [&](Flt threshold, int x, int y, const TColor& a, const TColor& b, const TColor& c, const TColor& d) {
        Flt xf, yf;
        getCoordsf(width, height, (Flt)x, (Flt)y, xf, yf);  // getCoords
        return decide(S, cam, recurs, threshold, xf, yf, a, b, c, d);
    };
    for (int x = xtmin; x <= xtmin + tw - 1; x += 2)
        for (int y = ytmin; y <= ytmin + th - 1; y += 2)
            if (((x - xtmin) + (y - ytmin)) % 4 == 0) {
                Flt xf, yf;
                getCoordsf(width, height, (Flt)x, (Flt)y, xf, yf);
                putc(v, x, y, get_color(S, cam, xf, yf, recurs));
            }
    for (int x = xtmin; x <= xtmin + tw - 1; x += 2)
        for (int y = ytmin; y <= ytmin + th - 1; y += 2)
            if (((x - xtmin) + (y - ytmin)) % 4 == 2) {
                TColor a = getc(x - 2, y), b = getc(x, y + 2), c = getc(x + 2, y), d = getc(x, y - 2);
                putc(v, x, y, decide_int(thr[0], x, y, a, b, c, d));
            }
    for (int x = xtmin + 1; x <= xtmin + tw - 1; x += 2)
        for (int y = ytmin + 1; y <= ytmin + th - 1; y += 2) {
            TColor a = getc(x - 1, y - 1), b = getc(x + 1, y - 1), c = getc(x + 1, y + 1), d = getc(x - 1, y + 1);
            putc(v, x, y, decide_int(thr[1], x, y, a, b, c, d));
        }
    for (int x = xtmin; x <= xtmin + tw - 1; x++)
        for (int y = ytmin; y <= ytmin + th - 1; y++)
            if (((x - xtmin) + (y - ytmin)) % 2 == 1) {
                TColor a = getc(x - 1, y), b = getc(x, y + 1), c = getc(x + 1, y), d = getc(x, y - 1);
                putc(v, x, y, decide_int(thr[2], x, y, a, b, c, d));
            }
    for (int x = xtmin; x <= xtmin + tw - 1; x++)
        for (int y = ytmin; y <= ytmin + th - 1; y++) {
            TColor a = getc(x, y), b = getc(x, y + 1), c = getc(x + 1, y + 1), d = getc(x + 1, y);
            Flt xf, yf;
            getCoordsf(width, height, (Flt)x + 0.5, (Flt)y + 0.5, xf, yf);
            TColor color = decide(S, cam, recurs, thr[3], xf, yf, a, b, c, d);
            if (x == xtmin + tw - 1) {
                if (y == ytmin + th - 1) putc(v2, x, y, color);
                else putc(v2, x, y, cAvg2(color, cAvg2(a, b)));
            } else {
                if (y == ytmin + th - 1) putc(v2, x, y, cAvg2(color, cAvg2(a, d)));
                else putc(v2, x, y, cAvg2(color, cAvg(a, b, c, d)));
            }
        }
    for (int y = 0; y < th; y++)
        for (int x = 0; x < tw; x++) img[(size_t)(ytmin + y) * width + (xtmin + x)] = v2[x + y * tw];
}
static inline Flt cap1(Flt x) { return (x >= 1) ? 1 - delta : x; }  // Glome.hs:98-101
static inline uint32_t rgbf(Flt r, Flt g, Flt b) {  // Glome.hs:107-110 (Word32 modular arithmetic)
    uint32_t R = (uint32_t)(int64_t)std::floor(cap1(r) * 256);
    uint32_t G = (uint32_t)(int64_t)std::floor(cap1(g) * 256);
    uint32_t B = (uint32_t)(int64_t)std::floor(cap1(b) * 256);
    return R * (256u * 256u) + G * 256u + B;
}

// ---- tree builders (host side in the reference; restated because no GHC exists here) ----------
struct BObj { Bbox bb; int idx; };
static inline Flt bbsa_p(const Bbox& b) { return hmax(0, bbsa(b)); }  // Bih.hs:208
struct BihOut {
    std::vector<GlomeBihNode> nodes;
    std::vector<int32_t> leaves;  // {first, count} into order
    std::vector<int32_t> order;
};
static inline Bbox vset_p2(const Bbox& b, int axis, Flt f) { Bbox r = b; if (axis == 0) r.p2.x = f; else if (axis == 1) r.p2.y = f; else r.p2.z = f; return r; }
static inline Bbox vset_p1(const Bbox& b, int axis, Flt f) { Bbox r = b; if (axis == 0) r.p1.x = f; else if (axis == 1) r.p1.y = f; else r.p1.z = f; return r; }
static int32_t bih_build_rec(BihOut& out, const std::vector<BObj>& objs, const Bbox& bb, const Vec& mid, int depth) {
    // Bih.hs:211-285.  Returns a child ref (>=0 branch, <0 leaf id encoded as ~leaf_index).
    size_t objcount = objs.size();
    auto build_leaf = [&]() -> int32_t {
        int32_t li = (int32_t)(out.leaves.size() / 2);
        out.leaves.push_back((int32_t)out.order.size());
        out.leaves.push_back((int32_t)objcount);
        for (size_t i = 0; i < objcount; i++) out.order.push_back(objs[i].idx);
        return ~li;
    };
    if (objcount <= 3) return build_leaf();
    Flt sa = bbsa_p(bb);
    std::vector<BObj> l[4], r[4];  // x, y, z, big/small   (partition is stable, Data.List)
    for (size_t i = 0; i < objcount; i++) {
        const BObj& o = objs[i];
        Vec m = bbmid(o.bb);
        (m.x < mid.x ? l[0] : r[0]).push_back(o);
        (m.y < mid.y ? l[1] : r[1]).push_back(o);
        (m.z < mid.z ? l[2] : r[2]).push_back(o);
        (bbsa_p(o.bb) > sa * 0.4 ? l[3] : r[3]).push_back(o);
    }
    Flt lmax[4], rmin[4];
    Bbox lbb[4], rbb[4];
    Flt cost[4];
    for (int k = 0; k < 4; k++) {
        int ax = (k == 3) ? 0 : k;  // big/small always uses x (Bih.hs:231-232)
        Flt lm = -infinity_, rm = infinity_;
        for (size_t i = 0; i < l[k].size(); i++) lm = fmax_(lm, va(l[k][i].bb.p2, ax));
        for (size_t i = 0; i < r[k].size(); i++) rm = fmin_(rm, va(r[k][i].bb.p1, ax));
        lmax[k] = lm; rmin[k] = rm;
        lbb[k] = vset_p2(bb, ax, lm);
        rbb[k] = vset_p1(bb, ax, rm);
        Flt fac = (k == 3) ? 1.2 : 1.1;
        cost[k] = ((bbsa_p(lbb[k]) * (Flt)l[k].size()) + (bbsa_p(rbb[k]) * (Flt)r[k].size())) * fac;
    }
    Flt costx = cost[0], costy = cost[1], costz = cost[2], costb = cost[3];
    Flt costorig = sa * (Flt)objcount;
    if (costorig < costx && costorig < costy && costorig < costz && costorig < costb) return build_leaf();
    int k;
    if (costx < costy && costx < costz && costx < costb) k = 0;
    else if (costy < costz && costy < costb) k = 1;
    else if (costy < costb) k = 2;  // sic: costy, not costz (Bih.hs:283)
    else k = 3;
    int axis = (k == 3) ? 0 : k;
    int32_t me = (int32_t)out.nodes.size();
    GlomeBihNode nd;
    memset(&nd, 0, sizeof(nd));
    nd.lsplit = lmax[k] + delta;
    nd.rsplit = rmin[k] - delta;
    nd.axis = axis;
    out.nodes.push_back(nd);
    // free the candidates we do not follow before recursing
    std::vector<BObj> lo, ro;
    lo.swap(l[k]); ro.swap(r[k]);
    for (int j = 0; j < 4; j++) { std::vector<BObj>().swap(l[j]); std::vector<BObj>().swap(r[j]); }
    Bbox lb = lbb[k], rb = rbb[k];
    int32_t lref = bih_build_rec(out, lo, lb, bbmid(lb), depth + 1);
    std::vector<BObj>().swap(lo);
    int32_t rref = bih_build_rec(out, ro, rb, bbmid(rb), depth + 1);
    out.nodes[me].left = lref;
    out.nodes[me].right = rref;
    return me;
}

struct BvhOut {
    std::vector<GlomeBvhNode> nodes;
    std::vector<int32_t> leafpool;  // {count, tri...} records
    std::vector<int32_t> leafoff;   // leaf index -> offset into leafpool
};
static Bbox bbpts3(const Vec& a, const Vec& b, const Vec& c) {  // bbpts [a,b,c] (Vec.hs:676-690)
    Bbox r = {vec(c.x - delta, c.y - delta, c.z - delta), vec(c.x + delta, c.y + delta, c.z + delta)};
    const Vec* ps[2] = {&b, &a};
    for (int i = 0; i < 2; i++) {
        const Vec& p = *ps[i];
        r.p1 = vec(fmin_(p.x - delta, r.p1.x), fmin_(p.y - delta, r.p1.y), fmin_(p.z - delta, r.p1.z));
        r.p2 = vec(fmax_(p.x + delta, r.p2.x), fmax_(p.y + delta, r.p2.y), fmax_(p.z + delta, r.p2.z));
    }
    return r;
}
static int32_t mesh_build_tree(BvhOut& out, const std::vector<Bbox>& alltribbs, const std::vector<int32_t>& tris,
                               const Bbox& bb) {
    // Mesh.hs:69-113
    size_t n = tris.size();
    auto leaf = [&]() -> int32_t {
        int32_t li = (int32_t)out.leafoff.size();
        out.leafoff.push_back((int32_t)out.leafpool.size());
        out.leafpool.push_back((int32_t)n);
        for (size_t i = 0; i < n; i++) out.leafpool.push_back(tris[i]);
        return ~li;
    };
    if (n < 3) return leaf();
    Vec mid = bbmid(bb);
    Flt sa = bbsa(bb);
    std::vector<int32_t> l[4], r[4];
    for (size_t i = 0; i < n; i++) {
        int32_t t = tris[i];
        const Bbox& tb = alltribbs[t];
        Vec m = bbmid(tb);
        (m.x < mid.x ? l[0] : r[0]).push_back(t);
        (m.y < mid.y ? l[1] : r[1]).push_back(t);
        (m.z < mid.z ? l[2] : r[2]).push_back(t);
        (bbsa(tb) > sa * 0.4 ? l[3] : r[3]).push_back(t);
    }
    Bbox lbb[4], rbb[4];
    Flt cost[4];
    for (int k = 0; k < 4; k++) {
        Bbox a = empty_bbox(), b = empty_bbox();  // trisbb: V.foldl' bbjoin empty_bbox (Mesh.hs:124-125)
        for (size_t i = 0; i < l[k].size(); i++) a = bbjoin(a, alltribbs[l[k][i]]);
        for (size_t i = 0; i < r[k].size(); i++) b = bbjoin(b, alltribbs[r[k][i]]);
        lbb[k] = a; rbb[k] = b;
        cost[k] = ((bbsa(a) * (Flt)l[k].size()) + (bbsa(b) * (Flt)r[k].size())) * 1.1;
    }
    Flt xcost = cost[0], ycost = cost[1], zcost = cost[2], bcost = cost[3];
    Flt lcost = bbsa(bb) * (Flt)n;
    if (lcost < xcost && lcost < ycost && lcost < zcost && lcost < bcost) return leaf();
    int k;
    if (xcost < ycost && xcost < zcost && xcost < bcost) k = 0;
    else if (ycost < zcost && ycost < bcost) k = 1;
    else if (zcost < bcost) k = 2;
    else k = 3;
    int32_t me = (int32_t)out.nodes.size();
    GlomeBvhNode nd;
    memset(&nd, 0, sizeof(nd));
    const Bbox& L = lbb[k]; const Bbox& R = rbb[k];
    nd.lbb[0] = L.p1.x; nd.lbb[1] = L.p1.y; nd.lbb[2] = L.p1.z; nd.lbb[3] = L.p2.x; nd.lbb[4] = L.p2.y; nd.lbb[5] = L.p2.z;
    nd.rbb[0] = R.p1.x; nd.rbb[1] = R.p1.y; nd.rbb[2] = R.p1.z; nd.rbb[3] = R.p2.x; nd.rbb[4] = R.p2.y; nd.rbb[5] = R.p2.z;
    out.nodes.push_back(nd);
    std::vector<int32_t> lo, ro;
    lo.swap(l[k]); ro.swap(r[k]);
    for (int j = 0; j < 4; j++) { std::vector<int32_t>().swap(l[j]); std::vector<int32_t>().swap(r[j]); }
    Bbox lb = lbb[k], rb = rbb[k];
    int32_t lref = mesh_build_tree(out, alltribbs, lo, lb);
    std::vector<int32_t>().swap(lo);
    int32_t rref = mesh_build_tree(out, alltribbs, ro, rb);
    out.nodes[me].left = lref;
    out.nodes[me].right = rref;
    return me;
}

}  // namespace orc

// ===============================================================================================
// C API (ctypes).  Everything here is test / baseline plumbing.
// ===============================================================================================
using namespace orc;

extern "C" {

struct OrcScene { Scene S; };

void* orc_scene_create(const GlomeFlatScene* d) {
    OrcScene* o = new OrcScene();
    Scene& S = o->S;
    S.root = d->root;
    S.nodes.assign(d->nodes, d->nodes + d->n_nodes);
    if (d->n_bihnodes) S.bih.assign(d->bihnodes, d->bihnodes + d->n_bihnodes);
    if (d->n_bvhnodes) S.bvh.assign(d->bvhnodes, d->bvhnodes + d->n_bvhnodes);
    if (d->n_ipool) S.ipool.assign(d->ipool, d->ipool + d->n_ipool);
    if (d->n_dpool) S.dpool.assign(d->dpool, d->dpool + d->n_dpool);
    if (d->n_textures) S.textures.assign(d->textures, d->textures + d->n_textures);
    if (d->n_materials) S.materials.assign(d->materials, d->materials + d->n_materials);
    if (d->n_lights) S.lights.assign(d->lights, d->lights + d->n_lights);
    if (d->n_lightsets) S.lightsets.assign(d->lightsets, d->lightsets + 2 * d->n_lightsets);
    else { S.lightsets.push_back(0); S.lightsets.push_back(d->n_lights); }
    S.total.clear();
    return o;
}
void orc_scene_destroy(void* h) { delete (OrcScene*)h; }

static void fill_hit(const Rayint& ri, GlomeHit* h) {
    memset(h, 0, sizeof(*h));
    h->t = ridepth(ri);
    h->hit = ri.hit ? 1 : 0;
    h->prim = ri.hit ? ri.prim : -1;
    h->sub = ri.hit ? ri.sub : -1;
    h->flags = ri.flags;
    if (ri.hit) {
        h->pos[0] = ri.pos.x; h->pos[1] = ri.pos.y; h->pos[2] = ri.pos.z;
        h->norm[0] = ri.norm.x; h->norm[1] = ri.norm.y; h->norm[2] = ri.norm.z;
        h->ntex = ri.tex.n; h->ntag = ri.tag.n;
        for (int i = 0; i < ri.tex.n && i < GLOME_MAX_STACK; i++) h->tex[i] = ri.tex.v[i];
        for (int i = 0; i < ri.tag.n && i < GLOME_MAX_STACK; i++) h->tag[i] = ri.tag.v[i];
    }
}

}  // extern "C"
template <typename F>
static void parallel_for(int64_t n, int threads, Scene& S, F f) {
    if (threads < 1) threads = 1;
    std::atomic<int64_t> next(0);
    std::vector<Stats> st(threads);
    auto worker = [&](int tid) {
        tl_stats.clear();
        const int64_t grain = 256;
        for (;;) {
            int64_t b = next.fetch_add(grain);
            if (b >= n) break;
            int64_t e = std::min(n, b + grain);
            for (int64_t i = b; i < e; i++) f(i);
        }
        st[tid] = tl_stats;
    };
    if (threads == 1) worker(0);
    else {
        std::vector<std::thread> th;
        for (int t = 0; t < threads; t++) th.emplace_back(worker, t);
        for (auto& t : th) t.join();
    }
    for (int t = 0; t < threads; t++) S.total.add(st[t]);
}

static inline Ray ldray(const double* rays, int64_t i) {
    Ray r = {vec(rays[6 * i], rays[6 * i + 1], rays[6 * i + 2]), vec(rays[6 * i + 3], rays[6 * i + 4], rays[6 * i + 5])};
    return r;
}

extern "C" {
void orc_rayint_batch(void* h, int64_t n, const double* rays, const double* tmax, int tmax_stride, GlomeHit* out,
                      int threads) {
    Scene& S = ((OrcScene*)h)->S;
    parallel_for(n, threads, S, [&](int64_t i) {
        List e1, e2;
        Rayint ri = rayint(S, S.root, ldray(rays, i), tmax[tmax_stride ? i : 0], e1, e2, 0);
        fill_hit(ri, &out[i]);
    });
}
void orc_debug_count_batch(void* h, int64_t n, const double* rays, const double* tmax, int tmax_stride, int32_t* out, int threads) {
    Scene& S = ((OrcScene*)h)->S;
    parallel_for(n, threads, S, [&](int64_t i) { out[i] = (int32_t)debug_count(S, S.root, ldray(rays, i), tmax[tmax_stride ? i : 0]); });
}
void orc_shadow_batch(void* h, int64_t n, const double* rays, const double* tmax, int tmax_stride, uint8_t* occ,
                      int threads) {
    Scene& S = ((OrcScene*)h)->S;
    parallel_for(n, threads, S, [&](int64_t i) { occ[i] = shadow(S, S.root, ldray(rays, i), tmax[tmax_stride ? i : 0], 0) ? 1 : 0; });
}
void orc_inside_batch(void* h, int64_t n, const double* pts, uint8_t* out, int threads) {
    Scene& S = ((OrcScene*)h)->S;
    parallel_for(n, threads, S, [&](int64_t i) { out[i] = inside(S, S.root, vec(pts[3 * i], pts[3 * i + 1], pts[3 * i + 2])) ? 1 : 0; });
}
void orc_trace_batch(void* h, int64_t n, const double* rays, const double* tmax, int tmax_stride, int recurs,
                     double* rgba, double* depth, GlomeHit* hits, int32_t* tags /* n*17: count + 16 */, int threads) {
    Scene& S = ((OrcScene*)h)->S;
    parallel_for(n, threads, S, [&](int64_t i) {
        TraceResult tr = trace(S, 0, S.root, ldray(rays, i), tmax[tmax_stride ? i : 0], recurs);
        rgba[4 * i] = tr.c.r; rgba[4 * i + 1] = tr.c.g; rgba[4 * i + 2] = tr.c.b; rgba[4 * i + 3] = tr.c.a;
        depth[i] = ridepth(tr.ri);
        if (hits) fill_hit(tr.ri, &hits[i]);
        if (tags) {
            tags[17 * i] = tr.tags.n;
            for (int k = 0; k < 16; k++) tags[17 * i + 1 + k] = k < tr.tags.n ? tr.tags.v[k] : -1;
        }
    });
}

int orc_tile_count(int width, int height, int blocksize) {
    std::vector<int> r;
    tile_list(width, height, blocksize, r);
    return (int)(r.size() / 4);
}
void orc_tile_rects(int width, int height, int blocksize, int32_t* out) {
    std::vector<int> r;
    tile_list(width, height, blocksize, r);
    for (size_t i = 0; i < r.size(); i++) out[i] = r[i];
}

// renderTiles (Glome.hs:379-386): one worker per thread pulling tiles from an atomic counter
// (mirrors parMap); max_tiles > 0 renders only the first max_tiles selected tiles (bounded sample).
void orc_render(void* h, const GlomeCamera* gc, int width, int height, const GlomeRenderOpts* o, double* tcolor,
                uint32_t* rgb8, int threads, int max_tiles) {
    Scene& S = ((OrcScene*)h)->S;
    Camera cam = {vec(gc->pos[0], gc->pos[1], gc->pos[2]), vec(gc->fwd[0], gc->fwd[1], gc->fwd[2]),
                  vec(gc->up[0], gc->up[1], gc->up[2]), vec(gc->right[0], gc->right[1], gc->right[2])};
    std::vector<int> rects;
    tile_list(width, height, o->blocksize, rects);
    std::vector<int> sel;
    int stride = o->tile_stride > 0 ? o->tile_stride : 1;
    for (int i = 0; i < (int)(rects.size() / 4); i++)
        if (i % stride == o->tile_first) sel.push_back(i);
    if (max_tiles > 0 && (int)sel.size() > max_tiles) sel.resize(max_tiles);
    TColor* img = (TColor*)tcolor;
    if (threads < 1) threads = 1;
    std::atomic<int> next(0);
    std::vector<Stats> st(threads);
    auto worker = [&](int tid) {
        tl_stats.clear();
        for (;;) {
            int k = next.fetch_add(1);
            if (k >= (int)sel.size()) break;
            const int* rect = &rects[4 * sel[k]];
            if (o->mode == GLOME_MODE_ADAPTIVE_AA) renderTileSubsample(S, cam, width, height, rect, o->recurs, o->thresholds, img);
            else renderTile(S, cam, width, height, rect, o->recurs, o->tint_depth, img, o->debug_heatmap);
            if (rgb8)
                for (int y = rect[1]; y < rect[1] + rect[3]; y++)
                    for (int x = rect[0]; x < rect[0] + rect[2]; x++) {
                        const TColor& c = img[(size_t)y * width + x];
                        rgb8[(size_t)y * width + x] = rgbf(c.r * c.a, c.g * c.a, c.b * c.a);  // blitTile (Glome.hs:353-358)
                    }
        }
        st[tid] = tl_stats;
    };
    if (threads == 1) worker(0);
    else {
        std::vector<std::thread> th;
        for (int t = 0; t < threads; t++) th.emplace_back(worker, t);
        for (auto& t : th) t.join();
    }
    for (int t = 0; t < threads; t++) S.total.add(st[t]);
}

// stats: out[] = {bih_branch, bvh_branch, bih_leaf_items, tri, trinorm, instance, rays_primary, rays_shadow,
//                 rays_secondary, overflow, perlin_range, prim[0..20]}
int orc_stats(void* h, int64_t* out, int reset) {
    Scene& S = ((OrcScene*)h)->S;
    const Stats& s = S.total;
    int k = 0;
    out[k++] = s.bih_branch; out[k++] = s.bvh_branch; out[k++] = s.bih_leaf_items; out[k++] = s.tri; out[k++] = s.trinorm;
    out[k++] = s.instance; out[k++] = s.rays_primary; out[k++] = s.rays_shadow; out[k++] = s.rays_secondary;
    out[k++] = s.overflow; out[k++] = s.perlin_range;
    for (int i = 0; i < GLOME_NODE_TYPE_COUNT; i++) out[k++] = s.prim[i];
    if (reset) S.total.clear();
    return k;
}

// bih (Bih.hs:309-324) over n bounding boxes.  Outputs malloc'ed; caller frees with orc_free.
// Node numbering: pre-order; leaves numbered left to right; order = concatenated leaf items.
int orc_bih_build(int64_t n, const double* bboxes, GlomeBihNode** nodes_out, int32_t* n_nodes_out, int32_t** leaves_out,
                  int32_t* n_leaves_out, int32_t** order_out, int32_t* root_ref_out, double* bb_out) {
    std::vector<BObj> objs((size_t)n);
    Bbox bb = empty_bbox();
    for (int64_t i = 0; i < n; i++) {
        const double* b = bboxes + 6 * i;
        objs[i].bb.p1 = vec(b[0], b[1], b[2]);
        objs[i].bb.p2 = vec(b[3], b[4], b[5]);
        objs[i].idx = (int)i;
        bb = bbjoin(bb, objs[i].bb);  // foldl' bbjoin empty_bbox
    }
    if (n > 0 && (bb.p1.x == -infinity_ || bb.p1.y == -infinity_ || bb.p1.z == -infinity_ || bb.p2.x == infinity_ ||
                  bb.p2.y == infinity_ || bb.p2.z == infinity_))
        return -1;  // error "bih: infinite bounding box" (Bih.hs:319-322)
    BihOut out;
    int32_t root = bih_build_rec(out, objs, bb, bbmid(bb), 0);
    *n_nodes_out = (int32_t)out.nodes.size();
    *n_leaves_out = (int32_t)(out.leaves.size() / 2);
    *nodes_out = (GlomeBihNode*)malloc(sizeof(GlomeBihNode) * (out.nodes.size() + 1));
    memcpy(*nodes_out, out.nodes.data(), sizeof(GlomeBihNode) * out.nodes.size());
    *leaves_out = (int32_t*)malloc(sizeof(int32_t) * (out.leaves.size() + 2));
    memcpy(*leaves_out, out.leaves.data(), sizeof(int32_t) * out.leaves.size());
    *order_out = (int32_t*)malloc(sizeof(int32_t) * (out.order.size() + 1));
    memcpy(*order_out, out.order.data(), sizeof(int32_t) * out.order.size());
    *root_ref_out = root;
    bb_out[0] = bb.p1.x; bb_out[1] = bb.p1.y; bb_out[2] = bb.p1.z; bb_out[3] = bb.p2.x; bb_out[4] = bb.p2.y; bb_out[5] = bb.p2.z;
    return 0;
}

// mesh (Mesh.hs:50-134): builds the BVH.  leafpool = {count, tri...} records, leafoff[leaf] = offset.
int orc_mesh_build(int64_t nverts, const double* verts, int64_t ntris, const int32_t* tris, GlomeBvhNode** nodes_out,
                   int32_t* n_nodes_out, int32_t** leafpool_out, int32_t* n_leafpool_out, int32_t** leafoff_out,
                   int32_t* n_leaves_out, int32_t* root_ref_out, double* bb_out) {
    // bbox = bbpts (V.toList verts)  (Mesh.hs:55): right fold, each point inflated by delta
    Bbox bb = empty_bbox();
    for (int64_t i = nverts - 1; i >= 0; i--) {
        Vec p = vec(verts[3 * i], verts[3 * i + 1], verts[3 * i + 2]);
        if (i == nverts - 1) {
            bb.p1 = vec(p.x - delta, p.y - delta, p.z - delta);
            bb.p2 = vec(p.x + delta, p.y + delta, p.z + delta);
        } else {
            bb.p1 = vec(fmin_(p.x - delta, bb.p1.x), fmin_(p.y - delta, bb.p1.y), fmin_(p.z - delta, bb.p1.z));
            bb.p2 = vec(fmax_(p.x + delta, bb.p2.x), fmax_(p.y + delta, bb.p2.y), fmax_(p.z + delta, bb.p2.z));
        }
    }
    std::vector<Bbox> alltribbs((size_t)ntris);
    std::vector<int32_t> all((size_t)ntris);
    for (int64_t i = 0; i < ntris; i++) {
        const int32_t* T = tris + 8 * i;
        Vec a = vec(verts[3 * T[0]], verts[3 * T[0] + 1], verts[3 * T[0] + 2]);
        Vec b = vec(verts[3 * T[1]], verts[3 * T[1] + 1], verts[3 * T[1] + 2]);
        Vec c = vec(verts[3 * T[2]], verts[3 * T[2] + 1], verts[3 * T[2] + 2]);
        alltribbs[i] = bbpts3(a, b, c);
        all[i] = (int32_t)i;
    }
    BvhOut out;
    int32_t root = mesh_build_tree(out, alltribbs, all, bb);
    *n_nodes_out = (int32_t)out.nodes.size();
    *nodes_out = (GlomeBvhNode*)malloc(sizeof(GlomeBvhNode) * (out.nodes.size() + 1));
    memcpy(*nodes_out, out.nodes.data(), sizeof(GlomeBvhNode) * out.nodes.size());
    *n_leafpool_out = (int32_t)out.leafpool.size();
    *leafpool_out = (int32_t*)malloc(sizeof(int32_t) * (out.leafpool.size() + 1));
    memcpy(*leafpool_out, out.leafpool.data(), sizeof(int32_t) * out.leafpool.size());
    *n_leaves_out = (int32_t)out.leafoff.size();
    *leafoff_out = (int32_t*)malloc(sizeof(int32_t) * (out.leafoff.size() + 1));
    memcpy(*leafoff_out, out.leafoff.data(), sizeof(int32_t) * out.leafoff.size());
    *root_ref_out = root;
    bb_out[0] = bb.p1.x; bb_out[1] = bb.p1.y; bb_out[2] = bb.p1.z; bb_out[3] = bb.p2.x; bb_out[4] = bb.p2.y; bb_out[5] = bb.p2.z;
    return 0;
}
void orc_free(void* p) { free(p); }

// small scalar probes for the known-answer tests
void orc_bbclip_ub(const double* ray, const double* bb, double* out) {
    Ray r = ldray(ray, 0);
    Bbox b = {vec(bb[0], bb[1], bb[2]), vec(bb[3], bb[4], bb[5])};
    bbclip_ub(r, b, out[0], out[1]);
}
void orc_getcoords(int w, int h, double x, double y, double* out) { getCoordsf(w, h, x, y, out[0], out[1]); }
double orc_perlin(const double* p) { return perlin(vec(p[0], p[1], p[2])); }
double orc_triangle_wave(double x) { return triangle_wave(x); }
uint32_t orc_rgbf(double r, double g, double b) { return rgbf(r, g, b); }
double orc_ccmp(const double* p, const double* q) {
    TColor a = {p[0], p[1], p[2], p[3], p[4]}, b = {q[0], q[1], q[2], q[3], q[4]};
    return cCmp(a, b);
}

}  // extern "C"
