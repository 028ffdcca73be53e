// glome_device.cuh -- device-side GlomeTrace: rayint / shadow / inside / get_metainfo over the
// FlatScene, trace + materialShader, for sm_100a.  FP64, compiled with -fmad=false.
//
// One source, two instantiations (template parameter L):
//   L >= 0  "flat" levels: no device recursion, everything inlines.  Level 0 = scene root
//           ({Tex,Tag}* then prim | Group | Bih | Mesh), level 1 = group child, level 2 = BIH leaf
//           item.  Eligible scenes are GLOME_CLASS_FLAT (host_builder.cpp: flat_class).
//   L == -1 general: the full scene graph (Instance, CSG, Bound, nested Bih ...) by true device
//           recursion, like the reference's recursive `rayint`.
//
// Traversal order is the reference's (near child first, `nearest` folded left to right, ties to
// the later operand).  The one deliberate difference: subtrees whose entry distance is already
// beyond the best hit found so far are skipped.  DESIGN.md ("best-hit culling") argues this never
// changes a result because every bound is padded by delta = 1e-4; tests/ check it ray for ray.
#pragma once
#include <stdint.h>

#include "../../include/glome_cuda.h"
#include "glome_math.h"

// Everything that is not a kernel is host+device so that the scene-graph machine (glome_gen.cuh) can also be compiled
// by g++ into the CPU debugging harness of tests/ (tests/tools/gen_host.cpp: test infrastructure, never the product).
#if defined(__CUDACC__)
#define GD_FN __host__ __device__ __forceinline__
#define GD_NOINLINE __host__ __device__ __noinline__
#else
#define GD_FN inline
#define GD_NOINLINE __attribute__((noinline))
struct double2 { double x, y; };
struct int2 { int x, y; };
struct int4 { int x, y, z, w; };
#endif

namespace gdev {

using namespace glm;

template <typename T>
GD_FN T gd_ldg(const T* p) {
#if defined(__CUDA_ARCH__)
    return __ldg(p);
#else
    return *p;
#endif
}

// Device records of the working precision.  FP64: the FlatScene's own records.  FP32 (GLOME_F32): payloads converted at
// upload -- a 16-byte BIH node (one load; axis rides in the low two bits of the right ref) and a 64-byte BVH node.
#ifdef GLOME_F32
typedef float2 Flt2;
struct DBihNode { float lsplit, rsplit; int32_t left, right_axis; };
struct DBvhNode { float lbb[6], rbb[6]; int32_t left, right; int32_t pad[2]; };
#define GLM_BIG FL(3.0e38)
#else
typedef double2 Flt2;
typedef GlomeBihNode DBihNode;
typedef GlomeBvhNode DBvhNode;
#define GLM_BIG FL(1.0e300)
#endif
static_assert(sizeof(DBihNode) % 16 == 0 && sizeof(DBvhNode) % 16 == 0, "node records are read with 16-byte loads");

struct DScene {
    const GlomeNode* __restrict__ nodes;
    const DBihNode* __restrict__ bih;
    const DBvhNode* __restrict__ bvh;
    const int32_t* __restrict__ ipool;
    const Flt* __restrict__ dpool;
    const GlomeTexture* __restrict__ textures;
    const GlomeMaterial* __restrict__ materials;
    const GlomeLight* __restrict__ lights;
    const int32_t* __restrict__ lightsets;
    const int32_t* __restrict__ tagvals;  // general-class scenes: dense tag id -> the caller's tag value (glome_gen.cuh)
    const int4* __restrict__ items;       // general-class scenes: one item descriptor per node (glome_tagmap.h)
    // Mesh leaves with their vertices inline (built at upload): 9 Flt {a, b, c} per leaf-pool slot, indexed by
    // (ipool position of the slot's triangle index) - leafv_base.  A leaf visit then reads its triangles' vertices from one
    // contiguous run instead of following triangle index -> Tri record -> three vertices.  nullptr: not built.
    const Flt* __restrict__ leafv;
    int leafv_base;
    int root;
    int n_lights;
};

#define GDEV_CSG_CAP 48     /* rayint_advance chain cap on the device (flagged, never silent) */
#define GDEV_BIH_STACK 64   /* traversal stack entries (thread-local memory) */

// traversal visit counters (flat kernels only): they feed the roofline's algorithmic-bytes figure
struct Cnt { unsigned int bih, prim, bvh, tri; };

struct Stk {
    int n;
    int v[GLOME_MAX_STACK];
};
GD_FN void stk_clear(Stk& s) { s.n = 0; }
// x : l  (head first); returns true on overflow
GD_FN bool stk_cons(Stk& out, int x, const Stk& l) {
    bool ovf = l.n >= GLOME_MAX_STACK;
    int n = ovf ? GLOME_MAX_STACK : l.n + 1;
#pragma unroll
    for (int i = GLOME_MAX_STACK - 1; i >= 1; i--) out.v[i] = l.v[i - 1];
    out.v[0] = x;
    out.n = n;
    return ovf;
}
// a ++ b
GD_FN bool stk_append(Stk& out, const Stk& a, const Stk& b) {
    Stk r = a;
    bool ovf = false;
    for (int i = 0; i < b.n; i++) {
        if (r.n < GLOME_MAX_STACK) r.v[r.n++] = b.v[i];
        else ovf = true;
    }
    out = r;
    return ovf;
}

// Rayint (Solid.hs:20-28)
struct Hit {
    Flt t;
    Vec pos, norm;
    Ray ray;
    int hit, prim, sub, flags;
    Stk tex, tag;
};
GD_FN void hit_clear(Hit& h) {
    h.t = GLM_INFINITY; h.hit = 0; h.prim = -1; h.sub = -1; h.flags = 0;
    h.tex.n = 0; h.tag.n = 0;
}
GD_FN Flt ridepth(const Hit& h) { return h.hit ? h.t : GLM_INFINITY; }  // Solid.hs:33
// nearest acc cand (Solid.hs:37-44): cand replaces acc unless acc is strictly nearer
GD_FN bool cand_wins(const Hit& acc, Flt t) { return !acc.hit || !(acc.t < t); }
GD_FN void fold_nearest(Hit& acc, const Hit& c) {
    int fl = acc.flags | c.flags;
    if (c.hit && cand_wins(acc, c.t)) acc = c;
    acc.flags = fl;
}

GD_FN Vec ldv(const Flt* __restrict__ p) { return vec(p[0], p[1], p[2]); }
GD_FN Bbox ldbb(const Flt* __restrict__ p) {
    // bbox records are 16-byte aligned (FP64): three 16-byte loads
    const Flt2* q = reinterpret_cast<const Flt2*>(p);
    Flt2 a = gd_ldg(q), b = gd_ldg(q + 1), c = gd_ldg(q + 2);
    return mkbb(vec(a.x, a.y, b.x), vec(b.y, c.x, c.y));
}
// order-preserving integer key of a non-negative depth (atomicMin on it) and back
GD_FN unsigned long long flt_key(Flt x) {
#if defined(__CUDA_ARCH__) && defined(GLOME_F32)
    return (unsigned long long)__float_as_uint(x);
#elif defined(__CUDA_ARCH__)
    return (unsigned long long)__double_as_longlong(x);
#else
    unsigned long long w = 0; __builtin_memcpy(&w, &x, sizeof(x)); return w;
#endif
}
GD_FN Flt key_flt(unsigned long long w) {
#if defined(__CUDA_ARCH__) && defined(GLOME_F32)
    return __uint_as_float((unsigned int)w);
#elif defined(__CUDA_ARCH__)
    return __longlong_as_double((long long)w);
#else
    Flt x; __builtin_memcpy(&x, &w, sizeof(x)); return x;
#endif
}
// one BIH branch (BihBranch lsplit rsplit axis l r, Bih.hs:55-57)
struct BihStep { Flt ls, rs; int axis, left, right; };
GD_FN BihStep ld_bih(const DBihNode* __restrict__ base, int ref) {
    BihStep b;
#ifdef GLOME_F32
    const int4 w = gd_ldg(reinterpret_cast<const int4*>(base + ref));
    b.ls = key_flt((unsigned int)w.x); b.rs = key_flt((unsigned int)w.y); b.left = w.z; b.axis = w.w & 3; b.right = w.w >> 2;
#else
    const double2* np = reinterpret_cast<const double2*>(base + ref);  // 32-byte node: two 16-byte loads
    const double2 sp2 = gd_ldg(np);
    const int4 ii = gd_ldg(reinterpret_cast<const int4*>(np + 1));
    b.ls = sp2.x; b.rs = sp2.y; b.axis = ii.x; b.left = ii.y; b.right = ii.z;
#endif
    return b;
}
// one Mesh BVH branch (Branch lbb rbb l r, Mesh.hs:36)
GD_FN void ld_bvh(const DBvhNode* __restrict__ base, int ref, Bbox& lbb, Bbox& rbb, int& left, int& right) {
    const Flt2* np = reinterpret_cast<const Flt2*>(base + ref);
    const Flt2 l0 = gd_ldg(np), l1 = gd_ldg(np + 1), l2 = gd_ldg(np + 2);
    const Flt2 r0 = gd_ldg(np + 3), r1 = gd_ldg(np + 4), r2 = gd_ldg(np + 5);
    const int2 kids = gd_ldg(reinterpret_cast<const int2*>(np + 6));
    lbb = mkbb(vec(l0.x, l0.y, l1.x), vec(l1.y, l2.x, l2.y));
    rbb = mkbb(vec(r0.x, r0.y, r1.x), vec(r1.y, r2.x, r2.y));
    left = kids.x; right = kids.y;
}

// ---------------------------------------------------------------------------------------------
// primitives.  FULL = also produce position and normal.
// ---------------------------------------------------------------------------------------------
template <bool FULL>
GD_FN bool prim_sphere(const Flt* __restrict__ p, const Ray& ray, Flt dist, Flt& t, Vec& pos,
                                            Vec& n) {
    // Sphere.hs:20-41.  {cx,cy,cz,r} is 32-byte aligned: two 16-byte loads
    const Flt2* q = reinterpret_cast<const Flt2*>(p);
    Flt2 c01 = gd_ldg(q), c23 = gd_ldg(q + 1);
    Vec center = vec(c01.x, c01.y, c23.x);
    Flt r = c23.y;
    Vec eo = vsub(center, ray.o);
    Flt v = vdot(eo, ray.d);
    Flt vsqr = v * v;
    Flt csqr = vdot(eo, eo);
    Flt rsqr = r * r;
    Flt disc = rsqr - (csqr - vsqr);
    if (disc < 0) return false;
    Flt d = sqrt(disc);
    Flt hitdist = ((v - d) > 0) ? (v - d) : (v + d);
    if ((hitdist < 0) || (hitdist > dist)) return false;
    t = hitdist;
    if (FULL) {
        pos = vscaleadd(ray.o, ray.d, hitdist);
        n = vnorm(vsub(pos, center));
    }
    return true;
}
GD_FN bool shadow_sphere(const Flt* __restrict__ p, const Ray& ray, Flt dist) {
    // Sphere.hs:51-71
    const Flt2* q = reinterpret_cast<const Flt2*>(p);
    Flt2 c01 = gd_ldg(q), c23 = gd_ldg(q + 1);
    Vec center = vec(c01.x, c01.y, c23.x);
    Flt r = c23.y;
    Vec eo = vsub(center, ray.o);
    Flt v = vdot(eo, ray.d);
    if ((dist >= (v - r)) && (v > 0)) {
        Flt vsqr = v * v;
        Flt csqr = vdot(eo, eo);
        Flt rsqr = r * r;
        Flt disc = rsqr - (csqr - vsqr);
        if (disc < 0) return false;
        Flt d = sqrt(disc);
        Flt hitdist = ((v - d) > 0) ? (v - d) : (v + d);
        if ((hitdist < 0) || (hitdist > dist)) return false;
        return true;
    }
    return false;
}

// Triangle.hs:45-73 / 109-141.  mode: 0 = flat normal, 1 = vertex normals (n1..n3 valid)
template <bool FULL>
GD_FN bool prim_triangle(const Vec& p1, const Vec& p2, const Vec& p3, bool smooth, const Vec& n1,
                                              const Vec& n2, const Vec& n3, const Ray& ray, Flt dist, Flt& t, Vec& pos,
                                              Vec& n) {
    Vec e1 = vsub(p2, p1);
    Vec e2 = vsub(p3, p1);
    Vec s1 = vcross(ray.d, e2);
    Flt divisor = vdot(s1, e1);
    if (divisor == 0) return false;
    Flt invdivisor = FL(1.0) / divisor;
    Vec d = vsub(ray.o, p1);
    Flt b1 = vdot(d, s1) * invdivisor;
    if (b1 < 0 || b1 > 1) return false;
    Vec s2 = vcross(d, e1);
    Flt b2 = vdot(ray.d, s2) * invdivisor;
    if (b2 < 0 || b1 + b2 > 1) return false;
    Flt tt = vdot(e2, s2) * invdivisor;
    if (tt < 0 || tt > dist) return false;
    t = tt;
    if (FULL) {
        pos = vscaleadd(ray.o, ray.d, tt);
        if (smooth) {
            Vec n1s = vscale(n1, 1 - (b1 + b2));
            Vec n2s = vscale(n2, b1);
            Vec n3s = vscale(n3, b2);
            n = vnorm(vadd3(n1s, n2s, n3s));
        } else {
            n = vnorm(vcross(e1, e2));
        }
    }
    return true;
}
GD_FN bool shadow_triangle(const Vec& p1, const Vec& p2, const Vec& p3, const Ray& ray, Flt dist) {
    // Triangle.hs:82-107
    Vec e1 = vsub(p2, p1);
    Vec e2 = vsub(p3, p1);
    Vec s1 = vcross(ray.d, e2);
    Flt divisor = vdot(s1, e1);
    if (divisor == 0) return false;
    Flt invdivisor = FL(1.0) / divisor;
    Vec d = vsub(ray.o, p1);
    Flt b1 = vdot(d, s1) * invdivisor;
    if ((b1 < 0) || (b1 > 1)) return false;
    Vec s2 = vcross(d, e1);
    Flt b2 = vdot(ray.d, s2) * invdivisor;
    if ((b2 < 0) || (b1 + b2 > 1)) return false;
    Flt tt = vdot(e2, s2) * invdivisor;
    return (tt >= 0) && (tt <= dist);
}

template <bool FULL>
GD_FN bool prim_box(const Flt* __restrict__ p, const Ray& r, Flt d, Flt& t, Vec& pos, Vec& n) {
    // Box.hs:18-54
    Bbox b = ldbb(p);
    Flt dx = r.d.x, dy = r.d.y, dz = r.d.z;
    Flt dxrcp = 1 / dx, dyrcp = 1 / dy, dzrcp = 1 / dz;
    Flt inx, outx, iny, outy, inz, outz;
    slab(dx > 0, b.p1.x, b.p2.x, r.o.x, dxrcp, inx, outx);
    slab(dy > 0, b.p1.y, b.p2.y, r.o.y, dyrcp, iny, outy);
    slab(dz > 0, b.p1.z, b.p2.z, r.o.z, dzrcp, inz, outz);
    Flt lastin = fmax3(inx, iny, inz);
    Flt firstout = fmin3(outx, outy, outz);
    if (lastin > firstout || firstout < 0 || lastin > d) return false;
    if (lastin < 0) {  // origin is inside
        t = firstout;
        if (FULL) {
            if (outx == firstout) n = (dx > 0) ? vec(1, 0, 0) : vec(-1, 0, 0);
            else if (outy == firstout) n = (dy > 0) ? vec(0, 1, 0) : vec(0, -1, 0);
            else n = (dz > 0) ? vec(0, 0, 1) : vec(0, 0, -1);
            pos = vscaleadd(r.o, r.d, firstout);
        }
    } else {
        t = lastin;
        if (FULL) {
            if (inx == lastin) n = (dx > 0) ? vec(-1, 0, 0) : vec(1, 0, 0);
            else if (iny == lastin) n = (dy > 0) ? vec(0, -1, 0) : vec(0, 1, 0);
            else n = (dz > 0) ? vec(0, 0, -1) : vec(0, 0, 1);
            pos = vscaleadd(r.o, r.d, lastin);
        }
    }
    return true;
}
// the same, with 1/d handed in: the three quotients depend on the ray only, so a caller testing many boxes with one ray
// (a group of boxes, a BIH) computes them once.  Bit-identical to prim_box.
template <bool FULL>
GD_FN bool prim_box_rcp(const Flt* __restrict__ p, const Ray& r, Flt dxrcp, Flt dyrcp, Flt dzrcp, Flt d, Flt& t, Vec& pos, Vec& n) {
    Bbox b = ldbb(p);
    Flt dx = r.d.x, dy = r.d.y, dz = r.d.z;
    Flt inx, outx, iny, outy, inz, outz;
    slab(dx > 0, b.p1.x, b.p2.x, r.o.x, dxrcp, inx, outx);
    slab(dy > 0, b.p1.y, b.p2.y, r.o.y, dyrcp, iny, outy);
    slab(dz > 0, b.p1.z, b.p2.z, r.o.z, dzrcp, inz, outz);
    Flt lastin = fmax3(inx, iny, inz);
    Flt firstout = fmin3(outx, outy, outz);
    if (lastin > firstout || firstout < 0 || lastin > d) return false;
    if (lastin < 0) {  // origin is inside
        t = firstout;
        if (FULL) {
            if (outx == firstout) n = (dx > 0) ? vec(1, 0, 0) : vec(-1, 0, 0);
            else if (outy == firstout) n = (dy > 0) ? vec(0, 1, 0) : vec(0, -1, 0);
            else n = (dz > 0) ? vec(0, 0, 1) : vec(0, 0, -1);
            pos = vscaleadd(r.o, r.d, firstout);
        }
    } else {
        t = lastin;
        if (FULL) {
            if (inx == lastin) n = (dx > 0) ? vec(-1, 0, 0) : vec(1, 0, 0);
            else if (iny == lastin) n = (dy > 0) ? vec(0, -1, 0) : vec(0, 1, 0);
            else n = (dz > 0) ? vec(0, 0, -1) : vec(0, 0, 1);
            pos = vscaleadd(r.o, r.d, lastin);
        }
    }
    return true;
}
// bbclip_ub (Vec.hs:743-762) with 1/d handed in
GD_FN void bbclip_ub_pre(const Ray& r, Flt dxrcp, Flt dyrcp, Flt dzrcp, const Bbox& b, Flt& near_, Flt& far_) {
    Flt inx, outx, iny, outy, inz, outz;
    slab(r.d.x > 0, b.p1.x, b.p2.x, r.o.x, dxrcp, inx, outx);
    slab(r.d.y > 0, b.p1.y, b.p2.y, r.o.y, dyrcp, iny, outy);
    slab(r.d.z > 0, b.p1.z, b.p2.z, r.o.z, dzrcp, inz, outz);
    near_ = fmax3(inx, iny, inz);
    far_ = fmin3(outx, outy, outz);
}
GD_FN bool shadow_box_rcp(const Flt* __restrict__ p, const Ray& r, Flt dxrcp, Flt dyrcp, Flt dzrcp, Flt d) {  // Box.hs:56-62
    Bbox b = ldbb(p);
    Flt near_, far_;
    bbclip_ub_pre(r, dxrcp, dyrcp, dzrcp, b, near_, far_);
    if ((near_ > far_) || far_ <= 0 || far_ > d) return false;
    return true;
}
GD_FN bool shadow_box(const Flt* __restrict__ p, const Ray& r, Flt d) {  // Box.hs:56-62
    Bbox b = ldbb(p);
    Flt near_, far_;
    bbclip_ub(r, b, near_, far_);
    if ((near_ > far_) || far_ <= 0 || far_ > d) return false;
    return true;
}

template <bool FULL>
GD_FN bool prim_plane(const Flt* __restrict__ p, const Ray& ray, Flt d, Flt& t, Vec& pos, Vec& n) {
    // Plane.hs:27-32
    Vec norm = ldv(p);
    Flt offset = p[3];
    Flt hit = -((vdot(norm, ray.o) - offset) / vdot(norm, ray.d));
    if (hit < 0 || hit > d) return false;
    t = hit;
    if (FULL) { pos = vscaleadd(ray.o, ray.d, hit); n = norm; }
    return true;
}

template <bool FULL>
GD_FN bool prim_disc_v(const Vec& point, const Vec& norm, Flt radius_sqr, const Ray& r, Flt d, Flt& t,
                                            Vec& pos, Vec& n) {
    // Cone.hs:69-79
    Flt dist = plane_int_dist(r, point, norm);
    if (dist < 0 || dist > d) return false;
    Vec p = vscaleadd(r.o, r.d, dist);
    Vec offset = vsub(p, point);
    if (vdot(offset, offset) > radius_sqr) return false;
    t = dist;
    if (FULL) { pos = p; n = norm; }
    return true;
}

template <bool FULL>
GD_FN bool prim_cylinder(const Flt* __restrict__ p, const Ray& ray, Flt d, Flt& t, Vec& pos, Vec& n) {
    // Cone.hs:104-139
    Flt r = p[0], h1 = p[1], h2 = p[2];
    Flt ox = ray.o.x, oy = ray.o.y, oz = ray.o.z, dx = ray.d.x, dy = ray.d.y, dz = ray.d.z;
    Flt a = dx * dx + dy * dy;
    Flt b = 2 * (dx * ox + dy * oy);
    Flt c = ox * ox + oy * oy - r * r;
    Flt disc = b * b - 4 * a * c;
    if (disc < 0) return false;
    Flt discsqrt = sqrt(disc);
    Flt q = (b < 0) ? (b - discsqrt) * FL(-0.5) : (b + discsqrt) * FL(-0.5);
    Flt t0p = q / a;
    Flt t1p = c / q;
    Flt t0 = fmin_(t0p, t1p);
    Flt t1 = fmax_(t0p, t1p);
    if (t1 < 0 || t0 > d) return false;
    Flt dist = (t0 < 0) ? t1 : t0;
    if (dist < 0 || dist > d) return false;
    Vec ps = vscaleadd(ray.o, ray.d, dist);
    if (ps.z > h1 && ps.z < h2) {
        t = dist;
        if (FULL) { pos = ps; n = vec(ps.x / r, ps.y / r, 0); }
        return true;
    }
    if (dz > 0) {
        if (oz < h1) return prim_disc_v<FULL>(vec(0, 0, h1), vec(0, 0, -1), r * r, ray, d, t, pos, n);
        return false;
    }
    if (oz > h2) return prim_disc_v<FULL>(vec(0, 0, h2), vec(0, 0, 1), r * r, ray, d, t, pos, n);
    return false;
}

// rayint_cone (Cone.hs:155-204) and shadow_cone (Cone.hs:206-245) share everything but the result
template <bool FULL>
GD_FN bool prim_cone(const Flt* __restrict__ p, const Ray& ray, Flt d, Flt& t, Vec& pos, Vec& n) {
    Flt r = p[0], clip1 = p[1], clip2 = p[2], height = p[3];
    Flt ox = ray.o.x, oy = ray.o.y, oz = ray.o.z, dx = ray.d.x, dy = ray.d.y, dz = ray.d.z;
    Flt kp = r / height;
    Flt k = kp * kp;
    Flt a = dx * dx + dy * dy - k * dz * dz;
    Flt b = 2 * (dx * ox + dy * oy - k * dz * (oz - height));
    Flt c = ox * ox + oy * oy - k * (oz - height) * (oz - height);
    Flt disc = b * b - 4 * a * c;
    if (disc < 0) return false;
    Flt discsqrt = sqrt(disc);
    Flt q = (b < 0) ? (b - discsqrt) * FL(-0.5) : (b + discsqrt) * FL(-0.5);
    Flt t0p = q / a;
    Flt t1p = c / q;
    Flt t0 = fmin_(t0p, t1p);
    Flt t1 = fmax_(t0p, t1p);
    if (t1 < 0 || t0 > d) return false;
    Flt dist = (t0 < 0) ? t1 : t0;
    if (dist < 0 || dist > d) return false;
    Vec ps = vscaleadd(ray.o, ray.d, dist);
    if (ps.z > clip1 && ps.z < clip2) {
        t = dist;
        if (FULL) {
            Flt invhyp = 1 / sqrt(height * height + r * r);
            Flt up = r * invhyp;
            Flt out = height * invhyp;
            Flt r_ = sqrt(ps.x * ps.x + ps.y * ps.y);
            Flt correction = out / r_;
            pos = ps;
            n = vec(ps.x * correction, ps.y * correction, up);
        }
        return true;
    }
    if (dz > 0) {
        if (oz < clip1) return prim_disc_v<FULL>(vec(0, 0, clip1), vec(0, 0, -1), r * r, ray, d, t, pos, n);
        return false;
    }
    if (oz > clip2) {
        Flt r2 = r * (1 - ((clip2 - clip1) / height));
        return prim_disc_v<FULL>(vec(0, 0, clip2), vec(0, 0, 1), r2 * r2, ray, d, t, pos, n);
    }
    return false;
}

GD_FN bool is_prim(int type) { return type >= GLOME_SPHERE && type <= GLOME_CONE; }

// rayint of a primitive node
template <bool FULL>
GD_FN bool prim_rayint(const DScene& S, const GlomeNode& nd, const Ray& r, Flt d, Flt& t, Vec& pos,
                                            Vec& n) {
    const Flt* p = S.dpool + nd.a;
    switch (nd.type) {
        case GLOME_SPHERE: return prim_sphere<FULL>(p, r, d, t, pos, n);
        case GLOME_TRIANGLE: {
            Vec z = vec(0, 0, 0);
            return prim_triangle<FULL>(ldv(p), ldv(p + 3), ldv(p + 6), false, z, z, z, r, d, t, pos, n);
        }
        case GLOME_TRIANGLENORM:
            return prim_triangle<FULL>(ldv(p), ldv(p + 3), ldv(p + 6), true, ldv(p + 9), ldv(p + 12), ldv(p + 15), r, d, t,
                                       pos, n);
        case GLOME_BOX: return prim_box<FULL>(p, r, d, t, pos, n);
        case GLOME_PLANE: return prim_plane<FULL>(p, r, d, t, pos, n);
        case GLOME_DISC: return prim_disc_v<FULL>(ldv(p), ldv(p + 3), p[6], r, d, t, pos, n);
        case GLOME_CYLINDER: return prim_cylinder<FULL>(p, r, d, t, pos, n);
        case GLOME_CONE: return prim_cone<FULL>(p, r, d, t, pos, n);
    }
    return false;
}
// shadow of a primitive node (default = rayint hit, Solid.hs:218-221)
GD_FN bool prim_shadow(const DScene& S, const GlomeNode& nd, const Ray& r, Flt d) {
    const Flt* p = S.dpool + nd.a;
    Flt t;
    Vec a, b;
    switch (nd.type) {
        case GLOME_SPHERE: return shadow_sphere(p, r, d);
        case GLOME_TRIANGLE:
        case GLOME_TRIANGLENORM: return shadow_triangle(ldv(p), ldv(p + 3), ldv(p + 6), r, d);
        case GLOME_BOX: return shadow_box(p, r, d);
        case GLOME_PLANE: return prim_plane<false>(p, r, d, t, a, b);
        case GLOME_DISC: return prim_disc_v<false>(ldv(p), ldv(p + 3), p[6], r, d, t, a, b);  // Cone.hs:81-91
        case GLOME_CYLINDER: return prim_cylinder<false>(p, r, d, t, a, b);
        case GLOME_CONE: return prim_cone<false>(p, r, d, t, a, b);  // shadow_cone == rayint_cone's hit test
    }
    return false;
}
GD_FN bool prim_inside(const DScene& S, const GlomeNode& nd, const Vec& pt) {
    const Flt* p = S.dpool + nd.a;
    switch (nd.type) {
        case GLOME_SPHERE: {  // Sphere.hs:73-76
            Vec offset = vsub(ldv(p), pt);
            return vdot(offset, offset) < p[3] * p[3];
        }
        case GLOME_BOX:  // Box.hs:64-68
            return pt.x > p[0] && pt.x < p[3] && pt.y > p[1] && pt.y < p[4] && pt.z > p[2] && pt.z < p[5];
        case GLOME_PLANE: {  // Plane.hs:34-38
            Vec norm = ldv(p);
            Vec onplane = vscale(norm, p[3]);
            Vec newvec = vsub(onplane, pt);
            return vdot(newvec, norm) > 0;
        }
        case GLOME_CYLINDER:  // Cone.hs:141-143
            return pt.z > p[1] && pt.z < p[2] && pt.x * pt.x + pt.y * pt.y < p[0] * p[0];
        case GLOME_CONE: {  // Cone.hs:248-251
            Flt r = p[0] * (1 - ((pt.z - p[1]) / p[3]));
            return pt.z > p[1] && pt.z < p[2] && pt.x * pt.x + pt.y * pt.y < r * r;
        }
    }
    return false;  // Triangle.hs:183, Cone.hs:100
}

// ---------------------------------------------------------------------------------------------
// the flat-scene interpreter
// ---------------------------------------------------------------------------------------------
// Flat-class scenes only ({Tex,Tag}* prim | Bih of those | Mesh | one group of those): three fixed levels, everything
// inlines, no recursion.  General scene graphs (Instance, CSG, Bound, nested Bih ...) are evaluated by the iterative
// machine of glome_gen.cuh.
template <int L>
GD_FN void rayint_node(const DScene& S, int ni, const Ray& r, Flt d, const Stk& tex, const Stk& tag, int csg, Hit& acc,
                       Cnt* cnt = nullptr);
template <int L>
GD_FN bool shadow_node(const DScene& S, int ni, const Ray& r, Flt d, int csg, Cnt* cnt = nullptr);

// record a winning primitive hit
GD_FN void take_hit(Hit& acc, Flt t, const Vec& pos, const Vec& n, const Ray& r, const Stk& tex,
                                         const Stk& tag, int prim, int sub) {
    acc.hit = 1; acc.t = t; acc.pos = pos; acc.norm = n; acc.ray = r; acc.tex = tex; acc.tag = tag;
    acc.prim = prim; acc.sub = sub;
}

struct TravEnt { int ref; Flt near_, far_; };

// rayint_bih (Bih.hs:332-368), iterative, reference order, best-hit culling
template <int L>
GD_FN void rayint_bih(const DScene& S, const GlomeNode& nd, const Ray& r, Flt d, const Stk& tex,
                                           const Stk& tag, int csg, Hit& acc, Cnt* cnt = nullptr) {
    Bbox bb = ldbb(S.dpool + nd.b);
    Flt near_, far_;
    bbclip_ub(r, bb, near_, far_);
    Flt dirr[3] = {1 / r.d.x, 1 / r.d.y, 1 / r.d.z};
    Flt org[3] = {r.o.x, r.o.y, r.o.z};
    far_ = fmin_(d, far_);  // traverse root near (fmin d far)  (Bih.hs:368)
    if (near_ < 0) near_ = 0;  // origin clamp: a subtree entirely behind the origin holds no valid hit
    TravEnt stack[GDEV_BIH_STACK];
    int sp = 0;
    int ref = nd.a;
    constexpr bool GEN = (L < 0);
    constexpr int LI = GEN ? -1 : 2;  // leaf items of a flat-class BIH are {Tex,Tag}* prim
    const bool linear = (nd.c & GLOME_BIH_LINEAR_SPHERES) != 0;
    // linear sphere block: item j's payload is at dpool[a0 + 4*(j - j0)]
    const int j0 = nd.c >> 4;
    const int a0 = linear ? S.nodes[j0].a : 0;
    for (;;) {
        bool pop = false;
        if (ref < 0) {
            // BihLeaf s -> rayint s r far t tags: list fold with the clipped far as max distance
            int2 lf;
            glome_bih_leaf(ref, S.ipool, &lf.x, &lf.y);
            for (int i = 0; i < lf.y; i++) {
                int item = lf.x + i;
                if (linear) {  // bare spheres: no node record to chase
                    if (cnt) cnt->prim++;
                    Flt t; Vec pos, n;
                    if (prim_sphere<GEN>(S.dpool + a0 + 4 * (item - j0), r, far_, t, pos, n) && cand_wins(acc, t))
                        take_hit(acc, t, pos, n, r, tex, tag, item, -1);
                } else {
                    rayint_node<LI>(S, item, r, far_, tex, tag, csg, acc, cnt);
                }
            }
            pop = true;
        } else {
            // 32-byte node: two 16-byte loads
            if (cnt) cnt->bih++;
            const BihStep bs_ = ld_bih(S.bih, ref);
            const Flt2 sp2 = {bs_.ls, bs_.rs};
            int4 ii; ii.x = bs_.axis; ii.y = bs_.left; ii.z = bs_.right; ii.w = 0;
            int axis = ii.x;
            Flt dr_ = dirr[axis], o = org[axis];
            Flt dl = (sp2.x - o) * dr_;
            Flt dr = (sp2.y - o) * dr_;
            if (near_ > far_) pop = true;
            else {
                int c1, c2;
                bool v1, v2;
                Flt n2;
                Flt f1;
                if (dr_ > 0) {
                    c1 = ii.y; v1 = near_ < dl; f1 = fmin_(dl, far_);
                    c2 = ii.z; v2 = dr < far_; n2 = fmax_(dr, near_);
                } else {
                    c1 = ii.z; v1 = near_ < dr; f1 = fmin_(dr, far_);
                    c2 = ii.y; v2 = dl < far_; n2 = fmax_(dl, near_);
                }
                if (v2 && acc.hit && n2 > acc.t) v2 = false;  // best-hit culling
                if (v1) {
                    if (v2) {
                        if (sp < GDEV_BIH_STACK) { stack[sp].ref = c2; stack[sp].near_ = n2; stack[sp].far_ = far_; sp++; }
                        else acc.flags |= GLOME_HITFLAG_STACK_OVERFLOW;
                    }
                    ref = c1; far_ = f1;
                } else if (v2) {
                    ref = c2; near_ = n2;
                } else pop = true;
            }
        }
        if (pop) {
            for (;;) {
                if (sp == 0) return;
                sp--;
                ref = stack[sp].ref; near_ = stack[sp].near_; far_ = stack[sp].far_;
                if (!(acc.hit && near_ > acc.t)) break;  // best-hit culling
            }
        }
    }
}

// shadow_bih (Bih.hs:510-544)
template <int L>
GD_FN bool shadow_bih(const DScene& S, const GlomeNode& nd, const Ray& r, Flt d, int csg,
                                           Cnt* cnt = nullptr) {
    Bbox bb = ldbb(S.dpool + nd.b);
    Flt near_, farp;
    bbclip_ub(r, bb, near_, farp);
    Flt far_ = fmin_(d, farp);
    if (near_ < 0) near_ = 0;  // origin clamp (see rayint_bih)
    Flt dirr[3] = {1 / r.d.x, 1 / r.d.y, 1 / r.d.z};
    Flt org[3] = {r.o.x, r.o.y, r.o.z};
    TravEnt stack[GDEV_BIH_STACK];
    int sp = 0;
    int ref = nd.a;
    constexpr int LI = (L < 0) ? -1 : 2;
    const bool linear = (nd.c & GLOME_BIH_LINEAR_SPHERES) != 0;
    const int j0 = nd.c >> 4;
    const int a0 = linear ? S.nodes[j0].a : 0;
    for (;;) {
        bool pop = false;
        if (ref < 0) {
            int2 lf;
            glome_bih_leaf(ref, S.ipool, &lf.x, &lf.y);
            Flt dd = fmin_(d, far_);  // shadow s r (fmin d far)  (Bih.hs:515)
            for (int i = 0; i < lf.y; i++) {
                if (linear) {
                    if (cnt) cnt->prim++;
                    if (shadow_sphere(S.dpool + a0 + 4 * (lf.x + i - j0), r, dd)) return true;
                } else if (shadow_node<LI>(S, lf.x + i, r, dd, csg, cnt)) return true;
            }
            pop = true;
        } else {
            if (cnt) cnt->bih++;
            const BihStep bs_ = ld_bih(S.bih, ref);
            const Flt2 sp2 = {bs_.ls, bs_.rs};
            int4 ii; ii.x = bs_.axis; ii.y = bs_.left; ii.z = bs_.right; ii.w = 0;
            int axis = ii.x;
            Flt dr_ = dirr[axis], o = org[axis];
            Flt dl = (sp2.x - o) * dr_;
            Flt dr = (sp2.y - o) * dr_;
            if (near_ > far_) pop = true;
            else {
                int c1, c2;
                bool v1, v2;
                Flt n2, f1;
                if (dr_ > 0) {
                    c1 = ii.y; v1 = near_ < dl; f1 = fmin_(dl, far_);
                    c2 = ii.z; v2 = dr < far_; n2 = fmax_(dr, near_);
                } else {
                    c1 = ii.z; v1 = near_ < dr; f1 = fmin_(dr, far_);
                    c2 = ii.y; v2 = dl < far_; n2 = fmax_(dl, near_);
                }
                if (v1) {
                    if (v2 && sp < GDEV_BIH_STACK) { stack[sp].ref = c2; stack[sp].near_ = n2; stack[sp].far_ = far_; sp++; }
                    ref = c1; far_ = f1;
                } else if (v2) {
                    ref = c2; near_ = n2;
                } else pop = true;
            }
        }
        if (pop) {
            if (sp == 0) return false;
            sp--;
            ref = stack[sp].ref; near_ = stack[sp].near_; far_ = stack[sp].far_;
        }
    }
}

// rayint_mesh (Mesh.hs:136-198), iterative.  The reference culls the second child by the first
// child's result depth; we cull by the best hit so far (a superset of that knowledge) while
// passing the unculled `far` down exactly as the reference does (Mesh.hs:178-184).
GD_FN void rayint_mesh(const DScene& S, int ni, const GlomeNode& nd, const Ray& ray, Flt depth,
                                            const Stk& texs, const Stk& tags, bool full, Hit& acc, Cnt* cnt = nullptr) {
    const GlomeMeshHeader* h = reinterpret_cast<const GlomeMeshHeader*>(S.ipool + nd.a);
    const int bb_off = h->bb_off, verts_off = h->verts_off, norms_off = h->norms_off, tris_off = h->tris_off;
    const int texs_off = h->texs_off, tags_off = h->tags_off;
    Bbox bb = ldbb(S.dpool + bb_off);
    Vec rcp = vrcp(ray.d);
    Flt near_, far_;
    bbclip_ub_rcp(ray.o, rcp, bb, near_, far_);
    if (near_ > far_ || near_ > depth || far_ < 0) return;
    TravEnt stack[GDEV_BIH_STACK];
    int sp = 0;
    int ref = h->root;
    for (;;) {
        bool pop = false;
        if (ref < 0) {
            int k = ~ref;
            int ntri = gd_ldg(S.ipool + k);
            for (int j = 0; j < ntri; j++) {
                int ti = gd_ldg(S.ipool + k + 1 + j);
                if (cnt) cnt->tri++;
                const int4* tp = reinterpret_cast<const int4*>(S.ipool + tris_off + 8 * ti);
                int4 t0 = gd_ldg(tp), t1 = gd_ldg(tp + 1);  // {a,b,c,na} {nb,nc,tex,tag}
                Vec a = ldv(S.dpool + verts_off + 3 * t0.x);
                Vec b = ldv(S.dpool + verts_off + 3 * t0.y);
                Vec c = ldv(S.dpool + verts_off + 3 * t0.z);
                bool smooth = t0.w != -1;
                Vec an = vec(0, 0, 0), bn = an, cn = an;
                if (smooth && full) {
                    an = ldv(S.dpool + norms_off + 3 * t0.w);
                    bn = ldv(S.dpool + norms_off + 3 * t1.x);
                    cn = ldv(S.dpool + norms_off + 3 * t1.y);
                }
                Flt t; Vec pos, n;
                bool hitp = full ? prim_triangle<true>(a, b, c, smooth, an, bn, cn, ray, far_, t, pos, n)
                                 : prim_triangle<false>(a, b, c, smooth, an, bn, cn, ray, far_, t, pos, n);
                if (hitp && cand_wins(acc, t)) {
                    Stk tx = texs, tg = tags;
                    if (t1.z != -1) { if (stk_cons(tx, gd_ldg(S.ipool + texs_off + t1.z), texs)) acc.flags |= GLOME_HITFLAG_STACK_OVERFLOW; }
                    if (t1.w != -1) { if (stk_cons(tg, gd_ldg(S.ipool + tags_off + t1.w), tags)) acc.flags |= GLOME_HITFLAG_STACK_OVERFLOW; }
                    take_hit(acc, t, pos, n, ray, tx, tg, ni, ti);
                }
            }
            pop = true;
        } else {
            // 128-byte node: two boxes + two child refs
            if (cnt) cnt->bvh++;
            Bbox lbb, rbb;
            int2 kids;
            ld_bvh(S.bvh, ref, lbb, rbb, kids.x, kids.y);
            Flt lnearp, lfarp, rnearp, rfarp;
            bbclip_ub_rcp(ray.o, rcp, lbb, lnearp, lfarp);
            bbclip_ub_rcp(ray.o, rcp, rbb, rnearp, rfarp);
            Flt lnear = hmax(near_, lnearp);
            Flt lfar = hmin(far_, lfarp);
            Flt rnear = hmax(near_, rnearp);
            Flt rfar = hmin(far_, rfarp);
            Flt best = ridepth(acc);
            int c1, c2;
            Flt n1, f1, n2, f2;
            if (lnear < rnear) { c1 = kids.x; n1 = lnear; f1 = lfar; c2 = kids.y; n2 = rnear; f2 = rfar; }
            else { c1 = kids.y; n1 = rnear; f1 = rfar; c2 = kids.x; n2 = lnear; f2 = lfar; }
            // first child: Mesh.hs:175 / 187; culled additionally by the best hit so far
            bool v1 = !(n1 > f1 || n1 > depth || f1 < 0) && !(acc.hit && n1 > best);
            // second child: entry test uses far culled by the (first child's) depth, Mesh.hs:178-182
            Flt f2c = hmin(f2, best);
            bool v2 = !(n2 > f2c || n2 > depth || f2c < 0);
            if (v1) {
                if (v2) {
                    if (sp < GDEV_BIH_STACK) { stack[sp].ref = c2; stack[sp].near_ = n2; stack[sp].far_ = f2; sp++; }
                    else acc.flags |= GLOME_HITFLAG_STACK_OVERFLOW;
                }
                ref = c1; near_ = n1; far_ = f1;
            } else if (v2) {
                ref = c2; near_ = n2; far_ = f2;
            } else pop = true;
        }
        if (pop) {
            for (;;) {
                if (sp == 0) return;
                sp--;
                ref = stack[sp].ref; near_ = stack[sp].near_; far_ = stack[sp].far_;
                // re-evaluate the second child's entry test with the hits found since it was pushed
                Flt f2c = hmin(far_, ridepth(acc));
                if (!(near_ > f2c || f2c < 0)) break;
            }
        }
    }
}

// General path: keep the traversal stacks out of the recursive rayint_node / shadow_node frames.

// class Solid: rayint (Solid.hs:146) for flat-class scenes.  Folds the node's result into acc with `nearest`.
// L = 0: scene root, 1: group child, 2: BIH leaf item.
template <int L>
GD_FN void rayint_node(const DScene& S, int ni, const Ray& r, Flt d, const Stk& tex_in, const Stk& tag_in, int csg, Hit& acc,
                       Cnt* cnt) {
    Stk tex = tex_in, tag = tag_in;
    GlomeNode nd = S.nodes[ni];
    // Tex / Tag / NoShadow wrappers: push and descend (Tex.hs:54,66,78)
    while (nd.type == GLOME_TEX || nd.type == GLOME_TAG || nd.type == GLOME_NOSHADOW) {
        if (nd.type == GLOME_TEX) { Stk t2; if (stk_cons(t2, nd.b, tex)) acc.flags |= GLOME_HITFLAG_STACK_OVERFLOW; tex = t2; }
        else if (nd.type == GLOME_TAG) { Stk t2; if (stk_cons(t2, nd.b, tag)) acc.flags |= GLOME_HITFLAG_STACK_OVERFLOW; tag = t2; }
        ni = nd.a;
        nd = S.nodes[ni];
    }
    if (is_prim(nd.type)) {
        Flt t; Vec pos, n;
        if (cnt) cnt->prim++;
        if (prim_rayint<false>(S, nd, r, d, t, pos, n) && cand_wins(acc, t)) take_hit(acc, t, pos, n, r, tex, tag, ni, -1);
        return;
    }
    switch (nd.type) {
        case GLOME_VOID:
        case GLOME_ONLYSHADOW: return;  // Solid.hs:354, Tex.hs:89
        case GLOME_BIH:
            if constexpr (L == 0 || L == 1) rayint_bih<L>(S, nd, r, d, tex, tag, csg, acc, cnt);
            return;
        case GLOME_MESH:
            if constexpr (L == 0 || L == 1) rayint_mesh(S, ni, nd, r, d, tex, tag, false, acc, cnt);
            return;
        case GLOME_GROUP:  // Solid.hs:327
            if constexpr (L == 0) {
                for (int i = 0; i < nd.b; i++) rayint_node<1>(S, nd.a + i, r, d, tex, tag, csg, acc, cnt);
            }
            return;
    }
}

// class Solid: shadow (Solid.hs:162) for flat-class scenes
template <int L>
GD_FN bool shadow_node(const DScene& S, int ni, const Ray& r, Flt d, int csg, Cnt* cnt) {
    GlomeNode nd = S.nodes[ni];
    while (nd.type == GLOME_TEX || nd.type == GLOME_TAG || nd.type == GLOME_ONLYSHADOW) {  // Tex.hs:57,69,92
        ni = nd.a;
        nd = S.nodes[ni];
    }
    if (is_prim(nd.type)) {
        if (cnt) cnt->prim++;
        return prim_shadow(S, nd, r, d);
    }
    switch (nd.type) {
        case GLOME_VOID:
        case GLOME_NOSHADOW:
        case GLOME_MESH: return false;  // Solid.hs:356, Tex.hs:81, Mesh.hs:210
        case GLOME_BIH:
            if constexpr (L == 0 || L == 1) return shadow_bih<L>(S, nd, r, d, csg, cnt);
            return false;
        case GLOME_GROUP:  // Solid.hs:330
            if constexpr (L == 0) {
                for (int i = 0; i < nd.b; i++)
                    if (shadow_node<1>(S, nd.a + i, r, d, csg, cnt)) return true;
            }
            return false;
    }
    return false;
}

// Flat scenes record t / prim / stacks while traversing and compute position + normal once, for
// the winner (same arithmetic as the eager path: prim_rayint<true> on the winning primitive).
GD_FN void finalize_flat(const DScene& S, const Ray& r, Hit& h) {
    if (!h.hit) return;
    GlomeNode nd = S.nodes[h.prim];
    Flt t; Vec pos, n;
    const Flt big = GLM_BIG;
    if (nd.type == GLOME_MESH) {
        const GlomeMeshHeader* mh = reinterpret_cast<const GlomeMeshHeader*>(S.ipool + nd.a);
        const int32_t* T = S.ipool + mh->tris_off + 8 * h.sub;
        Vec a = ldv(S.dpool + mh->verts_off + 3 * T[0]);
        Vec b = ldv(S.dpool + mh->verts_off + 3 * T[1]);
        Vec c = ldv(S.dpool + mh->verts_off + 3 * T[2]);
        bool smooth = T[3] != -1;
        Vec an = vec(0, 0, 0), bn = an, cn = an;
        if (smooth) {
            an = ldv(S.dpool + mh->norms_off + 3 * T[3]);
            bn = ldv(S.dpool + mh->norms_off + 3 * T[4]);
            cn = ldv(S.dpool + mh->norms_off + 3 * T[5]);
        }
        prim_triangle<true>(a, b, c, smooth, an, bn, cn, r, big, t, pos, n);
#ifdef GLOME_F32
        if (h.t > 1) {  // (see below)
            const Ray r2 = mkray(vscaleadd(r.o, r.d, h.t - 1), r.d);
            Flt t2; Vec p2, n2;
            if (prim_triangle<true>(a, b, c, smooth, an, bn, cn, r2, big, t2, p2, n2)) { pos = p2; n = n2; }
        }
#endif
    } else {
        prim_rayint<true>(S, nd, r, big, t, pos, n);
#ifdef GLOME_F32
        // FP32 only.  A depth of a few hundred units carries an error of a few float ulps of t -- about the reference's
        // delta (1e-4, Vec.hs:40), which is what lifts a shadow ray off the surface it starts on (Shader.hs:76): the point
        // would land inside its own primitive on a good share of the pixels ("shadow acne" that the Double reference does
        // not have).  The winner is therefore intersected once more from one unit in front of the hit, where the same
        // test resolves the surface to ~1e-7; the depth that took part in the nearest-hit fold stays as it was.
        if (h.t > 1) {
            const Ray r2 = mkray(vscaleadd(r.o, r.d, h.t - 1), r.d);
            Flt t2; Vec p2, n2;
            if (prim_rayint<true>(S, nd, r2, big, t2, p2, n2)) { pos = p2; n = n2; }
        }
#endif
    }
    h.pos = pos; h.norm = n; h.ray = r;
}

// rayint sld ray d [] []  (flat-class scenes)
GD_FN void rayint_scene_flat(const DScene& S, int sld, const Ray& r, Flt d, Hit& h, Cnt* cnt = nullptr) {
    hit_clear(h);
    Stk e;
    stk_clear(e);
    rayint_node<0>(S, sld, r, d, e, e, 0, h, cnt);
    finalize_flat(S, r, h);
}
GD_FN bool shadow_scene_flat(const DScene& S, int sld, const Ray& r, Flt d, Cnt* cnt = nullptr) {
    return shadow_node<0>(S, sld, r, d, 0, cnt);
}

// ---------------------------------------------------------------------------------------------
// Clr.hs, Texture.hs
// ---------------------------------------------------------------------------------------------
struct Color { Flt r, g, b; };
struct ColorA { Flt r, g, b, a; };
GD_FN ColorA mkca(Flt r, Flt g, Flt b, Flt a) { ColorA c; c.r = r; c.g = g; c.b = b; c.a = a; return c; }
GD_FN Flt aclamp(Flt x) { if (x > 1) return 1; if (x < 0) return 0; return x; }  // Clr.hs:75
GD_FN ColorA caweight(const ColorA& x, const ColorA& y, Flt w) {                 // Clr.hs:87
    return mkca((x.r * w) + (y.r * (1 - w)), (x.g * w) + (y.g * (1 - w)), (x.b * w) + (y.b * (1 - w)),
                (x.a * w) + (y.a * (1 - w)));
}
GD_FN ColorA cafold(const ColorA& x, const ColorA& y) {  // Clr.hs:106
    Flt trans = 1 - x.a;
    return mkca(x.r + (y.r * trans * y.a), x.g + (y.g * trans * y.a), x.b + (y.b * trans * y.a), x.a + (y.a * trans));
}
GD_FN Flt triangle_wave(Flt x) {  // Texture.hs:16
    Flt offset = x - floor(x);
    return (offset < FL(0.5)) ? (offset * 2) : (2 - (offset * 2));
}
GD_FN Flt omega(Flt t_) {  // Texture.hs:49
    Flt t = fabs_(t_);
    Flt tsqr = t * t;
    Flt tcube = tsqr * t;
    return (-6) * tcube * tsqr + 15 * tcube * t - 10 * tcube + 1;
}
// phi (Texture.hs:57) and the twelve gradient vectors (Texture.hs:60), 4 bits / 2 bits per entry
GD_FN int phi12(int i) { return (int)((0x69A8B5147203ULL >> (4 * i)) & 15); }
GD_FN void grad12(int c, Flt& gx, Flt& gy, Flt& gz) {
    // {-1,-1,0} {-1,0,-1} {-1,0,1} {-1,1,0} {0,-1,-1} {0,-1,1} {0,1,-1} {0,1,1} {1,-1,0} {1,0,-1} {1,0,1} {1,1,0}
    const int gxs[12] = {-1, -1, -1, -1, 0, 0, 0, 0, 1, 1, 1, 1};
    const int gys[12] = {-1, 0, 0, 1, -1, -1, 1, 1, -1, 0, 0, 1};
    const int gzs[12] = {0, -1, 1, 0, -1, 1, -1, 1, 0, -1, 1, 0};
    gx = (Flt)gxs[c]; gy = (Flt)gys[c]; gz = (Flt)gzs[c];
}
GD_FN long long iabs64(long long a) { return a < 0 ? -a : a; }
GD_FN Flt knot(long long i, long long j, long long k, const Vec& v) {  // Texture.hs:67-77
    int a = phi12((int)(iabs64(k) % 12));
    int b = phi12((int)(iabs64(j + a) % 12));
    int c = phi12((int)(iabs64(i + b) % 12));
    Vec g;
    grad12(c, g.x, g.y, g.z);
    return omega(v.x) * omega(v.y) * omega(v.z) * vdot(g, v);
}
GD_FN Flt noise(const Vec& p) {  // Texture.hs:92-107
    Flt fx = floor(p.x), fy = floor(p.y), fz = floor(p.z);
    long long i = (long long)fx, j = (long long)fy, k = (long long)fz;
    Flt u = p.x - fx, v = p.y - fy, w = p.z - fz;
    return knot(i, j, k, vec(u, v, w)) + knot(i + 1, j, k, vec(u - 1, v, w)) + knot(i, j + 1, k, vec(u, v - 1, w)) +
           knot(i, j, k + 1, vec(u, v, w - 1)) + knot(i + 1, j + 1, k, vec(u - 1, v - 1, w)) +
           knot(i + 1, j, k + 1, vec(u - 1, v, w - 1)) + knot(i, j + 1, k + 1, vec(u, v - 1, w - 1)) +
           knot(i + 1, j + 1, k + 1, vec(u - 1, v - 1, w - 1));
}

// ---------------------------------------------------------------------------------------------
// Trace.hs / Shader.hs
// ---------------------------------------------------------------------------------------------
struct RayCounters {  // per-thread, reduced by the caller
    unsigned int shadow, secondary, perlin_range;
    Cnt cnt;
};
// ctxb = [(Color, Vec)] (Shader.hs:65) as the set of lights that reach the point: bit i = light `first + i` passed the
// facing, radius and shadow tests of mpreshade.  (color * falloff, direction) are recomputed from the light and the
// point with mpreshade's own arithmetic when a Surface is shaded, so a light set may hold up to 64 lights.
#define GDEV_MAX_LIGHTS 64
struct LightSel {
    int done, first, count;
    unsigned long long mask;
};
struct MatVal {  // a Material value; Blend may carry a texture-computed weight
    int kind, a, b, c, d;
    Flt p[8];
};
GD_FN void mat_load(const DScene& S, int id, MatVal& m) {
    const GlomeMaterial* g = S.materials + id;
    m.kind = g->kind; m.a = g->a; m.b = g->b; m.c = g->c; m.d = g->d;
#pragma unroll
    for (int i = 0; i < 8; i++) m.p[i] = g->p[i];
}

// one light of mpreshade (Shader.hs:70-78): false when the light faces away or is out of range; otherwise the shadow
// ray (origin lifted by delta along the normal, length llen - 2 delta) and whether the light casts shadows at all
GD_FN bool light_probe(const GlomeLight* Lp, const Vec& pos, const Vec& norm, Ray& r, Flt& d, bool& need_shadow) {
    Vec lvec = vsub(vec(Lp->pos[0], Lp->pos[1], Lp->pos[2]), pos);
    if (vdot(lvec, norm) < 0) return false;
    Flt llen = vlen(lvec);
    Vec ldir = vscale(lvec, 1 / llen);
    if (llen > Lp->rad) return false;
    need_shadow = Lp->do_shadow != 0;
    r = mkray(vscaleadd(pos, norm, GLM_DELTA), ldir);
    d = llen - (2 * GLM_DELTA);
    return true;
}

// Surface color alpha amb kd ks shine (Shader.hs:90-105); p = {r,g,b,alpha,ambient,kd,ks,shine}
GD_FN void shade_surface(const DScene& S, const LightSel& L, const Vec& pos, const Flt* p, const Vec& n, const Vec& eyedir,
                         ColorA& outc) {
    Flt alpha = p[3], amb = p[4], kd = p[5], ks = p[6], shine = p[7];
    Flt ar = p[0] * amb, ag = p[1] * amb, ab = p[2] * amb;
    Flt dr = 0, dg = 0, db = 0;
    for (int i = 0; i < L.count; i++) {
        if (!((L.mask >> i) & 1ull)) continue;
        const GlomeLight* Lp = S.lights + L.first + i;
        Vec lvec = vsub(vec(Lp->pos[0], Lp->pos[1], Lp->pos[2]), pos);
        Flt llen = vlen(lvec);
        Vec ldir = vscale(lvec, 1 / llen);
        Flt fall = 1 / (llen * llen);  // Shader.hs:23
        Vec halfangle = bisect(ldir, eyedir);
        Flt ldotn = fmax_(0, vdot(ldir, n));
        Flt blinn;
        if (ks <= GLM_DELTA) blinn = 0;
        else {
            Flt b = fmax_(0, pow(vdot(halfangle, n), shine) * ldotn);
            blinn = (b != b) ? 0 : b;  // isNaN (Shader.hs:99)
        }
        Flt diffuse = vdot(ldir, n);
        Flt w = (blinn * ks) + (diffuse * kd);
        dr = dr + (Lp->color[0] * fall) * w;
        dg = dg + (Lp->color[1] * fall) * w;
        db = db + (Lp->color[2] * fall) * w;
    }
    outc = mkca(ar + dr, ag + dg, ab + db, alpha);
}

// tex ray ri  (Solid.hs:97) -> Material value
GD_FN void eval_texture(const DScene& S, int tex, const Vec& pos, MatVal& m, unsigned int& perlin_range) {
    const GlomeTexture* T = S.textures + tex;
    int kind = T->kind;
    if (kind == GLOME_TEX_UNIFORM) { mat_load(S, T->a, m); return; }
    Flt scale;
    if (kind == GLOME_TEX_STRIPE_BLEND) scale = triangle_wave(vdot(pos, vec(T->p[0], T->p[1], T->p[2])));  // TestScene.hs:225
    else {  // perlin (Texture.hs:109-116); out-of-range results are counted, not trapped
        scale = (noise(vscale(pos, (Flt)T->p[0])) + 1) * FL(0.5);
        if (scale > 1 || scale < 0) perlin_range++;
    }
    m.kind = GLOME_MAT_BLEND; m.a = T->a; m.b = T->b; m.c = 0; m.d = 0;
    m.p[0] = scale;
}

// mpreshade (Shader.hs:65-80) for a flat-class scene
GD_FN void mpreshade_flat(const DScene& S, int lightset, int scene, const Hit& ri, LightSel& L, RayCounters& rc) {
    L.done = 1;
    L.first = S.lightsets[2 * lightset];
    L.count = S.lightsets[2 * lightset + 1];
    L.mask = 0;
    for (int li = 0; li < L.count; li++) {
        Ray sr; Flt d; bool ns = false;
        if (!light_probe(S.lights + L.first + li, ri.pos, ri.norm, sr, d, ns)) continue;
        if (ns) {
            rc.shadow++;
            if (shadow_scene_flat(S, scene, sr, d, &rc.cnt)) continue;
        }
        L.mask |= 1ull << li;
    }
}

// Surface / Blend-of-Surface materials: all a flat-class scene holds (host_builder.cpp: flat_class)
GD_FN void mshade_flat(const DScene& S, const LightSel& L, const MatVal& m, const Hit& ri, const Vec& eyedir, ColorA& outc) {
    if (m.kind == GLOME_MAT_SURFACE) shade_surface(S, L, ri.pos, m.p, ri.norm, eyedir, outc);
    else if (m.kind == GLOME_MAT_BLEND) {  // Shader.hs:181-184
        ColorA ca, cb;
        MatVal ma, mb;  // (the payloads in the working precision)
        mat_load(S, m.a, ma);
        mat_load(S, m.b, mb);
        shade_surface(S, L, ri.pos, ma.p, ri.norm, eyedir, ca);
        shade_surface(S, L, ri.pos, mb.p, ri.norm, eyedir, cb);
        outc = caweight(ca, cb, m.p[0]);
    } else outc = mkca(0, 0, 0, 0);
}

// trace (Trace.hs:59-82) for a flat-class scene.  Its materials gather no tags (Shader.hs:92), so the TraceResult's tag
// list `ts ++ tags` is the hit's own tag stack.
GD_FN void trace_flat(const DScene& S, int lightset, int sld, const Ray& ray, Flt depth, int recurs, ColorA& outc, Hit& ri,
                      RayCounters& rc) {
    outc = mkca(0, 0, 0, 0);
    if (recurs == 0) { hit_clear(ri); return; }
    rayint_scene_flat(S, sld, ray, depth, ri, &rc.cnt);
    if (!ri.hit) return;  // mmissshade (Shader.hs:186)
    LightSel ctxb;
    ctxb.done = 0; ctxb.first = 0; ctxb.count = 0; ctxb.mask = 0;
    ColorA colora = mkca(0, 0, 0, 0);
    const Vec eyedir = vinvert(ray.d);
    for (int i = 0; i < ri.tex.n; i++) {
        if (colora.a + GLM_DELTA >= 1) continue;  // opaque (Trace.hs:50)
        MatVal m;
        eval_texture(S, ri.tex.v[i], ri.pos, m, rc.perlin_range);
        if (!ctxb.done) mpreshade_flat(S, lightset, sld, ri, ctxb, rc);  // ctxb is forced by the first Surface shaded
        ColorA colorb;
        mshade_flat(S, ctxb, m, ri, eyedir, colorb);
        colora = cafold(colora, colorb);
    }
    outc = colora;
}

// ---------------------------------------------------------------------------------------------
// Glome.hs: camera rays
// ---------------------------------------------------------------------------------------------
struct DCamera { Vec pos, fwd, up, right; };
GD_FN void getCoordsf(int width, int height, Flt xf, Flt yf, Flt& xc, Flt& yc) {  // Glome.hs:119-140
    Flt widthf = (Flt)width, heightf = (Flt)height;
    xc = (((xf / widthf) * 2) - 1) * (widthf / heightf);
    yc = -(((yf / heightf) * 2) - 1);
}
GD_FN Ray camera_ray(const DCamera& cam, Flt x, Flt y) {  // get_rayint (Glome.hs:27-33)
    Vec dir = vnorm(vadd3(cam.fwd, vscale(cam.right, -x), vscale(cam.up, y)));
    return mkray(cam.pos, dir);
}

}  // namespace gdev
