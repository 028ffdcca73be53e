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

namespace gdev {

using namespace glm;

struct DScene {
    const GlomeNode* __restrict__ nodes;
    const GlomeBihNode* __restrict__ bih;
    const GlomeBvhNode* __restrict__ bvh;
    const int32_t* __restrict__ ipool;
    const double* __restrict__ dpool;
    const GlomeTexture* __restrict__ textures;
    const GlomeMaterial* __restrict__ materials;
    const GlomeLight* __restrict__ lights;
    const int32_t* __restrict__ lightsets;
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
__device__ __forceinline__ void stk_clear(Stk& s) { s.n = 0; }
// x : l  (head first); returns true on overflow
__device__ __forceinline__ bool stk_cons(Stk& out, int x, const Stk& l) {
    bool ovf = l.n >= GLOME_MAX_STACK;
    int n = ovf ? GLOME_MAX_STACK : l.n + 1;
#pragma unroll
    for (int i = GLOME_MAX_STACK - 1; i >= 1; i--) out.v[i] = l.v[i - 1];
    out.v[0] = x;
    out.n = n;
    return ovf;
}
// a ++ b
__device__ __forceinline__ bool stk_append(Stk& out, const Stk& a, const Stk& b) {
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
__device__ __forceinline__ void hit_clear(Hit& h) {
    h.t = GLM_INFINITY; h.hit = 0; h.prim = -1; h.sub = -1; h.flags = 0;
    h.tex.n = 0; h.tag.n = 0;
}
__device__ __forceinline__ Flt ridepth(const Hit& h) { return h.hit ? h.t : GLM_INFINITY; }  // Solid.hs:33
// nearest acc cand (Solid.hs:37-44): cand replaces acc unless acc is strictly nearer
__device__ __forceinline__ bool cand_wins(const Hit& acc, Flt t) { return !acc.hit || !(acc.t < t); }
__device__ __forceinline__ void fold_nearest(Hit& acc, const Hit& c) {
    int fl = acc.flags | c.flags;
    if (c.hit && cand_wins(acc, c.t)) acc = c;
    acc.flags = fl;
}

__device__ __forceinline__ Vec ldv(const double* __restrict__ p) { return vec(p[0], p[1], p[2]); }
__device__ __forceinline__ Bbox ldbb(const double* __restrict__ p) {
    // bbox records are 16-byte aligned: three 16-byte loads
    const double2* q = reinterpret_cast<const double2*>(p);
    double2 a = __ldg(q), b = __ldg(q + 1), c = __ldg(q + 2);
    return mkbb(vec(a.x, a.y, b.x), vec(b.y, c.x, c.y));
}

// ---------------------------------------------------------------------------------------------
// primitives.  FULL = also produce position and normal.
// ---------------------------------------------------------------------------------------------
template <bool FULL>
__device__ __forceinline__ bool prim_sphere(const double* __restrict__ p, const Ray& ray, Flt dist, Flt& t, Vec& pos,
                                            Vec& n) {
    // Sphere.hs:20-41.  {cx,cy,cz,r} is 32-byte aligned: two 16-byte loads
    const double2* q = reinterpret_cast<const double2*>(p);
    double2 c01 = __ldg(q), c23 = __ldg(q + 1);
    Vec center = vec(c01.x, c01.y, c23.x);
    Flt r = c23.y;
    Vec eo = vsub(center, ray.o);
    Flt v = vdot(eo, ray.d);
    Flt vsqr = v * v;
    Flt csqr = vdot(eo, eo);
    Flt rsqr = r * r;
    Flt disc = rsqr - (csqr - vsqr);
    if (disc < 0.0) return false;
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
__device__ __forceinline__ bool shadow_sphere(const double* __restrict__ p, const Ray& ray, Flt dist) {
    // Sphere.hs:51-71
    const double2* q = reinterpret_cast<const double2*>(p);
    double2 c01 = __ldg(q), c23 = __ldg(q + 1);
    Vec center = vec(c01.x, c01.y, c23.x);
    Flt r = c23.y;
    Vec eo = vsub(center, ray.o);
    Flt v = vdot(eo, ray.d);
    if ((dist >= (v - r)) && (v > 0.0)) {
        Flt vsqr = v * v;
        Flt csqr = vdot(eo, eo);
        Flt rsqr = r * r;
        Flt disc = rsqr - (csqr - vsqr);
        if (disc < 0.0) return false;
        Flt d = sqrt(disc);
        Flt hitdist = ((v - d) > 0) ? (v - d) : (v + d);
        if ((hitdist < 0) || (hitdist > dist)) return false;
        return true;
    }
    return false;
}

// Triangle.hs:45-73 / 109-141.  mode: 0 = flat normal, 1 = vertex normals (n1..n3 valid)
template <bool FULL>
__device__ __forceinline__ bool prim_triangle(const Vec& p1, const Vec& p2, const Vec& p3, bool smooth, const Vec& n1,
                                              const Vec& n2, const Vec& n3, const Ray& ray, Flt dist, Flt& t, Vec& pos,
                                              Vec& n) {
    Vec e1 = vsub(p2, p1);
    Vec e2 = vsub(p3, p1);
    Vec s1 = vcross(ray.d, e2);
    Flt divisor = vdot(s1, e1);
    if (divisor == 0) return false;
    Flt invdivisor = 1.0 / divisor;
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
__device__ __forceinline__ bool shadow_triangle(const Vec& p1, const Vec& p2, const Vec& p3, const Ray& ray, Flt dist) {
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
    Flt tt = vdot(e2, s2) * invdivisor;
    return (tt >= 0) && (tt <= dist);
}

template <bool FULL>
__device__ __forceinline__ bool prim_box(const double* __restrict__ p, const Ray& r, Flt d, Flt& t, Vec& pos, Vec& n) {
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
__device__ __forceinline__ bool shadow_box(const double* __restrict__ p, const Ray& r, Flt d) {  // Box.hs:56-62
    Bbox b = ldbb(p);
    Flt near_, far_;
    bbclip_ub(r, b, near_, far_);
    if ((near_ > far_) || far_ <= 0 || far_ > d) return false;
    return true;
}

template <bool FULL>
__device__ __forceinline__ bool prim_plane(const double* __restrict__ p, const Ray& ray, Flt d, Flt& t, Vec& pos, Vec& n) {
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
__device__ __forceinline__ bool prim_disc_v(const Vec& point, const Vec& norm, Flt radius_sqr, const Ray& r, Flt d, Flt& t,
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
__device__ __forceinline__ bool prim_cylinder(const double* __restrict__ p, const Ray& ray, Flt d, Flt& t, Vec& pos, Vec& n) {
    // Cone.hs:104-139
    Flt r = p[0], h1 = p[1], h2 = p[2];
    Flt ox = ray.o.x, oy = ray.o.y, oz = ray.o.z, dx = ray.d.x, dy = ray.d.y, dz = ray.d.z;
    Flt a = dx * dx + dy * dy;
    Flt b = 2 * (dx * ox + dy * oy);
    Flt c = ox * ox + oy * oy - r * r;
    Flt disc = b * b - 4 * a * c;
    if (disc < 0) return false;
    Flt discsqrt = sqrt(disc);
    Flt q = (b < 0) ? (b - discsqrt) * (-0.5) : (b + discsqrt) * (-0.5);
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
__device__ __forceinline__ bool prim_cone(const double* __restrict__ p, const Ray& ray, Flt d, Flt& t, Vec& pos, Vec& n) {
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
    Flt q = (b < 0) ? (b - discsqrt) * (-0.5) : (b + discsqrt) * (-0.5);
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

__device__ __forceinline__ bool is_prim(int type) { return type >= GLOME_SPHERE && type <= GLOME_CONE; }

// rayint of a primitive node
template <bool FULL>
__device__ __forceinline__ bool prim_rayint(const DScene& S, const GlomeNode& nd, const Ray& r, Flt d, Flt& t, Vec& pos,
                                            Vec& n) {
    const double* p = S.dpool + nd.a;
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
__device__ __forceinline__ bool prim_shadow(const DScene& S, const GlomeNode& nd, const Ray& r, Flt d) {
    const double* p = S.dpool + nd.a;
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
__device__ __forceinline__ bool prim_inside(const DScene& S, const GlomeNode& nd, const Vec& pt) {
    const double* p = S.dpool + nd.a;
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
// the scene-graph interpreter
// ---------------------------------------------------------------------------------------------
template <int L>
__device__ void rayint_node(const DScene& S, int ni, const Ray& r, Flt d, const Stk& tex, const Stk& tag, int csg, Hit& acc,
                            Cnt* cnt = nullptr);
template <int L>
__device__ bool shadow_node(const DScene& S, int ni, const Ray& r, Flt d, int csg, Cnt* cnt = nullptr);
__device__ bool inside_node(const DScene& S, int ni, const Vec& pt);
__device__ void metainfo_node(const DScene& S, int ni, const Vec& v, Stk& texs, Stk& tags, int& flags);

// record a winning primitive hit
__device__ __forceinline__ void take_hit(Hit& acc, Flt t, const Vec& pos, const Vec& n, const Ray& r, const Stk& tex,
                                         const Stk& tag, int prim, int sub) {
    acc.hit = 1; acc.t = t; acc.pos = pos; acc.norm = n; acc.ray = r; acc.tex = tex; acc.tag = tag;
    acc.prim = prim; acc.sub = sub;
}

struct TravEnt { int ref; Flt near_, far_; };

// rayint_bih (Bih.hs:332-368), iterative, reference order, best-hit culling
template <int L>
__device__ __forceinline__ void rayint_bih(const DScene& S, const GlomeNode& nd, const Ray& r, Flt d, const Stk& tex,
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
            const double2* np = reinterpret_cast<const double2*>(S.bih + ref);
            double2 sp2 = __ldg(np);
            int4 ii = __ldg(reinterpret_cast<const int4*>(np + 1));
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
__device__ __forceinline__ bool shadow_bih(const DScene& S, const GlomeNode& nd, const Ray& r, Flt d, int csg,
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
            const double2* np = reinterpret_cast<const double2*>(S.bih + ref);
            double2 sp2 = __ldg(np);
            int4 ii = __ldg(reinterpret_cast<const int4*>(np + 1));
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
__device__ __forceinline__ void rayint_mesh(const DScene& S, int ni, const GlomeNode& nd, const Ray& ray, Flt depth,
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
            int ntri = __ldg(S.ipool + k);
            for (int j = 0; j < ntri; j++) {
                int ti = __ldg(S.ipool + k + 1 + j);
                if (cnt) cnt->tri++;
                const int4* tp = reinterpret_cast<const int4*>(S.ipool + tris_off + 8 * ti);
                int4 t0 = __ldg(tp), t1 = __ldg(tp + 1);  // {a,b,c,na} {nb,nc,tex,tag}
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
                    if (t1.z != -1) { if (stk_cons(tx, __ldg(S.ipool + texs_off + t1.z), texs)) acc.flags |= GLOME_HITFLAG_STACK_OVERFLOW; }
                    if (t1.w != -1) { if (stk_cons(tg, __ldg(S.ipool + tags_off + t1.w), tags)) acc.flags |= GLOME_HITFLAG_STACK_OVERFLOW; }
                    take_hit(acc, t, pos, n, ray, tx, tg, ni, ti);
                }
            }
            pop = true;
        } else {
            // 128-byte node: two boxes + two child refs
            if (cnt) cnt->bvh++;
            const double2* np = reinterpret_cast<const double2*>(S.bvh + ref);
            double2 l0 = __ldg(np), l1 = __ldg(np + 1), l2 = __ldg(np + 2);
            double2 r0 = __ldg(np + 3), r1 = __ldg(np + 4), r2 = __ldg(np + 5);
            int2 kids = __ldg(reinterpret_cast<const int2*>(np + 6));
            Bbox lbb = mkbb(vec(l0.x, l0.y, l1.x), vec(l1.y, l2.x, l2.y));
            Bbox rbb = mkbb(vec(r0.x, r0.y, r1.x), vec(r1.y, r2.x, r2.y));
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
__device__ __noinline__ void rayint_bih_gen(const DScene& S, const GlomeNode& nd, const Ray& r, Flt d, const Stk& tex,
                                            const Stk& tag, int csg, Hit& acc) {
    rayint_bih<-1>(S, nd, r, d, tex, tag, csg, acc);
}
__device__ __noinline__ bool shadow_bih_gen(const DScene& S, const GlomeNode& nd, const Ray& r, Flt d, int csg) {
    return shadow_bih<-1>(S, nd, r, d, csg);
}
__device__ __noinline__ void rayint_mesh_gen(const DScene& S, int ni, const GlomeNode& nd, const Ray& ray, Flt depth,
                                             const Stk& texs, const Stk& tags, Hit& acc) {
    rayint_mesh(S, ni, nd, ray, depth, texs, tags, true, acc);
}

// ---- general-only pieces (true recursion) -----------------------------------------------------
__device__ void rayint_advance(const DScene& S, int ni, const Ray& r, Flt d, const Stk& t, const Stk& tags, Flt adv, int csg,
                               Hit& out) {
    // Solid.hs:85-91
    hit_clear(out);
    if (csg >= GDEV_CSG_CAP) { out.flags |= GLOME_HITFLAG_CSG_OVERFLOW; return; }
    Flt a = adv + GLM_DELTA;
    rayint_node<-1>(S, ni, ray_move(r, a), d - a, t, tags, csg + 1, out);
    if (out.hit) out.t = out.t + a;
}

__device__ bool inside_isect(const DScene& S, int first, int count, const Vec& pt) {  // Csg.hs:99-101
    for (int i = 0; i < count; i++) {
        GlomeNode c = S.nodes[first + i];
        if (is_prim(c.type)) { if (!prim_inside(S, c, pt)) return false; continue; }  // planes of a polyhedron
        if (!inside_node(S, first + i, pt)) return false;
    }
    return true;
}

__device__ void rayint_isect(const DScene& S, int first, int count, const Ray& r, Flt d, const Stk& t, const Stk& tags,
                             int csg, Hit& out) {
    // Csg.hs:68-90 over the suffix [first, first+count)
    hit_clear(out);
    if (count == 0 || d < 0) return;
    if (count == 1) { rayint_node<-1>(S, first, r, d, t, tags, csg, out); return; }
    bool in = inside_node(S, first, r.o);
    Hit rs;
    hit_clear(rs);
    rayint_node<-1>(S, first, r, d, t, tags, csg, rs);
    if (in) {
        if (!rs.hit) {
            rayint_isect(S, first + 1, count - 1, r, d, t, tags, csg, out);
            out.flags |= rs.flags;
            return;
        }
        rayint_isect(S, first + 1, count - 1, r, rs.t, t, tags, csg, out);
        if (out.hit) return;
        // rayint_advance (SolidItem (Intersection slds)) r d t tags sd
        int fl = out.flags | rs.flags;
        hit_clear(out);
        out.flags = fl;
        if (csg >= GDEV_CSG_CAP) { out.flags |= GLOME_HITFLAG_CSG_OVERFLOW; return; }
        Flt a = rs.t + GLM_DELTA;
        rayint_isect(S, first, count, ray_move(r, a), d - a, t, tags, csg + 1, out);
        if (out.hit) out.t = out.t + a;
        out.flags |= fl;
        return;
    }
    if (!rs.hit) { out.flags |= rs.flags; return; }
    if (inside_isect(S, first + 1, count - 1, rs.pos)) {
        out = rs;
        out.ray = r;  // RayHit sd sp sn r vzero st stags  (Csg.hs:88)
        return;
    }
    int fl = rs.flags;
    if (csg >= GDEV_CSG_CAP) { out.flags |= fl | GLOME_HITFLAG_CSG_OVERFLOW; return; }
    Flt a = rs.t + GLM_DELTA;
    rayint_isect(S, first, count, ray_move(r, a), d - a, t, tags, csg + 1, out);
    if (out.hit) out.t = out.t + a;
    out.flags |= fl;
}

__device__ void rayint_difference(const DScene& S, int ni, const GlomeNode& nd, const Ray& r, Flt d, const Stk& t,
                                  const Stk& tags, int csg, Hit& out) {
    // Csg.hs:33-54
    hit_clear(out);
    int sa = nd.a, sb = nd.b;
    if (inside_node(S, sb, r.o)) {
        Hit rib;
        hit_clear(rib);
        rayint_node<-1>(S, sb, r, d, t, tags, csg, rib);
        if (!rib.hit) { out.flags |= rib.flags; return; }
        if (inside_node(S, sa, rib.pos) && !inside_node(S, sb, vscaleadd(rib.pos, r.d, GLM_DELTA))) {
            out = rib;
            out.norm = vinvert(rib.norm);
            if (nd.c != 0) {  // useatex: textures/tags come from get_metainfo sa bp ONLY
                int fl = 0;
                metainfo_node(S, sa, rib.pos, out.tex, out.tag, fl);
                out.flags |= fl;
            }
            return;
        }
        rayint_advance(S, ni, r, d, t, tags, rib.t, csg, out);
        out.flags |= rib.flags;
        return;
    }
    Hit ria;
    hit_clear(ria);
    rayint_node<-1>(S, sa, r, d, t, tags, csg, ria);
    if (!ria.hit) { out.flags |= ria.flags; return; }
    Hit rib;
    hit_clear(rib);
    rayint_node<-1>(S, sb, r, d, t, tags, csg, rib);
    if (rib.hit) {
        if (ria.t < rib.t) { out = ria; out.flags |= rib.flags; return; }
        rayint_advance(S, ni, r, d, t, tags, rib.t, csg, out);
        out.flags |= ria.flags | rib.flags;
        return;
    }
    out = ria;
    out.flags |= rib.flags;
}

// class Solid: rayint (Solid.hs:146).  Folds the node's result into acc with `nearest`.
template <int L>
__device__ void rayint_node(const DScene& S, int ni, const Ray& r, Flt d, const Stk& tex_in, const Stk& tag_in, int csg,
                            Hit& acc, Cnt* cnt) {
    constexpr bool GEN = (L < 0);
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
        if (prim_rayint<GEN>(S, nd, r, d, t, pos, n) && cand_wins(acc, t)) take_hit(acc, t, pos, n, r, tex, tag, ni, -1);
        return;
    }
    switch (nd.type) {
        case GLOME_VOID:
        case GLOME_ONLYSHADOW: return;  // Solid.hs:354, Tex.hs:89
        case GLOME_BIH:
            if constexpr (GEN) rayint_bih_gen(S, nd, r, d, tex, tag, csg, acc);
            else if constexpr (L == 0 || L == 1) rayint_bih<L>(S, nd, r, d, tex, tag, csg, acc, cnt);
            return;
        case GLOME_MESH:
            if constexpr (GEN) rayint_mesh_gen(S, ni, nd, r, d, tex, tag, acc);
            else if constexpr (L == 0 || L == 1) rayint_mesh(S, ni, nd, r, d, tex, tag, false, acc, cnt);
            return;
        case GLOME_GROUP:  // Solid.hs:327
            if constexpr (L == 0 || GEN) {
                for (int i = 0; i < nd.b; i++) {
                    if constexpr (GEN) {
                        // {Tex,Tag}* prim children inline (a 64-box chessboard is 64 calls otherwise)
                        int cj = nd.a + i;
                        GlomeNode c = S.nodes[cj];
                        int ntx = 0, ntg = 0;
                        while (c.type == GLOME_TEX || c.type == GLOME_TAG || c.type == GLOME_NOSHADOW) {
                            ntx += (c.type == GLOME_TEX); ntg += (c.type == GLOME_TAG);
                            cj = c.a; c = S.nodes[cj];
                        }
                        if (is_prim(c.type)) {
                            Flt t; Vec pos, n;
                            if (prim_rayint<true>(S, c, r, d, t, pos, n) && cand_wins(acc, t)) {
                                take_hit(acc, t, pos, n, r, tex, tag, cj, -1);
                                if (ntx | ntg) {  // rebuild the stacks of the winner only
                                    int wj = nd.a + i;
                                    GlomeNode w = S.nodes[wj];
                                    while (w.type == GLOME_TEX || w.type == GLOME_TAG || w.type == GLOME_NOSHADOW) {
                                        if (w.type == GLOME_TEX) { if (stk_cons(acc.tex, w.b, acc.tex)) acc.flags |= GLOME_HITFLAG_STACK_OVERFLOW; }
                                        else if (w.type == GLOME_TAG) { if (stk_cons(acc.tag, w.b, acc.tag)) acc.flags |= GLOME_HITFLAG_STACK_OVERFLOW; }
                                        wj = w.a; w = S.nodes[wj];
                                    }
                                }
                            }
                            continue;
                        }
                        if (c.type == GLOME_VOID) continue;
                    }
                    rayint_node<(GEN ? -1 : 1)>(S, nd.a + i, r, d, tex, tag, csg, acc, cnt);
                }
            }
            return;
    }
    if constexpr (GEN) {
        switch (nd.type) {
            case GLOME_INSTANCE: {  // Solid.hs:388-403
                const Flt* xfm = S.dpool + nd.b;
                Vec newdir = invxfm_vec(xfm, r.d);
                Vec neworig = invxfm_point(xfm, r.o);
                Flt lenscale = vlen(newdir);
                Flt invlenscale = 1 / lenscale;
                {   // {Tex,Tag}* prim child: no recursion (oak leaves, cone / cylinder constructors)
                    int cj = nd.a;
                    GlomeNode c = S.nodes[cj];
                    int nw = 0;
                    while (c.type == GLOME_TEX || c.type == GLOME_TAG || c.type == GLOME_NOSHADOW) { nw++; cj = c.a; c = S.nodes[cj]; }
                    if (is_prim(c.type)) {
                        Ray ir = mkray(neworig, vscale(newdir, invlenscale));
                        Flt t; Vec pos, n;
                        if (prim_rayint<true>(S, c, ir, d * lenscale, t, pos, n)) {
                            Flt tw = t * invlenscale;
                            if (cand_wins(acc, tw)) {
                                take_hit(acc, tw, xfm_point(xfm, pos), vnorm(invxfm_norm(xfm, n)), ir, tex, tag, cj, -1);
                                if (nw) {
                                    int wj = nd.a;
                                    GlomeNode w = S.nodes[wj];
                                    while (w.type == GLOME_TEX || w.type == GLOME_TAG || w.type == GLOME_NOSHADOW) {
                                        if (w.type == GLOME_TEX) { if (stk_cons(acc.tex, w.b, acc.tex)) acc.flags |= GLOME_HITFLAG_STACK_OVERFLOW; }
                                        else if (w.type == GLOME_TAG) { if (stk_cons(acc.tag, w.b, acc.tag)) acc.flags |= GLOME_HITFLAG_STACK_OVERFLOW; }
                                        wj = w.a; w = S.nodes[wj];
                                    }
                                }
                            }
                        }
                        return;
                    }
                }
                Hit h;
                hit_clear(h);
                rayint_node<-1>(S, nd.a, mkray(neworig, vscale(newdir, invlenscale)), d * lenscale, tex, tag, csg, h);
                acc.flags |= h.flags;
                if (h.hit) {
                    Flt t = h.t * invlenscale;
                    if (cand_wins(acc, t)) {
                        int fl = acc.flags;
                        acc = h;
                        acc.flags = fl;
                        acc.t = t;
                        acc.pos = xfm_point(xfm, h.pos);
                        acc.norm = vnorm(invxfm_norm(xfm, h.norm));
                    }
                }
                return;
            }
            case GLOME_DIFFERENCE: {
                Hit h;
                rayint_difference(S, ni, nd, r, d, tex, tag, csg, h);
                fold_nearest(acc, h);
                return;
            }
            case GLOME_INTERSECTION: {
                Hit h;
                rayint_isect(S, nd.a, nd.b, r, d, tex, tag, csg, h);
                fold_nearest(acc, h);
                return;
            }
            case GLOME_BOUND:  // Bound.hs:30-35
                if (inside_node(S, nd.a, r.o) || shadow_node<-1>(S, nd.a, r, d, csg)) {
                    Hit h;
                    hit_clear(h);
                    rayint_node<-1>(S, nd.b, r, d, tex, tag, csg, h);
                    fold_nearest(acc, h);
                }
                return;
            case GLOME_INNERBOUND: {  // Bound.hs:98-99
                Hit ha;
                hit_clear(ha);
                Stk e;
                stk_clear(e);
                rayint_node<-1>(S, nd.a, r, d, e, e, csg, ha);
                Hit h;
                hit_clear(h);
                rayint_node<-1>(S, nd.b, r, ridepth(ha), tex, tag, csg, h);
                fold_nearest(acc, h);
                return;
            }
        }
    }
}

// class Solid: shadow (Solid.hs:162)
template <int L>
__device__ bool shadow_node(const DScene& S, int ni, const Ray& r, Flt d, int csg, Cnt* cnt) {
    constexpr bool GEN = (L < 0);
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
            if constexpr (GEN) return shadow_bih_gen(S, nd, r, d, csg);
            else if constexpr (L == 0 || L == 1) return shadow_bih<L>(S, nd, r, d, csg, cnt);
            return false;
        case GLOME_GROUP:  // Solid.hs:330
            if constexpr (L == 0 || GEN) {
                for (int i = 0; i < nd.b; i++) {
                    if constexpr (GEN) {
                        GlomeNode c = S.nodes[nd.a + i];
                        while (c.type == GLOME_TEX || c.type == GLOME_TAG || c.type == GLOME_ONLYSHADOW) c = S.nodes[c.a];
                        if (is_prim(c.type)) { if (prim_shadow(S, c, r, d)) return true; continue; }
                        if (c.type == GLOME_VOID || c.type == GLOME_NOSHADOW || c.type == GLOME_MESH) continue;
                    }
                    if (shadow_node<(GEN ? -1 : 1)>(S, nd.a + i, r, d, csg, cnt)) return true;
                }
            }
            return false;
    }
    if constexpr (GEN) {
        switch (nd.type) {
            case GLOME_INSTANCE: {  // Solid.hs:464-471
                const Flt* xfm = S.dpool + nd.b;
                Vec newdir = invxfm_vec(xfm, r.d);
                Vec neworig = invxfm_point(xfm, r.o);
                Flt lenscale = vlen(newdir);
                Flt invlenscale = 1 / lenscale;
                {
                    GlomeNode c = S.nodes[nd.a];
                    while (c.type == GLOME_TEX || c.type == GLOME_TAG || c.type == GLOME_ONLYSHADOW) c = S.nodes[c.a];
                    if (is_prim(c.type)) return prim_shadow(S, c, mkray(neworig, vscale(newdir, invlenscale)), d * lenscale);
                }
                return shadow_node<-1>(S, nd.a, mkray(neworig, vscale(newdir, invlenscale)), d * lenscale, csg);
            }
            case GLOME_DIFFERENCE:
            case GLOME_INTERSECTION: {  // no shadow method: default falls back on rayint (Solid.hs:218-221)
                Hit h;
                hit_clear(h);
                Stk e;
                stk_clear(e);
                rayint_node<-1>(S, ni, r, d, e, e, csg, h);
                return h.hit != 0;
            }
            case GLOME_BOUND:  // Bound.hs:44-49
                if (inside_node(S, nd.a, r.o) || shadow_node<-1>(S, nd.a, r, d, csg)) return shadow_node<-1>(S, nd.b, r, d, csg);
                return false;
            case GLOME_INNERBOUND:  // Bound.hs:101-103
                return shadow_node<-1>(S, nd.a, r, d, csg) || shadow_node<-1>(S, nd.b, r, d, csg);
        }
    }
    return false;
}

// inside_bih (Bih.hs:550-565): point descent, both sides possible
__device__ bool inside_bih_rec(const DScene& S, int ref, const Vec& pt) {
    if (ref < 0) {
        int first, cnt;
        glome_bih_leaf(ref, S.ipool, &first, &cnt);
        for (int i = 0; i < cnt; i++)
            if (inside_node(S, first + i, pt)) return true;
        return false;
    }
    GlomeBihNode n = S.bih[ref];
    Flt o = va(pt, n.axis);
    if (o < n.lsplit && inside_bih_rec(S, n.left, pt)) return true;
    if (o > n.rsplit && inside_bih_rec(S, n.right, pt)) return true;
    return false;
}

// class Solid: inside (Solid.hs:166)
__device__ bool inside_node(const DScene& S, int ni, const Vec& pt) {
    GlomeNode nd = S.nodes[ni];
    while (nd.type == GLOME_TEX || nd.type == GLOME_TAG || nd.type == GLOME_NOSHADOW || nd.type == GLOME_ONLYSHADOW) {
        ni = nd.a;
        nd = S.nodes[ni];
    }
    if (is_prim(nd.type)) return prim_inside(S, nd, pt);
    switch (nd.type) {
        case GLOME_GROUP:  // Solid.hs:331
            for (int i = 0; i < nd.b; i++) {
                GlomeNode c = S.nodes[nd.a + i];
                while (c.type == GLOME_TEX || c.type == GLOME_TAG || c.type == GLOME_NOSHADOW || c.type == GLOME_ONLYSHADOW) c = S.nodes[c.a];
                if (is_prim(c.type)) { if (prim_inside(S, c, pt)) return true; continue; }
                if (inside_node(S, nd.a + i, pt)) return true;
            }
            return false;
        case GLOME_INSTANCE: {  // Solid.hs:473
            Vec q = invxfm_point(S.dpool + nd.b, pt);
            GlomeNode c = S.nodes[nd.a];
            while (c.type == GLOME_TEX || c.type == GLOME_TAG || c.type == GLOME_NOSHADOW || c.type == GLOME_ONLYSHADOW) c = S.nodes[c.a];
            if (is_prim(c.type)) return prim_inside(S, c, q);
            return inside_node(S, nd.a, q);
        }
        case GLOME_BIH: {
            const double* b = S.dpool + nd.b;
            return (pt.x > b[0]) && (pt.x < b[3]) && (pt.y > b[1]) && (pt.y < b[4]) && (pt.z > b[2]) && (pt.z < b[5]) &&
                   inside_bih_rec(S, nd.a, pt);
        }
        case GLOME_DIFFERENCE: return inside_node(S, nd.a, pt) && !inside_node(S, nd.b, pt);  // Csg.hs:92
        case GLOME_INTERSECTION: return inside_isect(S, nd.a, nd.b, pt);
        case GLOME_BOUND: return inside_node(S, nd.a, pt) && inside_node(S, nd.b, pt);       // Bound.hs:51
        case GLOME_INNERBOUND: return inside_node(S, nd.a, pt) || inside_node(S, nd.b, pt);  // Bound.hs:109
    }
    return false;  // Void, Mesh
}

__device__ void metainfo_list(const DScene& S, int first, int count, const Vec& v, Stk& texs, Stk& tags, int& flags) {
    // Solid.hs:337-339: later elements are prepended
    Stk at, ag;
    stk_clear(at);
    stk_clear(ag);
    for (int i = 0; i < count; i++) {
        if (inside_node(S, first + i, v)) {
            Stk xt, xg;
            metainfo_node(S, first + i, v, xt, xg, flags);
            if (stk_append(at, xt, at)) flags |= GLOME_HITFLAG_STACK_OVERFLOW;
            if (stk_append(ag, xg, ag)) flags |= GLOME_HITFLAG_STACK_OVERFLOW;
        }
    }
    texs = at;
    tags = ag;
}
__device__ void metainfo_bih_rec(const DScene& S, int ref, const Vec& pt, Stk& texs, Stk& tags, int& flags) {
    // Bih.hs:568-577
    if (ref < 0) {
        int first, cnt;
        glome_bih_leaf(ref, S.ipool, &first, &cnt);
        metainfo_list(S, first, cnt, pt, texs, tags, flags);
        return;
    }
    GlomeBihNode n = S.bih[ref];
    Flt o = va(pt, n.axis);
    Stk lt, lg, rt, rg;
    stk_clear(lt); stk_clear(lg); stk_clear(rt); stk_clear(rg);
    if (o < n.lsplit) metainfo_bih_rec(S, n.left, pt, lt, lg, flags);
    if (o > n.rsplit) metainfo_bih_rec(S, n.right, pt, rt, rg, flags);
    if (stk_append(texs, lt, rt)) flags |= GLOME_HITFLAG_STACK_OVERFLOW;
    if (stk_append(tags, lg, rg)) flags |= GLOME_HITFLAG_STACK_OVERFLOW;
}
// class Solid: get_metainfo (Solid.hs:200)
__device__ void metainfo_node(const DScene& S, int ni, const Vec& v, Stk& texs, Stk& tags, int& flags) {
    GlomeNode nd = S.nodes[ni];
    stk_clear(texs);
    stk_clear(tags);
    switch (nd.type) {
        case GLOME_GROUP: metainfo_list(S, nd.a, nd.b, v, texs, tags, flags); return;
        case GLOME_INSTANCE: metainfo_node(S, nd.a, invxfm_point(S.dpool + nd.b, v), texs, tags, flags); return;  // Solid.hs:517
        case GLOME_BIH: {  // Bih.hs:567-585
            const double* b = S.dpool + nd.b;
            if ((v.x > b[0]) && (v.x < b[3]) && (v.y > b[1]) && (v.y < b[4]) && (v.z > b[2]) && (v.z < b[5]))
                metainfo_bih_rec(S, nd.a, v, texs, tags, flags);
            return;
        }
        case GLOME_DIFFERENCE:  // Csg.hs:103-106
            if (inside_node(S, nd.a, v) && !inside_node(S, nd.b, v)) metainfo_node(S, nd.a, v, texs, tags, flags);
            return;
        case GLOME_INTERSECTION:  // Csg.hs:108-111
            if (inside_isect(S, nd.a, nd.b, v)) {
                for (int i = 0; i < nd.b; i++) {
                    Stk xt, xg;
                    metainfo_node(S, nd.a + i, v, xt, xg, flags);
                    if (stk_append(texs, texs, xt)) flags |= GLOME_HITFLAG_STACK_OVERFLOW;
                    if (stk_append(tags, tags, xg)) flags |= GLOME_HITFLAG_STACK_OVERFLOW;
                }
            }
            return;
        case GLOME_TEX: {  // Tex.hs:73-74
            Stk xt, xg;
            metainfo_node(S, nd.a, v, xt, xg, flags);
            if (stk_cons(texs, nd.b, xt)) flags |= GLOME_HITFLAG_STACK_OVERFLOW;
            tags = xg;
            return;
        }
        case GLOME_TAG: {  // Tex.hs:61-62
            Stk xt, xg;
            metainfo_node(S, nd.a, v, xt, xg, flags);
            texs = xt;
            if (stk_cons(tags, nd.b, xg)) flags |= GLOME_HITFLAG_STACK_OVERFLOW;
            return;
        }
        case GLOME_NOSHADOW:
        case GLOME_ONLYSHADOW: metainfo_node(S, nd.a, v, texs, tags, flags); return;
        case GLOME_BOUND:  // Bound.hs:54-58
            if (inside_node(S, nd.a, v)) metainfo_node(S, nd.b, v, texs, tags, flags);
            return;
        case GLOME_INNERBOUND: metainfo_node(S, nd.b, v, texs, tags, flags); return;  // Bound.hs:112
    }
}

// Flat scenes record t / prim / stacks while traversing and compute position + normal once, for
// the winner (same arithmetic as the eager path: prim_rayint<true> on the winning primitive).
__device__ __forceinline__ void finalize_flat(const DScene& S, const Ray& r, Hit& h) {
    if (!h.hit) return;
    GlomeNode nd = S.nodes[h.prim];
    Flt t; Vec pos, n;
    const Flt big = 1.0e300;
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
    } else {
        prim_rayint<true>(S, nd, r, big, t, pos, n);
    }
    h.pos = pos; h.norm = n; h.ray = r;
}

// rayint sld ray d [] []
template <bool GEN>
__device__ __forceinline__ void rayint_scene(const DScene& S, int sld, const Ray& r, Flt d, Hit& h, Cnt* cnt = nullptr) {
    hit_clear(h);
    Stk e;
    stk_clear(e);
    if (GEN) rayint_node<-1>(S, sld, r, d, e, e, 0, h, nullptr);
    else { rayint_node<0>(S, sld, r, d, e, e, 0, h, cnt); finalize_flat(S, r, h); }
}
template <bool GEN>
__device__ __forceinline__ bool shadow_scene(const DScene& S, int sld, const Ray& r, Flt d, Cnt* cnt = nullptr) {
    if (GEN) return shadow_node<-1>(S, sld, r, d, 0, nullptr);
    return shadow_node<0>(S, sld, r, d, 0, cnt);
}

// ---------------------------------------------------------------------------------------------
// Clr.hs, Texture.hs
// ---------------------------------------------------------------------------------------------
struct Color { Flt r, g, b; };
struct ColorA { Flt r, g, b, a; };
__device__ __forceinline__ ColorA mkca(Flt r, Flt g, Flt b, Flt a) { ColorA c; c.r = r; c.g = g; c.b = b; c.a = a; return c; }
__device__ __forceinline__ Flt aclamp(Flt x) { if (x > 1) return 1; if (x < 0) return 0; return x; }  // Clr.hs:75
__device__ __forceinline__ ColorA caweight(const ColorA& x, const ColorA& y, Flt w) {                 // Clr.hs:87
    return mkca((x.r * w) + (y.r * (1 - w)), (x.g * w) + (y.g * (1 - w)), (x.b * w) + (y.b * (1 - w)),
                (x.a * w) + (y.a * (1 - w)));
}
__device__ __forceinline__ ColorA cafold(const ColorA& x, const ColorA& y) {  // Clr.hs:106
    Flt trans = 1 - x.a;
    return mkca(x.r + (y.r * trans * y.a), x.g + (y.g * trans * y.a), x.b + (y.b * trans * y.a), x.a + (y.a * trans));
}
__device__ __forceinline__ Flt triangle_wave(Flt x) {  // Texture.hs:16
    Flt offset = x - floor(x);
    return (offset < 0.5) ? (offset * 2) : (2 - (offset * 2));
}
__device__ __forceinline__ Flt omega(Flt t_) {  // Texture.hs:49
    Flt t = fabs_(t_);
    Flt tsqr = t * t;
    Flt tcube = tsqr * t;
    return (-6) * tcube * tsqr + 15 * tcube * t - 10 * tcube + 1;
}
__constant__ int c_phi[12] = {3, 0, 2, 7, 4, 1, 5, 11, 8, 10, 9, 6};  // Texture.hs:57
__constant__ signed char c_grad[12][3] = {{-1, -1, 0}, {-1, 0, -1}, {-1, 0, 1}, {-1, 1, 0}, {0, -1, -1}, {0, -1, 1},
                                          {0, 1, -1},  {0, 1, 1},   {1, -1, 0}, {1, 0, -1}, {1, 0, 1},   {1, 1, 0}};  // Texture.hs:60
__device__ __forceinline__ long long iabs64(long long a) { return a < 0 ? -a : a; }
__device__ __forceinline__ Flt knot(long long i, long long j, long long k, const Vec& v) {  // Texture.hs:67-77
    int a = c_phi[iabs64(k) % 12];
    int b = c_phi[iabs64(j + a) % 12];
    int c = c_phi[iabs64(i + b) % 12];
    Vec g = vec((Flt)c_grad[c][0], (Flt)c_grad[c][1], (Flt)c_grad[c][2]);
    return omega(v.x) * omega(v.y) * omega(v.z) * vdot(g, v);
}
__device__ Flt noise(const Vec& p) {  // Texture.hs:92-107
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
#define GDEV_MAX_LIGHTS 8
struct LightCtx {  // ctxb = [(Color, Vec)] (Shader.hs:65), evaluated on first use like the lazy original
    int done, n;
    Color col[GDEV_MAX_LIGHTS];
    Vec dir[GDEV_MAX_LIGHTS];
};
struct MatVal {  // a Material value; Blend may carry a texture-computed weight
    int kind, a, b, c, d;
    Flt p[8];
};
__device__ __forceinline__ void mat_load(const DScene& S, int id, MatVal& m) {
    const GlomeMaterial* g = S.materials + id;
    m.kind = g->kind; m.a = g->a; m.b = g->b; m.c = g->c; m.d = g->d;
#pragma unroll
    for (int i = 0; i < 8; i++) m.p[i] = g->p[i];
}

// The [t] half of a TraceResult (Trace.hs:59-82; Shader.hs:116,154,171-183).  The render path discards it
// (Glome.hs:53-55), so the kernels that render instantiate trace / mpostshade with NoTags (an empty type: no register,
// no code); the pick query (getTags', Glome.hs:410-414) instantiates them with a TagList*.  Every function APPENDS
// its list to what the caller passed.  Capacity 16, the tail is dropped and flagged.
#define GDEV_TAGLIST_CAP 16
struct TagList { int n; int overflow; int v[GDEV_TAGLIST_CAP]; };
struct NoTags {};
__device__ __forceinline__ void tl_clear(TagList& t) { t.n = 0; t.overflow = 0; }
__device__ __forceinline__ void tl_push(TagList* t, int x) { if (t->n < GDEV_TAGLIST_CAP) t->v[t->n++] = x; else t->overflow = 1; }
__device__ __forceinline__ void tl_append(TagList* dst, const TagList& src) {  // dst ++ src
    for (int i = 0; i < src.n; i++) tl_push(dst, src.v[i]);
    dst->overflow |= src.overflow;
}
__device__ __forceinline__ void tl_append(NoTags, const TagList&) {}
template <typename TL> struct tl_on { static const bool value = true; };
template <> struct tl_on<NoTags> { static const bool value = false; };

template <bool GEN, typename TL = NoTags>
__device__ void trace(const DScene& S, int lightset, int sld, const Ray& ray, Flt depth, int recurs, ColorA& outc, Hit& ri,
                      RayCounters& rc, TL tl = TL());

template <bool GEN>
__device__ __forceinline__ void mpreshade(const DScene& S, int lightset, int scene, const Hit& ri, LightCtx& ctx,
                                          RayCounters& rc) {
    // Shader.hs:65-80
    ctx.done = 1;
    ctx.n = 0;
    int first = S.lightsets[2 * lightset], cnt = S.lightsets[2 * lightset + 1];
    for (int li = 0; li < cnt; li++) {
        const GlomeLight* Lp = S.lights + first + li;
        Vec lpos = vec(Lp->pos[0], Lp->pos[1], Lp->pos[2]);
        Vec lvec = vsub(lpos, ri.pos);
        if (vdot(lvec, ri.norm) < 0) continue;
        Flt llen = vlen(lvec);
        Vec ldir = vscale(lvec, 1 / llen);
        bool blocked = llen > Lp->rad;
        if (!blocked && Lp->do_shadow) {
            rc.shadow++;
            blocked = shadow_scene<GEN>(S, scene, mkray(vscaleadd(ri.pos, ri.norm, GLM_DELTA), ldir), llen - (2 * GLM_DELTA), &rc.cnt);
        }
        if (blocked) continue;
        Flt fall = 1 / (llen * llen);  // Shader.hs:23
        if (ctx.n < GDEV_MAX_LIGHTS) {
            ctx.col[ctx.n].r = Lp->color[0] * fall;
            ctx.col[ctx.n].g = Lp->color[1] * fall;
            ctx.col[ctx.n].b = Lp->color[2] * fall;
            ctx.dir[ctx.n] = ldir;
            ctx.n++;
        }
    }
}

// Surface color alpha amb kd ks shine (Shader.hs:90-105); p = {r,g,b,alpha,ambient,kd,ks,shine}
__device__ __forceinline__ void shade_surface(const LightCtx& lights, const Flt* p, const Vec& n, const Vec& eyedir,
                                              ColorA& outc) {
    Flt alpha = p[3], amb = p[4], kd = p[5], ks = p[6], shine = p[7];
    Flt ar = p[0] * amb, ag = p[1] * amb, ab = p[2] * amb;
    Flt dr = 0, dg = 0, db = 0;
    for (int i = 0; i < lights.n; i++) {
        Vec ldir = lights.dir[i];
        Vec halfangle = bisect(ldir, eyedir);
        Flt ldotn = fmax_(0, vdot(ldir, n));
        Flt blinn;
        if (ks <= GLM_DELTA) blinn = 0;
        else {
            Flt b = fmax_(0, pow(vdot(halfangle, n), shine) * ldotn);
            blinn = isnan(b) ? 0 : b;
        }
        Flt diffuse = vdot(ldir, n);
        Flt w = (blinn * ks) + (diffuse * kd);
        dr = dr + lights.col[i].r * w;
        dg = dg + lights.col[i].g * w;
        db = db + lights.col[i].b * w;
    }
    outc = mkca(ar + dr, ag + dg, ab + db, alpha);
}

template <bool GEN, typename TL = NoTags>
__device__ void mpostshade(const DScene& S, int ls, LightCtx& lights, const MatVal& mat, const Ray& ray, int s, const Hit& ri,
                           int recurs, ColorA& outc, RayCounters& rc, TL tl = TL()) {
    // Shader.hs:82-184 (ri is a RayHit here)
    const Vec dir = ray.d;
    const Vec n = ri.norm;
    const Vec p = ri.pos;
    Vec eyedir = vinvert(dir);
    switch (mat.kind) {
        case GLOME_MAT_SURFACE:
            if (!lights.done) mpreshade<GEN>(S, ls, s, ri, lights, rc);
            shade_surface(lights, mat.p, n, eyedir, outc);
            return;
        case GLOME_MAT_BLEND: {  // Shader.hs:181-184
            ColorA ca, cb;
            if constexpr (GEN) {
                MatVal m;
                mat_load(S, mat.a, m);
                mpostshade<true, TL>(S, ls, lights, m, ray, s, ri, recurs, ca, rc, tl);  // tagsa ++ tagsb (Shader.hs:184)
                mat_load(S, mat.b, m);
                mpostshade<true, TL>(S, ls, lights, m, ray, s, ri, recurs, cb, rc, tl);
            } else {
                // flat-class scenes only hold Surface materials: no recursion needed
                if (!lights.done) mpreshade<false>(S, ls, s, ri, lights, rc);
                shade_surface(lights, S.materials[mat.a].p, n, eyedir, ca);
                shade_surface(lights, S.materials[mat.b].p, n, eyedir, cb);
            }
            outc = caweight(ca, cb, mat.p[0]);
            return;
        }
    }
    if constexpr (GEN) {
        switch (mat.kind) {
            case GLOME_MAT_REFLECT: {  // Shader.hs:107-118
                Flt refl = mat.p[0];
                if ((refl > 0) && (recurs > 0)) {
                    Vec outdir = reflect(dir, n);
                    ColorA c;
                    Hit h;
                    rc.secondary++;
                    trace<true, TL>(S, ls, s, mkray(vscaleadd(p, outdir, GLM_DELTA), outdir), GLM_INFINITY, recurs - 1, c, h, rc, tl);  // refltags
                    outc = mkca(c.r, c.g, c.b, c.a * refl);
                } else outc = mkca(0, 0, 0, 1);
                return;
            }
            case GLOME_MAT_REFRACT: {  // Shader.hs:120-155
                Flt refl = mat.p[0], refr = mat.p[1], ior = mat.p[2];
                if ((refl > 0 || refr > 0) && (recurs > 0)) {
                    Vec outdir = reflect(dir, n);
                    ColorA a;
                    Hit h;
                    rc.secondary++;
                    trace<true, TL>(S, ls, s, mkray(vscaleadd(p, outdir, GLM_DELTA), outdir), GLM_INFINITY, recurs - 1, a, h, rc, tl);  // refltags ++
                    Flt eta = (vdot(n, eyedir) > 0) ? ior : 1 / ior;
                    Flt c1 = vdot(dir, n);
                    Flt cs2 = 1 - (eta * eta) * (1 - (c1 * c1));
                    ColorA b;
                    if (cs2 < 0) b = mkca(0, 0, 0, 1);
                    else {
                        Vec t = vadd(vscale(dir, eta), vscale(n, eta * c1 - sqrt(cs2)));
                        rc.secondary++;
                        trace<true, TL>(S, ls, s, mkray(vscaleadd(p, t, GLM_DELTA), t), GLM_INFINITY, recurs - 1, b, h, rc, tl);  // refrtags (Shader.hs:154)
                    }
                    outc = mkca(a.r * refl + b.r * refr, a.g * refl + b.g * refr, a.b * refl + b.b * refr, a.a * refl + b.a * refr);
                } else outc = mkca(0, 0, 0, 0);
                return;
            }
            case GLOME_MAT_WARP: {  // Shader.hs:157-175
                ColorA fc, wc;
                Hit fh, wh;
                rc.secondary += 2;
                Ray wr = xfm_ray(S.dpool + mat.d, mkray(ri.pos, vnorm(ray.d)));  // TestScene.hs:169-173
                if constexpr (tl_on<TL>::value) {  // (fcolor, ftags) or (wcolor, wtags) (Shader.hs:171-175)
                    TagList ft, wt;
                    tl_clear(ft); tl_clear(wt);
                    trace<true, TagList*>(S, ls, mat.a, ri.ray, GLM_INFINITY, recurs - 1, fc, fh, rc, &ft);
                    trace<true, TagList*>(S, mat.c, mat.b, wr, ridepth(fh), recurs - 1, wc, wh, rc, &wt);
                    tl_append(tl, (ridepth(fh) < ridepth(wh)) ? ft : wt);
                } else {
                    trace<true>(S, ls, mat.a, ri.ray, GLM_INFINITY, recurs - 1, fc, fh, rc);
                    trace<true>(S, mat.c, mat.b, wr, ridepth(fh), recurs - 1, wc, wh, rc);
                }
                if (ridepth(fh) < ridepth(wh)) outc = fc;
                else outc = wc;
                return;
            }
            case GLOME_MAT_ADDITIVE: {  // Shader.hs:177-179, casum Clr.hs:93
                Flt r = 0, g = 0, b = 0, prod = 1;
                for (int i = 0; i < mat.b; i++) {
                    ColorA c;
                    MatVal m;
                    mat_load(S, S.ipool[mat.a + i], m);
                    mpostshade<true, TL>(S, ls, lights, m, ray, s, ri, recurs, c, rc, tl);  // concat taglists (Shader.hs:179)
                    r = r + c.r * c.a; g = g + c.g * c.a; b = b + c.b * c.a;
                    prod = prod * (1 - aclamp(c.a));
                }
                outc = mkca(r, g, b, 1 - prod);
                return;
            }
        }
    }
    outc = mkca(0, 0, 0, 0);
}

// tex ray ri  (Solid.hs:97) -> Material value
__device__ __forceinline__ void eval_texture(const DScene& S, int tex, const Hit& ri, MatVal& m, RayCounters& rc) {
    const GlomeTexture* T = S.textures + tex;
    int kind = T->kind;
    if (kind == GLOME_TEX_UNIFORM) { mat_load(S, T->a, m); return; }
    Flt scale;
    if (kind == GLOME_TEX_STRIPE_BLEND) scale = triangle_wave(vdot(ri.pos, vec(T->p[0], T->p[1], T->p[2])));  // TestScene.hs:225
    else {  // perlin (Texture.hs:109-116); out-of-range results are counted, not trapped
        scale = (noise(vscale(ri.pos, T->p[0])) + 1) * 0.5;
        if (scale > 1 || scale < 0) rc.perlin_range++;
    }
    m.kind = GLOME_MAT_BLEND; m.a = T->a; m.b = T->b; m.c = 0; m.d = 0;
    m.p[0] = scale;
}

// trace (Trace.hs:59-82).  ri carries the primary hit's own tag stack; with a TagList* the TraceResult's tag list
// `ts ++ tags` is appended to *tl, where ts = tagsb_k ++ ... ++ tagsb_1 over the hit's textures (Trace.hs:68-79).
template <bool GEN, typename TL>
__device__ void trace(const DScene& S, int lightset, int sld, const Ray& ray, Flt depth, int recurs, ColorA& outc, Hit& ri,
                      RayCounters& rc, TL tl) {
    outc = mkca(0, 0, 0, 0);
    if (recurs == 0) { hit_clear(ri); return; }
    rayint_scene<GEN>(S, sld, ray, depth, ri, &rc.cnt);
    if (!ri.hit) return;  // mmissshade (Shader.hs:186)
    LightCtx ctxb;
    ctxb.done = 0;
    ctxb.n = 0;
    ColorA colora = mkca(0, 0, 0, 0);
    TagList ts;
    if constexpr (tl_on<TL>::value) tl_clear(ts);
    for (int i = 0; i < ri.tex.n; i++) {
        if (colora.a + GLM_DELTA >= 1) continue;  // opaque (Trace.hs:50)
        MatVal m;
        eval_texture(S, ri.tex.v[i], ri, m, rc);
        ColorA colorb;
        if constexpr (tl_on<TL>::value) {
            TagList tb;
            tl_clear(tb);
            mpostshade<GEN, TagList*>(S, lightset, ctxb, m, ray, sld, ri, recurs, colorb, rc, &tb);
            tl_append(&tb, ts);  // tagsb ++ tagsa
            ts = tb;
        } else mpostshade<GEN>(S, lightset, ctxb, m, ray, sld, ri, recurs, colorb, rc);
        colora = cafold(colora, colorb);
    }
    if constexpr (tl_on<TL>::value) {
        tl_append(tl, ts);
        for (int i = 0; i < ri.tag.n; i++) tl_push(tl, ri.tag.v[i]);  // ts ++ tags
    }
    outc = colora;
}

// ---------------------------------------------------------------------------------------------
// Glome.hs: camera rays
// ---------------------------------------------------------------------------------------------
// ---- rayint_debug: the Int half of (Rayint, Int) (get_color_debug's heat map, Glome.hs:57-60) --------------
// The count does not depend on what is hit.  Bih: +1 per branch entered on the reference's own (unculled,
// unclamped) walk, leaves get `fmin d far` (Bih.hs:378-412); [s]: sum (Solid.hs:312,329); Instance: transformed
// ray, d * lenscale (Solid.hs:447-461); Bound: +1 when the gate passes (Bound.hs:37-42); InnerBound: sb
// (Bound.hs:107); Tag / Tex / NoShadow pass through, OnlyShadow is 0 (Tex.hs:55,67,79,90); everything else is the
// class default 0 (Solid.hs:205).  A debugging aid, not tuned: plain device recursion below a Bih.
__device__ int debug_count_node(const DScene& S, int ni, const Ray& r, Flt d) {
    GlomeNode nd = S.nodes[ni];
    while (nd.type == GLOME_TEX || nd.type == GLOME_TAG || nd.type == GLOME_NOSHADOW) { ni = nd.a; nd = S.nodes[ni]; }
    switch (nd.type) {
        case GLOME_GROUP: {
            int c = 0;
            for (int i = 0; i < nd.b; i++) c += debug_count_node(S, nd.a + i, r, d);
            return c;
        }
        case GLOME_INSTANCE: {
            const Flt* xfm = S.dpool + nd.b;
            Vec newdir = invxfm_vec(xfm, r.d);
            Vec neworig = invxfm_point(xfm, r.o);
            Flt lenscale = vlen(newdir);
            Flt invlenscale = 1 / lenscale;
            return debug_count_node(S, nd.a, mkray(neworig, vscale(newdir, invlenscale)), d * lenscale);
        }
        case GLOME_BOUND:
            if (inside_node(S, nd.a, r.o) || shadow_node<-1>(S, nd.a, r, d, 0, nullptr)) return debug_count_node(S, nd.b, r, d) + 1;
            return 0;
        case GLOME_INNERBOUND: return debug_count_node(S, nd.b, r, d);
        case GLOME_BIH: {
            Bbox bb = ldbb(S.dpool + nd.b);
            Flt near_, far_;
            bbclip_ub(r, bb, near_, far_);  // no clip by d at the root (Bih.hs:381)
            const Flt drx = 1 / r.d.x, dry = 1 / r.d.y, drz = 1 / r.d.z;
            const bool linear = (nd.c & GLOME_BIH_LINEAR_SPHERES) != 0;
            TravEnt stack[64];
            int sp = 0, c = 0, ref = nd.a;
            for (;;) {
                bool pop = false;
                if (ref < 0) {
                    if (!linear) {  // a bare sphere counts 0
                        int lf, lc;
                        glome_bih_leaf(ref, S.ipool, &lf, &lc);
                        for (int i = 0; i < lc; i++) c += debug_count_node(S, lf + i, r, fmin_(d, far_));
                    }
                    pop = true;
                } else {
                    c++;
                    if (near_ > far_) pop = true;  // (RayMiss,0) wrapped with 1: only possible at the root
                    else {
                        const GlomeBihNode n = S.bih[ref];
                        Flt dr_ = (n.axis == 0) ? drx : ((n.axis == 1) ? dry : drz);
                        Flt o = (n.axis == 0) ? r.o.x : ((n.axis == 1) ? r.o.y : r.o.z);
                        Flt dl = (n.lsplit - o) * dr_;
                        Flt dr = (n.rsplit - o) * dr_;
                        const bool fwd = dr_ > 0;
                        const Flt dn = fwd ? dl : dr, df = fwd ? dr : dl;
                        const int c1 = fwd ? n.left : n.right, c2 = fwd ? n.right : n.left;
                        const bool v1 = near_ < dn, v2 = df < far_;
                        if (v1 && v2 && sp < 64) { stack[sp].ref = c2; stack[sp].near_ = fmax_(df, near_); stack[sp].far_ = far_; sp++; }
                        if (v1) { ref = c1; far_ = fmin_(dn, far_); }
                        else if (v2) { ref = c2; near_ = fmax_(df, near_); }
                        else pop = true;
                    }
                }
                if (pop) {
                    if (sp == 0) break;
                    sp--;
                    ref = stack[sp].ref; near_ = stack[sp].near_; far_ = stack[sp].far_;
                }
            }
            return c;
        }
        default: return 0;
    }
}

struct DCamera { Vec pos, fwd, up, right; };
__device__ __forceinline__ void getCoordsf(int width, int height, Flt xf, Flt yf, Flt& xc, Flt& yc) {  // Glome.hs:119-140
    Flt widthf = (Flt)width, heightf = (Flt)height;
    xc = (((xf / widthf) * 2) - 1) * (widthf / heightf);
    yc = -(((yf / heightf) * 2) - 1);
}
__device__ __forceinline__ Ray camera_ray(const DCamera& cam, Flt x, Flt y) {  // get_rayint (Glome.hs:27-33)
    Vec dir = vnorm(vadd3(cam.fwd, vscale(cam.right, -x), vscale(cam.up, y)));
    return mkray(cam.pos, dir);
}

}  // namespace gdev
