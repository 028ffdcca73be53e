// glome_math.h -- GlomeVec's scalar/vector algebra (GlomeVec/Data/Glome/Vec.hs) for the product,
// shared by the host scene-construction mirror and the sm_100a kernels.
//
// Bit-parity rules: the operation order of every expression follows the Haskell
// source; min/max are the reference's `>` chains (NaN behaviour differs from IEEE fmin/fmax); the
// translation units that include this header are compiled with -fmad=false (device) and
// -ffp-contract=off (host) because GHC never fuses a multiply-add.
#pragma once
#include <math.h>
#include <stdint.h>

#if defined(__CUDACC__)
#define GLM_HD __host__ __device__ __forceinline__
#else
#define GLM_HD inline
#endif

namespace glm {

// Vec.hs:7-9: `type Flt = Double`, with the note "make separate Float and Double instances of this library".
// GLOME_F32 (device translation unit glome_cuda_f32.cu only) is that Float instance: the optional FP32 mode.
#ifdef GLOME_F32
typedef float Flt;
#else
typedef double Flt;  // Vec.hs:9
#endif

#define FL(x) ((glm::Flt)(x))         /* a literal of the working precision (no silent promotion to double) */
#define GLM_INFINITY FL(1000000.0)  /* Vec.hs:14: not IEEE inf */
#define GLM_DELTA FL(0.0001)        /* Vec.hs:40 */

GLM_HD Flt fmin_(Flt a, Flt b) { return a > b ? b : a; }  // Vec.hs:44
GLM_HD Flt fmax_(Flt a, Flt b) { return a > b ? a : b; }  // Vec.hs:48
GLM_HD Flt fmin3(Flt a, Flt b, Flt c) {                   // Vec.hs:52
    if (a > b) return (b > c) ? c : b;
    return (a > c) ? c : a;
}
GLM_HD Flt fmax3(Flt a, Flt b, Flt c) {                   // Vec.hs:62
    if (a > b) return (a > c) ? a : c;
    return (b > c) ? b : c;
}
GLM_HD Flt fabs_(Flt a) { return a < 0 ? -a : a; }        // Vec.hs:80
GLM_HD Flt hmax(Flt x, Flt y) { return x <= y ? y : x; }  // Prelude max on Double
GLM_HD Flt hmin(Flt x, Flt y) { return x <= y ? x : y; }  // Prelude min on Double

struct Vec { Flt x, y, z; };
struct Ray { Vec o, d; };
struct Bbox { Vec p1, p2; };

GLM_HD Vec vec(Flt x, Flt y, Flt z) { Vec v; v.x = x; v.y = y; v.z = z; return v; }
GLM_HD Flt va(const Vec& v, int n) { return n == 0 ? v.x : (n == 1 ? v.y : v.z); }                 // Vec.hs:167
GLM_HD Flt vdot(const Vec& a, const Vec& b) { return (a.x * b.x) + (a.y * b.y) + (a.z * b.z); }    // Vec.hs:185
GLM_HD Vec vcross(const Vec& a, const Vec& b) {                                                    // Vec.hs:193
    return vec((a.y * b.z) - (a.z * b.y), (a.z * b.x) - (a.x * b.z), (a.x * b.y) - (a.y * b.x));
}
GLM_HD Vec vinvert(const Vec& a) { return vec(-a.x, -a.y, -a.z); }                                 // Vec.hs:213
GLM_HD Flt vlen(const Vec& a) { return sqrt(vdot(a, a)); }                                         // Vec.hs:222
GLM_HD Vec vadd(const Vec& a, const Vec& b) { return vec(a.x + b.x, a.y + b.y, a.z + b.z); }       // Vec.hs:226
GLM_HD Vec vadd3(const Vec& a, const Vec& b, const Vec& c) {                                       // Vec.hs:233
    return vec(a.x + b.x + c.x, a.y + b.y + c.y, a.z + b.z + c.z);
}
GLM_HD Vec vsub(const Vec& a, const Vec& b) { return vec(a.x - b.x, a.y - b.y, a.z - b.z); }       // Vec.hs:240
GLM_HD Vec vscale(const Vec& a, Flt f) { return vec(a.x * f, a.y * f, a.z * f); }                  // Vec.hs:294
GLM_HD Vec vscaleadd(const Vec& a, const Vec& b, Flt f) {                                          // Vec.hs:302
    return vec(a.x + (b.x * f), a.y + (b.y * f), a.z + (b.z * f));
}
GLM_HD Vec vnorm(const Vec& a) {                                                                   // Vec.hs:314
    Flt invlen = FL(1.0) / sqrt((a.x * a.x) + (a.y * a.y) + (a.z * a.z));
    return vec(a.x * invlen, a.y * invlen, a.z * invlen);
}
GLM_HD Vec bisect(const Vec& a, const Vec& b) { return vnorm(vadd(a, b)); }                        // Vec.hs:331
GLM_HD Vec reflect(const Vec& v, const Vec& n) { return vscaleadd(v, n, (-2) * vdot(v, n)); }      // Vec.hs:340
GLM_HD Vec vrcp(const Vec& a) { return vec(1 / a.x, 1 / a.y, 1 / a.z); }                           // Vec.hs:345
GLM_HD Ray mkray(const Vec& o, const Vec& d) { Ray r; r.o = o; r.d = d; return r; }
GLM_HD Ray ray_move(const Ray& r, Flt d) { return mkray(vscaleadd(r.o, r.d, d), r.d); }            // Vec.hs:361
GLM_HD Flt plane_int_dist(const Ray& r, const Vec& p, const Vec& norm) {                           // Vec.hs:391
    Vec newo = vsub(r.o, p);
    return -(vdot(norm, newo)) / (vdot(norm, r.d));
}

// Xfm = 24 doubles: forward 3x4 Matrix (row major) then inverse (Vec.hs:407-414)
GLM_HD Vec xfm_point(const Flt* m, const Vec& v) {                                                 // Vec.hs:502
    return vec(m[0] * v.x + m[1] * v.y + m[2] * v.z + m[3], m[4] * v.x + m[5] * v.y + m[6] * v.z + m[7],
               m[8] * v.x + m[9] * v.y + m[10] * v.z + m[11]);
}
GLM_HD Vec xfm_vec(const Flt* m, const Vec& v) {                                                   // Vec.hs:522
    return vec(m[0] * v.x + m[1] * v.y + m[2] * v.z, m[4] * v.x + m[5] * v.y + m[6] * v.z,
               m[8] * v.x + m[9] * v.y + m[10] * v.z);
}
GLM_HD Vec invxfm_point(const Flt* m, const Vec& v) { return xfm_point(m + 12, v); }               // Vec.hs:512
GLM_HD Vec invxfm_vec(const Flt* m, const Vec& v) { return xfm_vec(m + 12, v); }                   // Vec.hs:532
GLM_HD Vec invxfm_norm(const Flt* m, const Vec& v) {                                               // Vec.hs:543
    const Flt* i = m + 12;
    return vec(i[0] * v.x + i[4] * v.y + i[8] * v.z, i[1] * v.x + i[5] * v.y + i[9] * v.z,
               i[2] * v.x + i[6] * v.y + i[10] * v.z);
}
GLM_HD Ray xfm_ray(const Flt* m, const Ray& r) { return mkray(xfm_point(m, r.o), vnorm(xfm_vec(m, r.d))); }  // Vec.hs:553

GLM_HD Bbox mkbb(const Vec& p1, const Vec& p2) { Bbox b; b.p1 = p1; b.p2 = p2; return b; }
GLM_HD Bbox bbjoin(const Bbox& a, const Bbox& b) {                                                 // Vec.hs:652
    return mkbb(vec(fmin_(a.p1.x, b.p1.x), fmin_(a.p1.y, b.p1.y), fmin_(a.p1.z, b.p1.z)),
                vec(fmax_(a.p2.x, b.p2.x), fmax_(a.p2.y, b.p2.y), fmax_(a.p2.z, b.p2.z)));
}
GLM_HD Bbox bboverlap(const Bbox& a, const Bbox& b) {                                              // Vec.hs:657
    return mkbb(vec(fmax_(a.p1.x, b.p1.x), fmax_(a.p1.y, b.p1.y), fmax_(a.p1.z, b.p1.z)),
                vec(fmin_(a.p2.x, b.p2.x), fmin_(a.p2.y, b.p2.y), fmin_(a.p2.z, b.p2.z)));
}
GLM_HD Flt bbsa(const Bbox& b) {                                                                   // Vec.hs:694
    Vec d = vsub(b.p2, b.p1);
    return hmax(0, 2 * (d.x * d.y + d.x * d.z + d.y * d.z));
}
GLM_HD Bbox empty_bbox() {                                                                         // Vec.hs:706
    return mkbb(vec(GLM_INFINITY, GLM_INFINITY, GLM_INFINITY), vec(-GLM_INFINITY, -GLM_INFINITY, -GLM_INFINITY));
}
GLM_HD Bbox everything_bbox() {                                                                    // Vec.hs:712
    return mkbb(vec(-GLM_INFINITY, -GLM_INFINITY, -GLM_INFINITY), vec(GLM_INFINITY, GLM_INFINITY, GLM_INFINITY));
}
GLM_HD Vec bbmid(const Bbox& b) { return vscale(vadd(b.p1, b.p2), FL(0.5)); }                          // Bih.hs:162

// one slab of bbclip_ub / bbclip_ub_rcp: `pos` is the sign test the caller chose
GLM_HD void slab(bool pos, Flt p1, Flt p2, Flt o, Flt rcp, Flt& in, Flt& out) {
    if (pos) { in = (p1 - o) * rcp; out = (p2 - o) * rcp; }
    else { in = (p2 - o) * rcp; out = (p1 - o) * rcp; }
}
// bbclip_ub (Vec.hs:743-762): sign test on d (a +0.0 component always misses)
GLM_HD void bbclip_ub(const Ray& r, const Bbox& b, Flt& near_, Flt& far_) {
    Flt dxrcp = 1 / r.d.x, dyrcp = 1 / r.d.y, dzrcp = 1 / r.d.z;
    Flt inx, outx, iny, outy, inz, outz;
    slab(r.d.x > 0, b.p1.x, b.p2.x, r.o.x, dxrcp, inx, outx);
    slab(r.d.y > 0, b.p1.y, b.p2.y, r.o.y, dyrcp, iny, outy);
    slab(r.d.z > 0, b.p1.z, b.p2.z, r.o.z, dzrcp, inz, outz);
    near_ = fmax3(inx, iny, inz);
    far_ = fmin3(outx, outy, outz);
}
// bbclip_ub_rcp (Vec.hs:725-741): sign test on the reciprocal
GLM_HD void bbclip_ub_rcp(const Vec& o, const Vec& rcp, const Bbox& b, Flt& near_, Flt& far_) {
    Flt inx, outx, iny, outy, inz, outz;
    slab(rcp.x > 0, b.p1.x, b.p2.x, o.x, rcp.x, inx, outx);
    slab(rcp.y > 0, b.p1.y, b.p2.y, o.y, rcp.y, iny, outy);
    slab(rcp.z > 0, b.p1.z, b.p2.z, o.z, rcp.z, inz, outz);
    near_ = fmax3(inx, iny, inz);
    far_ = fmin3(outx, outy, outz);
}

}  // namespace glm
