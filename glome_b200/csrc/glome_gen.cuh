// glome_gen.cuh -- the general scene graph without recursion: one iterative machine per ray.
//
// The reference evaluates `rayint` / `shadow` / `inside` / `get_metainfo` and `trace` / `materialShader` by mutual
// recursion over the SolidItem graph (Solid.hs:138-275, Csg.hs, Bound.hs, Bih.hs, Trace.hs:59-82, Shader.hs:82-184).
// Round 1 ran that recursion literally on the device (40 KB of stack per thread, 882 KB of SASS, 20 GB of local-memory
// traffic per frame).  Here every one of those functions is a loop over a small explicit control stack:
//
//   QVM      rayint / shadow of any node: states ENTER (dispatch on the node), BRANCH (one BIH step, Bih.hs:332-368,
//            510-544), LIST (one item of a group / BIH leaf, Solid.hs:326-339) and RET (pop a continuation).  The two
//            hot shapes -- `{Tex,Tag}* prim` and `Instance ({Tex,Tag}* prim)` -- are tested in LIST without a frame, at
//            the only primitive-test site of the kernel.  Instance (Solid.hs:386-403, 464-471), Difference /
//            Intersection incl. rayint_advance (Csg.hs:33-90, Solid.hs:85-91), Bound / InnerBound (Bound.hs:30-49,
//            95-103) are continuations.  Hit records live in a small slot stack and are written only when a candidate
//            wins; the nearest depth of the current fold stays in a register.
//   inside   gq_inside: boolean expression walk with short circuit (Solid.hs:331, Csg.hs:92-101, Bih.hs:550-565)
//   metainfo gq_metainfo: depth-first walk that emits the texture / tag lists in their final order
//            (Solid.hs:337-339, Bih.hs:567-585, Csg.hs:103-111, Tex.hs:61-74, Bound.hs:54-58)
//   SHM      trace + materialShader: trace frames (one per reflection generation) and material continuations
//            (Blend, AdditiveLayers, Reflect, Refract, Warp); Surface shading asks the QVM for its shadow rays
//
// Arithmetic per node is the reference's, operation for operation (same primitives as the flat kernels); only the
// control flow is restated.  CSG keeps the reference's advance-and-retry marching, so results equal the oracle's bit
// for bit: (t_k + a_k) + ... + a_1 is summed innermost first from a stack of pending offsets.
//
// Everything here is host+device: tests/tools/gen_host.cpp compiles this file with g++ to check the machine against
// the oracle on the CPU (test infrastructure; the product only ever runs it inside the sm_100a kernels).
#pragma once
#include "glome_device.cuh"

namespace ggen {

using namespace gdev;

// ---------------------------------------------------------------------------------------------
// packed texture / tag stacks: 8 entries x 16 bits in two 64-bit registers.  Entry = id + 1, 0 = empty; the head
// (innermost = last pushed) sits in the low 16 bits of `lo`.  Texture ids are scene indices; tag values are arbitrary
// int32 in the reference's API, so general-class scenes carry dense tag ids on the device (DScene.tagvals maps back).
// ---------------------------------------------------------------------------------------------
#define GGEN_MAX_ID 65534
struct PStk { unsigned long long lo, hi; };
GD_FN PStk pstk_empty() { PStk s; s.lo = 0; s.hi = 0; return s; }
GD_FN bool pstk_cons(PStk& s, int x) {  // x : s ; true on overflow (the oldest entry falls off)
    const bool ovf = (s.hi >> 48) != 0;
    s.hi = (s.hi << 16) | (s.lo >> 48);
    s.lo = (s.lo << 16) | (unsigned long long)(unsigned int)(x + 1);
    return ovf;
}
GD_FN int pstk_get(const PStk& s, int i) {  // i-th entry from the head, -1 = none
    const unsigned long long w = i < 4 ? s.lo : s.hi;
    return (int)((w >> (16 * (i & 3))) & 0xffffull) - 1;
}
GD_FN int pstk_n(const PStk& s) {
    int n = 0;
    for (int i = 0; i < 8; i++) n += pstk_get(s, i) >= 0;
    return n;
}
GD_FN void pstk_to_stk(const PStk& p, Stk& s) {
    s.n = 0;
    for (int i = 0; i < GLOME_MAX_STACK; i++) {
        int v = pstk_get(p, i);
        s.v[i] = v >= 0 ? v : 0;
        if (v >= 0) s.n = i + 1;
    }
}
GD_FN bool pstk_from_stk(PStk& p, const Stk& s) {  // true on overflow
    p = pstk_empty();
    bool ovf = false;
    for (int i = s.n - 1; i >= 0; i--) ovf |= pstk_cons(p, s.v[i]);
    return ovf;
}

// Rayint (Solid.hs:20-28) as kept by the machine
struct GHit {
    Flt t;
    Vec pos, norm;
    Ray ray;       // riray: the ray the winning primitive was tested with (object space under an Instance)
    PStk tex, tag;
    int hit, prim, sub, flags;
};
GD_FN void ghit_clear(GHit& h) {
    h.t = GLM_INFINITY; h.hit = 0; h.prim = -1; h.sub = -1; h.flags = 0;
    h.tex = pstk_empty(); h.tag = pstk_empty();
}
GD_FN Flt ghit_depth(const GHit& h) { return h.hit ? h.t : (Flt)GLM_INFINITY; }  // ridepth (Solid.hs:33)

struct GCnt { unsigned int bih, prim, inst, csg, shadow, secondary, perlin, bvh, tri; };
GD_FN void gcnt_clear(GCnt& c) { c.bih = c.prim = c.inst = c.csg = c.shadow = c.secondary = c.perlin = c.bvh = c.tri = 0; }

GD_FN bool is_wrap_r(int t) { return t == GLOME_TEX || t == GLOME_TAG || t == GLOME_NOSHADOW; }    // transparent to rayint
GD_FN bool is_wrap_s(int t) { return t == GLOME_TEX || t == GLOME_TAG || t == GLOME_ONLYSHADOW; }  // transparent to shadow
GD_FN bool is_wrap_any(int t) { return t == GLOME_TEX || t == GLOME_TAG || t == GLOME_NOSHADOW || t == GLOME_ONLYSHADOW; }

// ---------------------------------------------------------------------------------------------
// inside (Solid.hs:166): iterative boolean walk
// ---------------------------------------------------------------------------------------------
#define GI_FRAMES 48
#define GI_PTS 8
enum { GI_OR_LIST = 0, GI_AND_LIST, GI_PT, GI_BIH_R, GI_DIFF, GI_NOT, GI_AND_NODE, GI_OR_NODE };

// inside s pt.  *ovf is set when the walk runs out of frames (the answer is then `false`).
GD_NOINLINE bool gq_inside(const DScene& S, int root, const Vec& pt0, int* ovf) {
    int fk[GI_FRAMES], fa[GI_FRAMES], fb[GI_FRAMES];
    Vec pts[GI_PTS];
    int fp = 0, np = 0;
    Vec pt = pt0;
    bool v = false;
    int n = root;      // node to evaluate (call) ...
    int bref = 0;      // ... or BIH ref to descend
    int mode = 0;      // 0 = call node n, 1 = call bih ref, 2 = return v
    for (;;) {
        if (mode == 0) {
            GlomeNode nd = S.nodes[n];
            while (is_wrap_any(nd.type)) { n = nd.a; nd = S.nodes[n]; }  // Tex.hs:59,71,83,94
            if (is_prim(nd.type)) { v = prim_inside(S, nd, pt); mode = 2; continue; }
            switch (nd.type) {
                case GLOME_GROUP:          // Solid.hs:331: or
                case GLOME_INTERSECTION:   // Csg.hs:99-101: and
                    if (nd.b == 0) { v = nd.type == GLOME_INTERSECTION; mode = 2; break; }
                    if (fp >= GI_FRAMES) { *ovf = 1; return false; }
                    fk[fp] = nd.type == GLOME_GROUP ? GI_OR_LIST : GI_AND_LIST; fa[fp] = nd.a + 1; fb[fp] = nd.a + nd.b; fp++;
                    n = nd.a;
                    break;
                case GLOME_INSTANCE:       // Solid.hs:473
                    if (fp >= GI_FRAMES || np >= GI_PTS) { *ovf = 1; return false; }
                    fk[fp] = GI_PT; fa[fp] = 0; fb[fp] = 0; fp++;
                    pts[np++] = pt;
                    pt = invxfm_point(S.dpool + nd.b, pt);
                    n = nd.a;
                    break;
                case GLOME_BIH: {          // Bih.hs:550-565: strict box test, then the point descent
                    const Flt* b = S.dpool + nd.b;
                    if ((pt.x > b[0]) && (pt.x < b[3]) && (pt.y > b[1]) && (pt.y < b[4]) && (pt.z > b[2]) && (pt.z < b[5])) {
                        bref = nd.a; mode = 1;
                    } else { v = false; mode = 2; }
                    break;
                }
                case GLOME_DIFFERENCE:     // Csg.hs:92: inside a && not (inside b)
                    if (fp >= GI_FRAMES) { *ovf = 1; return false; }
                    fk[fp] = GI_DIFF; fa[fp] = nd.b; fb[fp] = 0; fp++;
                    n = nd.a;
                    break;
                case GLOME_BOUND:          // Bound.hs:51
                case GLOME_INNERBOUND:     // Bound.hs:109
                    if (fp >= GI_FRAMES) { *ovf = 1; return false; }
                    fk[fp] = nd.type == GLOME_BOUND ? GI_AND_NODE : GI_OR_NODE; fa[fp] = nd.b; fb[fp] = 0; fp++;
                    n = nd.a;
                    break;
                default: v = false; mode = 2; break;  // Void, Mesh (Mesh.hs:211)
            }
            continue;
        }
        if (mode == 1) {
            if (bref < 0) {
                int first, cnt;
                glome_bih_leaf(bref, S.ipool, &first, &cnt);
                if (cnt == 0) { v = false; mode = 2; continue; }
                if (fp >= GI_FRAMES) { *ovf = 1; return false; }
                fk[fp] = GI_OR_LIST; fa[fp] = first + 1; fb[fp] = first + cnt; fp++;
                n = first; mode = 0;
                continue;
            }
            const BihStep bs_ = ld_bih(S.bih, bref);
            const BihStep* bn = &bs_;
            const Flt o = va(pt, bn->axis);
            const bool gl = o < bn->ls, gr = o > bn->rs;
            if (gl && gr) {
                if (fp >= GI_FRAMES) { *ovf = 1; return false; }
                fk[fp] = GI_BIH_R; fa[fp] = bn->right; fb[fp] = 0; fp++;
                bref = bn->left;
            } else if (gl) bref = bn->left;
            else if (gr) bref = bn->right;
            else { v = false; mode = 2; }
            continue;
        }
        // mode 2: hand v to the frame on top
        if (fp == 0) return v;
        fp--;
        const int k = fk[fp], a = fa[fp], b = fb[fp];
        switch (k) {
            case GI_OR_LIST:
                if (v || a == b) break;
                fk[fp] = GI_OR_LIST; fa[fp] = a + 1; fb[fp] = b; fp++;
                n = a; mode = 0;
                break;
            case GI_AND_LIST:
                if (!v || a == b) break;
                fk[fp] = GI_AND_LIST; fa[fp] = a + 1; fb[fp] = b; fp++;
                n = a; mode = 0;
                break;
            case GI_PT: pt = pts[--np]; break;
            case GI_BIH_R: if (!v) { bref = a; mode = 1; } break;
            case GI_DIFF:
                if (!v) break;
                fk[fp] = GI_NOT; fa[fp] = 0; fb[fp] = 0; fp++;
                n = a; mode = 0;
                break;
            case GI_NOT: v = !v; break;
            case GI_AND_NODE: if (v) { n = a; mode = 0; } break;
            case GI_OR_NODE: if (!v) { n = a; mode = 0; } break;
        }
    }
}

// inside of every element of [first, first+count)  (Csg.hs:99-101)
GD_FN bool gq_inside_all(const DScene& S, int first, int count, const Vec& pt, int* ovf) {
    for (int i = 0; i < count; i++) {
        const GlomeNode c = S.nodes[first + i];
        if (is_prim(c.type)) { if (!prim_inside(S, c, pt)) return false; continue; }  // the planes of a polyhedron
        if (!gq_inside(S, first + i, pt, ovf)) return false;
    }
    return true;
}

// ---------------------------------------------------------------------------------------------
// get_metainfo (Solid.hs:200): depth-first walk emitting both lists in their final order
// ---------------------------------------------------------------------------------------------
#define GM_ITEMS 48
#define GM_PTS 8
enum { GM_VISIT = 0, GM_LIST_DESC, GM_LIST_ASC, GM_BIHREF, GM_POP_PT };

GD_FN void gm_emit(Stk& s, int x, int& flags) {
    if (s.n < GLOME_MAX_STACK) s.v[s.n++] = x;
    else flags |= GLOME_HITFLAG_STACK_OVERFLOW;
}

GD_NOINLINE void gq_metainfo(const DScene& S, int root, const Vec& v, PStk& texs_out, PStk& tags_out, int& flags) {
    int ik[GM_ITEMS], ia[GM_ITEMS], ib[GM_ITEMS], ip[GM_ITEMS];
    Vec pts[GM_PTS];
    int sp = 0, np = 1, ovf = 0;
    Stk texs, tags;
    stk_clear(texs);
    stk_clear(tags);
    pts[0] = v;
#define GM_PUSH(K, A, B, P)                                                               \
    do {                                                                                  \
        if (sp >= GM_ITEMS) { flags |= GLOME_HITFLAG_STACK_OVERFLOW; sp = 0; goto gm_done; } \
        ik[sp] = (K); ia[sp] = (A); ib[sp] = (B); ip[sp] = (P); sp++;                     \
    } while (0)
    GM_PUSH(GM_VISIT, root, 0, 0);
    while (sp > 0) {
        sp--;
        const int k = ik[sp], a = ia[sp], b = ib[sp], p = ip[sp];
        const Vec pt = pts[p];
        if (k == GM_POP_PT) { np = a; continue; }
        if (k == GM_LIST_DESC) {
            // [s] (Solid.hs:337-339): every element that contains the point, later elements first; a = first, b = current
            if (b > a) GM_PUSH(GM_LIST_DESC, a, b - 1, p);
            if (gq_inside(S, b, pt, &ovf)) GM_PUSH(GM_VISIT, b, 0, p);
            continue;
        }
        if (k == GM_LIST_ASC) {  // Intersection (Csg.hs:108-111): all elements in order; a = current, b = end
            if (a + 1 < b) GM_PUSH(GM_LIST_ASC, a + 1, b, p);
            GM_PUSH(GM_VISIT, a, 0, p);
            continue;
        }
        if (k == GM_BIHREF) {  // Bih.hs:568-577
            if (a < 0) {
                int first, cnt;
                glome_bih_leaf(a, S.ipool, &first, &cnt);
                if (cnt > 0) GM_PUSH(GM_LIST_DESC, first, first + cnt - 1, p);
                continue;
            }
            const BihStep bs_ = ld_bih(S.bih, a);
            const BihStep* bn = &bs_;
            const Flt o = va(pt, bn->axis);
            if (o > bn->rs) GM_PUSH(GM_BIHREF, bn->right, 0, p);  // left part ++ right part: left is popped first
            if (o < bn->ls) GM_PUSH(GM_BIHREF, bn->left, 0, p);
            continue;
        }
        const GlomeNode nd = S.nodes[a];
        switch (nd.type) {
            case GLOME_GROUP:
                if (nd.b > 0) GM_PUSH(GM_LIST_DESC, nd.a, nd.a + nd.b - 1, p);
                break;
            case GLOME_INSTANCE:  // Solid.hs:517
                if (np >= GM_PTS) { flags |= GLOME_HITFLAG_STACK_OVERFLOW; break; }
                GM_PUSH(GM_POP_PT, np, 0, 0);
                pts[np] = invxfm_point(S.dpool + nd.b, pt);
                GM_PUSH(GM_VISIT, nd.a, 0, np);
                np++;
                break;
            case GLOME_BIH: {  // Bih.hs:579-585
                const Flt* bb = S.dpool + nd.b;
                if ((pt.x > bb[0]) && (pt.x < bb[3]) && (pt.y > bb[1]) && (pt.y < bb[4]) && (pt.z > bb[2]) && (pt.z < bb[5]))
                    GM_PUSH(GM_BIHREF, nd.a, 0, p);
                break;
            }
            case GLOME_DIFFERENCE:  // Csg.hs:103-106
                if (gq_inside(S, nd.a, pt, &ovf) && !gq_inside(S, nd.b, pt, &ovf)) GM_PUSH(GM_VISIT, nd.a, 0, p);
                break;
            case GLOME_INTERSECTION:
                if (nd.b > 0 && gq_inside_all(S, nd.a, nd.b, pt, &ovf)) GM_PUSH(GM_LIST_ASC, nd.a, nd.a + nd.b, p);
                break;
            case GLOME_TEX:  // Tex.hs:73-74: t : child's
                gm_emit(texs, nd.b, flags);
                GM_PUSH(GM_VISIT, nd.a, 0, p);
                break;
            case GLOME_TAG:  // Tex.hs:61-62
                gm_emit(tags, nd.b, flags);
                GM_PUSH(GM_VISIT, nd.a, 0, p);
                break;
            case GLOME_NOSHADOW:
            case GLOME_ONLYSHADOW: GM_PUSH(GM_VISIT, nd.a, 0, p); break;
            case GLOME_BOUND:  // Bound.hs:54-58
                if (gq_inside(S, nd.a, pt, &ovf)) GM_PUSH(GM_VISIT, nd.b, 0, p);
                break;
            case GLOME_INNERBOUND: GM_PUSH(GM_VISIT, nd.b, 0, p); break;  // Bound.hs:112
            default: break;  // primitives, Void, Mesh: ([],[])  (Solid.hs:254)
        }
    }
gm_done:
#undef GM_PUSH
    if (ovf) flags |= GLOME_HITFLAG_STACK_OVERFLOW;
    if (pstk_from_stk(texs_out, texs)) flags |= GLOME_HITFLAG_STACK_OVERFLOW;
    if (pstk_from_stk(tags_out, tags)) flags |= GLOME_HITFLAG_STACK_OVERFLOW;
}


// ---------------------------------------------------------------------------------------------
// QVM: rayint / shadow
// ---------------------------------------------------------------------------------------------
#ifndef GLOME_GROUP_ACCEL
#define GLOME_GROUP_ACCEL 2 /* GlomeNode.c of a GROUP: (ipool offset of its implicit-BIH record << 4) | 2 (glome_tagmap.h) */
#endif
#ifndef GQ_ROUND_BUDGET
#define GQ_ROUND_BUDGET 0
#endif
#define GQ_WORDS 224    /* control stack, 8-byte words */
#define GQ_SLOTS 12     /* hit slots */
#define GQ_ADV_CAP 4096 /* rayint_advance re-issues per query before the machine gives up (flagged) */

// Control-stack entries, bottom -> top; the LAST word of an entry is its header (opcode in the low 8 bits, two 28-bit
// signed ints above it).
enum {
    GF_ROOT = 1,    // bottom of a query                                        [hdr]
    GF_TRAV,        // pending BIH subtree                                      [far near hdr(ref in bits 32..63)]
    GF_BIH,         // bottom of a BIH's TRAV entries                           [hdr(previous BIH node)]
    GF_LIST,        // resume a list after a complex item                       [d ld hdr(next, end)]
    GF_CTX,         // restore the texture / tag context                        [tex.lo tex.hi tag.lo tag.hi hdr]
    GF_INST,        // Instance (Solid.hs:386-403, 464-471)                     [o.xyz d.xyz d invlenscale 1/d.xyz cull hdr(node, parent slot)]
    GF_MODE,        // a rayint stood in for a shadow (Solid.hs:218-221)        [hdr(slot, previous acc)]
    GF_GATE,        // Bound (Bound.hs:30-49): shadow of the bounding object    [hdr(bounded node, previous mode)]
    GF_OR,          // InnerBound shadow (Bound.hs:101-103)                     [hdr(second node)]
    GF_INNER,       // InnerBound rayint (Bound.hs:98-99)                       [tex.lo tex.hi tag.lo tag.hi d hdr(outer node, parent slot)]
    GF_SETD,        // restore the distance limit                               [d hdr]
    GF_DIFF,        // Difference (Csg.hs:33-54)          [o.xyz d h2(R, parent)] ADD* [hdr(node, phase)]
    GF_ADD,         // rayint_advance's depth fix-up (Solid.hs:91)              [a hdr]
    GF_ISECT_BASE,  // Intersection (Csg.hs:68-90)        [o.xyz d h2(R, parent) hdr(node)] (ADD | ELSE_ADV)* [GF_ISECT]
    GF_ISECT,       // Intersection: element k is being evaluated               [hdr(k, phase | in << 2)]
    GF_ELSE_ADV     // Intersection: "if the rest misses, advance past k"       [o.xyz d t hdr(k)]
};

struct QVM {
    unsigned long long cs[GQ_WORDS];
    GHit slot[GQ_SLOTS];
};

GD_FN unsigned long long gq_hdr(int op, int a, int b) {
    return (unsigned long long)(unsigned int)op | ((unsigned long long)((unsigned int)a & 0xfffffffu) << 8) |
           ((unsigned long long)((unsigned int)b & 0xfffffffu) << 36);
}
GD_FN int gq_op(unsigned long long h) { return (int)(h & 0xff); }
GD_FN int gq_a(unsigned long long h) { return ((int)(((unsigned int)(h >> 8)) << 4)) >> 4; }   // sign-extended 28 bits
GD_FN int gq_b(unsigned long long h) { return ((int)(((unsigned int)(h >> 36)) << 4)) >> 4; }
GD_FN unsigned long long gq_d2w(Flt x) { return flt_key(x); }  // the bits of a Flt in a control-stack word (any sign)
GD_FN Flt gq_w2d(unsigned long long w) { return key_flt(w); }

enum { GS_ENTER = 0, GS_BRANCH, GS_LIST, GS_RET, GS_DONE };

// item descriptor classes / flags (glome_tagmap.h: build_items)
enum { GI_DEAD = 0, GI_PRIM = 1, GI_INST_PRIM = 2, GI_COMPLEX = 3 };
#define GI_VIS_R 0x100
#define GI_VIS_S 0x200
#define GI_WRAPPED 0x400

// the primitive tests of the machine: rayint with position + normal, and shadow.  (rx, ry, rz) = 1 / ray.d, computed
// once per ray by the caller (only the box uses it).
#ifdef GQ_RARE_CALL
// the primitive kinds that are rare in a list next to spheres, boxes and planes: out of line, so that the list loop's
// own code stays small (the loop is instruction-fetch bound, profiles/README.md)
GD_NOINLINE bool gq_prim_rayint_rare(const DScene& S, int type, int payload, const Ray& r, Flt d, Flt& t, Vec& pos, Vec& n) {
    const Flt* p = S.dpool + payload;
    switch (type) {
        case GLOME_TRIANGLE: {
            Vec z = vec(0, 0, 0);
            return prim_triangle<true>(ldv(p), ldv(p + 3), ldv(p + 6), false, z, z, z, r, d, t, pos, n);
        }
        case GLOME_TRIANGLENORM:
            return prim_triangle<true>(ldv(p), ldv(p + 3), ldv(p + 6), true, ldv(p + 9), ldv(p + 12), ldv(p + 15), r, d, t, pos, n);
        case GLOME_DISC: return prim_disc_v<true>(ldv(p), ldv(p + 3), p[6], r, d, t, pos, n);
        case GLOME_CYLINDER: return prim_cylinder<true>(p, r, d, t, pos, n);
        case GLOME_CONE: return prim_cone<true>(p, r, d, t, pos, n);
    }
    return false;
}
#endif
GD_FN bool gq_prim_rayint(const DScene& S, int type, int payload, const Ray& r, Flt rx, Flt ry, Flt rz, Flt d, Flt& t, Vec& pos,
                          Vec& n) {
    const Flt* p = S.dpool + payload;
#ifdef GQ_RARE_CALL
    switch (type) {
        case GLOME_SPHERE: return prim_sphere<true>(p, r, d, t, pos, n);
        case GLOME_BOX: return prim_box_rcp<true>(p, r, rx, ry, rz, d, t, pos, n);
        case GLOME_PLANE: return prim_plane<true>(p, r, d, t, pos, n);
    }
    return gq_prim_rayint_rare(S, type, payload, r, d, t, pos, n);
#endif
    switch (type) {
        case GLOME_SPHERE: return prim_sphere<true>(p, r, d, t, pos, n);
        case GLOME_TRIANGLE: {
            Vec z = vec(0, 0, 0);
            return prim_triangle<true>(ldv(p), ldv(p + 3), ldv(p + 6), false, z, z, z, r, d, t, pos, n);
        }
        case GLOME_TRIANGLENORM:
            return prim_triangle<true>(ldv(p), ldv(p + 3), ldv(p + 6), true, ldv(p + 9), ldv(p + 12), ldv(p + 15), r, d, t, pos, n);
        case GLOME_BOX: return prim_box_rcp<true>(p, r, rx, ry, rz, d, t, pos, n);
        case GLOME_PLANE: return prim_plane<true>(p, r, d, t, pos, n);
        case GLOME_DISC: return prim_disc_v<true>(ldv(p), ldv(p + 3), p[6], r, d, t, pos, n);
        case GLOME_CYLINDER: return prim_cylinder<true>(p, r, d, t, pos, n);
        case GLOME_CONE: return prim_cone<true>(p, r, d, t, pos, n);
    }
    return false;
}
// the primitives with a shadow method of their own (Sphere.hs:51-71, Triangle.hs:82-107, Box.hs:56-62); the others fall
// back on rayint's hit flag (Solid.hs:218-221; shadow_cone, Cone.hs:206-245, is rayint_cone's hit test)
GD_FN bool gq_has_shadow_method(int type) { return type <= GLOME_BOX; }
GD_FN bool gq_prim_shadow(const DScene& S, int type, int payload, const Ray& r, Flt rx, Flt ry, Flt rz, Flt d) {
    const Flt* p = S.dpool + payload;
    switch (type) {
        case GLOME_SPHERE: return shadow_sphere(p, r, d);
        case GLOME_TRIANGLE:
        case GLOME_TRIANGLENORM: return shadow_triangle(ldv(p), ldv(p + 3), ldv(p + 6), r, d);
        case GLOME_BOX: return shadow_box_rcp(p, r, rx, ry, rz, d);
    }
    return false;
}

#define GI_HAS_META 0x800 /* a Tex or Tag somewhere below: get_metainfo can return something */

// A primitive hit as the test produces it: world-space depth / position / normal, and the ray it was tested with
struct SimpleHit { Flt t; Vec pos, norm; Ray ray; };

// rayint / shadow of a simple item -- `{Tex,Tag}* prim` or `{..}* Instance ({..}* prim)` (Solid.hs:388-403) -- described by
// its item record.  The ONE copy of the primitive tests in a kernel: the list loop and the CSG evaluators all call it.
// (rx, ry, rz) = 1 / r.d of the caller's ray.  smode: only the hit flag is wanted (shadow).
GD_FN bool gq_test_simple_inl(const DScene& S, int4 it, Flt ox, Flt oy, Flt oz, Flt dx, Flt dy, Flt dz, Flt rx, Flt ry, Flt rz, Flt d,
                              bool smode, SimpleHit* out, GCnt* cnt) {
    const int cls = it.x & 15, ptype = (it.x >> 4) & 15;
    Ray tr = mkray(vec(ox, oy, oz), vec(dx, dy, dz));
    Flt td = d, invls = 1;
    const Flt* xfm = nullptr;
    if (cls == GI_INST_PRIM) {
        cnt->inst++;
        xfm = S.dpool + it.w;
        const Vec newdir = invxfm_vec(xfm, tr.d);
        const Vec neworig = invxfm_point(xfm, tr.o);
        const Flt lenscale = vlen(newdir);
        invls = 1 / lenscale;
        tr = mkray(neworig, vscale(newdir, invls));
        td = d * lenscale;
        if (ptype == GLOME_BOX) { rx = 1 / tr.d.x; ry = 1 / tr.d.y; rz = 1 / tr.d.z; }
    }
    cnt->prim++;
    if (smode && gq_has_shadow_method(ptype)) return gq_prim_shadow(S, ptype, it.z, tr, rx, ry, rz, td);
    Flt t; Vec pos, n;
    if (!gq_prim_rayint(S, ptype, it.z, tr, rx, ry, rz, td, t, pos, n)) return false;
    if (smode) return true;
    out->ray = tr;
    if (xfm) { out->t = t * invls; out->pos = xfm_point(xfm, pos); out->norm = vnorm(invxfm_norm(xfm, n)); }
    else { out->t = t; out->pos = pos; out->norm = n; }
    return true;
}

// the shared copy for the CSG evaluators
GD_NOINLINE bool gq_test_simple(const DScene& S, int4 it, Flt ox, Flt oy, Flt oz, Flt dx, Flt dy, Flt dz, Flt rx, Flt ry, Flt rz, Flt d,
                                bool smode, SimpleHit* out, GCnt* cnt) {
    return gq_test_simple_inl(S, it, ox, oy, oz, dx, dy, dz, rx, ry, rz, d, smode, out, cnt);
}

// record a simple item's hit: the context stacks plus the item's own wrappers, outermost first (Tex.hs:54,66)
// (out of line: it runs only when a candidate wins, and the list loop around it is instruction-fetch bound -- config 1
// 11.4 -> 10.7 ms on B200; the rarer primitive kinds out of line as well, GQ_RARE_CALL, cost 3-8 %)
#ifndef GQ_FILL_INLINE
GD_NOINLINE
#else
GD_FN
#endif
void gq_fill_hit(const DScene& S, GHit& a, const int4& it, int item, const SimpleHit& h, const PStk& ctex, const PStk& ctag,
                       int& mflags) {
    a.hit = 1; a.t = h.t; a.pos = h.pos; a.norm = h.norm; a.ray = h.ray; a.prim = it.y; a.sub = -1;
    PStk wtx = ctex, wtg = ctag;
    if (it.x & GI_WRAPPED) {
        int wj = item;
        GlomeNode w = S.nodes[wj];
        for (;;) {
            if (w.type == GLOME_TEX) { if (pstk_cons(wtx, w.b)) mflags |= GLOME_HITFLAG_STACK_OVERFLOW; }
            else if (w.type == GLOME_TAG) { if (pstk_cons(wtg, w.b)) mflags |= GLOME_HITFLAG_STACK_OVERFLOW; }
            else if (w.type == GLOME_NOSHADOW || w.type == GLOME_INSTANCE) {}
            else break;
            wj = w.a; w = S.nodes[wj];
        }
    }
    a.tex = wtx; a.tag = wtg;
}

// inside s pt with the item record's short cuts
GD_FN bool gq_inside_fast(const DScene& S, int node, const Vec& pt, int* ovf) {
    const int4 it = gd_ldg(S.items + node);
    const int cls = it.x & 15;
    if (cls == GI_PRIM || cls == GI_INST_PRIM) {
        GlomeNode pn;
        pn.type = (it.x >> 4) & 15; pn.a = it.z; pn.b = 0; pn.c = 0;
        return prim_inside(S, pn, cls == GI_PRIM ? pt : invxfm_point(S.dpool + it.w, pt));
    }
    if (cls == GI_DEAD) return false;
    return gq_inside(S, node, pt, ovf);
}
GD_FN bool gq_inside_all_fast(const DScene& S, int first, int count, const Vec& pt, int* ovf) {  // Csg.hs:99-101
    for (int i = 0; i < count; i++)
        if (!gq_inside_fast(S, first + i, pt, ovf)) return false;
    return true;
}

// The machine's registers: everything a query keeps between two steps.  (A plain struct of scalars: after inlining the
// compiler keeps it in registers.)
struct QRegs {
    int sp;                 // control stack pointer (words)
    int nslots, acc;        // hit slots in use; the slot the current fold accumulates into
    Flt acc_t;              // cached slot[acc].t / .hit
    bool acc_hit;
    bool shadow_q;          // the query: shadow (true) or rayint
    bool smode;             // current evaluation: shadow (any hit) or rayint (closest hit)
    bool retb;              // a shadow result travelling down the stack
    bool cull;              // the current ray has unit length: best-hit culling and the origin clamp are safe (see qvm_start)
    int mflags;             // sticky overflow flags of this query
    int nadv;               // rayint_advance re-issues so far
    Ray r;
    Flt d;
    PStk ctex, ctag;
    int cur_bih;            // node of the BIH the lin_* registers belong to
    bool lin_ok;
    Flt drx, dry, drz;      // 1 / r.d: follows every change of r.d
    int lin_j0, lin_a0;     // linear sphere block of cur_bih (lin_a0 < 0: none)
    int ref;
    Flt near_, far_;
    int li, ln;             // list registers
    Flt ld;
    bool llin;              // the list is a leaf of a linear-sphere BIH
    int grp_orig;           // >= 0: cur_bih is the implicit BIH of a plain group (glome_tagmap.h): ipool offset of its list positions
    int grp_dup, grp_best;  // first copied item of that group; list position of the group's best hit so far (-1: none)
    int ni;
    int st;
};

// Start a query:  shadow_q = false: rayint root ray d [] []  -> result in vm.slot[0] when st == GS_DONE
//                 shadow_q = true : shadow root ray d        -> q.retb
GD_FN void qvm_start(QRegs& q, QVM& vm, int root, const Ray& qray, Flt qd, bool shadow_q) {
    q.sp = 0; q.nslots = 1; q.acc = 0; q.acc_t = GLM_INFINITY; q.acc_hit = false;
    q.shadow_q = shadow_q; q.smode = shadow_q; q.retb = false; q.mflags = 0; q.nadv = 0;
    q.r = qray; q.d = qd;
    // Best-hit culling and the origin clamp assume that a primitive's depth and a slab distance measure the
    // same thing.  rayint_sphere (Sphere.hs:20-41) takes |d| = 1 for granted, and one ray of the reference is not
    // normalised: Refract's refracted direction (Shader.hs:131-147).  For such a ray the walk follows the reference's own
    // schedule: both children whenever their intervals allow, no clamp.  Below an Instance the ray is normalised again.
    {
        const Flt dd = vdot(qray.d, qray.d);
        q.cull = (dd > 1 - 1e-12) && (dd < 1 + 1e-12);
    }
    q.ctex = pstk_empty(); q.ctag = pstk_empty();
    q.cur_bih = -1; q.lin_ok = false;
    q.drx = 1 / qray.d.x; q.dry = 1 / qray.d.y; q.drz = 1 / qray.d.z;
    q.lin_j0 = 0; q.lin_a0 = -1; q.ref = 0; q.near_ = 0; q.far_ = 0;
    q.li = 0; q.ln = 0; q.ld = 0; q.llin = false;
    q.grp_orig = -1; q.grp_dup = 0; q.grp_best = -1;
    q.ni = root; q.st = GS_ENTER;
    ghit_clear(vm.slot[0]);
    vm.cs[q.sp++] = gq_hdr(GF_ROOT, 0, 0);
}

// One round of the machine: BIH branch steps, then list items, then at most one ENTER and one RET.  The four parts are
// laid out one after the other so that the lanes of a warp that sit in the same state execute it together.
// PART 0: the whole round.  PART 1: BRANCH + LIST only, PART 2: ENTER + RET only -- the two halves a phase-locked block
// runs on either side of a barrier (k_gen_trace, GEN_PHASE_LOCK 2), so that the SM fetches one half's code at a time.
template <int PART>
GD_FN void qvm_step_part(const DScene& S, QRegs& q, QVM& vm, GCnt& cnt) {
    unsigned long long* cs = vm.cs;
    GHit* slot = vm.slot;
    int& sp = q.sp; int& nslots = q.nslots; int& acc = q.acc; Flt& acc_t = q.acc_t; bool& acc_hit = q.acc_hit;
    bool& smode = q.smode; bool& retb = q.retb; bool& cull = q.cull; int& mflags = q.mflags; int& nadv = q.nadv;
    Ray& r = q.r; Flt& d = q.d; PStk& ctex = q.ctex; PStk& ctag = q.ctag;
    int& cur_bih = q.cur_bih; bool& lin_ok = q.lin_ok; Flt& drx = q.drx; Flt& dry = q.dry; Flt& drz = q.drz;
    int& lin_j0 = q.lin_j0; int& lin_a0 = q.lin_a0; int& ref = q.ref; Flt& near_ = q.near_; Flt& far_ = q.far_;
    int& li = q.li; int& ln = q.ln; Flt& ld = q.ld; bool& llin = q.llin; int& ni = q.ni; int& st = q.st;
    int& grp_orig = q.grp_orig; int& grp_dup = q.grp_dup; int& grp_best = q.grp_best;

#define GQ_NEED(nw) if (sp + (nw) > GQ_WORDS) { mflags |= GLOME_HITFLAG_CSG_OVERFLOW; goto gq_abort; }
#define GQ_NEED_SLOT(k) if (nslots + (k) > GQ_SLOTS) { mflags |= GLOME_HITFLAG_CSG_OVERFLOW; goto gq_abort; }
#define GQ_SET_ACC(a_) do { acc = (a_); acc_t = slot[acc].t; acc_hit = slot[acc].hit != 0; } while (0)
#define GQ_LEAF()                                                              \
    do {                                                                       \
        int first_, count_;                                                    \
        glome_bih_leaf(ref, S.ipool, &first_, &count_);                        \
        li = first_; ln = first_ + count_;                                     \
        ld = smode ? fmin_(d, far_) : far_; /* Bih.hs:515 / :339 */            \
        llin = lin_a0 >= 0;                                                    \
        st = GS_LIST;                                                          \
    } while (0)


    if (PART != 2) {
#if GQ_ROUND_BUDGET > 0
    int budget_b = GQ_ROUND_BUDGET, budget_l = GQ_ROUND_BUDGET;  // steps a lane may take in one round (phase-locked blocks)
#define GQ_BUDGET(b) ((b)-- > 0)
#else
#define GQ_BUDGET(b) true
#endif
    // ---------------- BRANCH: BIH nodes (Bih.hs:340-366 / 516-542) ----------------
    while (st == GS_BRANCH && GQ_BUDGET(budget_b)) {
        cnt.bih++;
        const BihStep bs_ = ld_bih(S.bih, ref);
        const Flt2 sp2 = {bs_.ls, bs_.rs};
        int4 ii; ii.x = bs_.axis; ii.y = bs_.left; ii.z = bs_.right; ii.w = 0;
        const Flt dr_ = (ii.x == 0) ? drx : ((ii.x == 1) ? dry : drz);
        const Flt o = (ii.x == 0) ? r.o.x : ((ii.x == 1) ? r.o.y : r.o.z);
        const Flt dl = (sp2.x - o) * dr_;
        const Flt dr = (sp2.y - o) * dr_;
        const bool fwd = dr_ > 0;       // the near child is the left one iff dirr > 0
        const Flt dn = fwd ? dl : dr;
        const Flt df = fwd ? dr : dl;
        const int c1 = fwd ? ii.y : ii.z;
        const int c2 = fwd ? ii.z : ii.y;
        const bool v1 = near_ < dn;
        const Flt f1 = fmin_(dn, far_);
        bool v2 = df < far_;
        const Flt n2 = fmax_(df, near_);
        if (!smode && cull && v2 && acc_hit && n2 > acc_t) v2 = false;  // best-hit culling
        if (v1 && v2) {
            GQ_NEED(3);
            cs[sp] = gq_d2w(far_); cs[sp + 1] = gq_d2w(n2);
            cs[sp + 2] = (unsigned long long)GF_TRAV | ((unsigned long long)(unsigned int)c2 << 32);
            sp += 3;
        }
        if (v1) { ref = c1; far_ = f1; }
        else if (v2) { ref = c2; near_ = n2; }
        else { st = GS_RET; break; }
        if (ref < 0) GQ_LEAF();
    }
    // ---------------- LIST: the items of a group / BIH leaf (Solid.hs:326-331) ----------------
    while (st == GS_LIST && GQ_BUDGET(budget_l)) {
        if (li >= ln) { st = GS_RET; break; }
        const int item = li++;
        if (llin) {  // bare sphere of a linear block: no record to chase
            cnt.prim++;
            const Flt* sph = S.dpool + lin_a0 + 4 * (item - lin_j0);
            if (smode) {
                if (shadow_sphere(sph, r, ld)) { retb = true; st = GS_RET; }
            } else {
                Flt t; Vec pos, n;
                if (prim_sphere<true>(sph, r, ld, t, pos, n) && (!acc_hit || !(acc_t < t))) {
                    GHit& a = slot[acc];
                    a.hit = 1; a.t = t; a.pos = pos; a.norm = n; a.ray = r; a.tex = ctex; a.tag = ctag; a.prim = item; a.sub = -1;
                    acc_t = t; acc_hit = true;
                }
            }
            continue;
        }
        const int4 it = gd_ldg(S.items + item);
        const int cls = it.x & 15;
        if (cls == GI_DEAD || !(it.x & (smode ? GI_VIS_S : GI_VIS_R))) continue;  // Void; NoShadow / OnlyShadow / Mesh gates
        if (cls == GI_COMPLEX) {
            // remember where the list stands, evaluate the item with the list's distance
            GQ_NEED(3);
            cs[sp] = gq_d2w(d); cs[sp + 1] = gq_d2w(ld); cs[sp + 2] = gq_hdr(GF_LIST, li, ln);
            sp += 3;
            ni = item; d = ld; st = GS_ENTER;
            break;
        }
        SimpleHit sh_;
#ifdef GQ_LIST_CALL
        if (!gq_test_simple(S, it, r.o.x, r.o.y, r.o.z, r.d.x, r.d.y, r.d.z, drx, dry, drz, ld, smode, &sh_, &cnt)) continue;
#else
        if (!gq_test_simple_inl(S, it, r.o.x, r.o.y, r.o.z, r.d.x, r.d.y, r.d.z, drx, dry, drz, ld, smode, &sh_, &cnt)) continue;
#endif
        if (smode) { retb = true; st = GS_RET; continue; }
        if (acc_hit && acc_t < sh_.t) continue;
        if (grp_orig >= 0) {  // a plain group walked through its implicit BIH: of equal depths the LATER list element wins
            const int pos_ = S.ipool[grp_orig + (item - grp_dup)];
            if (acc_hit && acc_t == sh_.t && grp_best > pos_) continue;
            grp_best = pos_;
        }
        gq_fill_hit(S, slot[acc], it, item, sh_, ctex, ctag, mflags);
        acc_t = sh_.t; acc_hit = true;
    }
    }  // PART != 2
    if (PART == 1) return;
    // ---------------- ENTER: dispatch on a node ----------------
    if (st == GS_ENTER) {
        const int4 it = gd_ldg(S.items + ni);
        const int cls = it.x & 15;
        if (cls == GI_PRIM || cls == GI_INST_PRIM) {  // a one-element list: the primitive test lives in LIST only
            li = ni; ln = ni + 1; ld = d; llin = false; st = GS_LIST;
            return;
        }
        if (cls == GI_DEAD || !(it.x & (smode ? GI_VIS_S : GI_VIS_R))) { st = GS_RET; return; }  // Tex.hs:81,89; Mesh.hs:210
        if (!smode && (it.x & GI_WRAPPED)) {  // push the wrappers onto the context (Tex.hs:54,66); GF_CTX restores it
            GQ_NEED(5);
            cs[sp] = ctex.lo; cs[sp + 1] = ctex.hi; cs[sp + 2] = ctag.lo; cs[sp + 3] = ctag.hi; cs[sp + 4] = gq_hdr(GF_CTX, 0, 0);
            sp += 5;
            int wj = ni;
            while (wj != it.y) {
                const GlomeNode w = S.nodes[wj];
                if (w.type == GLOME_TEX) { if (pstk_cons(ctex, w.b)) mflags |= GLOME_HITFLAG_STACK_OVERFLOW; }
                else if (w.type == GLOME_TAG) { if (pstk_cons(ctag, w.b)) mflags |= GLOME_HITFLAG_STACK_OVERFLOW; }
                wj = w.a;
            }
        }
        ni = it.y;
        const GlomeNode nd = S.nodes[ni];
        switch (nd.type) {
            case GLOME_GROUP:  // Solid.hs:327-330
                if ((nd.c & GLOME_GROUP_ACCEL) && cull && r.d.x != 0 && r.d.y != 0 && r.d.z != 0 &&
                    fabs_(drx) < FL(1e30) && fabs_(dry) < FL(1e30) && fabs_(drz) < FL(1e30)) {
                    // a large plain group: walk its implicit BIH (glome_tagmap.h); same hits, list position as the tie key
                    const int ao = nd.c >> 4;
                    const Bbox bb = ldbb(S.dpool + S.ipool[ao + 1]);
                    Flt nr, fr;
                    bbclip_ub_pre(r, drx, dry, drz, bb, nr, fr);
                    fr = fmin_(d, fr);
                    if (nr < 0) nr = 0;
                    if (nr > fr) { st = GS_RET; break; }  // the ray misses the (padded) box of all elements
                    GQ_NEED(1);
                    cs[sp++] = gq_hdr(GF_BIH, cur_bih, 0);
                    cur_bih = ni;
                    lin_a0 = -1; lin_ok = true;
                    grp_orig = S.ipool[ao + 3]; grp_dup = S.ipool[ao + 2]; grp_best = -1;
                    ref = S.ipool[ao]; near_ = nr; far_ = fr;
                    if (ref < 0) GQ_LEAF();
                    else st = GS_BRANCH;
                    break;
                }
                li = nd.a; ln = nd.a + nd.b; ld = d; llin = false; st = GS_LIST;
                break;
            case GLOME_BIH: {  // Bih.hs:332-338, 368 / 510-515, 544
                const Bbox bb = ldbb(S.dpool + nd.b);
                Flt nr, fr;
                bbclip_ub_pre(r, drx, dry, drz, bb, nr, fr);
                fr = fmin_(d, fr);
                if (cull && nr < 0) nr = 0;  // origin clamp
                if (nd.a >= 0 && nr > fr) { st = GS_RET; break; }  // Bih.hs:347 at the root
                GQ_NEED(1);
                cs[sp++] = gq_hdr(GF_BIH, cur_bih, 0);
                cur_bih = ni;
                lin_a0 = -1;
                if (nd.c & GLOME_BIH_LINEAR_SPHERES) { lin_j0 = nd.c >> 4; lin_a0 = S.nodes[lin_j0].a; }
                lin_ok = true;
                ref = nd.a; near_ = nr; far_ = fr;
                if (ref < 0) GQ_LEAF();
                else st = GS_BRANCH;
                break;
            }
            case GLOME_MESH: {  // Mesh.hs:136-198; shadow = False (Mesh.hs:210)
                if (smode) { st = GS_RET; break; }
                Hit mh;
                hit_clear(mh);
                mh.hit = acc_hit ? 1 : 0; mh.t = acc_t;  // only a hit that beats the fold so far replaces it
                Stk tx, tg;
                pstk_to_stk(ctex, tx);
                pstk_to_stk(ctag, tg);
                Cnt mc = {0, 0, 0, 0};
                rayint_mesh(S, ni, nd, r, d, tx, tg, true, mh, &mc);
                cnt.bvh += mc.bvh; cnt.tri += mc.tri;
                mflags |= mh.flags;
                if (mh.sub >= 0) {
                    GHit& a = slot[acc];
                    a.hit = 1; a.t = mh.t; a.pos = mh.pos; a.norm = mh.norm; a.ray = r; a.prim = ni; a.sub = mh.sub;
                    if (pstk_from_stk(a.tex, mh.tex)) mflags |= GLOME_HITFLAG_STACK_OVERFLOW;
                    if (pstk_from_stk(a.tag, mh.tag)) mflags |= GLOME_HITFLAG_STACK_OVERFLOW;
                    acc_t = mh.t; acc_hit = true;
                }
                st = GS_RET;
                break;
            }
            case GLOME_INSTANCE: {  // Solid.hs:388-403 / 464-471
                cnt.inst++;
                const Flt* xfm = S.dpool + nd.b;
                const Vec newdir = invxfm_vec(xfm, r.d);
                const Vec neworig = invxfm_point(xfm, r.o);
                const Flt lenscale = vlen(newdir);
                const Flt invls = 1 / lenscale;
                GQ_NEED(13);
                cs[sp] = gq_d2w(r.o.x); cs[sp + 1] = gq_d2w(r.o.y); cs[sp + 2] = gq_d2w(r.o.z);
                cs[sp + 3] = gq_d2w(r.d.x); cs[sp + 4] = gq_d2w(r.d.y); cs[sp + 5] = gq_d2w(r.d.z);
                cs[sp + 6] = gq_d2w(d); cs[sp + 7] = gq_d2w(invls);
                cs[sp + 8] = gq_d2w(drx); cs[sp + 9] = gq_d2w(dry); cs[sp + 10] = gq_d2w(drz);
                cs[sp + 11] = cull ? 1ull : 0ull;
                cs[sp + 12] = gq_hdr(GF_INST, ni, acc);
                sp += 13;
                cull = true;  // the object-space ray is normalised (Solid.hs:392)
                if (!smode) {
                    GQ_NEED_SLOT(1);
                    ghit_clear(slot[nslots]);
                    GQ_SET_ACC(nslots);
                    nslots++;
                }
                r = mkray(neworig, vscale(newdir, invls));
                drx = 1 / r.d.x; dry = 1 / r.d.y; drz = 1 / r.d.z;
                d = d * lenscale;
                ni = nd.a;
                break;  // st stays GS_ENTER
            }
            case GLOME_DIFFERENCE:
            case GLOME_INTERSECTION: {
                if (smode) {  // no shadow method: the class default is a rayint with empty stacks (Solid.hs:218-221)
                    GQ_NEED(6);
                    GQ_NEED_SLOT(1);
                    cs[sp] = ctex.lo; cs[sp + 1] = ctex.hi; cs[sp + 2] = ctag.lo; cs[sp + 3] = ctag.hi; cs[sp + 4] = gq_hdr(GF_CTX, 0, 0);
                    cs[sp + 5] = gq_hdr(GF_MODE, nslots, acc);
                    sp += 6;
                    ctex = pstk_empty(); ctag = pstk_empty();
                    smode = false;
                    ghit_clear(slot[nslots]);
                    GQ_SET_ACC(nslots);
                    nslots++;
                }
                // result slot R; folded into `acc` when the node is done
                GQ_NEED_SLOT(1);
                GQ_NEED(7);
                const int R = nslots++;
                ghit_clear(slot[R]);
                cs[sp] = gq_d2w(r.o.x); cs[sp + 1] = gq_d2w(r.o.y); cs[sp + 2] = gq_d2w(r.o.z); cs[sp + 3] = gq_d2w(d);
                cs[sp + 4] = gq_hdr(0, R, acc);
                sp += 5;
                if (nd.type == GLOME_DIFFERENCE) { cs[sp] = gq_hdr(GF_DIFF, ni, 0); sp += 1; }
                else { cs[sp] = gq_hdr(GF_ISECT_BASE, ni, 0); cs[sp + 1] = gq_hdr(GF_ISECT, 0, 2); sp += 2; }
                st = GS_RET;  // the frame's first phase starts the node
                break;
            }
            case GLOME_BOUND: {  // Bound.hs:30-35 / 44-49
                int ovf = 0;
                const bool in = gq_inside(S, nd.a, r.o, &ovf);
                if (ovf) mflags |= GLOME_HITFLAG_STACK_OVERFLOW;
                if (in) { ni = nd.b; break; }
                GQ_NEED(1);
                cs[sp++] = gq_hdr(GF_GATE, nd.b, smode ? 1 : 0);
                smode = true;
                retb = false;
                ni = nd.a;
                break;
            }
            case GLOME_INNERBOUND: {
                if (smode) {  // Bound.hs:101-103
                    GQ_NEED(1);
                    cs[sp++] = gq_hdr(GF_OR, nd.b, 0);
                    ni = nd.a;
                    break;
                }
                // Bound.hs:98-99: the inner object (empty stacks) only supplies a depth limit for the outer one
                GQ_NEED(6);
                GQ_NEED_SLOT(1);
                cs[sp] = ctex.lo; cs[sp + 1] = ctex.hi; cs[sp + 2] = ctag.lo; cs[sp + 3] = ctag.hi; cs[sp + 4] = gq_d2w(d);
                cs[sp + 5] = gq_hdr(GF_INNER, nd.b, acc);
                sp += 6;
                ctex = pstk_empty(); ctag = pstk_empty();
                ghit_clear(slot[nslots]);
                GQ_SET_ACC(nslots);
                nslots++;
                ni = nd.a;
                break;
            }
            default: st = GS_RET; break;  // Void; OnlyShadow under rayint (Tex.hs:89); NoShadow under shadow (Tex.hs:81)
        }
        return;
    }
    // ---------------- RET: pop one continuation ----------------
    if (st == GS_RET) {
        const unsigned long long h = cs[--sp];
        switch (gq_op(h)) {
            case GF_ROOT:
                slot[0].flags |= mflags;
                st = GS_DONE;
                break;
            case GF_TRAV: {
                sp -= 2;
                if (smode && retb) break;  // an occluder was found: drop the pending subtrees
                const Flt nn = gq_w2d(cs[sp + 1]);
                if (!smode && cull && acc_hit && nn > acc_t) break;  // best-hit culling
                ref = (int)(unsigned int)(h >> 32);
                near_ = nn;
                far_ = gq_w2d(cs[sp]);
                if (!lin_ok) {  // a nested BIH used the registers: reload this BIH's
                    const GlomeNode bn = S.nodes[cur_bih];
                    lin_a0 = -1;
                    if (bn.c & GLOME_BIH_LINEAR_SPHERES) { lin_j0 = bn.c >> 4; lin_a0 = S.nodes[lin_j0].a; }
                    lin_ok = true;
                }
                if (ref < 0) GQ_LEAF();
                else st = GS_BRANCH;
                break;
            }
            case GF_BIH:
                cur_bih = gq_a(h);
                lin_ok = false;
                grp_orig = -1;  // (an implicit group BIH holds simple items only, so it is never the OUTER one)
                break;
            case GF_LIST:
                sp -= 2;
                d = gq_w2d(cs[sp]);
                if (smode && retb) break;
                li = gq_a(h); ln = gq_b(h); ld = gq_w2d(cs[sp + 1]);
                llin = false;  // a list holding a complex item is never a linear-sphere leaf
                st = GS_LIST;
                break;
            case GF_CTX:
                sp -= 4;
                ctex.lo = cs[sp]; ctex.hi = cs[sp + 1]; ctag.lo = cs[sp + 2]; ctag.hi = cs[sp + 3];
                break;
            case GF_INST: {
                sp -= 12;
                const int node = gq_a(h), parent = gq_b(h);
                r.o = vec(gq_w2d(cs[sp]), gq_w2d(cs[sp + 1]), gq_w2d(cs[sp + 2]));
                r.d = vec(gq_w2d(cs[sp + 3]), gq_w2d(cs[sp + 4]), gq_w2d(cs[sp + 5]));
                d = gq_w2d(cs[sp + 6]);
                drx = gq_w2d(cs[sp + 8]); dry = gq_w2d(cs[sp + 9]); drz = gq_w2d(cs[sp + 10]);
                cull = cs[sp + 11] != 0;
                if (smode) break;
                const Flt invls = gq_w2d(cs[sp + 7]);
                const GHit& c = slot[acc];
                GHit& p = slot[parent];
                if (c.hit) {
                    const Flt t = c.t * invls;
                    if (!p.hit || !(p.t < t)) {
                        const Flt* xfm = S.dpool + S.nodes[node].b;
                        p.hit = 1; p.t = t;
                        p.pos = xfm_point(xfm, c.pos);
                        p.norm = vnorm(invxfm_norm(xfm, c.norm));
                        p.ray = c.ray; p.tex = c.tex; p.tag = c.tag; p.prim = c.prim; p.sub = c.sub;
                    }
                }
                nslots--;
                GQ_SET_ACC(parent);
                break;
            }
            case GF_MODE: {  // the rayint that stood in for a shadow is done: its hit flag is the answer
                const int sl = gq_a(h);
                retb = slot[sl].hit != 0;
                nslots = sl;
                smode = true;
                GQ_SET_ACC(gq_b(h));
                break;
            }
            case GF_GATE: {
                const bool pass = retb;
                smode = gq_b(h) != 0;
                retb = false;
                if (pass) { ni = gq_a(h); st = GS_ENTER; }
                break;
            }
            case GF_OR:
                if (!retb) { ni = gq_a(h); st = GS_ENTER; }
                break;
            case GF_INNER: {
                sp -= 5;
                ctex.lo = cs[sp]; ctex.hi = cs[sp + 1]; ctag.lo = cs[sp + 2]; ctag.hi = cs[sp + 3];
                const Flt dsave = gq_w2d(cs[sp + 4]);
                const Flt dn = ghit_depth(slot[acc]);  // Bound.hs:99
                nslots--;
                GQ_SET_ACC(gq_b(h));
                cs[sp] = gq_d2w(dsave); cs[sp + 1] = gq_hdr(GF_SETD, 0, 0);  // (reuses the words just popped)
                sp += 2;
                d = dn;
                ni = gq_a(h);
                st = GS_ENTER;
                break;
            }
            case GF_SETD:
                sp -= 1;
                d = gq_w2d(cs[sp]);
                break;
            case GF_DIFF: {  // Csg.hs:33-54
                const int node = gq_a(h);
                int phase = gq_b(h);
                int bp = sp - 1;
                while (gq_op(cs[bp]) == GF_ADD) bp -= 2;
                const int R = gq_a(cs[bp]), parent = gq_b(cs[bp]);
                const GlomeNode nd = S.nodes[node];
                const int sa = nd.a, sb = nd.b;
                const int X = R + 1, Y = R + 2;  // sub-results live right above R
                int ovf = 0;
                // The phases run back to back while the operand to evaluate is a simple item (tested right here by
                // gq_test_simple); a complex operand is ENTERed with this frame re-pushed for the phase that follows.
                for (;;) {
                    bool finish = false, advance = false;
                    Flt adv = 0;
                    int want = -1, wslot = X, nphase = 0;  // operand to evaluate next, into which slot, continuing where
                    if (phase == 0) {
                        // inside sb (origin r)  ->  rib = rayint sb ... ; else ria = rayint sa ...
                        const bool inb = gq_inside_fast(S, sb, r.o, &ovf);
                        want = inb ? sb : sa; wslot = X; nphase = inb ? 1 : 2;
                    } else if (phase == 1) {  // origin inside sb; rib in X
                        const GHit& rib = slot[X];
                        if (!rib.hit) finish = true;
                        else if (gq_inside_fast(S, sa, rib.pos, &ovf) && !gq_inside_fast(S, sb, vscaleadd(rib.pos, r.d, GLM_DELTA), &ovf)) {
                            GHit& o = slot[R];
                            o = rib;
                            o.norm = vinvert(rib.norm);
                            if (nd.c != 0) {  // useatex: textures / tags come from get_metainfo sa bp ONLY
                                if (gd_ldg(S.items + sa).x & GI_HAS_META) {
                                    int fl = 0;
                                    gq_metainfo(S, sa, rib.pos, o.tex, o.tag, fl);
                                    mflags |= fl;
                                } else { o.tex = pstk_empty(); o.tag = pstk_empty(); }  // no Tex / Tag below sa: ([],[])
                            }
                            finish = true;
                        } else { advance = true; adv = rib.t; }
                    } else if (phase == 2) {  // origin outside sb; ria in X
                        if (!slot[X].hit) finish = true;
                        else { want = sb; wslot = Y; nphase = 3; }
                    } else {  // phase 3: ria in X, rib in Y
                        const GHit& ria = slot[X];
                        const GHit& rib = slot[Y];
                        if (rib.hit) {
                            if (ria.t < rib.t) { slot[R] = ria; finish = true; }
                            else { advance = true; adv = rib.t; }
                        } else { slot[R] = ria; finish = true; }
                    }
                    if (want >= 0) {
                        if (wslot + 1 > GQ_SLOTS) { mflags |= GLOME_HITFLAG_CSG_OVERFLOW; goto gq_abort; }
                        nslots = wslot + 1;
                        ghit_clear(slot[wslot]);
                        const int4 wit = gd_ldg(S.items + want);
                        const int wcls = wit.x & 15;
                        if (wcls != GI_COMPLEX) {  // rayint of a simple operand, no frame
                            if (wcls != GI_DEAD && (wit.x & GI_VIS_R)) {
                                SimpleHit sh_;
                                if (gq_test_simple(S, wit, r.o.x, r.o.y, r.o.z, r.d.x, r.d.y, r.d.z, drx, dry, drz, d, false, &sh_, &cnt))
                                    gq_fill_hit(S, slot[wslot], wit, want, sh_, ctex, ctag, mflags);
                            }
                            phase = nphase;
                            continue;
                        }
                        GQ_NEED(1);
                        GQ_SET_ACC(wslot);
                        cs[sp++] = gq_hdr(GF_DIFF, node, nphase);
                        ni = want; st = GS_ENTER;
                        break;
                    }
                    if (advance) {  // rayint_advance (Solid.hs:85-91): re-issue the whole node from adv + delta further on
                        const Flt a = adv + GLM_DELTA;
                        cnt.csg++;
                        if (++nadv > GQ_ADV_CAP) { mflags |= GLOME_HITFLAG_CSG_OVERFLOW; finish = true; }
                        else {
                            if (sp + 3 + 16 > GQ_WORDS) {
                                // no room for another pending offset: fold it into the one on top.  The sum's association
                                // changes (flagged), the geometry does not.
                                mflags |= GLOME_HITFLAG_CSG_OVERFLOW;
                                if (gq_op(cs[sp - 1]) == GF_ADD) cs[sp - 2] = gq_d2w(gq_w2d(cs[sp - 2]) + a);
                                else goto gq_abort;
                            } else { cs[sp] = gq_d2w(a); cs[sp + 1] = gq_hdr(GF_ADD, 0, 0); sp += 2; }
                            r = ray_move(r, a);
                            d = d - a;
                            nslots = R + 1;
                            phase = 0;
                            continue;
                        }
                    }
                    if (finish) {  // pending offsets innermost first, then the frame's fixed part
                        GHit& o = slot[R];
                        while (gq_op(cs[sp - 1]) == GF_ADD) {
                            if (o.hit) o.t = o.t + gq_w2d(cs[sp - 2]);
                            sp -= 2;
                        }
                        sp -= 5;
                        r.o = vec(gq_w2d(cs[sp]), gq_w2d(cs[sp + 1]), gq_w2d(cs[sp + 2]));
                        d = gq_w2d(cs[sp + 3]);
                        GHit& p = slot[parent];
                        if (o.hit && (!p.hit || !(p.t < o.t))) p = o;
                        nslots = R;
                        GQ_SET_ACC(parent);
                    }
                    break;
                }
                if (ovf) mflags |= GLOME_HITFLAG_STACK_OVERFLOW;
                break;
            }
            case GF_ISECT: {  // Csg.hs:68-90
                int k = gq_a(h);
                const int pb = gq_b(h);
                int phase = pb & 3, in = (pb >> 2) & 1;
                int bp = sp - 1;  // the base of this Intersection: below its ADD / ELSE_ADV entries
                for (;;) {
                    const int op = gq_op(cs[bp]);
                    if (op == GF_ISECT_BASE) break;
                    bp -= (op == GF_ADD) ? 2 : 6;
                }
                const int node = gq_a(cs[bp]);
                const int R = gq_a(cs[bp - 1]), parent = gq_b(cs[bp - 1]);
                const GlomeNode nd = S.nodes[node];
                const int first = nd.a, count = nd.b;
                const int X = R + 1;
                int ovf = 0;
                bool unwind = false;
                bool start = phase == 2;
                if (phase == 0) unwind = true;  // the last element wrote straight into R
                nslots = R + 1;
                // As in GF_DIFF: an element that is a simple item is tested right here and the loop goes on; a complex
                // element is ENTERed with a GF_ISECT frame for the phase that follows.
                for (;;) {
                    if (phase == 1) {          // rs = rayint s r d t tags, in X
                        phase = 3;
                        const GHit& rs = slot[X];
                        if (in) {
                            if (!rs.hit) { k++; start = true; }  // rayint_intersection ss r d
                            else {
                                // x = rayint_intersection ss r (ridepth rs); if x misses: advance past rs
                                GQ_NEED(6);
                                cs[sp] = gq_d2w(r.o.x); cs[sp + 1] = gq_d2w(r.o.y); cs[sp + 2] = gq_d2w(r.o.z); cs[sp + 3] = gq_d2w(d);
                                cs[sp + 4] = gq_d2w(rs.t); cs[sp + 5] = gq_hdr(GF_ELSE_ADV, k, 0);
                                sp += 6;
                                d = rs.t;
                                k++; start = true;
                            }
                        } else {
                            if (!rs.hit) unwind = true;
                            else if (gq_inside_all_fast(S, first + k + 1, count - k - 1, rs.pos, &ovf)) {
                                slot[R] = rs;
                                slot[R].ray = r;  // RayHit sd sp sn r vzero st stags (Csg.hs:88)
                                unwind = true;
                            } else {
                                const Flt a = rs.t + GLM_DELTA;
                                cnt.csg++;
                                if (++nadv > GQ_ADV_CAP) { mflags |= GLOME_HITFLAG_CSG_OVERFLOW; unwind = true; }
                                else {
                                    if (sp + 3 + 16 > GQ_WORDS) {
                                        mflags |= GLOME_HITFLAG_CSG_OVERFLOW;
                                        if (gq_op(cs[sp - 1]) == GF_ADD) cs[sp - 2] = gq_d2w(gq_w2d(cs[sp - 2]) + a);
                                        else goto gq_abort;
                                    } else { cs[sp] = gq_d2w(a); cs[sp + 1] = gq_hdr(GF_ADD, 0, 0); sp += 2; }
                                    r = ray_move(r, a);
                                    d = d - a;
                                    start = true;  // same k
                                }
                            }
                        }
                        nslots = R + 1;
                    }
                    if (start) {
                        start = false;
                        const int n = count - k;
                        if (n <= 0 || d < 0) unwind = true;
                        else {
                            const int el = first + k;
                            const int4 eit = gd_ldg(S.items + el);
                            const int ecls = eit.x & 15;
                            const bool last = n == 1;
                            const int tslot = last ? R : X;   // the last element's rayint is the result itself
                            bool inn = false;
                            if (!last) {
                                if (X + 1 > GQ_SLOTS) { mflags |= GLOME_HITFLAG_CSG_OVERFLOW; goto gq_abort; }
                                inn = gq_inside_fast(S, el, r.o, &ovf);
                                nslots = X + 1;
                            }
                            ghit_clear(slot[tslot]);
                            if (ecls != GI_COMPLEX) {
                                if (ecls != GI_DEAD && (eit.x & GI_VIS_R)) {
                                    SimpleHit sh_;
                                    if (gq_test_simple(S, eit, r.o.x, r.o.y, r.o.z, r.d.x, r.d.y, r.d.z, drx, dry, drz, d, false, &sh_, &cnt))
                                        gq_fill_hit(S, slot[tslot], eit, el, sh_, ctex, ctag, mflags);
                                }
                                if (last) unwind = true;
                                else { phase = 1; in = inn ? 1 : 0; }
                                continue;
                            }
                            GQ_NEED(1);
                            GQ_SET_ACC(tslot);
                            cs[sp++] = gq_hdr(GF_ISECT, k, last ? 0 : (1 | (inn ? 4 : 0)));
                            ni = el; st = GS_ENTER;
                            break;
                        }
                    }
                    if (!unwind) break;
                    // unwind this Intersection's continuations, top down
                    GHit& o = slot[R];
                    const int op = gq_op(cs[sp - 1]);
                    if (op == GF_ADD) {
                        if (o.hit) o.t = o.t + gq_w2d(cs[sp - 2]);
                        sp -= 2;
                        continue;
                    }
                    if (op == GF_ELSE_ADV) {
                        sp -= 6;
                        if (o.hit) continue;  // the rest did hit: that is the result
                        // rayint_advance (SolidItem (Intersection (s:ss))) r d t tags (ridepth rs)
                        const Flt a = gq_w2d(cs[sp + 4]) + GLM_DELTA;
                        cnt.csg++;
                        if (++nadv > GQ_ADV_CAP) { mflags |= GLOME_HITFLAG_CSG_OVERFLOW; continue; }
                        k = gq_a(cs[sp + 5]);
                        r.o = vscaleadd(vec(gq_w2d(cs[sp]), gq_w2d(cs[sp + 1]), gq_w2d(cs[sp + 2])), r.d, a);
                        d = gq_w2d(cs[sp + 3]) - a;
                        cs[sp] = gq_d2w(a); cs[sp + 1] = gq_hdr(GF_ADD, 0, 0);
                        sp += 2;
                        ghit_clear(o);
                        unwind = false; start = true;
                        continue;
                    }
                    // GF_ISECT_BASE: the node is done
                    sp -= 6;
                    r.o = vec(gq_w2d(cs[sp]), gq_w2d(cs[sp + 1]), gq_w2d(cs[sp + 2]));
                    d = gq_w2d(cs[sp + 3]);
                    GHit& p = slot[parent];
                    if (o.hit && (!p.hit || !(p.t < o.t))) p = o;
                    nslots = R;
                    GQ_SET_ACC(parent);
                    break;
                }
                if (ovf) mflags |= GLOME_HITFLAG_STACK_OVERFLOW;
                break;
            }
            default: mflags |= GLOME_HITFLAG_CSG_OVERFLOW; goto gq_abort;  // corrupt stack: cannot happen
        }
        return;
    }
    return;

gq_abort:
    // the control or slot stack is exhausted: a flagged miss rather than a wrong hit
    ghit_clear(slot[0]);
    slot[0].flags = mflags;
    retb = false;
    st = GS_DONE;
#undef GQ_NEED
#undef GQ_NEED_SLOT
#undef GQ_SET_ACC
#undef GQ_LEAF
#undef GQ_BUDGET
}

GD_FN void qvm_step(const DScene& S, QRegs& q, QVM& vm, GCnt& cnt) { qvm_step_part<0>(S, q, vm, cnt); }

// Evaluate one query to completion (batch kernels, the pick query, debug counts).
GD_NOINLINE bool gq_query(const DScene& S, QVM& vm, int root, const Ray& qray, Flt qd, bool shadow_q, GCnt& cnt) {
    QRegs q;
    qvm_start(q, vm, root, qray, qd, shadow_q);
    while (q.st != GS_DONE) qvm_step(S, q, vm, cnt);
    return shadow_q ? q.retb : (vm.slot[0].hit != 0);
}

// ---------------------------------------------------------------------------------------------
// rayint_debug's Int (Solid.hs:155; Bih.hs:378-412; Bound.hs:37-42; Tex.hs:55,67,79,90): the number of BIH boxes the
// reference's own (unculled, unclamped) walk enters.  A sum over the graph, so the order of the visits is free:
// a work list instead of the reference's recursion.
// ---------------------------------------------------------------------------------------------
#define GD_WORK 64
#define GD_RAYS 8
GD_NOINLINE int gq_debug_count(const DScene& S, QVM& vm, int root, const Ray& ray0, Flt d0, GCnt& cnt, int* ovf) {
    int wk[GD_WORK], wa[GD_WORK], wr[GD_WORK];
    Flt wd[GD_WORK];
    Ray rays[GD_RAYS];
    TravEnt trav[64];
    int sp = 0, nr = 1, total = 0;
    rays[0] = ray0;
#define GD_PUSH(K, A, R, D) do { if (sp >= GD_WORK) { *ovf = 1; } else { wk[sp] = (K); wa[sp] = (A); wr[sp] = (R); wd[sp] = (D); sp++; } } while (0)
    GD_PUSH(0, root, 0, d0);
    while (sp > 0) {
        sp--;
        if (wk[sp] == 1) { nr = wa[sp]; continue; }  // release the ray slots of a finished Instance
        int n = wa[sp];
        const int ri = wr[sp];
        const Flt d = wd[sp];
        const Ray r = rays[ri];
        GlomeNode nd = S.nodes[n];
        while (is_wrap_r(nd.type)) { n = nd.a; nd = S.nodes[n]; }  // OnlyShadow counts 0 (Tex.hs:90)
        switch (nd.type) {
            case GLOME_GROUP:  // Solid.hs:329: sum
                for (int i = 0; i < nd.b; i++) {
                    GlomeNode c = S.nodes[nd.a + i];
                    while (is_wrap_r(c.type)) c = S.nodes[c.a];
                    if (is_prim(c.type) || c.type == GLOME_VOID || c.type == GLOME_ONLYSHADOW || c.type == GLOME_MESH) continue;  // 0
                    GD_PUSH(0, nd.a + i, ri, d);
                }
                break;
            case GLOME_INSTANCE: {  // Solid.hs:447-461
                GlomeNode c = S.nodes[nd.a];
                while (is_wrap_r(c.type)) c = S.nodes[c.a];
                if (is_prim(c.type) || c.type == GLOME_VOID || c.type == GLOME_ONLYSHADOW || c.type == GLOME_MESH) break;
                if (nr >= GD_RAYS) { *ovf = 1; break; }
                const Flt* xfm = S.dpool + nd.b;
                const Vec newdir = invxfm_vec(xfm, r.d);
                const Vec neworig = invxfm_point(xfm, r.o);
                const Flt lenscale = vlen(newdir);
                const Flt invlenscale = 1 / lenscale;
                GD_PUSH(1, nr, 0, 0);
                rays[nr] = mkray(neworig, vscale(newdir, invlenscale));
                GD_PUSH(0, nd.a, nr, d * lenscale);
                nr++;
                break;
            }
            case GLOME_BOUND: {  // Bound.hs:37-42
                int o2 = 0;
                if (gq_inside(S, nd.a, r.o, &o2) || gq_query(S, vm, nd.a, r, d, true, cnt)) {
                    total += 1;
                    GD_PUSH(0, nd.b, ri, d);
                }
                if (o2) *ovf = 1;
                break;
            }
            case GLOME_INNERBOUND: GD_PUSH(0, nd.b, ri, d); break;  // Bound.hs:107
            case GLOME_BIH: {  // Bih.hs:378-412: no clip by d at the root, no origin clamp, no culling
                const Bbox bb = ldbb(S.dpool + nd.b);
                Flt near_, far_;
                bbclip_ub(r, bb, near_, far_);
                const Flt drx = 1 / r.d.x, dry = 1 / r.d.y, drz = 1 / r.d.z;
                const bool linear = (nd.c & GLOME_BIH_LINEAR_SPHERES) != 0;
                int tsp = 0, ref = nd.a;
                for (;;) {
                    bool pop = false;
                    if (ref < 0) {
                        if (!linear) {  // a bare sphere counts 0
                            int lf, lc;
                            glome_bih_leaf(ref, S.ipool, &lf, &lc);
                            for (int i = 0; i < lc; i++) {
                                GlomeNode c = S.nodes[lf + i];
                                while (is_wrap_r(c.type)) c = S.nodes[c.a];
                                if (is_prim(c.type) || c.type == GLOME_VOID || c.type == GLOME_ONLYSHADOW || c.type == GLOME_MESH) continue;
                                GD_PUSH(0, lf + i, ri, fmin_(d, far_));
                            }
                        }
                        pop = true;
                    } else {
                        total++;
                        if (near_ > far_) pop = true;  // (RayMiss,0) wrapped with 1: only possible at the root
                        else {
                            const BihStep bs_ = ld_bih(S.bih, ref);
                            struct { Flt lsplit, rsplit; int axis, left, right; } bn = {bs_.ls, bs_.rs, bs_.axis, bs_.left, bs_.right};
                            const Flt dr_ = (bn.axis == 0) ? drx : ((bn.axis == 1) ? dry : drz);
                            const Flt o = (bn.axis == 0) ? r.o.x : ((bn.axis == 1) ? r.o.y : r.o.z);
                            const Flt dl = (bn.lsplit - o) * dr_;
                            const Flt dr = (bn.rsplit - o) * dr_;
                            const bool fwd = dr_ > 0;
                            const Flt dn = fwd ? dl : dr, df = fwd ? dr : dl;
                            const int c1 = fwd ? bn.left : bn.right, c2 = fwd ? bn.right : bn.left;
                            const bool v1 = near_ < dn, v2 = df < far_;
                            if (v1 && v2) {
                                if (tsp < 64) { trav[tsp].ref = c2; trav[tsp].near_ = fmax_(df, near_); trav[tsp].far_ = far_; tsp++; }
                                else *ovf = 1;
                            }
                            if (v1) { ref = c1; far_ = fmin_(dn, far_); }
                            else if (v2) { ref = c2; near_ = fmax_(df, near_); }
                            else pop = true;
                        }
                    }
                    if (pop) {
                        if (tsp == 0) break;
                        tsp--;
                        ref = trav[tsp].ref; near_ = trav[tsp].near_; far_ = trav[tsp].far_;
                    }
                }
                break;
            }
            default: break;  // the class default: 0 (Solid.hs:205)
        }
    }
#undef GD_PUSH
    return total;
}

// ---------------------------------------------------------------------------------------------
// SHM: trace (Trace.hs:59-82) + materialShader (Shader.hs:65-189) as a machine over trace frames (one per
// generation of secondary rays) and material continuations.
// ---------------------------------------------------------------------------------------------
#define GS_MAX_RECURS 8                  /* `recurs` accepted by the general tracer */
#define GS_TFRAMES (GS_MAX_RECURS + 1)
#define GS_MFRAMES 24
#define GS_TAGCAP 512 /* the pick query only: one thread, so the arena can be generous */

struct TFrame {
    Ray ray;
    Flt dlimit;
    GHit ri;
    ColorA colora;
    LightSel L;
    int tex_i, sld, ls, recurs;
    int tagbase, nseg;      // TAGS only: arena position at the start of the frame; segment ends of the textures shaded so far
    int segend[GLOME_MAX_STACK];
};
enum { MO_TEXFOLD = 0, MO_BLEND, MO_ADD, MO_REFL, MO_REFR, MO_WARP };
struct MFrame {
    int op, a, phase, n;
    Flt w;
    ColorA c;
    int t0, t1;             // TAGS only (Warp): arena marks
};
struct SHM {
    QVM q;
    TFrame tf[GS_TFRAMES];
    MFrame mf[GS_MFRAMES];
};
struct TagArena { int n, overflow; int v[GS_TAGCAP]; };

enum { SS_T_BEGIN = 0, SS_T_HIT, SS_T_TEX, SS_M_EVAL, SS_LIGHTS, SS_LIGHT_RET, SS_M_RET, SS_T_RETURN, SS_FINISHED };

// the shading machine's registers
struct SRegs {
    int L, msp, st, fl;
    int m_id;        // material being evaluated (-1: a texture-computed Blend)
    int li;          // mpreshade's light loop
    ColorA rv_c;     // value returned by the last material / trace
    Flt rv_d;
};

// trace lights shader sld ray depth recurs: set up frame 0.  Then call shm_step until it returns false; whenever it
// returns true a query has been started in `q` and must be run to GS_DONE (qvm_step) before the next call.
GD_FN void shm_start(SRegs& s, SHM& sh, int lightset, int sld, const Ray& ray, Flt depth, int recurs, TagArena* ta) {
    s.L = 0; s.msp = 0; s.st = SS_T_BEGIN; s.fl = 0; s.m_id = -1; s.li = 0;
    s.rv_c = mkca(0, 0, 0, 0); s.rv_d = GLM_INFINITY;
    if (recurs > GS_MAX_RECURS) { recurs = GS_MAX_RECURS; s.fl |= GLOME_HITFLAG_CSG_OVERFLOW; }  // (the C-ABI rejects it earlier)
    TFrame& T = sh.tf[0];
    T.ray = ray; T.dlimit = depth; T.sld = sld; T.ls = lightset; T.recurs = recurs;
    if (ta) { ta->n = 0; ta->overflow = 0; }
}

template <bool TAGS>
GD_FN bool shm_step(const DScene& S, SRegs& s, SHM& sh, QRegs& q, GCnt& cnt, TagArena* ta) {
    TFrame* tf = sh.tf;
    MFrame* mf = sh.mf;
    int& L = s.L; int& msp = s.msp; int& st = s.st; int& fl = s.fl; int& m_id = s.m_id;
    ColorA& rv_c = s.rv_c; Flt& rv_d = s.rv_d;
    MatVal m;
    m.kind = -1; m.a = m.b = m.c = m.d = 0;

#define GS_CALL_TRACE(RAY, DLIM, REC, LS, SLD)                                                    \
    do {                                                                                          \
        TFrame& N_ = tf[L + 1];                                                                   \
        N_.ray = (RAY); N_.dlimit = (DLIM); N_.recurs = (REC); N_.ls = (LS); N_.sld = (SLD);      \
        L++; st = SS_T_BEGIN;                                                                     \
    } while (0)
#define GS_MPUSH() if (msp >= GS_MFRAMES) { fl |= GLOME_HITFLAG_STACK_OVERFLOW; rv_c = mkca(0, 0, 0, 0); st = SS_M_RET; break; }

    for (;;) {
        TFrame& T = tf[L];
        if (st == SS_T_BEGIN) {
            if (TAGS) { T.tagbase = ta->n; T.nseg = 0; }
            if (T.recurs == 0) {  // Trace.hs:60
                ghit_clear(T.ri);
                rv_c = mkca(0, 0, 0, 0); rv_d = GLM_INFINITY;
                st = SS_T_RETURN;
                continue;
            }
            qvm_start(q, sh.q, T.sld, T.ray, T.dlimit, false);
            st = SS_T_HIT;
            return true;
        }
        if (st == SS_T_HIT) {  // rayint sld ray depth [] [] is in slot 0
            T.ri = sh.q.slot[0];
            fl |= T.ri.flags;
            if (!T.ri.hit) {  // mmissshade (Shader.hs:186)
                rv_c = mkca(0, 0, 0, 0); rv_d = GLM_INFINITY;
                st = SS_T_RETURN;
                continue;
            }
            T.colora = mkca(0, 0, 0, 0);
            T.L.done = 0; T.L.first = 0; T.L.count = 0; T.L.mask = 0;
            T.tex_i = 0;
            st = SS_T_TEX;
            continue;
        }
        if (st == SS_T_TEX) {  // the fold over the hit's textures (Trace.hs:66-80)
            const int tex = T.tex_i < GLOME_MAX_STACK ? pstk_get(T.ri.tex, T.tex_i) : -1;
            if (tex < 0) {
                if (TAGS) {
                    // ts = tagsb_k ++ ... ++ tagsb_1 (Trace.hs:79): reverse the order of the segments gathered so far,
                    // then `ts ++ tags` (Trace.hs:82)
                    if (T.nseg > 1) {
                        int tmp[GS_TAGCAP];
                        const int b0 = T.tagbase, e0 = ta->n;
                        for (int i = b0; i < e0; i++) tmp[i] = ta->v[i];
                        int w = b0;
                        for (int sgi = T.nseg - 1; sgi >= 0; sgi--) {
                            const int sb = sgi == 0 ? b0 : T.segend[sgi - 1], se = T.segend[sgi];
                            for (int i = sb; i < se; i++) ta->v[w++] = tmp[i];
                        }
                    }
                    for (int i = 0; i < GLOME_MAX_STACK; i++) {
                        const int tg = pstk_get(T.ri.tag, i);
                        if (tg < 0) break;
                        if (ta->n < GS_TAGCAP) ta->v[ta->n++] = tg; else ta->overflow = 1;
                    }
                }
                rv_c = T.colora; rv_d = T.ri.t;
                st = SS_T_RETURN;
                continue;
            }
            if (T.colora.a + GLM_DELTA >= 1) {  // opaque (Trace.hs:50)
                if (TAGS && T.nseg < GLOME_MAX_STACK) T.segend[T.nseg++] = ta->n;
                T.tex_i++;
                continue;
            }
            const GlomeTexture* tx = S.textures + tex;
            m_id = tx->kind == GLOME_TEX_UNIFORM ? tx->a : -1;
            eval_texture(S, tex, T.ri.pos, m, cnt.perlin);
            if (msp >= GS_MFRAMES) { fl |= GLOME_HITFLAG_STACK_OVERFLOW; T.tex_i = GLOME_MAX_STACK; continue; }
            mf[msp].op = MO_TEXFOLD; msp++;
            st = SS_M_EVAL;
            continue;
        }
        if (st == SS_LIGHTS) {  // mpreshade (Shader.hs:65-80), one light per round; a shadow query may interrupt it
            const Vec p = T.ri.pos, n = T.ri.norm;
            bool asked = false;
            while (s.li < T.L.count) {
                Ray sr; Flt sd; bool ns = false;
                if (!light_probe(S.lights + T.L.first + s.li, p, n, sr, sd, ns)) { s.li++; continue; }
                if (ns) {
                    cnt.shadow++;
                    qvm_start(q, sh.q, T.sld, sr, sd, true);
                    st = SS_LIGHT_RET;
                    asked = true;
                    break;
                }
                T.L.mask |= 1ull << s.li;
                s.li++;
            }
            if (asked) return true;
            T.L.done = 1;
            mat_load(S, m_id, m);  // the Surface that forced ctxb
            st = SS_M_EVAL;
            continue;
        }
        if (st == SS_LIGHT_RET) {
            fl |= sh.q.slot[0].flags;
            if (!q.retb) T.L.mask |= 1ull << s.li;
            s.li++;
            st = SS_LIGHTS;
            continue;
        }
        if (st == SS_M_EVAL) {  // mpostshade (Shader.hs:82-184) of material m at the hit of frame L
            const Vec dir = T.ray.d, n = T.ri.norm, p = T.ri.pos;
            const Vec eyedir = vinvert(dir);
            switch (m.kind) {
                case GLOME_MAT_SURFACE: {
                    if (!T.L.done) {  // ctxb is forced by the first Surface that is shaded (Trace.hs:63, Shader.hs:92)
                        T.L.first = S.lightsets[2 * T.ls];
                        T.L.count = S.lightsets[2 * T.ls + 1];
                        T.L.mask = 0;
                        s.li = 0;
                        st = SS_LIGHTS;
                        break;
                    }
                    shade_surface(S, T.L, p, m.p, n, eyedir, rv_c);
                    st = SS_M_RET;
                    break;
                }
                case GLOME_MAT_BLEND: {  // Shader.hs:181-184
                    GS_MPUSH();
                    MFrame& f = mf[msp++];
                    f.op = MO_BLEND; f.a = m.b; f.phase = 0; f.w = m.p[0];
                    m_id = m.a;
                    mat_load(S, m_id, m);
                    break;
                }
                case GLOME_MAT_ADDITIVE: {  // Shader.hs:177-179
                    if (m.b <= 0) { rv_c = mkca(0, 0, 0, 0); st = SS_M_RET; break; }  // casum [] (Clr.hs:93)
                    GS_MPUSH();
                    MFrame& f = mf[msp++];
                    f.op = MO_ADD; f.a = m.a; f.phase = 0; f.n = m.b; f.c = mkca(0, 0, 0, 1);
                    m_id = S.ipool[m.a];
                    mat_load(S, m_id, m);
                    break;
                }
                case GLOME_MAT_REFLECT: {  // Shader.hs:107-118
                    const Flt refl = m.p[0];
                    if ((refl > 0) && (T.recurs > 0)) {
                        GS_MPUSH();
                        MFrame& f = mf[msp++];
                        f.op = MO_REFL; f.w = refl;
                        const Vec outdir = reflect(dir, n);
                        cnt.secondary++;
                        GS_CALL_TRACE(mkray(vscaleadd(p, outdir, GLM_DELTA), outdir), (Flt)GLM_INFINITY, T.recurs - 1, T.ls, T.sld);
                    } else { rv_c = mkca(0, 0, 0, 1); st = SS_M_RET; }
                    break;
                }
                case GLOME_MAT_REFRACT: {  // Shader.hs:120-155
                    const Flt refl = m.p[0], refr = m.p[1];
                    if ((refl > 0 || refr > 0) && (T.recurs > 0)) {
                        GS_MPUSH();
                        MFrame& f = mf[msp++];
                        f.op = MO_REFR; f.a = m_id; f.phase = 0;
                        const Vec outdir = reflect(dir, n);
                        cnt.secondary++;
                        GS_CALL_TRACE(mkray(vscaleadd(p, outdir, GLM_DELTA), outdir), (Flt)GLM_INFINITY, T.recurs - 1, T.ls, T.sld);
                    } else { rv_c = mkca(0, 0, 0, 0); st = SS_M_RET; }
                    break;
                }
                case GLOME_MAT_WARP: {  // Shader.hs:157-175
                    GS_MPUSH();
                    MFrame& f = mf[msp++];
                    f.op = MO_WARP; f.a = m_id; f.phase = 0;
                    if (TAGS) f.t0 = ta->n;
                    cnt.secondary += 2;
                    GS_CALL_TRACE(T.ri.ray, (Flt)GLM_INFINITY, T.recurs - 1, T.ls, m.a);
                    break;
                }
                default: rv_c = mkca(0, 0, 0, 0); st = SS_M_RET; break;
            }
            continue;
        }
        if (st == SS_M_RET) {  // a material value (rv_c) or a nested trace (rv_c, rv_d) returns to the frame on top
            MFrame& f = mf[msp - 1];
            switch (f.op) {
                case MO_TEXFOLD:
                    msp--;
                    T.colora = cafold(T.colora, rv_c);
                    if (TAGS && T.nseg < GLOME_MAX_STACK) T.segend[T.nseg++] = ta->n;
                    T.tex_i++;
                    st = SS_T_TEX;
                    break;
                case MO_BLEND:
                    if (f.phase == 0) {
                        f.c = rv_c; f.phase = 1;
                        m_id = f.a;
                        mat_load(S, m_id, m);
                        st = SS_M_EVAL;
                    } else {
                        rv_c = caweight(f.c, rv_c, f.w);
                        msp--;
                    }
                    break;
                case MO_ADD:  // casum (Clr.hs:93-103)
                    f.c.r = f.c.r + rv_c.r * rv_c.a; f.c.g = f.c.g + rv_c.g * rv_c.a; f.c.b = f.c.b + rv_c.b * rv_c.a;
                    f.c.a = f.c.a * (1 - aclamp(rv_c.a));
                    f.phase++;
                    if (f.phase < f.n) {
                        m_id = S.ipool[f.a + f.phase];
                        mat_load(S, m_id, m);
                        st = SS_M_EVAL;
                    } else {
                        rv_c = mkca(f.c.r, f.c.g, f.c.b, 1 - f.c.a);
                        msp--;
                    }
                    break;
                case MO_REFL:
                    rv_c = mkca(rv_c.r, rv_c.g, rv_c.b, rv_c.a * f.w);
                    msp--;
                    break;
                case MO_REFR: {
                    const GlomeMaterial* g = S.materials + f.a;
                    const Flt refl = g->p[0], refr = g->p[1], ior = g->p[2];
                    if (f.phase == 0) {
                        f.c = rv_c;
                        const Vec dir = T.ray.d, n = T.ri.norm;
                        const Vec eyedir = vinvert(dir);
                        const Flt eta = (vdot(n, eyedir) > 0) ? ior : 1 / ior;
                        const Flt c1 = vdot(dir, n);
                        const Flt cs2 = 1 - (eta * eta) * (1 - (c1 * c1));
                        if (cs2 < 0) {
                            const ColorA a = f.c, b = mkca(0, 0, 0, 1);
                            rv_c = mkca(a.r * refl + b.r * refr, a.g * refl + b.g * refr, a.b * refl + b.b * refr, a.a * refl + b.a * refr);
                            msp--;
                        } else {
                            f.phase = 1;
                            const Vec t = vadd(vscale(dir, eta), vscale(n, eta * c1 - sqrt(cs2)));
                            cnt.secondary++;
                            GS_CALL_TRACE(mkray(vscaleadd(T.ri.pos, t, GLM_DELTA), t), (Flt)GLM_INFINITY, T.recurs - 1, T.ls, T.sld);
                        }
                    } else {
                        const ColorA a = f.c, b = rv_c;
                        rv_c = mkca(a.r * refl + b.r * refr, a.g * refl + b.g * refr, a.b * refl + b.b * refr, a.a * refl + b.a * refr);
                        msp--;
                    }
                    break;
                }
                case MO_WARP: {
                    const GlomeMaterial* g = S.materials + f.a;
                    if (f.phase == 0) {
                        f.c = rv_c; f.w = rv_d; f.phase = 1;
                        if (TAGS) f.t1 = ta->n;
                        const Ray wr = xfm_ray(S.dpool + g->d, mkray(T.ri.pos, vnorm(T.ray.d)));  // TestScene.hs:169-173
                        GS_CALL_TRACE(wr, rv_d, T.recurs - 1, g->c, g->b);
                    } else {
                        if (f.w < rv_d) {  // the frame is nearer (Shader.hs:172-174)
                            rv_c = f.c;
                            if (TAGS) ta->n = f.t1;
                        } else if (TAGS) {
                            int w = f.t0;
                            for (int i = f.t1; i < ta->n; i++) ta->v[w++] = ta->v[i];
                            ta->n = w;
                        }
                        msp--;
                    }
                    break;
                }
            }
            continue;
        }
        // SS_T_RETURN: the trace of frame L is (rv_c, rv_d)
        if (L == 0) { st = SS_FINISHED; return false; }
        L--;
        st = SS_M_RET;
    }
#undef GS_CALL_TRACE
#undef GS_MPUSH
}

// trace to completion (batch kernels, the pick query, the CPU harness): colour, the TraceResult's tag list when TAGS;
// the primary Rayint is sh.tf[0].ri
template <bool TAGS>
GD_FN void gs_trace(const DScene& S, SHM& sh, int lightset, int sld, const Ray& ray, Flt depth, int recurs, ColorA& outc,
                    GCnt& cnt, int& flags_out, TagArena* ta) {
    SRegs s;
    QRegs q;
    shm_start(s, sh, lightset, sld, ray, depth, recurs, ta);
    while (shm_step<TAGS>(S, s, sh, q, cnt, ta)) {
        while (q.st != GS_DONE) qvm_step(S, q, sh.q, cnt);
    }
    outc = s.rv_c;
    flags_out = s.fl;
}

// GHit -> the C-ABI's GlomeHit; tag ids go back through the scene's tag table
GD_FN void ghit_out(const DScene& S, const GHit& h, GlomeHit* o) {
    o->t = ghit_depth(h);
    o->hit = h.hit;
    o->prim = h.hit ? h.prim : -1;
    o->sub = h.hit ? h.sub : -1;
    o->flags = h.flags;
    int nt = 0, ng = 0;
    for (int i = 0; i < 3; i++) { o->pos[i] = 0; o->norm[i] = 0; }
    if (h.hit) {
        o->pos[0] = h.pos.x; o->pos[1] = h.pos.y; o->pos[2] = h.pos.z;
        o->norm[0] = h.norm.x; o->norm[1] = h.norm.y; o->norm[2] = h.norm.z;
    }
    for (int i = 0; i < GLOME_MAX_STACK; i++) {
        const int tx = h.hit ? pstk_get(h.tex, i) : -1, tg = h.hit ? pstk_get(h.tag, i) : -1;
        o->tex[i] = tx >= 0 ? tx : 0;
        o->tag[i] = tg >= 0 ? (S.tagvals ? S.tagvals[tg] : tg) : 0;
        nt += tx >= 0; ng += tg >= 0;
    }
    o->ntex = nt; o->ntag = ng;
}

}  // namespace ggen
