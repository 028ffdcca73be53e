/*
 * glome_cuda.h -- C-ABI of libglomecuda.so, the B200 (sm_100a) backend for GlomeTrace's
 * ray-cast hot path.
 *
 * The reference (jimsnow/glome, Haskell) has no FFI of its own; the entry points below are what a
 * `GlomeTrace.CUDA` module binds with `foreign import ccall safe` (haskell/Data/Glome/CUDA.hs,
 * INTEGRATION.md).  Each one names the reference interface it stands in for.
 *
 * Conventions: every function returns 0 on success and a negative GLOME_E* code on failure;
 * glome_last_error() gives a thread-local message.  No exceptions cross the boundary.  All buffers
 * are caller-owned host memory unless a name ends in _dev.  A handle may be used by one host thread
 * at a time.  Flt = double everywhere (GlomeVec/Data/Glome/Vec.hs:9).
 */
#ifndef GLOME_CUDA_H
#define GLOME_CUDA_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define GLOME_FLAT_VERSION 1

/* error codes */
#define GLOME_OK 0
#define GLOME_EINVAL (-1)   /* bad argument / malformed FlatScene */
#define GLOME_ECUDA (-2)    /* CUDA runtime error (message has the cudaError string) */
#define GLOME_ENODEV (-3)   /* no CUDA device: the product path never falls back to the CPU */
#define GLOME_ELIMIT (-4)   /* scene exceeds a compiled limit (stack depth ...) */
#define GLOME_EBUILD (-5)   /* scene construction error, mirrors a Haskell `error` call */

/* ---------------------------------------------------------------------------------------------
 * FlatScene: the one contract between scene construction (Haskell `flatten`, or the C++ host
 * mirror in glome_b200/csrc/host) and the device.  All indices are 0-based int32.
 * ------------------------------------------------------------------------------------------- */

/* Scene-graph node: the existential SolidItem (Solid.hs:261) as a tag + payload indices. */
typedef struct GlomeNode {
    int32_t type;
    int32_t a, b, c;
} GlomeNode;

enum GlomeNodeType {
    GLOME_VOID = 0,         /* Solid.hs:349                                                     */
    GLOME_SPHERE = 1,       /* Sphere.hs:11      a = dpool offset of {cx,cy,cz,r}               */
    GLOME_TRIANGLE = 2,     /* Triangle.hs:13    a = dpool offset of {p1,p2,p3} (9)             */
    GLOME_TRIANGLENORM = 3, /* Triangle.hs:14    a = dpool offset of {p1,p2,p3,n1,n2,n3} (18)   */
    GLOME_BOX = 4,          /* Box.hs:7          a = dpool offset of {p1,p2} (6)                */
    GLOME_PLANE = 5,        /* Plane.hs:11       a = dpool offset of {nx,ny,nz,offset}          */
    GLOME_DISC = 6,         /* Cone.hs:21        a = dpool offset of {pos,norm,r*r} (7)         */
    GLOME_CYLINDER = 7,     /* Cone.hs:22        a = dpool offset of {r,h1,h2}                  */
    GLOME_CONE = 8,         /* Cone.hs:23        a = dpool offset of {r,clip1,clip2,height}     */
    GLOME_GROUP = 9,        /* Solid.hs:326      a = first child node (children contiguous), b = count */
    GLOME_INSTANCE = 10,    /* Solid.hs:386      a = child node, b = dpool offset of Xfm (fwd 12, inv 12) */
    GLOME_BIH = 11,         /* Bih.hs:51         a = root ref, b = dpool offset of bbox (6), c = flags */
    GLOME_MESH = 12,        /* Mesh.hs:42        a = ipool offset of GlomeMeshHeader            */
    GLOME_DIFFERENCE = 13,  /* Csg.hs:14         a = sa, b = sb, c = useatex (always 1 via `difference`) */
    GLOME_INTERSECTION = 14,/* Csg.hs:15         a = first child node (contiguous), b = count   */
    GLOME_TEX = 15,         /* Tex.hs:27         a = child, b = texture id                      */
    GLOME_TAG = 16,         /* Tex.hs:26         a = child, b = tag id                          */
    GLOME_NOSHADOW = 17,    /* Tex.hs:28         a = child                                      */
    GLOME_ONLYSHADOW = 18,  /* Tex.hs:29         a = child                                      */
    GLOME_BOUND = 19,       /* Bound.hs:20       a = bounding object, b = bounded object        */
    GLOME_INNERBOUND = 20,  /* Bound.hs:95       a = inner object, b = outer object             */
    GLOME_NODE_TYPE_COUNT = 21
};

/* BIH flag bits (GlomeNode.c of a GLOME_BIH node) */
#define GLOME_BIH_LINEAR_SPHERES 1 /* every leaf item is a bare GLOME_SPHERE node and
                                      node[j].a == node[j0].a + 4*(j-j0) over the BIH's item block;
                                      j0 (the block's first node) is stored in c >> 4 */

/* BihBranch lsplit rsplit axis l r (Bih.hs:55-57), 32 bytes = two 16-byte loads.
 * Child refs: >= 0 index into bihnodes[]; < 0 leaf (BihLeaf [s]), k = ~ref:
 *   (k & 7) != 7 : inline leaf, item count = k & 7 (0..6), first item node index = k >> 3
 *   (k & 7) == 7 : large leaf, k >> 3 is an ipool offset of {first item node index, item count}
 * Leaf items are contiguous GlomeNode records, so a leaf visit needs no extra dependent load. */
typedef struct GlomeBihNode {
    double lsplit, rsplit;
    int32_t axis;
    int32_t left, right;
    int32_t pad;
} GlomeBihNode;

#if defined(__CUDACC__)
#define GLOME_HD __host__ __device__
#else
#define GLOME_HD
#endif
static inline GLOME_HD void glome_bih_leaf(int32_t ref, const int32_t* ipool, int32_t* first, int32_t* count) {
    int32_t k = ~ref;
    if ((k & 7) != 7) { *first = k >> 3; *count = k & 7; }
    else { *first = ipool[k >> 3]; *count = ipool[(k >> 3) + 1]; }
}
static inline int32_t glome_bih_leaf_ref_inline(int32_t first, int32_t count) { return ~((first << 3) | count); }
static inline int32_t glome_bih_leaf_ref_escape(int32_t ipool_off) { return ~((ipool_off << 3) | 7); }

/* Mesh BVH Branch lbb rbb l r (Mesh.hs:36), 128 bytes.  Child refs: >= 0 index into bvhnodes[];
 * < 0 leaf, k = ~ref is an ipool offset of {count, tri index ...} (mesh-local tri indices). */
typedef struct GlomeBvhNode {
    double lbb[6];
    double rbb[6];
    int32_t left, right;
    int32_t pad[6];
} GlomeBvhNode;

/* Mesh verts norms tris texs tags bb bvh (Mesh.hs:42); lives in ipool as 12 int32. */
typedef struct GlomeMeshHeader {
    int32_t bb_off;    /* dpool offset of bbox (6)                                    */
    int32_t root;      /* BVH root ref                                                */
    int32_t verts_off; /* dpool offset, 3 doubles per vertex                          */
    int32_t norms_off; /* dpool offset, 3 doubles per normal                          */
    int32_t tris_off;  /* ipool offset, 8 int32 per Tri {a,b,c,na,nb,nc,tex,tag} (Mesh.hs:29) */
    int32_t texs_off;  /* ipool offset: mesh texture index -> scene texture id        */
    int32_t tags_off;  /* ipool offset: mesh tag index -> scene tag id                */
    int32_t ntris, nverts, nnorms, ntexs, ntags;
} GlomeMeshHeader;

/* Reified Material (Shader.hs:43-52). */
enum GlomeMaterialKind {
    GLOME_MAT_SURFACE = 0,  /* p = {r,g,b,alpha,ambient,kd,ks,shine}  (dielectric flag unused by the shader) */
    GLOME_MAT_REFLECT = 1,  /* p[0] = refl                                                  */
    GLOME_MAT_REFRACT = 2,  /* p = {refl, refr, ior}                                        */
    GLOME_MAT_WARP = 3,     /* a = frame node, b = scene node, c = lightset, d = dpool offset of Xfm:
                               new ray = xfm_ray X (Ray hitpos (vnorm dir))  (TestScene.hs:169-173) */
    GLOME_MAT_ADDITIVE = 4, /* a = ipool offset of material ids, b = count                  */
    GLOME_MAT_BLEND = 5     /* a = material A, b = material B, p[0] = weight                */
};
typedef struct GlomeMaterial {
    int32_t kind;
    int32_t a, b, c, d;
    int32_t pad[3];
    double p[8];
} GlomeMaterial; /* 96 bytes */

/* Reified Texture = Ray -> Rayint -> Material closure (Solid.hs:97). */
enum GlomeTextureKind {
    GLOME_TEX_UNIFORM = 0,      /* a = material                          (Shader.hs:55)        */
    GLOME_TEX_STRIPE_BLEND = 1, /* Blend a b (triangle_wave (pos . axis)); p[0..2] = axis (TestScene.hs:225-231) */
    GLOME_TEX_PERLIN_BLEND = 2  /* Blend a b (perlin (vscale pos p[0]))  (TestScene.hs:214-220) */
};
typedef struct GlomeTexture {
    int32_t kind;
    int32_t a, b, c;
    double p[4];
} GlomeTexture; /* 48 bytes */

/* Light (Shader.hs:13-19).  falloff: 0 = \x -> 1/(x*x) (the only one in the code base, Shader.hs:23) */
typedef struct GlomeLight {
    double pos[3];
    double color[3];
    double rad;
    int32_t falloff;
    int32_t do_shadow;
} GlomeLight; /* 64 bytes */

typedef struct GlomeFlatScene {
    int32_t version; /* GLOME_FLAT_VERSION */
    int32_t root;    /* node index of the scene's SolidItem */
    int32_t n_nodes;
    int32_t n_bihnodes;
    int32_t n_bvhnodes;
    int32_t n_ipool;
    int32_t n_textures;
    int32_t n_materials;
    int32_t n_lights;
    int32_t n_lightsets; /* lightsets[2*i] = first light, lightsets[2*i+1] = count; set 0 = the scene's */
    int64_t n_dpool;
    const GlomeNode* nodes;
    const GlomeBihNode* bihnodes;
    const GlomeBvhNode* bvhnodes;
    const int32_t* ipool;
    const double* dpool;
    const GlomeTexture* textures;
    const GlomeMaterial* materials;
    const GlomeLight* lights;
    const int32_t* lightsets;
    int32_t max_depth;  /* static nesting depth of the scene graph (informational)            */
    int32_t scene_class;/* GLOME_CLASS_*: which kernel family the scene is eligible for       */
} GlomeFlatScene;

#define GLOME_CLASS_GENERAL 0 /* needs the recursive scene-graph interpreter                   */
#define GLOME_CLASS_FLAT 1    /* {Tex,Tag}* over prim | Bih[{Tex,Tag}* prim] | Mesh | Group of those;
                                 all materials are Surface: eligible for the wavefront kernels */

#define GLOME_MAX_STACK 8 /* per-ray texture / tag stack capacity on the device */

/* Rayint (Solid.hs:20-28) as a plain record: the id / t parity surface. */
typedef struct GlomeHit {
    double t;        /* ridepth; 1e6 (= infinity, Vec.hs:14) on a miss */
    double pos[3];
    double norm[3];
    int32_t hit;     /* 1 = RayHit, 0 = RayMiss */
    int32_t prim;    /* node index of the primitive (or Mesh) that produced the hit, -1 on miss */
    int32_t sub;     /* mesh-local triangle index for a Mesh hit, else -1 */
    int32_t ntex, ntag;
    int32_t flags;   /* GLOME_HITFLAG_* */
    int32_t tex[GLOME_MAX_STACK]; /* ritex, head first (innermost first) */
    int32_t tag[GLOME_MAX_STACK]; /* ritag, head first */
} GlomeHit; /* 144 bytes */

#define GLOME_HITFLAG_STACK_OVERFLOW 1 /* tex/tag stack deeper than GLOME_MAX_STACK */
#define GLOME_HITFLAG_CSG_OVERFLOW 2   /* rayint_advance chain hit the iteration cap */

typedef struct GlomeCamera { /* Camera pos fwd up right (Scene.hs:35); up/right pre-scaled by tan(fov/2) */
    double pos[3], fwd[3], up[3], right[3];
} GlomeCamera;

#define GLOME_MODE_ONE_RAY 0     /* renderTile: one get_color per pixel        (Glome.hs:162-176) */
#define GLOME_MODE_ADAPTIVE_AA 1 /* renderTileSubsample: 5-pass adaptive AA    (Glome.hs:226-323) */
#define GLOME_MODE_ADAPTIVE_AA_STRICT 2 /* the same frame, and also the reference's ray schedule: passes 1-4 trace exactly the
                                           samples their decisions ask for.  Mode 1 may instead trace every pixel centre up
                                           front when that is faster on this device (get_color is a pure function of the
                                           sample position, so the frame is the same bit for bit; only GlomeRenderStats' ray
                                           counts differ).  bench.py counts a frame's rays in this mode. */

typedef struct GlomeRenderOpts {
    int32_t mode;          /* GLOME_MODE_*                                                   */
    int32_t blocksize;     /* 65  (Glome.hs:116)                                             */
    int32_t recurs;        /* maxdepth = 3 (Glome.hs:25)                                     */
    int32_t tint_depth;    /* 1: renderTile's r + d/400 (Glome.hs:174); 0: plain get_color   */
    double thresholds[4];  /* 0.14 0.15 0.16 0.18 (Glome.hs:221-224)                         */
    int32_t tile_first;    /* multi-GPU sharding: render tiles i with i % tile_stride == tile_first */
    int32_t tile_stride;   /* 1 = all tiles                                                  */
    int32_t want_rgb8;     /* also pack 0x00RRGGBB (rgbf, Glome.hs:107-110)                  */
    int32_t debug_heatmap; /* 1: get_color_debug (Glome.hs:57-60): r += (n mod 30)/60, g += n/1000 with n = rayint_debug's
                              count for the camera ray; GLOME_MODE_ONE_RAY only */
} GlomeRenderOpts;

typedef struct GlomeRenderStats {
    int64_t rays_primary;   /* get_color calls (camera rays traced)                   */
    int64_t rays_shadow;    /* shadow queries issued by mpreshade                     */
    int64_t rays_secondary; /* reflect / refract / warp traces                        */
    int64_t overflow_rays;  /* rays that raised a GLOME_HITFLAG_*                     */
    int64_t perlin_range;   /* perlin results outside [0,1] (Texture.hs:109-116 would `error`) */
    double kernel_ms;       /* device time of the last call, CUDA events              */
    int32_t launches;       /* kernels launched by the last call                      */
    int32_t reserved;
    /* traversal work actually done (flat-class scenes only; 0 for the general interpreter):
     * the N_* terms of the algorithmic-bytes formula, SURVEY.md section 8(d) */
    int64_t visits_bih;     /* BIH branch nodes loaded (32 B each)                    */
    int64_t tests_prim;     /* primitive records tested (sphere 32 B, ...)            */
    int64_t visits_bvh;     /* Mesh BVH branch nodes loaded (128 B each)              */
    int64_t tests_tri;      /* mesh triangles tested (32 B Tri + 72 B vertices)       */
    double traverse_ms;     /* device time of the traversal kernels alone (K1 + K1'), CUDA events   */
    int64_t traverse_launches;
    /* the same, per kernel family: [0] k_bih_traverse<closest>  [1] k_bih_traverse<any>  [2] k_bvh_closest
     * [3] k_gen_trace (the general-scene tracer: traversal and shading in one persistent kernel) */
    double family_ms[4];
    int64_t family_launches[4];
    int64_t visits_instance; /* Instance transforms applied (192 B Xfm each); general scenes          */
    int64_t csg_steps;       /* rayint_advance re-issues (Solid.hs:85-91); general scenes             */
} GlomeRenderStats;

typedef struct GlomeScene GlomeScene; /* opaque device-resident scene */

const char* glome_last_error(void);
int glome_device_count(void);

/* scene flatten + upload, once; `desc` is copied. */
int glome_scene_create(const GlomeFlatScene* desc, int device, GlomeScene** out);
int glome_scene_destroy(GlomeScene* s);
/* The optional FP32 mode (north_star; GlomeVec/Data/Glome/Vec.hs:7-9: `type Flt = Double` with the note "make separate
 * Float and Double instances of this library"): the same scene evaluated with Flt = Float.  Payloads are rounded to float
 * once, at upload (16-byte BIH nodes, 64-byte BVH nodes, 16-byte spheres); every kernel is the FP64 kernel's source
 * compiled for float.  The handle is used with the ordinary entry points below (render, batches, pick); rays, cameras,
 * frames and GlomeHit stay double at the boundary.  Results differ from FP64 within the FP32 tolerance the north_star
 * states (RGB 1e-3 max abs away from silhouette pixels; tests/test_gpu_f32.py reports the id agreement rate).
 * Not available through glome_multi_create (shard an FP32 scene with tile_first / tile_stride instead). */
int glome_scene_create_f32(const GlomeFlatScene* desc, int device, GlomeScene** out);
/* Run-time switches of one scene, for measurements and A/B runs (a frame never depends on them):
 *   GLOME_OPT_SEG_CONCURRENT  1 (default): a Bih and a Mesh at the top of a flat scene are walked side by side on two
 *                             streams; 0: one after the other, so that GlomeRenderStats.family_ms times each kernel alone
 *   GLOME_OPT_AA_SPECULATE    -1 (default): adaptive AA keeps the faster of its two schedules (timed on the device);
 *                             0 / 1: always the reference's waves / always every pixel centre up front */
#define GLOME_OPT_SEG_CONCURRENT 1
#define GLOME_OPT_AA_SPECULATE 2
int glome_scene_set_option(GlomeScene* s, int option, int value);

/* rayint sld ray d [] []  (Solid.hs:146-151).  rays = n*6 doubles {ox,oy,oz,dx,dy,dz};
 * tmax = n doubles, or 1 double when tmax_stride == 0. */
int glome_rayint_batch(GlomeScene* s, int64_t n, const double* rays, const double* tmax,
                       int tmax_stride, GlomeHit* out);
/* shadow sld ray d  (Solid.hs:162) */
int glome_shadow_batch(GlomeScene* s, int64_t n, const double* rays, const double* tmax,
                       int tmax_stride, uint8_t* occluded);
/* inside sld pt  (Solid.hs:166); pts = n*3 doubles */
int glome_inside_batch(GlomeScene* s, int64_t n, const double* pts, uint8_t* inside);
/* trace lights materialShader sld ray depth recurs  (Trace.hs:59-82).
 * rgba = n*4 doubles (ColorA, not premultiplied), depth = n doubles (ridepth of the primary hit). */
int glome_trace_batch(GlomeScene* s, int64_t n, const double* rays, const double* tmax,
                      int tmax_stride, int recurs, double* rgba, double* depth, GlomeHit* hits_or_null);
/* rayint_debug's Int for each ray (Solid.hs:155, Bih.hs:378-412, Bound.hs:37-42): the number of BIH boxes the
 * reference's own walk enters -- the input of GlomeView's heat map (get_color_debug, Glome.hs:57-60; render it
 * with GlomeRenderOpts.debug_heatmap). */
int glome_debug_count_batch(GlomeScene* s, int64_t n, const double* rays, const double* tmax, int tmax_stride,
                            int32_t* counts);
/* getTags' / get_tags (Glome.hs:69-72, 410-414): the tag list of the trace result under pixel (px, py), what
 * GlomeView prints on a click: `ts ++ tags` (Trace.hs:82), i.e. the tags gathered by Reflect / Refract / Warp
 * recursion followed by the hit's own tag stack, head first.  *ntags = list length (device capacity 16);
 * tags receives min(*ntags, max_tags) ids; *truncated = 1 when the list was cut.  hit_out (optional) = the primary Rayint. */
int glome_get_tags(GlomeScene* s, const GlomeCamera* cam, int width, int height, int px, int py, int recurs,
                   int32_t* tags, int max_tags, int* ntags, int* truncated, GlomeHit* hit_out);
/* renderTiles (+ blitTile)  (Glome.hs:379-386, 353-358).  tcolor = w*h*5 doubles (r,g,b,a,depth),
 * row-major, or NULL when only the packed image is wanted; rgb8 = w*h uint32 or NULL.  Pixels of tiles not selected by tile_first/tile_stride are
 * left untouched. */
int glome_render(GlomeScene* s, const GlomeCamera* cam, int width, int height,
                 const GlomeRenderOpts* opts, double* tcolor, uint32_t* rgb8, GlomeRenderStats* stats);

/* Device-resident variant used by the benchmark's kernel-only timing and the multi-GPU path:
 * results stay in HBM; tcolor_dev / rgb8_dev are device pointers (cudaMalloc'ed by the caller or
 * by glome_dev_alloc). `stream` is a cudaStream_t (0 = default). */
int glome_render_dev(GlomeScene* s, const GlomeCamera* cam, int width, int height,
                     const GlomeRenderOpts* opts, double* tcolor_dev, uint32_t* rgb8_dev,
                     GlomeRenderStats* stats, void* stream);
/* cumulative number of kernels this scene handle has launched (bench.py's gpu_launches) */
int64_t glome_scene_launches(GlomeScene* s);
int glome_dev_alloc(int device, int64_t bytes, void** out);
int glome_dev_free(int device, void* p);

void glome_render_opts_default(GlomeRenderOpts* o);

/* Multi-GPU plumbing (the reference has none).  Rank r of N renders tiles
 * i with i % N == r; its tiles are packed into ceil(T/N) equal slots of blocksize^2 elements so
 * that one all-gather of equal-sized buffers collects the frame; unpack scatters rank r's block
 * back into a full frame.  elem_bytes = 40 (TColor) or 4 (rgb8).  Device pointers. */
int glome_tile_slots(int width, int height, int blocksize, int tile_stride);
int glome_tiles_pack_dev(int width, int height, int blocksize, int tile_first, int tile_stride, int elem_bytes,
                         const void* frame_dev, void* packed_dev, void* stream);
int glome_tiles_unpack_dev(int width, int height, int blocksize, int tile_first, int tile_stride, int elem_bytes,
                           const void* packed_dev, void* frame_dev, void* stream);
/* the whole all-gathered buffer (tile_stride blocks of glome_tile_slots slots) in one launch; skip_rank's block
 * (already in place) is skipped, -1 = none */
int glome_tiles_unpack_all_dev(int width, int height, int blocksize, int tile_stride, int skip_rank, int elem_bytes,
                               const void* gathered_dev, void* frame_dev, void* stream);

/* One process driving several GPUs (what a Haskell program links: no torch, no NCCL).  The scene is replicated on
 * devices[0..ndev); glome_multi_render renders tile i on device i % ndev (same kernels as glome_render), gathers the
 * tiles on devices[0] with peer copies over NVLink and returns the frame in the caller's host buffers (either may be
 * NULL).  The frame equals the 1-GPU frame bit for bit.  opts->tile_first / tile_stride must be 0 / 1. */
typedef struct GlomeMulti GlomeMulti;
int glome_multi_create(const GlomeFlatScene* desc, int ndev, const int* devices, GlomeMulti** out);
int glome_multi_destroy(GlomeMulti* m);
int glome_multi_render(GlomeMulti* m, const GlomeCamera* cam, int width, int height, const GlomeRenderOpts* opts,
                       double* tcolor, uint32_t* rgb8, GlomeRenderStats* stats);

/* Tile list helpers (chunk, Glome.hs:371-377): number of tiles and tile i's rect, in the order
 * renderTiles enumerates them (x-major). */
int glome_tile_count(int width, int height, int blocksize);
int glome_tile_rect(int width, int height, int blocksize, int i, int32_t rect[4]);

/* ---------------------------------------------------------------------------------------------
 * Host-side scene construction mirror (C++ in glome_b200/csrc/host, exposed here so tests and
 * bench.py can build scenes without Haskell).  Names follow the GlomeTrace constructors.
 * A builder owns a pool of host "SolidItem"s identified by int ids.
 * ------------------------------------------------------------------------------------------- */
typedef struct GlomeBuilder GlomeBuilder;

int glome_builder_create(GlomeBuilder** out);
int glome_builder_destroy(GlomeBuilder* b);
/* device >= 0: `bih` and `mesh` (glome_sb_bih, glome_sb_mesh, the config scenes) build their trees on that GPU; -1 (default): on the host.
 * Same tree either way.  last_build_ms: the last bih's / mesh's {H2D, device build, D2H, wall} in milliseconds. */
int glome_builder_set_build_device(GlomeBuilder* b, int device);
int glome_builder_last_build_ms(GlomeBuilder* b, double out[4]);

/* constructors: return item id >= 0, or a negative error */
int glome_sb_void(GlomeBuilder* b);
int glome_sb_sphere(GlomeBuilder* b, const double c[3], double r);                        /* Sphere.hs:15 */
int glome_sb_spheres(GlomeBuilder* b, int64_t n, const double* centers, const double* radii,
                     int32_t* ids_out);                                                   /* bulk */
int glome_sb_triangle(GlomeBuilder* b, const double p[9]);                                /* Triangle.hs:18 */
int glome_sb_trianglenorm(GlomeBuilder* b, const double pn[18]);                          /* Triangle.hs:34 */
int glome_sb_box(GlomeBuilder* b, const double p1[3], const double p2[3]);                /* Box.hs:12 */
int glome_sb_plane(GlomeBuilder* b, const double orig[3], const double norm[3]);          /* Plane.hs:17 */
int glome_sb_plane_offset(GlomeBuilder* b, const double norm[3], double off);             /* Plane.hs:24 */
int glome_sb_disc(GlomeBuilder* b, const double pos[3], const double norm[3], double r);  /* Cone.hs:29 */
int glome_sb_cylinder(GlomeBuilder* b, const double p1[3], const double p2[3], double r); /* Cone.hs:40 */
int glome_sb_cone(GlomeBuilder* b, const double p1[3], double r1, const double p2[3], double r2); /* Cone.hs:52 */
int glome_sb_cylinder_z(GlomeBuilder* b, double r, double h1, double h2);                 /* Cone.hs:33 */
int glome_sb_cone_z(GlomeBuilder* b, double r, double h1, double h2, double height);      /* Cone.hs:36 */
int glome_sb_group(GlomeBuilder* b, int n, const int32_t* items);                         /* Solid.hs:293 */
int glome_sb_bih(GlomeBuilder* b, int64_t n, const int32_t* items);                       /* Bih.hs:309 */
int glome_sb_mesh(GlomeBuilder* b, int64_t nverts, const double* verts, int64_t nnorms,
                  const double* norms, int64_t ntris, const int32_t* tris /*8 per tri*/,
                  int ntexs, const int32_t* texs, int ntags, const int32_t* tags);        /* Mesh.hs:50 */
/* Raw forms for the Haskell `flatten` instances (haskell/): the fields an already-constructed value holds, taken as they
 * are -- no constructor logic is re-run, so the device sees the very numbers the CPU path uses. */
int glome_sb_list(GlomeBuilder* b, int n, const int32_t* items);                          /* [s], Solid.hs:326 (no group flattening) */
int glome_sb_instance(GlomeBuilder* b, int item, const double xfm[24]);                   /* Instance s xfm, Solid.hs:386 */
int glome_sb_disc_raw(GlomeBuilder* b, const double pos[3], const double norm[3], double rsqr); /* Disc pos norm (r*r), Cone.hs:21 */
int glome_sb_difference_ex(GlomeBuilder* b, int sa, int sb, int useatex);                 /* Difference a b Bool, Csg.hs:14,26-30 */
/* A `Bih bb root` / `Mesh` whose tree the Haskell constructor ALREADY built (Bih.hs:51-57 `BihBranch lsplit rsplit axis l r`
 * | `BihLeaf [s]`; Mesh.hs:36-42 `Branch lbb rbb l r` | `Leaf [Tri]`), imported as a PRE-ORDER stream of n_nodes records
 * (a branch is followed by all records of l, then those of r) instead of being rebuilt:
 *   kinds[i] >= 0  a branch.  Bih: axis = kinds[i], lsplit = splits[2i], rsplit = splits[2i+1].
 *                             Mesh: lbb = boxes[12i..12i+5], rbb = boxes[12i+6..12i+11] (p1 then p2); kinds[i] is ignored.
 *   kinds[i] <  0  a leaf holding the next -(kinds[i]+1) entries of `items` (Bih; items are listed in the order the
 *                  leaves hold them) or of `leaf_tris` (Mesh; mesh-local triangle indices).
 * bb = the root box (Bih.hs:323 / Mesh.hs:212).  The item equals the one glome_sb_bih / glome_sb_mesh would return had
 * their own builders produced that tree (tests/test_host_builder.py).  Malformed streams: GLOME_EBUILD. */
int glome_sb_bih_prebuilt(GlomeBuilder* b, int64_t n, const int32_t* items, int64_t n_nodes, const int32_t* kinds,
                          const double* splits, const double bb[6]);                       /* Bih.hs:51-57 */
int glome_sb_mesh_prebuilt(GlomeBuilder* b, int64_t nverts, const double* verts, int64_t nnorms, const double* norms,
                           int64_t ntris, const int32_t* tris /*8 per tri*/, int ntexs, const int32_t* texs, int ntags,
                           const int32_t* tags, int64_t n_nodes, const int32_t* kinds, const double* boxes,
                           int64_t n_leaf_tris, const int32_t* leaf_tris, const double bb[6]); /* Mesh.hs:36-42 */
int glome_sb_difference(GlomeBuilder* b, int sa, int sb);                                 /* Csg.hs:26 */
int glome_sb_intersection(GlomeBuilder* b, int n, const int32_t* items);                  /* Csg.hs:64 */
int glome_sb_tex(GlomeBuilder* b, int item, int texture);                                 /* Tex.hs:33 */
int glome_sb_tag(GlomeBuilder* b, int item, int tag);                                     /* Tex.hs:38 */
int glome_sb_noshadow(GlomeBuilder* b, int item);                                         /* Tex.hs:43 */
int glome_sb_onlyshadow(GlomeBuilder* b, int item);                                       /* Tex.hs:48 */
int glome_sb_bound_object(GlomeBuilder* b, int sa, int sb);                               /* Bound.hs:27 */
int glome_sb_innerbound(GlomeBuilder* b, int sa, int sb);                                 /* Bound.hs:116 */
/* transform item [xfm...]: xfms = nx * 24 doubles (fwd, inv), composed in list order (Solid.hs:184, Vec.hs:461) */
int glome_sb_transform(GlomeBuilder* b, int item, int nx, const double* xfms);
/* basic transforms (Vec.hs:564-598): write 24 doubles */
int glome_xfm_translate(const double v[3], double out[24]);
int glome_xfm_scale(const double v[3], double out[24]);
int glome_xfm_rotate(const double axis[3], double angle, double out[24]);
int glome_xfm_compose(int n, const double* xfms, double out[24]);
/* bih (tolist (SolidItem (flatten_transform item))) helper pieces (Solid.hs:177-246) */
int glome_sb_flatten_transform_bih(GlomeBuilder* b, int item);
/* bound item (Solid.hs:171): out = {p1, p2} */
int glome_sb_bound(GlomeBuilder* b, int item, double out[6]);

/* materials / textures / lights */
int glome_sb_mat_surface(GlomeBuilder* b, const double rgb[3], double alpha, double amb, double kd,
                         double ks, double shine);
int glome_sb_mat_reflect(GlomeBuilder* b, double refl);
int glome_sb_mat_refract(GlomeBuilder* b, double refl, double refr, double ior);
int glome_sb_mat_warp(GlomeBuilder* b, int frame_item, int scene_item, int lightset, const double xfm[24]);
int glome_sb_mat_additive(GlomeBuilder* b, int n, const int32_t* mats);
int glome_sb_mat_blend(GlomeBuilder* b, int ma, int mb, double weight);
int glome_sb_tex_uniform(GlomeBuilder* b, int mat);
int glome_sb_tex_stripe_blend(GlomeBuilder* b, int ma, int mb, const double axis[3]);
int glome_sb_tex_perlin_blend(GlomeBuilder* b, int ma, int mb, double scale);
int glome_sb_light(GlomeBuilder* b, const double pos[3], const double color[3]);          /* Shader.hs:22 */
int glome_sb_lightset(GlomeBuilder* b, int n, const int32_t* lights);
/* a Warp material may name the scene root before it exists (TestScene.hs:179): patch it afterwards */
int glome_sb_mat_warp_set_scene(GlomeBuilder* b, int mat, int scene_item);

/* camera pos at up angle (Scene.hs:48-57) */
int glome_camera(const double pos[3], const double at[3], const double up[3], double angle_deg,
                 GlomeCamera* out);

/* flatten the graph under `root` into builder-owned arrays; `out` stays valid until the next
 * flatten or builder_destroy. */
int glome_sb_flatten(GlomeBuilder* b, int root, GlomeFlatScene* out);

/* ready-made scenes for BASELINE.json's configs; return the root item, fill camera.
 *   1: TestScene.geom'' (TestScene.hs:183-197), oak with a substitute PRNG (seed)
 *   2: bih of n random spheres, 2 lights        3: n-triangle height-field Mesh (+ occluder bih)
 *   4: CSG-heavy grid                            (5 = scene 3, rendered with adaptive AA) */
int glome_sb_config_scene(GlomeBuilder* b, int config, int64_t n, uint64_t seed, GlomeCamera* cam,
                          int* recurs_out);

/* TestScene's oak draws from System.Random's StdGen (TestScene.hs:83-88,190).  Probe of the restated generator
 * (random-1.2 / splitmix, glome_b200/csrc/scenes.cpp) for known-answer tests and for comparison with a GHC run:
 * out = {seed, gamma of `mkStdGen n`; its next three Word64; seed, gamma of both halves of `split`},
 * dout = three successive `randomR (0, 0.5) :: Double`, vigna = first output of Vigna's splitmix64 from x = n. */
int glome_stdgen_probe(int64_t n, uint64_t out[9], double dout[3], uint64_t* vigna);

/* NFF (Neutral File Format, the SPD benchmark scenes) reader: Spd.hs:1-261, quirks included (groups and lights end
 * up in reverse file order, the last camera / background win, a "#" comment ends the scene; see nff.cpp).
 * Adds the scene's items, materials and lights to the builder and returns the root item
 * (`bih` of one `tex (bih prims) fill` per "f" group); cam / bg (3 doubles) / consumed (bytes parsed) may be NULL.
 * len < 0: text is NUL-terminated.  GLOME_EBUILD when the text has no camera or no background. */
int glome_sb_load_nff(GlomeBuilder* b, const char* text, int64_t len, GlomeCamera* cam, double bg[3], int64_t* consumed);

/* Tree builders on their own (host, multi-threaded): used by glome_sb_bih / glome_sb_mesh and
 * compared against the oracle's literal restatement in tests.
 * bboxes = n*6 doubles.  Outputs are malloc'ed; free with glome_free. */
int glome_bih_build(int64_t n, const double* bboxes, GlomeBihNode** nodes_out, int32_t* n_nodes_out,
                    int32_t** leaves_out /* {first,count} pairs in item_order */, int32_t* n_leaves_out,
                    int32_t** item_order_out /* n: leaf-ordered permutation of 0..n-1 */,
                    int32_t* root_ref_out, double bb_out[6]);
/* the same tree, built level by level on a GPU (glome_b200/csrc/glome_build.cu).
 * Replaces build_rec's list recursion (Bih.hs:211-285); output arrays are identical to glome_bih_build's.
 * timings_ms = {host->device, device build, device->host} or NULL.  GLOME_ENODEV without a device. */
int glome_bih_build_gpu(int64_t n, const double* bboxes, int device, GlomeBihNode** nodes_out, int32_t* n_nodes_out,
                        int32_t** leaves_out, int32_t* n_leaves_out, int32_t** item_order_out,
                        int32_t* root_ref_out, double bb_out[6], double timings_ms[3]);
/* mesh BVH (Mesh.hs:50-134): leafpool = {count, tri...} records, leafoff[leaf] = offset into leafpool */
int glome_mesh_build(int64_t nverts, const double* verts, int64_t ntris, const int32_t* tris /*8 per tri*/,
                     GlomeBvhNode** nodes_out, int32_t* n_nodes_out, int32_t** leafpool_out,
                     int32_t* n_leafpool_out, int32_t** leafoff_out, int32_t* n_leaves_out,
                     int32_t* root_ref_out, double bb_out[6]);
/* the same BVH built on a GPU (glome_build.cu); replaces build_tree's list recursion (Mesh.hs:69-113) */
int glome_mesh_build_gpu(int64_t nverts, const double* verts, int64_t ntris, const int32_t* tris /*8 per tri*/, int device,
                         GlomeBvhNode** nodes_out, int32_t* n_nodes_out, int32_t** leafpool_out,
                         int32_t* n_leafpool_out, int32_t** leafoff_out, int32_t* n_leaves_out,
                         int32_t* root_ref_out, double bb_out[6], double timings_ms[3]);
void glome_free(void* p);
/* The GPU tree builders keep their device work space (about 0.5 GB for 10^6 items) for the next build, because allocating
 * and freeing it costs several times the build itself.  This returns the cached blocks of `device` to the driver. */
int glome_build_release_cache(int device);

#ifdef __cplusplus
}
#endif

/* Layout contract of the structs that cross the boundary by address.  The Haskell binding (haskell/Data/Glome/CUDA.hs)
 * peeks / pokes them at these byte offsets (it has no hsc2hs step); a change here fails the build of the library. */
#include <stddef.h>
#ifdef __cplusplus
#define GLOME_LAYOUT_ASSERT(c) static_assert(c, #c)
#else
#define GLOME_LAYOUT_ASSERT(c) _Static_assert(c, #c)
#endif
GLOME_LAYOUT_ASSERT(sizeof(GlomeRenderOpts) == 64);
GLOME_LAYOUT_ASSERT(offsetof(GlomeRenderOpts, mode) == 0 && offsetof(GlomeRenderOpts, blocksize) == 4);
GLOME_LAYOUT_ASSERT(offsetof(GlomeRenderOpts, recurs) == 8 && offsetof(GlomeRenderOpts, tint_depth) == 12);
GLOME_LAYOUT_ASSERT(offsetof(GlomeRenderOpts, thresholds) == 16 && offsetof(GlomeRenderOpts, tile_first) == 48);
GLOME_LAYOUT_ASSERT(offsetof(GlomeRenderOpts, tile_stride) == 52 && offsetof(GlomeRenderOpts, want_rgb8) == 56);
GLOME_LAYOUT_ASSERT(offsetof(GlomeRenderOpts, debug_heatmap) == 60);
GLOME_LAYOUT_ASSERT(sizeof(GlomeHit) == 144);
GLOME_LAYOUT_ASSERT(offsetof(GlomeHit, t) == 0 && offsetof(GlomeHit, pos) == 8 && offsetof(GlomeHit, norm) == 32);
GLOME_LAYOUT_ASSERT(offsetof(GlomeHit, hit) == 56 && offsetof(GlomeHit, prim) == 60 && offsetof(GlomeHit, sub) == 64);
GLOME_LAYOUT_ASSERT(offsetof(GlomeHit, ntex) == 68 && offsetof(GlomeHit, ntag) == 72 && offsetof(GlomeHit, flags) == 76);
GLOME_LAYOUT_ASSERT(offsetof(GlomeHit, tex) == 80 && offsetof(GlomeHit, tag) == 112);
GLOME_LAYOUT_ASSERT(sizeof(GlomeCamera) == 96);   /* pos, fwd, up, right: 12 doubles */
GLOME_LAYOUT_ASSERT(sizeof(GlomeFlatScene) <= 256); /* the binding allocates 256 bytes for glome_sb_flatten's output */
GLOME_LAYOUT_ASSERT(sizeof(GlomeBihNode) == 32 && sizeof(GlomeBvhNode) == 128 && sizeof(GlomeNode) == 16);
#endif /* GLOME_CUDA_H */
