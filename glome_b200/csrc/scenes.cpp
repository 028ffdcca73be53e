// scenes.cpp -- the measurement scenes of BASELINE.json, written against the host mirror of the
// GlomeTrace constructors so they read like the reference's TestScene.hs.
//
//   1: TestScene.geom''  (GlomeView/TestScene.hs:183-197)  -- oak uses a substitute PRNG, see below
//   2: bih over n random spheres, 2 point lights with shadows
//   3: n-triangle height-field Mesh (shared vertex/normal arrays, per-triangle tex/tag) + occluder bih
//   4: CSG-heavy grid (nested difference/intersection of box, sphere, cylinder, cone), mirrors
#include <cmath>
#include <cstring>

#include "host_builder.h"

namespace glome_host {

using namespace glm;

static inline uint64_t splitmix64(uint64_t& s) {
    uint64_t z = (s += 0x9E3779B97F4A7C15ULL);
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ULL;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBULL;
    return z ^ (z >> 31);
}
static inline double u01(uint64_t& s) { return (double)(splitmix64(s) >> 11) * (1.0 / 9007199254740992.0); }
static inline double urange(uint64_t& s, double lo, double hi) { return lo + u01(s) * (hi - lo); }

// ---- some textures (TestScene.hs:201-245) ----
struct TestTex {
    int m_shiny_white, m_shiny_red, m_dull_gray, m_mirror;
    int t_shiny_white, t_shiny_red, t_mottled, t_stripe, t_mirror;
};
static int m_matte(Builder& b, Flt r, Flt g, Flt bl) { return b.mat_surface(r, g, bl, 1, 0.2, 1, 0, 0); }  // TestScene.hs:236
static int t_matte(Builder& b, Flt r, Flt g, Flt bl) { return b.tex_uniform(m_matte(b, r, g, bl)); }       // TestScene.hs:239
static TestTex make_testtex(Builder& b) {
    TestTex t;
    t.m_shiny_white = b.mat_surface(1, 1, 1, 1, 0.2, 0.8, 0.4, 10);  // TestScene.hs:202
    t.m_shiny_red = b.mat_surface(1, 0, 0, 1, 0.2, 0.8, 0.4, 10);    // TestScene.hs:204
    t.m_dull_gray = b.mat_surface(0.4, 0.3, 0.35, 1, 0.2, 0.8, 0, 0);  // TestScene.hs:212
    t.m_mirror = b.mat_reflect(0.8);                                 // TestScene.hs:243
    t.t_shiny_white = b.tex_uniform(t.m_shiny_white);
    t.t_shiny_red = b.tex_uniform(t.m_shiny_red);
    t.t_mottled = b.tex_perlin_blend(t.m_mirror, m_matte(b, 0.15, 0.3, 0.5), 3);        // TestScene.hs:214-220
    t.t_stripe = b.tex_stripe_blend(t.m_shiny_white, t.m_dull_gray, vec(4, 8, 5));      // TestScene.hs:225-231
    t.t_mirror = b.tex_uniform(t.m_mirror);
    return t;
}

enum { TAG_DODECAHEDRON = 1, TAG_ICOSAHEDRON = 2, TAG_TREE = 3, TAG_DOOR_FRAME = 4 };

static int dodecahedron(Builder& b, const Vec& pos, Flt r) {  // TestScene.hs:45-54
    Flt gr = (1 + sqrt(5.0)) / 2;
    Flt n11[2] = {-r, r};
    Flt ngrgr[2] = {(-gr) * r, gr * r};
    std::vector<Vec> points;
    for (int y = 0; y < 2; y++) for (int z = 0; z < 2; z++) points.push_back(vec(0, n11[y], ngrgr[z]));
    for (int z = 0; z < 2; z++) for (int x = 0; x < 2; x++) points.push_back(vec(ngrgr[x], 0, n11[z]));
    for (int x = 0; x < 2; x++) for (int y = 0; y < 2; y++) points.push_back(vec(n11[x], ngrgr[y], 0));
    std::vector<int32_t> xs;
    xs.push_back(b.sphere(pos, 1.26 * r));
    for (size_t i = 0; i < points.size(); i++) {
        Vec n = vnorm(points[i]);
        xs.push_back(b.plane_offset(n, r + vdot(n, pos)));
    }
    return b.tag(b.intersection(xs), TAG_DODECAHEDRON);
}
static int icosahedron(Builder& b, const Vec& pos, Flt r) {  // TestScene.hs:27-43
    Flt gr = (1 + sqrt(5.0)) / 2;
    Flt n11[2] = {-r, r};
    Flt ngrgr[2] = {(-gr) * r, gr * r};
    Flt grrcp[2] = {(-r) / gr, r / gr};
    std::vector<Vec> points;
    for (int x = 0; x < 2; x++) for (int y = 0; y < 2; y++) for (int z = 0; z < 2; z++) points.push_back(vec(n11[x], n11[y], n11[z]));
    for (int y = 0; y < 2; y++) for (int z = 0; z < 2; z++) points.push_back(vec(0, grrcp[y], ngrgr[z]));
    for (int x = 0; x < 2; x++) for (int y = 0; y < 2; y++) points.push_back(vec(grrcp[x], ngrgr[y], 0));
    for (int x = 0; x < 2; x++) for (int z = 0; z < 2; z++) points.push_back(vec(ngrgr[x], 0, grrcp[z]));
    std::vector<int32_t> xs;
    xs.push_back(b.sphere(pos, 1.26 * r));
    for (size_t i = 0; i < points.size(); i++) {
        Vec n = vnorm(points[i]);
        xs.push_back(b.plane_offset(n, r + vdot(n, pos)));
    }
    return b.tag(b.intersection(xs), TAG_ICOSAHEDRON);
}
static int lattice(Builder& b) {  // TestScene.hs:21-25
    std::vector<int32_t> xs;
    for (int x = -10; x <= 10; x++)
        for (int y = -10; y <= 10; y++)
            for (int z = -10; z <= 10; z++) xs.push_back(b.sphere(vec(x, y, z), 0.2));
    return b.bih(xs);
}

// oak (TestScene.hs:68-110).  The reference draws the branch parameters from System.Random's StdGen: `mkStdGen 42`,
// `split`, `randomR` (TestScene.hs:83-88, 190).  GlomeView.cabal:26 leaves the `random` version open; every GHC install of
// the last years resolves it to random >= 1.2, whose StdGen is splitmix's SMGen.  Restated here from the published
// algorithm (PINNED ASSUMPTION: random-1.2.1 with splitmix-0.1.x; the sources are not in this image, so this sub-object
// stays "parity unpinned" until bench/GlomeHeadless.hs prints the same first draws on a GHC box):
//   SMGen seed gamma;  mkSMGen s = SMGen (mix64 s) (mixGamma (s + goldenGamma))          (System.Random.SplitMix)
//   nextWord64 (SMGen s g) = (mix64 (s + g), SMGen (s + g) g)
//   splitSMGen (SMGen s g) = (SMGen s'' g, SMGen (mix64 s') (mixGamma s''))   with s' = s + g, s'' = s' + g
//   randomR (l, h) :: Double = x * l + (1 - x) * h,  x = fromIntegral w64 / fromIntegral (maxBound :: Word64)
//                                                                       (System.Random.Internal: uniformRM, uniformDouble01M)
struct StdGen { uint64_t seed, gamma; };
static inline uint64_t sm_shift_xor(int n, uint64_t w) { return w ^ (w >> n); }
static inline uint64_t sm_mix64(uint64_t z) {  // MurmurHash3's finaliser, as in java.util.SplittableRandom
    z = sm_shift_xor(33, z) * 0xff51afd7ed558ccdULL;
    z = sm_shift_xor(33, z) * 0xc4ceb9fe1a85ec53ULL;
    return sm_shift_xor(33, z);
}
static inline uint64_t sm_mix64variant13(uint64_t z) {  // Stafford's variant 13 (Vigna's splitmix64 output function)
    z = sm_shift_xor(30, z) * 0xbf58476d1ce4e5b9ULL;
    z = sm_shift_xor(27, z) * 0x94d049bb133111ebULL;
    return sm_shift_xor(31, z);
}
static inline uint64_t sm_mix_gamma(uint64_t z) {
    uint64_t g = sm_mix64variant13(z) | 1ULL;
    int n = __builtin_popcountll(g ^ (g >> 1));
    return n >= 24 ? g : g ^ 0xaaaaaaaaaaaaaaaaULL;
}
static const uint64_t SM_GOLDEN_GAMMA = 0x9e3779b97f4a7c15ULL;
StdGen mkStdGen(int64_t n) {
    uint64_t s = (uint64_t)n;
    StdGen g = {sm_mix64(s), sm_mix_gamma(s + SM_GOLDEN_GAMMA)};
    return g;
}
uint64_t stdgen_next_word64(StdGen& g) {
    g.seed += g.gamma;
    return sm_mix64(g.seed);
}
void stdgen_split(const StdGen& g, StdGen& a, StdGen& b) {
    uint64_t s1 = g.seed + g.gamma, s2 = s1 + g.gamma;
    a.seed = s2; a.gamma = g.gamma;
    b.seed = sm_mix64(s1); b.gamma = sm_mix_gamma(s2);
}
double stdgen_randomR_double(double l, double h, StdGen& g) {
    if (l == h) return l;
    double x = (double)stdgen_next_word64(g) / 18446744073709551615.0;  // fromIntegral w64 / fromIntegral (maxBound :: Word64)
    return x * l + (1 - x) * h;
}
uint64_t splitmix64_vigna(uint64_t& x) { return sm_mix64variant13(x += SM_GOLDEN_GAMMA); }  // known-answer hook (tests)

typedef StdGen OakRng;
static void oak_split(const OakRng& r, OakRng& a, OakRng& c) { stdgen_split(r, a, c); }
static int oak_tree(Builder& b, int n_, OakRng r, Flt season, int t_leaf) {
    const Flt thickness = 0.03;
    const Flt minbranch = deg(10), maxbranch = deg(25);
    if (n_ == 0) return b.void_();
    if (n_ == 1) return b.tex(b.sphere(vec(0, 0, 0), season), t_leaf);
    Flt nf = (Flt)n_;
    Flt height = nf;
    OakRng rng1, rng2, rng3, rng4;
    oak_split(r, rng1, rng2);
    oak_split(rng1, rng3, rng4);
    OakRng rng5 = rng4;  // (r1,rng5) = randomR (0,0.5) rng4 ... : one generator threaded through the three draws
    Flt r1 = stdgen_randomR_double(0, 0.5, rng5);
    Flt r2 = stdgen_randomR_double(minbranch, maxbranch, rng5);
    Flt r3 = stdgen_randomR_double(0.8, 0.95, rng5);
    // r4 :: Float in (0,1) is compared with 1 and never exceeds it (TestScene.hs:87,92-94): it selects nothing
    Flt seglen = 0.5 + r1, branchang = r2, scaling = r3;
    int n = n_;
    std::vector<int32_t> parts;
    parts.push_back(b.cone(vec(0, 0, 0), thickness * height, vec(0, seglen, 0), thickness * (height - 1) * scaling));
    for (int side = 0; side < 2; side++) {
        std::vector<Xfm> xs;
        xs.push_back(scale(vec(scaling, scaling, scaling)));
        xs.push_back(rotate(vec(0, 0, 1), side == 0 ? branchang : -branchang));
        xs.push_back(rotate(vec(0, 1, 0), deg(30)));
        xs.push_back(translate(vec(0, seglen, 0)));
        parts.push_back(b.transform(oak_tree(b, n - 1, side == 0 ? rng2 : rng3, season, t_leaf), xs));
    }
    return b.bound_object(b.sphere(vec(0, height / 2, 0), height / 2), b.group(parts));
}
static int oak(Builder& b, Flt age, uint64_t seed) {
    if (age < 0) return b.void_();
    int year = (int)floor(age);
    Flt season = age - (Flt)year;
    int t_leaf = t_matte(b, 0.2, 1, 0.4);
    OakRng r = mkStdGen((int64_t)seed);  // oak 11.4 (mkStdGen 42)  (TestScene.hs:190)
    int tree = oak_tree(b, year, r, season, t_leaf);
    // bih (tolist (SolidItem (flatten_transform (tree year rng))))
    std::vector<int32_t> leaves = b.tolist(b.list_raw(b.flatten_transform(tree)));
    return b.tag(b.tex(b.bih(leaves), t_matte(b, 0.8, 0.5, 0.4)), TAG_TREE);
}

static int chessboard(Builder& b, const TestTex& t) {  // TestScene.hs:140-150
    std::vector<int32_t> xs;
    for (int i = 0; i < 8; i++)
        for (int j = 0; j < 8; j++) {
            Flt x = -3.5 + i, z = -3.5 + j;
            Flt f = (x * z) / 40;
            long fl = (long)floor(x) + (long)floor(z);
            long m = ((fl % 2) + 2) % 2;
            int tx = (m == 0) ? t.t_shiny_white : t.t_mottled;
            xs.push_back(b.tex(b.box(vec(x - (1.0 / 2), -3, z - (1.0 / 2)), vec(x + (1.0 / 2), f, z + (1.0 / 2))), tx));
        }
    return b.group(xs);
}

static int portal(Builder& b, Flt height, Flt width, Flt thickness, int* warp_mat) {  // TestScene.hs:152-179
    int frame = b.tag(
        b.tex(b.difference(b.box(vec(-width, 0, -thickness), vec(width, height, thickness)),
                           b.box(vec(thickness - width, thickness, -(thickness + GLM_DELTA)),
                                 vec(width - thickness, height - thickness, thickness + GLM_DELTA))),
              t_matte(b, 0.4, 0.4, 0.8)),
        TAG_DOOR_FRAME);
    int surface = b.box(vec(-width, 0, -GLM_DELTA), vec(width, height - GLM_DELTA, GLM_DELTA));
    std::vector<Xfm> xs;
    xs.push_back(rotate(vec(1, 0, 0), deg(-85)));
    xs.push_back(translate(vec(8, 40, -4)));
    *warp_mat = b.mat_warp(frame, -1, 0, compose(xs));
    std::vector<int32_t> g;
    g.push_back(frame);
    g.push_back(b.tex(surface, b.tex_uniform(*warp_mat)));
    return b.group(g);
}

static int scene_testscene(Builder& b, uint64_t seed, GlomeCamera* cam, int* recurs) {
    // lights (TestScene.hs:17-19)
    b.light(vec(-100, 70, 140), 1 * 7000.0, 0.8 * 7000.0, 0.8 * 7000.0);
    b.light(vec(-3, 5, 8), 1.5 * 10.0, 2 * 10.0, 2 * 10.0);
    TestTex t = make_testtex(b);
    std::vector<int32_t> g;
    {
        std::vector<Xfm> xs(1, scale(vec(2, 1.2, 2)));
        g.push_back(b.difference(b.transform(chessboard(b, t), xs), b.tex(b.sphere(vec(4, 1.5, 3), 3.5), t.t_shiny_white)));
    }
    g.push_back(b.tex(dodecahedron(b, vec(-6, 3, 0), 1), t.t_stripe));
    {
        std::vector<Xfm> xs;
        xs.push_back(rotate(vec(0, 0, 1), deg(11)));
        xs.push_back(rotate(vec(1, 0, 0), deg(7)));
        g.push_back(b.tex(b.transform(icosahedron(b, vec(4, 1.5, 3), 1.5), xs), t.t_mottled));
    }
    g.push_back(b.cone(vec(-6, -1, 0), 0.7, vec(-6, 3, 0), 0));
    {
        std::vector<Xfm> xs;
        xs.push_back(scale(vec(2, 2, 2)));
        xs.push_back(translate(vec(2, -1, -8)));
        g.push_back(b.transform(oak(b, 11.4, seed), xs));
    }
    {
        std::vector<Xfm> xs;
        xs.push_back(rotate(vec(0, 0, 1), deg(23)));
        xs.push_back(rotate(vec(1, 0, 0), deg(43)));
        xs.push_back(scale(vec(3, 3, 3)));
        g.push_back(b.tex(b.difference(b.transform(lattice(b), xs), b.sphere(vec(0, 0, 0), 32)), t.t_shiny_red));
    }
    int warp_mat = -1;
    {
        std::vector<Xfm> xs;
        xs.push_back(rotate(vec(0, 1, 0), deg(8)));
        xs.push_back(translate(vec(-3, 0.5, -5)));
        g.push_back(b.transform(portal(b, 5, 2, 1.0 / 3, &warp_mat), xs));
    }
    {
        std::vector<Xfm> xs(1, scale(vec(1, 0.4, 1)));
        g.push_back(b.transform(b.tex(b.sphere(vec(-2.3, 0.3, 4.2), 1.7), b.tex_uniform(b.mat_refract(0.35, 0.8, 1.5))), xs));
    }
    int root = b.bih(g);
    b.materials[warp_mat].b = root;  // Warp frame geom'' lights xfm: the scene is geom'' itself
    make_camera(vec(-2, 4.3, 15), vec(0, 2, 0), vec(0, 1, 0), 45, cam);  // cust_cam (TestScene.hs:138)
    *recurs = 3;  // maxdepth (Glome.hs:25)
    return root;
}

// ---- config 2: bih of n random spheres ----
static int scene_sphere_cloud(Builder& b, int64_t n, uint64_t seed, GlomeCamera* cam, int* recurs) {
    if (n <= 0) n = 1000000;
    Flt half = 100.0 * cbrt((double)n / 1.0e6);  // constant density: 1e6 spheres in [-100,100]^3
    uint64_t s = seed;
    std::vector<int32_t> xs((size_t)n);
    b.items.reserve(b.items.size() + (size_t)n + 16);
    for (int64_t i = 0; i < n; i++) {
        Flt x = urange(s, -half, half), y = urange(s, -half, half), z = urange(s, -half, half);
        Flt r = urange(s, 0.05, 0.5);
        xs[(size_t)i] = b.sphere(vec(x, y, z), r);
    }
    Flt k = half * half;
    b.light(vec(3.0 * half, 4.0 * half, 1.0 * half), 15.0 * k, 15.0 * k, 14.0 * k);
    b.light(vec(-2.0 * half, 3.0 * half, 4.0 * half), 6.0 * k, 7.0 * k, 9.0 * k);
    int root = b.tex(b.bih(xs), t_matte(b, 0.8, 0.7, 0.5));
    make_camera(vec(1.7 * half, 1.2 * half, 1.45 * half), vec(0, 0, 0), vec(0, 1, 0), 45, cam);
    *recurs = 3;
    return root;
}

// ---- config 3: height-field Mesh ----
static inline double hash01(uint64_t seed, int64_t i, int64_t j) {
    uint64_t s = seed ^ ((uint64_t)i * 0x9E3779B97F4A7C15ULL) ^ ((uint64_t)j * 0xC2B2AE3D27D4EB4FULL);
    return u01(s);
}
static double value_noise(uint64_t seed, double x, double z) {
    double fx = floor(x), fz = floor(z);
    int64_t i = (int64_t)fx, j = (int64_t)fz;
    double u = x - fx, v = z - fz;
    u = u * u * (3 - 2 * u);
    v = v * v * (3 - 2 * v);
    double a = hash01(seed, i, j), bq = hash01(seed, i + 1, j), c = hash01(seed, i, j + 1), d = hash01(seed, i + 1, j + 1);
    return (a * (1 - u) + bq * u) * (1 - v) + (c * (1 - u) + d * u) * v;
}
static double terrain(uint64_t seed, double x, double z) {
    return 14.0 * value_noise(seed, x * 0.035, z * 0.035) + 5.0 * value_noise(seed + 1, x * 0.11, z * 0.11) +
           1.2 * value_noise(seed + 2, x * 0.37, z * 0.37);
}
static int scene_heightfield(Builder& b, int64_t ntris, uint64_t seed, GlomeCamera* cam, int* recurs) {
    if (ntris <= 0) ntris = 2000000;
    int g = (int)floor(sqrt((double)ntris / 2.0) + 0.5);
    if (g < 1) g = 1;
    int nv = (g + 1) * (g + 1);
    std::vector<double> verts((size_t)nv * 3), norms((size_t)nv * 3);
    const double half = 100.0;
    const double step = 2 * half / g;
    for (int j = 0; j <= g; j++)
        for (int i = 0; i <= g; i++) {
            double x = -half + i * step, z = -half + j * step;
            double y = terrain(seed, x, z);
            size_t k = (size_t)(j * (g + 1) + i) * 3;
            verts[k] = x; verts[k + 1] = y; verts[k + 2] = z;
            double e = 0.25 * step;
            double dx = (terrain(seed, x + e, z) - terrain(seed, x - e, z)) / (2 * e);
            double dz = (terrain(seed, x, z + e) - terrain(seed, x, z - e)) / (2 * e);
            Vec nn = vnorm(vec(-dx, 1, -dz));
            norms[k] = nn.x; norms[k + 1] = nn.y; norms[k + 2] = nn.z;
        }
    std::vector<int32_t> tris;
    tris.reserve((size_t)g * g * 16);
    for (int j = 0; j < g; j++)
        for (int i = 0; i < g; i++) {
            int v00 = j * (g + 1) + i, v10 = v00 + 1, v01 = v00 + (g + 1), v11 = v01 + 1;
            int tex = ((i / 8) + (j / 8)) % 4;
            int tag = ((i / 32) + 4 * (j / 32)) % 16;
            int32_t t1[8] = {v00, v01, v10, v00, v01, v10, tex, tag};
            int32_t t2[8] = {v10, v01, v11, v10, v01, v11, tex, tag};
            tris.insert(tris.end(), t1, t1 + 8);
            tris.insert(tris.end(), t2, t2 + 8);
        }
    int32_t texs[4] = {t_matte(b, 0.35, 0.55, 0.25), t_matte(b, 0.5, 0.45, 0.3), t_matte(b, 0.6, 0.6, 0.55),
                       b.tex_uniform(b.mat_surface(0.3, 0.4, 0.6, 1, 0.2, 0.8, 0.4, 10))};
    int32_t tags[16];
    for (int i = 0; i < 16; i++) tags[i] = 100 + i;
    int mesh = b.mesh(nv, verts.data(), nv, norms.data(), (int64_t)tris.size() / 8, tris.data(), 4, texs, 16, tags);
    // Mesh casts no shadows in the reference (Mesh.hs:210): give shadow rays real work with a bih of
    // occluder spheres floating above the terrain.
    uint64_t s = seed + 77;
    int nocc = 4096;
    std::vector<int32_t> occ((size_t)nocc);
    for (int i = 0; i < nocc; i++) {
        Flt x = urange(s, -half, half), z = urange(s, -half, half), y = urange(s, 24, 45), r = urange(s, 0.4, 1.6);
        occ[(size_t)i] = b.sphere(vec(x, y, z), r);
    }
    std::vector<int32_t> gl;
    gl.push_back(mesh);
    gl.push_back(b.tex(b.bih(occ), t_matte(b, 0.9, 0.3, 0.2)));
    b.light(vec(150, 400, 120), 1.9e5, 1.8e5, 1.6e5);
    b.light(vec(-220, 300, -80), 0.7e5, 0.8e5, 1.0e5);
    int root = b.group(gl);
    make_camera(vec(70, 52, 105), vec(-10, 0, -20), vec(0, 1, 0), 45, cam);
    *recurs = 3;
    return root;
}

// ---- config 4: CSG-heavy grid ----
static int scene_csg_grid(Builder& b, int64_t side, uint64_t seed, GlomeCamera* cam, int* recurs) {
    if (side <= 0) side = 16;
    uint64_t s = seed;
    TestTex t = make_testtex(b);
    int t_blue = t_matte(b, 0.2, 0.3, 0.8), t_green = t_matte(b, 0.3, 0.7, 0.3);
    std::vector<int32_t> cells;
    const Flt pitch = 3.0;
    for (int i = 0; i < side; i++)
        for (int j = 0; j < side; j++) {
            Flt cx = (i - (side - 1) * 0.5) * pitch + urange(s, -0.2, 0.2);
            Flt cz = (j - (side - 1) * 0.5) * pitch + urange(s, -0.2, 0.2);
            Flt cy = 1.0;
            Vec c = vec(cx, cy, cz);
            std::vector<int32_t> parts;
            parts.push_back(b.box(vec(cx - 1, cy - 1, cz - 1), vec(cx + 1, cy + 1, cz + 1)));
            parts.push_back(b.sphere(c, 1.3));
            parts.push_back(b.cylinder(vec(cx, cy - 1.5, cz), vec(cx, cy + 1.5, cz), urange(s, 1.05, 1.2)));
            int body = b.intersection(parts);
            // drill a cone through the top, then (every third cell) carve a sphere out of a corner
            int carved = b.difference(body, b.cone(vec(cx, cy + 1.2, cz), 0.7, vec(cx, cy - 0.4, cz), 0.1));
            if ((i + 2 * j) % 3 == 0) carved = b.difference(carved, b.sphere(vec(cx + 1, cy + 1, cz + 1), 0.8));
            int tex;
            switch ((i + j) % 4) {
                case 0: tex = t.t_mirror; break;
                case 1: tex = t_blue; break;
                case 2: tex = t.t_shiny_white; break;
                default: tex = t_green; break;
            }
            cells.push_back(b.tag(b.tex(carved, tex), 1000 + i * (int)side + j));
        }
    Flt ext = side * pitch * 0.5 + 4;
    cells.push_back(b.tex(b.box(vec(-ext, -0.5, -ext), vec(ext, 0, ext)), t.t_stripe));
    b.light(vec(-40, 60, 50), 5200, 5000, 4600);
    b.light(vec(30, 25, -35), 900, 1100, 1500);
    int root = b.bih(cells);
    make_camera(vec(ext * 0.9, ext * 0.55, ext * 1.1), vec(0, 0.5, 0), vec(0, 1, 0), 45, cam);
    *recurs = 5;  // primary + 4 reflected generations (Trace.hs:60, Shader.hs:107-118)
    return root;
}

int config_scene(Builder& b, int config, int64_t n, uint64_t seed, GlomeCamera* cam, int* recurs) {
    switch (config) {
        case 1: return scene_testscene(b, seed ? seed : 42, cam, recurs);
        case 2: return scene_sphere_cloud(b, n, seed ? seed : 2, cam, recurs);
        case 3:
        case 5: return scene_heightfield(b, n, seed ? seed : 3, cam, recurs);
        case 4: return scene_csg_grid(b, n, seed ? seed : 4, cam, recurs);
    }
    throw BuildError("unknown config scene");
}

}  // namespace glome_host
