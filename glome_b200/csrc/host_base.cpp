// host_base.cpp -- what libglomehost.so (scene construction, no CUDA) and libglomecuda.so (kernels) share: the
// thread-local error string of the C-ABI, the hooks through which the host builders reach the GPU tree builders when
// libglomecuda.so is loaded, and the pure-host helpers of include/glome_cuda.h (render options, the tile list of
// `chunk`, Glome.hs:371-377).
#include <string.h>

#include <string>

#include "glome_build.h"
#include "host_builder.h"

static thread_local std::string g_err;
std::string& glome_err_ref() { return g_err; }
void glome_set_error(const std::string& s) { g_err = s; }

namespace glome_host {
// set by libglomecuda.so's initialiser (glome_build.cu); null while only libglomehost.so is loaded
void (*hook_bih_build_gpu)(int64_t, const double*, int, BihTree&, double*) = nullptr;
void (*hook_mesh_build_gpu)(int64_t, const double*, int64_t, const int32_t*, int, MeshTree&, double*) = nullptr;

void bih_build_gpu(int64_t n, const double* bboxes, int device, BihTree& out, double* timings_ms) {
    if (!hook_bih_build_gpu) throw BuildError("building a tree on the GPU needs libglomecuda.so (only libglomehost.so is loaded)");
    hook_bih_build_gpu(n, bboxes, device, out, timings_ms);
}
void mesh_build_gpu(int64_t nverts, const double* verts, int64_t ntris, const int32_t* tris, int device, MeshTree& out,
                    double* timings_ms) {
    if (!hook_mesh_build_gpu) throw BuildError("building a tree on the GPU needs libglomecuda.so (only libglomehost.so is loaded)");
    hook_mesh_build_gpu(nverts, verts, ntris, tris, device, out, timings_ms);
}
}  // namespace glome_host

extern "C" {

const char* glome_last_error(void) { return g_err.c_str(); }

void glome_render_opts_default(GlomeRenderOpts* o) {
    memset(o, 0, sizeof(*o));
    o->mode = GLOME_MODE_ADAPTIVE_AA;  // the live path of the reference (Glome.hs:385)
    o->blocksize = 65;                 // Glome.hs:116
    o->recurs = 3;                     // Glome.hs:25
    o->tint_depth = 0;
    o->thresholds[0] = 0.14; o->thresholds[1] = 0.15; o->thresholds[2] = 0.16; o->thresholds[3] = 0.18;  // Glome.hs:221-224
    o->tile_first = 0; o->tile_stride = 1; o->want_rgb8 = 0; o->debug_heatmap = 0;
}

// chunk (Glome.hs:371-377): tiles start at k*bs, the last one of a row / column is partial
int glome_tile_count(int width, int height, int blocksize) {
    if (width <= 0 || height <= 0 || blocksize <= 0) return 0;
    return ((width + blocksize - 1) / blocksize) * ((height + blocksize - 1) / blocksize);
}
int glome_tile_rect(int width, int height, int blocksize, int i, int32_t rect[4]) {
    if (width <= 0 || height <= 0 || blocksize <= 0) { g_err = "bad geometry"; return GLOME_EINVAL; }
    const int nty = (height + blocksize - 1) / blocksize;
    if (i < 0 || i >= glome_tile_count(width, height, blocksize)) { g_err = "tile index out of range"; return GLOME_EINVAL; }
    int tx = i / nty, ty = i % nty;  // renderTiles enumerates x chunks outer, y chunks inner (Glome.hs:384)
    rect[0] = tx * blocksize; rect[1] = ty * blocksize;
    rect[2] = (blocksize < width - rect[0]) ? blocksize : width - rect[0];
    rect[3] = (blocksize < height - rect[1]) ? blocksize : height - rect[1];
    return GLOME_OK;
}
int glome_tile_slots(int width, int height, int blocksize, int tile_stride) {
    int n = glome_tile_count(width, height, blocksize);
    if (tile_stride <= 0) return 0;
    return (n + tile_stride - 1) / tile_stride;
}

}  // extern "C"
