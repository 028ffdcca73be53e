// glome_build.h -- construction of the BIH on the GPU.
#pragma once
#include "host_builder.h"

namespace glome_host {

// bih (Bih.hs:211-324) built level by level on `device`.  The result is the tree bih_build() makes, array
// for array (same pre-order node numbering, same leaf list, same leaf-ordered permutation), so every id the
// renderer reports stays what the reference's list order implies.
// timings_ms (optional, 3 doubles): host->device copy, device build, device->host copy.
// Throws BuildError (infinite bounding box, Bih.hs:319-322; recursion deeper than bih_build allows; no device).
void bih_build_gpu(int64_t n, const double* bboxes, int device, BihTree& out, double* timings_ms);

// mesh's BVH (Mesh.hs:50-134) the same way; identical to mesh_build()'s MeshTree.
void mesh_build_gpu(int64_t nverts, const double* verts, int64_t ntris, const int32_t* tris, int device, MeshTree& out,
                    double* timings_ms);

}  // namespace glome_host
