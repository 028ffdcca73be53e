// glome_cuda_f32.cu -- the FP32 twin of libglomecuda's kernels and scene-bound entry points.
//
// GlomeVec's `type Flt = Double` carries the note "make separate Float and Double instances of this library"
// (Vec.hs:7-9) and BASELINE.json's north_star asks for "an optional FP32 mode reported separately".  This translation
// unit is that Float instance: the same source (glome_cuda.cu and the headers it includes), compiled with
// Flt = float, FP32 device payloads (16-byte BIH nodes, 64-byte BVH nodes, 16-byte spheres) and `_f32` entry names.
// Rays, cameras, frames and hit records cross the C-ABI as doubles in both modes.
#define GLOME_F32 1
#include "glome_cuda.cu"
