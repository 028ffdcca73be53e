"""ctypes binding of libglomecuda.so (include/glome_cuda.h).

This is plumbing: every call below lands in the C-ABI that a Haskell `GlomeTrace.CUDA` module
binds with `foreign import ccall`.  There is no CPU fallback: if the shared
library is missing the import fails loudly, and scene upload fails with GLOME_ENODEV when no
CUDA device is present.
"""
import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("GLOME_LIB") or os.path.join(_HERE, "_build", "libglomecuda.so")  # GLOME_LIB: A/B builds
# libglomehost.so: scene construction only (the host mirror of the GlomeTrace constructors, tree builders, NFF reader); no
# CUDA in it.  GLOME_HOST_ONLY=1 makes load() hand out this library alone -- what bench.py's CPU reference arm does, so
# that the arm never maps the CUDA library.
HOST_LIB_PATH = os.path.join(os.path.dirname(LIB_PATH), "libglomehost.so")

GLOME_MAX_STACK = 8

# node types (enum GlomeNodeType)
(VOID, SPHERE, TRIANGLE, TRIANGLENORM, BOX, PLANE, DISC, CYLINDER, CONE, GROUP, INSTANCE, BIH, MESH,
 DIFFERENCE, INTERSECTION, TEX, TAG, NOSHADOW, ONLYSHADOW, BOUND, INNERBOUND) = range(21)
NODE_TYPE_COUNT = 21

CLASS_GENERAL, CLASS_FLAT = 0, 1
MODE_ONE_RAY, MODE_ADAPTIVE_AA, MODE_ADAPTIVE_AA_STRICT = 0, 1, 2
OPT_SEG_CONCURRENT, OPT_AA_SPECULATE = 1, 2

OK, EINVAL, ECUDA, ENODEV, ELIMIT, EBUILD = 0, -1, -2, -3, -4, -5
HITFLAG_STACK_OVERFLOW, HITFLAG_CSG_OVERFLOW = 1, 2


class GlomeNode(C.Structure):
    _fields_ = [("type", C.c_int32), ("a", C.c_int32), ("b", C.c_int32), ("c", C.c_int32)]


class GlomeBihNode(C.Structure):
    _fields_ = [("lsplit", C.c_double), ("rsplit", C.c_double), ("axis", C.c_int32), ("left", C.c_int32),
                ("right", C.c_int32), ("pad", C.c_int32)]


class GlomeBvhNode(C.Structure):
    _fields_ = [("lbb", C.c_double * 6), ("rbb", C.c_double * 6), ("left", C.c_int32), ("right", C.c_int32),
                ("pad", C.c_int32 * 6)]


class GlomeMaterial(C.Structure):
    _fields_ = [("kind", C.c_int32), ("a", C.c_int32), ("b", C.c_int32), ("c", C.c_int32), ("d", C.c_int32),
                ("pad", C.c_int32 * 3), ("p", C.c_double * 8)]


class GlomeTexture(C.Structure):
    _fields_ = [("kind", C.c_int32), ("a", C.c_int32), ("b", C.c_int32), ("c", C.c_int32), ("p", C.c_double * 4)]


class GlomeLight(C.Structure):
    _fields_ = [("pos", C.c_double * 3), ("color", C.c_double * 3), ("rad", C.c_double), ("falloff", C.c_int32),
                ("do_shadow", C.c_int32)]


class GlomeFlatScene(C.Structure):
    _fields_ = [("version", C.c_int32), ("root", C.c_int32), ("n_nodes", C.c_int32), ("n_bihnodes", C.c_int32),
                ("n_bvhnodes", C.c_int32), ("n_ipool", C.c_int32), ("n_textures", C.c_int32),
                ("n_materials", C.c_int32), ("n_lights", C.c_int32), ("n_lightsets", C.c_int32),
                ("n_dpool", C.c_int64),
                ("nodes", C.POINTER(GlomeNode)), ("bihnodes", C.POINTER(GlomeBihNode)),
                ("bvhnodes", C.POINTER(GlomeBvhNode)), ("ipool", C.POINTER(C.c_int32)),
                ("dpool", C.POINTER(C.c_double)), ("textures", C.POINTER(GlomeTexture)),
                ("materials", C.POINTER(GlomeMaterial)), ("lights", C.POINTER(GlomeLight)),
                ("lightsets", C.POINTER(C.c_int32)), ("max_depth", C.c_int32), ("scene_class", C.c_int32)]


class GlomeHit(C.Structure):
    _fields_ = [("t", C.c_double), ("pos", C.c_double * 3), ("norm", C.c_double * 3), ("hit", C.c_int32),
                ("prim", C.c_int32), ("sub", C.c_int32), ("ntex", C.c_int32), ("ntag", C.c_int32),
                ("flags", C.c_int32), ("tex", C.c_int32 * GLOME_MAX_STACK), ("tag", C.c_int32 * GLOME_MAX_STACK)]


class GlomeCamera(C.Structure):
    _fields_ = [("pos", C.c_double * 3), ("fwd", C.c_double * 3), ("up", C.c_double * 3), ("right", C.c_double * 3)]


class GlomeRenderOpts(C.Structure):
    _fields_ = [("mode", C.c_int32), ("blocksize", C.c_int32), ("recurs", C.c_int32), ("tint_depth", C.c_int32),
                ("thresholds", C.c_double * 4), ("tile_first", C.c_int32), ("tile_stride", C.c_int32),
                ("want_rgb8", C.c_int32), ("debug_heatmap", C.c_int32)]


class GlomeRenderStats(C.Structure):
    _fields_ = [("rays_primary", C.c_int64), ("rays_shadow", C.c_int64), ("rays_secondary", C.c_int64),
                ("overflow_rays", C.c_int64), ("perlin_range", C.c_int64), ("kernel_ms", C.c_double),
                ("launches", C.c_int32), ("reserved", C.c_int32), ("visits_bih", C.c_int64), ("tests_prim", C.c_int64),
                ("visits_bvh", C.c_int64), ("tests_tri", C.c_int64), ("traverse_ms", C.c_double),
                ("traverse_launches", C.c_int64), ("family_ms", C.c_double * 4), ("family_launches", C.c_int64 * 4),
                ("visits_instance", C.c_int64), ("csg_steps", C.c_int64)]


assert C.sizeof(GlomeHit) == 144 and C.sizeof(GlomeBihNode) == 32 and C.sizeof(GlomeBvhNode) == 128
assert C.sizeof(GlomeMaterial) == 96 and C.sizeof(GlomeTexture) == 48 and C.sizeof(GlomeLight) == 64

_P = C.POINTER
_dp, _ip, _vp = _P(C.c_double), _P(C.c_int32), C.c_void_p

# name -> (restype, argtypes): every symbol include/glome_cuda.h declares
SIGNATURES = {
    "glome_last_error": (C.c_char_p, []),
    "glome_device_count": (C.c_int, []),
    "glome_scene_create": (C.c_int, [_P(GlomeFlatScene), C.c_int, _P(_vp)]),
    "glome_scene_create_f32": (C.c_int, [_P(GlomeFlatScene), C.c_int, _P(_vp)]),
    "glome_scene_set_option": (C.c_int, [_vp, C.c_int, C.c_int]),
    "glome_build_release_cache": (C.c_int, [C.c_int]),
    "glome_scene_destroy": (C.c_int, [_vp]),
    "glome_rayint_batch": (C.c_int, [_vp, C.c_int64, _vp, _vp, C.c_int, _vp]),
    "glome_shadow_batch": (C.c_int, [_vp, C.c_int64, _vp, _vp, C.c_int, _vp]),
    "glome_inside_batch": (C.c_int, [_vp, C.c_int64, _vp, _vp]),
    "glome_trace_batch": (C.c_int, [_vp, C.c_int64, _vp, _vp, C.c_int, C.c_int, _vp, _vp, _vp]),
    "glome_render": (C.c_int, [_vp, _P(GlomeCamera), C.c_int, C.c_int, _P(GlomeRenderOpts), _vp, _vp,
                               _P(GlomeRenderStats)]),
    "glome_render_dev": (C.c_int, [_vp, _P(GlomeCamera), C.c_int, C.c_int, _P(GlomeRenderOpts), _vp, _vp,
                                   _P(GlomeRenderStats), _vp]),
    "glome_scene_launches": (C.c_int64, [_vp]),
    "glome_multi_create": (C.c_int, [_P(GlomeFlatScene), C.c_int, _ip, _P(_vp)]),
    "glome_multi_destroy": (C.c_int, [_vp]),
    "glome_multi_render": (C.c_int, [_vp, _P(GlomeCamera), C.c_int, C.c_int, _P(GlomeRenderOpts), _vp, _vp, _P(GlomeRenderStats)]),
    "glome_debug_count_batch": (C.c_int, [_vp, C.c_int64, _vp, _vp, C.c_int, _vp]),
    "glome_get_tags": (C.c_int, [_vp, _P(GlomeCamera), C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, _ip, C.c_int, _ip, _ip,
                                 _P(GlomeHit)]),
    "glome_dev_alloc": (C.c_int, [C.c_int, C.c_int64, _P(_vp)]),
    "glome_dev_free": (C.c_int, [C.c_int, _vp]),
    "glome_render_opts_default": (None, [_P(GlomeRenderOpts)]),
    "glome_tile_slots": (C.c_int, [C.c_int, C.c_int, C.c_int, C.c_int]),
    "glome_tiles_pack_dev": (C.c_int, [C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, _vp, _vp, _vp]),
    "glome_tiles_unpack_dev": (C.c_int, [C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, _vp, _vp, _vp]),
    "glome_tiles_unpack_all_dev": (C.c_int, [C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, _vp, _vp, _vp]),
    "glome_tile_count": (C.c_int, [C.c_int, C.c_int, C.c_int]),
    "glome_tile_rect": (C.c_int, [C.c_int, C.c_int, C.c_int, C.c_int, _ip]),
    "glome_builder_create": (C.c_int, [_P(_vp)]),
    "glome_builder_destroy": (C.c_int, [_vp]),
    "glome_builder_set_build_device": (C.c_int, [_vp, C.c_int]),
    "glome_builder_last_build_ms": (C.c_int, [_vp, _dp]),
    "glome_sb_void": (C.c_int, [_vp]),
    "glome_sb_sphere": (C.c_int, [_vp, _dp, C.c_double]),
    "glome_sb_spheres": (C.c_int, [_vp, C.c_int64, _vp, _vp, _vp]),
    "glome_sb_triangle": (C.c_int, [_vp, _dp]),
    "glome_sb_trianglenorm": (C.c_int, [_vp, _dp]),
    "glome_sb_box": (C.c_int, [_vp, _dp, _dp]),
    "glome_sb_plane": (C.c_int, [_vp, _dp, _dp]),
    "glome_sb_plane_offset": (C.c_int, [_vp, _dp, C.c_double]),
    "glome_sb_disc": (C.c_int, [_vp, _dp, _dp, C.c_double]),
    "glome_sb_cylinder": (C.c_int, [_vp, _dp, _dp, C.c_double]),
    "glome_sb_cone": (C.c_int, [_vp, _dp, C.c_double, _dp, C.c_double]),
    "glome_sb_cylinder_z": (C.c_int, [_vp, C.c_double, C.c_double, C.c_double]),
    "glome_sb_cone_z": (C.c_int, [_vp, C.c_double, C.c_double, C.c_double, C.c_double]),
    "glome_sb_group": (C.c_int, [_vp, C.c_int, _vp]),
    "glome_sb_bih": (C.c_int, [_vp, C.c_int64, _vp]),
    "glome_sb_mesh": (C.c_int, [_vp, C.c_int64, _vp, C.c_int64, _vp, C.c_int64, _vp, C.c_int, _vp, C.c_int, _vp]),
    "glome_sb_list": (C.c_int, [_vp, C.c_int, _vp]),
    "glome_sb_instance": (C.c_int, [_vp, C.c_int, _dp]),
    "glome_sb_disc_raw": (C.c_int, [_vp, _dp, _dp, C.c_double]),
    "glome_sb_difference_ex": (C.c_int, [_vp, C.c_int, C.c_int, C.c_int]),
    "glome_sb_bih_prebuilt": (C.c_int, [_vp, C.c_int64, _vp, C.c_int64, _vp, _vp, _vp]),
    "glome_sb_mesh_prebuilt": (C.c_int, [_vp, C.c_int64, _vp, C.c_int64, _vp, C.c_int64, _vp, C.c_int, _vp, C.c_int, _vp,
                                         C.c_int64, _vp, _vp, C.c_int64, _vp, _vp]),
    "glome_sb_difference": (C.c_int, [_vp, C.c_int, C.c_int]),
    "glome_sb_intersection": (C.c_int, [_vp, C.c_int, _vp]),
    "glome_sb_tex": (C.c_int, [_vp, C.c_int, C.c_int]),
    "glome_sb_tag": (C.c_int, [_vp, C.c_int, C.c_int]),
    "glome_sb_noshadow": (C.c_int, [_vp, C.c_int]),
    "glome_sb_onlyshadow": (C.c_int, [_vp, C.c_int]),
    "glome_sb_bound_object": (C.c_int, [_vp, C.c_int, C.c_int]),
    "glome_sb_innerbound": (C.c_int, [_vp, C.c_int, C.c_int]),
    "glome_sb_transform": (C.c_int, [_vp, C.c_int, C.c_int, _vp]),
    "glome_xfm_translate": (C.c_int, [_dp, _dp]),
    "glome_xfm_scale": (C.c_int, [_dp, _dp]),
    "glome_xfm_rotate": (C.c_int, [_dp, C.c_double, _dp]),
    "glome_xfm_compose": (C.c_int, [C.c_int, _vp, _dp]),
    "glome_sb_flatten_transform_bih": (C.c_int, [_vp, C.c_int]),
    "glome_sb_bound": (C.c_int, [_vp, C.c_int, _dp]),
    "glome_sb_mat_surface": (C.c_int, [_vp, _dp, C.c_double, C.c_double, C.c_double, C.c_double, C.c_double]),
    "glome_sb_mat_reflect": (C.c_int, [_vp, C.c_double]),
    "glome_sb_mat_refract": (C.c_int, [_vp, C.c_double, C.c_double, C.c_double]),
    "glome_sb_mat_warp": (C.c_int, [_vp, C.c_int, C.c_int, C.c_int, _dp]),
    "glome_sb_mat_additive": (C.c_int, [_vp, C.c_int, _vp]),
    "glome_sb_mat_blend": (C.c_int, [_vp, C.c_int, C.c_int, C.c_double]),
    "glome_sb_tex_uniform": (C.c_int, [_vp, C.c_int]),
    "glome_sb_tex_stripe_blend": (C.c_int, [_vp, C.c_int, C.c_int, _dp]),
    "glome_sb_tex_perlin_blend": (C.c_int, [_vp, C.c_int, C.c_int, C.c_double]),
    "glome_sb_light": (C.c_int, [_vp, _dp, _dp]),
    "glome_sb_lightset": (C.c_int, [_vp, C.c_int, _vp]),
    "glome_sb_mat_warp_set_scene": (C.c_int, [_vp, C.c_int, C.c_int]),
    "glome_camera": (C.c_int, [_dp, _dp, _dp, C.c_double, _P(GlomeCamera)]),
    "glome_sb_flatten": (C.c_int, [_vp, C.c_int, _P(GlomeFlatScene)]),
    "glome_sb_config_scene": (C.c_int, [_vp, C.c_int, C.c_int64, C.c_uint64, _P(GlomeCamera), _ip]),
    "glome_stdgen_probe": (C.c_int, [C.c_int64, _P(C.c_uint64), _dp, _P(C.c_uint64)]),
    "glome_sb_load_nff": (C.c_int, [_vp, C.c_char_p, C.c_int64, _P(GlomeCamera), _dp, _P(C.c_int64)]),
    "glome_bih_build": (C.c_int, [C.c_int64, _vp, _P(_P(GlomeBihNode)), _ip, _P(_ip), _ip, _P(_ip), _ip, _dp]),
    "glome_bih_build_gpu": (C.c_int, [C.c_int64, _vp, C.c_int, _P(_P(GlomeBihNode)), _ip, _P(_ip), _ip, _P(_ip), _ip, _dp, _dp]),
    "glome_mesh_build": (C.c_int, [C.c_int64, _vp, C.c_int64, _vp, _P(_P(GlomeBvhNode)), _ip, _P(_ip), _ip,
                                   _P(_ip), _ip, _ip, _dp]),
    "glome_mesh_build_gpu": (C.c_int, [C.c_int64, _vp, C.c_int64, _vp, C.c_int, _P(_P(GlomeBvhNode)), _ip, _P(_ip), _ip,
                                       _P(_ip), _ip, _ip, _dp, _dp]),
    "glome_free": (None, [_vp]),
}

_lib = None


class GlomeError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__("glome error %d: %s" % (code, msg))
        self.code = code


def load():
    """Load libglomecuda.so (built in-tree by __graft_entry__.build / make -C glome_b200/csrc)."""
    global _lib
    if _lib is not None:
        return _lib
    host_only = os.environ.get("GLOME_HOST_ONLY") == "1"
    path = HOST_LIB_PATH if host_only else LIB_PATH
    if not os.path.exists(path):
        raise ImportError("%s not built: run `python -c 'import __graft_entry__ as g; g.build()'` "
                          "(expected at %s); there is no CPU fallback" % (os.path.basename(path), path))
    lib = C.CDLL(path)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name, None)
        if fn is None:
            if host_only:
                continue  # a device entry point: absent from the host library by construction
            raise ImportError("libglomecuda.so does not export %s" % name)
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


def check(rc):
    if rc < 0:
        raise GlomeError(rc, load().glome_last_error().decode("utf-8", "replace"))
    return rc
