"""ctypes mirror of include/rt_b200.h (struct layouts and prototypes only — no arithmetic).

Each Structure matches the C struct of the same name field for field; tests/test_abi.py checks
sizes/offsets against the compiled library's view where that is observable.
"""
import ctypes as C

RT_B200_ABI_VERSION = 5

RT_OK = 0
RT_ERR_INVALID_ARGUMENT = -1
RT_ERR_CUDA = -2
RT_ERR_NO_DEVICE = -3
RT_ERR_OUT_OF_MEMORY = -4
RT_ERR_UNSUPPORTED = -5
RT_ERR_INTERNAL = -6

RT_MAT_LAMBERTIAN, RT_MAT_METAL, RT_MAT_DIELECTRIC, RT_MAT_DIFFUSE_LIGHT = 0, 1, 2, 3
RT_TEX_SOLID, RT_TEX_CHECKER, RT_TEX_IMAGE, RT_TEX_NOISE = 0, 1, 2, 3

RT_FLAG_NONE = 0
RT_FLAG_COUNT_WORK = 1
RT_FLAG_TILE_SPLIT = 2

_f3 = C.c_float * 3


class rt_sphere(C.Structure):
    _fields_ = [("cx", C.c_float), ("cy", C.c_float), ("cz", C.c_float), ("r", C.c_float),
                ("material", C.c_uint32)]


class rt_quad(C.Structure):
    _fields_ = [("q", _f3), ("u", _f3), ("v", _f3), ("material", C.c_uint32)]


class rt_material(C.Structure):
    _fields_ = [("kind", C.c_uint32), ("albedo", _f3), ("fuzz", C.c_float), ("ior", C.c_float),
                ("texture", C.c_uint32)]


class rt_texture(C.Structure):
    _fields_ = [("kind", C.c_uint32), ("a", _f3), ("b", _f3), ("scale", C.c_float),
                ("image", C.c_uint32), ("oob", _f3)]


class rt_image(C.Structure):
    _fields_ = [("w", C.c_int32), ("h", C.c_int32), ("rgb16", C.POINTER(C.c_uint16))]


class rt_perlin(C.Structure):
    _fields_ = [("vec", (C.c_float * 3) * 256), ("perm_x", C.c_uint8 * 256), ("perm_y", C.c_uint8 * 256),
                ("perm_z", C.c_uint8 * 256)]


class rt_scene_desc(C.Structure):
    _fields_ = [("abi_version", C.c_uint32), ("reserved", C.c_uint32),
                ("spheres", C.POINTER(rt_sphere)), ("n_spheres", C.c_uint64),
                ("materials", C.POINTER(rt_material)), ("n_materials", C.c_uint32),
                ("textures", C.POINTER(rt_texture)), ("n_textures", C.c_uint32),
                ("images", C.POINTER(rt_image)), ("n_images", C.c_uint32),
                ("ray_origin_radius", C.c_float),
                ("quads", C.POINTER(rt_quad)), ("n_quads", C.c_uint64),
                ("sphere_ids", C.POINTER(C.c_uint32)), ("quad_ids", C.POINTER(C.c_uint32)),
                ("perlins", C.POINTER(rt_perlin)), ("n_perlins", C.c_uint32), ("reserved2", C.c_uint32)]


class rt_camera(C.Structure):
    _fields_ = [("width", C.c_int32), ("height", C.c_int32), ("spp", C.c_int32),
                ("max_depth", C.c_int32), ("center", _f3), ("pixel00", _f3), ("pixel_du", _f3),
                ("pixel_dv", _f3), ("defocus_u", _f3), ("defocus_v", _f3),
                ("defocus_angle", C.c_float), ("background", _f3)]


class rt_camera_options(C.Structure):
    _fields_ = [("aspect_ratio", C.c_float), ("image_width", C.c_int32), ("spp", C.c_int32),
                ("max_depth", C.c_int32), ("fov_radians", C.c_float),
                ("defocus_angle_radians", C.c_float), ("focus_dist", C.c_float),
                ("look_from", _f3), ("look_at", _f3), ("vup", _f3), ("background", _f3)]


class rt_render_opts(C.Structure):
    _fields_ = [("seed", C.c_uint64), ("device", C.c_int32), ("sample_offset", C.c_int32),
                ("sample_count", C.c_int32), ("flags", C.c_int32), ("row_begin", C.c_int32),
                ("row_count", C.c_int32), ("row_step", C.c_int32), ("reserved", C.c_int32)]


class rt_stats(C.Structure):
    _fields_ = [("samples", C.c_uint64), ("rays", C.c_uint64), ("box_tests", C.c_uint64),
                ("sphere_tests", C.c_uint64), ("hits", C.c_uint64), ("ms_render", C.c_float),
                ("ms_total", C.c_float), ("kernel_launches", C.c_uint32),
                ("megakernel_launches", C.c_uint32), ("ms_megakernel", C.c_float), ("reserved", C.c_uint32),
                ("survivors", C.c_uint64), ("work_bytes", C.c_uint64)]


class rt_bvh_info(C.Structure):
    _fields_ = [("n_nodes", C.c_uint64), ("n_slots", C.c_uint64), ("max_depth", C.c_uint32),
                ("in_shared_memory", C.c_uint32), ("box_pad_min", C.c_float),
                ("box_pad_max", C.c_float), ("root_ref", C.c_uint32), ("built_on_device", C.c_uint32)]


# name -> (restype, argtypes); every symbol include/rt_b200.h declares.
PROTOTYPES = {
    "rt_last_error": (C.c_char_p, []),
    "rt_abi_version": (C.c_int, []),
    "rt_device_count": (C.c_int, []),
    "rt_scene_create": (C.c_int, [C.POINTER(rt_scene_desc), C.c_int, C.POINTER(C.c_void_p)]),
    "rt_scene_destroy": (None, [C.c_void_p]),
    "rt_scene_set_stream": (C.c_int, [C.c_void_p, C.c_void_p]),
    "rt_workspace_release": (None, [C.c_int]),
    "rt_render": (C.c_int, [C.c_void_p, C.POINTER(rt_camera), C.POINTER(rt_render_opts),
                            C.c_void_p, C.c_void_p, C.POINTER(rt_stats)]),
    "rt_render_multi": (C.c_int, [C.POINTER(rt_scene_desc), C.POINTER(rt_camera), C.POINTER(rt_render_opts),
                                  C.POINTER(C.c_int32), C.c_int32, C.c_void_p, C.c_void_p, C.POINTER(rt_stats)]),
    "rt_render_accum_device": (C.c_int, [C.c_void_p, C.POINTER(rt_camera),
                                         C.POINTER(rt_render_opts), C.c_void_p,
                                         C.POINTER(rt_stats)]),
    "rt_resolve_device": (C.c_int, [C.c_void_p, C.c_int32, C.c_int32, C.c_int32, C.c_int32,
                                    C.c_void_p, C.c_void_p]),
    "rt_trace": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.c_float, C.c_float,
                           C.c_void_p, C.c_void_p]),
    "rt_primary_rays": (C.c_int, [C.POINTER(rt_camera), C.POINTER(rt_render_opts), C.c_int64,
                                  C.c_int64, C.c_void_p, C.c_void_p]),
    "rt_camera_from_options": (C.c_int, [C.POINTER(rt_camera_options), C.POINTER(rt_camera)]),
    "rt_scene_bvh_info": (C.c_int, [C.c_void_p, C.POINTER(rt_bvh_info)]),
    "rt_scene_bvh_copy": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p]),
}


def bind(lib):
    """Attach restype/argtypes for every exported symbol; raises AttributeError if one is missing."""
    for name, (res, args) in PROTOTYPES.items():
        fn = getattr(lib, name)
        fn.restype = res
        fn.argtypes = args
    return lib
