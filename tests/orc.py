"""ctypes binding of the CPU oracle (oracle/liblum_oracle.so). TEST INFRASTRUCTURE ONLY."""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ORACLE_DIR = os.path.join(ROOT, "oracle")
LIB = os.path.join(ORACLE_DIR, "liblum_oracle.so")


class Vec3(C.Structure):
    _fields_ = [("x", C.c_float), ("y", C.c_float), ("z", C.c_float)]


class RGB(C.Structure):
    _fields_ = [("r", C.c_float), ("g", C.c_float), ("b", C.c_float)]


class Quat(C.Structure):
    _fields_ = [("x", C.c_float), ("y", C.c_float), ("z", C.c_float), ("w", C.c_float)]


class Quat16(C.Structure):
    _fields_ = [("x", C.c_uint16), ("y", C.c_uint16), ("z", C.c_uint16), ("w", C.c_uint16)]


class PathID(C.Structure):
    _fields_ = [("x", C.c_uint16), ("y", C.c_uint16), ("z", C.c_uint16)]


class Uint2(C.Structure):
    _fields_ = [("x", C.c_uint32), ("y", C.c_uint32)]


class Float2(C.Structure):
    _fields_ = [("x", C.c_float), ("y", C.c_float)]


class Mesh(C.Structure):
    _fields_ = [("num_tris", C.c_uint32), ("vertex", C.POINTER(C.c_float)), ("normal", C.POINTER(C.c_float)), ("uv", C.POINTER(C.c_float)),
                ("material", C.POINTER(C.c_uint16))]


class Transform(C.Structure):
    _fields_ = [("translation", Vec3), ("scale", Vec3), ("rotation", Quat16)]


class Instance(C.Structure):
    _fields_ = [("mesh_id", C.c_uint32), ("transform", Transform)]


class MaterialPacked(C.Structure):
    _fields_ = [("flags", C.c_uint8), ("roughness_clamp", C.c_uint8), ("metallic_tex", C.c_uint16), ("roughness", C.c_uint16),
                ("refraction_index", C.c_uint16), ("albedo_r", C.c_uint16), ("albedo_g", C.c_uint16), ("albedo_b", C.c_uint16),
                ("albedo_a", C.c_uint16), ("emission_r", C.c_uint16), ("emission_g", C.c_uint16), ("emission_b", C.c_uint16),
                ("emission_scale", C.c_uint16), ("albedo_tex", C.c_uint16), ("luminance_tex", C.c_uint16), ("roughness_tex", C.c_uint16),
                ("normal_tex", C.c_uint16)]


class MaterialDesc(C.Structure):
    _fields_ = [("base_substrate", C.c_uint32), ("albedo", C.c_float * 4), ("emission", C.c_float * 3), ("emission_scale", C.c_float),
                ("roughness", C.c_float), ("roughness_clamp", C.c_float), ("refraction_index", C.c_float), ("emission_active", C.c_bool),
                ("thin_walled", C.c_bool), ("metallic", C.c_bool), ("colored_transparency", C.c_bool), ("roughness_as_smoothness", C.c_bool),
                ("normal_map_is_compressed", C.c_bool), ("bidirectional_emission", C.c_bool)]


class Texture(C.Structure):
    _fields_ = [("width", C.c_uint32), ("height", C.c_uint32), ("pitch", C.c_uint32), ("type", C.c_uint32), ("num_components", C.c_uint32),
                ("wrap_u", C.c_uint32), ("wrap_v", C.c_uint32), ("filter", C.c_uint32), ("gamma", C.c_float), ("data", C.c_void_p)]


_TEX_TYPES = {np.dtype(np.float32): 0, np.dtype(np.uint8): 1, np.dtype(np.uint16): 2}


def make_texture(t: dict, keep: list) -> Texture:
    s = Texture()
    s.wrap_u, s.wrap_v = int(t.get("wrap_u", 0)), int(t.get("wrap_v", 0))
    s.filter = int(t.get("filter", 1))
    s.gamma = float(t.get("gamma", 1.0))
    data = t.get("data")
    if data is None:
        s.data = None
        return s
    a = np.ascontiguousarray(data)
    if a.ndim == 2:
        a = a[:, :, None]
    keep.append(a)
    s.height, s.width, s.num_components = a.shape
    s.type = _TEX_TYPES[a.dtype]
    s.pitch = a.strides[0]
    s.data = a.ctypes.data
    return s


def light_intensities(osc, mesh_ids, tri_ids) -> np.ndarray:
    """orc_light_intensity per (mesh, triangle): CPU restatement of the light_compute_intensity kernel."""
    L = lib()
    L.orc_light_intensity.argtypes = [C.c_void_p, C.c_uint32, C.c_uint32]
    L.orc_light_intensity.restype = C.c_float
    return np.array([L.orc_light_intensity(osc.handle, int(m), int(t)) for m, t in zip(mesh_ids, tri_ids)], np.float32)


def texture_mip_chain(t: dict):
    """[level 0, level 1, ...] arrays of a 4-component texture: orc_texture_next_mip applied floor(log2(min(w, h))) - 1 times."""
    L = lib()
    L.orc_texture_next_mip.argtypes = [C.POINTER(Texture), C.c_void_p]
    L.orc_texture_next_mip.restype = None
    levels = [np.ascontiguousarray(t["data"])]
    n = int(np.floor(np.log2(min(levels[0].shape[:2]))))
    for _ in range(max(n, 1) - 1):
        cur = dict(t, data=levels[-1])
        keep = []
        tex = make_texture(cur, keep)
        h, w = levels[-1].shape[0] >> 1, levels[-1].shape[1] >> 1
        out = np.zeros((h, w, 4), levels[-1].dtype)
        L.orc_texture_next_mip(C.byref(tex), out.ctypes.data_as(C.c_void_p))
        levels.append(out)
    return levels


def texture_fetch(t: dict, uv: np.ndarray) -> np.ndarray:
    """orc_texture_fetch over an (N, 2) array of (u, v): the CPU restatement of tex2D<float4>."""
    keep = []
    tex = make_texture(t, keep)
    uv = np.ascontiguousarray(uv, np.float32).reshape(-1, 2)
    out = np.empty((uv.shape[0], 4), np.float32)
    L = lib()
    L.orc_texture_fetch.argtypes = [C.POINTER(Texture), C.c_float, C.c_float, C.POINTER(C.c_float)]
    L.orc_texture_fetch.restype = None
    for i in range(uv.shape[0]):
        L.orc_texture_fetch(C.byref(tex), float(uv[i, 0]), float(uv[i, 1]), out[i].ctypes.data_as(C.POINTER(C.c_float)))
    return out


class AdaptiveParams(C.Structure):
    _fields_ = [("max_sampling_rate", C.c_uint32), ("avg_sampling_rate", C.c_uint32), ("update_interval", C.c_uint32), ("exposure", C.c_float),
                ("tonemap", C.c_uint32), ("agx_slope", C.c_float), ("agx_power", C.c_float), ("agx_saturation", C.c_float)]


def adaptive_params(max_sampling_rate=256, avg_sampling_rate=2, update_interval=64, exposure_aware=True, exposure=1.0, tonemap=4,
                    agx=(1.0, 1.0, 1.0)) -> AdaptiveParams:
    return AdaptiveParams(max_sampling_rate, avg_sampling_rate, update_interval, exposure if exposure_aware else 0.0, tonemap, agx[0], agx[1], agx[2])


def resolve(planes, width, height, words=None, executions=(0, 0, 0, 0, 0), uniform_count=1, mode=0, local_error_minimization=False, stage=0,
            params=None) -> np.ndarray:
    """orc_resolve: accumulation_generate_result in every output mode -> (3, h, w)."""
    L = lib()
    out = np.zeros((3, height, width), np.float32)
    ex = (C.c_uint32 * 5)(*[int(e) for e in executions])
    params = params or adaptive_params()
    w = None if words is None else np.ascontiguousarray(words, np.uint32).reshape(-1)
    L.orc_resolve.restype = None
    L.orc_resolve.argtypes = [C.POINTER(C.c_float), C.c_uint32, C.c_uint32, C.POINTER(C.c_uint32), C.POINTER(C.c_uint32), C.c_uint32, C.c_uint32, C.c_int,
                              C.c_uint32, C.POINTER(AdaptiveParams), C.POINTER(C.c_float)]
    L.orc_resolve(fptr(np.ascontiguousarray(planes, np.float32).reshape(-1)), width, height, None if w is None else uptr(w), ex, uniform_count, mode,
                  1 if local_error_minimization else 0, stage, C.byref(params), fptr(out))
    return out


def adaptive_stage_counts(planes: np.ndarray, width: int, height: int, words: np.ndarray, executions, stage: int, params: AdaptiveParams):
    """One stage build on given planes: -> (new words, block variance, sum). planes: 4 * w * h floats."""
    L = lib()
    bw, bh = (width + 3) >> 2, (height + 3) >> 2
    pl = np.ascontiguousarray(planes, np.float32).reshape(-1)
    w = np.ascontiguousarray(words, np.uint32).reshape(-1).copy()
    ex = (C.c_uint32 * 5)(*[int(e) for e in executions])
    var = np.zeros(bw * bh, np.float32)
    L.orc_adaptive_block_variance.restype = C.c_float
    L.orc_adaptive_block_variance.argtypes = [C.POINTER(C.c_float), C.c_uint32, C.c_uint32, C.POINTER(C.c_uint32), C.POINTER(C.c_uint32),
                                              C.POINTER(AdaptiveParams), C.POINTER(C.c_float)]
    total = L.orc_adaptive_block_variance(fptr(pl), width, height, uptr(w), ex, C.byref(params), fptr(var))
    L.orc_adaptive_stage_counts.restype = None
    L.orc_adaptive_stage_counts.argtypes = [C.POINTER(C.c_float), C.c_float, C.c_uint32, C.c_uint32, C.POINTER(AdaptiveParams), C.POINTER(C.c_uint32)]
    L.orc_adaptive_stage_counts(fptr(var), total, bw * bh, stage, C.byref(params), uptr(w))
    return w.reshape(bh, bw), var.reshape(bh, bw), float(total)


class Camera(C.Structure):
    _fields_ = [("pos", Vec3), ("rotation", Quat), ("fov", C.c_float), ("aperture_size", C.c_float), ("object_distance", C.c_float),
                ("camera_scale", C.c_float), ("russian_roulette_threshold", C.c_float), ("aperture_shape", C.c_uint32),
                ("aperture_blade_count", C.c_uint32)]


class Settings(C.Structure):
    _fields_ = [("width", C.c_uint32), ("height", C.c_uint32), ("max_ray_depth", C.c_uint32), ("sky_mode", C.c_uint32), ("sky_constant_color", RGB)]


class Hit(C.Structure):
    _fields_ = [("prim", C.c_uint32), ("t", C.c_float), ("u", C.c_float), ("v", C.c_float)]


class LightTree(C.Structure):
    _fields_ = [("root", C.c_void_p), ("nodes", C.c_void_p), ("tri_handle_map", C.POINTER(C.c_uint32)), ("num_lights", C.c_uint32)]


class SkyParams(C.Structure):  # OrcSkyParams
    _fields_ = [("geometry_offset", Vec3)] + [(n, C.c_float) for n in (
        "azimuth", "altitude", "moon_azimuth", "moon_altitude", "moon_tex_offset", "sun_strength", "base_density", "rayleigh_density", "mie_density",
        "ozone_density", "rayleigh_falloff", "mie_falloff", "mie_diameter", "ground_visibility", "ozone_layer_thickness", "multiscattering_factor",
        "stars_intensity")] + [(n, C.c_uint32) for n in ("steps", "ozone_absorption", "stars_count", "stars_seed", "aerial_perspective")]


def sky_params(sky: dict = None) -> SkyParams:
    """OrcSkyParams with the reference's defaults (sky.c:6-42), overridden by the entries of `sky`."""
    p = SkyParams()
    lib().orc_sky_params_default(C.byref(p))
    for k, v in (sky or {}).items():
        if k == "geometry_offset":
            p.geometry_offset = vec3(v)
        elif k not in ("mode", "hdri_dim", "hdri_samples"):
            setattr(p, k, v)
    return p


SKY_TM_SHAPE = (64, 256, 4)  # transmittance LUT: SKY_TM_TEX_HEIGHT x SKY_TM_TEX_WIDTH float4
SKY_MS_SHAPE = (32, 32, 4)   # multiscattering LUT


class RayCounts(C.Structure):
    _fields_ = [("closest_rays", C.c_uint64), ("shadow_rays", C.c_uint64), ("light_enum_rays", C.c_uint64)]


_lib = None
_bluenoise = None


# OrcVertexIn / OrcVertexOut of lum_oracle.h (all members are 2- or 4-byte aligned; sizes are asserted against the library)
_V3 = (np.float32, 3)
VERTEX_IN = np.dtype([("path_id", np.uint16, 3), ("state", np.uint16), ("origin", *_V3), ("ray", *_V3), ("prim", np.uint32), ("t", np.float32),
                      ("record", np.uint32, 2), ("medium_ior", np.uint32)])
VERTEX_OUT = np.dtype([("geo_light_id", np.uint32), ("geo_color", *_V3), ("geo_ray", *_V3), ("geo_dist", np.float32),
                       ("bsdf_weight", *_V3), ("bsdf_ray", *_V3), ("bsdf_root_sum", np.float32), ("bsdf_prob", np.float32),
                       ("amb_color", np.uint32, 2), ("amb_ray", np.uint32, 2), ("amb_valid", np.uint32), ("emission", *_V3),
                       ("bounce_alive", np.uint32), ("bounce_state", np.uint32), ("bounce_origin", *_V3), ("bounce_ray", *_V3),
                       ("bounce_record", np.uint32, 2), ("bounce_medium_ior", np.uint32), ("bounce_weight", *_V3), ("normal", *_V3),
                       ("hit_point", *_V3), ("is_transparent_pass", np.uint32), ("sun_color", np.uint32, 2), ("sun_ray", np.uint32, 2)])
NEE_SLOTS = 4  # ORC_NEE_SLOTS: light-tree light, BSDF-sampled light, ambient, sun


# OrcNeeSegment of lum_oracle.h
NEE_SEGMENT = np.dtype([("valid", np.uint32), ("ray", *_V3), ("dist", np.float32), ("color", *_V3), ("target_prim", np.uint32),
                        ("visibility", *_V3), ("enum_hits", np.uint32)])


def build() -> str:
    subprocess.check_call(["make", "-s", "-C", ORACLE_DIR])
    return LIB


def lib() -> C.CDLL:
    global _lib, _bluenoise
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB):
        build()
    L = C.CDLL(LIB)
    L.orc_path_id_get.restype = PathID
    L.orc_path_id_get.argtypes = [C.c_uint32] * 3
    L.orc_path_id_sample.restype = C.c_uint32
    L.orc_path_id_sample.argtypes = [PathID]
    L.orc_path_id_pixel.argtypes = [PathID, C.POINTER(C.c_uint32), C.POINTER(C.c_uint32)]
    L.orc_squares32.restype = C.c_uint32
    L.orc_squares32.argtypes = [C.c_uint32, C.c_uint32]
    L.orc_squares16.restype = C.c_uint16
    L.orc_squares16.argtypes = [C.c_uint32, C.c_uint32]
    L.orc_sobol.restype = Uint2
    L.orc_sobol.argtypes = [C.c_uint32, C.c_uint32]
    L.orc_random_2d_base.restype = Uint2
    L.orc_random_2d_base.argtypes = [C.c_uint32] * 5
    L.orc_u32_to_float.restype = C.c_float
    L.orc_u32_to_float.argtypes = [C.c_uint32]
    L.orc_pack_normal_host.restype = C.c_uint32
    L.orc_pack_normal_host.argtypes = [Vec3]
    L.orc_pack_normal.restype = C.c_uint32
    L.orc_pack_normal.argtypes = [Vec3]
    L.orc_unpack_normal.restype = Vec3
    L.orc_unpack_normal.argtypes = [C.c_uint32]
    L.orc_pack_uv.restype = C.c_uint32
    L.orc_pack_uv.argtypes = [C.c_float, C.c_float]
    L.orc_unpack_uv.restype = Float2
    L.orc_unpack_uv.argtypes = [C.c_uint32]
    L.orc_record_pack.restype = Uint2
    L.orc_record_pack.argtypes = [RGB]
    L.orc_record_unpack.restype = RGB
    L.orc_record_unpack.argtypes = [Uint2]
    L.orc_ray_pack.restype = Uint2
    L.orc_ray_pack.argtypes = [Vec3]
    L.orc_ray_unpack.restype = Vec3
    L.orc_ray_unpack.argtypes = [Uint2]
    L.orc_ior_compress.restype = C.c_uint32
    L.orc_ior_compress.argtypes = [C.c_float]
    L.orc_ior_decompress.restype = C.c_float
    L.orc_ior_decompress.argtypes = [C.c_uint32]
    L.orc_euler_to_quat.restype = Quat
    L.orc_euler_to_quat.argtypes = [Vec3]
    L.orc_quat_pack.restype = Quat16
    L.orc_quat_pack.argtypes = [Quat]
    L.orc_transform_apply.restype = Vec3
    L.orc_transform_apply.argtypes = [C.POINTER(Transform), Vec3]
    L.orc_transform_apply_inv.restype = Vec3
    L.orc_transform_apply_inv.argtypes = [C.POINTER(Transform), Vec3]
    L.orc_material_pack.argtypes = [C.POINTER(MaterialDesc), C.POINTER(MaterialPacked)]
    L.orc_camera_sample.argtypes = [C.POINTER(Camera), C.POINTER(Settings), PathID, C.POINTER(Vec3), C.POINTER(Vec3)]
    L.orc_scene_create.restype = C.c_void_p
    L.orc_scene_create.argtypes = [C.POINTER(Mesh), C.c_uint32, C.POINTER(Instance), C.c_uint32, C.POINTER(MaterialPacked), C.c_uint32]
    L.orc_scene_destroy.argtypes = [C.c_void_p]
    L.orc_scene_num_prims.restype = C.c_uint32
    L.orc_scene_num_prims.argtypes = [C.c_void_p]
    L.orc_scene_world_tris.restype = C.POINTER(C.c_float)
    L.orc_scene_world_tris.argtypes = [C.c_void_p]
    L.orc_tri_mt.restype = C.c_float
    L.orc_tri_mt.argtypes = [C.POINTER(C.c_float), Vec3, Vec3, C.POINTER(C.c_float), C.POINTER(C.c_float)]
    L.orc_tri_watertight.restype = C.c_bool
    L.orc_tri_watertight.argtypes = [C.POINTER(C.c_float), Vec3, Vec3, C.POINTER(C.c_float), C.POINTER(C.c_float), C.POINTER(C.c_float)]
    L.orc_closest_hit.restype = Hit
    L.orc_closest_hit.argtypes = [C.c_void_p, Vec3, Vec3, C.c_float, C.c_float, C.c_uint32, C.POINTER(C.c_uint64), C.POINTER(C.c_uint64)]
    L.orc_closest_hit_bruteforce.restype = Hit
    L.orc_closest_hit_bruteforce.argtypes = [C.c_void_p, Vec3, Vec3, C.c_float, C.c_float, C.c_uint32, C.c_int]
    L.orc_trace_primary.restype = C.c_double
    L.orc_trace_primary.argtypes = [C.c_void_p, C.POINTER(Camera), C.POINTER(Settings), C.c_uint32, C.POINTER(C.c_uint32), C.POINTER(C.c_uint32),
                                    C.POINTER(C.c_float), C.POINTER(C.c_float), C.POINTER(C.c_float), C.c_int, C.POINTER(C.c_uint64),
                                    C.POINTER(C.c_uint64)]
    L.orc_trace_rays.restype = C.c_double
    L.orc_trace_rays.argtypes = [C.c_void_p, C.POINTER(C.c_float), C.POINTER(C.c_float), C.c_uint32, C.POINTER(C.c_uint32), C.POINTER(C.c_float),
                                 C.POINTER(C.c_float), C.POINTER(C.c_float), C.c_int]
    for opt in ("orc_scene_set_light_tree", "orc_scene_set_bsdf_luts", "orc_bsdf_lut_generate", "orc_render", "orc_render_region"):
        if not hasattr(L, opt):
            continue
    if hasattr(L, "orc_render"):
        L.orc_scene_set_light_tree.argtypes = [C.c_void_p, C.POINTER(LightTree)]
        L.orc_scene_set_bsdf_luts.argtypes = [C.c_void_p] + [C.POINTER(C.c_uint16)] * 4
        L.orc_bsdf_lut_generate.argtypes = [C.POINTER(C.c_uint16)] * 4 + [C.c_uint32, C.c_int, C.c_int]
        L.orc_bsdf_lut_dielectric_texel.argtypes = [C.c_uint32, C.c_uint32, C.POINTER(C.c_uint16), C.POINTER(C.c_uint16)]
        L.orc_render.restype = C.c_double
        L.orc_render.argtypes = [C.c_void_p, C.POINTER(Camera), C.POINTER(Settings), C.c_uint32, C.c_uint32, C.POINTER(C.c_float), C.c_int,
                                 C.POINTER(RayCounts)]
        L.orc_render_region.restype = C.c_double
        L.orc_render_region.argtypes = [C.c_void_p, C.POINTER(Camera), C.POINTER(Settings), C.c_uint32, C.c_uint32, C.c_uint32, C.c_uint32,
                                        C.c_uint32, C.c_uint32, C.POINTER(C.c_float), C.c_int, C.POINTER(RayCounts)]
        if hasattr(L, "orc_render_debug"):
            L.orc_render_debug.restype = C.c_double
            L.orc_render_debug.argtypes = [C.c_void_p, C.POINTER(Camera), C.POINTER(Settings), C.c_uint32, C.c_uint32, C.c_uint32, C.POINTER(C.c_float),
                                           C.c_int]
        L.orc_path_vertices.argtypes = [C.c_void_p, C.POINTER(Camera), C.POINTER(Settings), C.c_uint32, C.c_uint32, C.c_void_p, C.c_void_p, C.c_int]
        L.orc_shade_vertices.argtypes = [C.c_void_p, C.POINTER(Camera), C.POINTER(Settings), C.c_uint32, C.c_uint32, C.c_void_p, C.c_void_p, C.c_int]
        L.orc_scene_prim_handle.argtypes = [C.c_void_p, C.c_uint32, C.POINTER(C.c_uint32), C.POINTER(C.c_uint32)]
        L.orc_nee_segments.argtypes = [C.c_void_p, C.POINTER(Camera), C.POINTER(Settings), C.c_uint32, C.c_uint32, C.c_void_p, C.c_void_p, C.c_int]
        L.orc_shadow_rays.argtypes = [C.c_void_p, C.c_uint32] + [C.c_void_p] * 6 + [C.c_int]
        L.orc_sizeof_vertex_in.restype = C.c_size_t
        L.orc_sizeof_vertex_out.restype = C.c_size_t
        assert L.orc_sizeof_vertex_in() == VERTEX_IN.itemsize and L.orc_sizeof_vertex_out() == VERTEX_OUT.itemsize
    _bluenoise = np.fromfile(os.path.join(ROOT, "luminary_b200", "data", "bluenoise_2D.bin"), dtype=np.uint32)
    L.orc_set_bluenoise.argtypes = [C.POINTER(C.c_uint32)]
    L.orc_set_bluenoise(_bluenoise.ctypes.data_as(C.POINTER(C.c_uint32)))
    _lib = L
    return L


def vec3(v) -> Vec3:
    return Vec3(float(v[0]), float(v[1]), float(v[2]))


def fptr(a: np.ndarray):
    return a.ctypes.data_as(C.POINTER(C.c_float))


def uptr(a: np.ndarray):
    return a.ctypes.data_as(C.POINTER(C.c_uint32))


def pack_material(m: dict) -> MaterialPacked:
    d = MaterialDesc()
    d.base_substrate = int(m["base_substrate"])
    d.albedo[:] = m["albedo"]
    d.emission[:] = m["emission"]
    d.emission_scale = m["emission_scale"]
    d.roughness = m["roughness"]
    d.roughness_clamp = m["roughness_clamp"]
    d.refraction_index = m["refraction_index"]
    for k in ("emission_active", "thin_walled", "metallic", "colored_transparency", "roughness_as_smoothness", "normal_map_is_compressed",
              "bidirectional_emission"):
        setattr(d, k, bool(m[k]))
    out = MaterialPacked()
    lib().orc_material_pack(C.byref(d), C.byref(out))
    for k in ("albedo_tex", "luminance_tex", "roughness_tex", "metallic_tex", "normal_tex"):  # device_structs.c:322-326
        setattr(out, k, int(m.get(k, 0xFFFF)))
    return out


def make_camera(cam: dict) -> Camera:
    L = lib()
    c = Camera()
    c.pos = vec3(cam["pos"])
    c.rotation = L.orc_euler_to_quat(vec3(cam["rotation"]))
    c.fov = cam["fov"]
    c.aperture_size = cam["aperture_size"]
    c.object_distance = cam["object_distance"]
    c.camera_scale = cam["camera_scale"]
    c.russian_roulette_threshold = cam["russian_roulette_threshold"]
    c.aperture_shape = cam["aperture_shape"]
    c.aperture_blade_count = cam["aperture_blade_count"]
    return c


def make_settings(scene) -> Settings:
    s = Settings()
    s.width, s.height, s.max_ray_depth = scene.width, scene.height, scene.max_ray_depth
    s.sky_mode = scene.sky_mode
    s.sky_constant_color = RGB(*[float(x) for x in scene.sky_color])
    return s


class OracleScene:
    """Owns an OrcScene built from a luminary_b200.scenes.Scene (only active instances are passed on;
    instance ids therefore refer to the list of active instances)."""

    def __init__(self, scene):
        L = lib()
        self.scene = scene
        self._keep = []
        meshes = (Mesh * len(scene.meshes))()
        for i, m in enumerate(scene.meshes):
            v = np.ascontiguousarray(m.vertex, np.float32).reshape(-1)
            n = np.ascontiguousarray(m.normal, np.float32).reshape(-1)
            t = np.ascontiguousarray(m.uv, np.float32).reshape(-1)
            mm = np.ascontiguousarray(m.material, np.uint16).reshape(-1)
            self._keep += [v, n, t, mm]
            meshes[i] = Mesh(m.num_tris, fptr(v), fptr(n), fptr(t), mm.ctypes.data_as(C.POINTER(C.c_uint16)))
        active = [ins for ins in scene.instances if ins.active]
        inst = (Instance * max(len(active), 1))()
        for i, ins in enumerate(active):
            inst[i].mesh_id = ins.mesh_id
            inst[i].transform.translation = vec3(ins.translation)
            inst[i].transform.scale = vec3(ins.scale)
            inst[i].transform.rotation = L.orc_quat_pack(L.orc_euler_to_quat(vec3(ins.rotation)))
        mats = (MaterialPacked * max(len(scene.materials), 1))()
        for i, m in enumerate(scene.materials):
            mats[i] = pack_material(m)
        self.materials_packed = mats
        self.handle = L.orc_scene_create(meshes, len(scene.meshes), inst, len(active), mats, len(scene.materials))
        textures = getattr(scene, "textures", None) or []
        if textures:
            arr = (Texture * len(textures))()
            for i, t in enumerate(textures):
                arr[i] = make_texture(t, self._keep)
            L.orc_scene_set_textures.argtypes = [C.c_void_p, C.POINTER(Texture), C.c_uint32]
            L.orc_scene_set_textures.restype = None
            L.orc_scene_set_textures(self.handle, arr, len(textures))
        self.camera = make_camera(scene.camera)
        self.settings = make_settings(scene)
        if scene.sky_mode in (0, 1) and getattr(scene, "sky", None) is not None:
            self.set_sky(scene.sky)

    def __del__(self):
        try:
            if self.handle:
                lib().orc_scene_destroy(self.handle)
                self.handle = None
        except Exception:
            pass

    def render_adaptive(self, params: AdaptiveParams, executions: int, state=None, threads: int = 0):
        """orc_render_adaptive from `state` (None = fresh): -> state dict(planes (4, h, w), words (bh, bw), executions [5], stage, paths)."""
        L = lib()
        w, h = self.settings.width, self.settings.height
        bw, bh = (w + 3) >> 2, (h + 3) >> 2
        if state is None:
            state = dict(planes=np.zeros((4, h, w), np.float32), words=np.zeros((bh, bw), np.uint32), executions=[0] * 5, stage=0, paths=0,
                         closest_rays=0, shadow_rays=0, light_enum_rays=0)
        planes = np.ascontiguousarray(state["planes"], np.float32)
        words = np.ascontiguousarray(state["words"], np.uint32)
        ex = (C.c_uint32 * 5)(*state["executions"])
        stage = C.c_uint32(state["stage"])
        counts = RayCounts()
        L.orc_render_adaptive.restype = C.c_uint64
        L.orc_render_adaptive.argtypes = [C.c_void_p, C.POINTER(Camera), C.POINTER(Settings), C.POINTER(AdaptiveParams), C.c_uint32, C.POINTER(C.c_float),
                                          C.POINTER(C.c_uint32), C.POINTER(C.c_uint32), C.POINTER(C.c_uint32), C.c_int, C.POINTER(RayCounts)]
        paths = L.orc_render_adaptive(self.handle, C.byref(self.camera), C.byref(self.settings), C.byref(params), executions, fptr(planes), uptr(words),
                                      ex, C.byref(stage), threads, C.byref(counts))
        return dict(planes=planes, words=words, executions=list(ex), stage=stage.value, paths=state["paths"] + int(paths),
                    closest_rays=state["closest_rays"] + counts.closest_rays, shadow_rays=state["shadow_rays"] + counts.shadow_rays,
                    light_enum_rays=state["light_enum_rays"] + counts.light_enum_rays)

    def adaptive_resolve(self, state) -> np.ndarray:
        L = lib()
        w, h = self.settings.width, self.settings.height
        out = np.zeros((3, h, w), np.float32)
        ex = (C.c_uint32 * 5)(*state["executions"])
        L.orc_adaptive_resolve.restype = None
        L.orc_adaptive_resolve.argtypes = [C.POINTER(C.c_float), C.c_uint32, C.c_uint32, C.POINTER(C.c_uint32), C.POINTER(C.c_uint32), C.POINTER(C.c_float)]
        L.orc_adaptive_resolve(fptr(np.ascontiguousarray(state["planes"], np.float32)), w, h, uptr(np.ascontiguousarray(state["words"], np.uint32)), ex,
                               fptr(out))
        return out

    def num_prims(self) -> int:
        return lib().orc_scene_num_prims(self.handle)

    def world_tris(self) -> np.ndarray:
        n = self.num_prims()
        p = lib().orc_scene_world_tris(self.handle)
        return np.ctypeslib.as_array(p, shape=(n, 3, 3)).copy()

    def trace_primary(self, sample_id: int = 0, threads: int = 0):
        n = self.scene.width * self.scene.height
        inst = np.empty(n, np.uint32)
        tri = np.empty(n, np.uint32)
        t = np.empty(n, np.float32)
        u = np.empty(n, np.float32)
        v = np.empty(n, np.float32)
        nv = C.c_uint64(0)
        tt = C.c_uint64(0)
        secs = lib().orc_trace_primary(self.handle, C.byref(self.camera), C.byref(self.settings), sample_id, uptr(inst), uptr(tri), fptr(t), fptr(u),
                                       fptr(v), threads, C.byref(nv), C.byref(tt))
        return dict(instance=inst, tri=tri, t=t, u=u, v=v, seconds=secs, nodes_visited=nv.value, tris_tested=tt.value)

    def trace_rays(self, origins, dirs, threads: int = 0):
        o = np.ascontiguousarray(origins, np.float32).reshape(-1, 3)
        d = np.ascontiguousarray(dirs, np.float32).reshape(-1, 3)
        n = o.shape[0]
        prim = np.empty(n, np.uint32)
        t = np.empty(n, np.float32)
        u = np.empty(n, np.float32)
        v = np.empty(n, np.float32)
        secs = lib().orc_trace_rays(self.handle, fptr(o), fptr(d), n, uptr(prim), fptr(t), fptr(u), fptr(v), threads)
        return dict(prim=prim, t=t, u=u, v=v, seconds=secs)

    def set_light_tree(self, root: bytes, nodes: bytes, handles: np.ndarray):
        self._lt_root = C.create_string_buffer(root, len(root))
        self._lt_nodes = C.create_string_buffer(nodes, max(len(nodes), 1))
        self._lt_handles = np.ascontiguousarray(handles, np.uint32).reshape(-1)
        lt = LightTree(C.cast(self._lt_root, C.c_void_p), C.cast(self._lt_nodes, C.c_void_p), uptr(self._lt_handles), self._lt_handles.size // 2)
        lib().orc_scene_set_light_tree(self.handle, C.byref(lt))

    def set_sky(self, sky: dict = None, enable: bool = True, threads: int = 0):
        """Attaches the procedural sky (reference defaults overridden by `sky`); enable=False removes it."""
        L = lib()
        L.orc_scene_set_sky.argtypes = [C.c_void_p, C.POINTER(SkyParams), C.c_int]
        L.orc_scene_set_sky.restype = None
        if not enable:
            L.orc_scene_set_sky(self.handle, None, threads)
            return
        p = sky_params(sky)
        L.orc_scene_set_sky(self.handle, C.byref(p), threads)

    def sky_luts(self):
        """-> (tm_low, tm_high (64, 256, 4), ms_low, ms_high (32, 32, 4)) copies of the oracle's LUTs"""
        L = lib()
        ptrs = [C.POINTER(C.c_float)() for _ in range(4)]
        L.orc_scene_sky_luts.argtypes = [C.c_void_p] + [C.POINTER(C.POINTER(C.c_float))] * 4
        L.orc_scene_sky_luts.restype = None
        L.orc_scene_sky_luts(self.handle, *[C.byref(q) for q in ptrs])
        shapes = (SKY_TM_SHAPE, SKY_TM_SHAPE, SKY_MS_SHAPE, SKY_MS_SHAPE)
        return tuple(np.ctypeslib.as_array(q, shape=sh).copy() for q, sh in zip(ptrs, shapes))

    def set_sky_luts(self, tm_low, tm_high, ms_low, ms_high):
        L = lib()
        arrs = [np.ascontiguousarray(a, np.float32) for a in (tm_low, tm_high, ms_low, ms_high)]
        assert arrs[0].size == arrs[1].size == 64 * 256 * 4 and arrs[2].size == arrs[3].size == 32 * 32 * 4
        L.orc_scene_set_sky_luts.argtypes = [C.c_void_p] + [C.POINTER(C.c_float)] * 4
        L.orc_scene_set_sky_luts.restype = None
        L.orc_scene_set_sky_luts(self.handle, *[fptr(a) for a in arrs])

    def sky_info(self):
        """-> dict(sun_pos, moon_pos (3,), stars (n, 4) [altitude, azimuth, radius, intensity], stars_offsets (64 * 32 + 1,))"""
        L = lib()
        sun, moon = (C.c_float * 3)(), (C.c_float * 3)()
        stars, offs, count = C.POINTER(C.c_float)(), C.POINTER(C.c_uint32)(), C.c_uint32(0)
        L.orc_scene_sky_info.argtypes = [C.c_void_p, C.c_float * 3, C.c_float * 3, C.POINTER(C.POINTER(C.c_float)), C.POINTER(C.POINTER(C.c_uint32)),
                                         C.POINTER(C.c_uint32)]
        L.orc_scene_sky_info.restype = None
        L.orc_scene_sky_info(self.handle, sun, moon, C.byref(stars), C.byref(offs), C.byref(count))
        n = count.value
        return dict(sun_pos=np.array(sun, np.float32), moon_pos=np.array(moon, np.float32),
                    stars=np.ctypeslib.as_array(stars, shape=(n, 4)).copy() if n else np.zeros((0, 4), np.float32),
                    stars_offsets=np.ctypeslib.as_array(offs, shape=(64 * 32 + 1,)).copy())

    def set_moon_textures(self, albedo: dict = None, normal: dict = None):
        """texture descriptions like scenes.Scene.textures (data (H, W, C), wrap_u, wrap_v, filter, gamma); None = absent"""
        L = lib()
        keep = []
        ta = make_texture(albedo, keep) if albedo is not None else None
        tn = make_texture(normal, keep) if normal is not None else None
        L.orc_scene_set_moon_textures.argtypes = [C.c_void_p, C.POINTER(Texture), C.POINTER(Texture)]
        L.orc_scene_set_moon_textures.restype = None
        L.orc_scene_set_moon_textures(self.handle, C.byref(ta) if ta is not None else None, C.byref(tn) if tn is not None else None)

    def build_sky_hdri(self, origin=None, dim: int = None, samples: int = None, threads: int = 0) -> np.ndarray:
        """sky_compute_hdri: bakes the sky seen from `origin` (default: the camera position) -> (dim, dim, 4)"""
        L = lib()
        sky = getattr(self.scene, "sky", None) or {}
        dim = dim if dim is not None else sky.get("hdri_dim", 2048)
        samples = samples if samples is not None else sky.get("hdri_samples", 32)
        o = (C.c_float * 3)(*(origin if origin is not None else self.scene.camera["pos"]))
        L.orc_scene_build_sky_hdri.argtypes = [C.c_void_p, C.c_float * 3, C.c_uint32, C.c_uint32, C.c_int]
        L.orc_scene_build_sky_hdri.restype = None
        L.orc_scene_build_sky_hdri(self.handle, o, dim, samples, threads)
        return self.sky_hdri()

    def sky_hdri(self) -> np.ndarray:
        L = lib()
        ptr, dim = C.POINTER(C.c_float)(), C.c_uint32(0)
        L.orc_scene_sky_hdri.argtypes = [C.c_void_p, C.POINTER(C.POINTER(C.c_float)), C.POINTER(C.c_uint32)]
        L.orc_scene_sky_hdri.restype = None
        L.orc_scene_sky_hdri(self.handle, C.byref(ptr), C.byref(dim))
        return np.ctypeslib.as_array(ptr, shape=(dim.value, dim.value, 4)).copy()

    def set_sky_hdri(self, color: np.ndarray):
        L = lib()
        a = np.ascontiguousarray(color, np.float32)
        assert a.ndim == 3 and a.shape[0] == a.shape[1] and a.shape[2] == 4
        L.orc_scene_set_sky_hdri.argtypes = [C.c_void_p, C.POINTER(C.c_float), C.c_uint32]
        L.orc_scene_set_sky_hdri.restype = None
        L.orc_scene_set_sky_hdri(self.handle, fptr(a), a.shape[0])

    def sky_colors(self, origins, rays, include_sun, random_offsets, threads: int = 0, mode: int = 0) -> np.ndarray:
        """sky_color_main of explicit rays (mode 0: ray march; mode 1: the baked table + the sun's disc) -> (n, 3)"""
        L = lib()
        if mode != 0:
            o = np.ascontiguousarray(origins, np.float32).reshape(-1, 3)
            d = np.ascontiguousarray(rays, np.float32).reshape(-1, 3)
            inc = np.ascontiguousarray(include_sun, np.uint32).reshape(-1)
            ro = np.ascontiguousarray(random_offsets, np.float32).reshape(-1)
            out = np.zeros((o.shape[0], 3), np.float32)
            L.orc_sky_colors_mode.argtypes = [C.c_void_p, C.c_uint32, C.c_uint32, C.POINTER(C.c_float), C.POINTER(C.c_float), C.POINTER(C.c_uint32),
                                              C.POINTER(C.c_float), C.POINTER(C.c_float), C.c_int]
            L.orc_sky_colors_mode.restype = None
            L.orc_sky_colors_mode(self.handle, mode, o.shape[0], fptr(o), fptr(d), uptr(inc), fptr(ro), fptr(out), threads)
            return out
        o = np.ascontiguousarray(origins, np.float32).reshape(-1, 3)
        d = np.ascontiguousarray(rays, np.float32).reshape(-1, 3)
        inc = np.ascontiguousarray(include_sun, np.uint32).reshape(-1)
        ro = np.ascontiguousarray(random_offsets, np.float32).reshape(-1)
        out = np.zeros((o.shape[0], 3), np.float32)
        L.orc_sky_colors.argtypes = [C.c_void_p, C.c_uint32, C.POINTER(C.c_float), C.POINTER(C.c_float), C.POINTER(C.c_uint32), C.POINTER(C.c_float),
                                     C.POINTER(C.c_float), C.c_int]
        L.orc_sky_colors.restype = None
        L.orc_sky_colors(self.handle, o.shape[0], fptr(o), fptr(d), uptr(inc), fptr(ro), fptr(out), threads)
        return out

    def set_bsdf_luts(self, conductor, glossy, dielectric, dielectric_inv):
        self._luts = [np.ascontiguousarray(a, np.uint16).reshape(-1) for a in (conductor, glossy, dielectric, dielectric_inv)]
        lib().orc_scene_set_bsdf_luts(self.handle, *[a.ctypes.data_as(C.POINTER(C.c_uint16)) for a in self._luts])

    def render(self, first_sample: int, num_samples: int, threads: int = 0, region=None):
        w, h = self.scene.width, self.scene.height
        planes = np.zeros((4, h, w), np.float32)
        counts = RayCounts()
        if region is None:
            region = (0, 0, w, h)
        secs = lib().orc_render_region(self.handle, C.byref(self.camera), C.byref(self.settings), first_sample, num_samples, region[0], region[1],
                                       region[2], region[3], fptr(planes), threads, C.byref(counts))
        return planes, dict(seconds=secs, closest_rays=counts.closest_rays, shadow_rays=counts.shadow_rays, light_enum_rays=counts.light_enum_rays)

    def render_debug(self, shading_mode: int, first_sample: int, num_samples: int, threads: int = 0):
        """Debug shading modes 1..5 (albedo, depth, normal, identification, lights): one bounce, planes summed like render()."""
        w, h = self.scene.width, self.scene.height
        planes = np.zeros((4, h, w), np.float32)
        lib().orc_render_debug(self.handle, C.byref(self.camera), C.byref(self.settings), shading_mode, first_sample, num_samples, fptr(planes), threads)
        return planes

    def path_vertices(self, sample_id: int, iteration: int, threads: int = 0):
        """(vertex inputs VERTEX_IN[n], pixel index[n]) of the paths of one sample that reach wavefront iteration `iteration` on geometry."""
        n = self.scene.width * self.scene.height
        vin = np.zeros(n, VERTEX_IN)
        valid = np.zeros(n, np.uint8)
        lib().orc_path_vertices(self.handle, C.byref(self.camera), C.byref(self.settings), sample_id, iteration, vin.ctypes.data, valid.ctypes.data,
                                threads)
        idx = np.nonzero(valid)[0]
        return vin[idx].copy(), idx

    def shade_vertices(self, vin: np.ndarray, depth: int, threads: int = 0) -> np.ndarray:
        vin = np.ascontiguousarray(vin, VERTEX_IN)
        out = np.zeros(vin.size, VERTEX_OUT)
        lib().orc_shade_vertices(self.handle, C.byref(self.camera), C.byref(self.settings), vin.size, depth, vin.ctypes.data, out.ctypes.data, threads)
        return out

    def nee_segments(self, vin: np.ndarray, depth: int, threads: int = 0) -> np.ndarray:
        """NEE_SEGMENT[n][NEE_SLOTS]: the shadow segments (light-tree light, BSDF-sampled light, ambient, sun) each vertex queues, with the
        oracle's transmittance along them."""
        vin = np.ascontiguousarray(vin, VERTEX_IN)
        out = np.zeros((vin.size, NEE_SLOTS), NEE_SEGMENT)
        lib().orc_nee_segments(self.handle, C.byref(self.camera), C.byref(self.settings), vin.size, depth, vin.ctypes.data, out.ctypes.data, threads)
        return out

    def shadow_rays(self, origins, dirs, limits, ignore_prims, target_prims, threads: int = 0) -> np.ndarray:
        """Transmittance (n, 3) of explicit shadow rays with the reference's any-hit rules."""
        o = np.ascontiguousarray(origins, np.float32).reshape(-1, 3)
        d = np.ascontiguousarray(dirs, np.float32).reshape(-1, 3)
        lim = np.ascontiguousarray(limits, np.float32)
        ig = np.ascontiguousarray(ignore_prims, np.uint32)
        tg = np.ascontiguousarray(target_prims, np.uint32)
        vis = np.zeros((o.shape[0], 3), np.float32)
        lib().orc_shadow_rays(self.handle, o.shape[0], o.ctypes.data, d.ctypes.data, lim.ctypes.data, ig.ctypes.data, tg.ctypes.data, vis.ctypes.data,
                              threads)
        return vis

    def prim_handles(self) -> np.ndarray:
        """(instance_id, tri_id) per flattened primitive."""
        n = self.num_prims()
        out = np.zeros((n, 2), np.uint32)
        a, b = C.c_uint32(), C.c_uint32()
        L = lib()
        for p in range(n):
            L.orc_scene_prim_handle(self.handle, p, C.byref(a), C.byref(b))
            out[p] = (a.value, b.value)
        return out

    def camera_rays(self, sample_id: int = 0):
        L = lib()
        w, h = self.scene.width, self.scene.height
        o = np.empty((h * w, 3), np.float32)
        d = np.empty((h * w, 3), np.float32)
        oo, dd = Vec3(), Vec3()
        for y in range(h):
            for x in range(w):
                L.orc_camera_sample(C.byref(self.camera), C.byref(self.settings), L.orc_path_id_get(x, y, sample_id), C.byref(oo), C.byref(dd))
                o[y * w + x] = (oo.x, oo.y, oo.z)
                d[y * w + x] = (dd.x, dd.y, dd.z)
        return o, d
