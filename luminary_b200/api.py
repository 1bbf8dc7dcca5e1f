"""ctypes mirror of include/lumb200.h.

`Device` follows the reference's per-GPU `Device` object (device/device.h:141-198): add_mesh, update_instances,
update_materials, update_light_tree, update_scene_entity (settings / camera / sky), build_bsdf_lut, start_render,
continue_render (render_samples), plus the parity hooks. Every call goes through the C ABI; nothing is computed in
Python. Errors raise `LuminaryError` carrying the LuminaryResult code (include/luminary/error.h numbering).
"""
from __future__ import annotations

import ctypes as C
import os
from typing import Dict, List, Optional, Sequence

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "liblumb200.so")
DATA_DIR = os.path.join(_HERE, "data")

RESULT_NAMES = {0: "SUCCESS", 1: "ARGUMENT_NULL", 2: "NOT_IMPLEMENTED", 3: "INVALID_API_ARGUMENT", 5: "OUT_OF_MEMORY",
                7: "API_EXCEPTION", 8: "CUDA", 12: "MISSING_DATA", 13: "INVALID_DEVICE"}


class LuminaryError(RuntimeError):
    def __init__(self, code: int, message: str):
        super().__init__(f"LUMINARY_ERROR_{RESULT_NAMES.get(code, code)}: {message}")
        self.code = code


class Mesh(C.Structure):
    _fields_ = [("triangle_count", C.c_uint32), ("vertex_buffer", C.POINTER(C.c_float)), ("normal_buffer", C.POINTER(C.c_float)),
                ("uv_buffer", C.POINTER(C.c_float)), ("material_id_buffer", C.POINTER(C.c_uint16))]


class Instance(C.Structure):
    _fields_ = [("mesh_id", C.c_uint32), ("translation", C.c_float * 3), ("rotation", C.c_float * 3), ("scale", C.c_float * 3),
                ("active", C.c_uint32)]


class Material(C.Structure):
    _fields_ = [("base_substrate", C.c_uint32), ("albedo", C.c_float * 4), ("emission", C.c_float * 3), ("emission_scale", C.c_float),
                ("roughness", C.c_float), ("roughness_clamp", C.c_float), ("refraction_index", C.c_float), ("emission_active", C.c_uint8),
                ("thin_walled", C.c_uint8), ("metallic", C.c_uint8), ("colored_transparency", C.c_uint8), ("roughness_as_smoothness", C.c_uint8),
                ("normal_map_is_compressed", C.c_uint8), ("bidirectional_emission", C.c_uint8), ("_pad", C.c_uint8),
                ("albedo_tex", C.c_uint16), ("luminance_tex", C.c_uint16), ("roughness_tex", C.c_uint16), ("metallic_tex", C.c_uint16),
                ("normal_tex", C.c_uint16), ("_pad2", C.c_uint16)]


TEXTURE_NONE = 0xFFFF
_TEX_TYPES = {np.dtype(np.float32): 0, np.dtype(np.uint8): 1, np.dtype(np.uint16): 2}


class Texture(C.Structure):
    _fields_ = [("width", C.c_uint32), ("height", C.c_uint32), ("pitch", C.c_uint32), ("type", C.c_uint32), ("num_components", C.c_uint32),
                ("wrap_mode_u", C.c_uint32), ("wrap_mode_v", C.c_uint32), ("filter", C.c_uint32), ("gamma", C.c_float), ("mipmap", C.c_uint32),
                ("data", C.c_void_p)]


class Settings(C.Structure):
    _fields_ = [("width", C.c_uint32), ("height", C.c_uint32), ("max_ray_depth", C.c_uint32), ("sort_by_material", C.c_uint32)]


class Camera(C.Structure):
    _fields_ = [("pos", C.c_float * 3), ("rotation", C.c_float * 3), ("fov", C.c_float), ("aperture_size", C.c_float),
                ("object_distance", C.c_float), ("camera_scale", C.c_float), ("russian_roulette_threshold", C.c_float),
                ("aperture_shape", C.c_uint32), ("aperture_blade_count", C.c_uint32)]


class Sky(C.Structure):  # Lumb200Sky
    _fields_ = [("mode", C.c_uint32), ("constant_color", C.c_float * 3), ("geometry_offset", C.c_float * 3)] + [(n, C.c_float) for n in (
        "azimuth", "altitude", "moon_azimuth", "moon_altitude", "moon_tex_offset", "sun_strength", "base_density", "rayleigh_density", "mie_density",
        "ozone_density", "rayleigh_falloff", "mie_falloff", "mie_diameter", "ground_visibility", "ozone_layer_thickness", "multiscattering_factor",
        "stars_intensity")] + [(n, C.c_uint32) for n in ("steps", "ozone_absorption", "aerial_perspective", "stars_count", "stars_seed", "hdri_dim", "hdri_samples")]


class LightTree(C.Structure):
    _fields_ = [("root_data", C.c_void_p), ("root_size", C.c_size_t), ("nodes_data", C.c_void_p), ("nodes_size", C.c_size_t),
                ("tri_handle_map", C.POINTER(C.c_uint32)), ("num_lights", C.c_uint32)]


class OutputParams(C.Structure):
    _fields_ = [("exposure", C.c_float), ("tonemap", C.c_uint32), ("agx_slope", C.c_float), ("agx_power", C.c_float), ("agx_saturation", C.c_float),
                ("dithering", C.c_uint32), ("purkinje", C.c_uint32), ("purkinje_kappa1", C.c_float), ("purkinje_kappa2", C.c_float),
                ("supersampling", C.c_uint32), ("local_error_minimization", C.c_uint32), ("bloom_blend", C.c_float), ("filter", C.c_uint32),
                ("use_color_correction", C.c_uint32), ("color_correction", C.c_float * 3), ("film_grain", C.c_float)]


class AdaptiveSampling(C.Structure):
    _fields_ = [("enable", C.c_uint32), ("max_sampling_rate", C.c_uint32), ("avg_sampling_rate", C.c_uint32), ("update_interval", C.c_uint32),
                ("exposure_aware", C.c_uint32), ("exposure", C.c_float), ("tonemap", C.c_uint32), ("agx_slope", C.c_float), ("agx_power", C.c_float),
                ("agx_saturation", C.c_float), ("output_mode", C.c_uint32)]


class AdaptiveState(C.Structure):
    _fields_ = [("stage_id", C.c_uint32), ("executions", C.c_uint32 * 5), ("tasks_per_execution", C.c_uint32), ("blocks_x", C.c_uint32),
                ("blocks_y", C.c_uint32), ("paths_traced", C.c_uint64)]


class Profile(C.Structure):
    _fields_ = [("milliseconds", C.c_double * 8), ("launches", C.c_uint64 * 8)]


class TraversalStats(C.Structure):
    _fields_ = [("closest_rays", C.c_uint64), ("closest_nodes", C.c_uint64), ("closest_tris", C.c_uint64), ("shadow_rays", C.c_uint64),
                ("shadow_nodes", C.c_uint64), ("shadow_tris", C.c_uint64), ("light_rays", C.c_uint64), ("shaded_vertices", C.c_uint64),
                ("light_tree_nodes", C.c_uint64), ("light_root_sections", C.c_uint32), ("_pad", C.c_uint32)]


KERNEL_CLASSES = ["raygen", "trace_closest", "sort", "shade", "trace_shadow", "accumulate", "trace_enum"]


class LightTreeBuffers(C.Structure):
    _fields_ = [("root_data", C.c_void_p), ("root_size", C.c_size_t), ("nodes_data", C.c_void_p), ("nodes_size", C.c_size_t),
                ("tri_handle_map", C.POINTER(C.c_uint32)), ("num_lights", C.c_uint32)]


class Stats(C.Structure):
    _fields_ = [("closest_rays", C.c_uint64), ("shadow_rays", C.c_uint64), ("light_rays", C.c_uint64), ("kernel_launches", C.c_uint64),
                ("render_seconds", C.c_double), ("accel_build_seconds", C.c_double), ("samples_done", C.c_uint32), ("bvh_nodes", C.c_uint32),
                ("bvh_tris", C.c_uint32), ("light_bvh_nodes", C.c_uint32), ("device_bytes", C.c_uint64), ("bvh_depth", C.c_uint32),
                ("light_bvh_depth", C.c_uint32), ("bvh_sah_cost", C.c_float), ("bvh_ploc_radius", C.c_uint32), ("stack_overflows", C.c_uint64), ("nonfinite_samples", C.c_uint64),
                ("nonfinite_pixel", C.c_uint32), ("reserved0", C.c_uint32)]


# every symbol include/lumb200.h declares (tests check that the library exports all of them)
EXPORTED_SYMBOLS = [
    "lumb200_last_error", "lumb200_get_device_count", "lumb200_device_create", "lumb200_device_destroy", "lumb200_device_load_bluenoise",
    "lumb200_device_add_mesh", "lumb200_device_update_instances", "lumb200_device_update_materials", "lumb200_device_update_materials_packed",
    "lumb200_device_update_light_tree", "lumb200_host_build_light_tree", "lumb200_host_free_light_tree", "lumb200_device_update_settings", "lumb200_device_update_camera", "lumb200_device_set_shading_mode", "lumb200_device_update_sky",
    "lumb200_sky_default", "lumb200_device_get_sky_lut", "lumb200_device_get_sky_info", "lumb200_device_build_sky_hdri", "lumb200_device_get_sky_hdri",
    "lumb200_device_load_moon_textures",
    "lumb200_device_build_bsdf_lut", "lumb200_device_get_bsdf_lut", "lumb200_device_set_bsdf_lut", "lumb200_device_build_accel",
    "lumb200_device_start_render", "lumb200_device_render_samples", "lumb200_device_sync", "lumb200_device_get_frame_planes",
    "lumb200_device_bind_frame_planes", "lumb200_device_download_frame_planes", "lumb200_device_download_result", "lumb200_device_trace_primary",
    "lumb200_device_trace_rays", "lumb200_device_download_bvh", "lumb200_device_get_stats", "lumb200_device_set_profiling", "lumb200_device_get_profile",
    "lumb200_device_measure_traversal", "lumb200_device_get_stream", "lumb200_device_time_primary_trace",
    "lumb200_device_load_bluenoise_1d", "lumb200_device_download_output_argb8", "lumb200_device_add_planes_from", "lumb200_get_device_properties",
    "lumb200_device_add_textures", "lumb200_device_sample_texture", "lumb200_device_compute_light_intensities",
    "lumb200_host_build_light_tree_textured", "lumb200_device_sample_texture_lod", "lumb200_device_update_adaptive_sampling",
    "lumb200_device_render_executions", "lumb200_device_get_adaptive_state", "lumb200_device_download_adaptive_words",
    "lumb200_device_set_adaptive_state", "lumb200_device_render_allocated_execution", "lumb200_device_build_adaptive_stage",
    "lumb200_device_download_result_async", "lumb200_device_wait_download",
    "lumb200_device_shade_vertices", "lumb200_device_trace_shadow_rays", "lumb200_host_pack_material", "lumb200_host_pack_triangles",
    "lumb200_host_pack_transform",
    "lumb200_comm_get_unique_id", "lumb200_comm_create_rank", "lumb200_comm_create_all", "lumb200_comm_destroy", "lumb200_comm_get_info",
    "lumb200_comm_reduce_planes", "lumb200_comm_reduce_planes_all", "lumb200_comm_broadcast_adaptive_words",
    "lumb200_comm_broadcast_adaptive_words_all", "lumb200_device_get_cuda_index", "lumb200_device_get_adaptive_words_device",
    "lumb200_device_adopt_adaptive_stage", "lumb200_device_clear_frame_planes", "lumb200_device_query_pixel",
]
COMM_ID_BYTES = 128

# Lumb200VertexIn / Lumb200NeeSegment / Lumb200VertexOut of include/lumb200.h as numpy record types (all members 4-byte aligned)
_V3 = (np.float32, 3)
VERTEX_IN = np.dtype([("pixel_x", np.uint32), ("pixel_y", np.uint32), ("state", np.uint32), ("origin", *_V3), ("ray", *_V3), ("prim", np.uint32),
                      ("t", np.float32), ("record", np.uint32, 2), ("medium", np.uint32)])
NEE_SEGMENT = np.dtype([("valid", np.uint32), ("ray", *_V3), ("dist", np.float32), ("color", *_V3), ("target_prim", np.uint32),
                        ("visible", *_V3)])
NEE_SLOTS = 4  # LB_NEE_SLOTS: light-tree light, BSDF-sampled light, ambient, sun
VERTEX_OUT = np.dtype([("nee", NEE_SEGMENT, NEE_SLOTS), ("emission", *_V3), ("alive", np.uint32), ("state", np.uint32), ("origin", *_V3), ("ray", *_V3),
                       ("record", np.uint32, 2), ("medium", np.uint32)])

_lib = None


def load_library() -> C.CDLL:
    """Loads liblumb200.so. Raises (never falls back) when the CUDA library has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    path = os.environ.get("LUMB200_LIBRARY", LIB_PATH)  # override: compile-time variants of the same library (tuning experiments)
    if not os.path.exists(path):
        raise ImportError(f"{path} is missing: build it with `python -m luminary_b200.build` (nvcc, sm_100a). There is no CPU fallback.")
    lib = C.CDLL(path)
    lib.lumb200_last_error.restype = C.c_char_p
    for name in EXPORTED_SYMBOLS:
        fn = getattr(lib, name)
        if name == "lumb200_host_free_light_tree":
            fn.restype = None
        elif name != "lumb200_last_error":
            fn.restype = C.c_uint64
    _lib = lib
    return lib


class Comm:
    """NCCL communicator of the C ABI (csrc/comm.cu). One-process-per-device flavour: rank 0 calls Comm.unique_id(), the bytes reach the
    other ranks by any transport, every rank builds Comm(device, world, rank, id). In-process flavour: Comm.create_all(devices)."""

    def __init__(self, device=None, world: int = 1, rank: int = 0, uid: bytes = b"", _handle=None):
        self._lib = load_library()
        if _handle is not None:
            self._h = _handle
            return
        assert len(uid) == COMM_ID_BYTES
        self._h = C.c_void_p()
        _check(self._lib.lumb200_comm_create_rank(C.byref(self._h), device._h, C.c_uint32(world), C.c_uint32(rank), C.c_char_p(uid)))

    @staticmethod
    def unique_id() -> bytes:
        buf = C.create_string_buffer(COMM_ID_BYTES)
        _check(load_library().lumb200_comm_get_unique_id(buf))
        return buf.raw

    @staticmethod
    def create_all(devices):
        lib = load_library()
        n = len(devices)
        handles = (C.c_void_p * n)()
        devs = (C.c_void_p * n)(*[d._h for d in devices])
        _check(lib.lumb200_comm_create_all(handles, devs, C.c_uint32(n)))
        return [Comm(_handle=C.c_void_p(handles[k])) for k in range(n)]

    def info(self) -> Dict:
        w, r, v = C.c_uint32(), C.c_uint32(), C.c_uint32()
        _check(self._lib.lumb200_comm_get_info(self._h, C.byref(w), C.byref(r), C.byref(v)))
        return dict(world=w.value, rank=r.value, nccl_version=v.value)

    def reduce_planes(self, root: int = 0) -> None:
        """in-place sum-reduce of this member's accumulation planes onto rank `root`, queued on the device's render stream"""
        _check(self._lib.lumb200_comm_reduce_planes(self._h, C.c_uint32(root)))

    def broadcast_adaptive_words(self, root: int = 0) -> None:
        _check(self._lib.lumb200_comm_broadcast_adaptive_words(self._h, C.c_uint32(root)))

    @staticmethod
    def reduce_planes_all(comms, root: int = 0) -> None:
        n = len(comms)
        arr = (C.c_void_p * n)(*[c._h for c in comms])
        _check(load_library().lumb200_comm_reduce_planes_all(arr, C.c_uint32(n), C.c_uint32(root)))

    @staticmethod
    def broadcast_adaptive_words_all(comms, root: int = 0) -> None:
        n = len(comms)
        arr = (C.c_void_p * n)(*[c._h for c in comms])
        _check(load_library().lumb200_comm_broadcast_adaptive_words_all(arr, C.c_uint32(n), C.c_uint32(root)))

    def destroy(self) -> None:
        if self._h:
            self._lib.lumb200_comm_destroy(C.byref(self._h))
            self._h = None


def _check(code: int) -> None:
    if code != 0:
        raise LuminaryError(int(code), load_library().lumb200_last_error().decode("utf-8", "replace"))


def device_count() -> int:
    n = C.c_uint32(0)
    _check(load_library().lumb200_get_device_count(C.byref(n)))
    return n.value


def load_bluenoise_2d() -> np.ndarray:
    """The reference's data/bluenoise/bluenoise_2D.bin (256 x 256 uint32), shipped as data with the package."""
    return np.fromfile(os.path.join(DATA_DIR, "bluenoise_2D.bin"), dtype=np.uint32)


def read_png_rgba8(path: str) -> np.ndarray:
    """Minimal PNG reader for the shipped 8-bit grey / RGB / RGBA, non-interlaced files -> (H, W, 4) uint8 the way the reference's png_load
    expands them (grey -> r = g = b, alpha 255)."""
    import struct
    import zlib

    raw = open(path, "rb").read()
    assert raw[:8] == b"\x89PNG\r\n\x1a\n", path
    pos, idat, hdr = 8, b"", None
    while pos < len(raw):
        n, typ = struct.unpack(">I4s", raw[pos:pos + 8])
        body = raw[pos + 8:pos + 8 + n]
        if typ == b"IHDR":
            hdr = struct.unpack(">IIBBBBB", body)
        elif typ == b"IDAT":
            idat += body
        pos += 12 + n
    w, h, depth, ctype, _c, _f, interlace = hdr
    assert depth == 8 and interlace == 0 and ctype in (0, 2, 6), (path, hdr)
    ch = {0: 1, 2: 3, 6: 4}[ctype]
    data = np.frombuffer(zlib.decompress(idat), np.uint8).reshape(h, 1 + w * ch)
    out = np.zeros((h, w * ch), np.uint8)
    prev = np.zeros(w * ch, np.int32)
    for y in range(h):
        f, line = int(data[y, 0]), data[y, 1:].astype(np.int32)
        if f == 0:
            cur = line
        elif f == 2:
            cur = (line + prev) & 0xFF
        else:  # sub / average / paeth carry a left-neighbour dependency: per channel, sequential in x
            cur = np.zeros(w * ch, np.int32)
            for x in range(w * ch):
                a = cur[x - ch] if x >= ch else 0
                b = prev[x]
                c = prev[x - ch] if x >= ch else 0
                if f == 1:
                    p = a
                elif f == 3:
                    p = (a + b) >> 1
                else:
                    pa, pb, pc = abs(b - c), abs(a - c), abs(a + b - 2 * c)
                    p = a if (pa <= pb and pa <= pc) else (b if pb <= pc else c)
                cur[x] = (line[x] + p) & 0xFF
        out[y] = cur
        prev = cur
    px = out.reshape(h, w, ch)
    rgba = np.full((h, w, 4), 255, np.uint8)
    if ch == 1:
        rgba[..., :3] = px
    else:
        rgba[..., :ch] = px
    return rgba


def load_moon_textures():
    """-> (albedo, normal) texture descriptions of the shipped moon surface, as the reference's png_load + texture_create deliver them
    (RGBA8, wrap addressing, linear filter, gamma 1: neither file carries a gAMA chunk). Decoded once and cached as .npy next to the files."""
    out = []
    for name in ("moon_albedo", "moon_normal"):
        png = os.path.join(os.path.dirname(os.path.abspath(__file__)), "data", name + ".png")
        cache = png[:-4] + ".rgba8.npy"
        if os.path.exists(cache) and os.path.getmtime(cache) >= os.path.getmtime(png):
            px = np.load(cache)
        else:
            px = read_png_rgba8(png)
            try:
                np.save(cache, px)
            except OSError:
                pass
        out.append(dict(data=px, wrap_u=0, wrap_v=0, filter=1, gamma=1.0))
    return out[0], out[1]


def load_bluenoise_1d() -> np.ndarray:
    """The reference's data/bluenoise/bluenoise_1D.bin (256 x 256 uint16): dither mask of the output chain."""
    return np.fromfile(os.path.join(DATA_DIR, "bluenoise_1D.bin"), dtype=np.uint16)


def _fptr(a: np.ndarray):
    return a.ctypes.data_as(C.POINTER(C.c_float))


def material_struct(m: Dict) -> Material:
    s = Material()
    s.base_substrate = int(m["base_substrate"])
    s.albedo[:] = m["albedo"]
    s.emission[:] = m["emission"]
    s.emission_scale = m["emission_scale"]
    s.roughness = m["roughness"]
    s.roughness_clamp = m["roughness_clamp"]
    s.refraction_index = m["refraction_index"]
    for k in ("emission_active", "thin_walled", "metallic", "colored_transparency", "roughness_as_smoothness", "normal_map_is_compressed",
              "bidirectional_emission"):
        setattr(s, k, 1 if m[k] else 0)
    for k in ("albedo_tex", "luminance_tex", "roughness_tex", "metallic_tex", "normal_tex"):
        setattr(s, k, int(m.get(k, TEXTURE_NONE)))
    return s


def texture_struct(t: Dict, keep: list) -> Texture:
    """t: dict(data = (H, W, C) array of uint8 / uint16 / float32 or None for an invalid texture, wrap_u, wrap_v, filter, gamma)."""
    s = Texture()
    s.wrap_mode_u = int(t.get("wrap_u", 0))
    s.wrap_mode_v = int(t.get("wrap_v", 0))
    s.filter = int(t.get("filter", 1))
    s.gamma = float(t.get("gamma", 1.0))
    s.mipmap = int(t.get("mipmap", 0))
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


def textured_emitter_triangles(scene):
    """(mesh_ids, triangle_ids) of every triangle whose material has an active luminance-textured emission: the tasks of
    the reference's _light_tree_queue_texture_integrations (device_light.c:1904-1950)."""
    mesh_ids, tri_ids = [], []
    textured = np.array([bool(m["emission_active"]) and m.get("luminance_tex", TEXTURE_NONE) != TEXTURE_NONE for m in scene.materials], bool)
    for mi, m in enumerate(scene.meshes):
        mid = np.asarray(m.material, np.int64)
        sel = np.nonzero(textured[np.minimum(mid, len(textured) - 1)] & (mid < len(textured)))[0]
        mesh_ids.append(np.full(sel.size, mi, np.uint32))
        tri_ids.append(sel.astype(np.uint32))
    return np.concatenate(mesh_ids) if mesh_ids else np.zeros(0, np.uint32), np.concatenate(tri_ids) if tri_ids else np.zeros(0, np.uint32)


def pack_material(m: Dict) -> bytes:
    """The product's host-side material packer (device_api.cu: pack_material) on one material dict: 32 bytes."""
    out = C.create_string_buffer(32)
    ms = material_struct(m)
    _check(load_library().lumb200_host_pack_material(C.byref(ms), out))
    return out.raw


def pack_triangles(mesh):
    """The product's host-side triangle packer on a scenes.Mesh: (vertices uint32[3 * T, 4], textris uint32[T, 4])."""
    v = np.ascontiguousarray(mesh.vertex, np.float32).reshape(-1)
    n = np.ascontiguousarray(mesh.normal, np.float32).reshape(-1)
    t = np.ascontiguousarray(mesh.uv, np.float32).reshape(-1)
    mm = np.ascontiguousarray(mesh.material, np.uint16).reshape(-1)
    ms = Mesh(mesh.num_tris, _fptr(v), _fptr(n), _fptr(t), mm.ctypes.data_as(C.POINTER(C.c_uint16)))
    verts = np.zeros((3 * mesh.num_tris, 4), np.uint32)
    tex = np.zeros((mesh.num_tris, 4), np.uint32)
    _check(load_library().lumb200_host_pack_triangles(C.byref(ms), C.c_void_p(verts.ctypes.data), C.c_void_p(tex.ctypes.data)))
    return verts, tex


def pack_transform(translation, rotation, scale) -> bytes:
    """The product's DeviceTransform packer (Quaternion16): 32 bytes."""
    ins = Instance()
    ins.mesh_id = 0
    ins.translation[:] = [float(x) for x in translation]
    ins.rotation[:] = [float(x) for x in rotation]
    ins.scale[:] = [float(x) for x in scale]
    ins.active = 1
    out = C.create_string_buffer(32)
    _check(load_library().lumb200_host_pack_transform(C.byref(ins), out))
    return out.raw


def build_light_tree(scene, triangle_intensities=None):
    """Runs the host-side light tree builder (C, csrc/host/light_tree.c) on a scenes.Scene.
    triangle_intensities: optional {mesh_id: float32 array of num_tris} from Device.compute_light_intensities (luminance-textured
    emitters). Returns (root bytes, nodes bytes, tri_handle_map uint32[num_lights, 2]) or None when the scene has no emitters."""
    lib = load_library()
    keep = []
    meshes = (Mesh * max(len(scene.meshes), 1))()
    for i, m in enumerate(scene.meshes):
        v = np.ascontiguousarray(m.vertex, np.float32).reshape(-1)
        n = np.ascontiguousarray(m.normal, np.float32).reshape(-1)
        t = np.ascontiguousarray(m.uv, np.float32).reshape(-1)
        mm = np.ascontiguousarray(m.material, np.uint16).reshape(-1)
        keep += [v, n, t, mm]
        meshes[i] = Mesh(m.num_tris, _fptr(v), _fptr(n), _fptr(t), mm.ctypes.data_as(C.POINTER(C.c_uint16)))
    inst = (Instance * max(len(scene.instances), 1))()
    for i, ins in enumerate(scene.instances):
        inst[i].mesh_id = ins.mesh_id
        inst[i].translation[:] = ins.translation
        inst[i].rotation[:] = ins.rotation
        inst[i].scale[:] = ins.scale
        inst[i].active = 1 if ins.active else 0
    mats = (Material * max(len(scene.materials), 1))()
    for i, m in enumerate(scene.materials):
        mats[i] = material_struct(m)
    out = LightTreeBuffers()
    ti = (C.POINTER(C.c_float) * max(len(scene.meshes), 1))()
    for mi, arr in (triangle_intensities or {}).items():
        a = np.ascontiguousarray(arr, np.float32)
        assert a.size == scene.meshes[mi].num_tris
        keep.append(a)
        ti[mi] = _fptr(a)
    _check(lib.lumb200_host_build_light_tree_textured(meshes, C.c_uint32(len(scene.meshes)), inst, C.c_uint32(len(scene.instances)), mats,
                                                      C.c_uint32(len(scene.materials)), ti if triangle_intensities else None, C.byref(out)))
    if out.num_lights == 0:
        return None
    root = C.string_at(out.root_data, out.root_size)
    nodes = C.string_at(out.nodes_data, out.nodes_size) if out.nodes_size else b""
    handles = np.ctypeslib.as_array(out.tri_handle_map, shape=(out.num_lights, 2)).copy()
    lib.lumb200_host_free_light_tree(C.byref(out))
    return root, nodes, handles


class Device:
    """One GPU. Mirrors the reference's Device object for the path-tracing hot path."""

    def __init__(self, cuda_index: int = 0, load_embedded_data: bool = True):
        self._lib = load_library()
        self._h = C.c_void_p()
        _check(self._lib.lumb200_device_create(C.byref(self._h), C.c_uint32(cuda_index)))
        self.width = 0
        self.height = 0
        self._keep = []
        if load_embedded_data:
            self.load_bluenoise(load_bluenoise_2d())

    def destroy(self) -> None:
        if self._h:
            _check(self._lib.lumb200_device_destroy(C.byref(self._h)))
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.destroy()
        except Exception:
            pass

    # -- embedded data / scene ---------------------------------------------------------------
    def load_bluenoise(self, table: np.ndarray) -> None:
        t = np.ascontiguousarray(table, dtype=np.uint32)
        _check(self._lib.lumb200_device_load_bluenoise(self._h, t.ctypes.data_as(C.POINTER(C.c_uint32)), C.c_size_t(t.size)))

    def add_mesh(self, vertex: np.ndarray, normal: np.ndarray, uv: np.ndarray, material: np.ndarray) -> int:
        v = np.ascontiguousarray(vertex, dtype=np.float32).reshape(-1)
        n = np.ascontiguousarray(normal, dtype=np.float32).reshape(-1)
        t = np.ascontiguousarray(uv, dtype=np.float32).reshape(-1)
        m = np.ascontiguousarray(material, dtype=np.uint16).reshape(-1)
        count = m.size
        if v.size != 9 * count or n.size != 9 * count or t.size != 6 * count:
            raise LuminaryError(3, "mesh buffers have inconsistent sizes")
        mesh = Mesh(count, _fptr(v), _fptr(n), _fptr(t), m.ctypes.data_as(C.POINTER(C.c_uint16)))
        mesh_id = C.c_uint32(0)
        _check(self._lib.lumb200_device_add_mesh(self._h, C.byref(mesh), C.byref(mesh_id)))
        return mesh_id.value

    def update_instances(self, instances: Sequence) -> None:
        arr = (Instance * max(len(instances), 1))()
        for i, ins in enumerate(instances):
            arr[i].mesh_id = ins.mesh_id
            arr[i].translation[:] = ins.translation
            arr[i].rotation[:] = ins.rotation
            arr[i].scale[:] = ins.scale
            arr[i].active = 1 if ins.active else 0
        _check(self._lib.lumb200_device_update_instances(self._h, arr, C.c_uint32(len(instances))))

    def update_materials(self, materials: Sequence[Dict]) -> None:
        arr = (Material * max(len(materials), 1))()
        for i, m in enumerate(materials):
            arr[i] = material_struct(m)
        _check(self._lib.lumb200_device_update_materials(self._h, arr, C.c_uint32(len(materials))))

    def add_textures(self, textures: Sequence[Dict]) -> None:
        keep = []
        arr = (Texture * max(len(textures), 1))()
        for i, t in enumerate(textures):
            arr[i] = texture_struct(t, keep)
        _check(self._lib.lumb200_device_add_textures(self._h, arr, C.c_uint32(len(textures))))

    def sample_texture(self, texture_id: int, uv: np.ndarray, lod: float = 0.0) -> np.ndarray:
        """Raw tex2DLod<float4> fetches (parity hook). uv: (N, 2) float32 -> (N, 4) float32."""
        uv = np.ascontiguousarray(uv, np.float32).reshape(-1, 2)
        out = np.empty((uv.shape[0], 4), np.float32)
        _check(self._lib.lumb200_device_sample_texture_lod(self._h, C.c_uint32(texture_id), _fptr(uv), C.c_uint32(uv.shape[0]), C.c_float(lod),
                                                           _fptr(out)))
        return out

    def update_light_tree(self, root: bytes, nodes: bytes, tri_handle_map: np.ndarray) -> None:
        hm = np.ascontiguousarray(tri_handle_map, dtype=np.uint32).reshape(-1)
        rb = C.create_string_buffer(root, len(root)) if len(root) else None
        nb = C.create_string_buffer(nodes, len(nodes)) if len(nodes) else None
        lt = LightTree(C.cast(rb, C.c_void_p) if rb else None, len(root), C.cast(nb, C.c_void_p) if nb else None, len(nodes),
                       hm.ctypes.data_as(C.POINTER(C.c_uint32)), hm.size // 2)
        _check(self._lib.lumb200_device_update_light_tree(self._h, C.byref(lt)))

    def update_settings(self, width: int, height: int, max_ray_depth: int, sort_by_material: bool = True) -> None:
        s = Settings(width, height, max_ray_depth, 1 if sort_by_material else 0)
        _check(self._lib.lumb200_device_update_settings(self._h, C.byref(s)))
        self.width, self.height = width, height

    def set_shading_mode(self, shading_mode: int) -> None:
        """LuminaryShadingMode: 0 path tracer, 1 albedo, 2 depth, 3 normal, 4 identification, 5 lights (one-bounce debug queue)."""
        _check(self._lib.lumb200_device_set_shading_mode(self._h, C.c_uint32(shading_mode)))

    def update_camera(self, cam: Dict) -> None:
        c = Camera()
        c.pos[:] = cam["pos"]
        c.rotation[:] = cam["rotation"]
        c.fov = cam["fov"]
        c.aperture_size = cam["aperture_size"]
        c.object_distance = cam["object_distance"]
        c.camera_scale = cam["camera_scale"]
        c.russian_roulette_threshold = cam["russian_roulette_threshold"]
        c.aperture_shape = cam["aperture_shape"]
        c.aperture_blade_count = cam["aperture_blade_count"]
        _check(self._lib.lumb200_device_update_camera(self._h, C.byref(c)))

    def update_sky(self, mode: int, color=(1.0, 1.0, 1.0), sky: Dict = None) -> None:
        """mode 2: constant colour. mode 0: the procedural atmosphere with the reference's defaults (sky.c:6-42) overridden by the
        entries of `sky` (field names of Lumb200Sky); builds the sky LUTs when a medium parameter changed. mode 1: the same
        atmosphere baked into a hdri_dim^2 table at the next start_render / build_sky_hdri."""
        s = Sky()
        self._lib.lumb200_sky_default.restype = None
        self._lib.lumb200_sky_default(C.byref(s))
        s.mode = mode
        s.constant_color[:] = color
        for k, v in (sky or {}).items():
            if k == "geometry_offset":
                s.geometry_offset[:] = v
            elif k != "mode":
                setattr(s, k, v)
        _check(self._lib.lumb200_device_update_sky(self._h, C.byref(s)))

    def get_sky_lut(self):
        """-> (tm_low, tm_high (64, 256, 4), ms_low, ms_high (32, 32, 4)): the transmittance / multiscattering tables of the procedural sky"""
        tm = [np.zeros((64, 256, 4), np.float32) for _ in range(2)]
        ms = [np.zeros((32, 32, 4), np.float32) for _ in range(2)]
        _check(self._lib.lumb200_device_get_sky_lut(self._h, _fptr(tm[0]), _fptr(tm[1]), _fptr(ms[0]), _fptr(ms[1])))
        return tm[0], tm[1], ms[0], ms[1]

    def load_moon_textures(self, albedo: Dict = "data", normal: Dict = "data") -> None:
        """The moon's surface textures (device_load_embedded_data). Default: the shipped data/moon_*.png; None = absent (black disc)."""
        if isinstance(albedo, str) or isinstance(normal, str):
            a, n = load_moon_textures()
            albedo = a if isinstance(albedo, str) else albedo
            normal = n if isinstance(normal, str) else normal
        keep = []
        ta = texture_struct(albedo, keep) if albedo is not None else None
        tn = texture_struct(normal, keep) if normal is not None else None
        _check(self._lib.lumb200_device_load_moon_textures(self._h, C.byref(ta) if ta is not None else None, C.byref(tn) if tn is not None else None))

    def build_sky_hdri(self) -> None:
        """Bakes the sky HDRI (sky mode 1) from the current camera position; start_render does it implicitly after a sky change."""
        _check(self._lib.lumb200_device_build_sky_hdri(self._h))

    def get_sky_hdri(self):
        """-> (color (dim, dim, 4) float32, origin (3,)) of the baked table"""
        dim = C.c_uint32(0)
        origin = np.zeros(3, np.float32)
        _check(self._lib.lumb200_device_get_sky_hdri(self._h, None, C.c_uint32(0), C.byref(dim), _fptr(origin)))
        color = np.zeros((dim.value, dim.value, 4), np.float32)
        _check(self._lib.lumb200_device_get_sky_hdri(self._h, _fptr(color), C.c_uint32(dim.value * dim.value), None, None))
        return color, origin

    def get_sky_info(self) -> Dict:
        """-> dict(sun_pos, moon_pos (3,), stars (n, 4) [altitude, azimuth, radius, intensity], stars_offsets (64 * 32 + 1,))"""
        sun, moon = np.zeros(3, np.float32), np.zeros(3, np.float32)
        n = C.c_uint32(0)
        offs = np.zeros(64 * 32 + 1, np.uint32)
        _check(self._lib.lumb200_device_get_sky_info(self._h, _fptr(sun), _fptr(moon), None, C.c_uint32(0), offs.ctypes.data_as(C.POINTER(C.c_uint32)),
                                                     C.byref(n)))
        stars = np.zeros((n.value, 4), np.float32)
        if n.value:
            _check(self._lib.lumb200_device_get_sky_info(self._h, None, None, _fptr(stars), n, None, None))
        return dict(sun_pos=sun, moon_pos=moon, stars=stars, stars_offsets=offs)

    def compute_light_intensities(self, mesh_ids: np.ndarray, tri_ids: np.ndarray) -> np.ndarray:
        mesh_ids = np.ascontiguousarray(mesh_ids, np.uint32)
        tri_ids = np.ascontiguousarray(tri_ids, np.uint32)
        out = np.zeros(mesh_ids.size, np.float32)
        p = lambda a: a.ctypes.data_as(C.POINTER(C.c_uint32))
        _check(self._lib.lumb200_device_compute_light_intensities(self._h, p(mesh_ids), p(tri_ids), C.c_uint32(mesh_ids.size), _fptr(out)))
        return out

    def build_light_tree(self, scene):
        """The reference's light tree flow for a scene whose meshes, textures and materials are already on this device
        (device_build_light_tree, device.h:162): integrate the luminance textures of textured emitters on the device, then
        build the tree on the host."""
        mesh_ids, tri_ids = textured_emitter_triangles(scene)
        intensities = None
        if mesh_ids.size:
            val = self.compute_light_intensities(mesh_ids, tri_ids)
            intensities = {}
            for mi in np.unique(mesh_ids):
                a = np.ones(scene.meshes[int(mi)].num_tris, np.float32)
                sel = mesh_ids == mi
                a[tri_ids[sel]] = val[sel]
                intensities[int(mi)] = a
        return build_light_tree(scene, intensities)

    def load_scene(self, scene, light_tree=None):
        """Uploads a luminary_b200.scenes.Scene the way the device manager does (device_manager.c:281-513).
        light_tree: a tuple from build_light_tree, None (no NEE), or "auto" = Device.build_light_tree after the upload.
        Returns the light tree that was uploaded (or None)."""
        for m in scene.meshes:
            self.add_mesh(m.vertex, m.normal, m.uv, m.material)
        self.update_instances(scene.instances)
        if getattr(scene, "textures", None):
            self.add_textures(scene.textures)
        self.update_materials(scene.materials)
        if isinstance(light_tree, str) and light_tree == "auto":
            light_tree = self.build_light_tree(scene)
        self.update_settings(scene.width, scene.height, scene.max_ray_depth)
        self.update_camera(scene.camera)
        self.update_sky(scene.sky_mode, scene.sky_color, getattr(scene, "sky", None))
        if light_tree is not None:
            self.update_light_tree(*light_tree)
        self.build_accel()
        return light_tree

    # -- builds -------------------------------------------------------------------------------
    def build_accel(self) -> None:
        _check(self._lib.lumb200_device_build_accel(self._h))

    def build_bsdf_lut(self) -> None:
        _check(self._lib.lumb200_device_build_bsdf_lut(self._h))

    def get_bsdf_lut(self):
        c = np.zeros(32 * 32, np.uint16)
        g = np.zeros(32 * 32, np.uint16)
        d = np.zeros(32 ** 3, np.uint16)
        di = np.zeros(32 ** 3, np.uint16)
        p = lambda a: a.ctypes.data_as(C.POINTER(C.c_uint16))
        _check(self._lib.lumb200_device_get_bsdf_lut(self._h, p(c), p(g), p(d), p(di)))
        return c, g, d, di

    def set_bsdf_lut(self, c, g, d, di) -> None:
        arrs = [np.ascontiguousarray(a, dtype=np.uint16).reshape(-1) for a in (c, g, d, di)]
        p = lambda a: a.ctypes.data_as(C.POINTER(C.c_uint16))
        _check(self._lib.lumb200_device_set_bsdf_lut(self._h, *[p(a) for a in arrs]))

    # -- rendering ----------------------------------------------------------------------------
    def start_render(self) -> None:
        _check(self._lib.lumb200_device_start_render(self._h))

    def render_samples(self, first_sample_id: int, count: int, stride: int = 1) -> None:
        _check(self._lib.lumb200_device_render_samples(self._h, C.c_uint32(first_sample_id), C.c_uint32(count), C.c_uint32(stride)))

    def update_adaptive_sampling(self, enable: bool = True, max_sampling_rate: int = 256, avg_sampling_rate: int = 2, update_interval: int = 64,
                                 exposure_aware: bool = True, exposure: float = 1.0, tonemap: int = 4, agx=(1.0, 1.0, 1.0), output_mode: int = 0) -> None:
        """Adaptive sampler settings (reference defaults, settings.c:15-19); latched by the next start_render."""
        p = AdaptiveSampling(1 if enable else 0, max_sampling_rate, avg_sampling_rate, update_interval, 1 if exposure_aware else 0, exposure, tonemap,
                             agx[0], agx[1], agx[2], output_mode)
        _check(self._lib.lumb200_device_update_adaptive_sampling(self._h, C.byref(p)))

    def render_executions(self, count: int) -> None:
        _check(self._lib.lumb200_device_render_executions(self._h, C.c_uint32(count)))

    def adaptive_state(self) -> Dict:
        st = AdaptiveState()
        _check(self._lib.lumb200_device_get_adaptive_state(self._h, C.byref(st)))
        return dict(stage_id=st.stage_id, executions=list(st.executions), tasks_per_execution=st.tasks_per_execution, blocks_x=st.blocks_x,
                    blocks_y=st.blocks_y, paths_traced=st.paths_traced)

    def set_adaptive_state(self, stage_id: int, executions, words=None) -> None:
        ex = (C.c_uint32 * 5)(*[int(e) for e in executions])
        if words is None:
            _check(self._lib.lumb200_device_set_adaptive_state(self._h, C.c_uint32(stage_id), ex, None, C.c_size_t(0)))
        else:
            w = np.ascontiguousarray(words, np.uint32).reshape(-1)
            _check(self._lib.lumb200_device_set_adaptive_state(self._h, C.c_uint32(stage_id), ex, w.ctypes.data_as(C.POINTER(C.c_uint32)),
                                                               C.c_size_t(w.size)))

    def render_allocated_execution(self, executions_before) -> None:
        ex = (C.c_uint32 * 5)(*[int(e) for e in executions_before])
        _check(self._lib.lumb200_device_render_allocated_execution(self._h, ex))

    def build_adaptive_stage(self) -> None:
        _check(self._lib.lumb200_device_build_adaptive_stage(self._h))

    def adaptive_words(self) -> np.ndarray:
        st = self.adaptive_state()
        out = np.zeros(st["blocks_x"] * st["blocks_y"], np.uint32)
        _check(self._lib.lumb200_device_download_adaptive_words(self._h, out.ctypes.data_as(C.POINTER(C.c_uint32)), C.c_size_t(out.size)))
        return out.reshape(st["blocks_y"], st["blocks_x"])

    def sync(self) -> None:
        _check(self._lib.lumb200_device_sync(self._h))

    def frame_planes_ptr(self):
        ptr = C.c_void_p()
        n = C.c_size_t()
        _check(self._lib.lumb200_device_get_frame_planes(self._h, C.byref(ptr), C.byref(n)))
        return ptr.value, n.value

    def bind_frame_planes(self, device_ptr: int, num_floats: int) -> None:
        _check(self._lib.lumb200_device_bind_frame_planes(self._h, C.c_void_p(device_ptr), C.c_size_t(num_floats)))

    def download_frame_planes(self) -> np.ndarray:
        out = np.empty(4 * self.width * self.height, dtype=np.float32)
        _check(self._lib.lumb200_device_download_frame_planes(self._h, _fptr(out), C.c_size_t(out.size)))
        return out.reshape(4, self.height, self.width)

    def download_result(self, sample_count: int) -> np.ndarray:
        out = np.empty(3 * self.width * self.height, dtype=np.float32)
        _check(self._lib.lumb200_device_download_result(self._h, C.c_uint32(sample_count), _fptr(out)))
        return out.reshape(3, self.height, self.width)

    def download_result_into(self, sample_count: int, host_ptr: int) -> None:
        """Same as download_result but into caller-owned host memory of 3 * width * height floats (e.g. a pinned buffer,
        which lets the D2H copy run at PCIe speed instead of being staged through pageable memory)."""
        _check(self._lib.lumb200_device_download_result(self._h, C.c_uint32(sample_count), C.cast(C.c_void_p(host_ptr), C.POINTER(C.c_float))))

    def download_result_async(self, sample_count: int, host_ptr: int, slot: int) -> None:
        """Resolves on the render stream and copies to host memory (3 * width * height floats, ideally pinned) on a second stream;
        wait_download(slot) before reading the buffer or reusing the slot."""
        _check(self._lib.lumb200_device_download_result_async(self._h, C.c_uint32(sample_count), C.cast(C.c_void_p(host_ptr), C.POINTER(C.c_float)),
                                                              C.c_uint32(slot)))

    def wait_download(self, slot: int) -> None:
        _check(self._lib.lumb200_device_wait_download(self._h, C.c_uint32(slot)))

    # -- parity / measurement hooks -----------------------------------------------------------------
    def load_bluenoise_1d(self, table: np.ndarray) -> None:
        t = np.ascontiguousarray(table, dtype=np.uint16)
        _check(self._lib.lumb200_device_load_bluenoise_1d(self._h, t.ctypes.data_as(C.POINTER(C.c_uint16)), C.c_size_t(t.size)))

    def download_output_argb8(self, sample_count: int, exposure: float = 1.0, tonemap: int = 0, agx=(1.0, 1.0, 1.0), dithering: bool = False,
                              purkinje=None, supersampling: int = 0, bloom_blend: float = 0.0, local_error_minimization: bool = False,
                              filter: int = 0, color_correction=None, film_grain: float = 0.0) -> np.ndarray:
        """ARGB8 output image (height >> s, width >> s, 4) with byte order b, g, r, a (LuminaryARGB8); purkinje = (kappa1, kappa2);
        bloom_blend > 0 runs the mip-chain bloom on the mean radiance first."""
        op = OutputParams(exposure, tonemap, agx[0], agx[1], agx[2], 1 if dithering else 0, 1 if purkinje else 0,
                          purkinje[0] if purkinje else 0.0, purkinje[1] if purkinje else 0.0, supersampling, 1 if local_error_minimization else 0,
                          bloom_blend, filter, 1 if color_correction is not None else 0,
                          (C.c_float * 3)(*(color_correction if color_correction is not None else (0.0, 0.0, 0.0))), film_grain)
        out = np.empty((self.height >> supersampling, self.width >> supersampling, 4), dtype=np.uint8)
        _check(self._lib.lumb200_device_download_output_argb8(self._h, C.c_uint32(sample_count), C.byref(op), out.ctypes.data_as(C.POINTER(C.c_uint8))))
        return out

    def add_planes_from(self, other: "Device") -> None:
        _check(self._lib.lumb200_device_add_planes_from(self._h, other._h))

    def trace_primary(self, sample_id: int = 0):
        n = self.width * self.height
        inst = np.empty(n, np.uint32)
        tri = np.empty(n, np.uint32)
        t = np.empty(n, np.float32)
        u = np.empty(n, np.float32)
        v = np.empty(n, np.float32)
        up = lambda a: a.ctypes.data_as(C.POINTER(C.c_uint32))
        _check(self._lib.lumb200_device_trace_primary(self._h, C.c_uint32(sample_id), up(inst), up(tri), _fptr(t), _fptr(u), _fptr(v)))
        return inst, tri, t, u, v

    def trace_rays(self, origins: np.ndarray, directions: np.ndarray):
        o = np.ascontiguousarray(origins, dtype=np.float32).reshape(-1, 3)
        d = np.ascontiguousarray(directions, dtype=np.float32).reshape(-1, 3)
        n = o.shape[0]
        inst = np.empty(n, np.uint32)
        tri = np.empty(n, np.uint32)
        t = np.empty(n, np.float32)
        u = np.empty(n, np.float32)
        v = np.empty(n, np.float32)
        up = lambda a: a.ctypes.data_as(C.POINTER(C.c_uint32))
        _check(self._lib.lumb200_device_trace_rays(self._h, _fptr(o), _fptr(d), C.c_uint32(n), up(inst), up(tri), _fptr(t), _fptr(u), _fptr(v)))
        return inst, tri, t, u, v

    def shade_vertices(self, vertices: np.ndarray, sample_id: int, rng_depth: int, is_last: bool = False) -> np.ndarray:
        """Runs sort -> k_shade -> k_trace_shadow on caller-supplied path vertices (VERTEX_IN records); returns VERTEX_OUT records."""
        vin = np.ascontiguousarray(vertices, VERTEX_IN)
        out = np.zeros(vin.size, VERTEX_OUT)
        _check(self._lib.lumb200_device_shade_vertices(self._h, C.c_uint32(sample_id), C.c_uint32(rng_depth), C.c_uint32(1 if is_last else 0),
                                                       C.c_void_p(vin.ctypes.data), C.c_uint32(vin.size), C.c_void_p(out.ctypes.data)))
        return out

    def trace_shadow_rays(self, origins, directions, max_dist, ignore_prims, target_prims) -> np.ndarray:
        """Transmittance (n, 3) of explicit shadow segments through k_trace_shadow."""
        o = np.ascontiguousarray(origins, np.float32).reshape(-1, 3)
        d = np.ascontiguousarray(directions, np.float32).reshape(-1, 3)
        m = np.ascontiguousarray(max_dist, np.float32).reshape(-1)
        ig = np.ascontiguousarray(ignore_prims, np.uint32).reshape(-1)
        tg = np.ascontiguousarray(target_prims, np.uint32).reshape(-1)
        n = o.shape[0]
        assert d.shape[0] == n and m.size == n and ig.size == n and tg.size == n
        vis = np.zeros((n, 3), np.float32)
        up = lambda a: a.ctypes.data_as(C.POINTER(C.c_uint32))
        _check(self._lib.lumb200_device_trace_shadow_rays(self._h, _fptr(o), _fptr(d), _fptr(m), up(ig), up(tg), C.c_uint32(n), _fptr(vis)))
        return vis

    def query_pixel(self, x: int, y: int, sample_id: int = 0) -> Dict:
        inst, tri, depth = C.c_uint32(), C.c_uint32(), C.c_float()
        ray = (C.c_float * 3)()
        _check(self._lib.lumb200_device_query_pixel(self._h, C.c_uint32(x), C.c_uint32(y), C.c_uint32(sample_id), C.byref(inst), C.byref(tri),
                                                    C.byref(depth), ray))
        return dict(instance=inst.value, tri=tri.value, depth=depth.value, ray=(ray[0], ray[1], ray[2]))

    def time_primary_trace(self, sample_id: int = 0, repeats: int = 10) -> float:
        ms = C.c_float(0)
        _check(self._lib.lumb200_device_time_primary_trace(self._h, C.c_uint32(sample_id), C.c_uint32(repeats), C.byref(ms)))
        return ms.value

    def download_bvh(self, which: int = 0):
        """Returns (nodes as (N, 80) uint8, triangles as (T, 12) float32) of the scene (0) or emitter (1) BVH8."""
        nn, nt = C.c_uint32(0), C.c_uint32(0)
        _check(self._lib.lumb200_device_download_bvh(self._h, C.c_uint32(which), None, C.c_size_t(0), None, C.c_size_t(0), C.byref(nn), C.byref(nt)))
        nodes = np.zeros((max(nn.value, 1), 80), np.uint8)
        tris = np.zeros((max(nt.value, 1), 12), np.float32)
        _check(self._lib.lumb200_device_download_bvh(self._h, C.c_uint32(which), nodes.ctypes.data_as(C.c_void_p), C.c_size_t(nn.value), _fptr(tris),
                                                     C.c_size_t(nt.value), C.byref(nn), C.byref(nt)))
        return nodes[:nn.value], tris[:nt.value]

    def stats(self) -> Dict:
        s = Stats()
        _check(self._lib.lumb200_device_get_stats(self._h, C.byref(s)))
        return {k: getattr(s, k) for k, _ in Stats._fields_}

    def set_profiling(self, enable: bool) -> None:
        _check(self._lib.lumb200_device_set_profiling(self._h, C.c_uint32(1 if enable else 0)))

    def profile(self) -> Dict:
        p = Profile()
        _check(self._lib.lumb200_device_get_profile(self._h, C.byref(p)))
        return {name: dict(ms=p.milliseconds[i], launches=int(p.launches[i])) for i, name in enumerate(KERNEL_CLASSES)}

    def measure_traversal(self, sample_id: int = 0) -> Dict:
        t = TraversalStats()
        _check(self._lib.lumb200_device_measure_traversal(self._h, C.c_uint32(sample_id), C.byref(t)))
        return {k: int(getattr(t, k)) for k, _ in TraversalStats._fields_}

    def stream(self) -> int:
        p = C.c_void_p()
        _check(self._lib.lumb200_device_get_stream(self._h, C.byref(p)))
        return p.value or 0
