"""ctypes access to oracle/_ref/libref_host.so: the REFERENCE's own host C code (struct packers, light-tree build) compiled
from /root/reference by oracle/ref/Makefile. Test infrastructure only. Tests that use it skip when the library was not built
(it can only be built where /root/reference exists; the built .so travels to the GPU box with the snapshot)."""
import ctypes as C
import os

import numpy as np

from luminary_b200 import api

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB_PATH = os.path.join(ROOT, "oracle", "_ref", "libref_host.so")
_lib = None


def available() -> bool:
    return os.path.exists(LIB_PATH)


def lib():
    global _lib
    if _lib is None:
        _lib = C.CDLL(LIB_PATH)
        _lib.refhost_pack_normal.restype = C.c_uint32
        _lib.refhost_pack_normal.argtypes = [C.c_float, C.c_float, C.c_float]
        _lib.refhost_pack_uv.restype = C.c_uint32
        _lib.refhost_pack_uv.argtypes = [C.c_float, C.c_float]
        _lib.refhost_sizeof_device_camera.restype = C.c_size_t
        _lib.refhost_free.argtypes = [C.c_void_p]
    return _lib


def material_convert(m: dict) -> bytes:
    out = C.create_string_buffer(32)
    s = api.material_struct(m)
    assert lib().refhost_material_convert(C.byref(s), out) == 0
    return out.raw


def instance_transform_convert(translation, rotation, scale) -> bytes:
    ins = api.Instance()
    ins.mesh_id = 0
    ins.translation[:] = translation
    ins.rotation[:] = rotation
    ins.scale[:] = scale
    ins.active = 1
    out = C.create_string_buffer(32)
    assert lib().refhost_instance_transform_convert(C.byref(ins), out) == 0
    return out.raw


def camera_struct(cam: dict) -> api.Camera:
    c = api.Camera()
    c.pos[:] = cam["pos"]
    c.rotation[:] = cam["rotation"]
    c.fov = cam["fov"]
    c.aperture_size = cam.get("aperture_size", 0.0)
    c.object_distance = cam.get("object_distance", 1.0)
    c.camera_scale = cam.get("camera_scale", 1.0)
    c.russian_roulette_threshold = cam.get("russian_roulette_threshold", 0.1)
    c.aperture_shape = cam.get("aperture_shape", 0)
    c.aperture_blade_count = cam.get("aperture_blade_count", 7)
    return c


def camera_convert(cam: dict) -> bytes:
    n = lib().refhost_sizeof_device_camera()
    out = C.create_string_buffer(n)
    c = camera_struct(cam)
    assert lib().refhost_camera_convert(C.byref(c), out, C.c_size_t(n)) == 0
    return out.raw


def _mesh_struct(m, keep):
    v = np.ascontiguousarray(m.vertex, np.float32).reshape(-1)
    n = np.ascontiguousarray(m.normal, np.float32).reshape(-1)
    t = np.ascontiguousarray(m.uv, np.float32).reshape(-1)
    mm = np.ascontiguousarray(m.material, np.uint16).reshape(-1)
    keep += [v, n, t, mm]
    return api.Mesh(m.num_tris, api._fptr(v), api._fptr(n), api._fptr(t), mm.ctypes.data_as(C.POINTER(C.c_uint16)))


def mesh_convert(m):
    """(vertices uint32[num_tris*3, 4], textris uint32[num_tris, 4]) as device_struct_vertex_convert / _triangle_texture_convert pack them."""
    keep = []
    ms = _mesh_struct(m, keep)
    verts = np.zeros((m.num_tris * 3, 4), np.uint32)
    tex = np.zeros((m.num_tris, 4), np.uint32)
    assert lib().refhost_mesh_convert(C.byref(ms), verts.ctypes.data_as(C.c_void_p), tex.ctypes.data_as(C.c_void_p)) == 0
    return verts, tex


def build_light_tree(scene):
    """The reference's light_tree_build on a scenes.Scene: (root bytes, nodes bytes, handles uint32[n,2], emitter vertices float32[n,3,4])."""
    keep = []
    meshes = (api.Mesh * max(len(scene.meshes), 1))()
    for i, m in enumerate(scene.meshes):
        meshes[i] = _mesh_struct(m, keep)
    inst = (api.Instance * max(len(scene.instances), 1))()
    for i, ins in enumerate(scene.instances):
        inst[i].mesh_id = ins.mesh_id
        inst[i].translation[:] = ins.translation
        inst[i].rotation[:] = ins.rotation
        inst[i].scale[:] = ins.scale
        inst[i].active = 1 if ins.active else 0
    mats = (api.Material * max(len(scene.materials), 1))()
    for i, m in enumerate(scene.materials):
        mats[i] = api.material_struct(m)
    out = api.LightTreeBuffers()
    bvh = C.c_void_p()
    bvh_size = C.c_size_t(0)
    rc = lib().refhost_light_tree_build(meshes, C.c_uint32(len(scene.meshes)), inst, C.c_uint32(len(scene.instances)), mats,
                                        C.c_uint32(len(scene.materials)), C.byref(out), C.byref(bvh), C.byref(bvh_size))
    assert rc == 0, rc
    if out.num_lights == 0:
        return None
    root = C.string_at(out.root_data, out.root_size)
    nodes = C.string_at(out.nodes_data, out.nodes_size) if out.nodes_size else b""
    handles = np.ctypeslib.as_array(out.tri_handle_map, shape=(out.num_lights, 2)).copy()
    verts = np.frombuffer(C.string_at(bvh, bvh_size.value), np.float32).reshape(out.num_lights, 3, 4).copy()
    for p in (out.root_data, out.nodes_data, C.cast(out.tri_handle_map, C.c_void_p), bvh):
        if p:
            lib().refhost_free(p)
    return root, nodes, handles, verts


class SkyParams(C.Structure):  # RefSkyParams of oracle/ref/ref_host_shim.c (= OrcSkyParams)
    _fields_ = [("geometry_offset", C.c_float * 3)] + [(n, C.c_float) for n in (
        "azimuth", "altitude", "moon_azimuth", "moon_altitude", "moon_tex_offset", "sun_strength", "base_density", "rayleigh_density", "mie_density",
        "ozone_density", "rayleigh_falloff", "mie_falloff", "mie_diameter", "ground_visibility", "ozone_layer_thickness", "multiscattering_factor",
        "stars_intensity")] + [(n, C.c_uint32) for n in ("steps", "ozone_absorption", "stars_count", "stars_seed", "aerial_perspective")]


def sky_params(sky: dict = None) -> SkyParams:
    """the reference's sky_get_default (sky.c:6-42) overridden by the entries of `sky`"""
    p = SkyParams()
    lib().refhost_sky_default_params(C.byref(p))
    for k, v in (sky or {}).items():
        if k == "geometry_offset":
            p.geometry_offset[:] = v
        elif k not in ("mode", "hdri_dim", "hdri_samples"):
            setattr(p, k, v)
    return p


def sky_convert(sky: dict = None, mode: int = 0, color=(1.0, 1.0, 1.0)) -> bytes:
    """device_struct_sky_convert (device_structs.c:107-172) -> DeviceSky bytes"""
    L = lib()
    L.refhost_sizeof_device_sky.restype = C.c_size_t
    n = L.refhost_sizeof_device_sky()
    out = C.create_string_buffer(n)
    p = sky_params(sky)
    assert L.refhost_sky_convert_params(C.byref(p), C.c_uint32(mode), (C.c_float * 3)(*color), out, C.c_size_t(n)) == 0
    return out.raw


def stars_generate(seed: int, count: int):
    """sky_stars_update (device_sky.c:470-572) -> (stars (count, 4) [altitude, azimuth, radius, intensity], offsets (64 * 32 + 1,))"""
    stars = np.zeros((count, 4), np.float32)
    offsets = np.zeros(64 * 32 + 1, np.uint32)
    assert lib().refhost_stars_generate(C.c_uint32(seed), C.c_uint32(count), stars.ctypes.data_as(C.c_void_p), offsets.ctypes.data_as(C.c_void_p)) == 0
    return stars, offsets
