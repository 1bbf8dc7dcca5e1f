"""ctypes access to the C host layer (libluminary_b200.so) for the tests: loaders, PNG writer, public API symbols."""
import ctypes as C
import os
import re
import struct
import zlib

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB_PATH = os.path.join(ROOT, "luminary_b200", "libluminary_b200.so")
CLI_PATH = os.path.join(ROOT, "luminary_b200", "LuminaryB200")
HEADER = os.path.join(ROOT, "include", "luminary", "luminary.h")


class Vec3(C.Structure):
    _fields_ = [("x", C.c_float), ("y", C.c_float), ("z", C.c_float)]


class RGBF(C.Structure):
    _fields_ = [("r", C.c_float), ("g", C.c_float), ("b", C.c_float)]


class RGBAF(C.Structure):
    _fields_ = [("r", C.c_float), ("g", C.c_float), ("b", C.c_float), ("a", C.c_float)]


class Material(C.Structure):  # LuminaryMaterial
    _fields_ = [("id", C.c_uint32), ("base_substrate", C.c_uint32), ("albedo", RGBAF), ("emission", RGBF), ("emission_scale", C.c_float),
                ("roughness", C.c_float), ("roughness_clamp", C.c_float), ("refraction_index", C.c_float), ("emission_active", C.c_bool),
                ("thin_walled", C.c_bool), ("metallic", C.c_bool), ("colored_transparency", C.c_bool), ("roughness_as_smoothness", C.c_bool),
                ("normal_map_is_compressed", C.c_bool), ("bidirectional_emission", C.c_bool), ("albedo_tex", C.c_uint16),
                ("luminance_tex", C.c_uint16), ("roughness_tex", C.c_uint16), ("metallic_tex", C.c_uint16), ("normal_tex", C.c_uint16)]


class Settings(C.Structure):  # LuminaryRendererSettings
    _fields_ = [("width", C.c_uint32), ("height", C.c_uint32), ("max_ray_depth", C.c_uint32), ("bridge_max_num_vertices", C.c_uint32),
                ("undersampling", C.c_uint32), ("supersampling", C.c_uint32), ("enable_adaptive_sampling", C.c_bool),
                ("adaptive_sampling_max_sampling_rate", C.c_uint32), ("adaptive_sampling_avg_sampling_rate", C.c_uint32),
                ("adaptive_sampling_update_interval", C.c_uint32), ("adaptive_sampling_exposure_aware", C.c_bool),
                ("adaptive_sampling_output_mode", C.c_uint32), ("shading_mode", C.c_uint32), ("region_x", C.c_float), ("region_y", C.c_float),
                ("region_width", C.c_float), ("region_height", C.c_float)]


class ThinLens(C.Structure):
    _fields_ = [("fov", C.c_float), ("aperture_size", C.c_float)]


class Physical(C.Structure):
    _fields_ = [("allow_reflections", C.c_bool), ("use_spectral_rendering", C.c_bool)] + [(n, C.c_float) for n in (
        "focal_length", "front_focal_point", "back_focal_point", "front_principal_point", "back_principal_point", "aperture_point",
        "aperture_diameter", "exit_pupil_point", "exit_pupil_diameter", "image_plane_distance", "sensor_width")]


class Camera(C.Structure):  # LuminaryCamera
    _fields_ = [("pos", Vec3), ("rotation", Vec3), ("aperture_shape", C.c_uint32), ("aperture_blade_count", C.c_uint32), ("exposure", C.c_float),
                ("tonemap", C.c_uint32), ("agx_custom_slope", C.c_float), ("agx_custom_power", C.c_float), ("agx_custom_saturation", C.c_float),
                ("filter", C.c_uint32), ("use_local_error_minimization", C.c_bool), ("bloom_blend", C.c_float), ("dithering", C.c_bool),
                ("purkinje", C.c_bool), ("purkinje_kappa1", C.c_float), ("purkinje_kappa2", C.c_float), ("wasd_speed", C.c_float),
                ("mouse_speed", C.c_float), ("smooth_movement", C.c_bool), ("smoothing_factor", C.c_float),
                ("russian_roulette_threshold", C.c_float), ("use_color_correction", C.c_bool), ("color_correction", RGBF),
                ("film_grain", C.c_float), ("camera_scale", C.c_float), ("object_distance", C.c_float), ("use_physical_camera", C.c_bool),
                ("thin_lens", ThinLens), ("physical", Physical)]


class Sky(C.Structure):  # LuminarySky
    _fields_ = [("geometry_offset", Vec3), ("azimuth", C.c_float), ("altitude", C.c_float), ("moon_azimuth", C.c_float),
                ("moon_altitude", C.c_float), ("moon_tex_offset", C.c_float), ("sun_strength", C.c_float), ("base_density", C.c_float),
                ("ozone_absorption", C.c_bool), ("steps", C.c_uint32), ("stars_count", C.c_uint32), ("stars_seed", C.c_uint32),
                ("stars_intensity", C.c_float), ("rayleigh_density", C.c_float), ("mie_density", C.c_float), ("ozone_density", C.c_float),
                ("rayleigh_falloff", C.c_float), ("mie_falloff", C.c_float), ("mie_diameter", C.c_float), ("ground_visibility", C.c_float),
                ("ozone_layer_thickness", C.c_float), ("multiscattering_factor", C.c_float), ("hdri_dim", C.c_uint32), ("hdri_samples", C.c_uint32),
                ("aerial_perspective", C.c_bool), ("constant_color", RGBF), ("mode", C.c_uint32)]


class WavefrontArgs(C.Structure):
    _fields_ = [("legacy_smoothness", C.c_bool), ("force_transparency_cutout", C.c_bool), ("emission_scale", C.c_float),
                ("force_bidirectional_emission", C.c_bool)]


class HostMesh(C.Structure):
    _fields_ = [("triangle_count", C.c_uint32), ("vertex_buffer", C.POINTER(C.c_float)), ("normal_buffer", C.POINTER(C.c_float)),
                ("uv_buffer", C.POINTER(C.c_float)), ("material_id_buffer", C.POINTER(C.c_uint16))]


class FileContent(C.Structure):
    _fields_ = [("settings", Settings), ("camera", Camera), ("sky", Sky), ("wavefront_args", WavefrontArgs),
                ("mesh_files", C.POINTER(C.c_char_p)), ("num_mesh_files", C.c_uint32)]


_lib = None


def lib() -> C.CDLL:
    global _lib
    if _lib is None:
        _lib = C.CDLL(LIB_PATH)
        _lib.luminary_b200_last_error.restype = C.c_char_p
        _lib.luminary_result_to_string.restype = C.c_char_p
        for n in ("lum_wavefront_load", "lum_file_read", "lum_png_write_argb8", "lum_png_read"):
            getattr(_lib, n).restype = C.c_uint64
        _lib.lum_host_mesh_free.restype = None
        _lib.lum_host_texture_free.restype = None
        _lib.lum_file_content_init.restype = None
        _lib.lum_file_content_free.restype = None
    return _lib


def declared_api_functions():
    """Names of every function the public headers (include/luminary/*.h) declare."""
    inc = os.path.dirname(HEADER)
    text = "".join(open(os.path.join(inc, f)).read() for f in sorted(os.listdir(inc)))
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(luminary_[a-z0-9_]+)\s*\(", text)))


def material_dict(m: Material) -> dict:
    return dict(base_substrate=int(m.base_substrate), albedo=(m.albedo.r, m.albedo.g, m.albedo.b, m.albedo.a),
                emission=(m.emission.r, m.emission.g, m.emission.b), emission_scale=m.emission_scale, roughness=m.roughness,
                roughness_clamp=m.roughness_clamp, refraction_index=m.refraction_index, emission_active=bool(m.emission_active),
                thin_walled=bool(m.thin_walled), metallic=bool(m.metallic), colored_transparency=bool(m.colored_transparency),
                roughness_as_smoothness=bool(m.roughness_as_smoothness), normal_map_is_compressed=bool(m.normal_map_is_compressed),
                bidirectional_emission=bool(m.bidirectional_emission), albedo_tex=int(m.albedo_tex), luminance_tex=int(m.luminance_tex),
                roughness_tex=int(m.roughness_tex), metallic_tex=int(m.metallic_tex), normal_tex=int(m.normal_tex))


class HostTexture(C.Structure):  # LumHostTexture
    _fields_ = [("width", C.c_uint32), ("height", C.c_uint32), ("pitch", C.c_uint32), ("type", C.c_uint32), ("num_components", C.c_uint32),
                ("gamma", C.c_float), ("data", C.c_void_p)]


def texture_dict(t: HostTexture) -> dict:
    """-> the dict form luminary_b200.scenes / api use (data None = invalid texture)"""
    if not t.data:
        return dict(data=None, gamma=float(t.gamma))
    dt = np.uint8 if t.type == 1 else np.uint16
    n = t.width * t.height * t.num_components
    a = np.ctypeslib.as_array(C.cast(t.data, C.POINTER(C.c_uint8 if t.type == 1 else C.c_uint16)), shape=(n,)).copy()
    return dict(data=a.astype(dt).reshape(t.height, t.width, t.num_components), wrap_u=0, wrap_v=0, filter=1, gamma=float(t.gamma))


def png_read(path: str):
    """lum_png_read -> (code, texture dict or None)"""
    L = lib()
    t = HostTexture()
    code = L.lum_png_read(path.encode(), C.byref(t))
    if code != 0:
        return code, None
    d = texture_dict(t)
    L.lum_host_texture_free(C.byref(t))
    return code, d


last_textures = []  # textures of the most recent wavefront_load call (list of dicts)


def wavefront_load(path: str, material_offset: int = 0, emission_scale: float = 1.0, bidirectional: bool = False, texture_offset: int = 0):
    """-> (code, has_mesh, vertex (T,3,3), normal (T,3,3), uv (T,3,2), material (T,), [material dicts], [ids]); the textures
    of the file are left in host_c.last_textures"""
    global last_textures
    last_textures = []
    L = lib()
    args = WavefrontArgs(False, False, emission_scale, bidirectional)
    mesh = HostMesh()
    has = C.c_bool(False)
    mats = C.POINTER(Material)()
    nm = C.c_uint32(0)
    texs = C.POINTER(HostTexture)()
    nt = C.c_uint32(0)
    code = L.lum_wavefront_load(path.encode(), args, C.c_uint32(material_offset), C.c_uint32(texture_offset), C.byref(mesh), C.byref(has),
                                C.byref(mats), C.byref(nm), C.byref(texs), C.byref(nt))
    if code != 0 or not has.value:
        return code, False, None, None, None, None, [], []
    t = mesh.triangle_count
    v = np.ctypeslib.as_array(mesh.vertex_buffer, shape=(t * 9,)).copy().reshape(t, 3, 3) if t else np.zeros((0, 3, 3), np.float32)
    n = np.ctypeslib.as_array(mesh.normal_buffer, shape=(t * 9,)).copy().reshape(t, 3, 3) if t else np.zeros((0, 3, 3), np.float32)
    uv = np.ctypeslib.as_array(mesh.uv_buffer, shape=(t * 6,)).copy().reshape(t, 3, 2) if t else np.zeros((0, 3, 2), np.float32)
    mid = np.ctypeslib.as_array(mesh.material_id_buffer, shape=(t,)).copy() if t else np.zeros((0,), np.uint16)
    md = [material_dict(mats[k]) for k in range(nm.value)]
    ids = [int(mats[k].id) for k in range(nm.value)]
    L.lum_host_mesh_free(C.byref(mesh))
    C.CDLL(None).free(mats)
    for k in range(nt.value):
        last_textures.append(texture_dict(texs[k]))
        L.lum_host_texture_free(C.byref(texs[k]))
    if nt.value:
        C.CDLL(None).free(texs)
    return code, True, v, n, uv, mid, md, ids


def lum_read(path: str):
    L = lib()
    c = FileContent()
    L.lum_file_content_init(C.byref(c))
    code = L.lum_file_read(path.encode(), C.byref(c))
    files = [c.mesh_files[k].decode() for k in range(c.num_mesh_files)] if code == 0 else []
    out = dict(code=code, settings=c.settings, camera=c.camera, sky=c.sky, args=c.wavefront_args, mesh_files=files)
    # copy the PODs before freeing
    out["settings"] = Settings.from_buffer_copy(c.settings)
    out["camera"] = Camera.from_buffer_copy(c.camera)
    out["sky"] = Sky.from_buffer_copy(c.sky)
    out["args"] = WavefrontArgs.from_buffer_copy(c.wavefront_args)
    L.lum_file_content_free(C.byref(c))
    return out


def png_decode_rgba(path: str) -> np.ndarray:
    """Decodes an 8-bit RGBA, non-interlaced PNG whose scanlines use filter 0 (what lum_png.c writes) -> (H, W, 4)."""
    data = open(path, "rb").read()
    assert data[:8] == b"\x89PNG\r\n\x1a\n"
    pos, idat, w, h = 8, b"", 0, 0
    while pos < len(data):
        (n,) = struct.unpack(">I", data[pos:pos + 4])
        typ = data[pos + 4:pos + 8]
        body = data[pos + 8:pos + 8 + n]
        (crc,) = struct.unpack(">I", data[pos + 8 + n:pos + 12 + n])
        assert zlib.crc32(typ + body) & 0xFFFFFFFF == crc, "bad chunk CRC"
        if typ == b"IHDR":
            w, h, depth, ctype, comp, flt, inter = struct.unpack(">IIBBBBB", body)
            assert (depth, ctype, comp, flt, inter) == (8, 6, 0, 0, 0)
        elif typ == b"IDAT":
            idat += body
        pos += 12 + n
    raw = np.frombuffer(zlib.decompress(idat), dtype=np.uint8).reshape(h, 1 + 4 * w)
    assert (raw[:, 0] == 0).all()
    return raw[:, 1:].reshape(h, w, 4).copy()


def write_lum(path: str, scene, obj_name: str, tonemap: int = 0, dither: int = 0, exposure: float = 1.0, bloom: float = 0.0) -> None:
    """A version-4 scene file for `scene` (camera, resolution, depth, constant sky) referencing obj_name."""
    c = scene.camera
    with open(path, "w") as f:
        f.write("Luminary\nVERSION 4\n# written by tests/host_c.py\n")
        f.write(f"GENERAL WIDTH___ {scene.width}\nGENERAL HEIGHT__ {scene.height}\nGENERAL BOUNCES_ {scene.max_ray_depth}\n")
        f.write(f"GENERAL MESHFILE {obj_name}\n")
        f.write("CAMERA POSITION %.9g %.9g %.9g\n" % tuple(c["pos"]))
        f.write("CAMERA ROTATION %.9g %.9g %.9g\n" % tuple(c["rotation"]))
        f.write("CAMERA FOV_____ %.9g\nCAMERA FOCALLEN %.9g\nCAMERA APERTURE %.9g\n" % (c["fov"], c["object_distance"], c["aperture_size"]))
        f.write("CAMERA EXPOSURE %.9g\nCAMERA TONEMAP_ %d\nCAMERA DITHER__ %d\nCAMERA PURKINJE 0\nCAMERA BLOOMBLE %.9g\n" % (exposure, tonemap, dither, bloom))
        f.write("CAMERA RUSSIANR %.9g\n" % c["russian_roulette_threshold"])
        f.write("SKY MODE____ %d\nSKY COLORCON %.9g %.9g %.9g\n" % ((scene.sky_mode,) + tuple(scene.sky_color)))
        keys = dict(azimuth="AZIMUTH_", altitude="ALTITUDE", moon_azimuth="MOONAZIM", moon_altitude="MOONALTI", sun_strength="SUNSTREN",
                    base_density="DENSITY_", steps="STEPS___", stars_seed="STARSEED", stars_count="STARNUM_", stars_intensity="STARINTE",
                    ozone_absorption="OZONEABS", rayleigh_density="RAYLEDEN", mie_density="MIEDENSI", ozone_density="OZONEDEN",
                    rayleigh_falloff="RAYLEFAL", mie_falloff="MIEFALLO", ground_visibility="GROUNDVI", mie_diameter="DIAMETER",
                    ozone_layer_thickness="OZONETHI", multiscattering_factor="MSFACTOR", hdri_dim="HDRIDIM_", hdri_samples="HDRISAMP")
        for k, v in (getattr(scene, "sky", None) or {}).items():
            if k == "geometry_offset":
                f.write("SKY OFFSET__ %.9g %.9g %.9g\n" % tuple(v))
            else:
                f.write(("SKY %s %d\n" if isinstance(v, int) else "SKY %s %.9g\n") % (keys[k], v))
