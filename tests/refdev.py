"""ctypes access to oracle/_ref/librefdev.so: the REFERENCE's own CUDA kernels (tasks_create, geometry_process_tasks,
sky_process_tasks, accumulation_*, bsdf_generate_*_lut) compiled for sm_100a from /root/reference by oracle/ref/ref_patch.sh and
launched unmodified by oracle/ref/ref_dev_harness.cu. Test infrastructure only. It can only be BUILT where /root/reference exists
(this container); the built .so travels to the GPU box with the snapshot. Tests skip when it is absent.

The reference keeps its task records warp-interleaved (cuda/memory.cuh:114-131): a record of C 16-byte chunks that belongs to
(buffer index b, task slot k, warp w, lane l) stores chunk c at float4 element  l + 32 * (c + C * (w + NW * (k + K * b))).
interleave() / deinterleave() convert between that layout and plain arrays [b][k][thread] of records."""
import ctypes as C
import os

import numpy as np

import refhost
from luminary_b200 import api

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB_PATH = os.path.join(ROOT, "oracle", "_ref", "librefdev.so")
THREADS_PER_BLOCK = 128
_lib = None

_V3 = (np.float32, 3)
# DeviceTaskState, device_utils.h:359-405 (80 bytes = 5 chunks)
TASK_STATE = np.dtype([("state", np.uint16), ("path_id", np.uint16, 3), ("origin", *_V3), ("ray", *_V3),
                       ("instance_id", np.uint32), ("tri_id", np.uint32), ("depth", np.float32), ("pad0", np.uint32),
                       ("record", np.uint32, 2), ("results_index", np.uint32), ("pad1", np.uint32),
                       ("ior", np.uint32), ("volume_id_01", np.uint32), ("volume_id_23", np.uint32), ("pad2", np.uint32)])
# DeviceTaskDirectLight, device_utils.h:411-469 (96 bytes = 6 chunks; geometry variant of the union)
DIRECT_LIGHT = np.dtype([("geo_light_id", np.uint32), ("geo_color", *_V3), ("geo_ray", *_V3), ("geo_dist", np.float32),
                         ("bsdf_weight", *_V3), ("bsdf_ray", *_V3), ("bsdf_root_sum", np.float32), ("bsdf_prob", np.float32),
                         ("sun_color", np.uint32, 2), ("sun_ray", np.uint32, 2), ("amb_color", np.uint32, 2), ("amb_ray", np.uint32, 2)])
# DeviceTaskResult, device_utils.h:399-403
RESULT = np.dtype([("color", *_V3), ("index", np.uint32)])
assert TASK_STATE.itemsize == 80 and DIRECT_LIGHT.itemsize == 96 and RESULT.itemsize == 16

SHADING_TASK_INDEX_GEOMETRY, SHADING_TASK_INDEX_SKY, SHADING_TASK_INDEX_TOTAL = 0, 4, 5
PRESORT, POSTSORT = 0, 1


def available() -> bool:
    return os.path.exists(LIB_PATH) and refhost.available()


def lib():
    global _lib
    if _lib is None:
        L = C.CDLL(LIB_PATH)
        L.refdev_sizeof.restype = C.c_size_t
        L.refdev_sizeof.argtypes = [C.c_char_p]
        L.refdev_buffer_size.restype = C.c_size_t
        L.refdev_buffer_size.argtypes = [C.c_char_p]
        L.refdev_set_settings.argtypes = [C.c_char_p, C.c_size_t]
        L.refdev_set_camera.argtypes = [C.c_char_p, C.c_size_t]
        L.refdev_set_sky.argtypes = [C.c_char_p, C.c_size_t]
        L.refdev_set_bluenoise.argtypes = [C.c_void_p, C.c_size_t, C.c_void_p, C.c_size_t]
        L.refdev_set_scene.argtypes = [C.c_uint32, C.POINTER(C.c_void_p), C.POINTER(C.c_void_p), C.c_void_p, C.c_uint32, C.c_void_p, C.c_void_p,
                                       C.c_uint32, C.c_void_p]
        L.refdev_set_light_tree.argtypes = [C.c_char_p, C.c_size_t, C.c_char_p, C.c_size_t, C.c_void_p, C.c_uint32]
        L.refdev_build_bsdf_lut.argtypes = [C.c_void_p] * 4
        L.refdev_configure.argtypes = [C.c_uint32, C.c_uint32]
        L.refdev_set_state.argtypes = [C.c_uint32] * 4
        L.refdev_upload.argtypes = [C.c_char_p, C.c_size_t, C.c_void_p, C.c_size_t]
        L.refdev_download.argtypes = [C.c_char_p, C.c_size_t, C.c_void_p, C.c_size_t]
        L.refdev_clear.argtypes = [C.c_char_p]
        if hasattr(L, "refdev_time_geometry_process_tasks"):
            L.refdev_time_geometry_process_tasks.argtypes = [C.c_int, C.POINTER(C.c_float)]
        assert L.refdev_sizeof(b"DeviceTaskState") == 80 and L.refdev_sizeof(b"DeviceTaskDirectLight") == 96
        _lib = L
    return _lib


def interleave(records: np.ndarray, num_threads: int) -> np.ndarray:
    """records[B][K][num_threads] (structured) -> raw uint32 words in the reference's warp-interleaved order."""
    b, k, t = records.shape
    assert t == num_threads and t % 32 == 0
    chunks = records.dtype.itemsize // 16
    words = np.ascontiguousarray(records).view(np.uint32).reshape(b, k, t // 32, 32, chunks, 4)
    return np.ascontiguousarray(words.transpose(0, 1, 2, 4, 3, 5)).reshape(-1)


def deinterleave(words: np.ndarray, dtype: np.dtype, num_buffers: int, tasks_per_thread: int, num_threads: int) -> np.ndarray:
    chunks = dtype.itemsize // 16
    words = np.ascontiguousarray(words).view(np.uint32)[:num_buffers * tasks_per_thread * num_threads * chunks * 4]  # buffers only ever grow
    w = words.reshape(num_buffers, tasks_per_thread, num_threads // 32, chunks, 32, 4)
    return np.ascontiguousarray(w.transpose(0, 1, 2, 4, 3, 5)).reshape(num_buffers, tasks_per_thread, num_threads, chunks * 4).view(dtype)[..., 0]


def tasks_from_vertices(vin, handles):
    """Oracle path vertices (orc.OracleScene.path_vertices) -> the reference's DeviceTaskState records (device_utils.h:365-420)."""
    t = np.zeros(vin.size, TASK_STATE)
    t["state"] = vin["state"]
    t["path_id"] = vin["path_id"]
    t["origin"] = vin["origin"]
    t["ray"] = vin["ray"]
    t["instance_id"] = handles[vin["prim"], 0]
    t["tri_id"] = handles[vin["prim"], 1]
    t["depth"] = vin["t"]
    t["record"] = vin["record"]
    t["ior"] = vin["medium_ior"]
    return t


class RefDevice:
    """The reference's device state for one luminary_b200.scenes.Scene, packed by the reference's own host code (refhost)."""

    def __init__(self, scene, light_tree="reference", cuda_index: int = 0):
        L = lib()
        self.scene = scene
        assert L.refdev_create(cuda_index) == 0
        s16 = C.create_string_buffer(16)
        assert refhost.lib().refhost_settings_convert(scene.width, scene.height, scene.max_ray_depth, s16) == 0
        assert L.refdev_set_settings(s16.raw, 16) == 0
        cam = refhost.camera_convert(scene.camera)
        assert L.refdev_set_camera(cam, len(cam)) == 0
        n = refhost.lib().refhost_sizeof_device_sky()
        sky = C.create_string_buffer(n)
        col = (C.c_float * 3)(*scene.sky_color)
        self.sky_params = refhost.sky_params(getattr(scene, "sky", None))
        assert refhost.lib().refhost_sky_convert_params(C.byref(self.sky_params), C.c_uint32(scene.sky_mode), col, sky, C.c_size_t(n)) == 0
        assert L.refdev_set_sky(sky.raw, n) == 0
        self.device_sky = sky.raw
        bn1, bn2 = api.load_bluenoise_1d(), api.load_bluenoise_2d()
        assert L.refdev_set_bluenoise(bn1.ctypes.data, bn1.size, bn2.ctypes.data, bn2.size) == 0

        keep = []
        nm = len(scene.meshes)
        vptr, tptr = (C.c_void_p * nm)(), (C.c_void_p * nm)()
        counts = np.zeros(nm, np.uint32)
        for i, m in enumerate(scene.meshes):
            v, t = refhost.mesh_convert(m)
            keep += [v, t]
            vptr[i], tptr[i], counts[i] = v.ctypes.data, t.ctypes.data, m.num_tris
        active = [ins for ins in scene.instances if ins.active]
        trans = b"".join(refhost.instance_transform_convert(i.translation, i.rotation, i.scale) for i in active)
        mesh_ids = np.array([i.mesh_id for i in active], np.uint32)
        mats = b"".join(refhost.material_convert(m) for m in scene.materials)
        assert L.refdev_set_scene(nm, vptr, tptr, counts.ctypes.data, len(active), trans, mesh_ids.ctypes.data, len(scene.materials), mats) == 0

        if light_tree == "reference":
            light_tree = refhost.build_light_tree(scene)
        self.light_tree = light_tree
        if light_tree is not None:
            root, nodes, handles = light_tree[0], light_tree[1], np.ascontiguousarray(light_tree[2], np.uint32)
            assert L.refdev_set_light_tree(root, len(root), nodes, len(nodes), handles.ctypes.data, handles.shape[0]) == 0
        else:
            assert L.refdev_set_light_tree(None, 0, None, 0, None, 0) == 0
        self.num_blocks = self.tasks_per_thread = 0

    def build_bsdf_lut(self):
        out = [np.zeros(32 * 32, np.uint16), np.zeros(32 * 32, np.uint16), np.zeros(32 ** 3, np.uint16), np.zeros(32 ** 3, np.uint16)]
        assert lib().refdev_build_bsdf_lut(*[a.ctypes.data for a in out]) == 0
        return out

    def build_sky_lut(self):
        """The reference's sky_compute_transmittance_lut / sky_compute_multiscattering_lut for the scene's sky; the tables stay bound as
        device.sky_lut_*_tex. -> (tm_low, tm_high (64, 256, 4), ms_low, ms_high (32, 32, 4))"""
        out = [np.zeros((64, 256, 4), np.float32), np.zeros((64, 256, 4), np.float32), np.zeros((32, 32, 4), np.float32), np.zeros((32, 32, 4), np.float32)]
        L = lib()
        L.refdev_build_sky_lut.argtypes = [C.c_void_p] * 4
        assert L.refdev_build_sky_lut(*[a.ctypes.data for a in out]) == 0
        return out

    def build_sky_hdri(self, dim: int, samples: int, origin) -> np.ndarray:
        """The reference's sky_compute_hdri seen from `origin` (world space); the table stays bound as device.sky_hdri_color_tex.
        -> (dim, dim, 4)"""
        out = np.zeros((dim, dim, 4), np.float32)
        L = lib()
        L.refdev_build_sky_hdri.argtypes = [C.c_uint32, C.c_uint32, C.c_float * 3, C.c_void_p]
        assert L.refdev_build_sky_hdri(dim, samples, (C.c_float * 3)(*origin), out.ctypes.data) == 0
        return out

    def set_moon_textures(self, albedo: np.ndarray = None, normal: np.ndarray = None):
        """(H, W, 4) uint8 texels of the moon's surface (as png_load expands the reference's data/moon/*.png), or None = absent"""
        L = lib()
        L.refdev_set_moon_textures.argtypes = [C.c_void_p, C.c_uint32, C.c_uint32, C.c_void_p, C.c_uint32, C.c_uint32]
        a = np.ascontiguousarray(albedo, np.uint8) if albedo is not None else None
        n = np.ascontiguousarray(normal, np.uint8) if normal is not None else None
        assert L.refdev_set_moon_textures(a.ctypes.data if a is not None else None, a.shape[1] if a is not None else 0, a.shape[0] if a is not None else 0,
                                          n.ctypes.data if n is not None else None, n.shape[1] if n is not None else 0, n.shape[0] if n is not None else 0) == 0

    def set_sky_lut(self, tm_low, tm_high, ms_low, ms_high):
        arrs = [np.ascontiguousarray(a, np.float32) for a in (tm_low, tm_high, ms_low, ms_high)]
        L = lib()
        L.refdev_set_sky_lut.argtypes = [C.c_void_p] * 4
        assert L.refdev_set_sky_lut(*[a.ctypes.data for a in arrs]) == 0

    def set_stars(self, stars: np.ndarray = None, offsets: np.ndarray = None):
        """The star catalogue; default: the reference host code's own (sky_stars_update through libref_host.so)."""
        if stars is None:
            stars, offsets = refhost.stars_generate(self.sky_params.stars_seed, self.sky_params.stars_count)
        stars = np.ascontiguousarray(stars, np.float32)
        offsets = np.ascontiguousarray(offsets, np.uint32)
        L = lib()
        L.refdev_set_stars.argtypes = [C.c_void_p, C.c_uint32, C.c_void_p]
        assert L.refdev_set_stars(stars.ctypes.data, stars.shape[0], offsets.ctypes.data) == 0
        return stars, offsets

    def sky(self, tasks: np.ndarray, depth: int) -> np.ndarray:
        """Runs sky_process_tasks on `tasks` (TASK_STATE[n], misses): -> colour written to each task's result record, (n, 3)."""
        T, K = self.num_threads, self.tasks_per_thread
        n = tasks.size
        assert n <= T * K
        slot, thread = np.arange(n) // T, np.arange(n) % T
        post = np.zeros((2, K, T), TASK_STATE)
        tt = tasks.copy()
        tt["results_index"] = thread + slot * T
        post[POSTSORT, slot, thread] = tt
        self.upload("task_states", interleave(post, T))
        res = np.zeros((1, K, T), RESULT)
        res["index"][0, slot, thread] = tt["path_id"][:, 0].astype(np.uint32) + tt["path_id"][:, 1].astype(np.uint32) * self.scene.width
        self.upload("task_results", interleave(res, T))
        counts = np.zeros((SHADING_TASK_INDEX_TOTAL, T), np.uint16)
        counts[SHADING_TASK_INDEX_SKY] = np.bincount(thread, minlength=T)
        self.upload("task_counts", counts)
        self.upload("task_offsets", np.zeros((SHADING_TASK_INDEX_TOTAL, T), np.uint16))
        self.set_state(depth, 0)
        self.launch("sky_process_tasks")
        return self.results()[slot, thread]["color"]

    def inscatter(self, tasks: np.ndarray, depth: int):
        """Runs sky_process_inscattering_events (aerial perspective) on `tasks` (TASK_STATE[n] with trace results, staged as the PRESORT
        tasks the trace kernel leaves): -> (colour added to each task's result record (n, 3), packed throughput after the step (n, 2))."""
        T, K = self.num_threads, self.tasks_per_thread
        n = tasks.size
        assert n <= T * K
        slot, thread = np.arange(n) // T, np.arange(n) % T
        pre = np.zeros((2, K, T), TASK_STATE)
        tt = tasks.copy()
        tt["results_index"] = thread + slot * T
        pre[PRESORT, slot, thread] = tt
        self.upload("task_states", interleave(pre, T))
        res = np.zeros((1, K, T), RESULT)
        res["index"][0, slot, thread] = tt["path_id"][:, 0].astype(np.uint32) + tt["path_id"][:, 1].astype(np.uint32) * self.scene.width
        self.upload("task_results", interleave(res, T))
        self.upload("trace_counts", np.bincount(thread, minlength=T).astype(np.uint16))
        self.set_state(depth, 0)
        self.launch("sky_process_inscattering_events")
        return self.results()[slot, thread]["color"], self.task_states()[PRESORT][slot, thread]["record"]

    def configure(self, num_blocks: int, tasks_per_thread: int):
        assert lib().refdev_configure(num_blocks, tasks_per_thread) == 0
        self.num_blocks, self.tasks_per_thread = num_blocks, tasks_per_thread

    @property
    def num_threads(self) -> int:
        return self.num_blocks * THREADS_PER_BLOCK

    def set_state(self, depth: int, sample_id: int, tile_id: int = 0, accumulated_samples: int = 0):
        assert lib().refdev_set_state(depth, tile_id, sample_id, accumulated_samples) == 0

    def upload(self, name: str, array: np.ndarray, offset: int = 0):
        a = np.ascontiguousarray(array)
        assert lib().refdev_upload(name.encode(), offset, a.ctypes.data, a.nbytes) == 0, name

    def download(self, name: str, dtype=np.uint32) -> np.ndarray:
        n = lib().refdev_buffer_size(name.encode())
        out = np.zeros(n // np.dtype(dtype).itemsize, dtype)
        assert lib().refdev_download(name.encode(), 0, out.ctypes.data, out.nbytes) == 0, name
        return out

    def clear(self, name: str):
        assert lib().refdev_clear(name.encode()) == 0, name

    def launch(self, kernel: str):
        assert getattr(lib(), "refdev_" + kernel)() == 0, kernel

    # ---- record-level helpers -------------------------------------------------------------------------------------
    def task_states(self) -> np.ndarray:
        """[PRESORT | POSTSORT][slot][thread] of TASK_STATE"""
        return deinterleave(self.download("task_states"), TASK_STATE, 2, self.tasks_per_thread, self.num_threads)

    def direct_light(self) -> np.ndarray:
        return deinterleave(self.download("task_direct_light"), DIRECT_LIGHT, 1, self.tasks_per_thread, self.num_threads)[0]

    def results(self) -> np.ndarray:
        return deinterleave(self.download("task_results"), RESULT, 1, self.tasks_per_thread, self.num_threads)[0]

    def _stage_shade(self, tasks: np.ndarray, depth: int):
        """Uploads `tasks` (TASK_STATE[n]) as the POSTSORT geometry tasks of one wavefront iteration: task i goes to thread i % T,
        slot i // T; results_index is assigned here."""
        T, K = self.num_threads, self.tasks_per_thread
        n = tasks.size
        assert n <= T * K
        slot, thread = np.arange(n) // T, np.arange(n) % T
        post = np.zeros((2, K, T), TASK_STATE)
        tt = tasks.copy()
        tt["results_index"] = thread + slot * T  # task_get_base_address<DeviceTaskResult>(slot, RESULT)
        post[POSTSORT, slot, thread] = tt
        self.upload("task_states", interleave(post, T))
        res = np.zeros((1, K, T), RESULT)
        res["index"][0, slot, thread] = tt["path_id"][:, 0].astype(np.uint32) + tt["path_id"][:, 1].astype(np.uint32) * self.scene.width
        self.upload("task_results", interleave(res, T))
        self.clear("task_direct_light")
        counts = np.zeros((SHADING_TASK_INDEX_TOTAL, T), np.uint16)
        counts[SHADING_TASK_INDEX_GEOMETRY] = np.bincount(thread, minlength=T)
        self.upload("task_counts", counts)
        self.upload("task_offsets", np.zeros((SHADING_TASK_INDEX_TOTAL, T), np.uint16))
        self.upload("trace_counts", np.zeros(T, np.uint16))
        self.set_state(depth, 0)
        return slot, thread

    def shade(self, tasks: np.ndarray, depth: int):
        """Runs geometry_process_tasks on `tasks`. Returns (direct_light[n], results[n] (emission only), bounce PRESORT task_states
        [slot][thread], trace_counts[thread])."""
        slot, thread = self._stage_shade(tasks, depth)
        self.launch("geometry_process_tasks")
        dl = self.direct_light()[slot, thread]
        rs = self.results()[slot, thread]
        bounce = self.task_states()[PRESORT]
        return dl, rs, bounce, self.download("trace_counts", np.uint16)

    def time_shade(self, tasks: np.ndarray, depth: int, repeats: int = 5) -> float:
        """Average milliseconds of one geometry_process_tasks launch over `tasks` (CUDA events, after one warm-up launch)."""
        self._stage_shade(tasks, depth)
        ms = C.c_float(0.0)
        assert lib().refdev_time_geometry_process_tasks(repeats, C.byref(ms)) == 0
        return float(ms.value)
