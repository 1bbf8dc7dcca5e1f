"""CPU tests of the C host layer: *.obj / *.mtl loader, *.lum v4 reader, PNG writer, exported public API.

The loaders restate src/luminary/host/wavefront.c and host/lum_v4.c of the reference (file:line in the C sources);
the reference ships no tests for them, so the expectations below are derived from its code paths."""
import ctypes as C
import math
import os

import numpy as np
import pytest

import host_c
from luminary_b200 import scenes


def test_library_exports_every_declared_api_function():
    L = host_c.lib()
    names = host_c.declared_api_functions()
    assert len(names) >= 50
    missing = [n for n in names if not hasattr(L, n)]
    assert not missing, missing
    assert L.luminary_result_to_string(C.c_uint64(7)).decode() == "API exception"
    assert L.luminary_result_to_string(C.c_uint64(7 | (1 << 63))).decode() == "API exception"


def test_obj_roundtrip_matches_scene_arrays(tmp_path):
    sc = scenes.example_with_light(width=64, height=36, sphere_subdiv=2)
    obj = str(tmp_path / "scene.obj")
    scenes.write_obj(sc, obj)
    code, has, v, n, uv, mid, mats, ids = host_c.wavefront_load(obj, material_offset=3, emission_scale=1.0, bidirectional=True)
    assert code == 0 and has
    ev = np.concatenate([m.vertex for m in sc.meshes]).astype(np.float32)
    en = np.concatenate([m.normal for m in sc.meshes]).astype(np.float32)
    euv = np.concatenate([m.uv for m in sc.meshes]).astype(np.float32)
    em = np.concatenate([m.material for m in sc.meshes]).astype(np.int64)
    assert v.shape == ev.shape
    assert np.array_equal(v, ev)  # %.9g round-trips float32 exactly
    assert np.array_equal(uv, euv)
    assert np.allclose(n, en, atol=2e-7)  # the loader re-normalises (wavefront.c:953-971)
    # material 0 of a file is its default material, newmtl k becomes 1 + k; ids are offset by the caller's material count
    assert np.array_equal(mid.astype(np.int64), em + 1 + 3)
    assert ids == list(range(3, 3 + 1 + len(sc.materials)))
    assert mats[0]["albedo"] == pytest.approx((0.9, 0.9, 0.9, 1.0)) and mats[0]["roughness"] == pytest.approx(0.7)
    for k, src in enumerate(sc.materials):
        m = mats[1 + k]
        assert m["albedo"] == pytest.approx(src["albedo"], abs=1e-6)
        assert m["roughness"] == pytest.approx(src["roughness"], abs=1e-6)  # 1 - Ns / 1000
        assert m["metallic"] == src["metallic"]
        assert m["emission_active"] == src["emission_active"]
        assert m["bidirectional_emission"] is True and m["roughness_clamp"] == pytest.approx(0.25)
        if src["emission_active"]:
            assert m["emission"] == pytest.approx(src["emission"], abs=1e-5)


def test_obj_statements_quads_negative_indices_and_fallback_normals(tmp_path):
    obj = tmp_path / "q.obj"
    (tmp_path / "q.mtl").write_text("newmtl red\nKd 1 0 0\nKe 2 3 4\nNs 250\nNi 1.5\nd 0.5\nKs 0.9 0.9 0.9\n")
    obj.write_text(
        "mtllib q.mtl\no thing\n"
        "v 0 0 0\nv 1 0 0\nv 1 1 0\nv 0 1 0\nv 5 5 5\n"
        "vt 0 0\nvt 1 0\nvt 1 1\nvt 0 1\n"
        "usemtl red\n"
        "f 1/1 2/2 3/3 4/4\n"        # quad with uv, no normals -> two triangles (0,1,2) (0,2,3), face normal
        "usemtl unknown\n"
        "f -5 -4 -3\n"               # negative indices against the final vertex count -> vertices 1,2,3
        "f 5 5 5\n"                   # degenerate: dropped
        "f 1 2 3 4 5\n"               # polygon: unsupported, skipped
    )
    code, has, v, n, uv, mid, mats, ids = host_c.wavefront_load(str(obj), emission_scale=2.0)
    assert code == 0 and has
    assert v.shape[0] == 3
    assert np.array_equal(v[0], [[0, 0, 0], [1, 0, 0], [1, 1, 0]])
    assert np.array_equal(v[1], [[0, 0, 0], [1, 1, 0], [0, 1, 0]])
    assert np.array_equal(v[2], [[0, 0, 0], [1, 0, 0], [1, 1, 0]])
    assert np.array_equal(uv[1], [[0, 0], [1, 1], [0, 1]])
    assert np.allclose(n, [0, 0, 1])
    assert list(mid) == [1, 1, 0]
    red = mats[1]
    assert red["albedo"] == pytest.approx((1, 0, 0, 0.5))
    assert red["emission"] == pytest.approx((4, 6, 8)) and red["emission_scale"] == pytest.approx(2.0) and red["emission_active"]
    assert red["roughness"] == pytest.approx(0.75) and red["refraction_index"] == pytest.approx(1.5) and red["metallic"]


def test_obj_without_object_statement_yields_no_mesh(tmp_path):
    obj = tmp_path / "noobj.obj"
    obj.write_text("v 0 0 0\nv 1 0 0\nv 0 1 0\nf 1 2 3\n")
    code, has, *_ = host_c.wavefront_load(str(obj))
    assert code == 0 and not has  # wavefront.c:845-848: only a warning
    code, has, *_ = host_c.wavefront_load(str(tmp_path / "missing.obj"))
    assert (code & 0xFF) == 7 and not has


def test_lum_v4_reader(tmp_path):
    lum = tmp_path / "s.lum"
    lum.write_text(
        "Luminary\nVERSION 4\n# comment\n"
        "GENERAL WIDTH___ 640\nGENERAL HEIGHT__ 360\nGENERAL BOUNCES_ 7\nGENERAL MESHFILE a.obj\nGENERAL MESHFILE sub/b.obj\n"
        "GENERAL SAMPLES_ 128\n"
        "MATERIAL EMISSION 2.5\nMATERIAL INTERTRO 1\n"
        "CAMERA POSITION 1.5 2.5 -3.5\nCAMERA ROTATION 0.1 0.2 0.3\nCAMERA FOV_____ 0.8\nCAMERA FOCALLEN 4.0\nCAMERA APERTURE 0.05\n"
        "CAMERA EXPOSURE 2.0\nCAMERA TONEMAP_ 1\nCAMERA DITHER__ 0\nCAMERA RUSSIANR 0.25\nCAMERA BLOOM___ 0\n"
        "SKY MODE____ 2\nSKY COLORCON 0.25 0.5 0.75\nSKY AZIMUTH_ 1.0\n"
        "CLOUD ACTIVE__ 1\nFOG ACTIVE__ 1\nOCEAN ACTIVE__ 1\nTOY ACTIVE__ 1\n"
    )
    r = host_c.lum_read(str(lum))
    assert r["code"] == 0
    assert (r["settings"].width, r["settings"].height, r["settings"].max_ray_depth) == (640, 360, 7)
    assert r["mesh_files"] == ["a.obj", "sub/b.obj"]
    cam = r["camera"]
    assert (cam.pos.x, cam.pos.y, cam.pos.z) == pytest.approx((1.5, 2.5, -3.5))
    assert cam.thin_lens.fov == pytest.approx(0.8) and cam.object_distance == pytest.approx(4.0) and cam.thin_lens.aperture_size == pytest.approx(0.05)
    assert cam.exposure == pytest.approx(math.log(2.0))  # legacy linear -> exponential scale
    assert cam.tonemap == 1 and cam.dithering is False and cam.russian_roulette_threshold == pytest.approx(0.25)
    assert cam.bloom_blend == 0.0  # BLOOM___ 0 forces the blend to 0
    assert r["sky"].mode == 2 and (r["sky"].constant_color.r, r["sky"].constant_color.g, r["sky"].constant_color.b) == pytest.approx((0.25, 0.5, 0.75))
    assert r["args"].emission_scale == pytest.approx(2.5) and r["args"].legacy_smoothness and r["args"].force_bidirectional_emission
    # defaults survive for keys the file does not set (camera.c:7-66)
    assert cam.aperture_blade_count == 7 and cam.camera_scale == pytest.approx(1.0)


def test_lum_reader_rejects_bad_headers(tmp_path):
    bad = tmp_path / "bad.lum"
    bad.write_text("NotLuminary\nVERSION 4\n")
    assert (host_c.lum_read(str(bad))["code"] & 0xFF) == 7
    old = tmp_path / "old.lum"
    old.write_text("Luminary\nVERSION 3\n")
    assert (host_c.lum_read(str(old))["code"] & 0xFF) == 7
    v5 = tmp_path / "v5.lum"
    v5.write_text("Luminary\nVERSION 5\n")
    assert (host_c.lum_read(str(v5))["code"] & 0xFF) == 2
    assert (host_c.lum_read(str(tmp_path / "missing.lum"))["code"] & 0xFF) == 7


def test_png_writer_roundtrip(tmp_path):
    rng = np.random.default_rng(5)
    for (w, h, ld) in ((5, 3, 5), (64, 40, 70), (300, 260, 300)):  # the last one spans several stored deflate blocks
        img = rng.integers(0, 256, size=(h, ld, 4), dtype=np.uint8)  # b, g, r, a
        path = str(tmp_path / f"t{w}.png")
        code = host_c.lib().lum_png_write_argb8(path.encode(), img.ctypes.data_as(C.POINTER(C.c_uint8)), C.c_uint32(w), C.c_uint32(h), C.c_size_t(ld))
        assert code == 0
        rgba = host_c.png_decode_rgba(path)
        assert rgba.shape == (h, w, 4)
        assert np.array_equal(rgba[..., 0], img[:, :w, 2]) and np.array_equal(rgba[..., 1], img[:, :w, 1])
        assert np.array_equal(rgba[..., 2], img[:, :w, 0]) and np.array_equal(rgba[..., 3], img[:, :w, 3])


# ---------------------------------------------------------------------------------------------
# textures: PNG reader (reference host/png.c:415-712) and map_ statements (host/wavefront.c:159-283)
# ---------------------------------------------------------------------------------------------
def _write_png(path, img, gamma=None, filter_type=None, level=6, idat_split=1):
    """Independent PNG encoder (zlib from the Python standard library). img: (H, W) or (H, W, C) uint8 / uint16."""
    import struct
    import zlib

    a = np.asarray(img)
    if a.ndim == 2:
        a = a[:, :, None]
    h, w, c = a.shape
    ctype = {1: 0, 2: 4, 3: 2, 4: 6}[c]
    depth = 8 if a.dtype == np.uint8 else 16
    raw_rows = a.astype(">u2").tobytes() if depth == 16 else a.tobytes()
    stride = w * c * depth // 8
    bpp = c * depth // 8
    rows = [bytearray(raw_rows[y * stride:(y + 1) * stride]) for y in range(h)]
    out = bytearray()
    for y in range(h):
        ft = (y % 5) if filter_type is None else filter_type
        cur, prev = rows[y], (rows[y - 1] if y else bytearray(stride))
        line = bytearray(stride)
        for i in range(stride):
            A = cur[i - bpp] if i >= bpp else 0
            B = prev[i]
            Cc = prev[i - bpp] if i >= bpp else 0
            if ft == 0:
                pred = 0
            elif ft == 1:
                pred = A
            elif ft == 2:
                pred = B
            elif ft == 3:
                pred = (A + B) >> 1
            else:
                p = A + B - Cc
                pa, pb, pc = abs(p - A), abs(p - B), abs(p - Cc)
                pred = A if (pa <= pb and pa <= pc) else (B if pb <= pc else Cc)
            line[i] = (cur[i] - pred) & 0xFF
        out.append(ft)
        out += line
    z = zlib.compress(bytes(out), level)

    def chunk(t, d):
        return struct.pack(">I", len(d)) + t + d + struct.pack(">I", zlib.crc32(t + d) & 0xFFFFFFFF)

    data = b"\x89PNG\r\n\x1a\n" + chunk(b"IHDR", struct.pack(">IIBBBBB", w, h, depth, ctype, 0, 0, 0))
    if gamma is not None:
        data += chunk(b"gAMA", struct.pack(">I", int(round(100000.0 / gamma))))
    step = max(1, len(z) // idat_split)
    for o in range(0, len(z), step):
        data += chunk(b"IDAT", z[o:o + step])
    data += chunk(b"IEND", b"")
    open(path, "wb").write(data)


@pytest.mark.parametrize("channels", [1, 2, 3, 4])
@pytest.mark.parametrize("dtype", [np.uint8, np.uint16])
def test_png_reader_decodes_every_supported_format(tmp_path, channels, dtype):
    rng = np.random.default_rng(channels * 10 + (dtype == np.uint16))
    h, w = 37, 53
    # smooth + noise so that the deflate stream uses dynamic Huffman blocks with real matches
    base = (np.add.outer(np.arange(h), np.arange(w))[:, :, None] * 3 + rng.integers(0, 9, (h, w, channels))) % 256
    img = base.astype(dtype) if dtype == np.uint8 else (base.astype(np.uint16) * 257 + rng.integers(0, 200, (h, w, channels)).astype(np.uint16))
    path = str(tmp_path / "t.png")
    _write_png(path, img if channels > 1 else img[:, :, 0], gamma=2.2, idat_split=3)
    code, tex = host_c.png_read(path)
    assert code == 0
    got = tex["data"]
    assert got.shape == (h, w, 4) and got.dtype == dtype  # always expanded to four components (png.c:613-705)
    full = np.iinfo(dtype).max
    if channels <= 2:
        assert np.array_equal(got[..., 0], img[..., 0]) and np.array_equal(got[..., 1], img[..., 0]) and np.array_equal(got[..., 2], img[..., 0])
        assert np.array_equal(got[..., 3], img[..., 1]) if channels == 2 else np.all(got[..., 3] == full)
    else:
        assert np.array_equal(got[..., :3], img[..., :3])
        assert np.array_equal(got[..., 3], img[..., 3]) if channels == 4 else np.all(got[..., 3] == full)
    assert abs(tex["gamma"] - 100000.0 / round(100000.0 / 2.2)) < 1e-6  # gAMA chunk, png.c:541


def test_png_reader_stored_fixed_and_dynamic_blocks_and_errors(tmp_path):
    rng = np.random.default_rng(5)
    img = rng.integers(0, 256, (16, 16, 4)).astype(np.uint8)
    for level in (0, 1, 9):  # stored blocks, fast (often fixed Huffman), best (dynamic Huffman)
        path = str(tmp_path / f"l{level}.png")
        _write_png(path, img, filter_type=0, level=level)
        code, tex = host_c.png_read(path)
        assert code == 0 and np.array_equal(tex["data"], img) and tex["gamma"] == 1.0
    tiny = np.zeros((2, 2, 4), np.uint8)  # a run of zeros: fixed-Huffman block with a long match
    _write_png(str(tmp_path / "z.png"), tiny, filter_type=0)
    assert np.array_equal(host_c.png_read(str(tmp_path / "z.png"))[1]["data"], tiny)
    # the writer of this library round-trips through the reader
    argb = rng.integers(0, 256, (5, 7, 4)).astype(np.uint8)
    p = str(tmp_path / "w.png")
    assert host_c.lib().lum_png_write_argb8(p.encode(), argb.ctypes.data_as(C.POINTER(C.c_uint8)), C.c_uint32(7), C.c_uint32(5), C.c_size_t(7)) == 0
    code, tex = host_c.png_read(p)
    assert code == 0 and np.array_equal(tex["data"], argb[..., [2, 1, 0, 3]])
    # errors: not a PNG, corrupted IHDR crc, interlaced, palette, truncated stream
    bad = tmp_path / "bad.png"
    bad.write_bytes(b"not a png at all, just some bytes that are long enough to pass the size check")
    assert host_c.png_read(str(bad))[0] != 0
    good = bytearray(open(str(tmp_path / "l9.png"), "rb").read())
    corrupt = bytearray(good)
    corrupt[20] ^= 0xFF
    bad.write_bytes(bytes(corrupt))
    assert host_c.png_read(str(bad))[0] != 0
    trunc = good[:len(good) - 40]
    bad.write_bytes(bytes(trunc))
    assert host_c.png_read(str(bad))[0] != 0
    assert host_c.png_read(str(tmp_path / "missing.png"))[0] != 0


def test_mtl_texture_maps(tmp_path):
    rng = np.random.default_rng(3)
    _write_png(str(tmp_path / "albedo.png"), rng.integers(0, 256, (8, 8, 4)).astype(np.uint8), gamma=2.2)
    _write_png(str(tmp_path / "rough.png"), rng.integers(0, 256, (4, 4)).astype(np.uint8))
    (tmp_path / "sub").mkdir()
    _write_png(str(tmp_path / "sub" / "normal.png"), rng.integers(0, 65536, (4, 4, 3)).astype(np.uint16))
    (tmp_path / "m.mtl").write_text(
        "newmtl a\nKd 0.5 0.5 0.5\nmap_Kd albedo.png\nmap_Ns rough.png\nmap_Bump -bm 0.5 sub/normal.png\nmap_Ka ignored.png\n"
        "newmtl b\nmap_Kd albedo.png\nmap_Ke missing.png\nmap_refl -o 1 2 3 -s 1 1 1 rough.png\n")
    (tmp_path / "t.obj").write_text("mtllib m.mtl\no tri\nv 0 0 0\nv 1 0 0\nv 0 1 0\nvt 0 0\nvt 1 0\nvt 0 1\nusemtl a\nf 1/1 2/2 3/3\nusemtl b\nf 1/1 3/3 2/2\n")
    code, has, v, n, uv, mid, mats, ids = host_c.wavefront_load(str(tmp_path / "t.obj"), material_offset=2, texture_offset=10)
    assert code == 0 and has
    tex = host_c.last_textures
    # one texture per distinct path, in order of first use; the missing file stays as an INVALID texture with its id
    assert len(tex) == 4
    assert tex[0]["data"].shape == (8, 8, 4) and abs(tex[0]["gamma"] - 100000.0 / round(100000.0 / 2.2)) < 1e-6
    assert tex[1]["data"].shape == (4, 4, 4) and tex[2]["data"].dtype == np.uint16 and tex[3]["data"] is None
    a, b = mats[1], mats[2]
    assert (a["albedo_tex"], a["roughness_tex"], a["normal_tex"], a["luminance_tex"], a["metallic_tex"]) == (10, 11, 12, 0xFFFF, 0xFFFF)
    assert (b["albedo_tex"], b["luminance_tex"], b["metallic_tex"], b["roughness_tex"], b["normal_tex"]) == (10, 13, 11, 0xFFFF, 0xFFFF)
    assert b["emission_active"] and not a["emission_active"]  # a luminance map activates emission (wavefront.c:810)
    assert mats[0]["albedo_tex"] == 0xFFFF  # the default material of the file
