"""Per-vertex parity of the PRODUCT's shading and shadow stages (material sort -> k_shade -> k_trace_shadow, reached through the
C-ABI hook lumb200_device_shade_vertices) against

  * the oracle (oracle/orc_shade.c: orc_shade_vertices + orc_nee_segments) on identical path vertices and random numbers, and
  * the REFERENCE's own `geometry_process_tasks` kernel (cuda/geometry.cuh:11-180), compiled unmodified for sm_100a into
    oracle/_ref/librefdev.so and launched on the same vertices,

at wavefront iterations 0..3 of a room that holds every material class of the path; and per-ray parity of k_trace_shadow
(lumb200_device_trace_shadow_rays) against the oracle's restatement of the shadow any-hit programs (optix_anyhit.cuh:49-139):
opaque blockers, plain and coloured transparency, alpha-textured cut-outs, the ignore / target primitive rules, tmax.

Tolerances (the device runs --use_fast_math like the reference, the oracle IEEE + libm): discrete decisions (which light, Russian
roulette, lobe, transparency pass) must agree on all but the stated small fraction of vertices - a random number within rounding
distance of a threshold flips them; continuous outputs are compared where the decisions agree."""
import numpy as np
import pytest

import orc
import refdev
from luminary_b200 import api, scenes
from test_ref_device_gpu import W, H, _ray_unpack, _record_unpack, _rel, parity_scene

pytestmark = pytest.mark.gpu

SAMPLE_ID = 3


def product_vertices(vin: np.ndarray) -> np.ndarray:
    """oracle VERTEX_IN records -> Lumb200VertexIn"""
    v = np.zeros(vin.size, api.VERTEX_IN)
    v["pixel_x"] = vin["path_id"][:, 0] & 0x3FFF
    v["pixel_y"] = vin["path_id"][:, 1] & 0x3FFF
    v["state"] = vin["state"]
    v["origin"] = vin["origin"]
    v["ray"] = vin["ray"]
    v["prim"] = vin["prim"]
    v["t"] = vin["t"]
    v["record"] = vin["record"]
    v["medium"] = vin["medium_ior"]
    return v


@pytest.fixture(scope="module")
def setup():
    sc = parity_scene()
    lt = api.build_light_tree(sc)
    dev = api.Device(0)
    dev.build_bsdf_lut()
    luts = dev.get_bsdf_lut()
    dev.load_scene(sc, light_tree=lt)
    osc = orc.OracleScene(sc)
    osc.set_light_tree(*lt)
    osc.set_bsdf_luts(*luts)
    yield sc, dev, osc, lt, luts
    dev.destroy()


_REF = {}


def _ref_device(sc):
    """one reference device for the module (librefdev.so keeps its state in globals)"""
    if "dev" not in _REF:
        _REF["dev"] = refdev.RefDevice(sc)
        _REF["dev"].build_bsdf_lut()
    return _REF["dev"]


def _depth_of(sc, iteration):
    return iteration if not (iteration == sc.max_ray_depth and iteration > 0) else iteration - 1


@pytest.mark.parametrize("iteration", [0, 1, 2, 3, 4])
def test_k_shade_and_k_trace_shadow_vs_oracle_per_vertex(setup, iteration):
    sc, dev, osc, lt, _ = setup
    vin, _pix = osc.path_vertices(SAMPLE_ID, iteration)
    n = vin.size
    assert n > 1000
    depth = _depth_of(sc, iteration)
    is_last = iteration == sc.max_ray_depth
    want = osc.shade_vertices(vin, depth)
    seg = osc.nee_segments(vin, depth)
    got = dev.shade_vertices(product_vertices(vin), SAMPLE_ID, depth, is_last)
    assert dev.stats()["stack_overflows"] == 0

    st = {}
    names = ("light-tree", "bsdf-light", "ambient")
    for s in range(3):
        gv, wv = got["nee"][:, s]["valid"] != 0, seg[:, s]["valid"] != 0
        st[f"{names[s]}: segment present equal"] = (gv == wv).mean()
        st[f"{names[s]}: present"] = wv.mean()
        both = gv & wv
        if not both.any():
            continue
        g, w = got["nee"][both, s], seg[both, s]
        same_target = g["target_prim"] == w["target_prim"]
        st[f"{names[s]}: target light equal"] = same_target.mean()
        ray_ok = same_target & (np.abs(g["ray"] - w["ray"]).max(axis=1) < 1e-3)
        st[f"{names[s]}: ray equal"] = ray_ok.mean()
        if s != 2:
            st[f"{names[s]}: dist p99 rel"] = np.percentile(_rel(g["dist"][ray_ok], w["dist"][ray_ok], 1e-3), 99)
        st[f"{names[s]}: color p99 rel"] = np.percentile(_rel(g["color"][ray_ok], w["color"][ray_ok], 1e-4).max(axis=1), 99)
        st[f"{names[s]}: color sum ratio"] = g["color"][ray_ok].sum() / max(w["color"][ray_ok].sum(), 1e-20)
        # shadow stage: visible = color x transmittance; the transmittance is bit-exact geometry + a table lookup
        vis_w = w["color"] * w["visibility"]
        blocked_equal = ((g["visible"] != 0).any(axis=1) == (vis_w != 0).any(axis=1))[ray_ok]
        st[f"{names[s]}: occlusion decision equal"] = blocked_equal.mean()
        sel = ray_ok & (g["visible"] != 0).any(axis=1) & (vis_w != 0).any(axis=1)
        if sel.any():
            st[f"{names[s]}: visible p99 rel"] = np.percentile(_rel(g["visible"][sel], vis_w[sel], 1e-4).max(axis=1), 99)
    # slots without a segment must not have gathered anything
    for s in range(3):
        none = got["nee"][:, s]["valid"] == 0
        assert not got["nee"][none, s]["visible"].any()

    st["emission max abs"] = np.abs(got["emission"] - want["emission"]).max()
    alive_w = (want["bounce_alive"] != 0) & (not is_last)
    alive_g = got["alive"] != 0
    st["rr decision equal"] = (alive_g == alive_w).mean()
    m = alive_g & alive_w
    if m.any():
        g, w = got[m], want[m]
        st["bounce state equal"] = (g["state"] == w["bounce_state"]).mean()
        st["bounce medium equal"] = (g["medium"] == w["bounce_medium_ior"]).mean()
        st["bounce origin max abs"] = np.abs(g["origin"] - w["bounce_origin"]).max()
        ray_ok = np.abs(g["ray"] - w["bounce_ray"]).max(axis=1) < 2e-3
        st["bounce ray equal"] = ray_ok.mean()
        rec = _rel(_record_unpack(g["record"]), _record_unpack(w["bounce_record"]), 1e-6).max(axis=1)[ray_ok]
        st["bounce record p99 rel"] = np.percentile(rec, 99)
    for k, v in st.items():
        print(f"  iter {iteration}: {k:45s} {v:.6g}")

    for s in range(3):
        assert st[f"{names[s]}: segment present equal"] >= 0.995
        if f"{names[s]}: ray equal" in st:
            assert st[f"{names[s]}: target light equal"] >= 0.99
            assert st[f"{names[s]}: ray equal"] >= 0.99
            assert st[f"{names[s]}: color p99 rel"] <= 2e-2
            assert abs(st[f"{names[s]}: color sum ratio"] - 1.0) <= 2e-3
            assert st[f"{names[s]}: occlusion decision equal"] >= 0.999
            if f"{names[s]}: visible p99 rel" in st:
                assert st[f"{names[s]}: visible p99 rel"] <= 2e-2
    assert st["light-tree: dist p99 rel"] <= 1e-3
    assert st["emission max abs"] <= 1e-5
    assert st["rr decision equal"] >= 0.995
    if m.any():
        assert st["bounce state equal"] >= 0.995 and st["bounce medium equal"] >= 0.995
        assert st["bounce origin max abs"] <= 1e-4
        assert st["bounce ray equal"] >= 0.99
        assert st["bounce record p99 rel"] <= 1e-3
    else:
        assert is_last


@pytest.mark.skipif(not refdev.available(), reason="oracle/_ref/librefdev.so not built (needs /root/reference)")
@pytest.mark.parametrize("iteration", [0, 1, 2, 3])
def test_k_shade_vs_reference_geometry_process_tasks(setup, iteration):
    """k_shade against the reference's own kernel on the same tasks: NEE through the light tree (light, ray, distance, colour),
    the BSDF-sampled direction, the ambient task, emission and the bounce task. The light tree, LUTs and packed scene of the
    reference device come from the reference's own host code (tests/refdev.py)."""
    sc, dev, osc, lt, luts = setup
    ref = _ref_device(sc)
    # identical tables on both sides: the product's light tree builder is byte-identical to the reference's (test_ref_host.py) and
    # the product's LUT kernels are bit-identical to the reference's (test_ref_device_gpu.py)
    assert bytes(ref.light_tree[0]) == bytes(lt[0]) and bytes(ref.light_tree[1]) == bytes(lt[1])
    handles = osc.prim_handles()
    light_prim = {}
    for lid, (inst, tri) in enumerate(np.asarray(lt[2]).reshape(-1, 2)):
        light_prim[lid] = int(np.nonzero((handles[:, 0] == inst) & (handles[:, 1] == tri))[0][0])
    vin, _ = osc.path_vertices(SAMPLE_ID, iteration)
    n = vin.size
    depth = _depth_of(sc, iteration)
    T = 8 * refdev.THREADS_PER_BLOCK
    ref.configure(T // refdev.THREADS_PER_BLOCK, -(-n // T))
    dl, res, bounce, trace_counts = ref.shade(refdev.tasks_from_vertices(vin, handles), depth)
    got = dev.shade_vertices(product_vertices(vin), SAMPLE_ID, depth, False)
    rec_in = _record_unpack(vin["record"])

    st = {}
    # light-tree NEE: the reference writes the task even when its colour is zero; the product queues a segment when colour x throughput != 0
    ref_valid = (dl["geo_light_id"] != 0xFFFFFFFF) & ((dl["geo_color"] * rec_in) != 0).any(axis=1)
    g0 = got["nee"][:, 0]
    st["geo segment present equal"] = ((g0["valid"] != 0) == ref_valid).mean()
    both = (g0["valid"] != 0) & ref_valid
    ref_prim = np.array([light_prim.get(int(l), -1) for l in dl["geo_light_id"][both]], np.int64)
    same = g0["target_prim"][both].astype(np.int64) == ref_prim
    st["geo light equal"] = same.mean()
    gb, rb, rin = g0[both][same], dl[both][same], rec_in[both][same]
    st["geo ray max abs"] = np.abs(gb["ray"] - rb["geo_ray"]).max()
    st["geo dist p99 rel"] = np.percentile(_rel(gb["dist"], rb["geo_dist"], 1e-3), 99)
    st["geo color p99 rel"] = np.percentile(_rel(gb["color"], rb["geo_color"] * rin, 1e-4).max(axis=1), 99)
    st["geo color sum ratio"] = gb["color"].sum() / max((rb["geo_color"] * rin).sum(), 1e-20)
    # ambient task: packed record x throughput, direction through the same 2 x 32 bit octahedral packing
    g2 = got["nee"][:, 2]
    amb_col = _record_unpack(dl["amb_color"]) * rec_in
    ref_amb = (dl["amb_color"] != 0).any(axis=1) & (amb_col != 0).any(axis=1)
    st["ambient segment present equal"] = ((g2["valid"] != 0) == ref_amb).mean()
    both = (g2["valid"] != 0) & ref_amb
    ray_ok = np.abs(g2["ray"][both] - _ray_unpack(dl["amb_ray"][both])).max(axis=1) < 1e-3
    st["ambient ray equal"] = ray_ok.mean()
    st["ambient color p99 rel"] = np.percentile(_rel(g2["color"][both][ray_ok], amb_col[both][ray_ok], 1e-4).max(axis=1), 99)
    # BSDF-sampled light: the product only keeps the direction when the enumeration found an emitter; compare where it did
    g1 = got["nee"][:, 1]
    both = (g1["valid"] != 0) & (dl["bsdf_prob"] != 0)
    st["bsdf segments compared"] = float(both.sum())
    if both.any():
        st["bsdf ray equal"] = (np.abs(g1["ray"][both] - dl["bsdf_ray"][both]).max(axis=1) < 1e-3).mean()
    assert not (g1["valid"][dl["bsdf_prob"] == 0]).any()  # no direction sampled by the reference -> no segment here
    # emission
    st["emission max abs"] = np.abs(got["emission"] - res["color"]).max()
    # bounce tasks, matched through the pixel
    K = ref.tasks_per_thread
    slot, _thread = np.meshgrid(np.arange(K), np.arange(T), indexing="ij")
    b = bounce[slot < trace_counts[None, :]]
    key_in = vin["path_id"][:, 0].astype(np.int64) + vin["path_id"][:, 1].astype(np.int64) * 65536
    key_out = b["path_id"][:, 0].astype(np.int64) + b["path_id"][:, 1].astype(np.int64) * 65536
    order = np.argsort(key_in)
    src = order[np.searchsorted(key_in[order], key_out)]
    alive_ref = np.zeros(n, bool)
    alive_ref[src] = True
    st["rr decision equal"] = ((got["alive"] != 0) == alive_ref).mean()
    m = (got["alive"][src] != 0)
    g, r = got[src[m]], b[m]
    st["bounce state equal"] = (g["state"] == r["state"]).mean()
    st["bounce medium equal"] = (g["medium"] == r["ior"]).mean()
    st["bounce origin max abs"] = np.abs(g["origin"] - r["origin"]).max()
    ray_ok = np.abs(g["ray"] - r["ray"]).max(axis=1) < 2e-3
    st["bounce ray equal"] = ray_ok.mean()
    st["bounce record p99 rel"] = np.percentile(_rel(_record_unpack(g["record"]), _record_unpack(r["record"]), 1e-6).max(axis=1)[ray_ok], 99)
    for k, v in st.items():
        print(f"  iter {iteration}: {k:45s} {v:.6g}")

    assert st["geo segment present equal"] >= 0.995 and st["geo light equal"] >= 0.99
    assert st["geo ray max abs"] <= 1e-3 and st["geo dist p99 rel"] <= 1e-3 and st["geo color p99 rel"] <= 2e-2
    assert abs(st["geo color sum ratio"] - 1.0) <= 2e-3
    assert st["ambient segment present equal"] >= 0.995 and st["ambient ray equal"] >= 0.995 and st["ambient color p99 rel"] <= 1e-3
    if "bsdf ray equal" in st:
        assert st["bsdf ray equal"] >= 0.99
    assert st["emission max abs"] <= 1e-5
    assert st["rr decision equal"] >= 0.995
    assert st["bounce state equal"] >= 0.995 and st["bounce medium equal"] >= 0.995
    assert st["bounce origin max abs"] <= 1e-4 and st["bounce ray equal"] >= 0.99 and st["bounce record p99 rel"] <= 1e-3


# ---------------------------------------------------------------------------------------------------------------------
# per-ray shadow transmittance
# ---------------------------------------------------------------------------------------------------------------------
def _shadow_scene(textured: bool):
    """A stack of horizontal sheets between y = 0 and y = 3 (each two triangles, 4 x 4 m), rays travel up and down through them:
    material 0 opaque, 1 plain 40 % transparent, 2 coloured transparent, 3 fully transparent (alpha 0, not coloured: ignored),
    4 coloured with alpha 0 (tints), 5 (textured variant only) alpha cut-out checkerboard texture."""
    mats = [
        scenes.default_material(albedo=(0.8, 0.8, 0.8, 1.0)),
        scenes.default_material(albedo=(0.9, 0.5, 0.2, 0.6)),
        scenes.default_material(albedo=(0.9, 0.5, 0.2, 0.3), colored_transparency=True),
        scenes.default_material(albedo=(0.3, 0.3, 0.3, 0.0)),
        scenes.default_material(albedo=(0.2, 0.7, 0.9, 0.0), colored_transparency=True),
    ]
    sheets = [(0.5, 1), (1.0, 2), (1.5, 3), (2.0, 4), (2.5, 1), (3.0, 0)]
    textures = []
    if textured:
        tex = np.zeros((8, 8, 4), np.uint8)
        tex[..., :3] = (200, 120, 40)
        tex[..., 3] = np.where((np.add.outer(np.arange(8), np.arange(8)) & 1) == 0, 255, 0)  # alpha checkerboard: opaque / cut out
        tex[0, 0, 3] = 128                                                                   # one half-transparent texel
        textures.append(dict(data=tex, wrap_u=0, wrap_v=0, filter=0, gamma=1.0))
        m = scenes.default_material(albedo=(1.0, 1.0, 1.0, 1.0))
        m["albedo_tex"] = 0
        mats.append(m)
        sheets.insert(2, (1.25, 5))
    meshes = [scenes.quad_uv((-2, y, -2), (2, y, -2), (2, y, 2), (-2, y, 2), mid, uv_scale=3.0) for y, mid in sheets]
    return scenes.Scene(name="shadow_sheets", width=256, height=128, max_ray_depth=1, meshes=meshes,
                        instances=[scenes.Instance(i) for i in range(len(meshes))], materials=mats, camera=dict(scenes.example(64, 36, 1).camera),
                        sky_mode=2, sky_color=(1.0, 1.0, 1.0), textures=textures)


@pytest.mark.parametrize("textured", [False, True])
def test_k_trace_shadow_per_ray_vs_oracle(setup, textured):
    _, _, _, _, luts = setup
    sc = _shadow_scene(textured)
    dev = api.Device(0)
    dev.set_bsdf_lut(*luts)
    dev.load_scene(sc, light_tree=None)
    osc = orc.OracleScene(sc)
    nprim = osc.num_prims()
    rng = np.random.default_rng(17)
    n = 20000
    o = np.stack([rng.uniform(-1.9, 1.9, n), np.where(rng.random(n) < 0.5, 0.0, 3.5), rng.uniform(-1.9, 1.9, n)], axis=1).astype(np.float32)
    d = rng.normal(size=(n, 3)).astype(np.float32)
    d[:, 1] = np.where(o[:, 1] < 1.0, np.abs(d[:, 1]) + 1.5, -np.abs(d[:, 1]) - 1.5)   # towards the stack
    d /= np.linalg.norm(d, axis=1, keepdims=True)
    d = d.astype(np.float32)
    limit = np.where(rng.random(n) < 0.3, rng.uniform(0.3, 3.0, n), np.float32(3.4e38)).astype(np.float32)   # some rays end inside the stack
    ignore = np.where(rng.random(n) < 0.3, rng.integers(0, nprim, n), 0xFFFFFFFF).astype(np.uint32)
    target = np.where(rng.random(n) < 0.3, rng.integers(0, nprim, n), 0xFFFFFFFF).astype(np.uint32)
    # rays that start exactly on a sheet they must ignore (tmin = eps does not exclude t = 0 hits of neighbouring geometry)
    o[:500, 1] = 1.0
    ignore[:500] = 2 + (np.arange(500) & 1)
    want = osc.shadow_rays(o, d, limit, ignore, target)
    got = dev.trace_shadow_rays(o, d, limit, ignore, target)
    assert dev.stats()["stack_overflows"] == 0
    dev.destroy()
    blocked_w, blocked_g = ~want.any(axis=1), ~got.any(axis=1)
    print(f"  textured={textured}: {n} rays, blocked {blocked_w.mean():.3f}, fully visible {(want == 1).all(axis=1).mean():.3f}, "
          f"attenuated {((want != 1).any(axis=1) & ~blocked_w).mean():.3f}")
    assert 0.1 < blocked_w.mean() < 0.9 and ((want != 1).any(axis=1) & ~blocked_w).mean() > 0.1   # the test exercises every branch
    if not textured:
        # the hits are bit-exact (same watertight test) and the response is a per-material table built from the same packed
        # 16-bit albedo: the products agree up to the order of the multiplications along the ray
        assert np.array_equal(blocked_w, blocked_g)
        assert np.abs(got - want).max() <= 2e-6
    else:
        # alpha from a point-sampled texture at the hit's interpolated uv: a hit within rounding distance of a texel edge may
        # read the neighbouring texel (bf16 vertex uv on both sides, fast-math interpolation on the device)
        same = blocked_w == blocked_g
        assert same.mean() >= 0.999
        assert (np.abs(got - want).max(axis=1) <= 2e-6).mean() >= 0.998


# ---------------------------------------------------------------------------------------------------------------------
# many lights: the reservoir chaos is a property of the reference's algorithm, not of this implementation
# ---------------------------------------------------------------------------------------------------------------------
def _many_lights_scene():
    """parity_scene() + a 6 x 5 grid of small ceiling emitters of different power: 64 lights, 8 root sections."""
    sc = parity_scene()
    rng = np.random.default_rng(41)
    base = len(sc.materials)
    for k in range(4):
        e = float(2.0 + 6.0 * k)
        sc.materials.append(scenes.default_material(albedo=(1.0, 1.0, 1.0, 1.0), emission=(e, e * 0.9, e * 0.7), emission_active=True, roughness=1.0))
    quads = []
    for ix in range(6):
        for iz in range(5):
            x, z = -1.7 + 0.65 * ix, -3.7 + 0.8 * iz
            h = 0.06 + 0.05 * rng.random()
            quads.append(scenes.quad((x - h, 2.97, z - h), (x + h, 2.97, z - h), (x + h, 2.97, z + h), (x - h, 2.97, z + h), base + int(rng.integers(0, 4))))
    sc.meshes.append(scenes.merge(quads))
    sc.instances.append(scenes.Instance(len(sc.meshes) - 1))
    sc.name = "parity_many_lights"
    return sc


@pytest.mark.skipif(not refdev.available(), reason="oracle/_ref/librefdev.so not built (needs /root/reference)")
def test_light_choice_chaos_is_inherent_to_the_reference_algorithm(setup):
    """On a scene with 64 emitters the product's light-tree NEE picks the oracle's light on only ~9 of 10 vertices (BASELINE scenes:
    70 - 91 %, tests/test_configs_gpu.py). This test shows where that comes from: the REFERENCE's own geometry_process_tasks,
    unmodified, launched on the same vertices, disagrees with the IEEE / libm oracle about as often, and the product agrees with the
    reference no worse than the reference agrees with the oracle - three implementations of one ill-conditioned computation (8
    reservoir lanes re-using one 23-bit random number over all root children, light_tree.cuh:191-262). All three conserve the NEE
    energy to a fraction of a percent, which is what makes the images agree in the mean."""
    _, _, _, _, luts = setup
    sc = _many_lights_scene()
    lt = api.build_light_tree(sc)
    assert np.asarray(lt[2]).reshape(-1, 2).shape[0] == 64
    dev = api.Device(0)
    dev.set_bsdf_lut(*luts)
    dev.load_scene(sc, light_tree=lt)
    osc = orc.OracleScene(sc)
    osc.set_light_tree(*lt)
    osc.set_bsdf_luts(*luts)
    _REF.pop("dev", None)
    ref = refdev.RefDevice(sc)
    ref.build_bsdf_lut()
    assert bytes(ref.light_tree[0]) == bytes(lt[0]) and bytes(ref.light_tree[1]) == bytes(lt[1])
    handles = osc.prim_handles()
    light_prim = np.array([int(np.nonzero((handles[:, 0] == i) & (handles[:, 1] == t))[0][0]) for i, t in np.asarray(lt[2]).reshape(-1, 2)], np.int64)

    vin, _ = osc.path_vertices(SAMPLE_ID, 1)
    n = vin.size
    want = osc.shade_vertices(vin, 1)
    T = 8 * refdev.THREADS_PER_BLOCK
    ref.configure(T // refdev.THREADS_PER_BLOCK, -(-n // T))
    dl, _res, _bounce, _tc = ref.shade(refdev.tasks_from_vertices(vin, handles), 1)
    got = dev.shade_vertices(product_vertices(vin), SAMPLE_ID, 1, False)
    dev.destroy()
    _REF.pop("dev", None)

    rec_in = _record_unpack(vin["record"])
    lit = (rec_in != 0).any(axis=1)                                  # the product queues nothing for paths without throughput
    o_id, r_id = want["geo_light_id"].astype(np.int64), dl["geo_light_id"].astype(np.int64)
    o_prim = np.where(o_id != 0xFFFFFFFF, light_prim[np.minimum(o_id, 63)], -1)
    r_prim = np.where(r_id != 0xFFFFFFFF, light_prim[np.minimum(r_id, 63)], -1)
    p_prim = np.where(got["nee"][:, 0]["valid"] != 0, got["nee"][:, 0]["target_prim"].astype(np.int64), -1)
    sel = lit & (o_prim >= 0) & (r_prim >= 0) & (p_prim >= 0)
    ref_orc = (r_prim[sel] == o_prim[sel]).mean()
    prod_orc = (p_prim[sel] == o_prim[sel]).mean()
    prod_ref = (p_prim[sel] == r_prim[sel]).mean()
    e_o = (want["geo_color"] * rec_in)[sel].sum()
    e_r = (dl["geo_color"] * rec_in)[sel].sum()
    e_p = got["nee"][:, 0]["color"][sel].sum()
    print(f"  64 lights, {int(sel.sum())} vertices: same light reference/oracle {ref_orc:.4f}, product/oracle {prod_orc:.4f}, product/reference {prod_ref:.4f}; "
          f"NEE energy oracle {e_o:.5g} reference {e_r:.5g} product {e_p:.5g}")
    assert sel.sum() > 5000
    assert ref_orc < 0.999, "the reference agrees with the oracle: the chaos argument would not hold"
    assert prod_orc >= ref_orc - 0.03 and prod_ref >= ref_orc - 0.03
    assert abs(e_p - e_o) <= 1e-2 * e_o and abs(e_r - e_o) <= 1e-2 * e_o and abs(e_p - e_r) <= 1e-2 * e_r
