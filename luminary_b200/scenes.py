"""Deterministic procedural scenes for the BASELINE.json configurations.

All generators are pure numpy, seeded with SplitMix64, and return a `Scene` whose meshes are the non-indexed
triangle soups the reference hands to `device_add_mesh` (mesh.h:8-20): 9 floats of position + 9 of normal +
6 of uv + one uint16 material id per triangle.

  S0  example()      Example.obj stand-in: Cornell-style box + 2 icospheres, 40 972 triangles (config 1)
  S1  atrium(...)    colonnaded hall, exactly 1 000 000 triangles at full size (configs 2 and 5)
  S2  terrain(...)   ridged-fBm heightfield + emissive lantern quads (config 3)
  S3  divergence(..) S1 geometry with per-triangle hashed materials incl. translucent + open ceiling (config 4)
"""
from __future__ import annotations

import dataclasses
import math
from typing import Dict, List, Optional

import numpy as np

MASK64 = (1 << 64) - 1


class SplitMix64:
    """SplitMix64 (Steele, Lea, Flood 2014); vectorised draws for numpy."""

    def __init__(self, seed: int):
        self.state = seed & MASK64

    def next_u64(self, n: int) -> np.ndarray:
        idx = np.arange(1, n + 1, dtype=np.uint64)
        with np.errstate(over="ignore"):
            z = np.uint64(self.state) + idx * np.uint64(0x9E3779B97F4A7C15)
            self.state = int(z[-1]) if n else self.state
            z = (z ^ (z >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)
            z = (z ^ (z >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)
            z = z ^ (z >> np.uint64(31))
        return z

    def uniform(self, n: int, lo: float = 0.0, hi: float = 1.0) -> np.ndarray:
        u = (self.next_u64(n) >> np.uint64(11)).astype(np.float64) * (1.0 / (1 << 53))
        return (lo + (hi - lo) * u).astype(np.float32)


def default_material(**kw) -> Dict:
    """material_get_default (reference material.c:5-29) + wavefront.c defaults."""
    m = dict(base_substrate=0, albedo=(0.9, 0.9, 0.9, 1.0), emission=(0.0, 0.0, 0.0), emission_scale=1.0, roughness=0.7,
             roughness_clamp=0.25, refraction_index=1.0, emission_active=False, thin_walled=False, metallic=False,
             colored_transparency=False, roughness_as_smoothness=False, normal_map_is_compressed=True, bidirectional_emission=False)
    m.update(kw)
    return m


@dataclasses.dataclass
class Mesh:
    vertex: np.ndarray    # (T, 3, 3) float32
    normal: np.ndarray    # (T, 3, 3) float32
    uv: np.ndarray        # (T, 3, 2) float32
    material: np.ndarray  # (T,) uint16

    @property
    def num_tris(self) -> int:
        return int(self.vertex.shape[0])


@dataclasses.dataclass
class Instance:
    mesh_id: int
    translation: tuple = (0.0, 0.0, 0.0)
    rotation: tuple = (0.0, 0.0, 0.0)
    scale: tuple = (1.0, 1.0, 1.0)
    active: bool = True


@dataclasses.dataclass
class Scene:
    name: str
    meshes: List[Mesh]
    instances: List[Instance]
    materials: List[Dict]
    camera: Dict
    width: int
    height: int
    max_ray_depth: int
    sky_mode: int = 2
    sky_color: tuple = (1.0, 1.0, 1.0)
    # procedural atmosphere (sky_mode 0): overrides of the reference's defaults (sky.c:6-42), field names of Lumb200Sky; None = defaults
    sky: Dict = None
    # material textures: dicts of data ((H, W, C) uint8 / uint16 / float32, C in {1, 2, 4}; None = invalid), wrap_u, wrap_v, filter, gamma
    textures: List[Dict] = dataclasses.field(default_factory=list)

    @property
    def num_tris(self) -> int:
        return sum(self.meshes[i.mesh_id].num_tris for i in self.instances if i.active)


def default_camera(**kw) -> Dict:
    c = dict(pos=(0.0, 0.0, 0.0), rotation=(0.0, 0.0, 0.0), fov=1.0, aperture_size=0.0, object_distance=1.0, camera_scale=1.0,
             russian_roulette_threshold=0.1, aperture_shape=0, aperture_blade_count=7)
    c.update(kw)
    return c


# ---------------------------------------------------------------------------------------------
# mesh helpers
# ---------------------------------------------------------------------------------------------
def _face_normals(v: np.ndarray) -> np.ndarray:
    e1 = v[:, 1] - v[:, 0]
    e2 = v[:, 2] - v[:, 0]
    n = np.cross(e1, e2)
    ln = np.linalg.norm(n, axis=1, keepdims=True)
    n = np.where(ln > 0, n / np.maximum(ln, 1e-30), np.array([0.0, 1.0, 0.0]))
    return np.repeat(n[:, None, :], 3, axis=1).astype(np.float32)


def mesh_from_tris(v: np.ndarray, material, normals: Optional[np.ndarray] = None, uv: Optional[np.ndarray] = None) -> Mesh:
    v = np.ascontiguousarray(v, dtype=np.float32).reshape(-1, 3, 3)
    t = v.shape[0]
    n = _face_normals(v) if normals is None else np.ascontiguousarray(normals, dtype=np.float32).reshape(-1, 3, 3)
    if uv is None:
        uv = np.zeros((t, 3, 2), dtype=np.float32)
        uv[:, 1, 0] = 1.0
        uv[:, 2, 1] = 1.0
    mat = np.full(t, material, dtype=np.uint16) if np.isscalar(material) else np.ascontiguousarray(material, dtype=np.uint16)
    return Mesh(v, n, np.ascontiguousarray(uv, dtype=np.float32), mat)


def merge(meshes: List[Mesh]) -> Mesh:
    return Mesh(np.concatenate([m.vertex for m in meshes]), np.concatenate([m.normal for m in meshes]),
                np.concatenate([m.uv for m in meshes]), np.concatenate([m.material for m in meshes]))


def quad(p0, p1, p2, p3, material) -> Mesh:
    """Two triangles (p0,p1,p2), (p0,p2,p3); normal = (p1-p0) x (p2-p0)."""
    p = [np.asarray(x, dtype=np.float32) for x in (p0, p1, p2, p3)]
    return mesh_from_tris(np.array([[p[0], p[1], p[2]], [p[0], p[2], p[3]]]), material)


def box_inward(lo, hi, materials) -> Mesh:
    """Axis-aligned room seen from inside: 6 quads = 12 triangles, normals pointing inwards.
    materials: (floor, ceiling, wall -x, wall +x, wall -z, wall +z)."""
    x0, y0, z0 = lo
    x1, y1, z1 = hi
    return merge([
        quad((x0, y0, z0), (x0, y0, z1), (x1, y0, z1), (x1, y0, z0), materials[0]),  # floor, +y
        quad((x0, y1, z0), (x1, y1, z0), (x1, y1, z1), (x0, y1, z1), materials[1]),  # ceiling, -y
        quad((x0, y0, z0), (x0, y1, z0), (x0, y1, z1), (x0, y0, z1), materials[2]),  # -x wall, +x
        quad((x1, y0, z0), (x1, y0, z1), (x1, y1, z1), (x1, y1, z0), materials[3]),  # +x wall, -x
        quad((x0, y0, z0), (x1, y0, z0), (x1, y1, z0), (x0, y1, z0), materials[4]),  # -z wall, +z
        quad((x0, y0, z1), (x0, y1, z1), (x1, y1, z1), (x1, y0, z1), materials[5]),  # +z wall, -z
    ])


def icosphere(subdiv: int, radius: float, center, material) -> Mesh:
    """20 * 4^subdiv triangles, smooth normals."""
    t = (1.0 + math.sqrt(5.0)) / 2.0
    verts = np.array([[-1, t, 0], [1, t, 0], [-1, -t, 0], [1, -t, 0], [0, -1, t], [0, 1, t], [0, -1, -t], [0, 1, -t],
                      [t, 0, -1], [t, 0, 1], [-t, 0, -1], [-t, 0, 1]], dtype=np.float64)
    verts /= np.linalg.norm(verts, axis=1, keepdims=True)
    faces = np.array([[0, 11, 5], [0, 5, 1], [0, 1, 7], [0, 7, 10], [0, 10, 11], [1, 5, 9], [5, 11, 4], [11, 10, 2], [10, 7, 6], [7, 1, 8],
                      [3, 9, 4], [3, 4, 2], [3, 2, 6], [3, 6, 8], [3, 8, 9], [4, 9, 5], [2, 4, 11], [6, 2, 10], [8, 6, 7], [9, 8, 1]])
    tris = verts[faces]  # (20, 3, 3)
    for _ in range(subdiv):
        a, b, c = tris[:, 0], tris[:, 1], tris[:, 2]
        ab = a + b
        bc = b + c
        ca = c + a
        ab /= np.linalg.norm(ab, axis=1, keepdims=True)
        bc /= np.linalg.norm(bc, axis=1, keepdims=True)
        ca /= np.linalg.norm(ca, axis=1, keepdims=True)
        tris = np.concatenate([np.stack([a, ab, ca], 1), np.stack([b, bc, ab], 1), np.stack([c, ca, bc], 1), np.stack([ab, bc, ca], 1)])
    normals = tris.astype(np.float32)
    pos = (tris * radius + np.asarray(center, dtype=np.float64)).astype(np.float32)
    return mesh_from_tris(pos, material, normals=normals)


def column(segments: int, rings: int, radius: float, height: float, material) -> Mesh:
    """Fluted column: `segments`-gon x `rings` rings x 2 triangles, smooth normals, base at y = 0."""
    ang = np.linspace(0.0, 2.0 * np.pi, segments + 1)
    ys = np.linspace(0.0, height, rings + 1)
    # entasis + fluting
    prof = radius * (1.0 - 0.12 * (ys / height) ** 2)
    flute = 1.0 + 0.03 * np.cos(ang * (segments // 4))
    r = prof[:, None] * flute[None, :]
    x = r * np.cos(ang)[None, :]
    z = r * np.sin(ang)[None, :]
    y = np.repeat(ys[:, None], segments + 1, axis=1)
    p = np.stack([x, y, z], axis=-1)
    nrm = np.stack([np.cos(ang)[None, :].repeat(rings + 1, 0), np.zeros_like(x), np.sin(ang)[None, :].repeat(rings + 1, 0)], axis=-1)
    p00, p01 = p[:-1, :-1], p[:-1, 1:]
    p10, p11 = p[1:, :-1], p[1:, 1:]
    n00, n01 = nrm[:-1, :-1], nrm[:-1, 1:]
    n10, n11 = nrm[1:, :-1], nrm[1:, 1:]
    t1 = np.stack([p00, p10, p11], axis=2).reshape(-1, 3, 3)
    t2 = np.stack([p00, p11, p01], axis=2).reshape(-1, 3, 3)
    m1 = np.stack([n00, n10, n11], axis=2).reshape(-1, 3, 3)
    m2 = np.stack([n00, n11, n01], axis=2).reshape(-1, 3, 3)
    return mesh_from_tris(np.concatenate([t1, t2]), material, normals=np.concatenate([m1, m2]))


def _fbm(x: np.ndarray, z: np.ndarray, octaves: int, seed: int, ridged: bool = False) -> np.ndarray:
    """Value-noise fBm on a lattice hashed with SplitMix64 constants (deterministic, numpy only)."""
    def lattice(ix, iz, o):
        with np.errstate(over="ignore"):
            h = (ix.astype(np.uint64) * np.uint64(0x9E3779B97F4A7C15)) ^ (iz.astype(np.uint64) * np.uint64(0xBF58476D1CE4E5B9)) ^ np.uint64((seed + 0x632BE5AB * (o + 1)) & MASK64)
            h = (h ^ (h >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)
            h = (h ^ (h >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)
            h = h ^ (h >> np.uint64(31))
        return (h >> np.uint64(11)).astype(np.float64) * (1.0 / (1 << 53))

    total = np.zeros_like(x, dtype=np.float64)
    amp, freq = 1.0, 1.0
    norm = 0.0
    for o in range(octaves):
        fx, fz = x * freq, z * freq
        ix, iz = np.floor(fx).astype(np.int64), np.floor(fz).astype(np.int64)
        tx, tz = fx - ix, fz - iz
        sx, sz = tx * tx * (3 - 2 * tx), tz * tz * (3 - 2 * tz)
        v00, v10 = lattice(ix, iz, o), lattice(ix + 1, iz, o)
        v01, v11 = lattice(ix, iz + 1, o), lattice(ix + 1, iz + 1, o)
        v = (v00 * (1 - sx) + v10 * sx) * (1 - sz) + (v01 * (1 - sx) + v11 * sx) * sz
        if ridged:
            v = 1.0 - np.abs(2.0 * v - 1.0)
        total += amp * v
        norm += amp
        amp *= 0.5
        freq *= 2.0
    return total / norm


def heightfield(nx: int, nz: int, x0: float, z0: float, x1: float, z1: float, height_fn, material, flip: bool = False) -> Mesh:
    """(nx x nz) quads = 2 nx nz triangles, smooth normals from central differences of the grid."""
    xs = np.linspace(x0, x1, nx + 1)
    zs = np.linspace(z0, z1, nz + 1)
    X, Z = np.meshgrid(xs, zs, indexing="ij")
    Y = height_fn(X, Z)
    P = np.stack([X, Y, Z], axis=-1)
    dx = np.gradient(Y, xs, axis=0)
    dz = np.gradient(Y, zs, axis=1)
    N = np.stack([-dx, np.ones_like(Y), -dz], axis=-1)
    N /= np.linalg.norm(N, axis=-1, keepdims=True)
    if flip:
        N = -N
    p00, p10, p01, p11 = P[:-1, :-1], P[1:, :-1], P[:-1, 1:], P[1:, 1:]
    n00, n10, n01, n11 = N[:-1, :-1], N[1:, :-1], N[:-1, 1:], N[1:, 1:]
    if not flip:  # +y facing: (p00, p01, p11), (p00, p11, p10)
        t1 = np.stack([p00, p01, p11], axis=2)
        t2 = np.stack([p00, p11, p10], axis=2)
        m1 = np.stack([n00, n01, n11], axis=2)
        m2 = np.stack([n00, n11, n10], axis=2)
    else:
        t1 = np.stack([p00, p11, p01], axis=2)
        t2 = np.stack([p00, p10, p11], axis=2)
        m1 = np.stack([n00, n11, n01], axis=2)
        m2 = np.stack([n00, n10, n11], axis=2)
    v = np.concatenate([t1.reshape(-1, 3, 3), t2.reshape(-1, 3, 3)])
    n = np.concatenate([m1.reshape(-1, 3, 3), m2.reshape(-1, 3, 3)])
    return mesh_from_tris(v, material, normals=n)


# ---------------------------------------------------------------------------------------------
# S0: Example.obj stand-in (config 1)
# ---------------------------------------------------------------------------------------------
def example(width: int = 960, height: int = 540, sphere_subdiv: int = 5) -> Scene:
    """Seed 0xB200E0: 4 x 3 x 4 m Cornell-style box (12 tris) + 2 icospheres (20 * 4^5 tris each) = 40 972 triangles."""
    rng = SplitMix64(0xB200E0)
    mats = [
        default_material(albedo=(0.73, 0.73, 0.73, 1.0), roughness=1.0),                       # 0 white
        default_material(albedo=(0.65, 0.05, 0.05, 1.0), roughness=1.0),                       # 1 red
        default_material(albedo=(0.12, 0.45, 0.15, 1.0), roughness=1.0),                       # 2 green
        default_material(albedo=(0.9, 0.9, 0.9, 1.0), roughness=0.15),                          # 3 glossy sphere
        default_material(albedo=(0.95, 0.8, 0.4, 1.0), roughness=0.3, metallic=True),          # 4 metal sphere
        default_material(albedo=(1.0, 1.0, 1.0, 1.0), emission=(12.0, 12.0, 12.0), emission_active=True, roughness=1.0),  # 5 light
    ]
    room = box_inward((-2.0, 0.0, -4.0), (2.0, 3.0, 0.0), (0, 0, 1, 2, 0, 0))
    jitter = rng.uniform(6, -0.05, 0.05)
    s1 = icosphere(sphere_subdiv, 0.6, (-0.8 + float(jitter[0]), 0.6, -2.6 + float(jitter[1])), 3)
    s2 = icosphere(sphere_subdiv, 0.45, (0.85 + float(jitter[2]), 0.45, -1.9 + float(jitter[3])), 4)
    meshes = [room, s1, s2]
    instances = [Instance(0), Instance(1), Instance(2)]
    cam = default_camera(pos=(0.0, 1.5, -0.15), rotation=(0.0, 0.0, 0.0), fov=0.9)
    return Scene("example", meshes, instances, mats, cam, width, height, max_ray_depth=0)


def example_with_light(width: int = 256, height: int = 144, sphere_subdiv: int = 3, max_ray_depth: int = 3) -> Scene:
    """Small lit variant of S0 for shading parity tests: adds one emissive ceiling quad (2 lights)."""
    sc = example(width, height, sphere_subdiv)
    light = quad((-0.5, 2.99, -2.5), (0.5, 2.99, -2.5), (0.5, 2.99, -1.5), (-0.5, 2.99, -1.5), 5)  # facing -y
    sc.meshes.append(light)
    sc.instances.append(Instance(3))
    sc.max_ray_depth = max_ray_depth
    sc.sky_color = (0.0, 0.0, 0.0)
    sc.name = "example_lit"
    return sc


def quad_uv(p0, p1, p2, p3, material, uv_scale: float = 1.0) -> Mesh:
    """quad() with texture coordinates (0,0), (s,0), (s,s), (0,s) on p0..p3."""
    p = [np.asarray(x, dtype=np.float32) for x in (p0, p1, p2, p3)]
    s = float(uv_scale)
    uv = np.array([[[0, 0], [s, 0], [s, s]], [[0, 0], [s, s], [0, s]]], dtype=np.float32)
    return mesh_from_tris(np.array([[p[0], p[1], p[2]], [p[0], p[2], p[3]]]), material, uv=uv)


def textured_example(width: int = 192, height: int = 108, sphere_subdiv: int = 2, max_ray_depth: int = 3, seed: int = 0xB200F1) -> Scene:
    """a lit room whose materials exercise every texture slot of the path - albedo (+ gamma), roughness and
    normal maps on the floor, an alpha cut-out screen (alpha 0 / 0.5 / 1 texels: closest-hit any-hit + shadow transparency),
    a luminance-textured emitter, a coloured-transparency pane with a textured albedo, an INVALID texture (default values)
    and all four address modes. Deterministic (SplitMix64)."""
    rng = SplitMix64(seed)

    def rnd_u8(shape, lo=0, hi=255):
        return (rng.uniform(int(np.prod(shape)), lo, hi + 0.999).astype(np.int32).clip(0, 255).astype(np.uint8)).reshape(shape)

    # 0: floor albedo, 64x64 RGBA8 checker with per-texel noise, sRGB-ish gamma, wrap
    yy, xx = np.mgrid[0:64, 0:64]
    chk = (((xx // 8) + (yy // 8)) & 1).astype(np.float32)
    alb = np.empty((64, 64, 4), np.uint8)
    noise = rnd_u8((64, 64, 3), 0, 40).astype(np.float32)
    alb[..., 0] = np.clip(60 + 150 * chk + noise[..., 0], 0, 255)
    alb[..., 1] = np.clip(70 + 120 * (1 - chk) + noise[..., 1], 0, 255)
    alb[..., 2] = np.clip(90 + noise[..., 2], 0, 255)
    alb[..., 3] = 255
    # 1: floor roughness, 32x32 single channel u8, mirror
    rough = rnd_u8((32, 32, 1), 40, 250)
    # 2: floor normal map, 32x32 RGBA8 "compressed" ([0,1] -> [-1,1]), wrap
    nx = rng.uniform(32 * 32, -0.35, 0.35).reshape(32, 32)
    ny = rng.uniform(32 * 32, -0.35, 0.35).reshape(32, 32)
    nz = np.sqrt(np.maximum(1.0 - nx * nx - ny * ny, 0.0))
    nmap = np.empty((32, 32, 4), np.uint8)
    nmap[..., 0] = np.round((nx * 0.5 + 0.5) * 255)
    nmap[..., 1] = np.round((ny * 0.5 + 0.5) * 255)
    nmap[..., 2] = np.round((nz * 0.5 + 0.5) * 255)
    nmap[..., 3] = 255
    # 3: cut-out screen, 32x32 RGBA8: 4x4-texel blocks of alpha 0 / 128 / 255, clamp
    blocks = (rng.uniform(8 * 8, 0.0, 3.0).astype(np.int32).clip(0, 2)).reshape(8, 8)
    a = np.array([0, 128, 255], np.uint8)[np.kron(blocks, np.ones((4, 4), np.int32))]
    cut = np.empty((32, 32, 4), np.uint8)
    cut[..., :3] = rnd_u8((32, 32, 3), 80, 240)
    cut[..., 3] = a
    # 4: emitter luminance, 8x8 RGBA fp32, border
    lum = np.ones((8, 8, 4), np.float32)
    lum[..., :3] = rng.uniform(8 * 8 * 3, 0.3, 1.0).reshape(8, 8, 3)
    # 5: stained glass albedo, 16x16 RGBA16 with alpha ~ 0.35..0.65, point filter
    glass = np.empty((16, 16, 4), np.uint16)
    glass[..., :3] = (rng.uniform(16 * 16 * 3, 0.2, 1.0) * 65535).astype(np.uint16).reshape(16, 16, 3)
    glass[..., 3] = (rng.uniform(16 * 16, 0.35, 0.65) * 65535).astype(np.uint16).reshape(16, 16)
    textures = [
        dict(data=alb, wrap_u=0, wrap_v=0, filter=1, gamma=2.2),
        dict(data=rough, wrap_u=2, wrap_v=2, filter=1, gamma=1.0),
        dict(data=nmap, wrap_u=0, wrap_v=0, filter=1, gamma=1.0),
        dict(data=cut, wrap_u=1, wrap_v=1, filter=1, gamma=2.2),
        dict(data=lum, wrap_u=3, wrap_v=3, filter=1, gamma=1.0),
        dict(data=glass, wrap_u=0, wrap_v=0, filter=0, gamma=1.0),
        dict(data=None),  # 6: invalid texture -> texture_load returns its default
    ]
    mats = [
        default_material(albedo=(0.7, 0.7, 0.7, 1.0), roughness=0.6, albedo_tex=0, roughness_tex=1, normal_tex=2),          # 0 floor
        default_material(albedo=(0.73, 0.73, 0.73, 1.0), roughness=1.0),                                                    # 1 walls
        default_material(albedo=(0.5, 0.5, 0.5, 1.0), roughness=0.8, albedo_tex=3),                                         # 2 cut-out screen
        default_material(albedo=(1.0, 1.0, 1.0, 1.0), emission=(12.0, 12.0, 12.0), emission_scale=12.0, emission_active=True, roughness=1.0,
                         luminance_tex=4),                                                                                  # 3 textured light
        default_material(albedo=(0.9, 0.9, 0.9, 0.5), roughness=0.3, colored_transparency=True, albedo_tex=5),              # 4 stained glass
        default_material(albedo=(0.9, 0.9, 0.9, 1.0), roughness=0.15),                                                      # 5 glossy sphere
        default_material(albedo=(0.1, 0.1, 0.9, 1.0), roughness=1.0, albedo_tex=6, roughness_tex=6, normal_tex=6),          # 6 invalid textures
    ]
    x0, y0, z0, x1, y1, z1 = -2.0, 0.0, -4.0, 2.0, 3.0, 0.0
    room = merge([
        quad_uv((x0, y0, z0), (x0, y0, z1), (x1, y0, z1), (x1, y0, z0), 0, 2.0),   # floor, +y, uv in [0, 2] (wrap / mirror)
        quad_uv((x0, y1, z0), (x1, y1, z0), (x1, y1, z1), (x0, y1, z1), 1),        # ceiling
        quad_uv((x0, y0, z0), (x0, y1, z0), (x0, y1, z1), (x0, y0, z1), 1),        # -x wall
        quad_uv((x1, y0, z0), (x1, y0, z1), (x1, y1, z1), (x1, y1, z0), 6),        # +x wall: invalid textures
        quad_uv((x0, y0, z0), (x1, y0, z0), (x1, y1, z0), (x0, y1, z0), 1),        # -z wall
        quad_uv((x0, y0, z1), (x0, y1, z1), (x1, y1, z1), (x1, y0, z1), 1),        # +z wall
    ])
    screen = quad_uv((-1.6, 0.2, -2.2), (0.2, 0.2, -2.2), (0.2, 2.2, -2.2), (-1.6, 2.2, -2.2), 2)            # facing +z (the camera)
    pane = quad_uv((0.5, 0.1, -1.6), (1.7, 0.1, -1.6), (1.7, 1.9, -1.6), (0.5, 1.9, -1.6), 4, 1.0)
    light = quad_uv((-0.8, 2.99, -3.0), (0.8, 2.99, -3.0), (0.8, 2.99, -1.4), (-0.8, 2.99, -1.4), 3, 1.25)   # facing -y; uv > 1: border
    ball = icosphere(sphere_subdiv, 0.5, (-0.6, 0.5, -3.1), 5)
    meshes = [room, screen, pane, light, ball]
    instances = [Instance(i) for i in range(len(meshes))]
    cam = default_camera(pos=(0.0, 1.4, -0.15), rotation=(0.0, 0.0, 0.0), fov=0.9)
    sc = Scene("textured_example", meshes, instances, mats, cam, width, height, max_ray_depth=max_ray_depth)
    sc.sky_color = (0.0, 0.0, 0.0)
    sc.textures = textures
    return sc


# ---------------------------------------------------------------------------------------------
# S1: atrium (configs 2 and 5)
# ---------------------------------------------------------------------------------------------
def atrium_materials(rng: SplitMix64) -> List[Dict]:
    mats = []
    alb = rng.uniform(8 * 3, 0.2, 0.8).reshape(8, 3)
    for k in range(8):  # diffuse
        mats.append(default_material(albedo=(float(alb[k, 0]), float(alb[k, 1]), float(alb[k, 2]), 1.0), roughness=1.0))
    alb = rng.uniform(6 * 3, 0.2, 0.8).reshape(6, 3)
    for k, rough in enumerate((0.7, 0.5, 0.35, 0.2, 0.1, 0.05)):  # glossy dielectric
        mats.append(default_material(albedo=(float(alb[k, 0]), float(alb[k, 1]), float(alb[k, 2]), 1.0), roughness=rough))
    mats.append(default_material(albedo=(0.95, 0.64, 0.54, 1.0), roughness=0.25, metallic=True))  # copper-ish
    mats.append(default_material(albedo=(0.91, 0.92, 0.92, 1.0), roughness=0.1, metallic=True))   # aluminium-ish
    mats.append(default_material(albedo=(1.0, 1.0, 1.0, 1.0), emission=(10.0, 10.0, 10.0), emission_active=True, roughness=1.0))  # 16 emissive
    return mats


def atrium(target_tris: int = 1_000_000, width: int = 1920, height: int = 1080, max_ray_depth: int = 5, seed: int = 0xB20001) -> Scene:
    """60 x 20 x 30 m hall, 6 x 12 colonnade, displaced floor and ceiling grids, 24 emissive ceiling quads.
    At target_tris = 1 000 000 the columns are 96-gon x 64 rings; smaller targets shrink the tessellation."""
    rng = SplitMix64(seed)
    mats = atrium_materials(rng)
    EMISSIVE = 16

    scale = min(1.0, math.sqrt(target_tris / 1_000_000.0))
    seg = max(8, int(round(96 * scale)) // 4 * 4)
    rings = max(4, int(round(64 * scale)))
    col = column(seg, rings, 0.55, 16.0, 8)
    n_cols = 72
    col_tris = col.num_tris * n_cols

    walls = box_inward((-30.0, 0.0, -15.0), (30.0, 20.0, 15.0), (0, 1, 2, 3, 4, 5))
    walls = Mesh(walls.vertex[4:], walls.normal[4:], walls.uv[4:], walls.material[4:])  # floor/ceiling come from the grids

    lights = []
    for k in range(24):
        cx = -27.5 + 5.0 * (k % 12)
        cz = -6.0 if k < 12 else 6.0
        lights.append(quad((cx - 1.0, 19.7, cz - 0.6), (cx + 1.0, 19.7, cz - 0.6), (cx + 1.0, 19.7, cz + 0.6), (cx - 1.0, 19.7, cz + 0.6), EMISSIVE))
    lights = merge(lights)

    remaining = target_tris - col_tris - walls.num_tris - lights.num_tris
    remaining = max(remaining, 16)
    per_grid = remaining // 2
    nz = max(2, int(math.sqrt(per_grid / 4.0)))
    nx = max(2, per_grid // (2 * nz))

    def floor_h(x, z):
        return 0.05 * (2.0 * _fbm(x * 0.5, z * 0.5, 5, seed) - 1.0)

    def ceil_h(x, z):
        return 20.0 + 0.05 * (2.0 * _fbm(x * 0.5 + 100.0, z * 0.5 - 50.0, 5, seed + 1) - 1.0)

    floor = heightfield(nx, nz, -30.0, -15.0, 30.0, 15.0, floor_h, 9)
    ceil = heightfield(nx, nz, -30.0, -15.0, 30.0, 15.0, ceil_h, 1, flip=True)

    static = [walls, lights, floor, ceil]
    used = col_tris + sum(m.num_tris for m in static)
    filler = target_tris - used
    if filler > 0:  # skirting strip of small quads along the -z wall to hit the exact count
        fq = []
        n_quads = filler // 2
        for k in range(n_quads):
            x0 = -30.0 + 60.0 * k / max(n_quads, 1)
            x1 = -30.0 + 60.0 * (k + 1) / max(n_quads, 1)
            fq.append(quad((x0, 0.0, -14.9), (x1, 0.0, -14.9), (x1, 0.3, -14.9), (x0, 0.3, -14.9), 12))
        if filler % 2:
            fq.append(mesh_from_tris(np.array([[[29.0, 0.3, -14.9], [30.0, 0.3, -14.9], [30.0, 0.6, -14.9]]]), 12))
        if fq:
            static.append(merge(fq))

    meshes = [merge(static), col]
    instances = [Instance(0)]
    rot = rng.uniform(n_cols, 0.0, 2.0 * math.pi)
    sc = rng.uniform(n_cols, 0.95, 1.05)
    k = 0
    for row in range(6):
        for c in range(12):
            x = -27.5 + 5.0 * c
            z = -12.5 + 5.0 * row
            instances.append(Instance(1, (x, 0.0, z), (0.0, float(rot[k]), 0.0), (float(sc[k]), 1.0 + 0.2 * (float(sc[k]) - 1.0), float(sc[k]))))
            k += 1
    # per-column material variety: materials 8..15 cycle through the column mesh rings
    col.material[:] = (8 + (np.arange(col.num_tris) // max(1, col.num_tris // 8)) % 8).astype(np.uint16)

    cam = default_camera(pos=(-26.0, 6.0, 2.0), rotation=(-0.12, -1.35, 0.0), fov=0.9)
    return Scene("atrium", meshes, instances, mats, cam, width, height, max_ray_depth)


def atrium_textured(target_tris: int = 1_000_000, width: int = 1920, height: int = 1080, max_ray_depth: int = 5, seed: int = 0xB20001) -> Scene:
    """S1 with material textures (SURVEY 8f rank 1 measured at the headline size): albedo + roughness + normal maps on the floor,
    an albedo map on the ceiling, perforated columns (albedo map with alpha cut-outs + normal map: the any-hit alpha test runs
    on 92 % of the triangles) and luminance-textured ceiling lights."""
    sc = atrium(target_tris, width, height, max_ray_depth, seed)
    rng = SplitMix64(seed ^ 0x7E57)

    def noise_u8(h, w, c, lo, hi, cell):
        """Value noise: random lattice of (h / cell, w / cell) up-sampled by repetition + per-texel jitter."""
        base = rng.uniform((h // cell) * (w // cell) * c, lo, hi).reshape(h // cell, w // cell, c)
        img = np.kron(base, np.ones((cell, cell, 1), np.float32))
        img += rng.uniform(h * w * c, -6.0, 6.0).reshape(h, w, c)
        return np.clip(img, 0, 255).astype(np.uint8)

    alb = np.concatenate([noise_u8(1024, 1024, 3, 60, 230, 16), np.full((1024, 1024, 1), 255, np.uint8)], axis=2)
    rough = noise_u8(512, 512, 1, 60, 250, 8)
    nxy = noise_u8(512, 512, 2, 96, 160, 4).astype(np.float32) / 255.0 * 2.0 - 1.0
    nz = np.sqrt(np.maximum(1.0 - (nxy ** 2).sum(axis=2, keepdims=True), 0.0))
    nmap = np.concatenate([np.round((np.concatenate([nxy, nz], axis=2) * 0.5 + 0.5) * 255).astype(np.uint8), np.full((512, 512, 1), 255, np.uint8)], axis=2)
    holes = rng.uniform(32 * 32, 0.0, 1.0).reshape(32, 32) < 0.3
    cut = np.concatenate([noise_u8(256, 256, 3, 90, 240, 8), np.where(np.kron(holes, np.ones((8, 8), bool)), 0, 255).astype(np.uint8)[..., None]], axis=2)
    lum = np.concatenate([noise_u8(64, 64, 3, 120, 255, 8), np.full((64, 64, 1), 255, np.uint8)], axis=2)
    sc.textures = [
        dict(data=alb, wrap_u=0, wrap_v=0, filter=1, gamma=2.2),
        dict(data=rough, wrap_u=0, wrap_v=0, filter=1, gamma=1.0),
        dict(data=nmap, wrap_u=0, wrap_v=0, filter=1, gamma=1.0),
        dict(data=cut, wrap_u=0, wrap_v=0, filter=1, gamma=2.2),
        dict(data=lum, wrap_u=0, wrap_v=0, filter=1, gamma=1.0),
    ]
    sc.materials[9].update(albedo_tex=0, roughness_tex=1, normal_tex=2)   # floor
    sc.materials[1].update(albedo_tex=0)                                   # ceiling
    sc.materials[8].update(albedo_tex=3, normal_tex=2)                     # columns: cut-outs
    sc.materials[16].update(luminance_tex=4, emission_scale=10.0)          # ceiling lights
    sc.name = "atrium_textured"
    return sc


# ---------------------------------------------------------------------------------------------
# S2: terrain + lanterns (config 3)
# ---------------------------------------------------------------------------------------------
def terrain(grid: int = 2236, n_lanterns: int = 50_000, width: int = 1920, height: int = 1080, max_ray_depth: int = 5,
            seed: int = 0xB20002) -> Scene:
    rng = SplitMix64(seed)
    mats = [
        default_material(albedo=(0.35, 0.3, 0.25, 1.0), roughness=1.0),
        default_material(albedo=(0.2, 0.4, 0.15, 1.0), roughness=0.8),
    ]

    def h(x, z):
        return 40.0 * _fbm(x * 0.01, z * 0.01, 6, seed, ridged=True)

    ground = heightfield(grid, grid, -500.0, -500.0, 500.0, 500.0, h, 0)
    slope = ground.normal[:, 0, 1]
    ground.material[:] = np.where(slope > 0.85, 1, 0).astype(np.uint16)

    # lantern quads: jittered grid (Poisson-disc-like minimum spacing) 0.5 m above the ground, facing down
    side = int(math.ceil(math.sqrt(n_lanterns)))
    cell = 1000.0 / side
    jx = rng.uniform(side * side, 0.15, 0.85).astype(np.float64)
    jz = rng.uniform(side * side, 0.15, 0.85).astype(np.float64)
    ke = rng.uniform(n_lanterns, 2.0, 20.0)
    gi, gk = np.divmod(np.arange(side * side), side)
    lx = (-500.0 + (gi + jx) * cell)[:n_lanterns]
    lz = (-500.0 + (gk + jz) * cell)[:n_lanterns]
    ly = h(lx, lz) + 1.5
    s = 0.15
    p0 = np.stack([lx - s, ly, lz - s], -1)
    p1 = np.stack([lx + s, ly, lz - s], -1)
    p2 = np.stack([lx + s, ly, lz + s], -1)
    p3 = np.stack([lx - s, ly, lz + s], -1)
    tv = np.concatenate([np.stack([p0, p1, p2], 1), np.stack([p0, p2, p3], 1)])  # normal -y
    # 16 emission levels keep the material table small
    levels = 16
    base = len(mats)
    for q in range(levels):
        e = 2.0 + 18.0 * (q + 0.5) / levels
        mats.append(default_material(albedo=(1.0, 1.0, 1.0, 1.0), emission=(e, 0.8 * e, 0.5 * e), emission_active=True, roughness=1.0))
    lvl = np.clip(((ke - 2.0) / 18.0 * levels).astype(np.int64), 0, levels - 1)
    lmat = (base + np.concatenate([lvl, lvl])).astype(np.uint16)
    lanterns = mesh_from_tris(tv, lmat)

    cam = default_camera(pos=(0.0, 70.0, 0.0), rotation=(-0.35, 0.6, 0.0), fov=0.9)
    sc = Scene("terrain", [ground, lanterns], [Instance(0), Instance(1)], mats, cam, width, height, max_ray_depth)
    sc.sky_color = (0.02, 0.03, 0.06)
    return sc


# ---------------------------------------------------------------------------------------------
# S3: divergence stress (config 4)
# ---------------------------------------------------------------------------------------------
def divergence(target_tris: int = 1_000_000, width: int = 1920, height: int = 1080, max_ray_depth: int = 8, seed: int = 0xB20003) -> Scene:
    sc = atrium(target_tris, width, height, max_ray_depth)
    rng = SplitMix64(seed)
    mats = []
    alb = rng.uniform(16 * 3, 0.3, 0.9).reshape(16, 3)
    for k in range(4):
        mats.append(default_material(albedo=(*map(float, alb[k]), 1.0), roughness=(0.05, 0.15, 0.3, 0.5)[k]))
    for k in range(4, 8):
        mats.append(default_material(base_substrate=1, albedo=(*map(float, alb[k]), 1.0), roughness=(0.02, 0.1, 0.2, 0.4)[k - 4], refraction_index=1.5))
    for k in range(8, 12):
        e = (2.0, 4.0, 6.0, 8.0)[k - 8]
        mats.append(default_material(albedo=(1.0, 1.0, 1.0, 1.0), emission=(e, e, e), emission_active=True, roughness=1.0))
    for k in range(12, 16):
        mats.append(default_material(albedo=(*map(float, alb[k]), 1.0), roughness=1.0))
    sc.materials = mats
    for mi, m in enumerate(sc.meshes):
        t = np.arange(m.num_tris, dtype=np.uint64)
        with np.errstate(over="ignore"):
            hsh = (t + np.uint64(mi * 7919 + seed)) * np.uint64(0x9E3779B97F4A7C15)
            hsh = (hsh ^ (hsh >> np.uint64(29))) * np.uint64(0xBF58476D1CE4E5B9)
            hsh ^= hsh >> np.uint64(32)
        cls = (hsh % np.uint64(64)).astype(np.int64)
        # 1/64 of the triangles emit; the rest split evenly over glossy / translucent / diffuse
        mat = np.where(cls == 0, 8 + (hsh >> np.uint64(8)) % np.uint64(4), 0).astype(np.int64)
        rest = cls > 0
        grp = (cls % 3)
        sub = ((hsh >> np.uint64(16)) % np.uint64(4)).astype(np.int64)
        mat = np.where(rest & (grp == 0), sub, mat)
        mat = np.where(rest & (grp == 1), 4 + sub, mat)
        mat = np.where(rest & (grp == 2), 12 + sub, mat)
        m.material[:] = mat.astype(np.uint16)
    # open ceiling: drop the ceiling grid (material 1 in the static mesh before re-hash is gone; cut by height)
    st = sc.meshes[0]
    keep = st.vertex[:, :, 1].min(axis=1) < 19.0
    sc.meshes[0] = Mesh(st.vertex[keep], st.normal[keep], st.uv[keep], st.material[keep])
    sc.name = "divergence"
    sc.sky_color = (1.0, 1.0, 1.0)
    return sc


# ---------------------------------------------------------------------------------------------
# .obj / .mtl / .lum writers (inputs of the reference's own loaders, host/wavefront.c and host/lum_v4.c)
# ---------------------------------------------------------------------------------------------
def write_png(path: str, img: np.ndarray, gamma: Optional[float] = None) -> None:
    """Minimal PNG encoder (zlib from the standard library): (H, W) or (H, W, C) uint8 / uint16, C in 1..4, filter 0."""
    import struct
    import zlib

    a = np.asarray(img)
    if a.ndim == 2:
        a = a[:, :, None]
    h, w, c = a.shape
    ctype = {1: 0, 2: 4, 3: 2, 4: 6}[c]
    depth = 8 if a.dtype == np.uint8 else 16
    rows = a.astype(">u2") if depth == 16 else a
    raw = b"".join(b"\x00" + rows[y].tobytes() for y in range(h))

    def chunk(t, d):
        return struct.pack(">I", len(d)) + t + d + struct.pack(">I", zlib.crc32(t + d) & 0xFFFFFFFF)

    data = b"\x89PNG\r\n\x1a\n" + chunk(b"IHDR", struct.pack(">IIBBBBB", w, h, depth, ctype, 0, 0, 0))
    if gamma is not None:
        data += chunk(b"gAMA", struct.pack(">I", int(round(100000.0 / gamma))))
    data += chunk(b"IDAT", zlib.compress(raw, 6)) + chunk(b"IEND", b"")
    with open(path, "wb") as f:
        f.write(data)


def write_obj(scene: Scene, obj_path: str, mtl_name: Optional[str] = None) -> None:
    """Writes all active instances baked to world space? No: writes mesh 0..n as objects in mesh space; callers
    that need instancing use the API. v/vt/vn per corner, usemtl per material run. Textures of the scene are written as
    tex<k>.png next to the file (fp32 data quantised to 16 bit; the sampler modes are not expressible in *.mtl: the loader
    gives every texture wrap + linear, texture.c:77-88) and referenced by map_Kd / map_Ke / map_Ns / map_refl / map_Bump."""
    import os
    mtl_name = mtl_name or (os.path.splitext(os.path.basename(obj_path))[0] + ".mtl")
    for k, t in enumerate(getattr(scene, "textures", None) or []):
        if t.get("data") is None:
            continue  # an invalid texture: the file is simply missing
        d = np.asarray(t["data"])
        if d.dtype == np.float32:
            d = np.round(np.clip(d, 0.0, 1.0) * 65535.0).astype(np.uint16)
        write_png(os.path.join(os.path.dirname(obj_path), f"tex{k}.png"), d, gamma=t.get("gamma", 1.0))
    with open(os.path.join(os.path.dirname(obj_path), mtl_name), "w") as f:
        for i, m in enumerate(scene.materials):
            f.write(f"newmtl mat{i}\n")
            f.write("Kd %.6f %.6f %.6f\n" % tuple(m["albedo"][:3]))
            f.write("d %.6f\n" % m["albedo"][3])
            if m["emission_active"]:
                f.write("Ke %.6f %.6f %.6f\n" % tuple(m["emission"]))
            # wavefront.c maps Ns -> roughness = 1 - Ns / 1000
            f.write("Ns %.6f\n" % ((1.0 - m["roughness"]) * 1000.0))
            f.write("Ni %.6f\n" % m["refraction_index"])
            if m["metallic"]:
                f.write("Ks 1.0 1.0 1.0\n")
            for key, stmt in (("albedo_tex", "map_Kd"), ("luminance_tex", "map_Ke"), ("roughness_tex", "map_Ns"), ("metallic_tex", "map_refl"),
                              ("normal_tex", "map_Bump")):
                if m.get(key, 0xFFFF) != 0xFFFF:
                    f.write(f"{stmt} tex{m[key]}.png\n")
            f.write("\n")
    with open(obj_path, "w") as f:
        f.write(f"mtllib {mtl_name}\n")
        base = 1
        for mi, m in enumerate(scene.meshes):
            f.write(f"o mesh{mi}\n")
            v = m.vertex.reshape(-1, 3)
            n = m.normal.reshape(-1, 3)
            t = m.uv.reshape(-1, 2)
            f.write("".join("v %.9g %.9g %.9g\n" % tuple(r) for r in v))
            f.write("".join("vt %.9g %.9g\n" % tuple(r) for r in t))
            f.write("".join("vn %.9g %.9g %.9g\n" % tuple(r) for r in n))
            cur = -1
            lines = []
            for k in range(m.num_tris):
                if int(m.material[k]) != cur:
                    cur = int(m.material[k])
                    lines.append(f"usemtl mat{cur}\n")
                a = base + 3 * k
                lines.append(f"f {a}/{a}/{a} {a + 1}/{a + 1}/{a + 1} {a + 2}/{a + 2}/{a + 2}\n")
            f.write("".join(lines))
            base += 3 * m.num_tris
