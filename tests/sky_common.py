"""Shared inputs of the procedural-sky tests (test_sky_oracle.py, test_sky_gpu.py, tests/golden/make_sky_golden.py): the sky variants,
a fixed set of miss rays, and conversions between the three implementations' record types.

Everything here is a pure function of its seeds, so the golden fixture tests/golden/sky_ref.npz (outputs of the REFERENCE's own
kernels for these inputs, made on a B200 by tests/golden/make_sky_golden.py) stays valid as long as this file does not change."""
import numpy as np

# name -> overrides of the reference's defaults (sky.c:6-42)
SKY_VARIANTS = {
    "default": {},
    # low sun, hazy, no ozone, offset observer, denser star field with its own seed
    "evening": dict(altitude=0.12, azimuth=1.0, mie_density=2.5, mie_diameter=1.2, ozone_absorption=0, rayleigh_falloff=7.0,
                    geometry_offset=(0.5, 1.5, -0.25), steps=24, stars_count=3000, stars_seed=11, stars_intensity=2.0, sun_strength=1.5,
                    ground_visibility=30.0, multiscattering_factor=0.8, base_density=1.2),
    # sun below the horizon, moon up: stars and multiscattering dominate
    "night": dict(altitude=-0.35, azimuth=4.0, moon_altitude=0.6, moon_azimuth=2.0, stars_count=20000, stars_seed=3, steps=16),
}

# HDRI mode (sky mode 1): name -> (table edge, samples per texel); 8 < 32 buckets, 40 gives the first 8 lanes two samples
HDRI_VARIANTS = {"default": (48, 8), "evening": (40, 40)}
HDRI_ORIGIN = (1.0, 2.0, -3.0)

STATE_DELTA_PATH, STATE_CAMERA_DIRECTION, STATE_ALLOW_EMISSION, STATE_ALLOW_AMBIENT = 0x01, 0x02, 0x08, 0x10


def star_direction(altitude, azimuth):
    """inverse of the lookup in sky_compute_atmosphere (sky.cuh:470-475): altitude = asin(ray.y), azimuth = atan2(-z, -x) + pi"""
    return np.stack([np.cos(azimuth) * np.cos(altitude), np.sin(altitude), np.sin(azimuth) * np.cos(altitude)], axis=-1).astype(np.float32)


def miss_rays(sun_pos, stars, width, height, seed=5, n_sphere=384, n_sun=96, n_horizon=96, n_stars=64):
    """Miss rays of a test: uniformly distributed directions, a cluster inside and around the sun's disc (angular radius 4.65 mrad),
    a band around the horizon, and rays aimed at catalogue stars. -> dict(origin (n, 3) world space [m], ray (n, 3), state (n,),
    pixel (n, 2), sample (n,))"""
    rng = np.random.default_rng(seed)
    d = rng.normal(size=(n_sphere, 3))
    sun_dir = np.asarray(sun_pos, np.float64) / np.linalg.norm(sun_pos)
    t1 = np.cross(sun_dir, [0.0, 1.0, 0.0])
    t1 /= np.linalg.norm(t1)
    t2 = np.cross(sun_dir, t1)
    ang = rng.uniform(0.0, 0.009, n_sun)          # up to about twice the disc radius
    phi = rng.uniform(0.0, 2.0 * np.pi, n_sun)
    ds = sun_dir[None] * np.cos(ang)[:, None] + (t1[None] * np.cos(phi)[:, None] + t2[None] * np.sin(phi)[:, None]) * np.sin(ang)[:, None]
    az = rng.uniform(0.0, 2.0 * np.pi, n_horizon)
    alt = rng.uniform(-0.03, 0.08, n_horizon)
    dh = np.stack([np.cos(az) * np.cos(alt), np.sin(alt), np.sin(az) * np.cos(alt)], axis=-1)
    above = stars[stars[:, 0] > 0.05]
    pick = above[rng.integers(0, max(len(above), 1), n_stars)] if len(above) else np.zeros((0, 4), np.float32)
    dst = star_direction(pick[:, 0].astype(np.float64), pick[:, 1].astype(np.float64)) if len(pick) else np.zeros((0, 3))
    ray = np.concatenate([d, ds, dh, dst]).astype(np.float64)
    ray = (ray / np.linalg.norm(ray, axis=1, keepdims=True)).astype(np.float32)
    n = ray.shape[0]
    origin = np.zeros((n, 3), np.float32)
    origin[:, 1] = rng.uniform(0.0, 30.0, n)      # metres above the scene origin
    origin[::7] = rng.uniform(-2000.0, 2000.0, (len(origin[::7]), 3)) + (0.0, 2500.0, 0.0)
    origin[5::31, 1] = 12000.0                   # an observer at 12 km
    state = np.full(n, STATE_ALLOW_AMBIENT | STATE_ALLOW_EMISSION | STATE_CAMERA_DIRECTION, np.uint32)
    state[1::3] = STATE_ALLOW_AMBIENT                       # bounce rays that may not see the sun's disc (NEE covers it)
    state[2::5] = STATE_ALLOW_AMBIENT | STATE_ALLOW_EMISSION
    state[3::17] = STATE_ALLOW_EMISSION                     # no ALLOW_AMBIENT: the miss adds nothing
    pixel = np.stack([rng.integers(0, width, n), rng.integers(0, height, n)], axis=-1).astype(np.uint32)
    sample = rng.integers(0, 4, n).astype(np.uint32) * 17  # few distinct ids: the product hook shades one sample id per call
    return dict(origin=origin, ray=ray, state=state, pixel=pixel, sample=sample)


def record_pack(rgb):
    """record_pack (math.cuh:1580-1593): 3 x 21-bit truncated floats in 2 words"""
    b = np.ascontiguousarray(rgb, np.float32).view(np.uint32) >> 11
    r, g, bl = b[:, 0].astype(np.uint64), b[:, 1].astype(np.uint64), b[:, 2].astype(np.uint64)
    x = (r | (g << np.uint64(21))) & np.uint64(0xFFFFFFFF)
    y = ((g >> np.uint64(11)) | (bl << np.uint64(10))) & np.uint64(0xFFFFFFFF)
    return np.stack([x, y], axis=-1).astype(np.uint32)


def record_unpack(rec):
    rec = np.asarray(rec, np.uint32)
    x, y = rec[:, 0].astype(np.uint64), rec[:, 1].astype(np.uint64)
    r = (x & np.uint64(0x1FFFFF)) << np.uint64(11)
    g = (((x >> np.uint64(21)) | ((y & np.uint64(0x3FF)) << np.uint64(11))) << np.uint64(11)) & np.uint64(0xFFFFFFFF)
    b = ((y >> np.uint64(10)) << np.uint64(11)) & np.uint64(0xFFFFFFFF)
    return np.stack([r, g, b], axis=-1).astype(np.uint32).view(np.float32)


def rel_err(got, want, floor):
    got, want = np.asarray(got, np.float64), np.asarray(want, np.float64)
    return np.abs(got - want) / np.maximum(np.abs(want), floor)


# the moon's surface (sky.cuh:440-475): a nearly full moon and rays across its disc (angular radius 4.5 mrad) from the scene origin
MOON_SKY = dict(altitude=0.1, azimuth=0.6, moon_altitude=0.3, moon_azimuth=3.74, moon_tex_offset=0.13, stars_count=0)


def moon_rays(moon_pos, width, height, n=512, seed=11):
    rng = np.random.default_rng(seed)
    moon_dir = np.asarray(moon_pos, np.float64) - (0.0, 6371.0 + 0.1, 0.0)   # observer at the origin, default geometry offset (0, 0.1 km, 0)
    moon_dir /= np.linalg.norm(moon_dir)
    t1 = np.cross(moon_dir, [0.0, 1.0, 0.0])
    t1 /= np.linalg.norm(t1)
    t2 = np.cross(moon_dir, t1)
    ang = rng.uniform(0.0, 0.0055, n)
    phi = rng.uniform(0.0, 2.0 * np.pi, n)
    ray = moon_dir[None] * np.cos(ang)[:, None] + (t1[None] * np.cos(phi)[:, None] + t2[None] * np.sin(phi)[:, None]) * np.sin(ang)[:, None]
    ray = (ray / np.linalg.norm(ray, axis=1, keepdims=True)).astype(np.float32)
    return dict(origin=np.zeros((n, 3), np.float32), ray=ray, angle=ang,
                state=np.full(n, STATE_ALLOW_AMBIENT | STATE_ALLOW_EMISSION | STATE_CAMERA_DIRECTION, np.uint32),
                pixel=np.stack([rng.integers(0, width, n), rng.integers(0, height, n)], axis=-1).astype(np.uint32), sample=np.zeros(n, np.uint32))


# aerial perspective (kernels.cuh:356-389): hit segments of 1 m .. 30 km under a 20 x denser atmosphere
AERIAL_SKY = dict(azimuth=1.2, altitude=1.0, aerial_perspective=1, base_density=20.0, mie_density=2.0, steps=120)


def aerial_segments(width, height, n=384, seed=23):
    """-> dict(origin (n, 3) [m], ray (n, 3), t (n,) [m], pixel (n, 2), sample (n,), record (n, 3))"""
    rng = np.random.default_rng(seed)
    d = rng.normal(size=(n, 3))
    d[:, 1] = np.abs(d[:, 1]) * 0.3            # mostly horizontal, never into the ground
    d = (d / np.linalg.norm(d, axis=1, keepdims=True)).astype(np.float32)
    origin = np.zeros((n, 3), np.float32)
    origin[:, 1] = rng.uniform(1.0, 500.0, n)
    t = np.exp(rng.uniform(np.log(1.0), np.log(30000.0), n)).astype(np.float32)
    return dict(origin=origin, ray=d, t=t, pixel=np.stack([rng.integers(0, width, n), rng.integers(0, height, n)], axis=-1).astype(np.uint32),
                sample=(rng.integers(0, 4, n) * 5).astype(np.uint32), record=rng.uniform(0.05, 1.0, (n, 3)).astype(np.float32))
