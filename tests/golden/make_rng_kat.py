"""Independent pure-Python restatement of the reference sampler (cuda/random.cuh:172-333) used to produce
rng_kat.json, the known-answer vectors of tests/test_oracle_core.py. Written from the reference source without
looking at oracle/orc_core.c; run `python tests/golden/make_rng_kat.py` to regenerate."""
import json
import os

import numpy as np

M = 0xFFFFFFFF
HERE = os.path.dirname(os.path.abspath(__file__))
BN = np.fromfile(os.path.join(HERE, "..", "..", "luminary_b200", "data", "bluenoise_2D.bin"), dtype=np.uint32)
TARGET_COUNT = 577


def swap(x):
    return ((x >> 16) | (x << 16)) & M


def squares32(key, ctr):
    x = (ctr * key) & M
    y = x
    z = (y + key) & M
    x = (x * x + y) & M
    x = swap(x)
    x = (x * x + z) & M
    x = swap(x)
    x = (x * x + y) & M
    x = swap(x)
    x = (x * x + z) & M
    z = x
    x = swap(x)
    return z ^ ((x * x + y) & M)


def squares16(key, ctr):
    x = (ctr * key) & M
    y = x
    z = (y + key) & M
    x = (x * x + y) & M
    x = swap(x)
    x = (x * x + z) & M
    x = swap(x)
    return ((x * x + y) & M) >> 16


def brev(x):
    return int(format(x & M, "032b")[::-1], 2)


def lk(x, seed):
    x = (x + seed) & M
    for c in (0x6C50B47C, 0xB82F1E52, 0xC7AFE638, 0x8D22F6E6):
        x ^= (x * c) & M
    return x


def owen(x, seed):
    return brev(lk(brev(x), seed))


def hcomb(seed, v):
    return seed ^ ((v + ((seed << 6) & M) + (seed >> 2)) & M)


def sobol_p(v):
    v ^= (v << 16) & M
    v ^= ((v & 0x00FF00FF) << 8) & M
    v ^= ((v & 0x0F0F0F0F) << 4) & M
    v ^= ((v & 0x33333333) << 2) & M
    v ^= ((v & 0x55555555) << 1) & M
    return v


def sobol(offset, dim):
    seed = squares32(0xFCBD6E15, dim)
    j = lk(brev(offset), seed)
    return owen(j, hcomb(seed, 0)), owen(sobol_p(j), hcomb(seed, 1))


def random_2d_base(target, px, py, seq, depth):
    dim = (target + depth * TARGET_COUNT) & M
    qx, qy = sobol(seq, dim)
    ox = (((1 + dim) * 3242174889) & M) >> 24
    oy = (((1 + dim) * 2447445413) & M) >> 24
    n = int(BN[((px + ox) & 0xFF) + ((py + oy) & 0xFF) * 256])
    return (qx + (n & 0xFFFF0000)) & M, (qy + ((n << 16) & M)) & M


def main():
    out = {"squares32": [], "squares16": [], "sobol": [], "random_2d_base": []}
    for key, ctr in [(0xFCBD6E15, 0), (0xFCBD6E15, 1), (0xFCBD6E15, 39), (0xFCBD6E15, 577 * 5 + 61), (0x12345679, 0xDEADBEEF), (1, 1)]:
        out["squares32"].append([key, ctr, squares32(key, ctr)])
        out["squares16"].append([key, ctr, squares16(key, ctr)])
    for off, dim in [(0, 0), (1, 0), (2, 39), (63, 51), (1023, 577 + 43), (0xFFFFF, 577 * 4 + 404), (12345, 63)]:
        x, y = sobol(off, dim)
        out["sobol"].append([off, dim, x, y])
    for args in [(63, 0, 0, 0, 0), (63, 0, 0, 17, 0), (39, 100, 200, 3, 0), (51, 1919, 1079, 63, 4), (387, 16383, 16383, 1048575, 5),
                 (404 + 7, 255, 256, 1000, 2), (575, 1, 2, 3, 1)]:
        x, y = random_2d_base(*args)
        out["random_2d_base"].append(list(args) + [x, y])
    with open(os.path.join(HERE, "rng_kat.json"), "w") as f:
        json.dump(out, f, indent=1)


if __name__ == "__main__":
    main()
