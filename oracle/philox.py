"""numpy restatement of the device ASE generator (TEST INFRASTRUCTURE): Philox4x32-10 keyed by
the 64-bit seed, counter = (sample index lo, hi, column, realization), two Box-Muller pairs per
counter -> complex standard normals for the X and Y polarizations (pmx_api.cu: philox4x32_10,
pmx_cnormal).  The reference draws ASE with randn (ampliflat.m:132-135); only the distribution
is shared, so parity of a noisy run is checked by feeding the same numbers through
ampliflat's options.noise."""
import numpy as np

M0, M1 = np.uint64(0xD2511F53), np.uint64(0xCD9E8D57)
W0, W1 = 0x9E3779B9, 0xBB67AE85
MASK = np.uint64(0xFFFFFFFF)


def philox4x32_10(c0, c1, c2, c3, k0, k1):
    c0, c1, c2, c3 = (np.asarray(c, dtype=np.uint64) & MASK for c in (c0, c1, c2, c3))
    k0, k1 = int(k0) & 0xFFFFFFFF, int(k1) & 0xFFFFFFFF
    for _ in range(10):
        p0, p1 = M0 * c0, M1 * c2
        hi0, lo0, hi1, lo1 = p0 >> np.uint64(32), p0 & MASK, p1 >> np.uint64(32), p1 & MASK
        c0, c1, c2, c3 = (hi1 ^ c1 ^ np.uint64(k0)) & MASK, lo1, (hi0 ^ c3 ^ np.uint64(k1)) & MASK, lo0
        k0, k1 = (k0 + W0) & 0xFFFFFFFF, (k1 + W1) & 0xFFFFFFFF
    return c0, c1, c2, c3


def cnormal(a, b):
    u1 = (a.astype(np.float64) + 0.5) / 4294967296.0
    u2 = (b.astype(np.float64) + 0.5) / 4294967296.0
    r = np.sqrt(-2.0 * np.log(u1))
    return r * np.cos(2 * np.pi * u2) + 1j * r * np.sin(2 * np.pi * u2)


def ase_normals(n, col, realization, seed):
    """-> (nx, ny) complex standard normals of one column of one realization, [n] each."""
    idx = np.arange(n, dtype=np.uint64)
    r = philox4x32_10(idx & MASK, idx >> np.uint64(32), np.full(n, col, dtype=np.uint64),
                      np.full(n, realization, dtype=np.uint64), seed & 0xFFFFFFFF, (seed >> 32) & 0xFFFFFFFF)
    return cnormal(r[0], r[1]), cnormal(r[2], r[3])
