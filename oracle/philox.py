"""numpy Philox4x32-10 + Box-Muller with the same stream addressing as
csrc/philox.cuh (uniforms are bit-identical; normals differ only by libm ulps)."""
import numpy as np

M0, M1 = np.uint64(0xD2511F53), np.uint64(0xCD9E8D57)
W0, W1 = 0x9E3779B9, 0xBB67AE85
MASK = np.uint64(0xFFFFFFFF)


def philox4x32(c0, c1, c2, c3, k0, k1):
    """All arguments uint64 arrays holding 32-bit values; returns 4 such arrays."""
    c0, c1, c2, c3 = (np.asarray(c, dtype=np.uint64) for c in (c0, c1, c2, c3))
    k0, k1 = int(k0) & 0xFFFFFFFF, int(k1) & 0xFFFFFFFF
    for _ in range(10):
        p0, p1 = M0 * c0, M1 * c2
        hi0, lo0 = p0 >> np.uint64(32), p0 & MASK
        hi1, lo1 = p1 >> np.uint64(32), p1 & MASK
        c0, c1, c2, c3 = hi1 ^ c1 ^ np.uint64(k0), lo1, hi0 ^ c3 ^ np.uint64(k1), lo0
        k0, k1 = (k0 + W0) & 0xFFFFFFFF, (k1 + W1) & 0xFFFFFFFF
    return c0, c1, c2, c3


def u52(hi, lo):
    """(k + 0.5) * 2^-52 with k = (low 20 bits of hi):(lo) - same definition as csrc/philox.cuh."""
    k = ((hi & np.uint64(0xFFFFF)) << np.uint64(32)) | lo
    return (k.astype(np.float64) + 0.5) * 2.220446049250313e-16


def _blocks(paths, block, kind, seed, stream):
    paths = np.asarray(paths, dtype=np.uint64)
    o = philox4x32(paths & MASK, paths >> np.uint64(32), np.full_like(paths, block), np.full_like(paths, kind),
                   seed, stream)
    return u52(o[0], o[1]), u52(o[2], o[3])


def normal_pair(paths, block, seed, stream):
    u1, u2 = _blocks(paths, block, 0, seed, stream)
    rad = np.sqrt(-2.0 * np.log(u1))
    # the angle uses k 2^-52 (the mantissa of the second word pair), without the half-step offset
    ang = 2.0 * np.pi * (u2 - 1.1102230246251565e-16)
    # sincospi(2 u2): evaluate through the reduced argument like the device does
    return rad * np.cos(ang), rad * np.sin(ang)


def normals(paths, n_sub, dim, seed, stream=0):
    """[n_sub, len(paths), dim] standard normals; normal #n of a path is element n&1 of block n>>1."""
    paths = np.asarray(paths, dtype=np.uint64)
    total = n_sub * dim
    flat = np.empty((total + 1, paths.shape[0]))
    for b in range((total + 1) // 2):
        z0, z1 = normal_pair(paths, b, seed, stream)
        flat[2 * b], flat[2 * b + 1] = z0, z1
    return flat[:total].reshape(n_sub, dim, paths.shape[0]).transpose(0, 2, 1).copy()


def uniforms(paths, count, seed, stream=0):
    """[count, len(paths)] uniforms (QE); uniform #m is element m&1 of block m>>1, kind 1."""
    paths = np.asarray(paths, dtype=np.uint64)
    out = np.empty((count + 1, paths.shape[0]))
    for b in range((count + 1) // 2):
        out[2 * b], out[2 * b + 1] = _blocks(paths, b, 1, seed, stream)
    return out[:count]
