"""TEST INFRASTRUCTURE ONLY — host restatement of the in-kernel noise of MIXGRPO_SRC_PHILOX (csrc/common.cuh,
philox4x32_10 / philox_normal4; the reference draws its noise with randn_tensor / randn_like, SU:189-194, SU:238, so
there is nothing in the reference to pin this to: the contract is "N(0,1), a pure function of (seed, offset, element)").
numpy, fp32 Box-Muller; the device uses __logf/__sincosf, so values agree to ~1e-6 absolute, not bitwise."""
import numpy as np

M0, M1, W0, W1 = np.uint64(0xD2511F53), np.uint64(0xCD9E8D57), np.uint32(0x9E3779B9), np.uint32(0xBB67AE85)
MASK = np.uint64(0xFFFFFFFF)


def philox4x32_10(c0, c1, c2, c3, k0, k1):
    c0, c1, c2, c3 = (a.astype(np.uint32) for a in (c0, c1, c2, c3))
    k0, k1 = np.uint32(k0), np.uint32(k1)
    with np.errstate(over="ignore"):
        for _ in range(10):
            p0 = M0 * c0.astype(np.uint64)
            p1 = M1 * c2.astype(np.uint64)
            hi0, lo0 = (p0 >> np.uint64(32)).astype(np.uint32), (p0 & MASK).astype(np.uint32)
            hi1, lo1 = (p1 >> np.uint64(32)).astype(np.uint32), (p1 & MASK).astype(np.uint32)
            c0, c1, c2, c3 = hi1 ^ c1 ^ k0, lo1, hi0 ^ c3 ^ k1, lo0
            k0, k1 = np.uint32(k0 + W0), np.uint32(k1 + W1)
    return c0, c1, c2, c3


def normal(seed: int, offset: int, numel: int) -> np.ndarray:
    """N(0,1) samples for elements 0..numel-1 of a tensor, fp32."""
    nq = (numel + 3) // 4
    q = np.arange(nq, dtype=np.uint64)
    off = np.uint64(offset)
    r = philox4x32_10(q & MASK, q >> np.uint64(32), np.full(nq, off & MASK, dtype=np.uint64), np.full(nq, off >> np.uint64(32), dtype=np.uint64),
                      seed & 0xFFFFFFFF, ((seed >> 32) & 0xFFFFFFFF) ^ 0x6d697867)
    u = [((a >> np.uint32(8)).astype(np.float32) + np.float32(0.5)) * np.float32(5.9604644775390625e-08) for a in r]
    ra = np.sqrt(np.float32(-2.0) * np.log(u[0])).astype(np.float32)
    rb = np.sqrt(np.float32(-2.0) * np.log(u[2])).astype(np.float32)
    two_pi = np.float32(6.283185307179586)
    z = np.stack([ra * np.cos(two_pi * u[1]), ra * np.sin(two_pi * u[1]), rb * np.cos(two_pi * u[3]), rb * np.sin(two_pi * u[3])], axis=1)
    return z.astype(np.float32).reshape(-1)[:numel]
