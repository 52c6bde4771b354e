"""TEST INFRASTRUCTURE (oracle): numpy restatement of the device noise generator of csrc/kernels_elem.cu (aid_philox_normal).

The reference draws sampler noise with torch's CPU generator and copies it to the device (diff_params/edm.py:94,
testing/edm_sampler_inpainting.py:212); the product can instead generate it on the device, keyed so that a clip's noise depends
only on (seed, stream id, global clip index, draw, element).  This file is the CPU definition of that stream:

  * Philox4x32-10 (Salmon, Moraes, Dror, Shaw: "Parallel random numbers: as easy as 1, 2, 3", SC'11): pinned below against the
    known-answer vectors of the Random123 distribution (kat_vectors: philox4x32 10);
  * counter = (element // 4, draw, clip, stream_id), key = (seed & 0xffffffff, seed >> 32);
  * u = ((r >> 8) + 0.5) * 2^-24; the outputs (r0, r1), (r2, r3) give two Box-Muller pairs
    n = sqrt(-2 ln u_a) * (cos, sin)(2 pi u_b) -> elements 4q, 4q+1, 4q+2, 4q+3.

Only tests/ may import this module.
"""
import numpy as np

M0, M1 = np.uint64(0xD2511F53), np.uint64(0xCD9E8D57)
W0, W1 = np.uint32(0x9E3779B9), np.uint32(0xBB67AE85)
MASK = np.uint64(0xFFFFFFFF)

# Random123 known-answer tests for philox4x32-10: (counter, key) -> output
KAT = [
    ((0x00000000, 0x00000000, 0x00000000, 0x00000000), (0x00000000, 0x00000000), (0x6627e8d5, 0xe169c58d, 0xbc57ac4c, 0x9b00dbd8)),
    ((0xffffffff, 0xffffffff, 0xffffffff, 0xffffffff), (0xffffffff, 0xffffffff), (0x408f276d, 0x41c83b0e, 0xa20bc7c6, 0x6d5451fd)),
    ((0x243f6a88, 0x85a308d3, 0x13198a2e, 0x03707344), (0xa4093822, 0x299f31d0), (0xd16cfe09, 0x94fdcceb, 0x5001e420, 0x24126ea1)),
]


def philox4x32_10(c0, c1, c2, c3, k0, k1):
    """Vectorised over numpy uint32 arrays (broadcast)."""
    c0, c1, c2, c3 = (np.asarray(v, dtype=np.uint32) for v in (c0, c1, c2, c3))
    c0, c1, c2, c3 = np.broadcast_arrays(c0, c1, c2, c3)
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


def uniform(r):
    return ((r >> np.uint32(8)).astype(np.float32) + np.float32(0.5)) * np.float32(2.0 ** -24)


def normals(seed, stream_id, clip, draw, L):
    """float32 [L]: the standard normals aid_philox_normal adds for one clip."""
    nq = (L + 3) // 4
    q = np.arange(nq, dtype=np.uint32)
    r0, r1, r2, r3 = philox4x32_10(q, np.uint32(draw), np.uint32(clip), np.uint32(stream_id), seed & 0xFFFFFFFF, (seed >> 32) & 0xFFFFFFFF)
    out = np.empty((nq, 4), dtype=np.float32)
    for j, (a, b) in enumerate(((r0, r1), (r2, r3))):
        rad = np.sqrt(np.float32(-2.0) * np.log(uniform(a).astype(np.float64))).astype(np.float32)
        ang = 2.0 * np.pi * uniform(b).astype(np.float64)
        out[:, 2 * j] = (rad * np.cos(ang)).astype(np.float32)
        out[:, 2 * j + 1] = (rad * np.sin(ang)).astype(np.float32)
    return out.reshape(-1)[:L]


def batch_normals(seed, stream_id, clip0, n_clips, draw, L):
    return np.stack([normals(seed, stream_id, clip0 + c, draw, L) for c in range(n_clips)]) if n_clips else np.zeros((0, L), np.float32)


class PhiloxNoise:
    """Iterator with the Sampler.noise_source protocol: draw d of clips [lo, hi) -> torch [hi-lo, L] (prior first)."""

    def __init__(self, seed, stream_id, lo, hi, L):
        self.seed, self.stream_id, self.lo, self.hi, self.L, self.draw = seed, stream_id, lo, hi, L, 0

    def __iter__(self):
        return self

    def __next__(self):
        import torch
        n = batch_normals(self.seed, self.stream_id, self.lo, self.hi - self.lo, self.draw, self.L)
        self.draw += 1
        return torch.from_numpy(n)
