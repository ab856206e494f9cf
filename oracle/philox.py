"""Oracle (test infrastructure): Philox4x32-10 counter-based generator + Box-Muller, in numpy.

The reference draws DP noise with ``torch.normal`` (``src/shared/privacy.py:212``), i.e. whatever
generator torch has on the device; it fixes no stream layout, so the CUDA sampler is validated
for (a) bit-exact Philox4x32-10 integer output against this restatement -- itself pinned by the
published Random123 known-answer vectors (Salmon et al., SC'11; ``kat_vectors``) -- and (b) the
distribution of the resulting normals.

Stream layout used by the CUDA kernels (csrc/philox.cuh), restated here:
  counter = (block_lo, block_hi, stream_lo, stream_hi), key = (seed_lo, seed_hi)
  block   = element_index // 4 ; the four 32-bit outputs feed elements 4*block .. 4*block+3
  u       = (x + 0.5) * 2^-32  in (0, 1)   [computed as fma(x, 2^-32, 2^-33) in fp32, clamped below 1]
  z0, z1  = sqrt(-2 ln u0) * (cos 2 pi u1, sin 2 pi u1) ; z2, z3 likewise from (u2, u3)
"""
from __future__ import annotations

import numpy as np

M0 = np.uint64(0xD2511F53)
M1 = np.uint64(0xCD9E8D57)
W0 = np.uint32(0x9E3779B9)
W1 = np.uint32(0xBB67AE85)
MASK = np.uint64(0xFFFFFFFF)


def philox4x32_10(ctr: np.ndarray, key: np.ndarray) -> np.ndarray:
    """ctr: [..., 4] uint32, key: [..., 2] uint32 (broadcastable) -> [..., 4] uint32."""
    c = np.array(ctr, dtype=np.uint32, copy=True)
    c0, c1, c2, c3 = (c[..., i].astype(np.uint64) for i in range(4))
    k = np.array(key, dtype=np.uint32, copy=True)
    k0 = np.broadcast_to(k[..., 0], c0.shape).astype(np.uint32)
    k1 = np.broadcast_to(k[..., 1], c0.shape).astype(np.uint32)
    with np.errstate(over="ignore"):
        for _ in range(10):
            p0 = M0 * c0
            p1 = M1 * c2
            hi0, lo0 = p0 >> np.uint64(32), p0 & MASK
            hi1, lo1 = p1 >> np.uint64(32), p1 & MASK
            n0 = hi1 ^ c1 ^ k0.astype(np.uint64)
            n2 = hi0 ^ c3 ^ k1.astype(np.uint64)
            c0, c1, c2, c3 = n0, lo1, n2, lo0
            k0 = (k0 + W0).astype(np.uint32)
            k1 = (k1 + W1).astype(np.uint32)
    return np.stack([c0, c1, c2, c3], axis=-1).astype(np.uint32)


def raw_blocks(n_blocks: int, seed: int, stream: int, first_block: int = 0) -> np.ndarray:
    """uint32 [n_blocks, 4] for blocks first_block .. first_block + n_blocks - 1."""
    b = np.arange(first_block, first_block + n_blocks, dtype=np.uint64)
    ctr = np.stack([(b & MASK), (b >> np.uint64(32)),
                    np.full_like(b, stream & 0xFFFFFFFF), np.full_like(b, (stream >> 32) & 0xFFFFFFFF)],
                   axis=-1).astype(np.uint32)
    key = np.array([seed & 0xFFFFFFFF, (seed >> 32) & 0xFFFFFFFF], dtype=np.uint32)
    return philox4x32_10(ctr, key)


def uniform01(x: np.ndarray) -> np.ndarray:
    u = x.astype(np.float32) * np.float32(2.0 ** -32) + np.float32(2.0 ** -33)
    return np.minimum(u, np.float32(1.0 - 2.0 ** -24))


def normals(n: int, seed: int, stream: int) -> np.ndarray:
    """fp32 [n] standard normals for element indices 0..n-1 of (seed, stream)."""
    nb = (n + 3) // 4
    u = uniform01(raw_blocks(nb, seed, stream)).astype(np.float64)
    r0 = np.sqrt(-2.0 * np.log(u[:, 0]))
    r1 = np.sqrt(-2.0 * np.log(u[:, 2]))
    t0 = 2.0 * np.pi * u[:, 1]
    t1 = 2.0 * np.pi * u[:, 3]
    z = np.stack([r0 * np.cos(t0), r0 * np.sin(t0), r1 * np.cos(t1), r1 * np.sin(t1)], axis=-1)
    return z.reshape(-1)[:n].astype(np.float32)


# Random123 known-answer vectors for philox4x32-10: (counter, key, expected)
KAT = [
    ((0x00000000, 0x00000000, 0x00000000, 0x00000000), (0x00000000, 0x00000000),
     (0x6627E8D5, 0xE169C58D, 0xBC57AC4C, 0x9B00DBD8)),
    ((0xFFFFFFFF, 0xFFFFFFFF, 0xFFFFFFFF, 0xFFFFFFFF), (0xFFFFFFFF, 0xFFFFFFFF),
     (0x408F276D, 0x41C83B0E, 0xA20BC7C6, 0x6D5451FD)),
    ((0x243F6A88, 0x85A308D3, 0x13198A2E, 0x03707344), (0xA4093822, 0x299F31D0),
     (0xD16CFE09, 0x94FDCCEB, 0x5001E420, 0x24126EA1)),
]
