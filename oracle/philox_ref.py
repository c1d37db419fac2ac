"""TEST INFRASTRUCTURE ONLY -- numpy restatement of the counter-based generator used by the
in-kernel Gaussian sketch (``parla_b200/csrc/philox.cuh``).

The reference draws its Gaussian operator with numpy's PCG64 + ziggurat
(parla/utils/sketching.py:20-31), which cannot be replayed on a GPU; the CUDA path therefore
defines its own *virtual* operator S(seed)[r, i] from Philox4x32-10 (Salmon et al., SC'11;
constants and round function as published, checked against the Random123 known-answer vectors
in tests/golden/philox_kat.npz).  This file states that definition on the CPU:

    q, lane   = i >> 2, i & 3                      (4 consecutive columns share one Philox block)
    o[0..3]   = philox4x32_10(ctr=(q & 0xffffffff, q >> 32, r, 0), key=(seed & 0xffffffff, seed >> 32))
    u_a, u_b  = (o[2p] + 0.5) * 2^-32, (o[2p+1] + 0.5) * 2^-32          for pair p = lane >> 1
    rad, ang  = sqrt(-2 ln u_a), pi * (2 u_b - 1)
    g         = rad * cos(ang) if lane is even else rad * sin(ang)       (Box-Muller)
    S[r, i]   = scale * g

The GPU evaluates ln/sin/cos with fp32 fast intrinsics, so device values agree with this
restatement to ~1e-6 absolute, not bit-for-bit; the integer stream is bit-exact.
"""
import numpy as np

M0 = np.uint64(0xD2511F53)
M1 = np.uint64(0xCD9E8D57)
W0 = 0x9E3779B9
W1 = 0xBB67AE85
MASK = np.uint64(0xFFFFFFFF)
SHIFT = np.uint64(32)


def philox4x32_10(ctr, key):
    """ctr: (N, 4) uint32, key: (2,) uint32 -> (N, 4) uint32."""
    c = np.asarray(ctr, dtype=np.uint32).astype(np.uint64)
    c0, c1, c2, c3 = (c[:, j].copy() for j in range(4))
    k0, k1 = int(key[0]), int(key[1])
    for _ in range(10):
        p0 = M0 * c0
        p1 = M1 * c2
        hi0, lo0 = p0 >> SHIFT, p0 & MASK
        hi1, lo1 = p1 >> SHIFT, p1 & MASK
        n0 = hi1 ^ c1 ^ np.uint64(k0)
        n2 = hi0 ^ c3 ^ np.uint64(k1)
        c0, c1, c2, c3 = n0, lo1, n2, lo0
        k0 = (k0 + W0) & 0xFFFFFFFF
        k1 = (k1 + W1) & 0xFFFFFFFF
    return np.stack([c0, c1, c2, c3], axis=1).astype(np.uint32)


def gaussian_block(seed, row0, n_rows, col0, n_cols, scale=1.0):
    """Rows [row0, row0+n_rows) x columns [col0, col0+n_cols) of the virtual operator S(seed)."""
    seed = int(seed) & 0xFFFFFFFFFFFFFFFF
    key = np.array([seed & 0xFFFFFFFF, seed >> 32], dtype=np.uint32)
    q_lo, q_hi = col0 >> 2, (col0 + n_cols - 1) >> 2
    qs = np.arange(q_lo, q_hi + 1, dtype=np.uint64)
    out = np.empty((n_rows, (q_hi - q_lo + 1) * 4))
    for rr in range(n_rows):
        ctr = np.zeros((qs.size, 4), dtype=np.uint32)
        ctr[:, 0] = (qs & MASK).astype(np.uint32)
        ctr[:, 1] = (qs >> SHIFT).astype(np.uint32)
        ctr[:, 2] = np.uint32(row0 + rr)
        o = philox4x32_10(ctr, key).astype(np.float64)
        u = (o + 0.5) * 2.0 ** -32
        rad_a = np.sqrt(-2.0 * np.log(u[:, 0]))
        rad_b = np.sqrt(-2.0 * np.log(u[:, 2]))
        ang_a = np.pi * (2.0 * u[:, 1] - 1.0)
        ang_b = np.pi * (2.0 * u[:, 3] - 1.0)
        g = np.stack([rad_a * np.cos(ang_a), rad_a * np.sin(ang_a),
                      rad_b * np.cos(ang_b), rad_b * np.sin(ang_b)], axis=1)
        out[rr] = g.reshape(-1)
    lo = col0 - 4 * q_lo
    return scale * out[:, lo:lo + n_cols]


def sjlt_columns(d, m, k, seed, col_offset=0):
    """Index form (rows[m,k] int32, signs[m,k] int8) of the native SJLT operator
    (``pla_sjlt_generate`` / parla_b200/csrc/sjlt.cu):

    for column gi = col_offset + i, Philox blocks with ctr = (gi_lo, gi_hi, call, 0x534A4C54) and
    key = (seed_lo, seed_hi), call = 0, 1, ...; the words are consumed in order.  Word 0 holds the
    sign bits (bit q set -> -1).  Row q is floor(word * d / 2^32), redrawn while it duplicates an
    earlier row of the same column (no redraw when d < k).  Integer arithmetic only: bit-exact.
    """
    seed = int(seed) & 0xFFFFFFFFFFFFFFFF
    key = np.array([seed & 0xFFFFFFFF, seed >> 32], dtype=np.uint32)
    rows = np.empty((m, k), dtype=np.int32)
    signs = np.empty((m, k), dtype=np.int8)
    gi = np.arange(col_offset, col_offset + m, dtype=np.uint64)
    ncalls = 24                                       # plenty of words for redraws even when d ~ k
    words = []
    for call in range(ncalls):
        ctr = np.zeros((m, 4), dtype=np.uint32)
        ctr[:, 0] = (gi & MASK).astype(np.uint32)
        ctr[:, 1] = (gi >> SHIFT).astype(np.uint32)
        ctr[:, 2] = call
        ctr[:, 3] = 0x534A4C54
        words.append(philox4x32_10(ctr, key))
    words = np.concatenate(words, axis=1).astype(np.uint64)     # (m, 4*ncalls)
    for i in range(m):
        wi = 1
        sbits = int(words[i, 0])
        picked = []
        for q in range(k):
            while True:
                if wi >= words.shape[1]:
                    raise RuntimeError("ran out of precomputed Philox words; raise ncalls")
                cand = int((int(words[i, wi]) * d) >> 32)
                wi += 1
                if d < k or cand not in picked:
                    break
            picked.append(cand)
            rows[i, q] = cand
            signs[i, q] = -1 if (sbits >> q) & 1 else 1
    return rows, signs
