"""numpy twin of the device counter RNG (TEST INFRASTRUCTURE).

Philox4x32-10 (Salmon et al., SC'11; Random123 constants).  The reference has no
counter RNG at all (it uses torch's global mt19937: BNN_BBB.py:25,
BNN_Dropout.py:74, util.py:45-50 via torch.rand); this twin pins the *device*
streams bit-for-bit: raw 32-bit words, the 24-bit uniforms, and the 0/1
Bernoulli keep masks are exact; normals and concrete masks go through
log/sincos/sigmoid and are compared within a few ulp (tolerance in the tests).

Counter layout (must match bnn_b200/csrc/philox.cuh):
    key     = (seed & 0xffffffff, seed >> 32)
    counter = (block & 0xffffffff, block >> 32, subsequence, stream_tag)
    element e of a stream lives in word (e & 3) of block (e >> 2).
"""
import numpy as np

M0, M1 = np.uint64(0xD2511F53), np.uint64(0xCD9E8D57)
W0, W1 = 0x9E3779B9, 0xBB67AE85
MASK32 = np.uint64(0xFFFFFFFF)

STREAM_REPARAM = 1   # eps for w = mu + sigma * eps
STREAM_BERNOULLI = 2  # MC-dropout keep masks
STREAM_CONCRETE = 3   # concrete-dropout uniforms
STREAM_SGLD = 4       # Langevin noise


def philox4x32_10(c0, c1, c2, c3, k0, k1):
    """Vectorised over counters; all inputs broadcastable uint32 arrays/ints."""
    c0, c1, c2, c3 = [np.asarray(v, dtype=np.uint64) & MASK32 for v in (c0, c1, c2, c3)]
    c0, c1, c2, c3 = np.broadcast_arrays(c0, c1, c2, c3)
    k0, k1 = int(k0) & 0xFFFFFFFF, int(k1) & 0xFFFFFFFF
    for _ in range(10):
        p0 = M0 * c0
        p1 = M1 * c2
        hi0, lo0 = p0 >> np.uint64(32), p0 & MASK32
        hi1, lo1 = p1 >> np.uint64(32), p1 & MASK32
        c0, c1, c2, c3 = (hi1 ^ c1 ^ np.uint64(k0)), lo1, (hi0 ^ c3 ^ np.uint64(k1)), lo0
        k0 = (k0 + W0) & 0xFFFFFFFF
        k1 = (k1 + W1) & 0xFFFFFFFF
    return [v.astype(np.uint32) for v in (c0, c1, c2, c3)]


def stream_words(seed, subsequence, tag, n):
    """First ``n`` 32-bit words of stream (seed, subsequence, tag)."""
    nblk = (n + 3) // 4
    blk = np.arange(nblk, dtype=np.uint64)
    w = philox4x32_10(blk & MASK32, blk >> np.uint64(32), subsequence, tag,
                      seed & 0xFFFFFFFF, (seed >> 32) & 0xFFFFFFFF)
    return np.stack(w, axis=1).reshape(-1)[:n]


def u24(words):
    """[0,1) uniforms with 24 random bits: (w >> 8) * 2**-24 (exact in float32)."""
    return ((words >> np.uint32(8)).astype(np.float32) * np.float32(2.0 ** -24)).astype(np.float32)


def stream_uniform(seed, subsequence, tag, n):
    return u24(stream_words(seed, subsequence, tag, n))


def stream_normal(seed, subsequence, tag, n):
    """Box-Muller, two normals per word pair: with m1 = w[2i] >> 8, m2 = w[2i+1] >> 8,
    r = sqrt(-2 ln((m1 + 1) * 2**-24)); z[2i] = r cos(2 pi m2 2**-24), z[2i+1] = r sin(..).
    Computed in float64 here and rounded; the device uses float32 logf/sincospif."""
    nw = ((n + 3) // 4) * 4
    w = stream_words(seed, subsequence, tag, nw).astype(np.uint64)
    m1 = (w[0::2] >> np.uint64(8)).astype(np.float64)
    m2 = (w[1::2] >> np.uint64(8)).astype(np.float64)
    r = np.sqrt(-2.0 * np.log((m1 + 1.0) * 2.0 ** -24))
    ang = 2.0 * np.pi * (m2 * 2.0 ** -24)
    z = np.empty(nw, dtype=np.float64)
    z[0::2] = r * np.cos(ang)
    z[1::2] = r * np.sin(ang)
    return z[:n].astype(np.float32)


def bernoulli_keep(seed, subsequence, n, p_drop):
    """0/1 keep mask: keep iff u24 < float32(1 - p)."""
    q = np.float32(1.0) - np.float32(p_drop)
    return (stream_uniform(seed, subsequence, STREAM_BERNOULLI, n) < q).astype(np.float32)
