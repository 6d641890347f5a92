"""Philox4x32-10 keyed RNG -- TEST INFRASTRUCTURE (oracle side).

This is the stateless counter-based generator that replaces the reference's
global ``np.random`` stream for the detection draw
(``env/flight_env_easy.py:238``, ``env/flight_env.py:248``) so that the draw for
(env, episode, t, agent i, target j) is the same number in the reference, the
oracle and the CUDA kernels.  Algorithm: Salmon et al., "Parallel random
numbers: as easy as 1, 2, 3" (SC'11), Philox-4x32 with 10 rounds; known-answer
vectors from the Random123 distribution are checked in tests/test_philox.py.

Nothing in the product package imports this file.
"""
import numpy as np

M0 = 0xD2511F53
M1 = 0xCD9E8D57
W0 = 0x9E3779B9
W1 = 0xBB67AE85
MASK = 0xFFFFFFFF

# stream tags (key word 1); the CUDA side uses the same values (csrc/philox.cuh)
STREAM_DETECT = 1     # detection uniforms
STREAM_TARGET = 2     # on-device target randomisation at reset
STREAM_POLICY = 3     # in-kernel uniform-random policy (step_k)
STREAM_SEARCH = 4     # search_env on-device target placement
STREAM_SPREAD = 5     # simple_spread on-device reset placement


def philox4x32(c0, c1, c2, c3, k0, k1, rounds=10):
    """Scalar Philox4x32 on Python ints. Returns 4 uint32 words."""
    c0 &= MASK; c1 &= MASK; c2 &= MASK; c3 &= MASK
    k0 &= MASK; k1 &= MASK
    for _ in range(rounds):
        p0 = M0 * c0
        p1 = M1 * c2
        c0, c1, c2, c3 = ((p1 >> 32) ^ c1 ^ k0) & MASK, p1 & MASK, \
                         ((p0 >> 32) ^ c3 ^ k1) & MASK, p0 & MASK
        k0 = (k0 + W0) & MASK
        k1 = (k1 + W1) & MASK
    return c0, c1, c2, c3


def philox4x32_np(c0, c1, c2, c3, k0, k1, rounds=10):
    """Vectorised Philox4x32 on uint64-held uint32 arrays (broadcasting)."""
    c0, c1, c2, c3 = [np.asarray(c, dtype=np.uint64) & np.uint64(MASK) for c in (c0, c1, c2, c3)]
    c0, c1, c2, c3 = np.broadcast_arrays(c0, c1, c2, c3)
    k0 = np.uint64(k0 & MASK)
    k1 = np.uint64(k1 & MASK)
    m = np.uint64(MASK)
    s32 = np.uint64(32)
    for _ in range(rounds):
        p0 = np.uint64(M0) * c0
        p1 = np.uint64(M1) * c2
        c0, c1, c2, c3 = ((p1 >> s32) ^ c1 ^ k0) & m, p1 & m, ((p0 >> s32) ^ c3 ^ k1) & m, p0 & m
        k0 = (k0 + np.uint64(W0)) & m
        k1 = (k1 + np.uint64(W1)) & m
    return c0, c1, c2, c3


def stream_key(stream, episode):
    """Key word 1 of a stream: the stream tag with the episode's bits 16..31 above it (csrc/cs_philox.cuh: cs_stream_key)."""
    return (stream | (((episode & MASK) >> 16) << 8)) & MASK


def detect_counter(env_id, episode, t, i, j):
    """Counter layout shared with the CUDA kernels (csrc/philox.cuh: cs_detect_words).
    One Philox block serves agents 4*(i>>2) .. 4*(i>>2)+3 for target j."""
    return env_id & MASK, ((episode & 0xFFFF) << 16) | (t & 0xFFFF), i >> 2, j


def detect_draw(seed, env_id, episode, t, i, j):
    """uint32 detection draw for pair (agent i, target j) of `env_id` at `_update_obs` call `t`.
    t = 0 is the call made inside reset(); step k (1-based) uses t = k."""
    c = detect_counter(env_id, episode, t, i, j)
    return philox4x32(c[0], c[1], c[2], c[3], seed, stream_key(STREAM_DETECT, episode))[i & 3]


def detect_threshold(detect_prob):
    """Largest uint32 r with r / 2**32 <= detect_prob; -1 if none (int64 domain)."""
    import math
    if detect_prob >= 1.0:
        return MASK
    if detect_prob < 0.0:
        return -1
    return int(math.floor(detect_prob * 4294967296.0))


def u53(hi, lo):
    """(0,1) double from two uint32 words: 53 random bits + half-ulp offset."""
    return (float((hi >> 5) * 67108864 + (lo >> 6)) + 0.5) / 9007199254740992.0
