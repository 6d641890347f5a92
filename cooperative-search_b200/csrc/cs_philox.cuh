// Philox4x32-10 counter-based RNG (Salmon et al., SC'11) -- device + host.
// Same constants/stream tags as oracle/philox.py; KATs are checked on the GPU in
// tests/test_gpu_philox.py through cs_debug_philox().
#pragma once
#include <stdint.h>

#if defined(__CUDACC__)
#define CS_HD __host__ __device__ __forceinline__
#else
#define CS_HD inline
#endif

#define CS_STREAM_DETECT 1u
#define CS_STREAM_TARGET 2u
#define CS_STREAM_POLICY 3u
#define CS_STREAM_SEARCH 4u
#define CS_STREAM_SPREAD 5u

struct cs_u4 { uint32_t x, y, z, w; };

CS_HD cs_u4 cs_philox4x32_10(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, uint32_t k0, uint32_t k1) {
    const uint32_t M0 = 0xD2511F53u, M1 = 0xCD9E8D57u, W0 = 0x9E3779B9u, W1 = 0xBB67AE85u;
#pragma unroll
    for (int r = 0; r < 10; ++r) {
        const uint64_t p0 = (uint64_t)M0 * c0;
        const uint64_t p1 = (uint64_t)M1 * c2;
        const uint32_t n0 = (uint32_t)(p1 >> 32) ^ c1 ^ k0;
        const uint32_t n2 = (uint32_t)(p0 >> 32) ^ c3 ^ k1;
        c1 = (uint32_t)p1;
        c3 = (uint32_t)p0;
        c0 = n0;
        c2 = n2;
        k0 += W0;
        k1 += W1;
    }
    cs_u4 o; o.x = c0; o.y = c1; o.z = c2; o.w = c3;
    return o;
}

// Key word 1 of a stream: the stream tag, with the episode's bits 16..31 above it -- the counter word holds
// (episode & 0xFFFF) << 16 | t, so together every episode of an env has its own streams (no repeat after 65536 episodes).
CS_HD uint32_t cs_stream_key(uint32_t stream, uint32_t episode) { return stream | ((episode >> 16) << 8); }

// Words serving agents 4*blk..4*blk+3 for target j at sensing call t of (env, episode).
CS_HD cs_u4 cs_detect_words(uint32_t seed, uint32_t env_id, uint32_t episode, uint32_t t, uint32_t blk, uint32_t j) {
    return cs_philox4x32_10(env_id, ((episode & 0xFFFFu) << 16) | (t & 0xFFFFu), blk, j, seed, cs_stream_key(CS_STREAM_DETECT, episode));
}

CS_HD uint32_t cs_word(const cs_u4& w, int k) { return k == 0 ? w.x : (k == 1 ? w.y : (k == 2 ? w.z : w.w)); }

// (0,1) double from two words: 53 random bits + half-ulp offset (oracle/philox.py: u53)
CS_HD double cs_u53(uint32_t hi, uint32_t lo) {
    const uint64_t bits = ((uint64_t)(hi >> 5) << 26) + (uint64_t)(lo >> 6);
    return ((double)bits + 0.5) / 9007199254740992.0;
}
