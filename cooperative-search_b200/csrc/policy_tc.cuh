// Agent network on the 5th-generation tensor cores (tcgen05 + TMEM), included by policy.cu.
//
// RNN.forward (network/base_net.py:30-47) for a tile of 128 (env, agent) rows is five small GEMMs with a shared M = 128:
//   fc1   [128 x 32] x [32 -> 64]                       -> +b, ReLU           -> x1     (input width <= 32, zero padded)
//   GRU   [128 x 64] x W_ih^T and [128 x 64] x W_hh^T   -> gates r | z | n    -> h'     (network/base_net.py:43, torch GRUCell)
//   fc2.0 [128 x 64] x [64 -> 64]                       -> +b, ReLU           -> f
//   fc2.2 [128 x 64] x [64 -> A (padded to 16)]         -> +b                 -> q
// One CTA of 256 threads owns a tile; threads t and t + 128 share row t -- each takes half of the columns of every
// epilogue (a warp reads the TMEM lanes 32 * (warp % 4) ..., so warps w and w + 4 see the same rows), which halves the
// dependent chain per tile and doubles the warps per SM that cover it (measured: 2.4e9 -> see DESIGN.md rows/s).  Operands are bf16 in shared memory in the canonical
// K-major, no-swizzle UMMA layout -- 16-byte chunks of 8 consecutive k, element (row, k) at
// chunk(k/8) * LBO + row * 16 + (k % 8) * 2 -- so that a thread writes its row's next operand with plain 16-byte
// stores; the weights (torch's (out, in) layout is already K-major) are packed into the same form by the host and
// stay resident in shared memory for the CTA's lifetime.  One elected thread issues the tcgen05.mma instructions
// (SASS: UTCHMMA), accumulators live in tensor memory (256 columns per CTA, two CTAs per SM), completion arrives on an
// mbarrier through tcgen05.commit, and every thread reads its row back with tcgen05.ld (SASS: LDTM) for the
// epilogue: bias, ReLU / sigmoid / tanh, the GRU blend, the bf16 operand of the next GEMM, and finally the masked
// argmax / epsilon-greedy / softmax-sampling action choice of agent/agent.py:66-97.  The r and z gates accumulate
// x1 W_ir^T + h W_hr^T in the same TMEM columns.  bf16 operands, fp32 accumulation, fp32 hidden state in HBM.
#pragma once
#include <cuda_bf16.h>

namespace cspol {

constexpr int kTcThreads = 256;                    // two threads per row
constexpr int kTcRows = 128;                       // rows per tile = UMMA M
constexpr int kTcK1 = 32;                          // fc1 input width, padded
constexpr int kTcNq = 16;                          // action-value width, padded (UMMA N is a multiple of 16)
constexpr int kTcTmemCols = 256;
constexpr int kALbo = (kTcRows + 1) * 16;          // bytes between K chunks of an A operand (129 rows: bank-conflict-free transposes)
// shared-memory map (bytes)
constexpr int kOffW1 = 0;                                   // [4 chunks][64][8] bf16
constexpr int kOffWih = kOffW1 + (kTcK1 / 8) * 64 * 16;     // [8][192][8]
constexpr int kOffWhh = kOffWih + 8 * 192 * 16;
constexpr int kOffW2a = kOffWhh + 8 * 192 * 16;             // [8][64][8]
constexpr int kOffW2b = kOffW2a + 8 * 64 * 16;              // [8][16][8]
constexpr int kOffBias = kOffW2b + 8 * kTcNq * 16;          // fp32: b1[64] | b_rz[128] | b_in[64] | b_hn[64] | b2a[64] | b2b[16]
constexpr int kBiasFloats = 64 + 128 + 64 + 64 + 64 + kTcNq;
constexpr int kTcWeightBytes = kOffBias + kBiasFloats * 4;  // what the host packs and every CTA copies
constexpr int kOffA0 = (kTcWeightBytes + 127) & ~127;       // [4][129 rows][8] bf16
constexpr int kOffAX = kOffA0 + (kTcK1 / 8) * kALbo;        // [8][129][8]: x1, later f
constexpr int kOffAH = kOffAX + 8 * kALbo;                  // [8][129][8]: h, later h'
constexpr int kOffBar = kOffAH + 8 * kALbo;                 // mbarrier (8 B) + TMEM base address (4 B)
constexpr int kTcSmemBytes = kOffBar + 16;

__device__ __forceinline__ uint32_t tc_smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

// K-major, no-swizzle shared-memory matrix descriptor: start address, leading-dimension byte offset (between the two
// 16-byte K chunks of one MMA), stride byte offset (between 8-row groups), descriptor version 1 (Blackwell)
__device__ __forceinline__ uint64_t tc_desc(uint32_t saddr, uint32_t lbo, uint32_t sbo) {
    return (uint64_t)((saddr & 0x3FFFFu) >> 4) | ((uint64_t)(lbo >> 4) << 16) | ((uint64_t)(sbo >> 4) << 32) | (1ull << 46);
}
// kind::f16 instruction descriptor: D fp32, A/B bf16, both K-major, M = 128
__device__ __forceinline__ constexpr uint32_t tc_idesc(int N) {
    return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(kTcRows >> 4) << 24);
}
__device__ __forceinline__ void tc_mma(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
                 ::"r"(d_tmem), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate) : "memory");
}
// D[128 x N] (+)= A[128 x 16*ksteps] * B[N x 16*ksteps]^T; B is rows n0.. of a weight block with `nrows` rows per chunk
__device__ __forceinline__ void tc_gemm(uint32_t d_tmem, uint32_t a_saddr, uint32_t b_saddr, int nrows, int n0, int N, int ksteps, bool accumulate) {
    const uint32_t idesc = tc_idesc(N);
    for (int k = 0; k < ksteps; ++k) {
        const uint64_t ad = tc_desc(a_saddr + (uint32_t)k * 2u * kALbo, kALbo, 128);
        const uint64_t bd = tc_desc(b_saddr + (uint32_t)n0 * 16u + (uint32_t)k * 2u * (uint32_t)nrows * 16u, (uint32_t)nrows * 16u, 128);
        tc_mma(d_tmem, ad, bd, idesc, (accumulate || k > 0) ? 1u : 0u);
    }
}
__device__ __forceinline__ void tc_commit(uint32_t bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void tc_wait(uint32_t bar, uint32_t parity) {
    uint32_t ok = 0, spins = 0;
    while (!ok) {
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                     : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
        if (!ok && ++spins > (1u << 26)) __trap();          // a lost completion must fail loudly, not hang the GPU
    }
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
}
// 8 / 16 consecutive fp32 columns of this thread's TMEM lane (= its row of the accumulator).  The loads are asynchronous:
// several are issued back to back and tc_ld_wait() orders them before the registers are read.
__device__ __forceinline__ void tc_ld8_issue(uint32_t taddr, float* v) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
                 : "=f"(v[0]), "=f"(v[1]), "=f"(v[2]), "=f"(v[3]), "=f"(v[4]), "=f"(v[5]), "=f"(v[6]), "=f"(v[7]) : "r"(taddr));
}
__device__ __forceinline__ void tc_ld16_issue(uint32_t taddr, float* v) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
                 : "=f"(v[0]), "=f"(v[1]), "=f"(v[2]), "=f"(v[3]), "=f"(v[4]), "=f"(v[5]), "=f"(v[6]), "=f"(v[7]),
                   "=f"(v[8]), "=f"(v[9]), "=f"(v[10]), "=f"(v[11]), "=f"(v[12]), "=f"(v[13]), "=f"(v[14]), "=f"(v[15]) : "r"(taddr));
}
__device__ __forceinline__ void tc_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
// ties 8 loaded registers to a point after tc_ld_wait(): an empty volatile asm that "rewrites" them, so that no consumer can be
// scheduled ahead of the wait (volatile asm statements keep their order)
__device__ __forceinline__ void tc_ld_tie8(float* v) {
    asm volatile("" : "+f"(v[0]), "+f"(v[1]), "+f"(v[2]), "+f"(v[3]), "+f"(v[4]), "+f"(v[5]), "+f"(v[6]), "+f"(v[7]));
}
__device__ __forceinline__ float tc_tanh(float x) {
    float y;
    asm("tanh.approx.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}
__device__ __forceinline__ float tc_sigmoid(float x) { return fmaf(0.5f, tc_tanh(0.5f * x), 0.5f); }
__device__ __forceinline__ uint32_t tc_pack2(float a, float b) {
    const __nv_bfloat162 v = __floats2bfloat162_rn(a, b);
    return *reinterpret_cast<const uint32_t*>(&v);
}
__device__ __forceinline__ uint4 tc_pack8(const float* v) {
    return make_uint4(tc_pack2(v[0], v[1]), tc_pack2(v[2], v[3]), tc_pack2(v[4], v[5]), tc_pack2(v[6], v[7]));
}
// the A operands just written with ordinary stores become visible to the tensor core (async proxy), TMEM reads are ordered
// before the next MMA, and the CTA meets
__device__ __forceinline__ void tc_operands_ready() {
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
}

template <class P>
__global__ void __launch_bounds__(kTcThreads, 2) policy_tc_kernel(const __grid_constant__ P p) {
    extern __shared__ __align__(128) unsigned char sm[];
    const int tid = threadIdx.x, warp = tid >> 5;
    const int row = tid & (kTcRows - 1), half = tid >> 7;             // this thread's row of the tile and its half of the columns
    const uint32_t sbase = tc_smem_u32(sm);
    const uint32_t bar = sbase + kOffBar;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(sm + kOffBar + 8);
    const float* bias = reinterpret_cast<const float*>(sm + kOffBias);
    const float *b1 = bias, *b_rz = bias + 64, *b_in = bias + 192, *b_hn = bias + 256, *b2a = bias + 320, *b2b = bias + 384;
    const int A = p.n_actions;

    // weights (bf16, canonical layout, packed by the host) + biases -> shared memory, once per CTA
    {
        const uint4* src = reinterpret_cast<const uint4*>(p.wtc);
        uint4* dst = reinterpret_cast<uint4*>(sm);
        for (int i = tid; i < kTcWeightBytes / 16; i += kTcThreads) dst[i] = __ldg(src + i);
    }
    if (tid == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(bar));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tc_smem_u32(tmem_slot)), "r"(kTcTmemCols) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tc_operands_ready();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem = *tmem_slot;
    const uint32_t trow = tmem + ((uint32_t)((warp & 3) * 32) << 16);  // this warp's 32 TMEM lanes
    uint32_t phase = 0;

    unsigned char* myA0 = sm + kOffA0 + row * 16;
    unsigned char* myAX = sm + kOffAX + row * 16;
    unsigned char* myAH = sm + kOffAH + row * 16;

    for (int tile = blockIdx.x; tile * kTcRows < p.rows; tile += gridDim.x) {
        const int r_raw = tile * kTcRows + row;
        const bool live = r_raw < p.rows;
        const int r = live ? r_raw : p.rows - 1;
        // ---- inputs: [conv features of the env's map ||] obs || last-action one-hot || agent-id one-hot
        //      (network/base_net.py:31-41, agent/agent.py:38-50)
        {
            const int la = (p.use_last && p.last_action) ? (int)p.last_action[r] : 255;
#pragma unroll
            for (int cc = 0; cc < kTcK1 / 16; ++cc) {
                const int c = (kTcK1 / 16) * half + cc;
                float x[8];
#pragma unroll
                for (int i = 0; i < 8; ++i) x[i] = policy_input(p, r, la, 8 * c + i);
                *reinterpret_cast<uint4*>(myA0 + c * kALbo) = tc_pack8(x);
            }
        }
        tc_operands_ready();
        // ---- fc1
        if (tid == 0) {
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            tc_gemm(tmem + 0, sbase + kOffA0, sbase + kOffW1, 64, 0, 64, kTcK1 / 16, false);
            tc_commit(bar);
        }
        // ... and while the tensor core works on fc1: this thread's half of the row's hidden state (units 32 * half ...), read
        // once (fp32, 8 loads in flight), kept in registers for the GRU blend, bf16 copy -> A operand of the GRU GEMMs
        float hreg[32];
        {
            const float4* hp = reinterpret_cast<const float4*>(p.hidden + (size_t)r * 64 + 32 * half);
#pragma unroll
            for (int c = 0; c < 8; ++c) {
                const float4 h4 = hp[c];
                hreg[4 * c] = h4.x; hreg[4 * c + 1] = h4.y; hreg[4 * c + 2] = h4.z; hreg[4 * c + 3] = h4.w;
            }
#pragma unroll
            for (int c = 0; c < 4; ++c) *reinterpret_cast<uint4*>(myAH + (4 * half + c) * kALbo) = tc_pack8(hreg + 8 * c);
            // the same half row of this CTA's NEXT tile: pull it into L2 now
            const long long rn = (long long)(tile + (int)gridDim.x) * kTcRows + row;
            if (rn < p.rows) {
                asm volatile("prefetch.global.L2 [%0];" ::"l"(p.hidden + (size_t)rn * 64 + 32 * half));
                if (half == 0) asm volatile("prefetch.global.L2 [%0];" ::"l"(p.obs + (size_t)rn * p.obs_dim));
            }
        }
        tc_wait(bar, phase); phase ^= 1u;
        {
            const int c = half;                                        // 32 of the 64 columns each
            float v[32];
            tc_ld16_issue(trow + 32 * c, v);
            tc_ld16_issue(trow + 32 * c + 16, v + 16);
            tc_ld_wait();
            tc_ld_tie8(v); tc_ld_tie8(v + 8); tc_ld_tie8(v + 16); tc_ld_tie8(v + 24);
#pragma unroll
            for (int i = 0; i < 32; ++i) v[i] = fmaxf(v[i] + b1[32 * c + i], 0.f);
#pragma unroll
            for (int q = 0; q < 4; ++q) *reinterpret_cast<uint4*>(myAX + (4 * c + q) * kALbo) = tc_pack8(v + 8 * q);
        }
        tc_operands_ready();
        // ---- GRU gates: r | z accumulate x1 W_i^T + h W_h^T in columns 0..127; n keeps its two halves apart
        if (tid == 0) {
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            tc_gemm(tmem + 0, sbase + kOffAX, sbase + kOffWih, 192, 0, 128, 4, false);
            tc_gemm(tmem + 0, sbase + kOffAH, sbase + kOffWhh, 192, 0, 128, 4, true);
            tc_gemm(tmem + 128, sbase + kOffAX, sbase + kOffWih, 192, 128, 64, 4, false);
            tc_gemm(tmem + 192, sbase + kOffAH, sbase + kOffWhh, 192, 128, 64, 4, false);
            tc_commit(bar);
        }
        tc_wait(bar, phase); phase ^= 1u;
        {
            float4* ho = reinterpret_cast<float4*>(p.hidden + (size_t)r * 64 + 32 * half);
#pragma unroll
            for (int cc = 0; cc < 4; ++cc) {
                const int c = 4 * half + cc;                           // hidden units 8c .. 8c+7
                float gr[8], gz[8], gi[8], gh[8];
                tc_ld8_issue(trow + 8 * c, gr);
                tc_ld8_issue(trow + 64 + 8 * c, gz);
                tc_ld8_issue(trow + 128 + 8 * c, gi);
                tc_ld8_issue(trow + 192 + 8 * c, gh);
                tc_ld_wait();
                tc_ld_tie8(gr); tc_ld_tie8(gz); tc_ld_tie8(gi); tc_ld_tie8(gh);
#pragma unroll
                for (int i = 0; i < 8; ++i) {
                    const int j = 8 * c + i, jl = 8 * cc + i;
                    const float rg = tc_sigmoid(gr[i] + b_rz[j]);
                    const float zg = tc_sigmoid(gz[i] + b_rz[64 + j]);
                    const float ng = tc_tanh(gi[i] + b_in[j] + rg * (gh[i] + b_hn[j]));
                    hreg[jl] = (1.f - zg) * ng + zg * hreg[jl];
                }
                if (live) {
                    ho[2 * cc] = make_float4(hreg[8 * cc], hreg[8 * cc + 1], hreg[8 * cc + 2], hreg[8 * cc + 3]);
                    ho[2 * cc + 1] = make_float4(hreg[8 * cc + 4], hreg[8 * cc + 5], hreg[8 * cc + 6], hreg[8 * cc + 7]);
                }
                *reinterpret_cast<uint4*>(myAH + c * kALbo) = tc_pack8(hreg + 8 * cc);
            }
        }
        tc_operands_ready();
        // ---- fc2.0
        if (tid == 0) {
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            tc_gemm(tmem + 0, sbase + kOffAH, sbase + kOffW2a, 64, 0, 64, 4, false);
            tc_commit(bar);
        }
        tc_wait(bar, phase); phase ^= 1u;
        {
            const int c = half;                                        // 32 of the 64 columns each
            float v[32];
            tc_ld16_issue(trow + 32 * c, v);
            tc_ld16_issue(trow + 32 * c + 16, v + 16);
            tc_ld_wait();
            tc_ld_tie8(v); tc_ld_tie8(v + 8); tc_ld_tie8(v + 16); tc_ld_tie8(v + 24);
#pragma unroll
            for (int i = 0; i < 32; ++i) v[i] = fmaxf(v[i] + b2a[32 * c + i], 0.f);
#pragma unroll
            for (int q = 0; q < 4; ++q) *reinterpret_cast<uint4*>(myAX + (4 * c + q) * kALbo) = tc_pack8(v + 8 * q);
        }
        tc_operands_ready();
        // ---- fc2.2 -> action values
        if (tid == 0) {
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            tc_gemm(tmem + 64, sbase + kOffAX, sbase + kOffW2b, kTcNq, 0, kTcNq, 4, false);
            tc_commit(bar);
        }
        tc_wait(bar, phase); phase ^= 1u;
        if (half == 0) {                                               // (warp-uniform: warps 0..3)
            float qv[8];
            tc_ld8_issue(trow + 64, qv);
            tc_ld_wait();
            tc_ld_tie8(qv);
#pragma unroll
            for (int a = 0; a < 8; ++a) qv[a] += b2b[a];
            if (live) policy_choose_action(p, r, qv);
        }
        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");     // the q read, before the next tile's first MMA
        __syncthreads();
    }
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(kTcTmemCols) : "memory");
}

}  // namespace cspol
