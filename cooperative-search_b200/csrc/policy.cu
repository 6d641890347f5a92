// Batched agent network + action selection (first version: CUDA cores, one warp per 4 rows).
//
// What is restated here (reference: WZN1ng/Cooperative-Search):
//   RNN.forward without the conv front end   network/base_net.py:30-47   fc1 -> ReLU -> GRUCell(64) -> Linear -> ReLU -> Linear
//   input assembly                           agent/agent.py:38-50        obs || last-action one-hot || agent-id one-hot
//   action choice                            agent/agent.py:66-75        q[avail == 0] = -inf; argmax, or a uniform available action with prob. epsilon
// The reference evaluates one (1, in) row per agent per step with a host<->device round trip each; here every (env, agent)
// row of a step is one warp-iteration of a persistent kernel whose CTAs keep the ~120 KB of weights in shared memory.
// fp32 with explicit FMAs, k ascending; sigmoid / tanh from expf / tanhf (no fast-math).
#include <math.h>
#include <new>
#include <vector>
#include "cs_common.cuh"
#include "cs_philox.cuh"

namespace {

constexpr int kH = 64;                 // args.rnn_hidden_dim (common/arguments.py)
constexpr int kPolicyThreads = 512;          // 16 warps per SM at 128 registers: one CTA per SM holds the weights
constexpr int kMaxIn = 32, kMaxActions = 8;
constexpr int kRowsPerWarp = 4;

struct PolicyParams {
    int rows, n_agents, obs_dim, n_actions, in_dim, use_last, use_id, evaluate;
    float epsilon;
    uint32_t seed, t;
    const float* w;              // packed weights, see pack order in cs_policy_create
    const float* obs;            // [rows][obs_dim]
    const uint8_t* last_action;  // [rows] or null (255 = none yet: zero one-hot)
    const uint8_t* avail;        // [rows][n_actions] or null (= all available)
    float* hidden;               // [rows][64], in/out
    float* q;                    // [rows][n_actions] or null
    uint8_t* actions;            // [rows] out (may alias last_action)
};

__device__ __forceinline__ float sigmoidf_(float x) { return 1.0f / (1.0f + expf(-x)); }

__global__ void __launch_bounds__(kPolicyThreads, 1) policy_kernel(const __grid_constant__ PolicyParams p) {
    extern __shared__ __align__(16) float wsm[];
    constexpr unsigned FULL = 0xffffffffu;
    const int in_dim = p.in_dim, A = p.n_actions;
    // packed layout (floats): W1T [in][64] | b1 [64] | WihT [64][192] | bih [192] | WhhT [64][192] | bhh [192] | W2T [64][64] | b2 [64] | W3 [A][64] | b3 [A]
    const int total = in_dim * kH + kH + 2 * (kH * 3 * kH + 3 * kH) + kH * kH + kH + A * kH + A;
    for (int i = threadIdx.x; i < total; i += kPolicyThreads) wsm[i] = p.w[i];
    __syncthreads();
    const float* W1T = wsm;
    const float* b1 = W1T + in_dim * kH;
    const float* WihT = b1 + kH;
    const float* bih = WihT + kH * 3 * kH;
    const float* WhhT = bih + 3 * kH;
    const float* bhh = WhhT + kH * 3 * kH;
    const float* W2T = bhh + 3 * kH;
    const float* b2 = W2T + kH * kH;
    const float* W3 = b2 + kH;
    const float* b3 = W3 + A * kH;

    // RPW consecutive rows per warp iteration: every weight read from shared memory serves RPW rows (the matvecs are
    // bound by the 12 shared-memory loads per k, not by the FMAs)
    constexpr int RPW = kRowsPerWarp;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, wpc = kPolicyThreads / 32;
    for (int rb = (blockIdx.x * wpc + warp) * RPW; rb < p.rows; rb += gridDim.x * wpc * RPW) {
        int row[RPW];
        float inv[RPW], x0[RPW], x1[RPW], h0[RPW], h1[RPW];
#pragma unroll
        for (int u = 0; u < RPW; ++u) {
            const int r = min(rb + u, p.rows - 1);                      // a short last block repeats its last row (writes are guarded)
            row[u] = r;
            const int a_id = r % p.n_agents;
            // input element `lane` (agent/agent.py:38-50)
            float v = 0.f;
            if (lane < p.obs_dim) v = p.obs[(size_t)r * p.obs_dim + lane];
            else if (p.use_last && lane < p.obs_dim + A) v = (p.last_action && p.last_action[r] == lane - p.obs_dim) ? 1.f : 0.f;
            else if (p.use_id && lane < in_dim) v = (a_id == lane - p.obs_dim - (p.use_last ? A : 0)) ? 1.f : 0.f;
            inv[u] = v;
            x0[u] = b1[lane]; x1[u] = b1[lane + 32];
            h0[u] = p.hidden[(size_t)r * kH + lane]; h1[u] = p.hidden[(size_t)r * kH + lane + 32];
        }
        __syncwarp();                                                    // all rows' last actions are read before any is rewritten
        // fc1 + ReLU: outputs j = lane, lane + 32
        for (int k = 0; k < in_dim; ++k) {
            const float w0 = W1T[k * kH + lane], w1 = W1T[k * kH + lane + 32];
#pragma unroll
            for (int u = 0; u < RPW; ++u) {
                const float v = __shfl_sync(FULL, inv[u], k);
                x0[u] = fmaf(v, w0, x0[u]);
                x1[u] = fmaf(v, w1, x1[u]);
            }
        }
#pragma unroll
        for (int u = 0; u < RPW; ++u) { x0[u] = fmaxf(x0[u], 0.f); x1[u] = fmaxf(x1[u], 0.f); }
        // GRUCell: gates r | z | n, outputs j = lane + 32 * o
        float gi[RPW][6], gh[RPW][6];
#pragma unroll
        for (int u = 0; u < RPW; ++u)
#pragma unroll
            for (int o = 0; o < 6; ++o) { gi[u][o] = bih[lane + 32 * o]; gh[u][o] = bhh[lane + 32 * o]; }
        for (int k = 0; k < kH; ++k) {
            float wi[6], wh[6];
#pragma unroll
            for (int o = 0; o < 6; ++o) { wi[o] = WihT[k * 3 * kH + lane + 32 * o]; wh[o] = WhhT[k * 3 * kH + lane + 32 * o]; }
#pragma unroll
            for (int u = 0; u < RPW; ++u) {
                const float xk = __shfl_sync(FULL, k < 32 ? x0[u] : x1[u], k & 31);
                const float hk = __shfl_sync(FULL, k < 32 ? h0[u] : h1[u], k & 31);
#pragma unroll
                for (int o = 0; o < 6; ++o) {
                    gi[u][o] = fmaf(xk, wi[o], gi[u][o]);
                    gh[u][o] = fmaf(hk, wh[o], gh[u][o]);
                }
            }
        }
        float hn0[RPW], hn1[RPW], y0[RPW], y1[RPW];
#pragma unroll
        for (int u = 0; u < RPW; ++u) {
            const float r0 = sigmoidf_(gi[u][0] + gh[u][0]), r1 = sigmoidf_(gi[u][1] + gh[u][1]);
            const float z0 = sigmoidf_(gi[u][2] + gh[u][2]), z1 = sigmoidf_(gi[u][3] + gh[u][3]);
            const float n0 = tanhf(gi[u][4] + r0 * gh[u][4]), n1 = tanhf(gi[u][5] + r1 * gh[u][5]);
            hn0[u] = (1.f - z0) * n0 + z0 * h0[u];
            hn1[u] = (1.f - z1) * n1 + z1 * h1[u];
            if (rb + u < p.rows) {
                p.hidden[(size_t)row[u] * kH + lane] = hn0[u];
                p.hidden[(size_t)row[u] * kH + lane + 32] = hn1[u];
            }
            y0[u] = b2[lane]; y1[u] = b2[lane + 32];
        }
        // fc2: Linear + ReLU, then Linear to the action values
        for (int k = 0; k < kH; ++k) {
            const float w0 = W2T[k * kH + lane], w1 = W2T[k * kH + lane + 32];
#pragma unroll
            for (int u = 0; u < RPW; ++u) {
                const float hk = __shfl_sync(FULL, k < 32 ? hn0[u] : hn1[u], k & 31);
                y0[u] = fmaf(hk, w0, y0[u]);
                y1[u] = fmaf(hk, w1, y1[u]);
            }
        }
#pragma unroll
        for (int u = 0; u < RPW; ++u) {
            const float ya = fmaxf(y0[u], 0.f), yb = fmaxf(y1[u], 0.f);
            const int r = row[u];
            float qv[kMaxActions];
#pragma unroll
            for (int a = 0; a < kMaxActions; ++a) {
                if (a >= A) { qv[a] = 0.f; continue; }
                float part = ya * W3[a * kH + lane] + yb * W3[a * kH + lane + 32];
                for (int o = 16; o > 0; o >>= 1) part += __shfl_xor_sync(FULL, part, o);
                qv[a] = part + b3[a];
            }
            if (lane == 0 && rb + u < p.rows) {
                int best = -1, navail = 0;
                float bq = -INFINITY;
#pragma unroll
                for (int a = 0; a < kMaxActions; ++a) {
                    if (a >= A) continue;
                    if (p.q) p.q[(size_t)r * A + a] = qv[a];
                    const bool ok = !p.avail || p.avail[(size_t)r * A + a] != 0;
                    navail += ok ? 1 : 0;
                    if (ok && (best < 0 || qv[a] > bq)) { best = a; bq = qv[a]; }     // first maximum, like torch.argmax
                }
                int act = best < 0 ? 0 : best;
                if (!p.evaluate && p.epsilon > 0.f && navail > 0) {
                    // agent.py:71-74: np.random.rand() >= epsilon -> argmax, else a uniform available action
                    const cs_u4 w = cs_philox4x32_10((uint32_t)r, p.t, 0u, 0u, p.seed, CS_STREAM_POLICY);
                    const float uu = (float)(w.x >> 8) * (1.0f / 16777216.0f);
                    if (uu < p.epsilon) {
                        int pick = (int)(w.y % (uint32_t)navail);
#pragma unroll
                        for (int a = 0; a < kMaxActions; ++a) {
                            if (a >= A) continue;
                            const bool ok = !p.avail || p.avail[(size_t)r * A + a] != 0;
                            if (ok && pick-- == 0) act = a;
                        }
                    }
                }
                p.actions[r] = (uint8_t)act;
            }
        }
    }
}

}  // namespace

struct cs_policy {
    int device, obs_dim, n_actions, n_agents, use_last, use_id, in_dim;
    size_t w_floats;
    float* d_w;
};

extern "C" {

int cs_policy_create(const cs_policy_cfg* cfg, const cs_policy_weights* hw, cs_policy** out) {
    CS_REQUIRE(cfg && hw && out, "cs_policy_create: null argument");
    CS_REQUIRE(cfg->struct_size == sizeof(cs_policy_cfg), "cs_policy_create: cfg.struct_size mismatch");
    CS_REQUIRE(cfg->hidden_dim == kH, "cs_policy_create: rnn_hidden_dim must be %d", kH);
    CS_REQUIRE(cfg->n_actions >= 1 && cfg->n_actions <= kMaxActions, "cs_policy_create: n_actions must be in 1..%d", kMaxActions);
    const int in_dim = cfg->obs_dim + (cfg->last_action ? cfg->n_actions : 0) + (cfg->reuse_network ? cfg->n_agents : 0);
    CS_REQUIRE(cfg->obs_dim >= 1 && in_dim <= kMaxIn, "cs_policy_create: input width %d exceeds %d (the conv front end is not part of this kernel)", in_dim, kMaxIn);
    CS_REQUIRE(hw->fc1_w && hw->fc1_b && hw->w_ih && hw->w_hh && hw->b_ih && hw->b_hh && hw->fc2a_w && hw->fc2a_b && hw->fc2b_w && hw->fc2b_b,
               "cs_policy_create: null weight pointer");
    const int A = cfg->n_actions;
    std::vector<float> pk;
    pk.reserve((size_t)in_dim * kH + kH + 2 * (kH * 3 * kH + 3 * kH) + kH * kH + kH + A * kH + A);
    // torch layouts: Linear.weight (out, in); GRUCell.weight_ih / weight_hh (3*hidden, hidden), gates r | z | n
    for (int k = 0; k < in_dim; ++k) for (int j = 0; j < kH; ++j) pk.push_back(hw->fc1_w[(size_t)j * in_dim + k]);
    for (int j = 0; j < kH; ++j) pk.push_back(hw->fc1_b[j]);
    for (int k = 0; k < kH; ++k) for (int j = 0; j < 3 * kH; ++j) pk.push_back(hw->w_ih[(size_t)j * kH + k]);
    for (int j = 0; j < 3 * kH; ++j) pk.push_back(hw->b_ih[j]);
    for (int k = 0; k < kH; ++k) for (int j = 0; j < 3 * kH; ++j) pk.push_back(hw->w_hh[(size_t)j * kH + k]);
    for (int j = 0; j < 3 * kH; ++j) pk.push_back(hw->b_hh[j]);
    for (int k = 0; k < kH; ++k) for (int j = 0; j < kH; ++j) pk.push_back(hw->fc2a_w[(size_t)j * kH + k]);
    for (int j = 0; j < kH; ++j) pk.push_back(hw->fc2a_b[j]);
    for (int a = 0; a < A; ++a) for (int j = 0; j < kH; ++j) pk.push_back(hw->fc2b_w[(size_t)a * kH + j]);
    for (int a = 0; a < A; ++a) pk.push_back(hw->fc2b_b[a]);

    cs_policy* h = new (std::nothrow) cs_policy();
    if (!h) { cs_set_error("out of host memory"); return CS_ERR_NOMEM; }
    h->device = cfg->device; h->obs_dim = cfg->obs_dim; h->n_actions = A; h->n_agents = cfg->n_agents;
    h->use_last = cfg->last_action; h->use_id = cfg->reuse_network; h->in_dim = in_dim; h->w_floats = pk.size(); h->d_w = nullptr;
    cudaError_t e = cudaSetDevice(cfg->device);
    if (e == cudaSuccess) e = cudaMalloc(&h->d_w, pk.size() * sizeof(float));
    if (e == cudaSuccess) e = cudaMemcpy(h->d_w, pk.data(), pk.size() * sizeof(float), cudaMemcpyHostToDevice);
    // the largest weight block any policy can have, so that handles of different widths coexist
    constexpr size_t kMaxWeights = (size_t)kMaxIn * kH + kH + 2 * (kH * 3 * kH + 3 * kH) + kH * kH + kH + kMaxActions * kH + kMaxActions;
    if (e == cudaSuccess) e = cudaFuncSetAttribute(policy_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(kMaxWeights * sizeof(float)));
    if (e != cudaSuccess) { cudaFree(h->d_w); delete h; }
    CS_CUDA(e);
    *out = h;
    return CS_OK;
}

void cs_policy_destroy(cs_policy* h) {
    if (!h) return;
    cudaSetDevice(h->device);
    cudaFree(h->d_w);
    delete h;
}

int cs_policy_act(cs_policy* h, const cs_policy_io* io, void* stream) {
    CS_REQUIRE(h && io && io->obs && io->hidden && io->actions && io->rows >= 0, "cs_policy_act: bad argument");
    if (io->rows == 0) return CS_OK;
    PolicyParams p;
    p.rows = io->rows; p.n_agents = h->n_agents; p.obs_dim = h->obs_dim; p.n_actions = h->n_actions; p.in_dim = h->in_dim;
    p.use_last = h->use_last; p.use_id = h->use_id; p.evaluate = io->evaluate; p.epsilon = io->epsilon; p.seed = io->seed; p.t = io->t;
    p.w = h->d_w; p.obs = io->obs; p.last_action = io->last_action; p.avail = io->avail; p.hidden = io->hidden; p.q = io->q;
    p.actions = io->actions;
    const int wpc = (kPolicyThreads / 32) * kRowsPerWarp;
    int grid = (io->rows + wpc - 1) / wpc;
    if (grid > CS_NUM_SMS_B200) grid = CS_NUM_SMS_B200;                   // persistent: one CTA per SM keeps the weights resident
    policy_kernel<<<grid, kPolicyThreads, h->w_floats * sizeof(float), (cudaStream_t)stream>>>(p);
    cs_count_launch(1);
    CS_CUDA(cudaGetLastError());
    return CS_OK;
}

}  // extern "C"
