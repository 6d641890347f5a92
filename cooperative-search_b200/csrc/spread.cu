// simple_spread (env/simple_spread.py of WZN1ng/Cooperative-Search), batched for B200: SURVEY.md section 8f rank 4.
//   reset (uniform placement of targets, then agents)   env/simple_spread.py:48-70
//   _agent_step (unit moves, clamped to the map)         :130-139
//   _update_obs / get_obs / get_state                    :78-128
//   reward (minus the distance of every target to its nearest agent, summed in target order) :141-167
//   step                                                 :169-180
// A group of G lanes per env (G = 8 where n_agents and target_num are <= 8: four envs per warp; else the whole warp).  The
// env's 2(n + m) float64 coordinates sit in a per-group shared-memory slice; lanes 0..n-1 of the group move their
// agent, every lane then writes observation / state elements (element q of the env's contiguous n * obs_shape block is a
// difference of two coordinates picked by index arithmetic: fully coalesced 128-byte stores), one lane per target finds its
// nearest agent and lane 0 adds the m distances in target order (the reference's summation order).  Coordinates and the
// reward are float64 like the reference's Python floats; observations and states leave as float32.  The file is compiled
// with -fmad=false.  There is no randomness in a step; reset placement comes from the keyed Philox stream
// (cs_philox.cuh, stream SPREAD) or is injected (CS_RESET_KEEP_TARGETS = keep every coordinate the caller wrote).
#include <new>
#include "cs_common.cuh"
#include "cs_philox.cuh"

namespace {

constexpr int kThreads = 128;
constexpr int kMaxEnt = CS_MAX_AGENTS + CS_MAX_TARGETS;

struct SpreadParams {
    int E, n, m, T, auto_reset;
    int obs_dim, state_dim;
    double Md, half_M;
    uint32_t seed, env_id_base;
    double* pos;            // [E][n + m][2]: agents, then targets
    int32_t* meta;          // [E][4]: time_step, episode, done flag, 0
    double* ep_reward;      // [E] total_reward of the running episode (simple_spread.py:175)
    float* obs;             // [E][n][obs_dim]
    float* state;           // [E][state_dim]
    float* reward;          // [E]
    double* reward64;       // [E] the same in the reference's precision
    uint8_t* terminated;    // [E]
    uint8_t* occupied;      // [E][m]   (:96-110; the reference only renders it)
    double* stats;          // [CS_NUM_STATS]
};

enum { MODE_STEP = 0, MODE_RESET = 1 };

template <int G>
__device__ __forceinline__ void place(const SpreadParams& p, double* xy, int e, uint32_t episode, int lane) {
    // entity k < m is target k, the rest are the agents, in the order the reference draws them (:57-68)
    for (int k = lane; k < p.n + p.m; k += G) {
        const cs_u4 w = cs_philox4x32_10(p.env_id_base + (uint32_t)e, (episode & 0xFFFFu) << 16, (uint32_t)k, 0u, p.seed,
                                         cs_stream_key(CS_STREAM_SPREAD, episode));
        const int slot = k < p.m ? p.n + k : k - p.m;
        xy[2 * slot] = p.Md * cs_u53(w.x, w.y);
        xy[2 * slot + 1] = p.Md * cs_u53(w.z, w.w);
    }
}

// slot s of agent i's observation row (element i * obs_dim + s of the env's observation block), simple_spread.py:82-101
__device__ __forceinline__ float obs_element(const SpreadParams& p, const double* xy, int i, int s) {
    const int n = p.n, m = p.m;
    const int k = s & 1;
    const double own = xy[2 * i + k];
    if (s < 2) return (float)(own - p.half_M);                                     // own position
    if (s < 2 + 2 * m) return (float)(xy[2 * (n + ((s - 2) >> 1)) + k] - own);       // targets, relative
    if (s < 2 + 2 * m + 2 * (n - 1)) {                                               // the other agents, relative
        int j = (s - 2 - 2 * m) >> 1;
        j += (j >= i);
        return (float)(xy[2 * j + k] - own);
    }
    return (float)(xy[2 * (n + ((s - 2 - 2 * m - 2 * (n - 1)) >> 1)) + k] - p.half_M);   // targets, absolute
}

template <int MODE, int G>
__global__ void __launch_bounds__(kThreads) spread_kernel(const SpreadParams p, const uint8_t* __restrict__ actions, const uint8_t* __restrict__ mask,
                                                         uint32_t rflags) {
    constexpr int kGroups = kThreads / G, kEnt = (G == 8) ? 16 : kMaxEnt, kTgt = (G == 8) ? 8 : CS_MAX_TARGETS;
    __shared__ double sxy[kGroups][2 * kEnt];
    __shared__ double smin[kGroups][kTgt];
    const int lane = threadIdx.x % G, w = threadIdx.x / G;
    const int e_raw = blockIdx.x * kGroups + w;
    const bool active = e_raw < p.E;
    const int e = active ? e_raw : p.E - 1;                                          // (idle groups keep the warp's barriers company)
    double* xy = sxy[w];
    const int n = p.n, m = p.m, ne = n + m;
    double* gp = p.pos + (size_t)e * 2 * ne;
    for (int c = lane; c < 2 * ne; c += G) xy[c] = gp[c];
    const int4 mt = *reinterpret_cast<const int4*>(p.meta + 4 * (size_t)e);
    int time_step = mt.x, episode = mt.y, done = mt.z;
    double ep_reward = p.ep_reward[e];
    __syncwarp();

    bool emit = false, have_result = false, term = false;
    double rew = 0.0;
    if (MODE == MODE_STEP) {
        have_result = true;
        const bool stepping = !done;
        if (stepping && lane < n) {                                                  // _agent_step (:130-139)
            const int a = actions != nullptr ? (int)actions[(size_t)e * n + lane]
                                             : (int)(cs_word(cs_philox4x32_10(p.env_id_base + (uint32_t)e, (((uint32_t)episode & 0xFFFFu) << 16) | ((uint32_t)(time_step + 1) & 0xFFFFu),
                                                                              (uint32_t)(lane >> 2), 0u, p.seed, cs_stream_key(CS_STREAM_POLICY, (uint32_t)episode)), lane & 3) % 5u);
            const double dx = (a == 1) ? 1.0 : ((a == 3) ? -1.0 : 0.0), dy = (a == 2) ? 1.0 : ((a == 4) ? -1.0 : 0.0);
            xy[2 * lane] = fmin(fmax(0.0, xy[2 * lane] + dx), p.Md);
            xy[2 * lane + 1] = fmin(fmax(0.0, xy[2 * lane + 1] + dy), p.Md);
        }
        __syncwarp();
        if (stepping) {
            for (int j = lane; j < m; j += G) {                                      // nearest agent of target j (:144-153)
                const double tx = xy[2 * (n + j)], ty = xy[2 * (n + j) + 1];
                double best = 0.0;
                bool occ = false;
                for (int i = 0; i < n; ++i) {
                    const double dx = xy[2 * i] - tx, dy = xy[2 * i + 1] - ty;
                    const double d = sqrt(dx * dx + dy * dy);
                    best = (i == 0 || d < best) ? d : best;
                    occ |= d < 6.0;                                                  // agent_radius (:24)
                }
                smin[w][j] = best;
                if (active) p.occupied[(size_t)e * m + j] = occ ? 1 : 0;
            }
        }
        __syncwarp();
        if (stepping) {
            for (int j = 0; j < m; ++j) rew -= smin[w][j];                           // every lane: the same sum, in target order
            ep_reward += rew;
            time_step += 1;
            term = time_step >= p.T;                                                 // (:176-178)
            done = term ? 1 : 0;
            emit = true;
        } else {
            term = true;                                                             // masked no-op on a finished env
        }
    }
    const bool do_reset = (MODE == MODE_RESET) ? (mask == nullptr || mask[e] != 0) : (p.auto_reset && term && emit);
    if (MODE == MODE_STEP && term && emit && lane == 0 && active) {                  // statistics of the episode that ends here
        atomicAdd(p.stats + CS_STAT_EPISODES, 1.0);
        atomicAdd(p.stats + CS_STAT_EP_REWARD, ep_reward);
        atomicAdd(p.stats + CS_STAT_EP_LEN, (double)time_step);
    }
    if (do_reset) {                                                                  // reset (:48-70)
        episode += (rflags & CS_RESET_KEEP_EPISODE) ? 0 : 1;
        time_step = 0; done = 0; ep_reward = 0.0;
    }
    __syncwarp();
    if (do_reset && !(rflags & CS_RESET_KEEP_TARGETS)) place<G>(p, xy, e, (uint32_t)episode, lane);
    __syncwarp();
    if (do_reset) {
        if (MODE == MODE_RESET && active) {
            for (int j = lane; j < m; j += G) {
                bool occ = false;
                for (int i = 0; i < n; ++i) {
                    const double dx = xy[2 * (n + j)] - xy[2 * i], dy = xy[2 * (n + j) + 1] - xy[2 * i + 1];
                    occ |= sqrt(dx * dx + dy * dy) < 6.0;
                }
                p.occupied[(size_t)e * m + j] = occ ? 1 : 0;
            }
        }
        emit = true;
    }
    if (emit && active) {
        for (int c = lane; c < (do_reset ? 2 * ne : 2 * n); c += G) gp[c] = xy[c];   // targets only move at a reset
        float* ob = p.obs + (size_t)e * n * p.obs_dim;
        int oi = lane / p.obs_dim, os = lane - oi * p.obs_dim;                        // (agent, slot) of element q, kept incrementally
        for (int q = lane; q < n * p.obs_dim; q += G) {
            ob[q] = obs_element(p, xy, oi, os);
            os += G;
            while (os >= p.obs_dim) { os -= p.obs_dim; ++oi; }
        }
        float* st = p.state + (size_t)e * p.state_dim;
        for (int q = lane; q < 2 * ne; q += G) st[q] = (float)(xy[q] - p.half_M);    // agents, then targets (:116-128)
        if (lane == 0) {
            *reinterpret_cast<int4*>(p.meta + 4 * (size_t)e) = make_int4(time_step, episode, done, 0);
            p.ep_reward[e] = ep_reward;
        }
    }
    if (have_result && lane == 0 && active) {
        p.reward[e] = (float)rew;
        p.reward64[e] = rew;
        p.terminated[e] = term ? 1 : 0;
    }
}

}  // namespace

struct cs_spread {
    cs_spread_cfg cfg;
    SpreadParams p;
    uint8_t* d_actions;
};

extern "C" {

void cs_spread_destroy(cs_spread* h) {
    if (!h) return;
    cudaSetDevice(h->cfg.device);
    cudaFree(h->p.pos); cudaFree(h->p.meta); cudaFree(h->p.ep_reward); cudaFree(h->p.obs); cudaFree(h->p.state); cudaFree(h->p.reward);
    cudaFree(h->p.reward64); cudaFree(h->p.terminated); cudaFree(h->p.occupied); cudaFree(h->p.stats); cudaFree(h->d_actions);
    delete h;
}

int cs_spread_create(const cs_spread_cfg* cfg, cs_spread** out) {
    CS_REQUIRE(cfg && out, "cs_spread_create: null argument");
    CS_REQUIRE(cfg->struct_size == sizeof(cs_spread_cfg), "cs_spread_create: cfg.struct_size %u != %zu (ABI mismatch)", cfg->struct_size, sizeof(cs_spread_cfg));
    CS_REQUIRE(cfg->num_envs > 0, "num_envs must be > 0");
    CS_REQUIRE(cfg->n_agents >= 1 && cfg->n_agents <= CS_MAX_AGENTS, "n_agents must be in 1..%d", CS_MAX_AGENTS);
    CS_REQUIRE(cfg->target_num >= 1 && cfg->target_num <= CS_MAX_TARGETS, "target_num must be in 1..%d", CS_MAX_TARGETS);
    CS_REQUIRE(cfg->map_size >= 1, "map_size must be >= 1");
    CS_REQUIRE(cfg->time_limit >= 1 && cfg->time_limit <= 65535, "time_limit must be in 1..65535");
    cs_spread* h = new (std::nothrow) cs_spread();
    if (!h) { cs_set_error("out of host memory"); return CS_ERR_NOMEM; }
    memset(h, 0, sizeof(*h));
    h->cfg = *cfg;
    SpreadParams& p = h->p;
    const int n = cfg->n_agents, m = cfg->target_num;
    const size_t E = (size_t)cfg->num_envs;
    p.E = cfg->num_envs; p.n = n; p.m = m; p.T = cfg->time_limit; p.auto_reset = cfg->auto_reset;
    p.obs_dim = 2 + (n - 1) * 2 + m * 4;                       // simple_spread.py:28
    p.state_dim = 2 * n + 2 * m;                               // :27
    p.Md = (double)cfg->map_size; p.half_M = 0.5 * (double)cfg->map_size;
    p.seed = cfg->seed; p.env_id_base = cfg->env_id_base;
    cudaError_t e = cudaSetDevice(cfg->device);
    if (e == cudaSuccess) e = cudaMalloc(&p.pos, E * 2 * (n + m) * sizeof(double));
    if (e == cudaSuccess) e = cudaMemset(p.pos, 0, E * 2 * (n + m) * sizeof(double));
    if (e == cudaSuccess) e = cudaMalloc(&p.meta, E * 4 * sizeof(int32_t));
    if (e == cudaSuccess) e = cudaMemset(p.meta, 0, E * 4 * sizeof(int32_t));
    // episode counter starts at -1 so that the first reset opens episode 0
    if (e == cudaSuccess) e = cudaMemset2D(p.meta + 1, 4 * sizeof(int32_t), 0xFF, sizeof(int32_t), E);
    if (e == cudaSuccess) e = cudaMalloc(&p.ep_reward, E * sizeof(double));
    if (e == cudaSuccess) e = cudaMemset(p.ep_reward, 0, E * sizeof(double));
    if (e == cudaSuccess) e = cudaMalloc(&p.obs, E * n * p.obs_dim * sizeof(float));
    if (e == cudaSuccess) e = cudaMemset(p.obs, 0, E * n * p.obs_dim * sizeof(float));
    if (e == cudaSuccess) e = cudaMalloc(&p.state, E * p.state_dim * sizeof(float));
    if (e == cudaSuccess) e = cudaMemset(p.state, 0, E * p.state_dim * sizeof(float));
    if (e == cudaSuccess) e = cudaMalloc(&p.reward, E * sizeof(float));
    if (e == cudaSuccess) e = cudaMemset(p.reward, 0, E * sizeof(float));
    if (e == cudaSuccess) e = cudaMalloc(&p.reward64, E * sizeof(double));
    if (e == cudaSuccess) e = cudaMemset(p.reward64, 0, E * sizeof(double));
    if (e == cudaSuccess) e = cudaMalloc(&p.terminated, E);
    if (e == cudaSuccess) e = cudaMemset(p.terminated, 0, E);
    if (e == cudaSuccess) e = cudaMalloc(&p.occupied, E * m);
    if (e == cudaSuccess) e = cudaMemset(p.occupied, 0, E * m);
    if (e == cudaSuccess) e = cudaMalloc(&p.stats, CS_NUM_STATS * sizeof(double));
    if (e == cudaSuccess) e = cudaMemset(p.stats, 0, CS_NUM_STATS * sizeof(double));
    if (e == cudaSuccess) e = cudaMalloc(&h->d_actions, E * n);
    if (e != cudaSuccess) {
        cs_set_error("cs_spread_create: %s", cudaGetErrorString(e));
        cs_spread_destroy(h);
        cudaGetLastError();
        return CS_ERR_CUDA;
    }
    *out = h;
    return CS_OK;
}

int cs_spread_env_info(const cs_spread* h, int32_t* out4) {
    CS_REQUIRE(h && out4, "cs_spread_env_info: null argument");
    out4[0] = 5;                       // n_actions       (simple_spread.py:26)
    out4[1] = h->p.state_dim;          // state_shape     (:27)
    out4[2] = h->p.obs_dim;            // obs_shape       (:28)
    out4[3] = h->p.T;                  // episode_limit   (:25,45)
    return CS_OK;
}

int cs_spread_buffers_get(cs_spread* h, cs_spread_buffers* b) {
    CS_REQUIRE(h && b, "cs_spread_buffers_get: null argument");
    const SpreadParams& p = h->p;
    b->pos = p.pos; b->meta = p.meta; b->obs = p.obs; b->state = p.state; b->reward = p.reward; b->reward64 = p.reward64;
    b->terminated = p.terminated; b->occupied = p.occupied; b->stats = p.stats; b->episode_reward = p.ep_reward;
    b->obs_dim = p.obs_dim; b->state_dim = p.state_dim;
    return CS_OK;
}

static cudaError_t spread_launch(cs_spread* h, int mode, const uint8_t* actions, const uint8_t* mask, uint32_t rflags, cudaStream_t st) {
    const bool small = h->p.n <= 8 && h->p.m <= 8;                   // 8 lanes per env, else the whole warp
    const int per_cta = kThreads / (small ? 8 : 32);
    const int grid = (h->p.E + per_cta - 1) / per_cta;
    if (small) {
        if (mode == MODE_STEP) spread_kernel<MODE_STEP, 8><<<grid, kThreads, 0, st>>>(h->p, actions, mask, rflags);
        else spread_kernel<MODE_RESET, 8><<<grid, kThreads, 0, st>>>(h->p, actions, mask, rflags);
    } else if (mode == MODE_STEP) spread_kernel<MODE_STEP, 32><<<grid, kThreads, 0, st>>>(h->p, actions, mask, rflags);
    else spread_kernel<MODE_RESET, 32><<<grid, kThreads, 0, st>>>(h->p, actions, mask, rflags);
    cs_count_launch(1);
    return cudaGetLastError();
}

int cs_spread_reset(cs_spread* h, const uint8_t* d_mask, uint32_t flags, void* stream) {
    CS_REQUIRE(h, "cs_spread_reset: null handle");
    CS_CUDA(spread_launch(h, MODE_RESET, nullptr, d_mask, flags, (cudaStream_t)stream));
    return CS_OK;
}

int cs_spread_step(cs_spread* h, const uint8_t* d_actions, void* stream) {
    CS_REQUIRE(h && d_actions, "cs_spread_step: null argument");
    CS_CUDA(spread_launch(h, MODE_STEP, d_actions, nullptr, 0u, (cudaStream_t)stream));
    return CS_OK;
}

int cs_spread_step_random(cs_spread* h, int32_t k, void* stream) {
    CS_REQUIRE(h && k >= 0, "cs_spread_step_random: bad argument");
    for (int i = 0; i < k; ++i) CS_CUDA(spread_launch(h, MODE_STEP, nullptr, nullptr, 0u, (cudaStream_t)stream));
    return CS_OK;
}

int cs_spread_step_host(cs_spread* h, const uint8_t* h_actions, float* h_reward, uint8_t* h_terminated, float* h_obs, float* h_state, void* stream) {
    CS_REQUIRE(h && h_actions, "cs_spread_step_host: null argument");
    const SpreadParams& p = h->p;
    const size_t E = (size_t)p.E;
    cudaStream_t st = (cudaStream_t)stream;
    CS_CUDA(cudaSetDevice(h->cfg.device));
    CS_CUDA(cudaMemcpyAsync(h->d_actions, h_actions, E * p.n, cudaMemcpyHostToDevice, st));
    CS_CUDA(spread_launch(h, MODE_STEP, h->d_actions, nullptr, 0u, st));
    if (h_reward) CS_CUDA(cudaMemcpyAsync(h_reward, p.reward, E * sizeof(float), cudaMemcpyDeviceToHost, st));
    if (h_terminated) CS_CUDA(cudaMemcpyAsync(h_terminated, p.terminated, E, cudaMemcpyDeviceToHost, st));
    if (h_obs) CS_CUDA(cudaMemcpyAsync(h_obs, p.obs, E * p.n * p.obs_dim * sizeof(float), cudaMemcpyDeviceToHost, st));
    if (h_state) CS_CUDA(cudaMemcpyAsync(h_state, p.state, E * p.state_dim * sizeof(float), cudaMemcpyDeviceToHost, st));
    CS_CUDA(cudaStreamSynchronize(st));
    return CS_OK;
}

int cs_spread_stats(cs_spread* h, double* h_out, void* stream) {
    CS_REQUIRE(h && h_out, "cs_spread_stats: null argument");
    cudaStream_t st = (cudaStream_t)stream;
    CS_CUDA(cudaSetDevice(h->cfg.device));
    CS_CUDA(cudaMemcpyAsync(h_out, h->p.stats, CS_NUM_STATS * sizeof(double), cudaMemcpyDeviceToHost, st));
    CS_CUDA(cudaStreamSynchronize(st));
    return CS_OK;
}

}  // extern "C"
