// Auxiliary kernels of the flight handles: reference-shaped map observation (de-tiling + TMA bulk stores), map
// export / import, episode-batch writer, live-step statistics.  See flight_common.cuh for the file map.
#include "flight_internal.h"

namespace csf {
namespace {

// ------------------------------------------------------------------------------------------------
// Reference-shaped observation of the flight variant: out[e][a] = prob_map[e].ravel() || (x^, y^, cos, sin)
// (flight_env.py:223-230), i.e. every map is read once and written n times.  The map lives in HBM as 4x4-cell tiles
// (flight_common.cuh); a persistent CTA de-tiles one env at a time into a row-major image in shared memory
// (coalesced 16-byte loads, two stages) and one elected thread then writes the image to the n observation rows with
// the TMA engine: n x cp.async.bulk.global.shared::cta [SASS: UBLKCP], one bulk group per env; a stage is refilled
// once the bulk group that read it has drained (cp.async.bulk.wait_group.read).  The bulk form needs
// (M*M*4) % 16 == 0; otherwise (or with obs_path = 1, the A/B hook) the image is written with plain stores.
// ------------------------------------------------------------------------------------------------
constexpr int kObsThreads = 256;

__global__ void __launch_bounds__(kObsThreads) flight_obs_full_kernel(const float* __restrict__ map, const float* __restrict__ obs,
                                                                      float* __restrict__ out, int E, int n, int M, int tiles,
                                                                      int map_stride, int stage_floats, int bulk) {
    extern __shared__ __align__(128) float img[];
    const int tid = threadIdx.x, MM = M * M;
    const size_t row_floats = (size_t)MM + 4;
    int it = 0;
    for (int e = blockIdx.x; e < E; e += gridDim.x, ++it) {
        float* st = img + (size_t)(it & 1) * stage_floats;
        if (bulk) {
            // the bulk group issued two iterations ago read this stage: it must have drained before the stage is rewritten
            if (tid == 0) asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory");
        }
        __syncthreads();
        const float4* src = reinterpret_cast<const float4*>(map + (size_t)e * map_stride);
        for (int f = tid; f < map_stride / 4; f += kObsThreads) {
            const int tile = f >> 2, r = f & 3;
            const int tr = tile / tiles, tc = tile - tr * tiles;
            const int i = 4 * tr + r, j0 = 4 * tc;
            if (i >= M) continue;
            const float4 v = __ldcs(src + f);
            const int d = i * M + j0;
            if (j0 + 4 <= M && (d & 3) == 0) {
                *reinterpret_cast<float4*>(st + d) = v;
            } else if (j0 + 4 <= M && (d & 1) == 0) {
                *reinterpret_cast<float2*>(st + d) = make_float2(v.x, v.y);
                *reinterpret_cast<float2*>(st + d + 2) = make_float2(v.z, v.w);
            } else {
                st[d] = v.x;
                if (j0 + 1 < M) st[d + 1] = v.y;
                if (j0 + 2 < M) st[d + 2] = v.z;
                if (j0 + 3 < M) st[d + 3] = v.w;
            }
        }
        // feature tails of this env: thread a writes the 4 floats after the map of row (e, a)
        for (int a = tid; a < n; a += kObsThreads) {
            const float4 f = reinterpret_cast<const float4*>(obs)[(size_t)e * n + a];
            float* dst = out + ((size_t)e * n + a) * row_floats + MM;
            dst[0] = f.x; dst[1] = f.y; dst[2] = f.z; dst[3] = f.w;
        }
        if (bulk) asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // the image, before the TMA engine reads it
        __syncthreads();
        if (bulk) {
            if (tid == 0) {
                const uint32_t s32 = smem_u32(st);
                for (int a = 0; a < n; ++a)
                    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;"
                                 ::"l"(out + ((size_t)e * n + a) * row_floats), "r"(s32), "r"((uint32_t)MM * 4u) : "memory");
                asm volatile("cp.async.bulk.commit_group;" ::: "memory");
            }
        } else {
            for (int idx = tid; idx < n * MM; idx += kObsThreads) {
                const int a = idx / MM, k = idx - a * MM;
                out[((size_t)e * n + a) * row_floats + k] = st[k];
            }
        }
    }
    if (bulk && tid == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");     // all stores complete before exit
}

// prob_map as the reference holds it, [E][M][M] row-major (prob_map[i][j], i <-> x), from / into the tiled device layout
__global__ void __launch_bounds__(256) flight_map_export_kernel(const float* __restrict__ map, float* __restrict__ out, int E, int M,
                                                                int tiles, int map_stride) {
    const long long total = (long long)E * M * M;
    for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (long long)gridDim.x * blockDim.x) {
        const long long e = idx / (M * M);
        const int c = (int)(idx - e * (M * M)), i = c / M, j = c - i * M;
        out[idx] = map[e * map_stride + tile_off(tiles, i, j)];
    }
}

__global__ void __launch_bounds__(256) flight_map_import_kernel(float* __restrict__ map, const float* __restrict__ in, int E, int M,
                                                                int tiles, int map_stride) {
    const long long total = (long long)E * M * M;
    for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (long long)gridDim.x * blockDim.x) {
        const long long e = idx / (M * M);
        const int c = (int)(idx - e * (M * M)), i = c / M, j = c - i * M;
        map[e * map_stride + tile_off(tiles, i, j)] = in[idx];
    }
}

// ------------------------------------------------------------------------------------------------
// Episode-batch writer: the padded 11-array layout RolloutWorker.generate_episode builds one env and one step at a
// time (common/rollout.py:43-132), for all envs of the handle, on the device.
//   begin : every array <- the padding of :105-116 (zeros, padded = terminated = 1); o[:,0], s[:,0] <- obs / state
//   record: after step t, for the envs that took it (time_step == t+1):  u, u_onehot, r, terminated, padded = 0,
//           avail_u[t] = avail_u_next[t] = 1, o_next[t] = s_next[t] = the new obs / state, and the same rows into
//           o[t+1], s[t+1] unless the episode ended (:79-97: "last obs" is only ever an *_next row)
// One thread per (env, output element); rows of different envs are contiguous, so the copies are coalesced.
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) flight_record_begin_kernel(const FlightParams p, cs_episode_buffers b, int T) {
    const int n = p.n, S = p.state_len, O = 4 * n, A = 3 * n;
    const int per_env = O + S;
    const long long total = (long long)p.E * per_env;
    for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (long long)gridDim.x * blockDim.x) {
        const int e = (int)(idx / per_env), k = (int)(idx - (long long)e * per_env);
        if (k < O) b.o[((size_t)e * T) * O + k] = p.obs[(size_t)e * O + k];
        else b.s[((size_t)e * T) * S + (k - O)] = p.state[(size_t)e * p.state_stride + (k - O)];
    }
    (void)A;
}

__global__ void __launch_bounds__(256) flight_record_kernel(const FlightParams p, cs_episode_buffers b, int t, int T,
                                                            const uint8_t* __restrict__ actions) {
    const int n = p.n, S = p.state_len, O = 4 * n, A = 3 * n;
    const int per_env = O + S + 1;
    const long long total = (long long)p.E * per_env;
    for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (long long)gridDim.x * blockDim.x) {
        const int e = (int)(idx / per_env), k = (int)(idx - (long long)e * per_env);
        uint4 mq0, mq1;
        meta_ld(p, e, &mq0, &mq1);
        const uint32_t mt[CS_META_WORDS] = {mq0.x, mq0.y, mq0.z, mq0.w, mq1.x, mq1.y, mq1.z, mq1.w};
        if (mt[CS_META_TIME] != (uint32_t)(t + 1)) continue;           // this env did not take step t (episode over)
        const bool over = (mt[CS_META_FLAGS] & CS_FLAG_DONE) != 0;
        const size_t row = (size_t)e * T + t;
        if (k < O) {
            const float v = p.obs[(size_t)e * O + k];
            b.o_next[row * O + k] = v;
            if (!over && t + 1 < T) b.o[(row + 1) * O + k] = v;
        } else if (k < O + S) {
            const float v = p.state[(size_t)e * p.state_stride + (k - O)];
            b.s_next[row * S + (k - O)] = v;
            if (!over && t + 1 < T) b.s[(row + 1) * S + (k - O)] = v;
        } else {
            for (int a = 0; a < n; ++a) {
                const uint8_t act = actions[(size_t)e * n + a];
                b.u[row * n + a] = act;
                for (int c = 0; c < 3; ++c) {
                    b.u_onehot[row * A + 3 * a + c] = (c == act) ? 1 : 0;
                    b.avail_u[row * A + 3 * a + c] = 1;                // get_avail_agent_actions: ones (flight_env_easy.py:184-188)
                    b.avail_u_next[row * A + 3 * a + c] = 1;
                }
            }
            b.r[row] = p.reward[e];
            b.terminated[row] = p.terminated[e];
            b.padded[row] = 0;
            if (over || t + 1 == T) {                                   // episode summary (rollout.py:64,79,137-140)
                b.episode_reward[e] = __uint_as_float(mt[CS_META_EPREWARD]);
                b.win_tag[e] = (over && (mt[CS_META_FLAGS] & CS_FLAG_WIN)) ? 1 : 0;
                b.targets_find[e] = p.target_find[e];
                b.length[e] = t + 1;
            }
        }
    }
}

// sum of the time_step words of all envs (steps of the episodes still running), for cs_flight_stats
__global__ void __launch_bounds__(256) flight_live_steps_kernel(const double* __restrict__ dyn, int E, long long rs, long long es, int meta_off,
                                                                double* __restrict__ out) {
    unsigned long long acc = 0;
    for (int e = blockIdx.x * blockDim.x + threadIdx.x; e < E; e += gridDim.x * blockDim.x) {
        const uint2 w23 = *reinterpret_cast<const uint2*>(dyn + (size_t)(meta_off + 1) * rs + (size_t)e * es);     // out mask, time_step
        const uint2 w45 = *reinterpret_cast<const uint2*>(dyn + (size_t)(meta_off + 2) * rs + (size_t)e * es);     // episode, flags
        if (!(w45.y & CS_FLAG_DONE)) acc += w23.y;                          // a finished episode is already in CS_STAT_EP_LEN
    }
    for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
    if ((threadIdx.x & 31) == 0 && acc) atomicAdd(out, (double)acc);
}

// ------------------------------------------------------------------------------------------------
// Compact step results for the host-buffer path.  What changes in an env's reference-shaped outputs every step is small:
// reward / target_find / terminated / win, the n agent rows (x^, y^, cos, sin) and a few find flags; the 2m target
// coordinates of a state row change only when the env is reset.  One record per env:
//   { float reward; uint32 found; uint8 target_find, terminated, win, reset; uint32 0 }   (16 bytes; with rec_bytes > 16
//   followed by the n agent rows -- the host path has the copy engine scatter those straight from `obs`, flight_hostio.cu)
// and, for the envs that were reset inside this call (auto-reset), one entry { int32 env; float xy[2m] } in a small
// side region claimed with an atomic counter.  The host expands this into full rows (flight_hostio.cu).
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128) flight_pack_kernel(const FlightParams p, unsigned char* __restrict__ out, int rec_bytes,
                                                          unsigned char* __restrict__ entries, int ent_bytes, int cap,
                                                          unsigned int* __restrict__ counter) {
    const int e = blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= p.E) return;
    uint4 m0, m1;
    meta_ld(p, e, &m0, &m1);
    const uint32_t term = p.terminated[e], reset = (term && p.auto_reset) ? 1u : 0u;
    uint4* rec = reinterpret_cast<uint4*>(out + (size_t)e * rec_bytes);
    const uint32_t bytes = (uint32_t)p.target_find[e] | (term << 8) | ((uint32_t)p.win[e] << 16) | (reset << 24);
    rec[0] = make_uint4(__float_as_uint(p.reward[e]), m0.x, bytes, 0u);
    if (rec_bytes > 16) {
        const uint4* row = reinterpret_cast<const uint4*>(p.state + (size_t)e * p.state_stride);
        for (int a = 0; a < p.n; ++a) rec[1 + a] = row[a];
    }
    if (reset) {
        const unsigned slot = atomicAdd(counter, 1u);
        if (slot < (unsigned)cap) {
            int* ent = reinterpret_cast<int*>(entries + (size_t)slot * ent_bytes);
            ent[0] = e;
            const float* tr = p.state + (size_t)e * p.state_stride + 4 * p.n;
            float* xy = reinterpret_cast<float*>(ent + 1);
            for (int j = 0; j < p.m; ++j) { xy[2 * j] = tr[3 * j]; xy[2 * j + 1] = tr[3 * j + 1]; }
        }
    }
}

// The same for all batches of a host pool in ONE launch (blockIdx.y = batch), as dense arrays at the batch's env offset:
// reward / target_find / terminated / win / found mask land in the pool's result block exactly as the host reads them
// (its views are the pinned mirror of that block), the agent rows in one contiguous array for the strided copy; reset
// entries carry the pooled env index.
struct PoolOut { float* reward; int32_t* target_find; uint8_t* terminated; uint8_t* win; uint32_t* found; uint4* agent; };

__global__ void __launch_bounds__(128) flight_pack_pool_kernel(const __grid_constant__ FlightParams common, const __grid_constant__ GroupTable tab,
                                                               const __grid_constant__ PoolGeom geo, const PoolOut out,
                                                               unsigned char* __restrict__ entries, int ent_bytes, int cap,
                                                               unsigned int* __restrict__ counter) {
    const GroupEntry& g = tab.h[blockIdx.y];
    const int e = blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= g.E) return;
    const int ge = geo.first[blockIdx.y] + e;
    const uint32_t term = g.terminated[e];
    out.reward[ge] = g.reward[e];
    out.target_find[ge] = g.target_find[e];
    out.terminated[ge] = (uint8_t)term;
    out.win[ge] = g.win[e];
    out.found[ge] = reinterpret_cast<const uint2*>(g.dyn + (size_t)common.meta_off * g.dyn_rs + (size_t)e * g.dyn_es)->x;
    const uint4* orow = reinterpret_cast<const uint4*>(g.obs) + (size_t)e * common.n;
    uint4* arow = out.agent + (size_t)ge * common.n;
    for (int a = 0; a < common.n; ++a) arow[a] = orow[a];
    if (term && common.auto_reset) {
        const unsigned slot = atomicAdd(counter, 1u);
        if (slot < (unsigned)cap) {
            int* ent = reinterpret_cast<int*>(entries + (size_t)slot * ent_bytes);
            ent[0] = ge;
            const float* tr = g.state + (size_t)e * common.state_stride + 4 * common.n;
            float* xy = reinterpret_cast<float*>(ent + 1);
            for (int j = 0; j < common.m; ++j) { xy[2 * j] = tr[3 * j]; xy[2 * j + 1] = tr[3 * j + 1]; }
        }
    }
}

inline int capped_grid(long long items, int per_block, int max_blocks) {
    const long long g = (items + per_block - 1) / per_block;
    return (int)(g < (long long)max_blocks ? (g < 1 ? 1 : g) : max_blocks);
}

}  // namespace

cudaError_t launch_obs_full(cs_flight* h, float* d_out, cudaStream_t st) {
    const FlightParams& p = h->p;
    const int MM = p.M * p.M;
    const int stage_floats = (MM + 31) & ~31;                         // stages start 128-byte aligned
    const size_t smem = 2 * (size_t)stage_floats * sizeof(float);
    const int bulk = (MM % 4 == 0 && h->obs_path != 1) ? 1 : 0;
    static size_t cur_limit = 48 * 1024;                              // per kernel, only ever raised
    if (smem > cur_limit) {
        const cudaError_t e = cudaFuncSetAttribute(flight_obs_full_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
        cur_limit = smem;
    }
    int per_sm = (int)((200 * 1024) / (smem + 1024));
    per_sm = per_sm > 4 ? 4 : (per_sm < 1 ? 1 : per_sm);
    const int grid = p.E < CS_NUM_SMS_B200 * per_sm ? p.E : CS_NUM_SMS_B200 * per_sm;
    flight_obs_full_kernel<<<grid, kObsThreads, smem, st>>>(p.prob_map, p.obs, d_out, p.E, p.n, p.M, p.tiles, p.map_stride, stage_floats, bulk);
    cs_count_launch(1);
    return cudaGetLastError();
}

cudaError_t launch_map_export(cs_flight* h, float* d_out, cudaStream_t st) {
    const FlightParams& p = h->p;
    flight_map_export_kernel<<<capped_grid((long long)p.E * p.M * p.M, 256, CS_NUM_SMS_B200 * 8), 256, 0, st>>>(p.prob_map, d_out, p.E, p.M, p.tiles, p.map_stride);
    cs_count_launch(1);
    return cudaGetLastError();
}

cudaError_t launch_map_import(cs_flight* h, const float* d_in, cudaStream_t st) {
    const FlightParams& p = h->p;
    flight_map_import_kernel<<<capped_grid((long long)p.E * p.M * p.M, 256, CS_NUM_SMS_B200 * 8), 256, 0, st>>>(p.prob_map, d_in, p.E, p.M, p.tiles, p.map_stride);
    cs_count_launch(1);
    return cudaGetLastError();
}

cudaError_t launch_record_begin(cs_flight* h, const cs_episode_buffers& b, int T, cudaStream_t st) {
    const FlightParams& p = h->p;
    flight_record_begin_kernel<<<capped_grid((long long)p.E * (4 * p.n + p.state_len), 256, CS_NUM_SMS_B200 * 8), 256, 0, st>>>(p, b, T);
    cs_count_launch(1);
    return cudaGetLastError();
}

cudaError_t launch_record(cs_flight* h, const cs_episode_buffers& b, int t, int T, const uint8_t* actions, cudaStream_t st) {
    const FlightParams& p = h->p;
    flight_record_kernel<<<capped_grid((long long)p.E * (4 * p.n + p.state_len + 1), 256, CS_NUM_SMS_B200 * 8), 256, 0, st>>>(p, b, t, T, actions);
    cs_count_launch(1);
    return cudaGetLastError();
}

cudaError_t launch_pack(cs_flight* h, cudaStream_t st) {
    cs_flight_compact* c = h->hc;
    cudaError_t e = cudaMemsetAsync(c->d_pack + c->off_counter, 0, 16, st);
    if (e != cudaSuccess) return e;
    flight_pack_kernel<<<(h->p.E + 127) / 128, 128, 0, st>>>(h->p, c->d_pack, (int)c->rec_bytes, c->d_pack + c->off_entries, (int)c->ent_bytes, c->cap,
                                                          reinterpret_cast<unsigned int*>(c->d_pack + c->off_counter));
    cs_count_launch(1);
    return cudaGetLastError();
}

cudaError_t launch_pack_pool(cs_flight_host_pool* pl, cudaStream_t st) {
    cudaError_t e = cudaMemsetAsync(pl->d_rec + pl->off_counter, 0, 16, st);
    if (e != cudaSuccess) return e;
    int max_E = 0;
    for (int i = 0; i < pl->count; ++i) max_E = pl->envs[i]->p.E > max_E ? pl->envs[i]->p.E : max_E;
    const dim3 grid((unsigned)((max_E + 127) / 128), (unsigned)pl->count);
    PoolOut out;
    out.reward = reinterpret_cast<float*>(pl->d_rec);
    out.target_find = reinterpret_cast<int32_t*>(pl->d_rec + pl->off_tf);
    out.terminated = pl->d_rec + pl->off_term;
    out.win = pl->d_rec + pl->off_win;
    out.found = reinterpret_cast<uint32_t*>(pl->d_rec + pl->off_found);
    out.agent = reinterpret_cast<uint4*>(pl->d_agent);
    flight_pack_pool_kernel<<<grid, 128, 0, st>>>(pl->envs[0]->p, pl->table, pl->geo, out,
                                                  pl->d_rec + pl->off_entries, (int)pl->ent_bytes, pl->cap,
                                                  reinterpret_cast<unsigned int*>(pl->d_rec + pl->off_counter));
    cs_count_launch(1);
    return cudaGetLastError();
}

cudaError_t launch_live_steps(cs_flight* h, cudaStream_t st) {
    flight_live_steps_kernel<<<capped_grid(h->p.E, 256, CS_NUM_SMS_B200 * 4), 256, 0, st>>>(h->p.dyn, h->p.E, h->p.dyn_rs, h->p.dyn_es, h->p.meta_off, h->d_live);
    cs_count_launch(1);
    return cudaGetLastError();
}

}  // namespace csf
