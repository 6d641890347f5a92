// Compact host-buffer path of the flight envs: the call a CPU-side rollout makes every step
// (reward, terminated, info = env.step(actions); env.get_obs(); env.get_state(), common/rollout.py:45-63) moves
// 16 + 16n bytes per env over PCIe instead of the full reference-shaped state row (4(4n + 3m) + 10 bytes) and keeps the
// reference-shaped rows up to date in (pinned) host memory:
//   * the agent part of every row (= the get_obs rows, 16n bytes) is written IN PLACE by the copy engine: one strided
//     device-to-host copy (cudaMemcpy2DAsync, rows state_stride floats apart) -- measured on B200 / PCIe 5: 262144 rows of
//     48 bytes in 0.45 ms, where host threads need ~2 ms for the same scatter (one cache miss per row);
//   * a 16-byte record per env {reward, found mask, target_find, terminated, win, reset} comes back in one flat copy; host
//     threads spread it over the result arrays and flip the find flags whose bit changed;
//   * the 2m target coordinates are rewritten only for envs that were reset inside the call (a small side list).
// Results live in library-owned host arrays (cs_flight_host_views), valid until the next step.
#include <atomic>
#include <condition_variable>
#include <functional>
#include <mutex>
#include <new>
#include <thread>
#include <vector>
#include "flight_internal.h"

using namespace csf;

namespace {

// fork-join pool: run(fn) calls fn(thread index, thread count) on every worker and returns when all are done
class HostPool {
public:
    static HostPool& get() {
        // never destroyed: the workers are detached and wait on the condition variable for the life of the process
        // (destroying a condition variable with waiters blocks at exit)
        static HostPool* pool = new HostPool();
        return *pool;
    }
    int size() const { return (int)workers_.size() + 1; }
    // fn(thread index, thread count) on every worker; workers with index >= the count the caller wants simply return
    void run(const std::function<void(int, int)>& fn) {
        const int T = size();
        if (T == 1) { fn(0, 1); return; }
        std::lock_guard<std::mutex> one_at_a_time(run_m_);           // callers on several host threads take turns
        {
            std::lock_guard<std::mutex> lk(m_);
            fn_ = &fn;
            pending_ = T - 1;
            ++gen_;
        }
        cv_.notify_all();
        fn(0, T);
        std::unique_lock<std::mutex> lk(m_);
        done_.wait(lk, [&] { return pending_ == 0; });
    }

private:
    HostPool() {
        int t = 0;
        if (const char* s = getenv("CS_HOST_THREADS")) t = atoi(s);
        if (t <= 0) {
            int hw = (int)std::thread::hardware_concurrency();
            int local = 1;
            if (const char* s = getenv("LOCAL_WORLD_SIZE")) local = atoi(s) > 0 ? atoi(s) : 1;   // one process per GPU: share the cores
            t = hw / local;
            if (t > 16) t = 16;
        }
        if (t < 1) t = 1;
        for (int i = 1; i < t; ++i) workers_.emplace_back([this, i, t] { loop(i, t); });
        for (auto& w : workers_) w.detach();
    }
    void loop(int idx, int T) {
        unsigned long long seen = 0;
        for (;;) {
            const std::function<void(int, int)>* fn;
            {
                std::unique_lock<std::mutex> lk(m_);
                cv_.wait(lk, [&] { return gen_ != seen; });
                seen = gen_;
                fn = fn_;
            }
            (*fn)(idx, T);
            {
                std::lock_guard<std::mutex> lk(m_);
                if (--pending_ == 0) done_.notify_one();
            }
        }
    }
    std::vector<std::thread> workers_;
    std::mutex m_, run_m_;
    std::condition_variable cv_, done_;
    const std::function<void(int, int)>* fn_ = nullptr;
    unsigned long long gen_ = 0;
    int pending_ = 0;
};

struct PackHdr { float reward; uint32_t found; uint8_t target_find, terminated, win, reset; uint32_t pad; };
static_assert(sizeof(PackHdr) == 16, "PackHdr");

// rebuilds the host rows of envs [e0, e1) from the packed records
void expand_range(const FlightParams& p, cs_flight_compact* c, int e0, int e1) {
    const int n = p.n, m = p.m, stride = p.state_stride;
    for (int e = e0; e < e1; ++e) {
        const unsigned char* rec = c->h_pack + (size_t)e * c->rec_bytes;
        const PackHdr* hd = reinterpret_cast<const PackHdr*>(rec);
        c->reward[e] = hd->reward;
        c->target_find[e] = hd->target_find;
        c->terminated[e] = hd->terminated;
        c->win[e] = hd->win;
        if (!hd->reset) {
            uint32_t diff = hd->found ^ c->shadow_found[e];
            float* row = c->state + (size_t)e * stride;                // (the agent part arrives by DMA)
            while (diff) {                                           // find flags that changed (:206-209)
                const int j = __builtin_ctz(diff);
                diff &= diff - 1;
                row[4 * n + 3 * j + 2] = ((hd->found >> j) & 1u) ? 1.0f : 0.0f;
            }
            c->shadow_found[e] = hd->found;
        }
    }
    (void)m;
}

void apply_reset_entries(cs_flight* h, unsigned count) {
    cs_flight_compact* c = h->hc;
    const FlightParams& p = h->p;
    const int n = p.n, m = p.m;
    for (unsigned k = 0; k < count; ++k) {
        const int* ent = reinterpret_cast<const int*>(c->h_pack + c->off_entries + (size_t)k * c->ent_bytes);
        const int e = ent[0];
        const float* xy = reinterpret_cast<const float*>(ent + 1);
        const uint32_t found = reinterpret_cast<const PackHdr*>(c->h_pack + (size_t)e * c->rec_bytes)->found;
        float* tr = c->state + (size_t)e * p.state_stride + 4 * n;
        for (int j = 0; j < m; ++j) {                                // the new episode's targets (:201-211)
            tr[3 * j] = xy[2 * j];
            tr[3 * j + 1] = xy[2 * j + 1];
            tr[3 * j + 2] = ((found >> j) & 1u) ? 1.0f : 0.0f;
        }
        c->shadow_found[e] = found;
    }
}

void free_compact(cs_flight_compact* c) {
    if (!c) return;
    if (c->pooled) {                                // the pool owns the arrays; it must not touch this env any more
        if (c->pool) c->pool->envs[c->pool_index] = nullptr;
        delete c;
        return;
    }
    cudaFree(c->d_pack);
    cudaFreeHost(c->h_pack);
    cudaFreeHost(c->state);
    free(c->reward); free(c->target_find); free(c->terminated); free(c->win); free(c->shadow_found);
    delete c;
}

}  // namespace

namespace csf {
void flight_compact_release(cs_flight* h) {
    free_compact(h->hc);
    h->hc = nullptr;
}
void flight_compact_mark_dirty(cs_flight* h) {
    if (h->hc) h->hc->dirty = true;
}
}  // namespace csf

extern "C" {

int cs_flight_host_compact_begin(cs_flight* h, cs_flight_host_views* out) {
    CS_REQUIRE(h && out, "cs_flight_host_compact_begin: null argument");
    const FlightParams& p = h->p;
    if (!h->hc) {
        cs_flight_compact* c = new (std::nothrow) cs_flight_compact();
        if (!c) { cs_set_error("out of host memory"); return CS_ERR_NOMEM; }
        memset(c, 0, sizeof(*c));
        const size_t E = (size_t)p.E;
        c->rec_bytes = 16;
        c->ent_bytes = (4 + 8 * (size_t)p.m + 15) & ~(size_t)15;
        c->cap = (int)(E / 16 > 64 ? E / 16 : (E < 64 ? E : 64));
        c->off_entries = (E * c->rec_bytes + 255) & ~(size_t)255;
        c->off_counter = c->off_entries + (size_t)c->cap * c->ent_bytes;
        c->pack_bytes = c->off_counter + 16;
        c->dirty = true;
        h->hc = c;
        cudaError_t e = cudaSetDevice(h->cfg.device);
        if (e == cudaSuccess) e = cudaMalloc(&c->d_pack, c->pack_bytes);
        if (e == cudaSuccess) e = cudaMemset(c->d_pack, 0, c->pack_bytes);
        if (e == cudaSuccess) e = cudaHostAlloc(reinterpret_cast<void**>(&c->h_pack), c->pack_bytes, cudaHostAllocDefault);
        if (e == cudaSuccess) e = cudaHostAlloc(reinterpret_cast<void**>(&c->state), E * p.state_stride * sizeof(float), cudaHostAllocDefault);   // DMA target
        if (e == cudaSuccess) memset(c->state, 0, E * p.state_stride * sizeof(float));
        c->reward = (float*)calloc(E, sizeof(float));
        c->target_find = (int32_t*)calloc(E, sizeof(int32_t));
        c->terminated = (uint8_t*)calloc(E, 1);
        c->win = (uint8_t*)calloc(E, 1);
        c->shadow_found = (uint32_t*)calloc(E, sizeof(uint32_t));
        if (e != cudaSuccess || !c->reward || !c->target_find || !c->terminated || !c->win || !c->state || !c->shadow_found) {
            flight_compact_release(h);
            if (e != cudaSuccess) CS_CUDA(e);
            cs_set_error("out of host memory");
            return CS_ERR_NOMEM;
        }
    }
    cs_flight_compact* c = h->hc;
    out->reward = c->reward; out->target_find = c->target_find; out->terminated = c->terminated; out->win = c->win;
    out->state = c->state; out->state_stride = p.state_stride;
    out->h2d_bytes_per_step = (uint64_t)p.E * p.n;
    out->d2h_bytes_per_step = (c->pooled ? (uint64_t)p.E * c->rec_bytes : c->pack_bytes) + (uint64_t)p.E * 16 * p.n;
    return CS_OK;
}

namespace {
// first part of an expansion: wait for the transfer, refresh the rows in full where needed; returns the number of reset
// entries to apply afterwards (0 after a full refresh) through *entries
int expand_prepare(cs_flight* h, cudaStream_t st, int sync, unsigned* entries) {
    cs_flight_compact* c = h->hc;
    const FlightParams& p = h->p;
    if (sync) CS_CUDA(cudaStreamSynchronize(st));
    const unsigned count = *reinterpret_cast<const unsigned*>(c->h_pack + c->off_counter);
    const bool overflow = count > (unsigned)c->cap;
    *entries = overflow ? 0u : count;
    if (c->dirty || overflow) {
        // full refresh: the device holds the complete rows (first step, after a reset / import, or more envs were reset
        // inside one call than the side region holds)
        CS_CUDA(cudaSetDevice(h->cfg.device));
        CS_CUDA(cudaMemcpyAsync(c->state, p.state, (size_t)p.E * p.state_stride * sizeof(float), cudaMemcpyDeviceToHost, st));
        CS_CUDA(cudaStreamSynchronize(st));
        for (int e = 0; e < p.E; ++e)
            c->shadow_found[e] = reinterpret_cast<const PackHdr*>(c->h_pack + (size_t)e * c->rec_bytes)->found;
        c->dirty = false;
    }
    return CS_OK;
}
}  // namespace

int cs_flight_host_expand(cs_flight* h, void* stream, int32_t sync) {
    CS_REQUIRE(h && h->hc, "cs_flight_host_expand: call cs_flight_host_compact_begin first");
    cs_flight* one[1] = {h};
    void* st[1] = {stream};
    return cs_flight_host_expand_many(one, 1, st, 1, sync);
}

int cs_flight_step_host_compact(cs_flight* h, const uint8_t* h_actions, uint32_t flags, void* stream) {
    CS_REQUIRE(h && h_actions && h->hc, "cs_flight_step_host_compact: bad argument (cs_flight_host_compact_begin first)");
    CS_REQUIRE(!h->hc->pooled, "cs_flight_step_host_compact: the env belongs to a host pool -- step it with cs_flight_host_pool_step");
    cs_flight_compact* c = h->hc;
    const FlightParams& p = h->p;
    cudaStream_t st = (cudaStream_t)stream;
    CS_CUDA(cudaSetDevice(h->cfg.device));
    CS_CUDA(cudaMemcpyAsync(h->d_actions, h_actions, (size_t)p.E * p.n, cudaMemcpyHostToDevice, st));
    CS_CUDA(flight_dispatch(h, MODE_STEP, h->d_actions, nullptr, 0u, st));
    CS_CUDA(flight_map_join(h, st));      // a host-buffer step is complete when it returns: the belief map too
    CS_CUDA(launch_pack(h, st));
    CS_CUDA(cudaMemcpyAsync(c->h_pack, c->d_pack, c->pack_bytes, cudaMemcpyDeviceToHost, st));
    // the agent part of every host row, in place: E pieces of 16n bytes, state_stride floats apart
    CS_CUDA(cudaMemcpy2DAsync(c->state, (size_t)p.state_stride * sizeof(float), p.obs, 16 * (size_t)p.n, 16 * (size_t)p.n, (size_t)p.E,
                              cudaMemcpyDeviceToHost, st));
    CS_CUDA(cudaGetLastError());
    if (!(flags & CS_HOST_NO_SYNC)) return cs_flight_host_expand(h, stream, 1);
    return CS_OK;
}

// `count` independent env batches in one call: every batch is enqueued on streams[i % n_streams]; then batch by batch
// the stream is synchronised and the rows are rebuilt, so that the host work of batch i overlaps the transfers of i+1.
int cs_flight_step_host_compact_many(cs_flight* const* envs, const uint8_t* const* h_actions, int32_t count, void* const* streams,
                                     int32_t n_streams, uint32_t flags) {
    CS_REQUIRE(envs && h_actions && streams && count >= 0 && n_streams >= 1, "cs_flight_step_host_compact_many: bad argument");
    for (int i = 0; i < count; ++i) {
        const int rc = cs_flight_step_host_compact(envs[i], h_actions[i], CS_HOST_NO_SYNC, streams[i % n_streams]);
        if (rc != CS_OK) return rc;
    }
    if (flags & CS_HOST_NO_SYNC) return CS_OK;
    return cs_flight_host_expand_many(envs, count, streams, n_streams, 1);
}

// All batches share ONE fork-join of the host pool: the work items are (batch, env range) pairs of ~4096 envs, handed out
// through an atomic counter (a fork-join per batch costs more than rebuilding a few thousand rows).
int cs_flight_host_expand_many(cs_flight* const* envs, int32_t count, void* const* streams, int32_t n_streams, int32_t sync) {
    CS_REQUIRE(envs && streams && count >= 0 && n_streams >= 1, "cs_flight_host_expand_many: bad argument");
    std::vector<unsigned> entries((size_t)count, 0u);
    std::vector<int> first((size_t)count + 1, 0);
    constexpr int kChunk = 4096;
    for (int i = 0; i < count; ++i) {
        CS_REQUIRE(envs[i] && envs[i]->hc, "cs_flight_host_expand_many: call cs_flight_host_compact_begin first (batch %d)", i);
        const int rc = expand_prepare(envs[i], (cudaStream_t)streams[i % n_streams], sync, &entries[(size_t)i]);
        if (rc != CS_OK) return rc;
        first[(size_t)i + 1] = first[(size_t)i] + (envs[i]->p.E + kChunk - 1) / kChunk;
    }
    const int items = first[(size_t)count];
    std::atomic<int> next{0};
    auto work = [&](int, int) {
        for (;;) {
            const int it = next.fetch_add(1, std::memory_order_relaxed);
            if (it >= items) break;
            int b = 0;
            while (first[(size_t)b + 1] <= it) ++b;
            const int e0 = (it - first[(size_t)b]) * kChunk;
            const int e1 = e0 + kChunk < envs[b]->p.E ? e0 + kChunk : envs[b]->p.E;
            expand_range(envs[b]->p, envs[b]->hc, e0, e1);
        }
    };
    if (items <= 2) work(0, 1);
    else HostPool::get().run(work);
    for (int i = 0; i < count; ++i) apply_reset_entries(envs[i], entries[(size_t)i]);
    return CS_OK;
}

}  // extern "C"

// ---- pooled host buffers: every batch's arrays are segments of single allocations ------------------------------------
namespace {

void pool_free(cs_flight_host_pool* pl) {
    if (!pl) return;
    for (int i = 0; i < pl->count; ++i)
        if (pl->envs[i] && pl->envs[i]->hc && pl->envs[i]->hc->pooled) flight_compact_release(pl->envs[i]);
    if (pl->group) cs_flight_group_destroy(pl->group);
    cudaFreeHost(pl->h_actions); cudaFree(pl->d_actions); cudaFree(pl->d_rec); cudaFreeHost(pl->h_rec); cudaFree(pl->d_agent);
    cudaFreeHost(pl->h_state);
    if (pl->ev_rec) cudaEventDestroy(pl->ev_rec);
    free(pl->shadow_found);
    delete pl;
}

inline const uint32_t* pool_found(const cs_flight_host_pool* pl) { return reinterpret_cast<const uint32_t*>(pl->h_rec + pl->off_found); }

// rows of the envs that were reset inside the call: the new episode's targets (flight_env_easy.py:201-211)
void pool_apply_entries(cs_flight_host_pool* pl, unsigned count) {
    const FlightParams& p = pl->envs[0]->p;
    const int n = p.n, m = p.m;
    const uint32_t* found_all = pool_found(pl);
    for (unsigned k = 0; k < count; ++k) {
        const int* ent = reinterpret_cast<const int*>(pl->h_rec + pl->off_entries + (size_t)k * pl->ent_bytes);
        const int ge = ent[0];
        const float* xy = reinterpret_cast<const float*>(ent + 1);
        const uint32_t found = found_all[ge];
        float* tr = pl->h_state + (size_t)ge * p.state_stride + 4 * n;
        for (int j = 0; j < m; ++j) {
            tr[3 * j] = xy[2 * j];
            tr[3 * j + 1] = xy[2 * j + 1];
            tr[3 * j + 2] = ((found >> j) & 1u) ? 1.0f : 0.0f;
        }
        pl->shadow_found[ge] = found;
    }
}

// find flags whose bit changed, envs [e0, e1) (:206-209)
void pool_scan_found(cs_flight_host_pool* pl, int e0, int e1) {
    const FlightParams& p = pl->envs[0]->p;
    const uint32_t* found = pool_found(pl);
    uint32_t* shadow = pl->shadow_found;
    const int n = p.n, stride = p.state_stride;
    int e = e0;
    while (e < e1) {
        const int blk = e + 16 <= e1 ? 16 : e1 - e;
        if (memcmp(found + e, shadow + e, (size_t)blk * 4) != 0) {
            for (int k = e; k < e + blk; ++k) {
                uint32_t diff = found[k] ^ shadow[k];
                if (!diff) continue;
                float* row = pl->h_state + (size_t)k * stride;
                while (diff) {
                    const int j = __builtin_ctz(diff);
                    diff &= diff - 1;
                    row[4 * n + 3 * j + 2] = ((found[k] >> j) & 1u) ? 1.0f : 0.0f;
                }
                shadow[k] = found[k];
            }
        }
        e += blk;
    }
}

}  // namespace

extern "C" {

int cs_flight_host_pool_create(cs_flight* const* envs, int32_t count, cs_flight_host_pool** out) {
    CS_REQUIRE(envs && out && count >= 1 && count <= kMaxGroup, "cs_flight_host_pool_create: count must be in 1..%d", kMaxGroup);
    for (int i = 0; i < count; ++i) {
        CS_REQUIRE(envs[i] && !envs[i]->hc, "cs_flight_host_pool_create: batch %d is null or already has compact host buffers", i);
        const FlightParams &a = envs[0]->p, &b = envs[i]->p;
        CS_REQUIRE(a.n == b.n && a.m == b.m && a.state_stride == b.state_stride && a.auto_reset == b.auto_reset && a.meta_off == b.meta_off &&
                   envs[0]->cfg.device == envs[i]->cfg.device, "cs_flight_host_pool_create: batch %d differs from batch 0 in n_agents / target_num / auto_reset / device", i);
    }
    cs_flight_host_pool* pl = new (std::nothrow) cs_flight_host_pool();
    if (!pl) { cs_set_error("out of host memory"); return CS_ERR_NOMEM; }
    memset(pl, 0, sizeof(*pl));
    pl->count = count;
    pl->device = envs[0]->cfg.device;
    const FlightParams& p0 = envs[0]->p;
    for (int i = 0; i < count; ++i) {
        pl->envs[i] = envs[i];
        pl->geo.first[i + 1] = pl->geo.first[i] + envs[i]->p.E;
        const FlightParams& p = envs[i]->p;
        GroupEntry& t = pl->table.h[i];
        t.E = p.E; t.env_id_base = p.env_id_base; t.seed = p.seed; t.dyn_rs = p.dyn_rs; t.dyn_es = p.dyn_es; t.tgt_rs = p.tgt_rs; t.tgt_es = p.tgt_es;
        t.dyn = p.dyn; t.tgt = p.tgt; t.obs = p.obs; t.state = p.state; t.reward = p.reward; t.terminated = p.terminated; t.win = p.win;
        t.target_find = p.target_find; t.stats = p.stats; t.tmpl = p.tmpl;
    }
    pl->total = pl->geo.first[count];
    const size_t T = (size_t)pl->total, n = (size_t)p0.n;
    auto al = [](size_t v) { return (v + 255) & ~(size_t)255; };
    pl->ent_bytes = (4 + 8 * (size_t)p0.m + 15) & ~(size_t)15;
    pl->cap = (int)(T / 16 > 64 ? T / 16 : (T < 64 ? T : 64));
    // a step copies the result arrays, the counter and the first cap_fast reset entries; the rest of the list only when it
    // is used (many envs reset inside one call)
    pl->cap_fast = (int)(T / 48 > 64 ? T / 48 : 64);
    if (pl->cap_fast > pl->cap) pl->cap_fast = pl->cap;
    pl->off_tf = al(T * 4);
    pl->off_term = pl->off_tf + al(T * 4);
    pl->off_win = pl->off_term + al(T);
    pl->off_found = pl->off_win + al(T);
    pl->off_counter = pl->off_found + al(T * 4);
    pl->off_entries = pl->off_counter + 256;
    pl->fast_bytes = pl->off_entries + (size_t)pl->cap_fast * pl->ent_bytes;
    pl->rec_block_bytes = pl->off_entries + (size_t)pl->cap * pl->ent_bytes;
    cudaError_t e = cudaSetDevice(pl->device);
    if (e == cudaSuccess) e = cudaHostAlloc(reinterpret_cast<void**>(&pl->h_actions), T * n, cudaHostAllocDefault);
    if (e == cudaSuccess) e = cudaMalloc(&pl->d_actions, T * n);
    if (e == cudaSuccess) e = cudaMalloc(&pl->d_rec, pl->rec_block_bytes);
    if (e == cudaSuccess) e = cudaMemset(pl->d_rec, 0, pl->rec_block_bytes);
    if (e == cudaSuccess) e = cudaHostAlloc(reinterpret_cast<void**>(&pl->h_rec), pl->rec_block_bytes, cudaHostAllocDefault);
    if (e == cudaSuccess) e = cudaMalloc(&pl->d_agent, T * 16 * n);
    if (e == cudaSuccess) e = cudaHostAlloc(reinterpret_cast<void**>(&pl->h_state), T * p0.state_stride * sizeof(float), cudaHostAllocDefault);
    if (e == cudaSuccess) e = cudaEventCreateWithFlags(&pl->ev_rec, cudaEventDisableTiming);
    pl->shadow_found = (uint32_t*)calloc(T, sizeof(uint32_t));
    bool ok = e == cudaSuccess && pl->shadow_found;
    if (ok) {
        memset(pl->h_actions, 0, T * n);
        memset(pl->h_rec, 0, pl->rec_block_bytes);
        memset(pl->h_state, 0, T * p0.state_stride * sizeof(float));
        for (int i = 0; i < count && ok; ++i) {
            cs_flight_compact* c = new (std::nothrow) cs_flight_compact();
            if (!c) { ok = false; break; }
            memset(c, 0, sizeof(*c));
            const size_t f = (size_t)pl->geo.first[i];
            c->pooled = true; c->dirty = true; c->rec_bytes = 14;
            c->pool = pl; c->pool_index = i;
            c->reward = reinterpret_cast<float*>(pl->h_rec) + f;
            c->target_find = reinterpret_cast<int32_t*>(pl->h_rec + pl->off_tf) + f;
            c->terminated = pl->h_rec + pl->off_term + f;
            c->win = pl->h_rec + pl->off_win + f;
            c->state = pl->h_state + f * p0.state_stride; c->shadow_found = pl->shadow_found + f;
            envs[i]->hc = c;
            pl->act_ptrs[i] = pl->d_actions + f * n;
        }
    }
    if (!ok) {
        pool_free(pl);
        if (e != cudaSuccess) CS_CUDA(e);
        cs_set_error("out of host memory");
        return CS_ERR_NOMEM;
    }
    // one grouped step launch when the handles allow it (flight_easy, thread-per-env kernel, same configuration)
    if (cs_flight_group_create(envs, count, &pl->group) != CS_OK) pl->group = nullptr;
    *out = pl;
    return CS_OK;
}

void cs_flight_host_pool_destroy(cs_flight_host_pool* pl) {
    if (!pl) return;
    cudaSetDevice(pl->device);
    pool_free(pl);
}

int cs_flight_host_pool_views(cs_flight_host_pool* pl, int32_t i, cs_flight_host_views* out, uint8_t** h_actions) {
    CS_REQUIRE(pl && out && i >= 0 && i < pl->count && pl->envs[i], "cs_flight_host_pool_views: bad argument");
    if (h_actions) *h_actions = pl->h_actions + (size_t)pl->geo.first[i] * pl->envs[i]->p.n;
    return cs_flight_host_compact_begin(pl->envs[i], out);
}

int cs_flight_host_pool_expand(cs_flight_host_pool* pl, void* stream, int32_t sync) {
    CS_REQUIRE(pl, "cs_flight_host_pool_expand: null pool");
    for (int i = 0; i < pl->count; ++i) CS_REQUIRE(pl->envs[i], "cs_flight_host_pool_expand: batch %d was destroyed", i);
    cudaStream_t st = (cudaStream_t)stream;
    const FlightParams& p0 = pl->envs[0]->p;
    bool any_dirty = false;
    for (int i = 0; i < pl->count; ++i) any_dirty |= pl->envs[i]->hc->dirty;
    // sync = 1: the host work below only needs the result block -- it runs while the copy engine still scatters the agent
    // rows (different bytes of the same host rows; DMA writes are coherent with the cores' caches), and the stream is
    // joined at the end.  A full refresh needs everything first.
    if (sync) {
        if (any_dirty) CS_CUDA(cudaStreamSynchronize(st));
        else CS_CUDA(cudaEventSynchronize(pl->ev_rec));
    }
    const unsigned count = *reinterpret_cast<const unsigned*>(pl->h_rec + pl->off_counter);
    const bool overflow = count > (unsigned)pl->cap;
    if (!overflow && count > (unsigned)pl->cap_fast) {          // the tail of the reset list (rare)
        CS_CUDA(cudaSetDevice(pl->device));
        CS_CUDA(cudaStreamSynchronize(st));
        CS_CUDA(cudaMemcpy(pl->h_rec + pl->fast_bytes, pl->d_rec + pl->fast_bytes, (size_t)(count - pl->cap_fast) * pl->ent_bytes, cudaMemcpyDeviceToHost));
    }
    bool refreshed = false;
    if (any_dirty || overflow) {
        if (sync && !any_dirty) CS_CUDA(cudaStreamSynchronize(st));
        for (int i = 0; i < pl->count; ++i) {
            cs_flight* h = pl->envs[i];
            if (!h->hc->dirty && !overflow) continue;
            // full refresh of this batch's rows (first step, after a reset / import, or more resets than the side list holds)
            if (!refreshed) CS_CUDA(cudaSetDevice(pl->device));
            CS_CUDA(cudaMemcpyAsync(h->hc->state, h->p.state, (size_t)h->p.E * p0.state_stride * sizeof(float), cudaMemcpyDeviceToHost, st));
            refreshed = true;
        }
    }
    if (refreshed) {
        CS_CUDA(cudaStreamSynchronize(st));
        const uint32_t* found = pool_found(pl);
        for (int i = 0; i < pl->count; ++i) {
            cs_flight_compact* c = pl->envs[i]->hc;
            if (!c->dirty && !overflow) continue;
            memcpy(c->shadow_found, found + pl->geo.first[i], (size_t)pl->envs[i]->p.E * 4);
            c->dirty = false;
        }
    }
    if (!overflow) pool_apply_entries(pl, count);
    // few threads on purpose: a D2H copy into lines that sit in many cores' caches runs at a fraction of its speed
    // (measured: the 4 MB result block in 95 us after one thread read it, 530 us after sixteen did)
    constexpr int kScanThreads = 2;
    const int chunk = (pl->total + kScanThreads - 1) / kScanThreads;
    if (pl->total < 65536 || HostPool::get().size() < kScanThreads) pool_scan_found(pl, 0, pl->total);
    else HostPool::get().run([&](int idx, int) { if (idx < kScanThreads) pool_scan_found(pl, idx * chunk, (idx + 1) * chunk < pl->total ? (idx + 1) * chunk : pl->total); });
    if (sync) CS_CUDA(cudaStreamSynchronize(st));               // the agent rows are in place
    return CS_OK;
}

int cs_flight_host_pool_step(cs_flight_host_pool* pl, const uint8_t* h_actions, uint32_t flags, void* stream) {
    CS_REQUIRE(pl, "cs_flight_host_pool_step: null pool");
    for (int i = 0; i < pl->count; ++i) CS_REQUIRE(pl->envs[i], "cs_flight_host_pool_step: batch %d was destroyed", i);
    cudaStream_t st = (cudaStream_t)stream;
    const FlightParams& p0 = pl->envs[0]->p;
    const size_t T = (size_t)pl->total, n = (size_t)p0.n;
    CS_CUDA(cudaSetDevice(pl->device));
    CS_CUDA(cudaMemcpyAsync(pl->d_actions, h_actions ? h_actions : pl->h_actions, T * n, cudaMemcpyHostToDevice, st));
    if (pl->group) {
        bool was_dirty[kMaxGroup];
        for (int i = 0; i < pl->count; ++i) was_dirty[i] = pl->envs[i]->hc->dirty;
        const int rc = cs_flight_group_step(pl->group, pl->act_ptrs, stream);      // (marks the host rows stale; this step refreshes them)
        for (int i = 0; i < pl->count; ++i) pl->envs[i]->hc->dirty = was_dirty[i];
        if (rc != CS_OK) return rc;
    } else {
        for (int i = 0; i < pl->count; ++i) {
            CS_CUDA(flight_dispatch(pl->envs[i], MODE_STEP, pl->act_ptrs[i], nullptr, 0u, st));
            CS_CUDA(flight_map_join(pl->envs[i], st));
        }
    }
    CS_CUDA(launch_pack_pool(pl, st));
    CS_CUDA(cudaMemcpyAsync(pl->h_rec, pl->d_rec, pl->fast_bytes, cudaMemcpyDeviceToHost, st));
    CS_CUDA(cudaEventRecord(pl->ev_rec, st));
    // the agent part of every host row, in place: `total` pieces of 16n bytes, state_stride floats apart
    CS_CUDA(cudaMemcpy2DAsync(pl->h_state, (size_t)p0.state_stride * sizeof(float), pl->d_agent, 16 * n, 16 * n, T, cudaMemcpyDeviceToHost, st));
    if (!(flags & CS_HOST_NO_SYNC)) return cs_flight_host_pool_expand(pl, stream, 1);
    return CS_OK;
}

}  // extern "C"
