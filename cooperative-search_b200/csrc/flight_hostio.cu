// Compact host-buffer path of the flight envs: the call a CPU-side rollout makes every step
// (reward, terminated, info = env.step(actions); env.get_obs(); env.get_state(), common/rollout.py:45-63) moves
// 16 + 16n bytes per env over PCIe instead of the full reference-shaped state row (4(4n + 3m) + 10 bytes) and rebuilds
// the rows on the host: the agent part is overwritten, find flags flip where the found mask changed, and the 2m target
// coordinates are rewritten only for envs that were reset inside the call.  The rebuild is spread over a small pool
// of host threads.  Results live in library-owned host arrays (cs_flight_host_views), valid until the next step.
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
    void run(const std::function<void(int, int)>& fn) {
        const int T = size();
        if (T == 1) { fn(0, 1); return; }
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
    std::mutex m_;
    std::condition_variable cv_, done_;
    const std::function<void(int, int)>* fn_ = nullptr;
    unsigned long long gen_ = 0;
    int pending_ = 0;
};

struct PackHdr { float reward; uint32_t found; uint8_t target_find, terminated, win, reset; uint32_t pad; };
static_assert(sizeof(PackHdr) == 16, "PackHdr");

// rebuilds the host rows of envs [e0, e1) from the packed records
void expand_range(cs_flight* h, int e0, int e1) {
    cs_flight_compact* c = h->hc;
    const FlightParams& p = h->p;
    const int n = p.n, m = p.m, stride = p.state_stride;
    for (int e = e0; e < e1; ++e) {
        const unsigned char* rec = c->h_pack + (size_t)e * c->rec_bytes;
        const PackHdr* hd = reinterpret_cast<const PackHdr*>(rec);
        c->reward[e] = hd->reward;
        c->target_find[e] = hd->target_find;
        c->terminated[e] = hd->terminated;
        c->win[e] = hd->win;
        float* row = c->state + (size_t)e * stride;
        {                                                            // agent part = get_obs rows (flight_env_easy.py:192-193)
            typedef float v4 __attribute__((vector_size(16), aligned(4)));
            const v4* src = reinterpret_cast<const v4*>(rec + 16);
            v4* dst = reinterpret_cast<v4*>(row);
            for (int a = 0; a < n; ++a) dst[a] = src[a];
        }
        if (!hd->reset) {
            uint32_t diff = hd->found ^ c->shadow_found[e];
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
    cudaFree(c->d_pack);
    cudaFreeHost(c->h_pack);
    free(c->reward); free(c->target_find); free(c->terminated); free(c->win); free(c->state); free(c->shadow_found);
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
        c->rec_bytes = 16 + 16 * (size_t)p.n;
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
        c->reward = (float*)calloc(E, sizeof(float));
        c->target_find = (int32_t*)calloc(E, sizeof(int32_t));
        c->terminated = (uint8_t*)calloc(E, 1);
        c->win = (uint8_t*)calloc(E, 1);
        c->state = (float*)calloc(E * p.state_stride, sizeof(float));
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
    out->d2h_bytes_per_step = c->pack_bytes;
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
    cs_flight_compact* c = h->hc;
    const FlightParams& p = h->p;
    cudaStream_t st = (cudaStream_t)stream;
    CS_CUDA(cudaSetDevice(h->cfg.device));
    CS_CUDA(cudaMemcpyAsync(h->d_actions, h_actions, (size_t)p.E * p.n, cudaMemcpyHostToDevice, st));
    CS_CUDA(flight_dispatch(h, MODE_STEP, h->d_actions, nullptr, 0u, st));
    CS_CUDA(flight_map_join(h, st));      // a host-buffer step is complete when it returns: the belief map too
    CS_CUDA(launch_pack(h, st));
    CS_CUDA(cudaMemcpyAsync(c->h_pack, c->d_pack, c->pack_bytes, cudaMemcpyDeviceToHost, st));
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
            expand_range(envs[b], e0, e1);
        }
    };
    if (items <= 2) work(0, 1);
    else HostPool::get().run(work);
    for (int i = 0; i < count; ++i) apply_reset_entries(envs[i], entries[(size_t)i]);
    return CS_OK;
}

}  // extern "C"
