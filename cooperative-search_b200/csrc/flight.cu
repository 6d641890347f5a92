// flight_easy / flight env-step hot path for B200 (sm_100a).
//
// What is restated here (reference: WZN1ng/Cooperative-Search, pure Python):
//   _agent_step + _potential_energy_force   env/flight_env_easy.py:255-301, env/flight_env.py:305-355
//   _update_obs (detection, reward, win)    env/flight_env_easy.py:223-253, env/flight_env.py:232-266
//   _update_prob_map / _percent_in_...      env/flight_env.py:275-303
//   step / reset / get_obs / get_state      env/flight_env_easy.py:79-221,303-314
//
// Design (DESIGN.md has the full account):
//   * one GROUP of LPE lanes (1..32, template) per env instance; a CTA of 128 threads stages the
//     fp64 state records of its 128/LPE envs in shared memory with coalesced 16-byte loads, the
//     groups work on the staged records, and the CTA writes state + fp32 outputs back coalesced.
//   * positions / headings / targets are fp64 so that the in-range test and the wall test take the
//     same branch as the reference's Python floats; compiled with -fmad=false so a*b+c keeps the
//     reference's two roundings.
//   * detection draws are keyed Philox words (cs_philox.cuh), order independent.
//   * the belief-map update (variant 1) runs warp-per-env inside the same kernel, half-warp per
//     map row so that every load/store instruction covers one contiguous 64-byte run of a row.
#include <cuda.h>
#include <math.h>
#include <stdlib.h>
#include <new>
#include <vector>
#include "cs_common.cuh"
#include "cs_philox.cuh"

namespace {

constexpr int kThreads = 128;
struct FlightParams {
    int E, n, m, M, T;
    int variant, auto_reset, agent_mode, target_mode, count_touched;
    // per-env record geometry (doubles)
    int rec, yaw_off, meta_off, state_len;
    long long dyn_rs, dyn_es, tgt_rs, tgt_es;   // strides (in doubles) of a row / an env in dyn and tgt
    int state_stride;            // floats per state row in HBM (state_len rounded up to a multiple of 4)
    // per-warp shared-memory scratch of the step kernel (doubles): the heading-table index
    int s_lut, s_warp;
    int span_cap, span_shift;    // power of two >= 2R: corner rows per agent in the interval pass (and its log2)
    int rs_shift;                // log2 of the row slots per agent box in the map sweep (power of two >= 2R+1)
    int lps_shift;               // log2 of the lanes per row run in the map sweep (16 cells per lane, >= 2R+2 cells)
    int ms_own, ms_col, ms_box, ms_xy, ms_hit, ms_warp;   // map kernel: per-warp scratch offsets / size in 8-byte words
    int mt_R, mt_box, mt_xy, mt_hit, mt_bar, mt_group, tile_hshift, tile_stride;   // TMA map kernel: per-env scratch offsets / size in bytes, log2 of the row pairs per tile, bytes between tiles (multiple of 128)
    int pre_stride;              // doubles per env in `pre`
    double* pre;                 // [E][pre_stride]: agent xy (2n) | int nh, hit cells -- the sensing before an in-call auto-reset
    double Md, half_M, inv_half, R, R2, v, fk, fd2, near2, q_miss;
    double turn, pi, two_pi, three_pi, half_pi;
    long long thr;
    uint32_t seed, env_id_base;
    double* dyn;
    double* tgt;
    float* obs;
    float* state;
    float* reward;
    uint8_t* terminated;
    uint8_t* win;
    int32_t* target_find;
    float* prob_map;
    double* stats;
    const double* tmpl;   // [m][5]: x, y, sx, sy, random   (already scaled by a = M/10)
    // heading-lattice trig table (see HeadingLut below)
    const longlong2* lut_meta;   // [37]: x = bit pattern of the cluster centre, y = base | (half << 32)
    const double2* lut;          // (sin, cos) of every bit pattern in every cluster window, from the HOST libm
    double inv_turn;             // 18/pi
    double cos0, sin0;           // cos/sin of the start heading of agent_mode, from the host libm
    double lin[CS_MAX_AGENTS];   // i*map_size/(n-1) (map_size/2 for n == 1), flight_env_easy.py:140-143
};

// ------------------------------------------------------------------------------------------------
// State layout, chosen per handle (cs_flight_create): element (row r, env e) of dyn lives at dyn[r*dyn_rs + e*dyn_es],
// of tgt at tgt[r*tgt_rs + e*tgt_es].  Rows of dyn: 0..2n-1 = x0, y0, x1, y1, ..., 2n..3n-1 the headings,
// meta_off..meta_off+3 the 8 uint32 meta words (two per row); rows of tgt: 2j = x of target j, 2j+1 = y.
//   * structure of arrays (rs = E, es = 1), from 32768 envs per handle (= one thread per env, flight_tpe_kernel<N, 1>): with one thread per env every load and store
//     of a warp is one contiguous 256-byte run (the record layout cost the thread-per-env kernel ~1100 L1 wavefronts
//     per warp and bounded it: 1M envs 5.05e9 -> 5.5e9, 65536 envs 2.67e9 -> 2.9e9 env-steps/s);
//   * record per env (rs = 1, es = rows), below: a launch of a few thousand envs is latency-bound and 12 % faster
//     when a warp's state is a handful of consecutive lines instead of 43 rows 32 KB apart.
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ double* dyn_at(const FlightParams& p, int row, int e) {
    return p.dyn + (size_t)row * p.dyn_rs + (size_t)e * p.dyn_es;
}
// record layout (row stride 1): pairs of rows are adjacent and 16-byte aligned -> one 16-byte access
__device__ __forceinline__ double2 xy_ld(const FlightParams& p, int a, int e) {
    const double* q = dyn_at(p, 2 * a, e);
    if (p.dyn_rs == 1) return *reinterpret_cast<const double2*>(q);
    return make_double2(q[0], q[p.dyn_rs]);
}
__device__ __forceinline__ void xy_st(const FlightParams& p, int a, int e, double x, double y) {
    double* q = dyn_at(p, 2 * a, e);
    if (p.dyn_rs == 1) { *reinterpret_cast<double2*>(q) = make_double2(x, y); return; }
    q[0] = x;
    q[p.dyn_rs] = y;
}
__device__ __forceinline__ double2 tgt_ld(const FlightParams& p, int j, int e) {
    const double* t = p.tgt + (size_t)(2 * j) * p.tgt_rs + (size_t)e * p.tgt_es;
    if (p.tgt_rs == 1) return *reinterpret_cast<const double2*>(t);
    return make_double2(t[0], t[p.tgt_rs]);
}
__device__ __forceinline__ void tgt_st(const FlightParams& p, int j, int e, double2 v) {
    double* t = p.tgt + (size_t)(2 * j) * p.tgt_rs + (size_t)e * p.tgt_es;
    if (p.tgt_rs == 1) { *reinterpret_cast<double2*>(t) = v; return; }
    t[0] = v.x;
    t[p.tgt_rs] = v.y;
}
__device__ __forceinline__ uint2* meta_at(const FlightParams& p, int pair, int e) {       // words 2*pair, 2*pair+1
    return reinterpret_cast<uint2*>(dyn_at(p, p.meta_off + pair, e));
}
__device__ __forceinline__ void meta_ld(const FlightParams& p, int e, uint4* m0, uint4* m1) {
    if (p.dyn_rs == 1) {
        const uint4* mp = reinterpret_cast<const uint4*>(dyn_at(p, p.meta_off, e));
        *m0 = mp[0]; *m1 = mp[1];
        return;
    }
    const uint2 a = *meta_at(p, 0, e), b = *meta_at(p, 1, e), c = *meta_at(p, 2, e), d = *meta_at(p, 3, e);
    *m0 = make_uint4(a.x, a.y, b.x, b.y);
    *m1 = make_uint4(c.x, c.y, d.x, d.y);
}
__device__ __forceinline__ void meta_st(const FlightParams& p, int e, uint4 m0, uint4 m1) {
    if (p.dyn_rs == 1) {
        uint4* mp = reinterpret_cast<uint4*>(dyn_at(p, p.meta_off, e));
        mp[0] = m0; mp[1] = m1;
        return;
    }
    *meta_at(p, 0, e) = make_uint2(m0.x, m0.y);
    *meta_at(p, 1, e) = make_uint2(m0.z, m0.w);
    *meta_at(p, 2, e) = make_uint2(m1.x, m1.y);
    *meta_at(p, 3, e) = make_uint2(m1.z, m1.w);
}

// The same accessors with the layout known at compile time (thread-per-env kernel: structure of arrays with one thread
// per env, record per env with 4 threads per env): no stride arithmetic, 16-byte accesses in the record layout.
template <bool SOA>
struct Lay {
    static __device__ __forceinline__ double2 xy_ld(const FlightParams& p, int a, int e) {
        if (SOA) { const double* q = p.dyn + (size_t)(2 * a) * p.E + e; return make_double2(q[0], q[p.E]); }
        return *reinterpret_cast<const double2*>(p.dyn + (size_t)e * p.rec + 2 * a);
    }
    static __device__ __forceinline__ void xy_st(const FlightParams& p, int a, int e, double x, double y) {
        if (SOA) { double* q = p.dyn + (size_t)(2 * a) * p.E + e; q[0] = x; q[p.E] = y; return; }
        *reinterpret_cast<double2*>(p.dyn + (size_t)e * p.rec + 2 * a) = make_double2(x, y);
    }
    static __device__ __forceinline__ double* yaw_at(const FlightParams& p, int a, int e) {
        return SOA ? p.dyn + (size_t)(p.yaw_off + a) * p.E + e : p.dyn + (size_t)e * p.rec + p.yaw_off + a;
    }
    static __device__ __forceinline__ double2 tgt_ld(const FlightParams& p, int j, int e) {
        if (SOA) { const double* t = p.tgt + (size_t)(2 * j) * p.E + e; return make_double2(t[0], t[p.E]); }
        return *reinterpret_cast<const double2*>(p.tgt + ((size_t)e * p.m + j) * 2);
    }
    static __device__ __forceinline__ void meta_ld(const FlightParams& p, int e, uint4* m0, uint4* m1) {
        if (SOA) {
            const uint2* q = reinterpret_cast<const uint2*>(p.dyn + (size_t)p.meta_off * p.E + e);
            const uint2 a = q[0], b = q[p.E], c = q[2 * (size_t)p.E], d = q[3 * (size_t)p.E];
            *m0 = make_uint4(a.x, a.y, b.x, b.y);
            *m1 = make_uint4(c.x, c.y, d.x, d.y);
            return;
        }
        const uint4* mp = reinterpret_cast<const uint4*>(p.dyn + (size_t)e * p.rec + p.meta_off);
        *m0 = mp[0]; *m1 = mp[1];
    }
    static __device__ __forceinline__ void meta_st(const FlightParams& p, int e, uint4 m0, uint4 m1) {
        if (SOA) {
            uint2* q = reinterpret_cast<uint2*>(p.dyn + (size_t)p.meta_off * p.E + e);
            q[0] = make_uint2(m0.x, m0.y); q[p.E] = make_uint2(m0.z, m0.w);
            q[2 * (size_t)p.E] = make_uint2(m1.x, m1.y); q[3 * (size_t)p.E] = make_uint2(m1.z, m1.w);
            return;
        }
        uint4* mp = reinterpret_cast<uint4*>(p.dyn + (size_t)e * p.rec + p.meta_off);
        mp[0] = m0; mp[1] = m1;
    }
};

// ------------------------------------------------------------------------------------------------
// cos/sin of a heading.
//
// Headings only ever take the values reachable from {0, pi/2, pi} under +-pi/18 turns, the 2*pi wrap and
// the wall reflection (flight_env_easy.py:259-266,281-284): 37 clusters of fp64 values, each a few hundred
// ulps wide after 200 steps (rounding drift ~1.7e-16 per step).  Whether an agent that comes back to a
// wall ends at y = 0.0 or y = -5e-17 -- and therefore its out-of-map flag, its reward and its reflected
// heading -- depends on the LAST BIT of sin/cos, and the reference's bits are those of the host libm
// (numpy -> glibc), which is not correctly rounded (measured: 2 of 652 evaluations).  So the handle tabulates
// the host libm's sin/cos for every bit pattern in every cluster window at create time and the kernel looks
// the pair up (one 16-byte load) -- bit-identical to the reference and cheaper than evaluating sincos().
// Off-lattice headings (user-injected state) fall back to CUDA's sincos (<= 2 ulp).
// ------------------------------------------------------------------------------------------------
__device__ __noinline__ void offlattice_sincos(double h, double* sn, double* c) { sincos(h, sn, c); }

__device__ __forceinline__ void heading_sincos(const FlightParams& p, const longlong2* lutm, double h, double* sn, double* c) {
    const int k = __double2int_rn(h * p.inv_turn);
    if (k == 0 && fabs(h) < 7.450580596923828e-09) {   // |h| < 2^-27: libm returns sin = h, cos = 1
        *sn = h;
        *c = 1.0;
        return;
    }
    if (k >= 1 && k <= 36) {
        const longlong2 mt = lutm[k];                      // cluster index staged in shared memory by the warp
        const long long off = __double_as_longlong(h) - mt.x;
        const long long half = mt.y >> 32;
        if (off >= -half && off <= half) {
            const double2 v = __ldg(p.lut + ((mt.y & 0xffffffffLL) + half + off));
            *sn = v.x;
            *c = v.y;
            return;
        }
    }
    offlattice_sincos(h, sn, c);
}

// ------------------------------------------------------------------------------------------------
// wall handling of one agent: env/flight_env_easy.py:278-290 ('>' test), env/flight_env.py:328 ('>=')
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ bool wall_reg(const FlightParams& p, double& x, double& y, double& h, double& c) {
    const double Md = p.Md;
    const bool outside = p.variant ? (x < 0.0 || x >= Md || y < 0.0 || y >= Md)
                                   : (x < 0.0 || x > Md || y < 0.0 || y > Md);
    if (outside) {
        x = fmin(fmax(x, 0.0), Md);
        y = fmin(fmax(y, 0.0), Md);
        h = (h <= p.pi) ? (p.pi - h) : (p.three_pi - h);
        c = -c;   // cos(pi - h) = cos(3pi - h) = -cos(h); sin unchanged (only the fp32 outputs use them)
    }
    return outside;
}

// Reset-time target placement (env/flight_env_easy.py:95-127) from the keyed stream; cold, out of line.
__device__ __noinline__ double2 draw_target(const FlightParams& p, uint32_t env_id, uint32_t episode, int j) {
    const cs_u4 w = cs_philox4x32_10(env_id, (episode & 0xFFFFu) << 16, (uint32_t)j, 0u, p.seed, CS_STREAM_TARGET);
    const double u1 = cs_u53(w.x, w.y), u2 = cs_u53(w.z, w.w);
    double x, y;
    if (p.target_mode == 0) {
        const double* row = p.tmpl + 5 * j;
        x = row[0];
        y = row[1];
        if (row[4] != 0.0) {                      // deter == 'f' (:106-110)
            const double rad = sqrt(-2.0 * log(u1));
            double sn, c;
            sincos(2.0 * p.pi * u2, &sn, &c);
            x += row[2] * 2.0 * (rad * c - 0.5);
            y += row[3] * 2.0 * (rad * sn - 0.5);
        }
    } else {                                      // target_mode 1 (:122-127)
        x = p.Md * u1;
        y = p.Md * u2;
    }
    return make_double2(x, y);
}

// Integer corner coordinates c with fl((c - a)^2) < R^2 -- a necessary condition for a corner in that
// row/column to be inside the agent's disc, evaluated with the SAME floating-point expression as the corner test
// (an agent at x = 1.0000000000000004 has corner x = 4 inside although fl(x + 3) = 4.0; bounds derived from
// floor/ceil of a +- R alone lose such corners).  Returns [lo, hi], at most 2R wide.
__device__ __forceinline__ void corner_span(double a, double R, double R2, int* lo, int* hi) {
    int c = (int)floor(a - R);
    double d = (double)c - a;
    *lo = (d * d < R2) ? c : c + 1;
    c = (int)ceil(a + R);
    d = (double)c - a;
    *hi = (d * d < R2) ? c : c - 1;
}

__device__ __forceinline__ bool corner_pred(double A, double cy, double ay, double R2) {
    const double dy = cy - ay;
    return A + dy * dy < R2;                                   // strict '<' (:300)
}

// The reference's own grouping (1-d)*p + (1-p) matters: 1-p is exact near p = 1.  A cell at exactly 1 (a found
// target) must map to exactly percent -- in particular stay exactly 1 while all four corners are in view --
// because the map's derivative there is 10 and any seed error would grow tenfold per step; that case is taken
// exactly, everything else uses the fast reciprocal (values never approach 1 from below: p' <= p).
__device__ __forceinline__ float belief_update(float pv, int cnt, float qf) {
    const float frac = 0.25f * (float)cnt;
    const float num = (frac * qf) * pv;                        // percent*(1-d)*p            (:292)
    const float den = fmaf(qf, pv, 1.0f - pv);                 // (1-d)*p + (1-p)
    return (pv == 1.0f && qf != 0.0f) ? frac : __fdividef(num, den);     // detect_prob = 1: 0/0 = nan, like the reference
}

// ------------------------------------------------------------------------------------------------
// belief map: _update_prob_map + _percent_in_agent_viewrange (env/flight_env.py:275-303).
//
// Runs as its OWN kernel right after the step / reset kernel, 16 lanes per env (two envs per warp).  The step kernel leaves a job for
// every env it sensed: a non-zero meta word CS_META_SENSE (cleared again by the next call that does not sense the env), the agent positions in the state record,
// the targets found by that call in CS_META_NEWFOUND; an env that was auto-reset inside the call was sensed twice
// (flight_env.py:266 runs inside reset() too) and its first job -- positions and hit cells before the reset -- sits in
// the `pre` side buffer.  Few registers (no Philox, no fp64 state in flight) and nothing but the map traffic
// outstanding: the SMs hold 48 warps each, which is what covers the HBM latency of the scattered 64-byte runs.
//
// The reference classifies the 4 corners of every cell against every agent (2500 x 4 x n fp64 tests).  Here:
//  (1) corner classification by ROW INTERVALS: for agent a and integer corner row cx, the corner columns cy with
//      fl(fl((cx-ax)^2) + fl((cy-ay)^2)) < R^2 form an interval (the expression is monotone in |cy-ay|).  Its
//      ends come from one fp32 sqrt; only when an end lies within 1e-3 of an integer is the reference's exact
//      fp64 predicate evaluated there.  One lane per (agent, corner row): <= 2R*n tasks.  The interval is OR-ed as a
//      bit mask into R[cx] (bit cy), the union over agents ("any agent", the `break` of :299-302).
//  (2) percent of cell (i,j) = popc of bits j,j+1 of R[i] and R[i+1]; touched <=> any of them set.
//  (3) sweep: one lane per (agent, box row): its 16-cell run (64 bytes) is loaded with 8 independent float2 loads, all
//      in flight together, updated in registers and stored.  A pair of cells belongs to the
//      FIRST agent whose (pair-aligned) box holds it, so every touched pair is loaded, updated and stored by exactly
//      one lane; untouched pairs are neither read nor written.  The update itself is fp32 and branch free
//      (DESIGN.md 4.4).
// Needs map_size <= 63 (one 64-bit mask per corner row); larger maps take flight_map_wide_kernel.
// ------------------------------------------------------------------------------------------------
constexpr int kMapThreads = 128;
#ifndef CS_MAP_MIN_CTAS
#define CS_MAP_MIN_CTAS 8      // resident CTAs per SM the map kernel is compiled for (register budget 65536/(128*N))
#endif

// cell of a target found by the sensing call: [min(int(x), M-1), min(int(y), M-1)], Python int() truncates toward
// zero (flight_env.py:279); negative indices never match a swept cell
__device__ __forceinline__ int hit_cell(const FlightParams& p, double tx, double ty) {
    const int ci = min((int)fmin(tx, p.Md), p.M - 1), cj = min((int)fmin(ty, p.Md), p.M - 1);
    return (ci < 0 || cj < 0) ? -1 : ci * p.M + cj;
}

// Loads job `job` of env e into the group's scratch: agent positions -> xy[2n], hit cells -> hit[]; returns the number
// of hit cells.  job 0 = the sensing before an in-call auto-reset (side buffer), job 1 = the state record as the step
// / reset kernel left it.  GL lanes cooperate (lane = index inside the group).
template <int GL>
__device__ __forceinline__ int map_job_load(const FlightParams& p, int e, int job, int lane, uint32_t newf, double* xy, int* hit) {
    const int n = p.n, m = p.m;
    int nh;
    if (job == 0) {
        const double* pj = p.pre + (size_t)e * p.pre_stride;
        for (int a = lane; a < n; a += GL) {
            const double2 v = reinterpret_cast<const double2*>(pj)[a];
            xy[2 * a] = v.x; xy[2 * a + 1] = v.y;
        }
        const int* ph = reinterpret_cast<const int*>(pj + 2 * n);
        nh = ph[0];
        for (int k = lane; k < nh; k += GL) hit[k] = ph[1 + k];
    } else {
        for (int a = lane; a < n; a += GL) {
            const double2 v = xy_ld(p, a, e);
            xy[2 * a] = v.x; xy[2 * a + 1] = v.y;
        }
        nh = __popc(newf);
        for (int j = lane; j < m; j += GL)
            if ((newf >> j) & 1u) {
                const double2 t = tgt_ld(p, j, e);
                hit[__popc(newf & ((1u << j) - 1u))] = hit_cell(p, t.x, t.y);
            }
    }
    return nh;
}

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ float rcp_approx(float x) {
    float r;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
    return r;
}

// one cell of the sweep, branch free: `on` selects between the updated and the old value
__device__ __forceinline__ float belief_cell(float pv, unsigned AB, unsigned cmask, bool on, float kq, float qf) {
    const float cnt = __uint_as_float(0x4B000000u | (unsigned)__popc(AB & cmask)) - 8388608.0f;   // corners of this cell in view, 0..4
    const float num = (cnt * kq) * pv;                          // percent*(1-d)*p            (:292), kq = (1-d)/4
    const float den = fmaf(qf, pv, 1.0f - pv);                  // (1-d)*p + (1-p)
    float r = num * rcp_approx(den);
    r = (pv == 1.0f && qf != 0.0f) ? 0.25f * cnt : r;           // see belief_update; detect_prob = 1: 0/0 = nan, like the reference
    return on ? r : pv;
}

// same, from the number of corners in view (0 = untouched)
__device__ __forceinline__ float belief_cell(float pv, int pc, float kq, float qf) {
    const float cnt = __uint_as_float(0x4B000000u | (unsigned)pc) - 8388608.0f;
    const float num = (cnt * kq) * pv;
    const float den = fmaf(qf, pv, 1.0f - pv);
    float r = num * rcp_approx(den);
    r = (pv == 1.0f && qf != 0.0f) ? 0.25f * cnt : r;
    return pc ? r : pv;
}

constexpr int kMapGL = 16;                                      // lanes per env in the map kernel
constexpr int kMapEnvsPerCta = kMapThreads / kMapGL;

// Direct form (any map_size <= 63), the default.  PAIRS: even map_size (float2 accesses).
template <bool PAIRS>
__global__ void __launch_bounds__(kMapThreads, CS_MAP_MIN_CTAS) flight_map_kernel(const __grid_constant__ FlightParams p, uint32_t seq) {
    constexpr int GL = kMapGL;
    constexpr unsigned FULL = 0xffffffffu;
    extern __shared__ __align__(16) unsigned long long msm[];
    const int grp = threadIdx.x / GL, lane = threadIdx.x % GL;
    const int e_raw = blockIdx.x * kMapEnvsPerCta + grp;
    const int e = min(e_raw, p.E - 1);
    const int n = p.n, M = p.M;
    uint4 m0, m1;
    meta_ld(p, e, &m0, &m1);
    const bool sensed = e_raw < p.E && (m1.w >> 1) == seq;  // else: not sensed by this call (finished env, masked reset)
    if (!__any_sync(FULL, sensed)) return;

    unsigned long long* S = msm + (size_t)grp * p.ms_warp;
    unsigned long long* R = S;                              // [M+2]  corner-row masks
    unsigned long long* own = S + p.ms_own;                 // [n<<rs_shift] touched cells of (agent, box row) that agent's sweep owns
    unsigned long long* col = S + p.ms_col;                 // [n]    columns of each agent's box (pair aligned)
    int4* box = reinterpret_cast<int4*>(S + p.ms_box);      // [n]    i0, i1, first column of the window, first corner row
    double* xy = reinterpret_cast<double*>(S + p.ms_xy);    // [2n]
    int* hit = reinterpret_cast<int*>(S + p.ms_hit);        // [m]
    float* map = p.prob_map + (size_t)e * M * M;
    const float qf = (float)p.q_miss, kq = 0.25f * qf;
    const int lps_shift = p.lps_shift, rs_shift = p.rs_shift;
    const int total = n << rs_shift;
    const int sub = lane & ((1 << lps_shift) - 1), seg = lane >> lps_shift;
    unsigned touched = 0;

    // both envs of the warp run the phases together (the barriers are whole-warp); `run` tells the groups apart
    for (int job = 0; job < 2; ++job) {
        const bool run = sensed && (job == 1 || (m1.w & 1u));
        if (!__any_sync(FULL, run)) continue;
        int nh = 0;
        if (run) {
            nh = map_job_load<GL>(p, e, job, lane, m0.y, xy, hit);
            for (int r = lane; r <= M + 1; r += GL) R[r] = 0ull;
        }
        __syncwarp();
        if (run) {
            for (int a = lane; a < n; a += GL) {
                const double ax = xy[2 * a], ay = xy[2 * a + 1];
                int lo, hi, clo;
                corner_span(ax, p.R, p.R2, &lo, &hi);
                clo = lo;
                const int i0 = max(0, lo - 1), i1 = min(M - 1, hi);      // cell i has corners i and i+1
                corner_span(ay, p.R, p.R2, &lo, &hi);
                int j0 = max(0, lo - 1), j1 = min(M - 1, hi);
                if (PAIRS) { j0 &= ~1; j1 |= 1; }                        // M is even: j1|1 <= M-1
                box[a] = make_int4(i0, i1, j0, clo);
                col[a] = (j0 <= j1 && i0 <= i1) ? (((2ull << j1) - 1ull) & ~((1ull << j0) - 1ull)) : 0ull;
            }
        }
        __syncwarp();
        // (1) corner-row intervals
        const int span = p.span_cap;                                     // power of two >= 2R
        for (int t = lane; t < n * span; t += GL) {
            if (!run) continue;
            const int a = t >> p.span_shift, r = t & (span - 1);
            const double axa = xy[2 * a], aya = xy[2 * a + 1];
            const int cx = box[a].w + r;
            const double dx = (double)cx - axa;
            const double A = dx * dx;
            if (!(A < p.R2) || cx < 0 || cx > M) continue;               // past the last corner row of this agent
            // candidate ends in fp32 (ay <= map_size: absolute error ~4e-6, far inside the 1e-3 guard band)
            const float wf = sqrtf(fmaxf((float)(p.R2 - A), 0.0f));
            const float ayf = (float)aya;
            const float yh = ayf + wf, yl = ayf - wf;
            float fh = floorf(yh), cl = ceilf(yl);
            if (yh - fh < 1e-3f || yh - fh > 1.0f - 1e-3f) {             // end within 1e-3 of an integer: decide exactly
                const double Y = (double)rintf(yh);
                fh = (float)(corner_pred(A, Y, aya, p.R2) ? Y : Y - 1.0);
            }
            if (cl - yl < 1e-3f || cl - yl > 1.0f - 1e-3f) {
                const double Y = (double)rintf(yl);
                cl = (float)(corner_pred(A, Y, aya, p.R2) ? Y : Y + 1.0);
            }
            const int yhi = min((int)fh, M), ylo = max((int)cl, 0);
            if (ylo <= yhi) {
                const unsigned long long mk = ((2ull << yhi) - 1ull) & ~((1ull << ylo) - 1ull);
                unsigned* w = reinterpret_cast<unsigned*>(&R[cx]);
                if ((unsigned)mk) atomicOr(w, (unsigned)mk);
                if ((unsigned)(mk >> 32)) atomicOr(w + 1, (unsigned)(mk >> 32));
            }
        }
        __syncwarp();
        // (2) slot t = (agent a, box row r): the touched cells of that map row inside a's box that no earlier agent's
        //     box holds; one lane per slot
        for (int t = lane; t < total; t += GL) {
            if (!run) continue;
            const int a = t >> rs_shift, r = t & ((1 << rs_shift) - 1);
            const int4 bx = box[a];
            const int i = bx.x + r;
            unsigned long long T = 0ull;
            if (i <= bx.y) {
                const unsigned long long Ri = R[i], Ri1 = R[i + 1];
                T = (Ri | (Ri >> 1) | Ri1 | (Ri1 >> 1)) & col[a];
                for (int q = 0; q < a; ++q) {
                    const int4 bq = box[q];
                    if (i >= bq.x && i <= bq.y) T &= ~col[q];
                }
            }
            own[t] = T;
        }
        __syncwarp();
        // (3) sweep: 16 cells per lane, (1 << lps_shift) lanes per slot; all of a lane's loads are in flight together
        for (int t = seg; t < total; t += GL >> lps_shift) {
            const int4 bx = box[t >> rs_shift];
            const int i = bx.x + (t & ((1 << rs_shift) - 1));
            const int j = bx.z + 16 * sub;
            const unsigned tm = (run && j < M) ? ((unsigned)(own[t] >> j) & 0xFFFFu) : 0u;     // percent == 0 -> untouched (:285-286)
            if (!tm) continue;
            const unsigned A = (unsigned)(R[i] >> j) & 0x1FFFFu, B = (unsigned)(R[i + 1] >> j) & 0x1FFFFu;
            const int c0 = i * M + j;
            float* rowp = map + c0;
            float2 v[8];
            if (PAIRS) {
#pragma unroll
                for (int c = 0; c < 8; ++c)
                    if ((tm >> (2 * c)) & 3u) v[c] = *reinterpret_cast<const float2*>(rowp + 2 * c);
            } else {
#pragma unroll
                for (int c = 0; c < 8; ++c) {
                    if ((tm >> (2 * c)) & 1u) v[c].x = rowp[2 * c];
                    if ((tm >> (2 * c)) & 2u) v[c].y = rowp[2 * c + 1];
                }
            }
#pragma unroll
            for (int h = 0; h < 2; ++h) {
                const unsigned AB = ((A >> (8 * h)) & 0x1FFu) | (((B >> (8 * h)) & 0x1FFu) << 16);
                const unsigned th = tm >> (8 * h);
#pragma unroll
                for (int c = 0; c < 4; ++c) {
                    v[4 * h + c].x = belief_cell(v[4 * h + c].x, AB, 0x00030003u << (2 * c), (th >> (2 * c)) & 1u, kq, qf);
                    v[4 * h + c].y = belief_cell(v[4 * h + c].y, AB, 0x00060006u << (2 * c), (th >> (2 * c)) & 2u, kq, qf);
                }
            }
            if (nh) {                                                    // targets found by THIS call -> 1 (:288-289)
                for (int k = 0; k < nh; ++k) {
                    const unsigned d = (unsigned)(hit[k] - c0);
                    if (d < 16u && ((tm >> d) & 1u)) {
#pragma unroll
                        for (int c = 0; c < 8; ++c) {
                            if (d == 2u * c) v[c].x = 1.0f;
                            if (d == 2u * c + 1u) v[c].y = 1.0f;
                        }
                    }
                }
            }
            if (PAIRS) {
#pragma unroll
                for (int c = 0; c < 8; ++c)
                    if ((tm >> (2 * c)) & 3u) *reinterpret_cast<float2*>(rowp + 2 * c) = v[c];
            } else {
#pragma unroll
                for (int c = 0; c < 8; ++c) {
                    if ((tm >> (2 * c)) & 1u) rowp[2 * c] = v[c].x;
                    if ((tm >> (2 * c)) & 2u) rowp[2 * c + 1] = v[c].y;
                }
            }
            touched += __popc(tm);
        }
        __syncwarp();                                                    // the scratch is reused by the next job
    }
    if (p.count_touched) {
        const unsigned tot = __reduce_add_sync(FULL, touched);
        if ((threadIdx.x & 31) == 0 && tot) atomicAdd(p.stats + CS_STAT_TOUCHED, (double)tot);
    }
}

// ------------------------------------------------------------------------------------------------
// belief map, TMA form (even map_size in 10..63, 2R+2 <= 16; opt-in with CS_MAP_TMA=1, measured 55.6 us against 49.6 us
// per 16384-env step for the direct form): the SMs issue no global load or store for the map.
//
// The map of an env is viewed as a 3-D tensor (2M, M/2, E): one "row" of the view is a PAIR of map rows, so that the
// view's strides are multiples of 16 bytes although a map row (4M bytes) is not.  The box of agent a then is two
// tiles of H row-pairs x 20 positions -- its even rows and its odd rows -- that ONE lane per warp (TMA instructions run
// on the uniform datapath) moves with cp.async.bulk.tensor: global -> shared memory (completion on an mbarrier) while
// the other lanes build the corner masks, and shared memory -> global after the update.  The inner coordinate of a
// tile must be a multiple of 16 bytes (measured: an illegal-instruction trap otherwise, tools/probes/tma_probe.cu),
// hence 20 positions for a box of at most 16 columns; where 4M is not a multiple of 16 a float4 of a tile can hold
// the last cells of an even row and the first cells of the odd row after it (position x of a row pair = cell
// (2*pair + (x >= M), x mod M)), which the sweep handles cell by cell.
// Every cell of every tile gets the same treatment, new = touched ? f(old) : old with `touched` and the corner count
// taken from the UNION masks R, so where tiles of two agents overlap both copies hold identical values and both
// stores write the same bytes: no ownership, no ordering between the stores.  Positions of a tile that fall outside
// the map are zero-filled on load and clipped on store by the TMA unit.
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    uint32_t ok = 0, spins = 0;
    while (!ok) {
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                     : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
        if (!ok && ++spins > (1u << 26)) __trap();          // a lost copy must fail loudly, not hang the GPU
    }
}

__global__ void __launch_bounds__(kMapThreads, 7) flight_map_tma_kernel(const __grid_constant__ FlightParams p,
                                                                        const __grid_constant__ CUtensorMap tmap, uint32_t seq) {
    constexpr int GL = kMapGL;
    constexpr unsigned FULL = 0xffffffffu;
    extern __shared__ __align__(128) unsigned char tsm[];
    const int grp = threadIdx.x / GL, lane = threadIdx.x % GL;
    const int e_raw = blockIdx.x * kMapEnvsPerCta + grp;
    const int e = min(e_raw, p.E - 1);
    const int n = p.n, M = p.M;
    uint4 m0, m1;
    meta_ld(p, e, &m0, &m1);
    const bool sensed = e_raw < p.E && (m1.w >> 1) == seq;  // else: not sensed by this call (finished env, masked reset)
    if (!__any_sync(FULL, sensed)) return;

    unsigned char* G = tsm + (size_t)grp * p.mt_group;                                  // tiles [n][2][H][20] first
    unsigned long long* R = reinterpret_cast<unsigned long long*>(G + p.mt_R);          // [M+2] corner-row masks
    int4* box = reinterpret_cast<int4*>(G + p.mt_box);                                  // [n] tile coordinates: c0, c1 of the even-row tile, of the odd-row tile
    double* xy = reinterpret_cast<double*>(G + p.mt_xy);                                // [2n]
    int* hit = reinterpret_cast<int*>(G + p.mt_hit);                                    // [m]
    int* clo = hit + p.m;                                                               // [n] first corner row of each agent
    // TMA instructions run on the uniform datapath: exactly one lane of a warp may issue them, so lane 0 of the warp
    // moves the tiles of both of its envs and both groups wait on one mbarrier (the even group's slot)
    const int lane32 = threadIdx.x & 31, grp0 = grp & ~1;
    const uint32_t bar = smem_u32(tsm + (size_t)grp0 * p.mt_group + p.mt_bar);
    const float qf = (float)p.q_miss, kq = 0.25f * qf;
    const int hs = p.tile_hshift, H = 1 << hs;                                          // row pairs per tile
    const uint32_t tile_bytes = (uint32_t)H * 80u, tile_stride = (uint32_t)p.tile_stride;
    const uint64_t tm_ptr = reinterpret_cast<uint64_t>(&tmap);
    uint32_t phase = 0;

    if (lane32 == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(bar));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncwarp();
    for (int job = 0; job < 2; ++job) {
        const bool run = sensed && (job == 1 || (m1.w & 1u));
        const unsigned runmask = __ballot_sync(FULL, run);
        if (!runmask) continue;
        int nh = 0;
        if (run) {
            nh = map_job_load<GL>(p, e, job, lane, m0.y, xy, hit);
            for (int r = lane; r <= M + 1; r += GL) R[r] = 0ull;
        }
        __syncwarp();
        if (run) {
            for (int a = lane; a < n; a += GL) {
                int lo, hi;
                corner_span(xy[2 * a], p.R, p.R2, &lo, &hi);
                clo[a] = lo;
                const int i0 = max(0, lo - 1), i1 = min(M - 1, hi);      // cell i has corners i and i+1
                corner_span(xy[2 * a + 1], p.R, p.R2, &lo, &hi);
                const int j0 = max(0, lo - 1);
                // tile (a, par): the rows of parity par from the first such row >= i0, H of them; an agent whose box
                // is empty (injected position outside the map) gets tiles outside the tensor: zero fill, clipped store
                const int far = (i0 <= i1) ? 0 : M;
                box[a] = make_int4(j0 & ~3, ((i0 + (i0 & 1)) >> 1) + far, (M + j0) & ~3, (i0 >> 1) + far);
            }
        }
        __syncwarp();
        if (lane32 == 0) {
            const uint32_t ngrp = (runmask & 1u) + ((runmask >> GL) & 1u);
            asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(ngrp * 2u * (uint32_t)n * tile_bytes) : "memory");
            for (int g2 = 0; g2 < 32 / GL; ++g2) {
                if (!((runmask >> (g2 * GL)) & 1u)) continue;
                unsigned char* G2 = tsm + (size_t)(grp0 + g2) * p.mt_group;
                const int2* tc = reinterpret_cast<const int2*>(G2 + p.mt_box);
                const int e2 = blockIdx.x * kMapEnvsPerCta + grp0 + g2;
                uint32_t dst = smem_u32(G2);
                for (int t = 0; t < 2 * n; ++t, dst += tile_stride) {
                    const int2 c = tc[t];
                    asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];"
                                 ::"r"(dst), "l"(tm_ptr), "r"(c.x), "r"(c.y), "r"(e2), "r"(bar) : "memory");
                }
            }
        }
        // (1) corner-row intervals, while the tiles are in flight
        const int span = p.span_cap;                                     // power of two >= 2R
        for (int t = lane; t < n * span; t += GL) {
            if (!run) continue;
            const int a = t >> p.span_shift, r = t & (span - 1);
            const double axa = xy[2 * a], aya = xy[2 * a + 1];
            const int cx = clo[a] + r;
            const double dx = (double)cx - axa;
            const double A = dx * dx;
            if (!(A < p.R2) || cx < 0 || cx > M) continue;               // past the last corner row of this agent
            // candidate ends in fp32 (ay <= map_size: absolute error ~4e-6, far inside the 1e-3 guard band)
            const float wf = sqrtf(fmaxf((float)(p.R2 - A), 0.0f));
            const float ayf = (float)aya;
            const float yh = ayf + wf, yl = ayf - wf;
            float fh = floorf(yh), cl = ceilf(yl);
            if (yh - fh < 1e-3f || yh - fh > 1.0f - 1e-3f) {             // end within 1e-3 of an integer: decide exactly
                const double Y = (double)rintf(yh);
                fh = (float)(corner_pred(A, Y, aya, p.R2) ? Y : Y - 1.0);
            }
            if (cl - yl < 1e-3f || cl - yl > 1.0f - 1e-3f) {
                const double Y = (double)rintf(yl);
                cl = (float)(corner_pred(A, Y, aya, p.R2) ? Y : Y + 1.0);
            }
            const int yhi = min((int)fh, M), ylo = max((int)cl, 0);
            if (ylo <= yhi) {
                const unsigned long long mk = ((2ull << yhi) - 1ull) & ~((1ull << ylo) - 1ull);
                unsigned* w = reinterpret_cast<unsigned*>(&R[cx]);
                if ((unsigned)mk) atomicOr(w, (unsigned)mk);
                if ((unsigned)(mk >> 32)) atomicOr(w + 1, (unsigned)(mk >> 32));
            }
        }
        __syncwarp();
        mbar_wait(bar, phase);
        phase ^= 1u;
        // (2) sweep of the tiles in shared memory: one tile row (20 positions of one row pair) per lane and pass
        for (int q5 = lane; q5 < (n << (hs + 1)); q5 += GL) {
            if (!run) continue;
            const int2 tc = reinterpret_cast<const int2*>(box)[q5 >> hs];
            const int rp = tc.y + (q5 & (H - 1)), c0 = tc.x;             // row pair and first position of this tile row
            if (2 * rp >= M) continue;
            float4* rowp = reinterpret_cast<float4*>(G + (size_t)(q5 >> hs) * tile_stride + (size_t)(q5 & (H - 1)) * 80);
            // corner bits of the 20 positions as two strings PA (upper corner row of the cells) / PB (lower): cell at
            // offset u of the tile row uses bits u, u+1 -- or u+1, u+2 once past the end of the even row (s = M - c0),
            // where one bit is left out so that the last even-row cell and the first odd-row cell do not share a bit
            const unsigned long long R0 = R[2 * rp], R1 = R[2 * rp + 1], R2 = R[2 * rp + 2];   // corner rows of map rows 2rp, 2rp+1
            unsigned PA, PB;
            int s;
            if (c0 >= M) {
                PA = (unsigned)(R1 >> (c0 - M)); PB = (unsigned)(R2 >> (c0 - M)); s = 64;
            } else {
                s = M - c0;
                const unsigned keep = s < 31 ? (2u << s) - 1u : 0xffffffffu;             // corners c0 .. M of the even row
                const unsigned up = s < 31 ? s + 1 : 31;
                PA = ((unsigned)(R0 >> c0) & keep) | (s < 31 ? (unsigned)R1 << up : 0u);
                PB = ((unsigned)(R1 >> c0) & keep) | (s < 31 ? (unsigned)R2 << up : 0u);
            }
#pragma unroll
            for (int g = 0; g < 5; ++g) {
                if (c0 + 4 * g >= 2 * M) continue;
                const int sh = 4 * g + (4 * g >= s ? 1 : 0);
                const bool split = (4 * g + 2 == s);                     // two cells before, two after the end of the even row
                const unsigned AB = ((PA >> sh) & 0x7Fu) | (((PB >> sh) & 0x7Fu) << 16);
                const int pc0 = __popc(AB & 0x00030003u), pc1 = __popc(AB & 0x00060006u);
                const int pc2 = __popc(AB & (split ? 0x00180018u : 0x000C000Cu)), pc3 = __popc(AB & (split ? 0x00300030u : 0x00180018u));
                if (!(pc0 | pc1 | pc2 | pc3)) continue;                  // percent == 0 -> untouched (:285-286)
                float4 v = rowp[g];
                v.x = belief_cell(v.x, pc0, kq, qf);
                v.y = belief_cell(v.y, pc1, kq, qf);
                v.z = belief_cell(v.z, pc2, kq, qf);
                v.w = belief_cell(v.w, pc3, kq, qf);
                rowp[g] = v;
            }
        }
        __syncwarp();
        // (3) targets found by THIS call -> 1 where the cell was touched (:288-289), in every tile that holds the cell
        if (__any_sync(FULL, run && nh != 0)) {
            for (int h = lane; h < nh; h += GL) {
                if (!run || hit[h] < 0) continue;
                const int ci = hit[h] / M, cj = hit[h] - ci * M;
                if (!(((unsigned)(R[ci] >> cj) | (unsigned)(R[ci + 1] >> cj)) & 3u)) continue;
                const int x = (ci & 1) * M + cj, rp = ci >> 1;
                for (int t = 0; t < 2 * n; ++t) {
                    const int2 tc = reinterpret_cast<const int2*>(box)[t];
                    const unsigned dr = (unsigned)(rp - tc.y), dc = (unsigned)(x - tc.x);
                    if (dr < (unsigned)H && dc < 20u)
                        reinterpret_cast<float*>(G + (size_t)t * tile_stride)[dr * 20 + dc] = 1.0f;
                }
            }
        }
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");    // the tile writes above, before the TMA unit reads them
        __syncwarp();
        if (lane32 == 0) {
            for (int g2 = 0; g2 < 32 / GL; ++g2) {
                if (!((runmask >> (g2 * GL)) & 1u)) continue;
                unsigned char* G2 = tsm + (size_t)(grp0 + g2) * p.mt_group;
                const int2* tc = reinterpret_cast<const int2*>(G2 + p.mt_box);
                const int e2 = blockIdx.x * kMapEnvsPerCta + grp0 + g2;
                uint32_t src = smem_u32(G2);
                for (int t = 0; t < 2 * n; ++t, src += tile_stride) {
                    const int2 c = tc[t];
                    asm volatile("cp.async.bulk.tensor.3d.global.shared::cta.tile.bulk_group [%0, {%2, %3, %4}], [%1];"
                                 ::"l"(tm_ptr), "r"(src), "r"(c.x), "r"(c.y), "r"(e2) : "memory");
                }
            }
            asm volatile("cp.async.bulk.commit_group;" ::: "memory");
            asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");   // the next job (or the exit) may reuse the tiles
        }
        __syncwarp();
    }
}

// map_size > 63: per-cell corner tests (same results, more arithmetic), half-warp per map row, one warp per env
__global__ void __launch_bounds__(kMapThreads) flight_map_wide_kernel(const __grid_constant__ FlightParams p, uint32_t seq) {
    extern __shared__ __align__(16) unsigned long long msm[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int e = blockIdx.x * (kMapThreads / 32) + warp;
    if (e >= p.E) return;
    const int n = p.n, M = p.M;
    uint4 m0, m1;
    meta_ld(p, e, &m0, &m1);
    if ((m1.w >> 1) != seq) return;
    unsigned long long* S = msm + (size_t)warp * p.ms_warp;
    int4* box = reinterpret_cast<int4*>(S + p.ms_box);
    double* xy = reinterpret_cast<double*>(S + p.ms_xy);
    int* hit = reinterpret_cast<int*>(S + p.ms_hit);
    float* map = p.prob_map + (size_t)e * M * M;
    const float qf = (float)p.q_miss;
    const int colk = lane & 15, half = lane >> 4;
    unsigned touched = 0;
    for (int job = (m1.w & 1u) ? 0 : 1; job < 2; ++job) {
        const int nh = map_job_load<32>(p, e, job, lane, m0.y, xy, hit);
        __syncwarp();
        if (lane < n) {
            int lo, hi;
            corner_span(xy[2 * lane], p.R, p.R2, &lo, &hi);
            const int i0 = max(0, lo - 1), i1 = min(M - 1, hi);
            corner_span(xy[2 * lane + 1], p.R, p.R2, &lo, &hi);
            box[lane] = make_int4(i0, i1, max(0, lo - 1), min(M - 1, hi));
        }
        __syncwarp();
        for (int a = 0; a < n; ++a) {
            const int4 bx = box[a];
            for (int jc = bx.z; jc <= bx.w; jc += 16) {
                const int j = jc + colk;
                if (j > bx.w) continue;
                const double y0 = (double)j, y1 = (double)(j + 1);
                for (int i = bx.x + half; i <= bx.y; i += 2) {
                    bool mine = true;
                    for (int b = 0; b < a; ++b) {
                        const int4 bb = box[b];
                        mine &= !(i >= bb.x && i <= bb.y && j >= bb.z && j <= bb.w);
                    }
                    if (!mine) continue;
                    const double x0 = (double)i, x1 = (double)(i + 1);
                    uint32_t bits = 0;
                    for (int q = 0; q < n; ++q) {
                        const double qx = xy[2 * q], qy = xy[2 * q + 1];
                        const double dx0 = x0 - qx, dx1 = x1 - qx, dy0 = y0 - qy, dy1 = y1 - qy;
                        const double sx0 = dx0 * dx0, sx1 = dx1 * dx1, sy0 = dy0 * dy0, sy1 = dy1 * dy1;
                        bits |= (sx0 + sy0 < p.R2) ? 1u : 0u;
                        bits |= (sx1 + sy0 < p.R2) ? 2u : 0u;
                        bits |= (sx0 + sy1 < p.R2) ? 4u : 0u;
                        bits |= (sx1 + sy1 < p.R2) ? 8u : 0u;
                    }
                    if (!bits) continue;
                    ++touched;
                    const int cell = i * M + j;
                    float v = belief_update(map[cell], __popc(bits), qf);
                    for (int k = 0; k < nh; ++k)
                        if (hit[k] == cell) v = 1.0f;
                    map[cell] = v;
                }
            }
        }
        __syncwarp();
    }
    if (p.count_touched) {
        const unsigned tot = __reduce_add_sync(0xffffffffu, touched);
        if (lane == 0 && tot) atomicAdd(p.stats + CS_STAT_TOUCHED, (double)tot);
    }
}

enum { MODE_STEP = 0, MODE_RESET = 1 };

// ------------------------------------------------------------------------------------------------
// The step / reset kernel.  A group of LPE lanes (LPE = power of two >= max(n_agents, target_num)) owns one env:
// lane a < n holds agent a (x, y, heading, cos, sin) and lane j < m holds target j, all in REGISTERS; positions
// travel between lanes by warp shuffles.  Every global load is issued up front; nothing is staged through shared
// memory except the heading-table index (per warp) and the scratch of the warp-wide belief-map pass.
// Every warp collective uses the FULL mask and sits in warp-uniform control flow (groups are told apart by
// predicates, not branches), so no partial-mask MATCH/REDUX sequences or divergence barriers are generated.
//   pass 0 (STEP): _agent_step -> _update_obs -> step bookkeeping            (flight_env_easy.py:255-314)
//   pass 1       : reset (selected envs in RESET mode; just-terminated envs under auto_reset) -> _update_obs (:79-182)
// actions == nullptr in MODE_STEP: uniform-random policy drawn in-kernel (alg=random, agent/agent.py:34-36)
// ------------------------------------------------------------------------------------------------
template <int LPE, int MODE, bool MAP>
__global__ void __launch_bounds__(kThreads, 8) flight_kernel(const __grid_constant__ FlightParams p, const uint8_t* __restrict__ actions,
                                                             const uint8_t* __restrict__ mask, uint32_t rflags, uint32_t seq) {
    constexpr int EPW = 32 / LPE;
    constexpr unsigned FULL = 0xffffffffu;
    constexpr unsigned GBITS = (LPE == 32) ? 0xffffffffu : ((1u << (LPE & 31)) - 1u);
    extern __shared__ __align__(16) double smem[];

    const int tid = threadIdx.x, warp = tid >> 5, lane32 = tid & 31;
    const int wenv0 = (blockIdx.x * (kThreads / 32) + warp) * EPW;      // first env of this warp
    const int wcnt = min(EPW, p.E - wenv0);
    if (wcnt <= 0) return;                                               // whole warp idle (nothing is block-synchronised)
    double* W = smem + (size_t)warp * p.s_warp;
    longlong2* lutm = reinterpret_cast<longlong2*>(W + p.s_lut);
    const int n = p.n, m = p.m;
    const int g = lane32 / LPE, lane = lane32 % LPE, gbase = lane32 - lane;
    const bool active = g < wcnt;
    const int e = wenv0 + (active ? g : 0);
    const bool is_agent = active && lane < n, is_tgt = active && lane < m;
    const uint32_t env_id = p.env_id_base + (uint32_t)e;
#define GSHFL(v, q) __shfl_sync(FULL, (v), gbase + (q))
#define GBALLOT(pred) ((__ballot_sync(FULL, (pred)) >> gbase) & GBITS)

    // ---- every load of the step, issued before anything is consumed ----------------------------------------
    double ax = 0.0, ay = 0.0, yaw = 0.0, tx = 0.0, ty = 0.0;
    uint4 m0 = make_uint4(0, 0, 0, 0), m1 = make_uint4(0, 0, 0, 0);
    int act = 0;
    if (is_agent) {
        const double2 v = xy_ld(p, lane, e);
        ax = v.x; ay = v.y;
        yaw = *dyn_at(p, p.yaw_off + lane, e);
        if (MODE == MODE_STEP && actions != nullptr) act = actions[(size_t)e * n + lane];
    }
    if (is_tgt) {
        const double2 v = tgt_ld(p, lane, e);
        tx = v.x; ty = v.y;
    }
    if (active) {
        meta_ld(p, e, &m0, &m1);                                               // same addresses for the whole group
    }
    if (MODE == MODE_STEP) {
        const longlong2 v0 = __ldg(p.lut_meta + lane32);
        longlong2 v1 = make_longlong2(0, 0);
        if (lane32 + 32 < 37) v1 = __ldg(p.lut_meta + lane32 + 32);
        lutm[lane32] = v0;
        if (lane32 + 32 < 37) lutm[lane32 + 32] = v1;
        __syncwarp();
    }
    uint32_t found = m0.x, newf_last = m0.y, outmask = m0.z, time_step = m0.w;
    uint32_t episode = m1.x, flags = m1.y;
    float ep_reward = __uint_as_float(m1.z);

    double c_h = 0.0, s_h = 0.0;                 // cos/sin of the heading (agent lanes), for the fp32 outputs
    bool done = (flags & CS_FLAG_DONE) != 0;
    bool do_sense = false, emit = false, state_full = false, tgt_dirty = false, have_result = false;
    float res_reward = 0.f;
    uint32_t res_term = 0, res_win = 0, res_found = 0, t_key = 0;   // what step()/reset() report for this env
    float st_eps = 0.f, st_rew = 0.f, st_found = 0.f, st_wins = 0.f, st_len = 0.f;
    uint32_t sense_word = m1.w, prejob = 0;     // CS_META_SENSE: (call number << 1) | job parked in `pre`

    // ---- _agent_step -------------------------------------------------------------------------------------------
    if (MODE == MODE_STEP) {
        const bool stepping = active && !done;
        if (stepping && is_agent) {
            if (actions == nullptr) {
                // one Philox block serves 4 agents; action = word % 3 (np.random.randint(0, 3), agent.py:36)
                const cs_u4 w = cs_philox4x32_10(env_id, ((episode & 0xFFFFu) << 16) | ((time_step + 1u) & 0xFFFFu),
                                                 (uint32_t)(lane >> 2), 0u, p.seed, CS_STREAM_POLICY);
                act = (int)(cs_word(w, lane & 3) % 3u);
            }
            double h = yaw + ((act == 1) ? p.turn : ((act == 2) ? -p.turn : 0.0));   // dyaw = [0, pi/18, -pi/18] (:259-262)
            if (h > p.two_pi) h -= p.two_pi;                                          // strict tests (:263-266)
            else if (h < 0.0) h += p.two_pi;
            heading_sincos(p, lutm, h, &s_h, &c_h);
            yaw = h;
        }
        // Can any repulsion term be non-zero this step?  If every pair of OLD positions is farther apart than
        // force_dist + |v| (with slack), no agent receives a force, every displacement is <= |v|, and by
        // induction over the sequential update order no later agent does either.
        bool close = false;
        for (int q = 1; q < n; ++q) {
            const double xq = GSHFL(ax, q), yq = GSHFL(ay, q);
            const double dx = xq - ax, dy = yq - ay;
            close |= (q > lane && dx * dx + dy * dy < p.near2);
        }
        const bool close_g = GBALLOT(close && stepping && is_agent) != 0u;     // some pair of this env is close
        bool outside = false;
        if (stepping && is_agent && !close_g) {
            ax = ax + p.v * c_h;                          // x += v*cos(yaw)   (:267-268)
            ay = ay + p.v * s_h;
            outside = wall_reg(p, ax, ay, yaw, c_h);
        }
        if (__any_sync(FULL, close_g)) {
            // Repulsion path: the reference's sequential, in-place update (:271,:293-301) -- agent k sees its own
            // OLD position and the already-moved j<k.  Lanes keep their CURRENT position in registers, so iterating
            // k = 0..n-1 and letting every lane q evaluate its term against agent k's old position reproduces
            // exactly that; the terms of the (few) lanes in range are added on lane k in ascending q, the
            // reference's summation order.
            for (int k = 0; k < n; ++k) {
                const double x0 = GSHFL(ax, k), y0 = GSHFL(ay, k);
                const double dxq = ax - x0, dyq = ay - y0;
                const bool inr = close_g && is_agent && lane != k && (dxq * dxq + dyq * dyq < p.fd2) && (ax != x0 || ay != y0);
                uint32_t near_all = __ballot_sync(FULL, inr);
                double tfx = 0.0, tfy = 0.0;
                if (inr) {
                    const double ex = x0 - ax, ey = y0 - ay;
                    const double r2 = ex * ex + ey * ey;
                    tfx = p.fk * ex / r2;
                    tfy = p.fk * ey / r2;
                }
                double fx = 0.0, fy = 0.0;
                while (near_all) {                                          // warp-uniform; ascending lane = ascending q
                    const int src = __ffs(near_all) - 1;
                    near_all &= near_all - 1;
                    const double vx = __shfl_sync(FULL, tfx, src), vy = __shfl_sync(FULL, tfy, src);
                    if ((src & ~(LPE - 1)) == gbase) { fx += vx; fy += vy; }
                }
                if (close_g && is_agent && lane == k) {
                    ax = ax + p.v * c_h;
                    ay = ay + p.v * s_h;
                    ax += fx;
                    ay += fy;
                    outside = wall_reg(p, ax, ay, yaw, c_h);
                }
            }
        }
        if (stepping) {
            do_sense = true;
            t_key = time_step + 1u;
        } else if (active) {
            have_result = true;                            // masked no-op on a finished env
            res_reward = 0.f;
            res_term = 1;
            res_win = flags & CS_FLAG_WIN;
            res_found = (uint32_t)__popc(found);
        }
        const uint32_t ob = GBALLOT(outside);
        if (stepping) outmask = ob;
    }

    for (int pass = 0; pass < 2; ++pass) {
        if (pass == 1) {
            const bool do_reset = active && ((MODE == MODE_RESET) ? (mask == nullptr || mask[e] != 0) : (p.auto_reset && done));
            if (!__any_sync(FULL, do_reset)) break;
            do_sense = do_reset;
            if (do_reset) {                                                   // reset (:79-180)
                episode += (rflags & CS_RESET_KEEP_EPISODE) ? 0u : 1u;
                found = 0; outmask = 0; time_step = 0; flags = 0; ep_reward = 0.f; done = false;
                if (!(rflags & CS_RESET_KEEP_TARGETS)) {
                    if (is_tgt) {
                        const double2 t = draw_target(p, env_id, episode, lane);
                        tx = t.x; ty = t.y;
                    }
                    tgt_dirty = true;
                }
                if (is_agent) {
                    const double lin = p.lin[lane];                                       // (:140-143)
                    switch (p.agent_mode) {
                        case 0: ax = lin; ay = 0.0; yaw = p.half_pi; break;
                        case 1: ax = lin; ay = p.Md / 2.0; yaw = p.half_pi; break;
                        case 2: ax = 0.0; ay = lin; yaw = 0.0; break;
                        default: ax = p.Md; ay = lin; yaw = p.pi; break;
                    }
                    c_h = p.cos0; s_h = p.sin0;
                }
                if (MAP && (rflags & CS_RESET_INIT)) {
                    float* map = p.prob_map + (size_t)e * p.M * p.M;
                    for (int c = lane; c < p.M * p.M; c += LPE) map[c] = 0.5f;            // flight_env.py:84-86
                }
                t_key = 0;
                emit = true;
                state_full = true;
                if (MODE == MODE_RESET) { have_result = true; res_reward = 0.f; res_term = 0; }
            }
        } else if (MODE == MODE_RESET) {
            continue;
        }

        // ---- _update_obs: detection + reward + win (:223-253) ----------------------------------------------
        uint32_t amask = 0;
        for (int q = 0; q < n; ++q) {
            const double xq = GSHFL(ax, q), yq = GSHFL(ay, q);
            const double dx = tx - xq, dy = ty - yq;
            if (dx * dx + dy * dy <= p.R2) amask |= 1u << q;                   // '<=' (:237)
        }
        bool got = false;
        if (do_sense && is_tgt && amask && !((found >> lane) & 1u)) {          // draw is irrelevant once found (:239)
            for (int blk = 0; 4 * blk < n && !got; ++blk) {
                const uint32_t bits = (amask >> (4 * blk)) & 0xFu;
                if (!bits) continue;
                const cs_u4 w = cs_detect_words(p.seed, env_id, episode, t_key, (uint32_t)blk, (uint32_t)lane);
                got = ((bits & 1u) && (long long)w.x <= p.thr) || ((bits & 2u) && (long long)w.y <= p.thr) ||
                      ((bits & 4u) && (long long)w.z <= p.thr) || ((bits & 8u) && (long long)w.w <= p.thr);
            }
        }
        const uint32_t newf = GBALLOT(got);
        int rew = 0;
        if (do_sense) {
            found |= newf;
            newf_last = newf;
            const int c = __popc(newf);
            rew = -1 + 10 * c;                                                 // MOVE_COST + FIND_ONE_TGT (:228,:241)
            if (c > 0 && __popc(found) == m && !(flags & CS_FLAG_WIN)) {
                rew += 100;                                                    // FIND_ALL_TGT (:244-246)
                flags |= CS_FLAG_WIN;
            }
            rew -= __popc(outmask);                                            // OUT_PUNISH per agent outside (:249-250)
        }
        if (MODE == MODE_RESET && do_sense) {
            res_win = flags & CS_FLAG_WIN;
            res_found = (uint32_t)__popc(found);
        }
        if (pass == 0 && do_sense) {                                           // step bookkeeping (:308-314)
            time_step += 1u;
            ep_reward += (float)rew;
            const int nfound = __popc(found);
            const bool term = (nfound >= m) || ((int)time_step >= p.T);
            if (term) flags |= CS_FLAG_DONE;
            done = term;
            emit = true;
            have_result = true;
            res_reward = (float)rew;
            res_term = term ? 1u : 0u;
            res_win = flags & CS_FLAG_WIN;          // of the episode this step belongs to, also when auto_reset follows
            res_found = (uint32_t)nfound;
            if (lane == 0 && term) {
                st_eps = 1.f; st_rew = ep_reward; st_found = (float)nfound; st_wins = (flags & CS_FLAG_WIN) ? 1.f : 0.f;
                st_len = (float)time_step;
            }
        }
        // ---- belief map (flight_env.py:266,:275-303): a job for flight_map_kernel, which runs next on the stream.
        //      Normally the job IS the state record (positions, CS_META_NEWFOUND); an env about to be reset inside this
        //      call parks the job of its last step in the side buffer, because pass 1 overwrites the record.
        if (MAP && do_sense) {
            sense_word = seq << 1;
            if (pass == 0 && done && p.auto_reset) {
                double* pj = p.pre + (size_t)e * p.pre_stride;
                int* ph = reinterpret_cast<int*>(pj + 2 * n);
                if (is_agent) *reinterpret_cast<double2*>(pj + 2 * lane) = make_double2(ax, ay);
                if (lane == 0) ph[0] = __popc(newf);
                if (is_tgt && ((newf >> lane) & 1u)) ph[1 + __popc(newf & ((1u << lane) - 1u))] = hit_cell(p, tx, ty);
                prejob = 1u;
            }
        }
    }
#undef GSHFL
#undef GBALLOT

    // ---- outputs, straight from registers -------------------------------------------------------------------
    if (active && emit) {
        if (is_agent) {
            xy_st(p, lane, e, ax, ay);
            *dyn_at(p, p.yaw_off + lane, e) = yaw;
            // get_obs row = agent part of get_state (flight_env_easy.py:218-221, :192-193)
            const float4 o = make_float4((float)((ax - p.half_M) * p.inv_half), (float)((ay - p.half_M) * p.inv_half),
                                         (float)c_h, (float)s_h);
            reinterpret_cast<float4*>(p.obs)[(size_t)e * n + lane] = o;
            reinterpret_cast<float4*>(p.state + (size_t)e * p.state_stride)[lane] = o;
        }
        if (lane == 0) {
            meta_st(p, e, make_uint4(found, newf_last, outmask, time_step),
                    make_uint4(episode, flags, __float_as_uint(ep_reward), MAP ? (sense_word | prejob) : 0u));
        }
        if (is_tgt) {
            // target part of the state row (:201-211): rewritten in full after a reset, otherwise only the 'find'
            // entry of a target found by this call
            float* srow = p.state + (size_t)e * p.state_stride + 4 * n + 3 * lane;
            if (state_full) {
                srow[0] = (float)((tx - p.half_M) * p.inv_half);
                srow[1] = (float)((ty - p.half_M) * p.inv_half);
                srow[2] = ((found >> lane) & 1u) ? 1.0f : 0.0f;
            } else if ((newf_last >> lane) & 1u) {
                srow[2] = 1.0f;
            }
            if (tgt_dirty) tgt_st(p, lane, e, make_double2(tx, ty));
        }
    }
    if (MAP && active && !emit && lane == 0 && m1.w != 0u)
        meta_at(p, 3, e)->y = 0u;      // not sensed by THIS call: no job for the map kernel
    if (active && have_result && lane == 0) {
        p.reward[e] = res_reward;
        p.terminated[e] = (uint8_t)res_term;
        p.win[e] = res_win ? 1 : 0;
        p.target_find[e] = (int32_t)res_found;
    }
    // ---- statistics of episodes that ended in this call: warp reduction, then one atomic per statistic per warp.
    //      Nothing is accumulated for ordinary steps (a same-address atomic per warp per step serialises in L2 and
    //      was the floor of the small configurations); env_steps = sum of finished episode lengths + live time_steps
    //      is assembled by cs_flight_stats.
    if (MODE == MODE_STEP) {
        if (__any_sync(FULL, st_eps != 0.f)) {
            for (int o = 16; o > 0; o >>= 1) {
                st_eps += __shfl_xor_sync(FULL, st_eps, o);
                st_rew += __shfl_xor_sync(FULL, st_rew, o);
                st_found += __shfl_xor_sync(FULL, st_found, o);
                st_wins += __shfl_xor_sync(FULL, st_wins, o);
                st_len += __shfl_xor_sync(FULL, st_len, o);
            }
            if (lane32 == 0) {
                atomicAdd(p.stats + CS_STAT_EPISODES, (double)st_eps);
                atomicAdd(p.stats + CS_STAT_EP_REWARD, (double)st_rew);
                atomicAdd(p.stats + CS_STAT_TARGETS_FOUND, (double)st_found);
                atomicAdd(p.stats + CS_STAT_WINS, (double)st_wins);
                atomicAdd(p.stats + CS_STAT_EP_LEN, (double)st_len);
            }
        }
    }
}

// ------------------------------------------------------------------------------------------------
// The step / reset kernel, THREAD-PER-ENV form (n_agents <= kTpeMaxAgents; the default).
//
// flight_kernel above spends ~380 warp instructions per env-step with most lanes idle in the agent phases and every
// reduction a shuffle or a ballot: on B200 that is issue-bound at 14 % of the HBM roofline.  Here one THREAD owns one
// env: agent state lives in registers (arrays indexed by compile-time constants), the targets stream through from
// global memory, the agents x targets test and the reward are plain loops, and nothing is exchanged between lanes --
// a warp advances 32 envs per instruction.  K > 1 splits the long part, the loop over the targets, over K threads
// that each repeat the (short) agent phase: the same latency as flight_kernel with 32/K envs per warp, for launches too
// small to fill the GPU with one thread per env.  Only the rare heavy pieces are warp-cooperative: the Box-Muller target
// redraw of a reset env (one lane per target) and the 0.5 fill of a belief map.
// Same arithmetic, same operation order, same Philox counters as flight_kernel: the two are interchangeable bit for
// bit (tests/test_gpu_flight_easy.py compares them).
//   pass 0 (STEP): _agent_step -> _update_obs -> step bookkeeping            (flight_env_easy.py:255-314)
//   pass 1       : reset (selected envs in RESET mode; just-terminated envs under auto_reset) -> _update_obs (:79-182)
// ------------------------------------------------------------------------------------------------
#ifndef CS_TPE_THREADS
#define CS_TPE_THREADS 64
#endif
constexpr int kTpeThreads = CS_TPE_THREADS;
#ifndef CS_TPE_TU
#define CS_TPE_TU 5            // targets a thread loads together in the sensing loop
#endif
#ifndef CS_TPE_MIN_CTAS
#define CS_TPE_MIN_CTAS 8      // resident CTAs per SM the thread-per-env kernel is compiled for (register budget 65536/(64*N))
#endif
constexpr int kTpeMaxAgents = 8;

template <int N, int K, int MODE, bool MAP>
__device__ __forceinline__ void flight_tpe_body(const FlightParams& p, const int block, const uint8_t* __restrict__ actions,
                                                const uint8_t* __restrict__ mask, uint32_t rflags, uint32_t seq) {
    constexpr unsigned FULL = 0xffffffffu;
    __shared__ longlong2 lutm[40];
    static_assert(K == 1 || K == 4, "K");
    using L = Lay<K == 1>;                                              // one thread per env <-> structure of arrays (cs_flight_create)
    constexpr uint32_t MINE = 0xFFFFFFFFu / ((1u << K) - 1u);          // targets j with j % K == 0
    const int tid = threadIdx.x, lane32 = tid & 31, kk = tid % K;       // kk: which of the env's K threads this is
    const int e_raw = (block * kTpeThreads + tid) / K;
    const bool active = e_raw < p.E;
    const int e = active ? e_raw : p.E - 1;
    const int m = p.m;
    const uint32_t env_id = p.env_id_base + (uint32_t)e;
    if (MODE == MODE_STEP) {
        if (tid < 37) lutm[tid] = __ldg(p.lut_meta + tid);
        __syncthreads();
    }

    // ---- state of this env ---------------------------------------------------------------------------------
    if (MODE == MODE_STEP) {
        // the targets are needed after the agent phase: start pulling this warp's target rows into L1 now, together with
        // the state loads below, so that the sensing loop pays no further HBM round trip
        if (K == 1) {
            const double* tp = p.tgt + e;
            for (int r = 0; r < 2 * m; ++r) asm volatile("prefetch.global.L1 [%0];" ::"l"(tp + (size_t)r * p.E));
        } else {
            const char* tb = reinterpret_cast<const char*>(p.tgt + (size_t)e * m * 2);
            for (int off = 128 * kk; off < m * 16; off += 128 * K) asm volatile("prefetch.global.L1 [%0];" ::"l"(tb + off));
            if (kk == K - 1) asm volatile("prefetch.global.L1 [%0];" ::"l"(tb + m * 16 - 1));
        }
    }
    double ax[N], ay[N], yaw[N], c_h[N], s_h[N];
#pragma unroll
    for (int a = 0; a < N; ++a) {
        const double2 v = L::xy_ld(p, a, e);
        ax[a] = v.x; ay[a] = v.y;
        yaw[a] = *L::yaw_at(p, a, e);
        c_h[a] = 0.0; s_h[a] = 0.0;
    }
    uint4 m0, m1;
    L::meta_ld(p, e, &m0, &m1);
    uint32_t found = m0.x, newf_last = m0.y, outmask = m0.z, time_step = m0.w;
    uint32_t episode = m1.x, flags = m1.y;
    float ep_reward = __uint_as_float(m1.z);
    uint32_t sense_word = m1.w, prejob = 0;     // CS_META_SENSE: (call number << 1) | job parked in `pre`

    bool done = (flags & CS_FLAG_DONE) != 0;
    bool do_sense = false, emit = false, state_full = false, have_result = false;
    float res_reward = 0.f;
    uint32_t res_term = 0, res_win = 0, res_found = 0, t_key = 0;   // what step()/reset() report for this env
    float st_eps = 0.f, st_rew = 0.f, st_found = 0.f, st_wins = 0.f, st_len = 0.f;

    // ---- _agent_step -------------------------------------------------------------------------------------------
    if (MODE == MODE_STEP) {
        const bool stepping = active && !done;
        if (stepping) {
            cs_u4 pw = {0u, 0u, 0u, 0u};
#pragma unroll
            for (int a = 0; a < N; ++a) {
                int act;
                if (actions != nullptr) {
                    act = actions[(size_t)e * N + a];
                } else {
                    // one Philox block serves 4 agents; action = word % 3 (np.random.randint(0, 3), agent.py:36)
                    if ((a & 3) == 0)
                        pw = cs_philox4x32_10(env_id, ((episode & 0xFFFFu) << 16) | ((time_step + 1u) & 0xFFFFu),
                                              (uint32_t)(a >> 2), 0u, p.seed, CS_STREAM_POLICY);
                    act = (int)(cs_word(pw, a & 3) % 3u);
                }
                double h = yaw[a] + ((act == 1) ? p.turn : ((act == 2) ? -p.turn : 0.0));   // dyaw = [0, pi/18, -pi/18] (:259-262)
                if (h > p.two_pi) h -= p.two_pi;                                             // strict tests (:263-266)
                else if (h < 0.0) h += p.two_pi;
                double sn, cs;
                heading_sincos(p, lutm, h, &sn, &cs);
                s_h[a] = sn; c_h[a] = cs;
                yaw[a] = h;
            }
            // Can any repulsion term be non-zero this step?  (see flight_kernel / DESIGN.md 4.2)
            bool close = false;
#pragma unroll
            for (int a = 0; a < N; ++a)
#pragma unroll
                for (int q = a + 1; q < N; ++q) {
                    const double dx = ax[q] - ax[a], dy = ay[q] - ay[a];
                    close |= (dx * dx + dy * dy < p.near2);
                }
            uint32_t ob = 0;
            if (!close) {
#pragma unroll
                for (int a = 0; a < N; ++a) {
                    ax[a] = ax[a] + p.v * c_h[a];                 // x += v*cos(yaw)   (:267-268)
                    ay[a] = ay[a] + p.v * s_h[a];
                    if (wall_reg(p, ax[a], ay[a], yaw[a], c_h[a])) ob |= 1u << a;
                }
            } else {
                // the reference's sequential, in-place update (:271,:293-301): agent k sees its own OLD position and the
                // already-moved q < k; the terms are added in ascending q
#pragma unroll
                for (int k = 0; k < N; ++k) {
                    const double x0 = ax[k], y0 = ay[k];
                    double fx = 0.0, fy = 0.0;
#pragma unroll
                    for (int q = 0; q < N; ++q) {
                        if (q == k) continue;
                        const double dxq = ax[q] - x0, dyq = ay[q] - y0;
                        if ((dxq * dxq + dyq * dyq < p.fd2) && (ax[q] != x0 || ay[q] != y0)) {
                            const double ex = x0 - ax[q], ey = y0 - ay[q];
                            const double r2 = ex * ex + ey * ey;
                            fx += p.fk * ex / r2;
                            fy += p.fk * ey / r2;
                        }
                    }
                    double nx = x0 + p.v * c_h[k], ny = y0 + p.v * s_h[k];
                    nx += fx;
                    ny += fy;
                    ax[k] = nx; ay[k] = ny;
                    if (wall_reg(p, ax[k], ay[k], yaw[k], c_h[k])) ob |= 1u << k;
                }
            }
            outmask = ob;
            do_sense = true;
            t_key = time_step + 1u;
        } else if (active) {
            have_result = true;                            // masked no-op on a finished env
            res_reward = 0.f;
            res_term = 1;
            res_win = flags & CS_FLAG_WIN;
            res_found = (uint32_t)__popc(found);
        }
    }

    for (int pass = 0; pass < 2; ++pass) {
        if (pass == 1) {
            const bool do_reset = active && ((MODE == MODE_RESET) ? (mask == nullptr || mask[e] != 0) : (p.auto_reset && done));
            const unsigned rmask = __ballot_sync(FULL, do_reset && kk == 0);
            if (!rmask) break;
            do_sense = do_reset;
            if (do_reset) {                                                   // reset (:79-180)
                episode += (rflags & CS_RESET_KEEP_EPISODE) ? 0u : 1u;
                found = 0; outmask = 0; time_step = 0; flags = 0; ep_reward = 0.f; done = false;
#pragma unroll
                for (int a = 0; a < N; ++a) {
                    const double lin = p.lin[a];                                          // (:140-143)
                    switch (p.agent_mode) {
                        case 0: ax[a] = lin; ay[a] = 0.0; yaw[a] = p.half_pi; break;
                        case 1: ax[a] = lin; ay[a] = p.Md / 2.0; yaw[a] = p.half_pi; break;
                        case 2: ax[a] = 0.0; ay[a] = lin; yaw[a] = 0.0; break;
                        default: ax[a] = p.Md; ay[a] = lin; yaw[a] = p.pi; break;
                    }
                    c_h[a] = p.cos0; s_h[a] = p.sin0;
                }
                t_key = 0;
                emit = true;
                state_full = true;
                if (MODE == MODE_RESET) { have_result = true; res_reward = 0.f; res_term = 0; }
            }
            // warp-cooperative heavy parts of a reset, one resetting env at a time: its targets (one lane per
            // target, :95-127) and, for reset(init=True), its belief map (flight_env.py:84-86)
            if (!(rflags & CS_RESET_KEEP_TARGETS) || (MAP && (rflags & CS_RESET_INIT))) {
                unsigned left = rmask;
                const int t_first = block * kTpeThreads + (tid & ~31);
                while (left) {
                    const int src = __ffs(left) - 1;
                    left &= left - 1;
                    const int es = (t_first + src) / K;
                    const uint32_t ep_s = __shfl_sync(FULL, episode, src);
                    if (!(rflags & CS_RESET_KEEP_TARGETS)) {
                        for (int j = lane32; j < m; j += 32) {
                            const double2 t = draw_target(p, p.env_id_base + (uint32_t)es, ep_s, j);
                            tgt_st(p, j, es, t);
                        }
                    }
                    if (MAP && (rflags & CS_RESET_INIT)) {
                        float* map = p.prob_map + (size_t)es * p.M * p.M;
                        for (int c = lane32; c < p.M * p.M; c += 32) map[c] = 0.5f;
                    }
                }
                __syncwarp();                                   // the owners read the new targets back below
            }
        } else if (MODE == MODE_RESET) {
            continue;
        }

        // ---- _update_obs: detection + reward + win (:223-253) ----------------------------------------------
        uint32_t newf = 0;
        if (do_sense) {
            // the env's K threads share the targets; TU of a thread's targets are loaded together so that their
            // latency is paid once per block, not once per target
            constexpr int TU = CS_TPE_TU;
            for (int j0 = kk; j0 < m; j0 += K * TU) {
                double2 t[TU];
#pragma unroll
                for (int u = 0; u < TU; ++u) {
                    const int j = min(j0 + u * K, m - 1);
                    t[u] = L::tgt_ld(p, j, e);
                }
#pragma unroll
                for (int u = 0; u < TU; ++u) {
                    const int j = j0 + u * K;
                    if (j >= m) break;
                    uint32_t amask = 0;
#pragma unroll
                    for (int a = 0; a < N; ++a) {
                        const double dx = t[u].x - ax[a], dy = t[u].y - ay[a];
                        if (dx * dx + dy * dy <= p.R2) amask |= 1u << a;               // '<=' (:237)
                    }
                    if (amask && !((found >> j) & 1u)) {                               // draw is irrelevant once found (:239)
                        bool got = false;
#pragma unroll
                        for (int blk = 0; 4 * blk < N; ++blk) {
                            const uint32_t bits = (amask >> (4 * blk)) & 0xFu;
                            if (!bits || got) continue;
                            const cs_u4 w = cs_detect_words(p.seed, env_id, episode, t_key, (uint32_t)blk, (uint32_t)j);
                            got = ((bits & 1u) && (long long)w.x <= p.thr) || ((bits & 2u) && (long long)w.y <= p.thr) ||
                                  ((bits & 4u) && (long long)w.z <= p.thr) || ((bits & 8u) && (long long)w.w <= p.thr);
                        }
                        if (got) newf |= 1u << j;
                    }
                }
            }
        }
#pragma unroll
        for (int o = K / 2; o > 0; o >>= 1) newf |= __shfl_xor_sync(FULL, newf, o);
        int rew = 0;
        if (do_sense) {
            found |= newf;
            newf_last = newf;
            const int c = __popc(newf);
            rew = -1 + 10 * c;                                                 // MOVE_COST + FIND_ONE_TGT (:228,:241)
            if (c > 0 && __popc(found) == m && !(flags & CS_FLAG_WIN)) {
                rew += 100;                                                    // FIND_ALL_TGT (:244-246)
                flags |= CS_FLAG_WIN;
            }
            rew -= __popc(outmask);                                            // OUT_PUNISH per agent outside (:249-250)
        }
        if (MODE == MODE_RESET && do_sense) {
            res_win = flags & CS_FLAG_WIN;
            res_found = (uint32_t)__popc(found);
        }
        if (pass == 0 && do_sense) {                                           // step bookkeeping (:308-314)
            time_step += 1u;
            ep_reward += (float)rew;
            const int nfound = __popc(found);
            const bool term = (nfound >= m) || ((int)time_step >= p.T);
            if (term) flags |= CS_FLAG_DONE;
            done = term;
            emit = true;
            have_result = true;
            res_reward = (float)rew;
            res_term = term ? 1u : 0u;
            res_win = flags & CS_FLAG_WIN;          // of the episode this step belongs to, also when auto_reset follows
            res_found = (uint32_t)nfound;
            if (term && kk == 0) {
                st_eps = 1.f; st_rew = ep_reward; st_found = (float)nfound; st_wins = (flags & CS_FLAG_WIN) ? 1.f : 0.f;
                st_len = (float)time_step;
            }
        }
        // ---- belief map job for flight_map_kernel (see flight_kernel) -------------------------------------------
        if (MAP && do_sense) {
            sense_word = seq << 1;
            if (pass == 0 && done && p.auto_reset && kk == 0) {
                double* pj = p.pre + (size_t)e * p.pre_stride;
                int* ph = reinterpret_cast<int*>(pj + 2 * N);
#pragma unroll
                for (int a = 0; a < N; ++a) *reinterpret_cast<double2*>(pj + 2 * a) = make_double2(ax[a], ay[a]);
                int k = 0;
                for (uint32_t left = newf; left; left &= left - 1) {
                    const int j = __ffs(left) - 1;
                    const double2 t = L::tgt_ld(p, j, e);
                    ph[1 + k++] = hit_cell(p, t.x, t.y);
                }
                ph[0] = k;
            }
            if (pass == 0 && done && p.auto_reset) prejob = 1u;
        }
    }

    // ---- outputs ---------------------------------------------------------------------------------------------
    if (active && emit) {
        float* srow = p.state + (size_t)e * p.state_stride;
#pragma unroll
        for (int a = 0; a < N; ++a) {
            if (a % K != kk) continue;                                  // the env's K threads share the rows
            L::xy_st(p, a, e, ax[a], ay[a]);
            *L::yaw_at(p, a, e) = yaw[a];
            // get_obs row = agent part of get_state (flight_env_easy.py:218-221, :192-193)
            const float4 o = make_float4((float)((ax[a] - p.half_M) * p.inv_half), (float)((ay[a] - p.half_M) * p.inv_half),
                                         (float)c_h[a], (float)s_h[a]);
            reinterpret_cast<float4*>(p.obs)[(size_t)e * N + a] = o;
            reinterpret_cast<float4*>(srow)[a] = o;
        }
        if (kk == 0) {
            L::meta_st(p, e, make_uint4(found, newf_last, outmask, time_step),
                    make_uint4(episode, flags, __float_as_uint(ep_reward), MAP ? (sense_word | prejob) : 0u));
        }
        // target part of the state row (:201-211): rewritten in full after a reset, otherwise only the 'find' entry of
        // a target found by this call
        if (state_full) {
            for (int j = kk; j < m; j += K) {
                const double2 t = L::tgt_ld(p, j, e);
                float* s3 = srow + 4 * N + 3 * j;
                const float nx = (float)((t.x - p.half_M) * p.inv_half), ny = (float)((t.y - p.half_M) * p.inv_half);
                const float fj = ((found >> j) & 1u) ? 1.0f : 0.0f;
                s3[0] = nx; s3[1] = ny; s3[2] = fj;
            }
        } else {
            for (uint32_t left = newf_last & (MINE << kk); left; left &= left - 1) srow[4 * N + 3 * (__ffs(left) - 1) + 2] = 1.0f;
        }
    }
    if (MAP && active && !emit && kk == 0 && m1.w != 0u)
        meta_at(p, 3, e)->y = 0u;      // not sensed by THIS call: no job for the map kernel
    if (active && have_result && kk == 0) {
        p.reward[e] = res_reward;
        p.terminated[e] = (uint8_t)res_term;
        p.win[e] = res_win ? 1 : 0;
        p.target_find[e] = (int32_t)res_found;
    }
    // ---- statistics of episodes that ended in this call: warp reduction, then one atomic per statistic per warp
    if (MODE == MODE_STEP) {
        if (__any_sync(FULL, st_eps != 0.f)) {
            for (int o = 16; o > 0; o >>= 1) {
                st_eps += __shfl_xor_sync(FULL, st_eps, o);
                st_rew += __shfl_xor_sync(FULL, st_rew, o);
                st_found += __shfl_xor_sync(FULL, st_found, o);
                st_wins += __shfl_xor_sync(FULL, st_wins, o);
                st_len += __shfl_xor_sync(FULL, st_len, o);
            }
            if (lane32 == 0) {
                atomicAdd(p.stats + CS_STAT_EPISODES, (double)st_eps);
                atomicAdd(p.stats + CS_STAT_EP_REWARD, (double)st_rew);
                atomicAdd(p.stats + CS_STAT_TARGETS_FOUND, (double)st_found);
                atomicAdd(p.stats + CS_STAT_WINS, (double)st_wins);
                atomicAdd(p.stats + CS_STAT_EP_LEN, (double)st_len);
            }
        }
    }
}

template <int N, int K, int MODE, bool MAP>
__global__ void __launch_bounds__(kTpeThreads, CS_TPE_MIN_CTAS) flight_tpe_kernel(const __grid_constant__ FlightParams p, const uint8_t* __restrict__ actions,
                                                                 const uint8_t* __restrict__ mask, uint32_t rflags, uint32_t seq) {
    flight_tpe_body<N, K, MODE, MAP>(p, (int)blockIdx.x, actions, mask, rflags, seq);
}

// Grouped step of several handles (independent env batches of the same shape: rollout workers) in ONE launch:
// blockIdx.y picks the handle, whose parameter block comes from a device table into shared memory.  A launch of a few
// thousand envs is bound by the launch path (2.2 us per 4096-env launch inside a 64-node graph);
// grouped, the same batches fill the GPU like one large handle.  flight_easy variant only.
constexpr int kMaxGroup = 128;
struct GroupActions { const uint8_t* a[kMaxGroup]; };

template <int N, int K>
__global__ void __launch_bounds__(kTpeThreads, CS_TPE_MIN_CTAS) flight_tpe_group_kernel(const FlightParams* __restrict__ table,
                                                                                       const __grid_constant__ GroupActions acts) {
    __shared__ FlightParams sp;
    {
        const uint32_t* src = reinterpret_cast<const uint32_t*>(table + blockIdx.y);
        uint32_t* dst = reinterpret_cast<uint32_t*>(&sp);
        for (int i = threadIdx.x; i < (int)(sizeof(FlightParams) / 4); i += kTpeThreads) dst[i] = src[i];
    }
    __syncthreads();
    if ((long long)blockIdx.x * kTpeThreads >= (long long)sp.E * K) return;         // handles may differ in num_envs
    flight_tpe_body<N, K, MODE_STEP, false>(sp, (int)blockIdx.x, acts.a[blockIdx.y], nullptr, 0u, 0u);
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
// Reference-shaped observation of the flight variant with the TMA (bulk async copy) engine:
// out[e][a] = prob_map[e].ravel() || (x^, y^, cos, sin)   (flight_env.py:223-230), i.e. every map is read once and
// written n times.  One elected thread per CTA drives a ring of kStages shared-memory buffers:
//   cp.async.bulk.shared::cluster.global (map e -> smem, completion on an mbarrier)   [SASS: UBLKCP]
//   n x cp.async.bulk.global.shared::cta (smem -> row (e,a)), one bulk group per env
// and refills a stage once the bulk group that read it has drained.  No SM load/store instruction touches the map;
// the 16-byte feature tails are written by the other lanes.  Needs (M*M*4) % 16 == 0.
// ------------------------------------------------------------------------------------------------
constexpr int kObsStages = 4;

__global__ void __launch_bounds__(32) flight_obs_full_tma_kernel(const float* __restrict__ map, const float* __restrict__ obs,
                                                                 float* __restrict__ out, int E, int n, uint32_t map_bytes) {
    extern __shared__ __align__(128) unsigned char tma_smem[];
    __shared__ __align__(8) unsigned long long full_bar[kObsStages];
    const int lane = threadIdx.x;
    const uint32_t stage_bytes = (map_bytes + 127u) & ~127u;
    const size_t row_bytes = (size_t)map_bytes + 16;
    const int my_count = (E - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x;      // envs of this CTA
    if (my_count <= 0) return;
    if (lane == 0) {
        for (int s = 0; s < kObsStages; ++s)
            asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&full_bar[s])));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncwarp();
    auto issue_load = [&](int it) {                      // env of iteration `it` -> stage it % kObsStages
        const int s = it % kObsStages;
        const size_t e = (size_t)blockIdx.x + (size_t)it * gridDim.x;
        const uint32_t bar = smem_u32(&full_bar[s]), dst = smem_u32(tma_smem + (size_t)s * stage_bytes);
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(map_bytes) : "memory");
        asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                     ::"r"(dst), "l"(reinterpret_cast<const unsigned char*>(map) + e * map_bytes), "r"(map_bytes), "r"(bar)
                     : "memory");
    };
    if (lane == 0)
        for (int it = 0; it < kObsStages && it < my_count; ++it) issue_load(it);
    for (int it = 0; it < my_count; ++it) {
        const int s = it % kObsStages;
        const size_t e = (size_t)blockIdx.x + (size_t)it * gridDim.x;
        // feature tails of this env: lane a writes the 16 bytes after the map of row (e, a)
        for (int a = lane; a < n; a += 32) {
            const float4 f = reinterpret_cast<const float4*>(obs)[e * n + a];
            *reinterpret_cast<float4*>(reinterpret_cast<unsigned char*>(out) + (e * n + a) * row_bytes + map_bytes) = f;
        }
        if (lane == 0) {
            const uint32_t bar = smem_u32(&full_bar[s]), src = smem_u32(tma_smem + (size_t)s * stage_bytes);
            const uint32_t parity = (uint32_t)(it / kObsStages) & 1u;
            uint32_t ok = 0, spins = 0;
            while (!ok) {
                asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                             : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
                if (!ok && ++spins > (1u << 26)) __trap();          // a lost copy must fail loudly, not hang the GPU
            }
            for (int a = 0; a < n; ++a)
                asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;"
                             ::"l"(reinterpret_cast<unsigned char*>(out) + (e * n + a) * row_bytes), "r"(src), "r"(map_bytes) : "memory");
            asm volatile("cp.async.bulk.commit_group;" ::: "memory");
            // the stage used by the PREVIOUS iteration may be refilled once its bulk group has finished reading it
            if (it >= 1 && it - 1 + kObsStages < my_count) {
                asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory");
                issue_load(it - 1 + kObsStages);
            }
        }
    }
    if (lane == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");     // all stores complete before exit
}

// Reference-shaped observation of the flight variant: [E][n][M*M+4] (flight_env.py:223-230), plain-copy form
__global__ void __launch_bounds__(256) flight_obs_full_kernel(const float* __restrict__ map, const float* __restrict__ obs,
                                                              float* __restrict__ out, int E, int n, int cells4) {
    // one row = cells4 float4 of the map + 1 float4 of agent features; rows = E*n
    const long long rows = (long long)E * n;
    const int row4 = cells4 + 1;
    const long long total = rows * row4;
    const float4* m4 = reinterpret_cast<const float4*>(map);
    const float4* o4 = reinterpret_cast<const float4*>(obs);
    float4* out4 = reinterpret_cast<float4*>(out);
    for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < total;
         idx += (long long)gridDim.x * blockDim.x) {
        const long long r = idx / row4;
        const int k = (int)(idx - r * row4);
        const long long e = r / n;
        out4[idx] = (k < cells4) ? m4[e * cells4 + k] : o4[r];
    }
}

__global__ void flight_obs_full_scalar_kernel(const float* __restrict__ map, const float* __restrict__ obs,
                                              float* __restrict__ out, int E, int n, int cells) {
    const long long rows = (long long)E * n;
    const int rowlen = cells + 4;
    const long long total = rows * rowlen;
    for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < total;
         idx += (long long)gridDim.x * blockDim.x) {
        const long long r = idx / rowlen;
        const int k = (int)(idx - r * rowlen);
        const long long e = r / n;
        out[idx] = (k < cells) ? map[e * cells + k] : obs[r * 4 + (k - cells)];
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

}  // namespace

// =================================================================================================
// host side
// =================================================================================================
struct cs_flight {
    cs_flight_cfg cfg;
    FlightParams p;
    int lpe;
    bool tpe;             // thread-per-env step kernel (n_agents <= kTpeMaxAgents and no explicit lanes_per_env)
    int tpe_k;            // threads that share one env's target loop in that kernel (1 or 4)
    size_t smem_bytes, map_smem;
    int grid, map_grid;
    bool map_tma;         // the TMA form of the map kernel applies (even map_size in 16..63, 2R+2 <= 16) and the tensor map exists
    CUtensorMap tmap;     // (2M, M/2, E) view of prob_map, box 20 x H x 1
    size_t map_tma_smem;
    bool map_use_tma;     // opt-in (CS_MAP_TMA=1): the TMA form measured 12 % slower than the direct form on the c4 workload
    uint32_t seq;         // value of CS_META_SENSE >> 1 that marks "sensed by the latest call" (constant: launches captured in CUDA graphs replay it)
    double* d_tmpl;
    uint8_t* d_actions;   // device staging of the *_host entry point's actions
    double* d_live;       // scratch of cs_flight_stats
    int obs_path;         // 0 = TMA bulk-copy kernel for the map observation, 1 = plain float4 copy kernel (A/B measurement)
    uint8_t* d_slab;      // one allocation behind reward | target_find | terminated | win | obs | state
    size_t slab_bytes, host_bytes, off_reward, off_tf, off_term, off_win, off_obs, off_state;
    longlong2* d_lut_meta;
    double2* d_lut;
    bool have_tmpl;
};

struct cs_flight_group {
    int count, n, k, grid_x, device;
    cs_flight* envs[kMaxGroup];
    FlightParams* d_table;
};

namespace {

// lanes per env: the smallest power of two that gives every agent and every target its own lane
int pick_lpe(const cs_flight_cfg& c) {
    int need = c.n_agents > c.target_num ? c.n_agents : c.target_num;
    if (c.lanes_per_env > need) need = c.lanes_per_env;
    int lpe = 1;
    while (lpe < need) lpe <<= 1;
    return lpe > 32 ? 32 : lpe;
}

// the belief maps of the envs the step / reset kernel just sensed (flight_env.py:266)
void launch_map(cs_flight* h, cudaStream_t st);

template <int N, int K>
cudaError_t launch_tpe_k(cs_flight* h, int mode, const uint8_t* actions, const uint8_t* mask, uint32_t rflags, cudaStream_t st) {
    const long long threads = (long long)h->p.E * K;
    const int grid = (int)((threads + kTpeThreads - 1) / kTpeThreads);
    if (h->p.variant) {
        if (mode == MODE_STEP)
            flight_tpe_kernel<N, K, MODE_STEP, true><<<grid, kTpeThreads, 0, st>>>(h->p, actions, mask, rflags, h->seq);
        else
            flight_tpe_kernel<N, K, MODE_RESET, true><<<grid, kTpeThreads, 0, st>>>(h->p, actions, mask, rflags, h->seq);
        launch_map(h, st);
    } else {
        if (mode == MODE_STEP)
            flight_tpe_kernel<N, K, MODE_STEP, false><<<grid, kTpeThreads, 0, st>>>(h->p, actions, mask, rflags, 0u);
        else
            flight_tpe_kernel<N, K, MODE_RESET, false><<<grid, kTpeThreads, 0, st>>>(h->p, actions, mask, rflags, 0u);
    }
    cs_count_launch(1);
    return cudaGetLastError();
}

template <int N>
cudaError_t launch_tpe(cs_flight* h, int mode, const uint8_t* actions, const uint8_t* mask, uint32_t rflags, cudaStream_t st) {
    switch (h->tpe_k) {
        case 1: return launch_tpe_k<N, 1>(h, mode, actions, mask, rflags, st);
        default: return launch_tpe_k<N, 4>(h, mode, actions, mask, rflags, st);
    }
}

template <int LPE>
cudaError_t launch_flight(cs_flight* h, int mode, const uint8_t* actions, const uint8_t* mask, uint32_t rflags,
                          cudaStream_t st) {
    if (h->p.variant) {
        if (mode == MODE_STEP)
            flight_kernel<LPE, MODE_STEP, true><<<h->grid, kThreads, h->smem_bytes, st>>>(h->p, actions, mask, rflags, h->seq);
        else
            flight_kernel<LPE, MODE_RESET, true><<<h->grid, kThreads, h->smem_bytes, st>>>(h->p, actions, mask, rflags, h->seq);
        launch_map(h, st);
    } else {
        if (mode == MODE_STEP)
            flight_kernel<LPE, MODE_STEP, false><<<h->grid, kThreads, h->smem_bytes, st>>>(h->p, actions, mask, rflags, 0u);
        else
            flight_kernel<LPE, MODE_RESET, false><<<h->grid, kThreads, h->smem_bytes, st>>>(h->p, actions, mask, rflags, 0u);
    }
    cs_count_launch(1);
    return cudaGetLastError();
}

void launch_map(cs_flight* h, cudaStream_t st) {
    if (h->p.M > 63)
        flight_map_wide_kernel<<<h->map_grid, kMapThreads, h->map_smem, st>>>(h->p, h->seq);
    else if (h->map_tma && !h->p.count_touched && h->map_use_tma)
        flight_map_tma_kernel<<<h->map_grid, kMapThreads, h->map_tma_smem, st>>>(h->p, h->tmap, h->seq);
    else if (h->p.M & 1)
        flight_map_kernel<false><<<h->map_grid, kMapThreads, h->map_smem, st>>>(h->p, h->seq);
    else
        flight_map_kernel<true><<<h->map_grid, kMapThreads, h->map_smem, st>>>(h->p, h->seq);
    cs_count_launch(1);
}

template <int LPE>
cudaError_t set_smem_attr(size_t bytes) {
    cudaError_t e = cudaFuncSetAttribute(flight_kernel<LPE, MODE_STEP, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes);
    if (e == cudaSuccess) e = cudaFuncSetAttribute(flight_kernel<LPE, MODE_RESET, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes);
    if (e == cudaSuccess) e = cudaFuncSetAttribute(flight_kernel<LPE, MODE_STEP, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes);
    if (e == cudaSuccess) e = cudaFuncSetAttribute(flight_kernel<LPE, MODE_RESET, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes);
    return e;
}

cudaError_t dispatch(cs_flight* h, int mode, const uint8_t* actions, const uint8_t* mask, uint32_t rflags,
                     cudaStream_t st) {
    if (h->tpe) {
        switch (h->p.n) {
            case 1: return launch_tpe<1>(h, mode, actions, mask, rflags, st);
            case 2: return launch_tpe<2>(h, mode, actions, mask, rflags, st);
            case 3: return launch_tpe<3>(h, mode, actions, mask, rflags, st);
            case 4: return launch_tpe<4>(h, mode, actions, mask, rflags, st);
            case 5: return launch_tpe<5>(h, mode, actions, mask, rflags, st);
            case 6: return launch_tpe<6>(h, mode, actions, mask, rflags, st);
            case 7: return launch_tpe<7>(h, mode, actions, mask, rflags, st);
            default: return launch_tpe<8>(h, mode, actions, mask, rflags, st);
        }
    }
    switch (h->lpe) {
        case 1: return launch_flight<1>(h, mode, actions, mask, rflags, st);
        case 2: return launch_flight<2>(h, mode, actions, mask, rflags, st);
        case 4: return launch_flight<4>(h, mode, actions, mask, rflags, st);
        case 8: return launch_flight<8>(h, mode, actions, mask, rflags, st);
        case 16: return launch_flight<16>(h, mode, actions, mask, rflags, st);
        default: return launch_flight<32>(h, mode, actions, mask, rflags, st);
    }
}

cudaError_t dispatch_attr(int lpe, size_t bytes) {
    switch (lpe) {
        case 1: return set_smem_attr<1>(bytes);
        case 2: return set_smem_attr<2>(bytes);
        case 4: return set_smem_attr<4>(bytes);
        case 8: return set_smem_attr<8>(bytes);
        case 16: return set_smem_attr<16>(bytes);
        default: return set_smem_attr<32>(bytes);
    }
}

template <int N>
cudaError_t launch_group(const cs_flight_group* g, const GroupActions& acts, cudaStream_t st) {
    const dim3 grid((unsigned)g->grid_x, (unsigned)g->count);
    if (g->k == 1) flight_tpe_group_kernel<N, 1><<<grid, kTpeThreads, 0, st>>>(g->d_table, acts);
    else flight_tpe_group_kernel<N, 4><<<grid, kTpeThreads, 0, st>>>(g->d_table, acts);
    cs_count_launch(1);
    return cudaGetLastError();
}

inline int up2(int v) { return (v + 1) & ~1; }

// Host-side heading table (see heading_sincos): clusters k = 1..36 around k*pi/18, window = rounding drift after
// `time_limit` steps with a 1.5x margin (measured drift: +-3.4e-14 after 200 steps ~ 1.7e-16 per step).
struct HeadingLut {
    std::vector<longlong2> meta;
    std::vector<double2> tab;
};

inline double ulp_of(double v) {
    long long b;
    memcpy(&b, &v, 8);
    ++b;
    double w;
    memcpy(&w, &b, 8);
    return w - v;
}

void build_heading_lut(int time_limit, HeadingLut* out) {
    const double drift = ((double)time_limit + 16.0) * 2.6e-16;
    out->meta.assign(37, make_longlong2(0, 0));
    out->tab.clear();
    for (int k = 1; k <= 36; ++k) {
        const double centre = (double)k * M_PI / 18.0;
        long long cb;
        memcpy(&cb, &centre, 8);
        long long half = (long long)ceil(drift / ulp_of(centre * 0.999)) + 4;
        if (half > 16384) half = 16384;             // very long episodes: the tail falls back to sincos()
        const long long base = (long long)out->tab.size();
        for (long long off = -half; off <= half; ++off) {
            const long long bits = cb + off;
            double h;
            memcpy(&h, &bits, 8);
            out->tab.push_back(make_double2(sin(h), cos(h)));     // HOST libm: the reference's own bits
        }
        out->meta[k] = make_longlong2(cb, base | (half << 32));
    }
}

// host mirror of the device lookup (tests call it through cs_debug_heading_lut)
void host_heading_sincos(const HeadingLut& lut, double h, double* sn, double* c, int* from_table) {
    const int k = (int)nearbyint(h * (18.0 / M_PI));
    *from_table = 1;
    if (k == 0 && fabs(h) < 7.450580596923828e-09) { *sn = h; *c = 1.0; return; }
    if (k >= 1 && k <= 36) {
        long long hb;
        memcpy(&hb, &h, 8);
        const long long off = hb - lut.meta[k].x, half = lut.meta[k].y >> 32;
        if (off >= -half && off <= half) {
            const double2 v = lut.tab[(size_t)((lut.meta[k].y & 0xffffffffLL) + half + off)];
            *sn = v.x; *c = v.y;
            return;
        }
    }
    *from_table = 0;
    *sn = sin(h); *c = cos(h);
}

}  // namespace

extern "C" {

int cs_flight_create(const cs_flight_cfg* cfg, cs_flight** out) {
    CS_REQUIRE(cfg && out, "cs_flight_create: null argument");
    CS_REQUIRE(cfg->struct_size == sizeof(cs_flight_cfg), "cs_flight_create: cfg.struct_size %u != %zu (ABI mismatch)",
               cfg->struct_size, sizeof(cs_flight_cfg));
    CS_REQUIRE(cfg->num_envs > 0, "num_envs must be > 0");
    CS_REQUIRE(cfg->n_agents >= 1 && cfg->n_agents <= CS_MAX_AGENTS, "n_agents must be in 1..%d", CS_MAX_AGENTS);
    CS_REQUIRE(cfg->target_num >= 1 && cfg->target_num <= CS_MAX_TARGETS, "target_num must be in 1..%d", CS_MAX_TARGETS);
    CS_REQUIRE(cfg->map_size >= 2 && cfg->map_size <= 4096, "map_size out of range");
    CS_REQUIRE(cfg->view_range >= 1, "view_range must be >= 1");
    CS_REQUIRE(cfg->time_limit >= 1 && cfg->time_limit <= 65535, "time_limit must be in 1..65535");
    CS_REQUIRE(cfg->agent_mode >= 0 && cfg->agent_mode <= 3, "No such agent mode");      // flight_env_easy.py:180
    CS_REQUIRE(cfg->target_mode == 0 || cfg->target_mode == 1, "No such target mode");   // flight_env_easy.py:136
    CS_REQUIRE(cfg->variant == 0 || cfg->variant == 1, "variant must be 0 (flight_easy) or 1 (flight)");
    const int l = cfg->lanes_per_env;
    CS_REQUIRE(l == 0 || l == 1 || l == 2 || l == 4 || l == 8 || l == 16 || l == 32, "lanes_per_env must be 0 or a power of two <= 32");

    cs_flight* h = new (std::nothrow) cs_flight();
    if (!h) { cs_set_error("out of host memory"); return CS_ERR_NOMEM; }
    memset(h, 0, sizeof(*h));
    h->cfg = *cfg;
    CS_CUDA(cudaSetDevice(cfg->device));

    FlightParams& p = h->p;
    const int n = cfg->n_agents, m = cfg->target_num, M = cfg->map_size;
    p.E = cfg->num_envs; p.n = n; p.m = m; p.M = M; p.T = cfg->time_limit;
    p.variant = cfg->variant; p.auto_reset = cfg->auto_reset; p.agent_mode = cfg->agent_mode;
    p.target_mode = cfg->target_mode; p.count_touched = cfg->count_touched;
    p.yaw_off = 2 * n;
    p.meta_off = up2(3 * n);
    p.rec = p.meta_off + CS_META_WORDS / 2;
    p.state_len = 4 * n + 3 * m;
    p.state_stride = (p.state_len + 3) & ~3;
    // constants, computed exactly as the reference's Python floats are
    p.Md = (double)M;
    p.half_M = 0.5 * (double)M;
    p.inv_half = 1.0 / ((double)M / 2.0);
    p.R = (double)cfg->view_range;
    p.R2 = (double)cfg->view_range * (double)cfg->view_range;
    p.v = cfg->velocity;
    p.fk = cfg->safe_dist * 0.8 * cfg->velocity;                 // safe_dist*POTENTIAL_FORCE_FACTOR*velocity (:299)
    p.fd2 = cfg->force_dist * cfg->force_dist;
    const double reach = cfg->force_dist + 1.01 * fabs(cfg->velocity) + 1e-9;
    p.near2 = reach * reach;
    p.q_miss = 1.0 - cfg->detect_prob;                           // (1 - detect_prob) (flight_env.py:292)
    p.pi = M_PI; p.two_pi = 2 * M_PI; p.three_pi = 3 * M_PI; p.half_pi = M_PI / 2; p.turn = M_PI / 18;
    if (cfg->detect_prob >= 1.0) p.thr = 0xFFFFFFFFLL;
    else if (cfg->detect_prob < 0.0) p.thr = -1;
    else p.thr = (long long)floor(cfg->detect_prob * 4294967296.0);
    p.seed = cfg->seed; p.env_id_base = cfg->env_id_base;
    p.inv_turn = 18.0 / M_PI;
    for (int a = 0; a < CS_MAX_AGENTS; ++a)
        p.lin[a] = (n != 1) ? (double)(a * M) / (double)(n - 1) : (double)M / 2.0;
    {
        const double h0 = (cfg->agent_mode <= 1) ? M_PI / 2 : (cfg->agent_mode == 2 ? 0.0 : M_PI);
        p.cos0 = cos(h0);
        p.sin0 = sin(h0);
    }

    p.span_cap = 1; p.span_shift = 0;
    while (p.span_cap < 2 * cfg->view_range) { p.span_cap <<= 1; ++p.span_shift; }
    h->lpe = pick_lpe(*cfg);
    // lanes_per_env: 0 = automatic; 1 or 4 = thread-per-env kernel with that many threads per env; larger = lane-per-agent kernel
    h->tpe = n <= kTpeMaxAgents && (cfg->lanes_per_env == 0 || cfg->lanes_per_env == 1 || cfg->lanes_per_env == 4);
    // measured on B200 (tools/sweep_step.sh): one thread per env wins from ~32k envs per launch (2.7e9 against 1.7e9
    // env-steps/s at 65536 envs, 5.0e9 against 2.4e9 at 1M); below that a launch cannot fill the GPU with one thread
    // per env and 4 threads per env match the lane-per-agent kernel's latency
    h->tpe_k = p.E >= 32768 ? 1 : 4;
    if (cfg->lanes_per_env == 1 || cfg->lanes_per_env == 4) h->tpe_k = cfg->lanes_per_env;
    if (const char* kenv = getenv("CS_TPE_K")) {                               // tuning sweeps only
        const int kv = atoi(kenv);
        if (kv == 1 || kv == 4) h->tpe_k = kv;
    }
    p.s_lut = 0;
    p.s_warp = 76;                                        // 37 x 16 B heading-table index, padded
    {
        // map kernel geometry: row slots per agent box (>= 2R+1 rows), lanes per row run (16 cells each, >= 2R+2 cells,
        // never more than the pair-aligned map row), per-warp scratch in 8-byte words
        const int rows = 2 * cfg->view_range + 1 < M ? 2 * cfg->view_range + 1 : M;
        const int cells = 2 * cfg->view_range + 2 < M + 1 ? 2 * cfg->view_range + 2 : M + 1;
        p.rs_shift = 0;
        while ((1 << p.rs_shift) < rows) ++p.rs_shift;
        p.lps_shift = 0;
        while ((16 << p.lps_shift) < cells && p.lps_shift < 4) ++p.lps_shift;
        p.ms_own = (M <= 63) ? M + 2 : 0;
        p.ms_col = p.ms_own + ((M <= 63) ? (n << p.rs_shift) : 0);
        p.ms_box = up2(p.ms_col + n);
        p.ms_xy = p.ms_box + 2 * n;
        p.ms_hit = p.ms_xy + 2 * n;
        p.ms_warp = up2(p.ms_hit + (m + 1) / 2);
        const int per_cta = (M <= 63) ? kMapEnvsPerCta : kMapThreads / 32;     // wide kernel: one warp per env
        h->map_smem = (size_t)per_cta * p.ms_warp * sizeof(unsigned long long);
        h->map_grid = (p.E + per_cta - 1) / per_cta;
        p.pre_stride = up2(2 * n + (m + 2) / 2);
    }
    h->smem_bytes = (size_t)(kThreads / 32) * p.s_warp * sizeof(double);
    const int env_per_cta = (kThreads / 32) * (32 / h->lpe);
    h->grid = (p.E + env_per_cta - 1) / env_per_cta;
    CS_CUDA(dispatch_attr(h->lpe, h->smem_bytes));
    if (cfg->variant && h->map_smem > 48 * 1024) {
        CS_CUDA(cudaFuncSetAttribute(flight_map_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)h->map_smem));
        CS_CUDA(cudaFuncSetAttribute(flight_map_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)h->map_smem));
        CS_CUDA(cudaFuncSetAttribute(flight_map_wide_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)h->map_smem));
    }

    const size_t E = (size_t)p.E;
    CS_CUDA(cudaMalloc(&p.dyn, E * p.rec * sizeof(double)));
    CS_CUDA(cudaMemset(p.dyn, 0, E * p.rec * sizeof(double)));
    // episode counter starts at -1 so that the first reset opens episode 0
    // structure of arrays exactly where one thread owns one env (flight_tpe_kernel<N, 1>), records otherwise
    if (h->tpe && h->tpe_k == 1) { p.dyn_rs = (long long)E; p.dyn_es = 1; p.tgt_rs = (long long)E; p.tgt_es = 1; }
    else { p.dyn_rs = 1; p.dyn_es = p.rec; p.tgt_rs = 1; p.tgt_es = 2 * m; }
    CS_CUDA(cudaMemset2D(reinterpret_cast<uint32_t*>(p.dyn + (size_t)(p.meta_off + CS_META_EPISODE / 2) * p.dyn_rs) + (CS_META_EPISODE & 1),
                         (size_t)p.dyn_es * sizeof(double), 0xFF, sizeof(uint32_t), E));
    CS_CUDA(cudaMalloc(&p.tgt, E * 2 * m * sizeof(double)));
    CS_CUDA(cudaMemset(p.tgt, 0, E * 2 * m * sizeof(double)));
    {
        // all step outputs live in one slab so that the host-buffer step can fetch them with a single D2H copy
        auto al = [](size_t v) { return (v + 255) & ~(size_t)255; };
        size_t off = 0;
        h->off_reward = off; off = al(off + E * sizeof(float));
        h->off_tf = off; off = al(off + E * sizeof(int32_t));
        h->off_term = off; off = al(off + E);
        h->off_win = off; off = al(off + E);
        h->off_state = off; off = al(off + E * p.state_stride * sizeof(float));
        h->host_bytes = off;                  // what the host-buffer step copies: obs rows are the state rows' first 4n floats
        h->off_obs = off; off = al(off + E * 4 * n * sizeof(float));
        h->slab_bytes = off;
        CS_CUDA(cudaMalloc(&h->d_slab, off));
        CS_CUDA(cudaMemset(h->d_slab, 0, off));
        p.reward = reinterpret_cast<float*>(h->d_slab + h->off_reward);
        p.target_find = reinterpret_cast<int32_t*>(h->d_slab + h->off_tf);
        p.terminated = h->d_slab + h->off_term;
        p.win = h->d_slab + h->off_win;
        p.obs = reinterpret_cast<float*>(h->d_slab + h->off_obs);
        p.state = reinterpret_cast<float*>(h->d_slab + h->off_state);
    }
    CS_CUDA(cudaMalloc(&h->d_live, sizeof(double)));
    CS_CUDA(cudaMalloc(&p.stats, CS_NUM_STATS * sizeof(double)));
    CS_CUDA(cudaMemset(p.stats, 0, CS_NUM_STATS * sizeof(double)));
    CS_CUDA(cudaMalloc(&h->d_tmpl, (size_t)m * 5 * sizeof(double)));
    CS_CUDA(cudaMemset(h->d_tmpl, 0, (size_t)m * 5 * sizeof(double)));
    p.tmpl = h->d_tmpl;
    if (cfg->variant) {
        CS_CUDA(cudaMalloc(&p.prob_map, E * M * M * sizeof(float)));
        CS_CUDA(cudaMemset(p.prob_map, 0, E * M * M * sizeof(float)));
        CS_CUDA(cudaMalloc(&p.pre, E * p.pre_stride * sizeof(double)));
        CS_CUDA(cudaMemset(p.pre, 0, E * p.pre_stride * sizeof(double)));
    }
    CS_CUDA(cudaMalloc(&h->d_actions, E * n));
    {
        HeadingLut lut;
        build_heading_lut(cfg->time_limit, &lut);
        CS_CUDA(cudaMalloc(&h->d_lut_meta, lut.meta.size() * sizeof(longlong2)));
        CS_CUDA(cudaMemcpy(h->d_lut_meta, lut.meta.data(), lut.meta.size() * sizeof(longlong2), cudaMemcpyHostToDevice));
        CS_CUDA(cudaMalloc(&h->d_lut, lut.tab.size() * sizeof(double2)));
        CS_CUDA(cudaMemcpy(h->d_lut, lut.tab.data(), lut.tab.size() * sizeof(double2), cudaMemcpyHostToDevice));
        p.lut_meta = h->d_lut_meta;
        p.lut = h->d_lut;
    }
    h->seq = 1u;
    h->map_use_tma = getenv("CS_MAP_TMA") != nullptr;         // A/B measurement and tests
    if (cfg->variant && !(M & 1) && M >= 10 && M <= 63 && 2 * cfg->view_range + 2 <= 16) {
        // TMA form of the map kernel: tensor map over prob_map viewed as (2M, M/2, E), tiles of 20 positions x H row pairs
        p.tile_hshift = p.rs_shift > 0 ? p.rs_shift - 1 : 0;
        p.tile_stride = (int)(((1u << p.tile_hshift) * 80u + 127u) & ~127u);
        const size_t tiles = (size_t)n * 2 * p.tile_stride;                   // [n][2] tiles of H x 20 floats
        p.mt_R = (int)tiles;
        p.mt_box = p.mt_R + (M + 2) * 8;
        p.mt_xy = p.mt_box + n * 16;
        p.mt_hit = p.mt_xy + n * 16;
        p.mt_bar = (p.mt_hit + (m + n) * 4 + 7) & ~7;
        p.mt_group = (p.mt_bar + 8 + 127) & ~127;
        h->map_tma_smem = (size_t)kMapEnvsPerCta * p.mt_group;
        typedef CUresult (*encode_fn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                      const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                      CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
        void* fn = nullptr;
        cudaDriverEntryPointQueryResult qres;
        if (h->map_tma_smem <= 200 * 1024 &&
            cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres) == cudaSuccess && fn &&
            qres == cudaDriverEntryPointSuccess) {
            const cuuint64_t gdim[3] = {(cuuint64_t)2 * M, (cuuint64_t)M / 2, (cuuint64_t)p.E};
            const cuuint64_t gstr[2] = {(cuuint64_t)2 * M * 4, (cuuint64_t)M * M * 4};
            const cuuint32_t bdim[3] = {20u, 1u << p.tile_hshift, 1u};
            const cuuint32_t estr[3] = {1u, 1u, 1u};
            const CUresult r = ((encode_fn)fn)(&h->tmap, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, p.prob_map, gdim, gstr, bdim, estr,
                                               CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE,
                                               CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
            h->map_tma = (r == CUDA_SUCCESS);
            if (h->map_tma && h->map_tma_smem > 48 * 1024)
                CS_CUDA(cudaFuncSetAttribute(flight_map_tma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)h->map_tma_smem));
        } else {
            cudaGetLastError();
        }
    }
    *out = h;
    return CS_OK;
}

void cs_flight_destroy(cs_flight* h) {
    if (!h) return;
    cudaSetDevice(h->cfg.device);
    cudaFree(h->p.dyn); cudaFree(h->p.tgt); cudaFree(h->d_slab); cudaFree(h->p.stats); cudaFree(h->d_live);
    cudaFree(h->d_tmpl); cudaFree(h->p.prob_map); cudaFree(h->p.pre); cudaFree(h->d_actions); cudaFree(h->d_lut_meta); cudaFree(h->d_lut);
    delete h;
}

int cs_flight_buffers_get(cs_flight* h, cs_flight_buffers* b) {
    CS_REQUIRE(h && b, "cs_flight_buffers_get: null argument");
    const FlightParams& p = h->p;
    b->dyn_row_stride = p.dyn_rs; b->dyn_env_stride = p.dyn_es; b->tgt_row_stride = p.tgt_rs; b->tgt_env_stride = p.tgt_es;
    b->dyn = p.dyn; b->dyn_doubles = p.rec; b->yaw_off = p.yaw_off; b->meta_off = p.meta_off; b->state_len = p.state_len; b->state_stride = p.state_stride;
    b->tgt = p.tgt; b->obs = p.obs; b->state = p.state; b->reward = p.reward; b->terminated = p.terminated;
    b->win = p.win; b->target_find = p.target_find; b->prob_map = p.prob_map; b->stats = p.stats;
    return CS_OK;
}

int cs_flight_env_info(const cs_flight* h, int32_t* out4) {
    CS_REQUIRE(h && out4, "cs_flight_env_info: null argument");
    out4[0] = 3;                       // n_actions       (flight_env_easy.py:32)
    out4[1] = h->p.state_len;          // state_shape     (:33)
    out4[2] = 4;                       // obs_shape       (:35)
    out4[3] = h->p.T;                  // episode_limit   (:76)
    return CS_OK;
}

int cs_flight_lanes_per_env(const cs_flight* h) { return h ? (h->tpe ? h->tpe_k : h->lpe) : CS_ERR_INVALID; }

// tuning / measurement hook: which kernel cs_flight_obs_full uses (0 = TMA bulk copies, 1 = plain float4 copies)
int cs_debug_flight_obs_path(cs_flight* h, int32_t path) {
    CS_REQUIRE(h && (path == 0 || path == 1), "cs_debug_flight_obs_path: bad argument");
    h->obs_path = path;
    return CS_OK;
}

// test hook (host only, no GPU needed): the heading table's sin/cos for `count` headings
int cs_debug_heading_lut(int32_t time_limit, const double* h_in, int32_t count, double* sin_out, double* cos_out,
                         int32_t* from_table) {
    CS_REQUIRE(h_in && sin_out && cos_out && time_limit >= 1, "cs_debug_heading_lut: bad argument");
    HeadingLut lut;
    build_heading_lut(time_limit, &lut);
    for (int i = 0; i < count; ++i) {
        int ft = 0;
        host_heading_sincos(lut, h_in[i], &sin_out[i], &cos_out[i], &ft);
        if (from_table) from_table[i] = ft;
    }
    return (int)lut.tab.size();
}

int cs_flight_set_target_template(cs_flight* h, const double* rows, int32_t nrows) {
    CS_REQUIRE(h && rows, "cs_flight_set_target_template: null argument");
    CS_REQUIRE(nrows >= h->p.m, "target template has %d rows, target_num is %d", nrows, h->p.m);
    const int m = h->p.m;
    double* tmp = new (std::nothrow) double[(size_t)m * 5];
    if (!tmp) { cs_set_error("out of host memory"); return CS_ERR_NOMEM; }
    const double a = (double)h->p.M / 10.0;                  // a = map_size/10  (flight_env_easy.py:97)
    for (int j = 0; j < m; ++j) {
        tmp[5 * j + 0] = a * rows[5 * j + 0];
        tmp[5 * j + 1] = a * rows[5 * j + 1];
        tmp[5 * j + 2] = a * rows[5 * j + 2];
        tmp[5 * j + 3] = a * rows[5 * j + 3];
        tmp[5 * j + 4] = rows[5 * j + 4];
    }
    cudaError_t e = cudaSetDevice(h->cfg.device);
    if (e == cudaSuccess) e = cudaMemcpy(h->d_tmpl, tmp, (size_t)m * 5 * sizeof(double), cudaMemcpyHostToDevice);
    delete[] tmp;
    CS_CUDA(e);
    h->have_tmpl = true;
    return CS_OK;
}

int cs_flight_reset(cs_flight* h, const uint8_t* d_mask, uint32_t flags, void* stream) {
    CS_REQUIRE(h, "cs_flight_reset: null handle");
    CS_REQUIRE((flags & CS_RESET_KEEP_TARGETS) || h->p.target_mode == 1 || h->have_tmpl,
               "target_mode 0 needs cs_flight_set_target_template before reset");
    CS_CUDA(dispatch(h, MODE_RESET, nullptr, d_mask, flags, (cudaStream_t)stream));
    return CS_OK;
}

int cs_flight_step(cs_flight* h, const uint8_t* d_actions, void* stream) {
    CS_REQUIRE(h && d_actions, "cs_flight_step: null argument");
    CS_CUDA(dispatch(h, MODE_STEP, d_actions, nullptr, 0u, (cudaStream_t)stream));
    return CS_OK;
}

int cs_flight_step_random(cs_flight* h, int32_t k, void* stream) {
    CS_REQUIRE(h && k >= 0, "cs_flight_step_random: bad argument");
    for (int i = 0; i < k; ++i) CS_CUDA(dispatch(h, MODE_STEP, nullptr, nullptr, 0u, (cudaStream_t)stream));
    return CS_OK;
}

int cs_flight_obs_full(cs_flight* h, float* d_out, void* stream) {
    CS_REQUIRE(h && d_out, "cs_flight_obs_full: null argument");
    CS_REQUIRE(h->p.variant == 1, "cs_flight_obs_full: only the flight (prob map) variant has a map observation");
    const FlightParams& p = h->p;
    const int cells = p.M * p.M;
    const long long total = (long long)p.E * p.n * (cells + 4);
    const size_t stage_bytes = ((size_t)cells * 4 + 127) & ~(size_t)127;
    if (cells % 4 == 0 && h->obs_path != 1 && kObsStages * stage_bytes <= 200 * 1024) {
        // TMA path: persistent single-warp CTAs, a few per SM, each streaming whole maps through shared memory
        const size_t smem = kObsStages * stage_bytes;
        int per_sm = (int)((200 * 1024) / smem);
        if (per_sm > 4) per_sm = 4;
        if (per_sm < 1) per_sm = 1;
        int grid = CS_NUM_SMS_B200 * per_sm;
        if (grid > p.E) grid = p.E;
        CS_CUDA(cudaFuncSetAttribute(flight_obs_full_tma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        flight_obs_full_tma_kernel<<<grid, 32, smem, (cudaStream_t)stream>>>(p.prob_map, p.obs, d_out, p.E, p.n, (uint32_t)cells * 4u);
    } else if (cells % 4 == 0) {
        const long long t4 = total / 4;
        const int grid = (int)((t4 + 255) / 256 < (long long)CS_NUM_SMS_B200 * 16 ? (t4 + 255) / 256 : CS_NUM_SMS_B200 * 16);
        flight_obs_full_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(p.prob_map, p.obs, d_out, p.E, p.n, cells / 4);
    } else {
        const int grid = (int)((total + 255) / 256 < (long long)CS_NUM_SMS_B200 * 16 ? (total + 255) / 256 : CS_NUM_SMS_B200 * 16);
        flight_obs_full_scalar_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(p.prob_map, p.obs, d_out, p.E, p.n, cells);
    }
    cs_count_launch(1);
    CS_CUDA(cudaGetLastError());
    return CS_OK;
}

int cs_flight_slab_layout(const cs_flight* h, uint64_t* out8) {
    CS_REQUIRE(h && out8, "cs_flight_slab_layout: null argument");
    out8[0] = h->host_bytes; out8[1] = h->off_reward; out8[2] = h->off_tf; out8[3] = h->off_term; out8[4] = h->off_win;
    out8[5] = h->off_obs; out8[6] = h->off_state; out8[7] = (uint64_t)h->p.state_stride * sizeof(float);
    return CS_OK;
}

int cs_flight_step_host(cs_flight* h, const cs_flight_host_io* io, void* stream) {
    CS_REQUIRE(h && io && io->actions, "cs_flight_step_host: null argument");
    const FlightParams& p = h->p;
    const size_t E = (size_t)p.E;
    cudaStream_t st = (cudaStream_t)stream;
    CS_CUDA(cudaSetDevice(h->cfg.device));
    CS_CUDA(cudaMemcpyAsync(h->d_actions, io->actions, E * p.n, cudaMemcpyHostToDevice, st));
    CS_CUDA(dispatch(h, MODE_STEP, h->d_actions, nullptr, 0u, st));
    if (io->slab) {
        // one copy for everything: reward | target_find | terminated | win | state (cs_flight_slab_layout); the obs
        // rows are the first 4n floats of the state rows and are not sent twice
        CS_CUDA(cudaMemcpyAsync(io->slab, h->d_slab, h->host_bytes, cudaMemcpyDeviceToHost, st));
    } else {
        if (io->reward) CS_CUDA(cudaMemcpyAsync(io->reward, p.reward, E * sizeof(float), cudaMemcpyDeviceToHost, st));
        if (io->terminated) CS_CUDA(cudaMemcpyAsync(io->terminated, p.terminated, E, cudaMemcpyDeviceToHost, st));
        if (io->win) CS_CUDA(cudaMemcpyAsync(io->win, p.win, E, cudaMemcpyDeviceToHost, st));
        if (io->obs) CS_CUDA(cudaMemcpyAsync(io->obs, p.obs, E * 4 * p.n * sizeof(float), cudaMemcpyDeviceToHost, st));
        if (io->state)   // compact [E][state_len] on the host, padded rows on the device
            CS_CUDA(cudaMemcpy2DAsync(io->state, p.state_len * sizeof(float), p.state, p.state_stride * sizeof(float),
                                      p.state_len * sizeof(float), E, cudaMemcpyDeviceToHost, st));
    }
    if (!(io->flags & CS_HOST_NO_SYNC)) CS_CUDA(cudaStreamSynchronize(st));
    return CS_OK;
}

// Many independent env batches (rollout workers) in one call: batch i is enqueued on streams[i % n_streams] without
// synchronising, then every stream is synchronised once (unless all ios carry CS_HOST_NO_SYNC).  Saves the per-call
// overhead of the host language, which dominates a host-buffer step of a few thousand envs.
int cs_flight_step_host_many(cs_flight* const* envs, const cs_flight_host_io* ios, int32_t count, void* const* streams, int32_t n_streams) {
    CS_REQUIRE(envs && ios && streams && count >= 0 && n_streams >= 1, "cs_flight_step_host_many: bad argument");
    bool sync = false;
    for (int i = 0; i < count; ++i) {
        cs_flight_host_io io = ios[i];
        sync |= !(io.flags & CS_HOST_NO_SYNC);
        io.flags |= CS_HOST_NO_SYNC;
        const int rc = cs_flight_step_host(envs[i], &io, streams[i % n_streams]);
        if (rc != CS_OK) return rc;
    }
    if (sync)
        for (int s = 0; s < n_streams && s < count; ++s) CS_CUDA(cudaStreamSynchronize((cudaStream_t)streams[s]));
    return CS_OK;
}

// ---- grouped device step (flight_tpe_group_kernel) ------------------------------------------------------------
int cs_flight_group_create(cs_flight* const* envs, int32_t count, cs_flight_group** out) {
    CS_REQUIRE(envs && out && count >= 1 && count <= kMaxGroup, "cs_flight_group_create: count must be in 1..%d", kMaxGroup);
    cs_flight_group* g = new (std::nothrow) cs_flight_group();
    if (!g) { cs_set_error("out of host memory"); return CS_ERR_NOMEM; }
    memset(g, 0, sizeof(*g));
    std::vector<FlightParams> table((size_t)count);
    for (int i = 0; i < count; ++i) {
        cs_flight* h = envs[i];
        const bool ok = h && h->tpe && h->p.variant == 0 && (i == 0 || (h->p.n == g->n && h->tpe_k == g->k && h->cfg.device == g->device));
        if (!ok) {
            delete g;
            cs_set_error("cs_flight_group_create: handle %d is not a flight_easy handle with n_agents <= %d, or differs from handle 0 in n_agents / threads per env / device", i, kTpeMaxAgents);
            return CS_ERR_INVALID;
        }
        if (i == 0) { g->n = h->p.n; g->k = h->tpe_k; g->device = h->cfg.device; }
        const long long threads = (long long)h->p.E * h->tpe_k;
        const int gx = (int)((threads + kTpeThreads - 1) / kTpeThreads);
        if (gx > g->grid_x) g->grid_x = gx;
        g->envs[i] = h;
        table[(size_t)i] = h->p;
    }
    g->count = count;
    cudaError_t e = cudaSetDevice(g->device);
    if (e == cudaSuccess) e = cudaMalloc(&g->d_table, (size_t)count * sizeof(FlightParams));
    if (e == cudaSuccess) e = cudaMemcpy(g->d_table, table.data(), (size_t)count * sizeof(FlightParams), cudaMemcpyHostToDevice);
    if (e != cudaSuccess) { cudaFree(g->d_table); delete g; }
    CS_CUDA(e);
    *out = g;
    return CS_OK;
}

void cs_flight_group_destroy(cs_flight_group* g) {
    if (!g) return;
    cudaSetDevice(g->device);
    cudaFree(g->d_table);
    delete g;
}

int cs_flight_group_step(cs_flight_group* g, const uint8_t* const* d_actions, void* stream) {
    CS_REQUIRE(g && d_actions, "cs_flight_group_step: null argument");
    GroupActions acts;
    for (int i = 0; i < g->count; ++i) {
        CS_REQUIRE(d_actions[i] != nullptr, "cs_flight_group_step: null actions for handle %d", i);
        acts.a[i] = d_actions[i];
    }
    cudaStream_t st = (cudaStream_t)stream;
    cudaError_t e;
    switch (g->n) {
        case 1: e = launch_group<1>(g, acts, st); break;
        case 2: e = launch_group<2>(g, acts, st); break;
        case 3: e = launch_group<3>(g, acts, st); break;
        case 4: e = launch_group<4>(g, acts, st); break;
        case 5: e = launch_group<5>(g, acts, st); break;
        case 6: e = launch_group<6>(g, acts, st); break;
        case 7: e = launch_group<7>(g, acts, st); break;
        default: e = launch_group<8>(g, acts, st); break;
    }
    CS_CUDA(e);
    return CS_OK;
}

// Episode-batch writer (see flight_record_kernel).  Buffers are caller-owned device memory.
int cs_flight_record_begin(cs_flight* h, const cs_episode_buffers* b, int32_t T, void* stream) {
    CS_REQUIRE(h && b && T >= 1, "cs_flight_record_begin: bad argument");
    const FlightParams& p = h->p;
    cudaStream_t st = (cudaStream_t)stream;
    const size_t ET = (size_t)p.E * T, O = 4 * (size_t)p.n, S = (size_t)p.state_len, A = 3 * (size_t)p.n;
    CS_CUDA(cudaSetDevice(h->cfg.device));
    CS_CUDA(cudaMemsetAsync(b->o, 0, ET * O * sizeof(float), st));
    CS_CUDA(cudaMemsetAsync(b->s, 0, ET * S * sizeof(float), st));
    CS_CUDA(cudaMemsetAsync(b->o_next, 0, ET * O * sizeof(float), st));
    CS_CUDA(cudaMemsetAsync(b->s_next, 0, ET * S * sizeof(float), st));
    CS_CUDA(cudaMemsetAsync(b->u, 0, ET * p.n, st));
    CS_CUDA(cudaMemsetAsync(b->u_onehot, 0, ET * A, st));
    CS_CUDA(cudaMemsetAsync(b->avail_u, 0, ET * A, st));
    CS_CUDA(cudaMemsetAsync(b->avail_u_next, 0, ET * A, st));
    CS_CUDA(cudaMemsetAsync(b->r, 0, ET * sizeof(float), st));
    CS_CUDA(cudaMemsetAsync(b->padded, 1, ET, st));                       // rollout.py:115-116
    CS_CUDA(cudaMemsetAsync(b->terminated, 1, ET, st));
    CS_CUDA(cudaMemsetAsync(b->episode_reward, 0, (size_t)p.E * sizeof(float), st));
    CS_CUDA(cudaMemsetAsync(b->win_tag, 0, (size_t)p.E, st));
    CS_CUDA(cudaMemsetAsync(b->targets_find, 0, (size_t)p.E * sizeof(int32_t), st));
    CS_CUDA(cudaMemsetAsync(b->length, 0, (size_t)p.E * sizeof(int32_t), st));
    const long long total = (long long)p.E * (O + S);
    const int grid = (int)((total + 255) / 256 < (long long)CS_NUM_SMS_B200 * 8 ? (total + 255) / 256 : CS_NUM_SMS_B200 * 8);
    flight_record_begin_kernel<<<grid, 256, 0, st>>>(p, *b, T);
    cs_count_launch(1);
    CS_CUDA(cudaGetLastError());
    return CS_OK;
}

int cs_flight_record(cs_flight* h, const cs_episode_buffers* b, int32_t t, int32_t T, const uint8_t* d_actions, void* stream) {
    CS_REQUIRE(h && b && d_actions && t >= 0 && t < T, "cs_flight_record: bad argument");
    const FlightParams& p = h->p;
    const long long total = (long long)p.E * (4 * p.n + p.state_len + 1);
    const int grid = (int)((total + 255) / 256 < (long long)CS_NUM_SMS_B200 * 8 ? (total + 255) / 256 : CS_NUM_SMS_B200 * 8);
    flight_record_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(p, *b, t, T, d_actions);
    cs_count_launch(1);
    CS_CUDA(cudaGetLastError());
    return CS_OK;
}

int cs_flight_stats(cs_flight* h, double* h_out, void* stream) {
    CS_REQUIRE(h && h_out, "cs_flight_stats: null argument");
    cudaStream_t st = (cudaStream_t)stream;
    CS_CUDA(cudaSetDevice(h->cfg.device));
    // env_steps = lengths of the finished episodes + steps of the episodes still running
    CS_CUDA(cudaMemsetAsync(h->d_live, 0, sizeof(double), st));
    const int grid = (h->p.E + 255) / 256 < CS_NUM_SMS_B200 * 4 ? (h->p.E + 255) / 256 : CS_NUM_SMS_B200 * 4;
    flight_live_steps_kernel<<<grid, 256, 0, st>>>(h->p.dyn, h->p.E, h->p.dyn_rs, h->p.dyn_es, h->p.meta_off, h->d_live);
    cs_count_launch(1);
    CS_CUDA(cudaGetLastError());
    double live = 0.0;
    CS_CUDA(cudaMemcpyAsync(h_out, h->p.stats, CS_NUM_STATS * sizeof(double), cudaMemcpyDeviceToHost, st));
    CS_CUDA(cudaMemcpyAsync(&live, h->d_live, sizeof(double), cudaMemcpyDeviceToHost, st));
    CS_CUDA(cudaStreamSynchronize(st));
    h_out[CS_STAT_ENV_STEPS] = h_out[CS_STAT_EP_LEN] + live;
    return CS_OK;
}

}  // extern "C"
