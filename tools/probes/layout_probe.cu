// Does HBM care whether a warp's 43 state rows are 43 separate 256-byte pieces 8 MB apart (structure of arrays) or one
// contiguous 11 KB block (blocked structure of arrays)?  Reads R rows per env, writes W rows per env, 1M envs, no arithmetic
// worth mentioning.  nvcc -arch=sm_100a -O3 -o /tmp/layout_probe tools/probes/layout_probe.cu && /tmp/layout_probe
#include <cstdio>
#include <cuda_runtime.h>
constexpr int R = 43, W = 13;
template <bool BLOCKED, int B>
__global__ void __launch_bounds__(64, 8) probe(const double* __restrict__ in, double* __restrict__ out, int E) {
    const int e = blockIdx.x * 64 + threadIdx.x;
    if (e >= E) return;
    const size_t base = BLOCKED ? (size_t)(e / B) * R * B + (e % B) : (size_t)e;
    const size_t rs = BLOCKED ? B : (size_t)E;
    double v[R];
#pragma unroll
    for (int r = 0; r < R; ++r) v[r] = in[base + r * rs];
    double s = 0;
#pragma unroll
    for (int r = W; r < R; ++r) s += v[r];
    const size_t obase = BLOCKED ? (size_t)(e / B) * W * B + (e % B) : (size_t)e;
#pragma unroll
    for (int r = 0; r < W; ++r) out[obase + r * rs] = v[r] + s;
}
template <bool BLOCKED, int B>
float run(const double* in, double* out, int E, int reps) {
    cudaEvent_t a, b;
    cudaEventCreate(&a); cudaEventCreate(&b);
    for (int i = 0; i < 3; ++i) probe<BLOCKED, B><<<(E + 63) / 64, 64>>>(in, out, E);
    cudaEventRecord(a);
    for (int i = 0; i < reps; ++i) probe<BLOCKED, B><<<(E + 63) / 64, 64>>>(in, out, E);
    cudaEventRecord(b);
    cudaEventSynchronize(b);
    float ms;
    cudaEventElapsedTime(&ms, a, b);
    return ms / reps;
}
int main() {
    const int E = 1 << 20;
    double *in, *out;
    cudaMalloc(&in, (size_t)E * R * 8);
    cudaMalloc(&out, (size_t)E * W * 8);
    cudaMemset(in, 0, (size_t)E * R * 8);
    const double bytes = (double)E * (R + W) * 8;
    float t;
    t = run<false, 32>(in, out, E, 20); printf("soa            %.1f us  %.0f GB/s\n", t * 1e3, bytes / t / 1e6);
    t = run<true, 32>(in, out, E, 20);  printf("blocked B=32   %.1f us  %.0f GB/s\n", t * 1e3, bytes / t / 1e6);
    t = run<true, 64>(in, out, E, 20);  printf("blocked B=64   %.1f us  %.0f GB/s\n", t * 1e3, bytes / t / 1e6);
    t = run<true, 128>(in, out, E, 20); printf("blocked B=128  %.1f us  %.0f GB/s\n", t * 1e3, bytes / t / 1e6);
    return 0;
}
