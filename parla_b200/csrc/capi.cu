// Library-level plumbing: version, thread-local error string, device queries.
#include <stdarg.h>
#include <atomic>
#include "common.cuh"
#include "../../include/parla_b200.h"

namespace pla {

static thread_local char g_err[512] = "";

void set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

static std::atomic<long long> g_launches{0};
void note_launch(int n) { g_launches.fetch_add(n, std::memory_order_relaxed); }

int num_sms() {
    static thread_local int cached_dev = -1, cached = 0;
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) return 148;
    if (dev != cached_dev) {
        int v = 0;
        if (cudaDeviceGetAttribute(&v, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || v <= 0) v = 148;
        cached = v;
        cached_dev = dev;
    }
    return cached;
}

// ------------------------------------------------------------------ sum of squares (two-stage, fixed order)
__global__ void __launch_bounds__(256) sumsq_stage1(const double* __restrict__ x, long long n, double* part) {
    __shared__ double scratch[33];
    double acc = 0.0;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
        acc = fma(x[i], x[i], acc);
    acc = block_sum(acc, scratch);
    if (threadIdx.x == 0) part[blockIdx.x] = acc;
}
__global__ void __launch_bounds__(256) sumsq_stage2(const double* part, int nb, double* out) {
    __shared__ double scratch[33];
    double acc = 0.0;
    for (int i = threadIdx.x; i < nb; i += blockDim.x) acc += part[i];
    acc = block_sum(acc, scratch);
    if (threadIdx.x == 0) out[0] = acc;
}

// ------------------------------------------------------------------ FP64 tensor pipe probe (measurement hook)
__global__ void __launch_bounds__(256) dmma_probe_kernel(int iters, double* sink) {
    double c[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) c[i] = 0.0;
    const double a = 1.0 + threadIdx.x * 1e-9, b = 1.0 - threadIdx.x * 1e-9;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < 8; ++i) dmma884(c[2 * i], c[2 * i + 1], a, b);
    }
    double s = 0.0;
#pragma unroll
    for (int i = 0; i < 16; ++i) s += c[i];
    if (s == 123.456) sink[0] = s;
}

}  // namespace pla

using namespace pla;

// Runs `iters` x 8 independent DMMA.8x8x4 per warp on every SM (8 warps x `ctas_per_sm` CTAs).
// flops = 2 * 256 * 8 * iters * (8 warps) * grid.  Timing is the caller's (CUDA events).
extern "C" int pla_dmma_probe(int iters, int ctas_per_sm, double* sink, void* stream) {
    PLA_CHECK_ARG(iters >= 1, 1, "iters < 1");
    PLA_CHECK_ARG(ctas_per_sm >= 1 && ctas_per_sm <= 8, 2, "ctas_per_sm out of range");
    PLA_CHECK_ARG(sink != nullptr, 3, "sink is null");
    dmma_probe_kernel<<<num_sms() * ctas_per_sm, 256, 0, (cudaStream_t)stream>>>(iters, sink);
    PLA_LAUNCH_CHECK();
    return 0;
}


extern "C" int pla_version(void) { return 100; }
extern "C" const char* pla_last_error(void) { return g_err; }
extern "C" int pla_num_sms(void) { return num_sms(); }
extern "C" long long pla_launch_count(void) { return g_launches.load(std::memory_order_relaxed); }
extern "C" void pla_note_launches(long long n) { g_launches.fetch_add(n, std::memory_order_relaxed); }

extern "C" size_t pla_sumsq_workspace_bytes(int64_t n) {
    (void)n;
    return (size_t)1024 * sizeof(double);
}

extern "C" int pla_sumsq_f64(const double* x, int64_t n, double* out, void* ws, size_t ws_bytes, void* stream) {
    PLA_CHECK_ARG(x != nullptr, 1, "x is null");
    PLA_CHECK_ARG(n >= 0, 2, "n < 0");
    PLA_CHECK_ARG(out != nullptr, 3, "out is null");
    PLA_CHECK_ARG(ws != nullptr && ws_bytes >= pla_sumsq_workspace_bytes(n), 5, "workspace too small");
    cudaStream_t st = (cudaStream_t)stream;
    int nb = (int)((n + 255) / 256);
    if (nb < 1) nb = 1;
    if (nb > 1024) nb = 1024;
    sumsq_stage1<<<nb, 256, 0, st>>>(x, n, (double*)ws);
    PLA_LAUNCH_CHECK();
    sumsq_stage2<<<1, 256, 0, st>>>((const double*)ws, nb, out);
    PLA_LAUNCH_CHECK();
    return 0;
}
