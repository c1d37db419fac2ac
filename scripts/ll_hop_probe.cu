// Probe: latency of LL-line exchanges between co-resident CTAs (one per SM) on B200.
//   mode 0: all-to-all, every CTA publishes Q lines and polls Q x G lines (thread t -> lines t, t+256, ...)
//   mode 1: reducer: CTA q sums quantity q over the G CTAs and publishes the total; all CTAs poll Q totals (2 hops)
//   mode 2: 1 line per CTA all-to-all (pure barrier latency)
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#include <cooperative_groups.h>
struct __align__(16) LL { uint32_t lo, f1, hi, f2; };
__device__ __forceinline__ void ll_store(LL* p, double v, uint32_t flag) {
    const uint32_t lo = (uint32_t)__double2loint(v), hi = (uint32_t)__double2hiint(v);
    asm volatile("st.relaxed.gpu.global.v4.u32 [%0], {%1, %2, %3, %4};" ::"l"(p), "r"(lo), "r"(flag), "r"(hi), "r"(flag) : "memory");
}
__device__ __forceinline__ bool ll_try(const LL* p, uint32_t flag, double& v) {
    uint32_t lo, f1, hi, f2;
    asm volatile("ld.relaxed.gpu.global.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(lo), "=r"(f1), "=r"(hi), "=r"(f2) : "l"(p) : "memory");
    v = __hiloint2double((int)hi, (int)lo);
    return f1 == flag && f2 == flag;
}
constexpr int MAXG = 160, Q = 16;
// mode 3: reducer scheme with PRIVATE lines: partial (q, b) sits in its own 32-byte sector and is read only by reducer q;
//         reducer q writes one copy of its total per reader (priv[par][reader][q]); a reader polls its own 16 lines.
__global__ void __launch_bounds__(256, 1) probe_private(LL* part, LL* priv, long long* cyc, int iters) {
    __shared__ double red[Q];
    __shared__ double rs[8];
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5, G = gridDim.x, b = blockIdx.x;
    double acc = 0.0;
    long long t0 = clock64();
    for (int it = 1; it <= iters; ++it) {
        const uint32_t epoch = (uint32_t)it;
        const int par = it & 1;
        if (tid < Q) ll_store(part + 2 * (((size_t)(par * Q + tid)) * MAXG + b), 1.0 + tid + acc * 1e-30, epoch);
        if (b < Q) {
            double v = 0.0;
            if (tid < G) while (!ll_try(part + 2 * (((size_t)(par * Q + b)) * MAXG + tid), epoch, v)) {}
            v += __shfl_xor_sync(0xffffffffu, v, 16); v += __shfl_xor_sync(0xffffffffu, v, 8);
            v += __shfl_xor_sync(0xffffffffu, v, 4); v += __shfl_xor_sync(0xffffffffu, v, 2); v += __shfl_xor_sync(0xffffffffu, v, 1);
            if (lane == 0) rs[wid] = v;
            __syncthreads();
            const double t = rs[0] + rs[1] + rs[2] + rs[3] + rs[4] + rs[5] + rs[6] + rs[7];
            if (tid < G) ll_store(priv + ((size_t)(par * MAXG + tid)) * Q + b, t, epoch);
            __syncthreads();
        }
        if (tid < Q) { double v; while (!ll_try(priv + ((size_t)(par * MAXG + b)) * Q + tid, epoch, v)) {} red[tid] = v; }
        __syncthreads();
        acc += red[tid & 15];
        __syncthreads();
    }
    long long t1 = clock64();
    if (tid == 0) cyc[b] = t1 - t0;
    if (acc == 12345.678) cyc[b] = 0;
}
template <int MODE> __global__ void __launch_bounds__(256, 1) probe(LL* step, LL* tot, long long* cyc, int iters) {
    __shared__ double red[Q];
    __shared__ double rs[8];
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5, G = gridDim.x, b = blockIdx.x;
    double acc = 0.0;
    long long t0 = clock64();
    for (int it = 1; it <= iters; ++it) {
        const uint32_t epoch = (uint32_t)it;
        const int par = it & 1;
        const int nq = MODE == 2 ? 1 : Q;
        if (tid < nq) ll_store(step + ((size_t)(par * Q + tid)) * MAXG + b, 1.0 + tid + acc * 1e-30, epoch);
        if (MODE == 0 || MODE == 2) {
            double s = 0.0;
            for (int L = tid; L < nq * G; L += 256) {
                const int q = L / G, bb = L - q * G;
                double v; while (!ll_try(step + ((size_t)(par * Q + q)) * MAXG + bb, epoch, v)) {}
                s += v;
            }
            s += __shfl_xor_sync(0xffffffffu, s, 16); s += __shfl_xor_sync(0xffffffffu, s, 8);
            s += __shfl_xor_sync(0xffffffffu, s, 4); s += __shfl_xor_sync(0xffffffffu, s, 2); s += __shfl_xor_sync(0xffffffffu, s, 1);
            if (lane == 0) rs[wid] = s;
            __syncthreads();
            acc += rs[0] + rs[1] + rs[2] + rs[3] + rs[4] + rs[5] + rs[6] + rs[7];
            __syncthreads();
        } else {
            if (b < Q) {
                double v = 0.0;
                if (tid < G) while (!ll_try(step + ((size_t)(par * Q + b)) * MAXG + tid, epoch, v)) {}
                v += __shfl_xor_sync(0xffffffffu, v, 16); v += __shfl_xor_sync(0xffffffffu, v, 8);
                v += __shfl_xor_sync(0xffffffffu, v, 4); v += __shfl_xor_sync(0xffffffffu, v, 2); v += __shfl_xor_sync(0xffffffffu, v, 1);
                if (lane == 0) rs[wid] = v;
                __syncthreads();
                if (tid == 0) ll_store(tot + par * Q + b, rs[0] + rs[1] + rs[2] + rs[3] + rs[4] + rs[5] + rs[6] + rs[7], epoch);
                __syncthreads();
            }
            if (tid < Q) { double v; while (!ll_try(tot + par * Q + tid, epoch, v)) {} red[tid] = v; }
            __syncthreads();
            acc += red[tid & 15];
            __syncthreads();
        }
    }
    long long t1 = clock64();
    if (tid == 0) cyc[b] = t1 - t0;
    if (acc == 12345.678) cyc[b] = 0;
}
int main() {
    LL *step, *tot; long long* cyc; long long h[MAXG];
    cudaMalloc(&step, 4 * Q * MAXG * sizeof(LL)); cudaMalloc(&tot, 2 * Q * MAXG * sizeof(LL)); cudaMalloc(&cyc, MAXG * 8);
    int iters = 2000;
    for (int G : {2, 16, 32, 64, 86, 128, 148}) for (int mode = 0; mode < 4; ++mode) {
        if ((mode == 1 || mode == 3) && G < Q) continue;          // (the reducer scheme needs one CTA per quantity)
        cudaMemset(step, 0, 4 * Q * MAXG * sizeof(LL)); cudaMemset(tot, 0, 2 * Q * MAXG * sizeof(LL));
        void* args[] = {&step, &tot, &cyc, &iters};
        const void* f = mode == 0 ? (const void*)probe<0> : mode == 1 ? (const void*)probe<1> : mode == 2 ? (const void*)probe<2> : (const void*)probe_private;
        cudaError_t e = cudaLaunchCooperativeKernel(f, dim3(G), dim3(256), args, 0, 0);
        cudaDeviceSynchronize();
        cudaMemcpy(h, cyc, G * 8, cudaMemcpyDeviceToHost);
        printf("G %3d  mode %d (%s)  %8.1f cycles per exchange  (%s)\n", G, mode,
               mode == 0 ? "all-to-all 16 q" : mode == 1 ? "reducer 16 q, 2 hops" : mode == 2 ? "all-to-all 1 line" : "reducer, private lines", (double)h[0] / iters,
               cudaGetErrorString(e));
    }
    return 0;
}
