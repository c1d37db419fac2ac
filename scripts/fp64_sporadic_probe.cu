// Probe: cost of a sporadic FP64 sqrt + 2 divisions by ONE thread between block barriers (pattern of the QR column step).
#include <cstdio>
#include <cuda_runtime.h>
__device__ __forceinline__ double soft_rsqrt(double a) {      // no MUFU.RSQ64H: fp32 seed + 3 Newton steps
    double y = (double)rsqrtf((float)a);
#pragma unroll
    for (int i = 0; i < 3; ++i) y = y * fma(-0.5 * a, y * y, 1.5);
    return y;
}
template <int MODE> __global__ void probe(double* out, long long* cyc, double s0, int iters) {
    __shared__ double hh[4];
    double acc = 0.0, sig = s0, alpha = 0.3;
    long long t0 = clock64();
    for (int i = 0; i < iters; ++i) {
        __syncthreads();
        if (threadIdx.x == 0) {
            double beta, tau, scale;
            if (MODE == 0) {
                const double nrm = sqrt(fma(alpha, alpha, sig));
                beta = -nrm; tau = (beta - alpha) / beta; scale = 1.0 / (alpha - beta);
            } else if (MODE == 1) {
                const double a = fma(alpha, alpha, sig);
                const double r = soft_rsqrt(a);
                const double nrm = a * r;
                beta = -nrm;
                tau = (beta - alpha) * (-r);
                const double dnm = alpha - beta;
                double y = (double)(1.0f / (float)dnm);
                y = y * (2.0 - dnm * y); y = y * (2.0 - dnm * y); y = y * (2.0 - dnm * y);
                scale = y;
            } else {
                beta = -sig; tau = alpha * 0.5; scale = sig * 0.25;      // no special functions at all
            }
            hh[0] = beta; hh[1] = tau; hh[2] = scale;
        }
        __syncthreads();
        acc = fma(hh[0], hh[1], acc) + hh[2];
        sig = sig * 1.0000001 + 1e-7 * threadIdx.x;
    }
    long long t1 = clock64();
    out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
    if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}
int main() {
    double* out; long long* cyc; long long h[1];
    cudaMalloc(&out, 1 << 22); cudaMalloc(&cyc, 1 << 12);
    const char* names[3] = {"sqrt + 2 div (MUFU.64H)", "fp32-seeded Newton", "no special functions"};
    for (int grid : {1, 86}) for (int mode = 0; mode < 3; ++mode) {
        const int iters = 2000;
        if (mode == 0) probe<0><<<grid, 256>>>(out, cyc, 1.1, iters);
        if (mode == 1) probe<1><<<grid, 256>>>(out, cyc, 1.1, iters);
        if (mode == 2) probe<2><<<grid, 256>>>(out, cyc, 1.1, iters);
        cudaDeviceSynchronize();
        cudaMemcpy(h, cyc, 8, cudaMemcpyDeviceToHost);
        printf("grid %3d  %-26s %8.1f cycles per barrier-separated step\n", grid, names[mode], (double)h[0] / iters);
    }
    return 0;
}
