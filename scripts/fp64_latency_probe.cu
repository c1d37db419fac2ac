// Micro-probe: cycles per dependent FP64 op with 1 / 8 warps per SM doing the same chain (B200, sm_100a).
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o /tmp/fp64probe scripts/fp64_latency_probe.cu
#include <cstdio>
#include <cuda_runtime.h>
template <int OP> __global__ void probe(double* out, long long* cyc, double x0, int iters) {
    double x = x0 + threadIdx.x * 1e-9;
    long long t0 = clock64();
    for (int i = 0; i < iters; ++i) {
        if (OP == 0) x = fma(x, 1.0000001, 1e-9);
        if (OP == 1) x = sqrt(x) + 1.5;
        if (OP == 2) x = 1.0 / x + 1.5;
        if (OP == 3) x = x + 1e-9;
    }
    long long t1 = clock64();
    out[blockIdx.x * blockDim.x + threadIdx.x] = x;
    if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}
int main() {
    double* out; long long* cyc; long long h[4];
    cudaMalloc(&out, 1 << 20); cudaMalloc(&cyc, 64);
    const char* names[4] = {"DFMA", "sqrt", "div", "DADD"};
    for (int threads : {32, 256, 1024}) {
        for (int op = 0; op < 4; ++op) {
            const int iters = 2000;
            if (op == 0) probe<0><<<1, threads>>>(out, cyc, 1.1, iters);
            if (op == 1) probe<1><<<1, threads>>>(out, cyc, 1.1, iters);
            if (op == 2) probe<2><<<1, threads>>>(out, cyc, 1.1, iters);
            if (op == 3) probe<3><<<1, threads>>>(out, cyc, 1.1, iters);
            cudaDeviceSynchronize();
            cudaMemcpy(h, cyc, 8, cudaMemcpyDeviceToHost);
            printf("%4d threads/SM  %-5s %8.1f cycles per dependent op\n", threads, names[op], (double)h[0] / iters);
        }
    }
    return 0;
}
