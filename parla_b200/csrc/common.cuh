// Shared device/host helpers for libparla_b200 (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>

namespace pla {

// ------------------------------------------------------------------ host-side error plumbing
void set_error(const char* fmt, ...);          // capi.cu (thread-local message)
int num_sms();                                  // capi.cu (cached per current device)
void note_launch(int n = 1);                    // capi.cu (process-wide kernel-launch counter)

#define PLA_CHECK_ARG(cond, idx, msg)                                             \
    do { if (!(cond)) { ::pla::set_error("%s: bad argument %d: %s", __func__, (idx), (msg)); \
                        return -(idx); } } while (0)

#define PLA_CUDA(call)                                                            \
    do { cudaError_t e__ = (call);                                                \
         if (e__ != cudaSuccess) {                                                \
             ::pla::set_error("%s: %s failed: %s", __func__, #call, cudaGetErrorString(e__)); \
             return (int)e__; } } while (0)

#define PLA_LAUNCH_CHECK()                                                        \
    do { cudaError_t e__ = cudaGetLastError();                                    \
         if (e__ != cudaSuccess) {                                                \
             ::pla::set_error("%s: kernel launch failed: %s", __func__, cudaGetErrorString(e__)); \
             return (int)e__; }                                                   \
         ::pla::note_launch(); } while (0)

// ------------------------------------------------------------------ device helpers
#ifdef __CUDACC__

__device__ __forceinline__ uint32_t smem_u32(const void* p) {
    return (uint32_t)__cvta_generic_to_shared(p);
}

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// Sum NV (<= 32) per-lane values across the warp in one butterfly: 31 shuffles instead of 5 * NV.
// On return lane l (< NV) holds the warp total of value l in v[0].  Fixed order => deterministic.
template <int NV>
__device__ __forceinline__ void warp_multi_sum(double (&v)[32]) {
    const int lane = threadIdx.x & 31;
#pragma unroll
    for (int i = NV; i < 32; ++i) v[i] = 0.0;
#pragma unroll
    for (int off = 16; off >= 1; off >>= 1) {
        const bool upper = (lane & off) != 0;
#pragma unroll
        for (int i = 0; i < off; ++i) {
            const double send = upper ? v[i] : v[i + off];
            const double keep = upper ? v[i + off] : v[i];
            v[i] = keep + __shfl_xor_sync(0xffffffffu, send, off);
        }
    }
}

// Block-wide sum for blockDim.x <= 1024 (multiple of 32); result valid in every thread.
// `scratch` must hold 33 doubles. Deterministic (fixed tree).
__device__ __forceinline__ double block_sum(double v, double* scratch) {
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
    v = warp_sum(v);
    __syncthreads();                 // protect scratch reuse
    if (lane == 0) scratch[wid] = v;
    __syncthreads();
    if (wid == 0) {
        double t = lane < nw ? scratch[lane] : 0.0;
        t = warp_sum(t);
        if (lane == 0) scratch[32] = t;
    }
    __syncthreads();
    return scratch[32];
}

// ---- mbarrier / bulk-copy (TMA engine, non-tensor form) ----
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void fence_mbar_init() {
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void fence_proxy_async() {
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes)
                 : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
    uint32_t ok;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok)
        : "r"(smem_u32(bar)), "r"(parity)
        : "memory");
    return ok != 0;
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    while (!mbar_try_wait(bar, parity)) {}
}
__device__ __forceinline__ uint64_t l2_policy_evict_first() {
    uint64_t pol;
    asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol));
    return pol;
}
// global -> shared bulk copy, completion signalled on `bar` (complete_tx). 16-byte aligned, size % 16 == 0.
__device__ __forceinline__ void bulk_g2s(void* dst_smem, const void* src_gmem, uint32_t bytes, uint64_t* bar,
                                         uint64_t policy) {
    asm volatile(
        "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1], %2, [%3], %4;"
        ::"r"(smem_u32(dst_smem)), "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar)), "l"(policy)
        : "memory");
}

// ---- cp.async (LDGSTS) ----
__device__ __forceinline__ void cp_async8(void* dst_smem, const void* src, bool pred) {
    int sz = pred ? 8 : 0;   // src-size 0 => zero fill
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8, %2;" ::"r"(smem_u32(dst_smem)), "l"(src), "r"(sz) : "memory");
}
__device__ __forceinline__ void cp_async16(void* dst_smem, const void* src, bool pred) {
    int sz = pred ? 16 : 0;
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(smem_u32(dst_smem)), "l"(src), "r"(sz) : "memory");
}
// 16-byte slot, `bytes` (0, 8 or 16) read from global and the rest zero-filled: the last chunk of an odd-length row
__device__ __forceinline__ void cp_async16_n(void* dst_smem, const void* src, int bytes) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(smem_u32(dst_smem)), "l"(src), "r"(bytes) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

// ---- FP64 tensor-core MMA: D(8x8) += A(8x4, row) * B(4x8, col).  SASS: DMMA.8x8x4 ----
// lane l holds  a = A[l/4][l%4],  b = B[l%4][l/4],  c0/c1 = C[l/4][2*(l%4) + {0,1}].
__device__ __forceinline__ void dmma884(double& c0, double& c1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                 : "+d"(c0), "+d"(c1)
                 : "d"(a), "d"(b));
}

#endif  // __CUDACC__
}  // namespace pla
