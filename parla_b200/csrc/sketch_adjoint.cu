// Adjoint application of the sketching operators to ONE vector:  out = S^T v  (m outputs).
//
// Needed by the saddle-point driver SPS2 to fold the lower right-hand side c into b
// (parla/drivers/saddlesys.py:291-292, `b_aug[:m] -= S.T @ v[:d]`; scipy.sparse CSC transpose
// product or a dense dgemv in the reference).  Both kernels are deterministic gathers: one thread
// (SJLT) or one thread column (Gaussian) owns an output entry and sums in a fixed order.
#include "common.cuh"
#include "philox.cuh"
#include "../../include/parla_b200.h"

namespace pla {

// out[i] = scale * sum_q signs[i,q] * v[rows[i,q]]     (SJLT in index form, k nonzeros per column)
__global__ void __launch_bounds__(256) sjlt_rmatvec_kernel(const int* __restrict__ rows,
                                                           const signed char* __restrict__ signs, long long m, int k,
                                                           const double* __restrict__ v, double scale,
                                                           double* __restrict__ out) {
    const long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (i >= m) return;
    double acc = 0.0;
    for (int q = 0; q < k; ++q) {
        const double vi = __ldg(v + rows[i * k + q]);
        acc += signs[i * k + q] < 0 ? -vi : vi;
    }
    out[i] = scale * acc;
}

// out[j] = scale * sum_{r < d} G(seed)[r, col_offset + j] * v[r]  for the virtual Gaussian operator of
// philox.cuh.  Block = 32 column quads x GR_SLICES row slices; a warp shares r, so v[r] is a broadcast.
constexpr int GR_SLICES = 8;

__global__ void __launch_bounds__(32 * GR_SLICES) gauss_rmatvec_kernel(long long d, long long m, uint64_t seed,
                                                                       long long col_offset, double scale,
                                                                       const double* __restrict__ v,
                                                                       double* __restrict__ out) {
    __shared__ double part[GR_SLICES][32][4];
    const int tx = threadIdx.x, ty = threadIdx.y;
    const long long ql = (long long)blockIdx.x * 32 + tx;           // quad within this shard
    const uint64_t q = (uint64_t)((col_offset >> 2) + ql);
    double acc[4] = {0.0, 0.0, 0.0, 0.0};
    if (4 * ql < m) {
        for (long long r = ty; r < d; r += GR_SLICES) {
            double g[4];
            philox_normal4(seed, (uint32_t)r, q, g);
            const double vr = __ldg(v + r);
#pragma unroll
            for (int j = 0; j < 4; ++j) acc[j] = fma(g[j], vr, acc[j]);
        }
    }
#pragma unroll
    for (int j = 0; j < 4; ++j) part[ty][tx][j] = acc[j];
    __syncthreads();
    if (ty == 0) {
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            double tot = 0.0;
#pragma unroll
            for (int s = 0; s < GR_SLICES; ++s) tot += part[s][tx][j];
            const long long c = 4 * ql + j;
            if (c < m) out[c] = scale * tot;
        }
    }
}

}  // namespace pla

using namespace pla;

extern "C" int pla_sjlt_rmatvec_f64(const int32_t* rows, const int8_t* signs, int64_t m, int64_t k, int64_t d,
                                    const double* v, double scale, double* out, void* stream) {
    PLA_CHECK_ARG(rows && signs, 1, "null index form");
    PLA_CHECK_ARG(m >= 0 && k >= 1 && d >= 1, 3, "bad shape");
    PLA_CHECK_ARG(v && out, 6, "null vector");
    if (m == 0) return 0;
    sjlt_rmatvec_kernel<<<(unsigned)((m + 255) / 256), 256, 0, (cudaStream_t)stream>>>(
        rows, reinterpret_cast<const signed char*>(signs), m, (int)k, v, scale, out);
    PLA_LAUNCH_CHECK();
    return 0;
}

extern "C" int pla_gauss_rmatvec_f64(int64_t d, int64_t m, uint64_t seed, int64_t col_offset, double scale,
                                     const double* v, double* out, void* stream) {
    PLA_CHECK_ARG(d >= 1 && d <= (1LL << 32), 1, "d out of range");
    PLA_CHECK_ARG(m >= 0, 2, "m < 0");
    PLA_CHECK_ARG(col_offset >= 0 && (col_offset & 3) == 0, 4, "col_offset must be a non-negative multiple of 4");
    PLA_CHECK_ARG(v && out, 6, "null vector");
    if (m == 0) return 0;
    const long long quads = (m + 3) / 4;
    gauss_rmatvec_kernel<<<(unsigned)((quads + 31) / 32), dim3(32, GR_SLICES), 0, (cudaStream_t)stream>>>(
        d, m, seed, col_offset, scale, v, out);
    PLA_LAUNCH_CHECK();
    return 0;
}
