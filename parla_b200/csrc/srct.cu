// Subsampled randomized cosine transform (SRCT) building blocks.
//
// The reference applies  S = R . DCT-II(ortho) . diag(e) . P  with scipy.fft.dct over ALL m rows and then
// keeps the d sampled rows (parla/utils/sketching.py:106-176, `apply_srct`).  Only d << m output rows are
// needed, so the transform is evaluated as a PRUNED two-level DCT built from FP64 tensor-core GEMMs
// (host side: parla_b200/utils/sketching.py::SRCTOperator):
//     j = j1 + m1 j2 :  cos((2j+1) t_k) = cos((2 j1 + 1) t_k) cos(pi k j2 / m2) - sin((2 j1 + 1) t_k) sin(pi k j2 / m2)
//   level 1: Y = F (2 m2 x m2 cos/sin table) times the permuted, sign-flipped A viewed as m2 x (m1 n)   [pla_gemm_f64]
//   level 2: one GEMM per frequency class k mod 2 m2 with the weights generated HERE.
// All angles are reduced in exact integer arithmetic ((2j+1) k mod 4m) before cospi/sinpi, so every
// weight is correct to an ulp for any m < 2^31.
#include "common.cuh"
#include "../../include/parla_b200.h"

namespace pla {

// out[i, c]          =  f(k_i) cos(pi (2 j + 1) k_i / (2m)) * e[j]              j = jmap ? jmap[j0 + c] : j0 + c
// out[i, ncols + c]  = -sgn_i f(k_i) sin(pi (2 j + 1) k_i / (2m)) * e[j]        (with_sin only)
// f(0) = sqrt(1/m), f(k > 0) = sqrt(2/m)   (scipy.fft.dct type 2, norm='ortho')
__global__ void __launch_bounds__(256) srct_weights_kernel(const long long* __restrict__ k, long long g, long long m,
                                                           long long j0, long long ncols,
                                                           const long long* __restrict__ jmap,
                                                           const double* __restrict__ e,
                                                           const double* __restrict__ sgn, int with_sin,
                                                           double* __restrict__ out, long long ldo) {
    const long long total = g * ncols;
    const unsigned long long four_m = 4ULL * (unsigned long long)m;
    const double inv_2m = 1.0 / (2.0 * (double)m);
    for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < total;
         idx += (long long)gridDim.x * blockDim.x) {
        const long long i = idx / ncols, c = idx - i * ncols;
        const long long j = jmap ? jmap[j0 + c] : j0 + c;
        const unsigned long long ki = (unsigned long long)k[i];
        const unsigned long long p = ((2ULL * (unsigned long long)j + 1ULL) * ki) % four_m;   // exact, < 2^63
        const double x = (double)p * inv_2m;                                                 // angle / pi in [0, 2)
        double f = (ki == 0ULL) ? sqrt(1.0 / (double)m) : sqrt(2.0 / (double)m);
        if (e) f *= e[j];
        double sn, cs;
        sincospi(x, &sn, &cs);
        out[i * ldo + c] = f * cs;
        if (with_sin) out[i * ldo + ncols + c] = -(sgn ? sgn[i] : 1.0) * f * sn;
    }
}

// out[t, c] = e[t] * A[perm[t], c0 + c],  t in [0, rows), c in [0, nb)   (out row-major, leading dimension ldo)
__global__ void __launch_bounds__(256) gather_rows_scale_kernel(const double* __restrict__ A, long long lda,
                                                                const long long* __restrict__ perm,
                                                                const double* __restrict__ e, long long rows,
                                                                long long c0, long long nb, double* __restrict__ out,
                                                                long long ldo, int vec2) {
    const int lane = threadIdx.x & 31;
    const long long warp = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const long long nwarps = ((long long)gridDim.x * blockDim.x) >> 5;
    for (long long t = warp; t < rows; t += nwarps) {
        const double s = e ? e[t] : 1.0;
        const double* src = A + (perm ? perm[t] : t) * lda + c0;
        double* dst = out + t * ldo;
        if (vec2) {
            for (long long c = 2 * lane; c < nb; c += 64) {
                const double2 a = __ldg(reinterpret_cast<const double2*>(src + c));
                *reinterpret_cast<double2*>(dst + c) = make_double2(s * a.x, s * a.y);
            }
        } else {
            for (long long c = lane; c < nb; c += 32) dst[c] = s * __ldg(src + c);
        }
    }
}

}  // namespace pla

using namespace pla;

extern "C" int pla_srct_weights_f64(const int64_t* k, int64_t g, int64_t m, int64_t j0, int64_t ncols,
                                    const int64_t* jmap, const double* e, const double* sgn, int with_sin,
                                    double* out, int64_t ldo, void* stream) {
    PLA_CHECK_ARG(k != nullptr && g >= 1, 1, "no frequencies");
    PLA_CHECK_ARG(m >= 1 && m < (1LL << 31), 3, "m out of range (1 .. 2^31)");
    PLA_CHECK_ARG(j0 >= 0 && ncols >= 1, 4, "bad column range");
    PLA_CHECK_ARG(out != nullptr && ldo >= (with_sin ? 2 : 1) * ncols, 10, "bad out / ldo");
    const long long total = g * ncols;
    long long nb = (total + 255) / 256;
    if (nb > 16LL * num_sms()) nb = 16LL * num_sms();
    srct_weights_kernel<<<(unsigned)nb, 256, 0, (cudaStream_t)stream>>>(
        reinterpret_cast<const long long*>(k), g, m, j0, ncols, reinterpret_cast<const long long*>(jmap), e, sgn,
        with_sin, out, ldo);
    PLA_LAUNCH_CHECK();
    return 0;
}

extern "C" int pla_gather_rows_scale_f64(const double* A, int64_t lda, const int64_t* perm, const double* e,
                                         int64_t rows, int64_t c0, int64_t nb, double* out, int64_t ldo,
                                         void* stream) {
    PLA_CHECK_ARG(A != nullptr && lda >= c0 + nb, 1, "bad A / lda");
    PLA_CHECK_ARG(rows >= 1 && c0 >= 0 && nb >= 1, 5, "bad block");
    PLA_CHECK_ARG(out != nullptr && ldo >= nb, 8, "bad out / ldo");
    const bool vec2 = (nb % 2 == 0) && (lda % 2 == 0) && (ldo % 2 == 0) && (c0 % 2 == 0) &&
                      ((reinterpret_cast<uintptr_t>(A) & 15) == 0) && ((reinterpret_cast<uintptr_t>(out) & 15) == 0);
    long long ctas = (rows + 7) / 8;
    if (ctas > 32LL * num_sms()) ctas = 32LL * num_sms();
    gather_rows_scale_kernel<<<(unsigned)ctas, 256, 0, (cudaStream_t)stream>>>(
        A, lda, reinterpret_cast<const long long*>(perm), e, rows, c0, nb, out, ldo, vec2 ? 1 : 0);
    PLA_LAUNCH_CHECK();
    return 0;
}
