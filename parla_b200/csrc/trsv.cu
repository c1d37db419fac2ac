// K3b: upper-triangular solve with one right-hand side (R row-major n x n).
//
// Replaces scipy.linalg.solve_triangular(R, ., lower=False [, trans='T']) on the LSQR path
// (parla/comps/preconditioning.py:28,37,40,41; parla/drivers/least_squares.py:316,363).
// One CTA, blocked by 32: the sequential 32x32 diagonal solves run in one warp on registers /
// shuffles; the off-diagonal products are spread over all 32 warps with coalesced row reads of R
// (R is L2-resident: n^2*8 bytes = 32 MiB at n = 2048).
#include "common.cuh"
#include "../../include/parla_b200.h"

namespace pla {

constexpr int TS_THREADS = 1024;
constexpr int TS_NB = 32;

// trans = 0: back substitution, left-looking by block rows.
__global__ void __launch_bounds__(TS_THREADS) trsv_upper_n_kernel(const double* __restrict__ R, long long n,
                                                                  long long ldr, const double* b, double* xout,
                                                                  const int* istop) {
    if (istop != nullptr && *istop != 0) return;
    extern __shared__ double ts_smem[];
    double* x = ts_smem;                       // n (padded to multiple of 32)
    double* diag = x + ((n + 31) / 32) * 32;   // 32 x 33
    double* rhs = diag + 32 * 33;              // 32
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    const long long nblk = (n + TS_NB - 1) / TS_NB;
    for (long long kb = nblk - 1; kb >= 0; --kb) {
        const long long i0 = kb * TS_NB;
        const long long i = i0 + wid;                      // warp <-> row of the block
        double acc = 0.0;
        if (i < n) {
            const double* row = R + i * ldr;
            for (long long c = i0 + TS_NB + lane; c < n; c += 32) acc = fma(row[c], x[c], acc);
            // diagonal block row (coalesced), identity padding outside the matrix
            const long long c = i0 + lane;
            diag[wid * 33 + lane] = (c < n && c >= i) ? row[c] : 0.0;
        } else {
            diag[wid * 33 + lane] = (wid == lane) ? 1.0 : 0.0;
        }
        acc = warp_sum(acc);
        if (lane == 0) rhs[wid] = (i < n ? b[i] : 0.0) - acc;
        __syncthreads();
        if (wid == 0) {
            double r = rhs[lane];
            for (int j = TS_NB - 1; j >= 0; --j) {
                const double xj = __shfl_sync(0xffffffffu, r, j) / diag[j * 33 + j];
                if (lane == j) r = xj;
                else if (lane < j) r = fma(-diag[lane * 33 + j], xj, r);
            }
            x[i0 + lane] = r;
        }
        __syncthreads();
    }
    for (long long c = threadIdx.x; c < n; c += blockDim.x) xout[c] = x[c];
}

// trans = 1: forward substitution on R^T, right-looking (thread <-> column keeps its own rhs entry).
__global__ void __launch_bounds__(TS_THREADS) trsv_upper_t_kernel(const double* __restrict__ R, long long n,
                                                                  long long ldr, const double* b, double* xout,
                                                                  const int* istop) {
    if (istop != nullptr && *istop != 0) return;
    extern __shared__ double ts_smem[];
    double* x = ts_smem;                       // running rhs, becomes the solution
    double* diag = x + ((n + 31) / 32) * 32;
    double* xk = diag + 32 * 33;               // solved block
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    for (long long c = threadIdx.x; c < n; c += blockDim.x) x[c] = b[c];
    const long long nblk = (n + TS_NB - 1) / TS_NB;
    for (long long kb = 0; kb < nblk; ++kb) {
        const long long i0 = kb * TS_NB;
        {   // load the diagonal block: warp <-> row
            const long long i = i0 + wid, c = i0 + lane;
            double v = (wid == lane) ? 1.0 : 0.0;
            if (i < n && c < n) v = (c >= i) ? R[i * ldr + c] : 0.0;
            diag[wid * 33 + lane] = v;
        }
        __syncthreads();
        if (wid == 0) {
            // solve (R_kk)^T y = rhs : y_j = (rhs_j - sum_{i<j} R[i][j] y_i) / R[j][j]
            double r = (i0 + lane < n) ? x[i0 + lane] : 0.0;
            for (int j = 0; j < TS_NB; ++j) {
                const double yj = __shfl_sync(0xffffffffu, r, j) / diag[j * 33 + j];
                if (lane == j) r = yj;
                else if (lane > j) r = fma(-diag[j * 33 + lane], yj, r);
            }
            xk[lane] = r;
            if (i0 + lane < n) x[i0 + lane] = r;
        }
        __syncthreads();
        // trailing update: rhs[c] -= sum_i R[i0+i][c] * y_i for c beyond this block
        const long long rows = min((long long)TS_NB, n - i0);
        for (long long c = i0 + TS_NB + threadIdx.x; c < n; c += blockDim.x) {
            double acc = x[c];
            for (long long i = 0; i < rows; ++i) acc = fma(-R[(i0 + i) * ldr + c], xk[i], acc);
            x[c] = acc;
        }
        __syncthreads();
    }
    for (long long c = threadIdx.x; c < n; c += blockDim.x) xout[c] = x[c];
}

// Inverse of each 32x32 diagonal block of an upper-triangular matrix (base case of the blocked
// triangular inversion; the off-diagonal blocks are filled by DMMA GEMMs from the host side:
// X12 = -X11 (R12 X22)).  One CTA per block, thread j <-> column j (back substitution in smem).
constexpr int TI_BS = 32;
__global__ void __launch_bounds__(TI_BS) trtri_diag_kernel(const double* __restrict__ R, long long n, long long ldr,
                                                           double* X, long long ldx) {
    __shared__ double Rb[TI_BS][TI_BS + 1];
    __shared__ double Xb[TI_BS][TI_BS + 1];
    const long long i0 = (long long)blockIdx.x * TI_BS;
    const int j = threadIdx.x;
    const int bs = (int)min((long long)TI_BS, n - i0);
    for (int r = 0; r < TI_BS; ++r) {
        double v = (r == j) ? 1.0 : 0.0;                       // identity padding past the matrix edge
        if (r < bs && j < bs) v = (j >= r) ? R[(i0 + r) * ldr + i0 + j] : 0.0;
        Rb[r][j] = v;
    }
    __syncthreads();
    for (int i = TI_BS - 1; i >= 0; --i) {
        double acc = (i == j) ? 1.0 : 0.0;
        if (i <= j) {
            for (int l = i + 1; l <= j; ++l) acc = fma(-Rb[i][l], Xb[l][j], acc);
            acc /= Rb[i][i];
        } else {
            acc = 0.0;
        }
        Xb[i][j] = acc;                                        // column j only reads its own column of Xb
    }
    __syncthreads();
    for (int r = 0; r < bs; ++r)
        if (j < bs) X[(i0 + r) * ldx + i0 + j] = Xb[r][j];
}

// One level of the recursion for SMALL blocks, all pairs of the level in ONE launch: with the s x s diagonal blocks
// X11, X22 already inverted, X12 = -X11 (R12 X22).  CTA <-> pair; the three blocks live in shared memory.  (Issued as
// individual GEMMs from the host the levels s = 32, 64 cost ~100 launches of a few microseconds each.)
constexpr int TM_THREADS = 256;
__global__ void __launch_bounds__(TM_THREADS) trtri_merge_kernel(const double* __restrict__ R, long long n, long long ldr,
                                                                 double* X, long long ldx, int s) {
    extern __shared__ double tm_smem[];
    const int ld = s + 1;
    double* X11 = tm_smem;                 // s x s upper triangular
    double* R12 = X11 + s * ld;            // s x w
    double* X22 = R12 + s * ld;            // w x w upper triangular   (then T = R12 X22 overwrites R12)
    const long long i0 = (long long)blockIdx.x * 2 * s, i1 = i0 + s;
    if (i1 >= n) return;
    const int w = (int)min((long long)s, n - i1);
    for (int idx = threadIdx.x; idx < s * s; idx += TM_THREADS) {
        const int r = idx / s, c = idx - r * s;
        X11[r * ld + c] = (c >= r) ? X[(i0 + r) * ldx + i0 + c] : 0.0;
        R12[r * ld + c] = (c < w) ? R[(i0 + r) * ldr + i1 + c] : 0.0;
        X22[r * ld + c] = (r < w && c < w && c >= r) ? X[(i1 + r) * ldx + i1 + c] : 0.0;
    }
    __syncthreads();
    // T = R12 X22 (X22 upper triangular: l <= c); every thread keeps its outputs in registers until all are computed
    double t[16];
    int cnt = 0;
    for (int idx = threadIdx.x; idx < s * w; idx += TM_THREADS, ++cnt) {
        const int r = idx / w, c = idx - r * w;
        double acc = 0.0;
        for (int l = 0; l <= c; ++l) acc = fma(R12[r * ld + l], X22[l * ld + c], acc);
        t[cnt] = acc;
    }
    __syncthreads();
    cnt = 0;
    for (int idx = threadIdx.x; idx < s * w; idx += TM_THREADS, ++cnt) {
        const int r = idx / w, c = idx - r * w;
        R12[r * ld + c] = t[cnt];
    }
    __syncthreads();
    // X12 = -X11 T (X11 upper triangular: l >= r)
    for (int idx = threadIdx.x; idx < s * w; idx += TM_THREADS) {
        const int r = idx / w, c = idx - r * w;
        double acc = 0.0;
        for (int l = r; l < s; ++l) acc = fma(X11[r * ld + l], R12[l * ld + c], acc);
        X[(i0 + r) * ldx + i1 + c] = -acc;
    }
}

}  // namespace pla

using namespace pla;

extern "C" int pla_trtri_merge_f64(const double* R, int64_t n, int64_t ldr, double* X, int64_t ldx, int64_t s, void* stream) {
    PLA_CHECK_ARG(R != nullptr, 1, "R is null");
    PLA_CHECK_ARG(n >= 1 && ldr >= n, 3, "bad n / ldr");
    PLA_CHECK_ARG(X != nullptr && ldx >= n, 4, "bad X / ldx");
    PLA_CHECK_ARG(s == 32 || s == 64, 6, "level size must be 32 or 64");
    const long long pairs = (n + 2 * s - 1) / (2 * s);
    const size_t smem = (size_t)3 * s * (s + 1) * sizeof(double);
    PLA_CUDA(cudaFuncSetAttribute(trtri_merge_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    trtri_merge_kernel<<<(unsigned)pairs, TM_THREADS, smem, (cudaStream_t)stream>>>(R, n, ldr, X, ldx, (int)s);
    PLA_LAUNCH_CHECK();
    return 0;
}

extern "C" int pla_trtri_diag_f64(const double* R, int64_t n, int64_t ldr, double* X, int64_t ldx, void* stream) {
    PLA_CHECK_ARG(R != nullptr, 1, "R is null");
    PLA_CHECK_ARG(n >= 1, 2, "n < 1");
    PLA_CHECK_ARG(ldr >= n, 3, "ldr < n");
    PLA_CHECK_ARG(X != nullptr && ldx >= n, 4, "bad X / ldx");
    trtri_diag_kernel<<<(unsigned)((n + TI_BS - 1) / TI_BS), TI_BS, 0, (cudaStream_t)stream>>>(R, n, ldr, X, ldx);
    PLA_LAUNCH_CHECK();
    return 0;
}


extern "C" int pla_trsv_upper_f64(const double* R, int64_t n, int64_t ldr, int trans, const double* b, double* x,
                                  const int* istop_dev, void* stream) {
    PLA_CHECK_ARG(R != nullptr, 1, "R is null");
    PLA_CHECK_ARG(n >= 1 && n <= 24576, 2, "n out of range (1..24576)");
    PLA_CHECK_ARG(ldr >= n, 3, "ldr < n");
    PLA_CHECK_ARG(trans == 0 || trans == 1, 4, "trans must be 0 or 1");
    PLA_CHECK_ARG(b != nullptr && x != nullptr, 5, "null vector");
    const size_t smem = (size_t)(((n + 31) / 32) * 32 + 32 * 33 + 32) * sizeof(double);
    cudaStream_t st = (cudaStream_t)stream;
    if (trans == 0) {
        PLA_CUDA(cudaFuncSetAttribute(trsv_upper_n_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        trsv_upper_n_kernel<<<1, TS_THREADS, smem, st>>>(R, n, ldr, b, x, istop_dev);
    } else {
        PLA_CUDA(cudaFuncSetAttribute(trsv_upper_t_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        trsv_upper_t_kernel<<<1, TS_THREADS, smem, st>>>(R, n, ldr, b, x, istop_dev);
    }
    PLA_LAUNCH_CHECK();
    return 0;
}
