// FP64 tensor-core GEMM (mma.sync m8n8k4 -> SASS DMMA.8x8x4) for row-major operands, and the
// Gaussian sketch whose operator tiles are generated in shared memory from Philox (never in HBM).
//
// C[M x N] = alpha * op(A) op(B) + beta * C.  128x128x16 CTA tile, 8 warps (2 x 4), warp tile 64 x 32,
// cp.async ring of 3 stages, fragment reads are bank-conflict free by construction of the padded
// shared-memory strides (see frag_* below).  K may be split across gridDim.z; the partial tiles are
// summed in a fixed order by a second kernel, so results are run-to-run deterministic.
//
// Reference call sites replaced: see include/parla_b200.h (pla_gemm_f64 / pla_sketch_gauss_f64).
#include "common.cuh"
#include "philox.cuh"
#include "../../include/parla_b200.h"

namespace pla {

constexpr int GM_BM = 128, GM_BN = 128, GM_BK = 16;
constexpr int GM_THREADS = 256;
constexpr int GM_STAGES = 3;
constexpr int GM_LDK = GM_BK + 4;      // row stride of an [x][k] tile  (20 doubles: 4r + c distinct mod 16)
constexpr int GM_LDX = GM_BM + 4;      // row stride of a  [k][x] tile  (132 doubles: 4c + r distinct mod 16)
constexpr int GM_TILE = (GM_BM * GM_LDK > GM_BK * GM_LDX) ? GM_BM * GM_LDK : GM_BK * GM_LDX;   // doubles

struct GemmParams {
    const double* A; long long lda;
    const double* B; long long ldb;
    double* C; long long ldc;
    double* part;                     // split-K partials [splits][M][N] or null
    long long M, N, K;
    long long Nb;                     // columns physically present in B (N = Nb + 1 when xcol != null)
    const double* xcol;               // optional extra logical column of B (length K)
    double alpha, beta;
    int ktiles_per_split;
    // Gaussian-operator mode (A operand generated): S[r][kglobal]
    uint64_t seed; long long col_offset;
};

// ---- tile loaders -----------------------------------------------------------------------------
// Source stored [X][K] (k contiguous): smem tile[x][GM_LDK].   rows x0.., cols k0..
template <bool VEC16>
__device__ __forceinline__ void load_xk(double* tile, const double* __restrict__ src, long long ld, long long X,
                                        long long K, long long x0, long long k0) {
    if (VEC16) {
        // 128 rows x 8 chunks(16B)
        for (int idx = threadIdx.x; idx < GM_BM * (GM_BK / 2); idx += GM_THREADS) {
            const int r = idx >> 3, ch = idx & 7;
            const long long gx = x0 + r, gk = k0 + 2 * ch;
            const bool ok = gx < X && gk < K;
            cp_async16(tile + r * GM_LDK + 2 * ch, ok ? src + gx * ld + gk : src, ok);
        }
    } else {
        for (int idx = threadIdx.x; idx < GM_BM * GM_BK; idx += GM_THREADS) {
            const int r = idx >> 4, c = idx & 15;
            const long long gx = x0 + r, gk = k0 + c;
            const bool ok = gx < X && gk < K;
            cp_async8(tile + r * GM_LDK + c, ok ? src + gx * ld + gk : src, ok);
        }
    }
}
// Source stored [K][X] (x contiguous): smem tile[k][GM_LDX].  `xcol` (optional, length K) is a
// logical extra column at index X (used to sketch [A | b] in one launch).
template <bool VEC16>
__device__ __forceinline__ void load_kx(double* tile, const double* __restrict__ src, long long ld, long long X,
                                        long long K, long long x0, long long k0,
                                        const double* __restrict__ xcol = nullptr) {
    if (VEC16) {
        // 16 rows x 64 chunks(16B)
        for (int idx = threadIdx.x; idx < GM_BK * (GM_BM / 2); idx += GM_THREADS) {
            const int r = idx >> 6, ch = idx & 63;
            const long long gk = k0 + r, gx = x0 + 2 * ch;
            double* dst = tile + r * GM_LDX + 2 * ch;
            if (gx < X || xcol == nullptr || gx != X) {
                const bool ok = gk < K && gx < X;
                cp_async16(dst, ok ? src + gk * ld + gx : src, ok);
            } else {
                const bool ok = gk < K;
                cp_async8(dst, ok ? xcol + gk : src, ok);
                cp_async8(dst + 1, src, false);
            }
        }
    } else {
        for (int idx = threadIdx.x; idx < GM_BK * GM_BM; idx += GM_THREADS) {
            const int r = idx >> 7, c = idx & 127;
            const long long gk = k0 + r, gx = x0 + c;
            double* dst = tile + r * GM_LDX + c;
            if (xcol != nullptr && gx == X) {
                const bool ok = gk < K;
                cp_async8(dst, ok ? xcol + gk : src, ok);
            } else {
                const bool ok = gk < K && gx < X;
                cp_async8(dst, ok ? src + gk * ld + gx : src, ok);
            }
        }
    }
}
// Generated operator tile: rows = operator rows x0.., k = global column (col_offset + k0 ..). [x][GM_LDK]
__device__ __forceinline__ void gen_xk(double* tile, uint64_t seed, long long col_offset, long long X, long long K,
                                       long long x0, long long k0) {
    // 128 rows x 4 quads; k0 and col_offset are multiples of 4 by construction
    for (int idx = threadIdx.x; idx < GM_BM * (GM_BK / 4); idx += GM_THREADS) {
        const int r = idx >> 2, qd = idx & 3;
        const long long gx = x0 + r, gk = k0 + 4 * qd;
        double g[4] = {0.0, 0.0, 0.0, 0.0};
        if (gx < X && gk < K) {
            philox_normal4(seed, (uint32_t)gx, (uint64_t)(col_offset + gk) >> 2, g);
#pragma unroll
            for (int j = 0; j < 4; ++j) if (gk + j >= K) g[j] = 0.0;
        }
        double* dst = tile + r * GM_LDK + 4 * qd;
        *reinterpret_cast<double2*>(dst) = make_double2(g[0], g[1]);
        *reinterpret_cast<double2*>(dst + 2) = make_double2(g[2], g[3]);
    }
}

// TA: 0 = A stored [M][K], 1 = A stored [K][M], 2 = generated Gaussian operator.
// TB: 0 = B stored [K][N], 1 = B stored [N][K].
template <int TA, int TB, bool VEC16>
__global__ void __launch_bounds__(GM_THREADS, 1) gemm_f64_kernel(const GemmParams p) {
    extern __shared__ __align__(16) double gm_smem[];
    double* sA = gm_smem;                               // [STAGES][GM_TILE]
    double* sB = gm_smem + GM_STAGES * GM_TILE;
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    const int wm = wid >> 2, wn = wid & 3;              // warp grid 2 x 4
    const long long m0 = (long long)blockIdx.y * GM_BM, n0 = (long long)blockIdx.x * GM_BN;
    const long long ktiles = (p.K + GM_BK - 1) / GM_BK;
    const long long kt_begin = (long long)blockIdx.z * p.ktiles_per_split;
    long long kt_end = kt_begin + p.ktiles_per_split;
    if (kt_end > ktiles) kt_end = ktiles;
    const int nkt = (int)(kt_end - kt_begin);

    double acc[8][4][2];
#pragma unroll
    for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j][0] = acc[i][j][1] = 0.0;

    auto issue = [&](int kt_local) {
        const int slot = kt_local % GM_STAGES;
        const long long k0 = (kt_begin + kt_local) * GM_BK;
        double* a = sA + slot * GM_TILE;
        double* b = sB + slot * GM_TILE;
        if (TA == 0) load_xk<VEC16>(a, p.A, p.lda, p.M, p.K, m0, k0);
        else if (TA == 1) load_kx<VEC16>(a, p.A, p.lda, p.M, p.K, m0, k0);
        else gen_xk(a, p.seed, p.col_offset, p.M, p.K, m0, k0);
        if (TB == 0) load_kx<VEC16>(b, p.B, p.ldb, p.Nb, p.K, n0, k0, p.xcol);
        else load_xk<VEC16>(b, p.B, p.ldb, p.Nb, p.K, n0, k0);
    };

#pragma unroll
    for (int s = 0; s < GM_STAGES - 1; ++s) {
        if (s < nkt) issue(s);
        cp_async_commit();
    }
    const int fr = lane >> 2, fc = lane & 3;            // fragment row / k (A), k / col (B)
    for (int kt = 0; kt < nkt; ++kt) {
        cp_async_wait<GM_STAGES - 2>();
        __syncthreads();
        if (kt + GM_STAGES - 1 < nkt) issue(kt + GM_STAGES - 1);
        cp_async_commit();
        const double* a = sA + (kt % GM_STAGES) * GM_TILE;
        const double* b = sB + (kt % GM_STAGES) * GM_TILE;
#pragma unroll
        for (int kk = 0; kk < GM_BK; kk += 4) {
            double af[8], bf[4];
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                const int mm = wm * 64 + i * 8 + fr;
                af[i] = (TA == 1) ? a[(kk + fc) * GM_LDX + mm] : a[mm * GM_LDK + kk + fc];
            }
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const int nn = wn * 32 + j * 8 + fr;
                bf[j] = (TB == 0) ? b[(kk + fc) * GM_LDX + nn] : b[nn * GM_LDK + kk + fc];
            }
#pragma unroll
            for (int i = 0; i < 8; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j) dmma884(acc[i][j][0], acc[i][j][1], af[i], bf[j]);
        }
    }
    cp_async_wait<0>();

    // ---- epilogue: lane holds C[m][n..n+1], m = fr, n = 2*fc within each 8x8 fragment
    const bool split = p.part != nullptr;
    double* out = split ? p.part + (size_t)blockIdx.z * p.M * p.N : p.C;
    const long long ldo = split ? p.N : p.ldc;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        const long long gm = m0 + wm * 64 + i * 8 + fr;
        if (gm >= p.M) continue;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const long long gn = n0 + wn * 32 + j * 8 + 2 * fc;
#pragma unroll
            for (int e = 0; e < 2; ++e) {
                if (gn + e < p.N) {
                    double* dst = out + gm * ldo + gn + e;
                    if (split) *dst = acc[i][j][e];
                    else *dst = (p.beta == 0.0) ? p.alpha * acc[i][j][e] : fma(p.alpha, acc[i][j][e], p.beta * *dst);
                }
            }
        }
    }
}

__global__ void __launch_bounds__(256) gemm_splitk_reduce(const double* __restrict__ part, int splits, long long M,
                                                          long long N, double alpha, double beta, double* C,
                                                          long long ldc) {
    const long long total = M * N;
    for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < total;
         idx += (long long)gridDim.x * blockDim.x) {
        double acc = 0.0;
        for (int s = 0; s < splits; ++s) acc += part[(size_t)s * total + idx];
        const long long r = idx / N, c = idx - r * N;
        double* dst = C + r * ldc + c;
        *dst = (beta == 0.0) ? alpha * acc : fma(alpha, acc, beta * *dst);
    }
}

static int choose_splits(long long M, long long N, long long K) {
    const long long tiles = ((M + GM_BM - 1) / GM_BM) * ((N + GM_BN - 1) / GM_BN);
    const long long ktiles = (K + GM_BK - 1) / GM_BK;
    const int sms = num_sms();
    if (tiles >= sms || ktiles < 32) return 1;
    long long s = (2LL * sms + tiles - 1) / tiles;      // aim at ~2 CTAs' worth of work per SM
    const long long max_by_k = ktiles / 16;             // keep >= 16 k-tiles per split
    if (s > max_by_k) s = max_by_k;
    if (s > 64) s = 64;
    if (s < 1) s = 1;
    return (int)s;
}

template <int TA, int TB>
static int launch_gemm(GemmParams& p, void* ws, size_t ws_bytes, bool vec16, cudaStream_t st) {
    const int splits = choose_splits(p.M, p.N, p.K);
    const long long ktiles = (p.K + GM_BK - 1) / GM_BK;
    p.ktiles_per_split = (int)((ktiles + splits - 1) / splits);
    const int zs = (int)((ktiles + p.ktiles_per_split - 1) / p.ktiles_per_split);
    p.part = nullptr;
    if (zs > 1) {
        const size_t need = (size_t)zs * p.M * p.N * sizeof(double);
        if (ws == nullptr || ws_bytes < need) { set_error("gemm: workspace too small (%zu < %zu)", ws_bytes, need); return -15; }
        p.part = (double*)ws;
    }
    dim3 grid((unsigned)((p.N + GM_BN - 1) / GM_BN), (unsigned)((p.M + GM_BM - 1) / GM_BM), (unsigned)zs);
    const size_t smem = (size_t)2 * GM_STAGES * GM_TILE * sizeof(double);
    cudaError_t e;
    if (vec16) {
        e = cudaFuncSetAttribute(gemm_f64_kernel<TA, TB, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e == cudaSuccess) { gemm_f64_kernel<TA, TB, true><<<grid, GM_THREADS, smem, st>>>(p); e = cudaGetLastError(); note_launch(); }
    } else {
        e = cudaFuncSetAttribute(gemm_f64_kernel<TA, TB, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e == cudaSuccess) { gemm_f64_kernel<TA, TB, false><<<grid, GM_THREADS, smem, st>>>(p); e = cudaGetLastError(); note_launch(); }
    }
    if (e != cudaSuccess) { set_error("gemm: launch failed: %s", cudaGetErrorString(e)); return (int)e; }
    if (zs > 1) {
        long long total = p.M * p.N;
        int nb = (int)((total + 255) / 256);
        if (nb > 4 * num_sms()) nb = 4 * num_sms();
        gemm_splitk_reduce<<<nb, 256, 0, st>>>(p.part, zs, p.M, p.N, p.alpha, p.beta, p.C, p.ldc);
        e = cudaGetLastError();
        note_launch();
        if (e != cudaSuccess) { set_error("gemm: reduce launch failed: %s", cudaGetErrorString(e)); return (int)e; }
    }
    return 0;
}

static bool aligned16(const void* ptr, long long ld) { return ((reinterpret_cast<uintptr_t>(ptr) & 15) == 0) && (ld % 2 == 0); }

}  // namespace pla

using namespace pla;

extern "C" size_t pla_gemm_workspace_bytes(int64_t M, int64_t N, int64_t K) {
    const int s = choose_splits(M, N, K);
    return s > 1 ? (size_t)s * M * N * sizeof(double) : 0;
}

extern "C" int pla_gemm_f64(int transa, int transb, int64_t M, int64_t N, int64_t K, double alpha, const double* A,
                            int64_t lda, const double* B, int64_t ldb, double beta, double* C, int64_t ldc, void* ws,
                            size_t ws_bytes, void* stream) {
    PLA_CHECK_ARG(transa == 0 || transa == 1, 1, "transa must be 0/1");
    PLA_CHECK_ARG(transb == 0 || transb == 1, 2, "transb must be 0/1");
    PLA_CHECK_ARG(M >= 1 && N >= 1 && K >= 1, 3, "empty dimension");
    PLA_CHECK_ARG(A != nullptr && lda >= (transa ? M : K), 8, "bad A / lda");
    PLA_CHECK_ARG(B != nullptr && ldb >= (transb ? K : N), 10, "bad B / ldb");
    PLA_CHECK_ARG(C != nullptr && ldc >= N, 13, "bad C / ldc");
    GemmParams p;
    p.A = A; p.lda = lda; p.B = B; p.ldb = ldb; p.C = C; p.ldc = ldc; p.M = M; p.N = N; p.K = K;
    p.alpha = alpha; p.beta = beta; p.seed = 0; p.col_offset = 0; p.part = nullptr; p.Nb = N; p.xcol = nullptr;
    // 16-byte copies need even leading dimensions AND even extents along the contiguous axis
    const bool a_ok = aligned16(A, lda) && ((transa ? M : K) % 2 == 0);
    const bool b_ok = aligned16(B, ldb) && ((transb ? K : N) % 2 == 0);
    const bool vec16 = a_ok && b_ok;
    cudaStream_t st = (cudaStream_t)stream;
    if (!transa && !transb) return launch_gemm<0, 0>(p, ws, ws_bytes, vec16, st);
    if (transa && !transb) return launch_gemm<1, 0>(p, ws, ws_bytes, vec16, st);
    if (!transa && transb) return launch_gemm<0, 1>(p, ws, ws_bytes, vec16, st);
    return launch_gemm<1, 1>(p, ws, ws_bytes, vec16, st);
}

extern "C" size_t pla_sketch_gauss_workspace_bytes(int64_t d, int64_t n, int64_t m) {
    return pla_gemm_workspace_bytes(d, n + 1, m);     // sized for the optional rhs column
}

extern "C" int pla_sketch_gauss_f64(const double* A, int64_t m, int64_t n, int64_t lda, const double* bvec, int64_t d,
                                    uint64_t seed, int64_t col_offset, double scale, double beta, double* out,
                                    int64_t ldo, void* ws, size_t ws_bytes, void* stream) {
    PLA_CHECK_ARG(A != nullptr, 1, "A is null");
    PLA_CHECK_ARG(m >= 1 && n >= 1, 2, "empty A");
    PLA_CHECK_ARG(lda >= n, 4, "lda < n");
    PLA_CHECK_ARG(d >= 1 && d < (1LL << 32), 6, "d out of range");
    PLA_CHECK_ARG(col_offset >= 0 && col_offset % 4 == 0, 8, "col_offset must be a non-negative multiple of 4");
    const int64_t ncols = n + (bvec != nullptr ? 1 : 0);
    PLA_CHECK_ARG(out != nullptr && ldo >= ncols, 11, "bad out / ldo");
    GemmParams p;
    p.A = nullptr; p.lda = 0; p.B = A; p.ldb = lda; p.C = out; p.ldc = ldo; p.M = d; p.N = ncols; p.K = m;
    p.Nb = n; p.xcol = bvec;
    p.alpha = scale; p.beta = beta; p.seed = seed; p.col_offset = col_offset; p.part = nullptr;
    const bool vec16 = aligned16(A, lda) && (n % 2 == 0);
    return launch_gemm<2, 0>(p, ws, ws_bytes, vec16, (cudaStream_t)stream);
}

namespace pla {
__global__ void __launch_bounds__(256) philox_fill_kernel(double* out, long long rows, long long cols, long long ldo,
                                                          uint64_t seed, long long row_offset, long long col_offset,
                                                          double scale) {
    const long long quads = (cols + 3) / 4 + 1;          // +1: col_offset may start mid-quad
    const long long total = rows * quads;
    for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < total;
         idx += (long long)gridDim.x * blockDim.x) {
        const long long r = idx / quads, qi = idx - r * quads;
        const long long q = (col_offset >> 2) + qi;
        double g[4];
        philox_normal4(seed, (uint32_t)(row_offset + r), (uint64_t)q, g);
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const long long c = 4 * q + j - col_offset;
            if (c >= 0 && c < cols) out[r * ldo + c] = scale * g[j];
        }
    }
}
}  // namespace pla

extern "C" int pla_philox_normal_fill_f64(double* out, int64_t rows, int64_t cols, int64_t ldo, uint64_t seed,
                                          int64_t row_offset, int64_t col_offset, double scale, void* stream) {
    PLA_CHECK_ARG(out != nullptr, 1, "out is null");
    PLA_CHECK_ARG(rows >= 1 && cols >= 1, 2, "empty block");
    PLA_CHECK_ARG(ldo >= cols, 4, "ldo < cols");
    PLA_CHECK_ARG(row_offset >= 0 && row_offset + rows <= (1LL << 32), 6, "row range");
    PLA_CHECK_ARG(col_offset >= 0, 7, "col_offset < 0");
    long long total = rows * ((cols + 3) / 4 + 1);
    int nb = (int)((total + 255) / 256);
    if (nb > 8 * num_sms()) nb = 8 * num_sms();
    philox_fill_kernel<<<nb, 256, 0, (cudaStream_t)stream>>>(out, rows, cols, ldo, seed, row_offset, col_offset, scale);
    PLA_LAUNCH_CHECK();
    return 0;
}
