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
    int ctas_in_flight;               // CACC kernels: tiles computed concurrently (one per SM); 0 = no next-tile prefetch
    // Gaussian-operator mode (A operand generated): S[r][kglobal]
    uint64_t seed; long long col_offset;
};

// ---- tile loaders -----------------------------------------------------------------------------
// Each thread copies the same (row, chunk) slots of every k-tile, so addresses are one base pointer
// plus compile-time multiples of the leading dimension and the bounds tests are hoisted.
// Source stored [X][K] (k contiguous): smem tile[x][GM_LDK].   rows x0.., cols k0..
// VEC16: 0 = 8-byte copies; 1 = 16-byte copies, even extent along the contiguous axis; 2 = 16-byte copies whose last
// slot of an odd-length row is half filled (cp.async src-size 8).  (2 costs a select per copy: 5 % on the big products.)
template <int VEC16>
__device__ __forceinline__ void load_xk(double* tile, const double* __restrict__ src, long long ld, long long X,
                                        long long K, long long x0, long long k0) {
    const int tid = threadIdx.x;
    if (VEC16) {                       // 128 rows x 8 chunks(16B): thread -> chunk tid&7 of rows (tid>>3) + 32 i
        const int r0 = tid >> 3, c2 = 2 * (tid & 7);
        const int kbytes = (VEC16 == 1 || k0 + c2 + 1 < K) ? 16 : 8;
        const bool kok = k0 + c2 < K;
        const double* p = src + (x0 + r0) * ld + k0 + c2;
        double* d = tile + r0 * GM_LDK + c2;
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const bool ok = kok && (x0 + r0 + 32 * i < X);
            if (VEC16 == 1) cp_async16(d + 32 * i * GM_LDK, ok ? p + 32 * i * ld : src, ok);
            else cp_async16_n(d + 32 * i * GM_LDK, ok ? p + 32 * i * ld : src, ok ? kbytes : 0);
        }
    } else {                           // 128 rows x 16 doubles: thread -> column tid&15 of rows (tid>>4) + 16 i
        const int r0 = tid >> 4, c = tid & 15;
        const bool kok = k0 + c < K;
        const double* p = src + (x0 + r0) * ld + k0 + c;
        double* d = tile + r0 * GM_LDK + c;
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            const bool ok = kok && (x0 + r0 + 16 * i < X);
            cp_async8(d + 16 * i * GM_LDK, ok ? p + 16 * i * ld : src, ok);
        }
    }
}
// Source stored [K][X] (x contiguous): smem tile[k][GM_LDX].  `xcol` (optional, length K) is a
// logical extra column at index X (used to sketch [A | b] in one launch).
template <int VEC16>
__device__ __forceinline__ void load_kx(double* tile, const double* __restrict__ src, long long ld, long long X,
                                        long long K, long long x0, long long k0,
                                        const double* __restrict__ xcol = nullptr) {
    const int tid = threadIdx.x;
    if (VEC16) {                       // 16 rows x 64 chunks(16B): thread -> chunk tid&63 of rows (tid>>6) + 4 i
        const int r0 = tid >> 6, c2 = 2 * (tid & 63);
        const long long gx = x0 + c2;
        const int xbytes = (VEC16 == 1 || gx + 1 < X) ? 16 : 8;               // odd X (xcol == null): half a slot
        const bool xok = gx < X;
        const bool extra = (xcol != nullptr) && (gx == X);
        const double* p = src + (k0 + r0) * ld + gx;
        double* d = tile + r0 * GM_LDX + c2;
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const bool kok = k0 + r0 + 4 * i < K;
            if (!extra) {
                const bool ok = xok && kok;
                if (VEC16 == 1) cp_async16(d + 4 * i * GM_LDX, ok ? p + 4 * i * ld : src, ok);
                else cp_async16_n(d + 4 * i * GM_LDX, ok ? p + 4 * i * ld : src, ok ? xbytes : 0);
            } else {
                cp_async8(d + 4 * i * GM_LDX, kok ? xcol + k0 + r0 + 4 * i : src, kok);
                cp_async8(d + 4 * i * GM_LDX + 1, src, false);
            }
        }
    } else {                           // 16 rows x 128 doubles: thread -> column tid&127 of rows (tid>>7) + 2 i
        const int r0 = tid >> 7, c = tid & 127;
        const long long gx = x0 + c;
        const bool xok = gx < X;
        const bool extra = (xcol != nullptr) && (gx == X);
        const double* p = src + (k0 + r0) * ld + gx;
        double* d = tile + r0 * GM_LDX + c;
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            const bool kok = k0 + r0 + 2 * i < K;
            if (!extra) {
                const bool ok = xok && kok;
                cp_async8(d + 2 * i * GM_LDX, ok ? p + 2 * i * ld : src, ok);
            } else {
                cp_async8(d + 2 * i * GM_LDX, kok ? xcol + k0 + r0 + 2 * i : src, kok);
            }
        }
    }
}
// Generated operator tile: rows = operator rows x0.., k = global column (col_offset + k0 ..). [x][GM_LDK]
// half = 0/1 generates rows [0,64) / [64,128) of the tile (one Philox block per thread), so the
// generator's dependent integer chain can be split around DMMA groups; half < 0 does both.
__device__ __forceinline__ void gen_xk(double* tile, uint64_t seed, long long col_offset, long long X, long long K,
                                       long long x0, long long k0, int half) {
    // 128 rows x 4 quads; thread -> quad tid&3 of rows (tid>>2) + 64 i.  k0, col_offset multiples of 4.
    const int tid = threadIdx.x;
    const int r0 = tid >> 2, qd = tid & 3;
    const long long gk = k0 + 4 * qd;
    const uint64_t q = (uint64_t)(col_offset + gk) >> 2;
#pragma unroll
    for (int i = 0; i < 2; ++i) {
        if (half >= 0 && half != i) continue;
        const long long gx = x0 + r0 + 64 * i;
        double g[4] = {0.0, 0.0, 0.0, 0.0};
        if (gx < X && gk < K) {
            philox_normal4(seed, (uint32_t)gx, q, g);
            if (gk + 4 > K) {
#pragma unroll
                for (int j = 0; j < 4; ++j) if (gk + j >= K) g[j] = 0.0;
            }
        }
        double* dst = tile + (r0 + 64 * i) * GM_LDK + 4 * qd;
        *reinterpret_cast<double2*>(dst) = make_double2(g[0], g[1]);
        *reinterpret_cast<double2*>(dst + 2) = make_double2(g[2], g[3]);
    }
}

// TA: 0 = A stored [M][K], 1 = A stored [K][M], 2 = generated Gaussian operator.
// TB: 0 = B stored [K][N], 1 = B stored [N][K].
// CACC: the beta != 0 update C = alpha op(A) op(B) + beta C without split-K, with a prologue / epilogue built for it
// (the rank-128 block-reflector update of the QR).  It is a separate instantiation because any change to the
// prologue / epilogue code perturbs ptxas' schedule of the main loop of the other variants by +-5 %.
template <int TA, int TB, int VEC16, bool CACC>
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
    // beta != 0: the accumulators START at (beta / alpha) C, so the 64 loads of C per thread are in flight together
    // with the first operand tiles and the epilogue only stores alpha * acc.  (Reading C in the epilogue serialised
    // 64 load -> fma -> store round trips per thread: the K = 128 block-reflector update of the QR ran at 12-20 TF.)
    const bool c_in_acc = p.part == nullptr && p.beta != 0.0 && p.alpha != 0.0;
    if (!CACC && c_in_acc) {
        const double f = p.beta / p.alpha;
        const int fr0 = lane >> 2, fc0 = lane & 3;
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            const long long gm = m0 + wm * 64 + i * 8 + fr0;
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const long long gn = n0 + wn * 32 + j * 8 + 2 * fc0;
#pragma unroll
                for (int e = 0; e < 2; ++e)
                    if (gm < p.M && gn + e < p.N) acc[i][j][e] = f * p.C[gm * p.ldc + gn + e];
            }
        }
    }
    // CACC: interior tiles (the CTA-uniform common case) use unpredicated, 16-byte accesses when C allows them: with
    // per-element predicates ptxas issued the loads in small dependent groups (14 us per tile against a 19 us main loop).
    const bool interior = (m0 + GM_BM <= p.M) && (n0 + GM_BN <= p.N);
    const bool c_vec = interior && ((reinterpret_cast<uintptr_t>(p.C) & 15) == 0) && ((p.ldc & 1) == 0);
    const long long c_off = (m0 + wm * 64 + (lane >> 2)) * p.ldc + n0 + wn * 32 + 2 * (lane & 3);   // fragment (0, 0)
    if (CACC) {
        if (c_vec) {
            const double* cp = p.C + c_off;
#pragma unroll
            for (int i = 0; i < 8; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    const double2 v = *reinterpret_cast<const double2*>(cp + (long long)(i * 8) * p.ldc + j * 8);
                    acc[i][j][0] = v.x; acc[i][j][1] = v.y;
                }
        } else if (interior) {
            const double* cp = p.C + c_off;
#pragma unroll
            for (int i = 0; i < 8; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    acc[i][j][0] = cp[(long long)(i * 8) * p.ldc + j * 8];
                    acc[i][j][1] = cp[(long long)(i * 8) * p.ldc + j * 8 + 1];
                }
        } else {
            const int fr0 = lane >> 2, fc0 = lane & 3;
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                const long long gm = m0 + wm * 64 + i * 8 + fr0;
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    const long long gn = n0 + wn * 32 + j * 8 + 2 * fc0;
#pragma unroll
                    for (int e = 0; e < 2; ++e)
                        if (gm < p.M && gn + e < p.N) acc[i][j][e] = p.C[gm * p.ldc + gn + e];
                }
            }
        }
    }

    // The two operand tiles of a k-step are fetched by separate helpers so that their address /
    // predicate / Philox instructions can be interleaved with DMMA groups (tensor pipe stays busy).
    auto issue_a = [&](int kt_local, int half) {        // half: -1 whole tile, 0/1 first/second part
        const long long k0 = (kt_begin + kt_local) * GM_BK;
        double* a = sA + (kt_local % GM_STAGES) * GM_TILE;
        if (TA == 2) { gen_xk(a, p.seed, p.col_offset, p.M, p.K, m0, k0, half); return; }
        if (half > 0) return;                            // copied tiles are issued in one piece
        if (TA == 0) load_xk<VEC16>(a, p.A, p.lda, p.M, p.K, m0, k0);
        else load_kx<VEC16>(a, p.A, p.lda, p.M, p.K, m0, k0);
    };
    auto issue_b = [&](int kt_local) {
        const long long k0 = (kt_begin + kt_local) * GM_BK;
        double* b = sB + (kt_local % GM_STAGES) * GM_TILE;
        if (TB == 0) load_kx<VEC16>(b, p.B, p.ldb, p.Nb, p.K, n0, k0, p.xcol);
        else load_xk<VEC16>(b, p.B, p.ldb, p.Nb, p.K, n0, k0);
    };

#pragma unroll
    for (int s = 0; s < GM_STAGES - 1; ++s) {
        if (s < nkt) { issue_a(s, -1); issue_b(s); }
        cp_async_commit();
    }
    if (CACC) {
        // The C tile of the CTA that will follow on this SM (linear tile id + #CTAs in flight = one per SM) is pulled
        // into L2 while this tile computes, so its 128 KB of loads do not start a wave with a DRAM round trip.
        const long long nxt = (long long)blockIdx.y * gridDim.x + blockIdx.x + p.ctas_in_flight;
        const long long ny = nxt / gridDim.x, nx = nxt - ny * gridDim.x;
        if (p.ctas_in_flight > 0 && ny < gridDim.y) {
            const long long pm0 = ny * GM_BM, pn0 = nx * GM_BN;
#pragma unroll
            for (int i = 0; i < 4; ++i) {                    // 128 rows x 8 lines of 128 bytes
                const int line = tid + GM_THREADS * i;
                const long long r = pm0 + (line >> 3), cc = pn0 + 16 * (line & 7);
                if (r < p.M && cc < p.N) asm volatile("prefetch.global.L2 [%0];" ::"l"(p.C + r * p.ldc + cc));
            }
        }
    }
    if (CACC) {                                          // scale once the operand copies are on their way
        const double f = p.beta / p.alpha;
#pragma unroll
        for (int i = 0; i < 8; ++i)
#pragma unroll
            for (int j = 0; j < 4; ++j) { acc[i][j][0] *= f; acc[i][j][1] *= f; }
    }
    const int fr = lane >> 2, fc = lane & 3;            // fragment row / k (A), k / col (B)
    for (int kt = 0; kt < nkt; ++kt) {
        cp_async_wait<GM_STAGES - 2>();
        __syncthreads();                                 // tile kt landed; slot (kt-1)%STAGES is free again
        const double* a = sA + (kt % GM_STAGES) * GM_TILE;
        const double* b = sB + (kt % GM_STAGES) * GM_TILE;
        const bool more = kt + GM_STAGES - 1 < nkt;
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
            // prefetch of tile kt+STAGES-1 rides in the shadow of the DMMAs just issued
            if (kk == 0 && more) issue_a(kt + GM_STAGES - 1, 0);
            if (kk == 4 && more) issue_b(kt + GM_STAGES - 1);
            if (kk == 4) cp_async_commit();
            if (kk == 8 && more) issue_a(kt + GM_STAGES - 1, 1);
        }
    }
    cp_async_wait<0>();

    // ---- epilogue: lane holds C[m][n..n+1], m = fr, n = 2*fc within each 8x8 fragment
    if (CACC) {
        // CTA-uniform shapes (per-element branching on split / beta made ptxas reload alpha from constant memory
        // before every store)
        const double alpha = p.alpha;
        double* cp = p.C + c_off;
        if (c_vec) {
#pragma unroll
            for (int i = 0; i < 8; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j)
                    *reinterpret_cast<double2*>(cp + (long long)(i * 8) * p.ldc + j * 8) =
                        make_double2(alpha * acc[i][j][0], alpha * acc[i][j][1]);
        } else if (interior) {
#pragma unroll
            for (int i = 0; i < 8; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    cp[(long long)(i * 8) * p.ldc + j * 8] = alpha * acc[i][j][0];
                    cp[(long long)(i * 8) * p.ldc + j * 8 + 1] = alpha * acc[i][j][1];
                }
        } else {
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                const long long gm = m0 + wm * 64 + i * 8 + fr;
                if (gm >= p.M) continue;
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    const long long gn = n0 + wn * 32 + j * 8 + 2 * fc;
#pragma unroll
                    for (int e = 0; e < 2; ++e)
                        if (gn + e < p.N) p.C[gm * p.ldc + gn + e] = alpha * acc[i][j][e];
                }
            }
        }
        return;
    }
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
                    else if (c_in_acc || p.beta == 0.0) *dst = p.alpha * acc[i][j][e];
                    else *dst = fma(p.alpha, acc[i][j][e], p.beta * *dst);
                }
            }
        }
    }
}

// ---- Gaussian sketch, dedicated tile shape -------------------------------------------------------
// out[M x N] (+)= G(seed)[M x K] * B[K x N] with a 64 x 256 x 16 CTA tile (8 warps side by side, warp
// tile 64 x 32): the generated 64 x 16 operator tile is amortised over 256 output columns, i.e. half
// the Philox / Box-Muller work per FMA of the square 128 x 128 tile, and one Philox block per thread
// per k-step hides completely behind the DMMA groups.
constexpr int GS_BM = 64, GS_BN = 256, GS_STAGES = 4;
constexpr int GS_LDX = GS_BN + 4;                               // 260: 4c + r distinct mod 16
constexpr int GS_TILE_A = GS_BM * GM_LDK, GS_TILE_B = GM_BK * GS_LDX;

template <bool VEC16>
__device__ __forceinline__ void gs_load_b(double* tile, const double* __restrict__ src, long long ld, long long X,
                                          long long K, long long x0, long long k0, const double* __restrict__ xcol) {
    const int tid = threadIdx.x;
    if (VEC16) {                       // 16 rows x 128 chunks(16B): thread -> chunk tid&127 of rows (tid>>7) + 2 i
        const int r0 = tid >> 7, c2 = 2 * (tid & 127);
        const long long gx = x0 + c2;
        const bool xok = gx < X;
        const bool extra = (xcol != nullptr) && (gx == X);
        const double* p = src + (k0 + r0) * ld + gx;
        double* d = tile + r0 * GS_LDX + c2;
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            const bool kok = k0 + r0 + 2 * i < K;
            if (!extra) {
                const bool ok = xok && kok;
                cp_async16(d + 2 * i * GS_LDX, ok ? p + 2 * i * ld : src, ok);
            } else {
                cp_async8(d + 2 * i * GS_LDX, kok ? xcol + k0 + r0 + 2 * i : src, kok);
                cp_async8(d + 2 * i * GS_LDX + 1, src, false);
            }
        }
    } else {                           // 16 rows x 256 doubles: thread -> column tid of rows i
        const long long gx = x0 + tid;
        const bool xok = gx < X;
        const bool extra = (xcol != nullptr) && (gx == X);
        const double* p = src + k0 * ld + gx;
        double* d = tile + tid;
#pragma unroll
        for (int i = 0; i < GM_BK; ++i) {
            const bool kok = k0 + i < K;
            if (!extra) {
                const bool ok = xok && kok;
                cp_async8(d + i * GS_LDX, ok ? p + i * ld : src, ok);
            } else {
                cp_async8(d + i * GS_LDX, kok ? xcol + k0 + i : src, kok);
            }
        }
    }
}

// operator-tile generation in two phases: (1) Philox rounds -> 4 words in registers, (2) Box-Muller + store
__device__ __forceinline__ Philox4 gs_gen_words(uint64_t seed, long long col_offset, long long x0, long long k0) {
    const int tid = threadIdx.x;                      // 64 rows x 4 quads: one Philox block per thread
    const long long gx = x0 + (tid >> 2), gk = k0 + 4 * (tid & 3);
    const uint64_t q = (uint64_t)(col_offset + gk) >> 2;
    return philox4x32_10((uint32_t)q, (uint32_t)(q >> 32), (uint32_t)gx, 0u, (uint32_t)seed, (uint32_t)(seed >> 32));
}
__device__ __forceinline__ void gs_store_a(double* tile, const Philox4& words, long long X, long long K, long long x0,
                                           long long k0) {
    const int tid = threadIdx.x;
    const int r = tid >> 2, qd = tid & 3;
    const long long gx = x0 + r, gk = k0 + 4 * qd;
    double g[4] = {0.0, 0.0, 0.0, 0.0};
    if (gx < X && gk < K) {
        philox_words_to_normal4(words, g);
        if (gk + 4 > K) {
#pragma unroll
            for (int j = 0; j < 4; ++j) if (gk + j >= K) g[j] = 0.0;
        }
    }
    double* dst = tile + r * GM_LDK + 4 * qd;
    *reinterpret_cast<double2*>(dst) = make_double2(g[0], g[1]);
    *reinterpret_cast<double2*>(dst + 2) = make_double2(g[2], g[3]);
}

template <bool VEC16>
__global__ void __launch_bounds__(GM_THREADS, 1) gauss_sketch_kernel(const GemmParams p) {
    extern __shared__ __align__(16) double gm_smem[];
    double* sA = gm_smem;                               // [STAGES][GS_TILE_A]
    double* sB = gm_smem + GS_STAGES * GS_TILE_A;       // [STAGES][GS_TILE_B]
    const int tid = threadIdx.x, lane = tid & 31, wn = tid >> 5;     // 8 warps along N
    const long long m0 = (long long)blockIdx.y * GS_BM, n0 = (long long)blockIdx.x * GS_BN;
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

    const bool c_in_acc = p.part == nullptr && p.beta != 0.0 && p.alpha != 0.0;     // see gemm_f64_kernel
    if (c_in_acc) {
        const double f = p.beta / p.alpha;
        const int fr0 = lane >> 2, fc0 = lane & 3;
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            const long long gm = m0 + i * 8 + fr0;
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const long long gn = n0 + wn * 32 + j * 8 + 2 * fc0;
#pragma unroll
                for (int e = 0; e < 2; ++e)
                    if (gm < p.M && gn + e < p.N) acc[i][j][e] = f * p.C[gm * p.ldc + gn + e];
            }
        }
    }

    Philox4 words{0u, 0u, 0u, 0u};
    auto gen_words = [&](int kt_local) { words = gs_gen_words(p.seed, p.col_offset, m0, (kt_begin + kt_local) * GM_BK); };
    auto gen_store = [&](int kt_local) {
        gs_store_a(sA + (kt_local % GS_STAGES) * GS_TILE_A, words, p.M, p.K, m0, (kt_begin + kt_local) * GM_BK);
    };
    auto issue_a = [&](int kt_local) { gen_words(kt_local); gen_store(kt_local); };
    // (staggering the generation between the two warps of a scheduler, or splitting it further across
    //  k-steps with run-time phases, measured slower: 25.8 vs 27.1 TF -- the extra branches cost more)
    auto issue_b = [&](int kt_local) {
        gs_load_b<VEC16>(sB + (kt_local % GS_STAGES) * GS_TILE_B, p.B, p.ldb, p.Nb, p.K, n0,
                         (kt_begin + kt_local) * GM_BK, p.xcol);
    };
#pragma unroll
    for (int s = 0; s < GS_STAGES - 1; ++s) {
        if (s < nkt) { issue_a(s); issue_b(s); }
        cp_async_commit();
    }
    const int fr = lane >> 2, fc = lane & 3;
    for (int kt = 0; kt < nkt; ++kt) {
        cp_async_wait<GS_STAGES - 2>();
        __syncthreads();
        const double* a = sA + (kt % GS_STAGES) * GS_TILE_A;
        const double* b = sB + (kt % GS_STAGES) * GS_TILE_B;
        const bool more = kt + GS_STAGES - 1 < nkt;
#pragma unroll
        for (int kk = 0; kk < GM_BK; kk += 4) {
            double af[8], bf[4];
#pragma unroll
            for (int i = 0; i < 8; ++i) af[i] = a[(i * 8 + fr) * GM_LDK + kk + fc];
#pragma unroll
            for (int j = 0; j < 4; ++j) bf[j] = b[(kk + fc) * GS_LDX + wn * 32 + j * 8 + fr];
#pragma unroll
            for (int i = 0; i < 8; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j) dmma884(acc[i][j][0], acc[i][j][1], af[i], bf[j]);
            if (kk == 0 && more) issue_b(kt + GS_STAGES - 1);
            if (kk == 0) cp_async_commit();
            if (kk == 4 && more) gen_words(kt + GS_STAGES - 1);
            if (kk == 8 && more) gen_store(kt + GS_STAGES - 1);
        }
    }
    cp_async_wait<0>();

    const bool split = p.part != nullptr;
    double* out = split ? p.part + (size_t)blockIdx.z * p.M * p.N : p.C;
    const long long ldo = split ? p.N : p.ldc;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        const long long gm = m0 + i * 8 + fr;
        if (gm >= p.M) continue;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const long long gn = n0 + wn * 32 + j * 8 + 2 * fc;
#pragma unroll
            for (int e = 0; e < 2; ++e) {
                if (gn + e < p.N) {
                    double* dst = out + gm * ldo + gn + e;
                    if (split) *dst = acc[i][j][e];
                    else if (c_in_acc || p.beta == 0.0) *dst = p.alpha * acc[i][j][e];
                    else *dst = fma(p.alpha, acc[i][j][e], p.beta * *dst);
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
    const long long sms = num_sms();
    if (tiles >= 4 * sms || ktiles < 16) return 1;
    // Few output tiles, long K (Gram matrices, the 128 x nc products inside the QR): split K over gridDim.z.  The
    // split count minimises a small cost model in microseconds -- waves x (k-tiles per CTA x 2.4 + 6 fixed) for the
    // product (one CTA per SM; 2.4 us per 128 x 128 x 16 step at the DMMA rate) plus, when split, the reduce kernel
    // (6 us + the partial tiles written and read back at ~4 TB/s).  Filling the last wave exactly is not worth four
    // times as many partials: 37 splits of 128 x 1921 x 8192 measured 203 us where the model's 9 splits take ~155.
    long long smax = ktiles / 8;
    if (smax > 64) smax = 64;
    int best = 1;
    double best_t = 0.0;
    for (long long sp = 1; sp <= smax || sp == 1; ++sp) {
        const long long kps = (ktiles + sp - 1) / sp;
        const long long zs = (ktiles + kps - 1) / kps;           // splits actually launched
        if (zs != sp && sp > 1) continue;
        const long long waves = (tiles * zs + sms - 1) / sms;
        double t = (double)waves * ((double)kps * 2.4 + 6.0);
        if (zs > 1) t += 6.0 + (double)zs * (double)M * (double)N * 16.0 / 4.0e6;
        if (sp == 1 || t < 0.97 * best_t) { best_t = t; best = (int)zs; }     // more partials only for a real gain
    }
    return best;
}

template <int TA, int TB>
static int launch_gemm(GemmParams& p, void* ws, size_t ws_bytes, int vec16, cudaStream_t st) {
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
    // the update form (beta != 0, no split-K) of the plain product has its own instantiation
    const bool cacc = (TA == 0 && TB == 0) && zs == 1 && p.beta != 0.0 && p.alpha != 0.0;
    static const bool pf = [] { const char* e = getenv("PLA_GEMM_PREFETCH"); return !(e && e[0] == '0'); }();
    p.ctas_in_flight = pf ? num_sms() : 0;
#define PLA_GEMM_LAUNCH(V, C)                                                                                             \
    do {                                                                                                                  \
        e = cudaFuncSetAttribute(gemm_f64_kernel<TA, TB, V, C>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);  \
        if (e == cudaSuccess) { gemm_f64_kernel<TA, TB, V, C><<<grid, GM_THREADS, smem, st>>>(p); e = cudaGetLastError(); note_launch(); } \
    } while (0)
    if (cacc) {
        if constexpr (TA == 0 && TB == 0) {
            if (vec16 == 1) PLA_GEMM_LAUNCH(1, true); else if (vec16 == 2) PLA_GEMM_LAUNCH(2, true); else PLA_GEMM_LAUNCH(0, true);
        }
    } else {
        if (vec16 == 1) PLA_GEMM_LAUNCH(1, false); else if (vec16 == 2) PLA_GEMM_LAUNCH(2, false); else PLA_GEMM_LAUNCH(0, false);
    }
#undef PLA_GEMM_LAUNCH
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

// out_b[r] (+)= scale * sum_i G[r, off+i] b[i]: one warp per (operator row, k-split); partials reduced in order.
__global__ void __launch_bounds__(256) gauss_matvec_kernel(const double* __restrict__ b, long long m, long long d,
                                                           uint64_t seed, long long col_offset, long long k_per_split,
                                                           double* part) {
    const int lane = threadIdx.x & 31;
    const long long r = (long long)blockIdx.x * 8 + (threadIdx.x >> 5);
    if (r >= d) return;
    const long long kb = (long long)blockIdx.y * k_per_split;
    long long ke = kb + k_per_split;
    if (ke > m) ke = m;
    double acc = 0.0;
    for (long long k = kb + 4 * lane; k < ke; k += 128) {
        double g[4];
        philox_normal4(seed, (uint32_t)r, (uint64_t)(col_offset + k) >> 2, g);
#pragma unroll
        for (int j = 0; j < 4; ++j)
            if (k + j < ke) acc = fma(g[j], b[k + j], acc);
    }
    acc = warp_sum(acc);
    if (lane == 0) part[(size_t)blockIdx.y * d + r] = acc;
}
__global__ void __launch_bounds__(256) gauss_matvec_reduce(const double* __restrict__ part, int splits, long long d,
                                                           double scale, double beta, double* out, long long ldo) {
    const long long r = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= d) return;
    double acc = 0.0;
    for (int s = 0; s < splits; ++s) acc += part[(size_t)s * d + r];
    out[r * ldo] = (beta == 0.0) ? scale * acc : fma(scale, acc, beta * out[r * ldo]);
}

static int choose_splits_tiles(long long tiles, long long K) {
    const long long ktiles = (K + GM_BK - 1) / GM_BK;
    const long long sms = num_sms();
    if (tiles >= 4 * sms || ktiles < 64) return 1;
    long long smax = ktiles / 32;
    if (smax > 64) smax = 64;
    int best = 1;
    double best_eff = (double)tiles / (double)(((tiles + sms - 1) / sms) * sms);
    for (long long sp = 2; sp <= smax; ++sp) {
        const long long ctas = tiles * sp;
        const double eff = (double)ctas / (double)(((ctas + sms - 1) / sms) * sms);
        if (eff > best_eff + 0.02) { best_eff = eff; best = (int)sp; }
        if (ctas >= 8 * sms) break;
    }
    return best;
}

static int launch_gauss(GemmParams& p, void* ws, size_t ws_bytes, bool vec16, cudaStream_t st) {
    const long long tiles = ((p.M + GS_BM - 1) / GS_BM) * ((p.N + GS_BN - 1) / GS_BN);
    const int splits = choose_splits_tiles(tiles, p.K);
    const long long ktiles = (p.K + GM_BK - 1) / GM_BK;
    p.ktiles_per_split = (int)((ktiles + splits - 1) / splits);
    const int zs = (int)((ktiles + p.ktiles_per_split - 1) / p.ktiles_per_split);
    p.part = nullptr;
    if (zs > 1) {
        const size_t need = (size_t)zs * p.M * p.N * sizeof(double);
        if (ws == nullptr || ws_bytes < need) { set_error("sketch_gauss: workspace too small (%zu < %zu)", ws_bytes, need); return -13; }
        p.part = (double*)ws;
    }
    dim3 grid((unsigned)((p.N + GS_BN - 1) / GS_BN), (unsigned)((p.M + GS_BM - 1) / GS_BM), (unsigned)zs);
    const size_t smem = (size_t)GS_STAGES * (GS_TILE_A + GS_TILE_B) * sizeof(double);
    cudaError_t e;
    if (vec16) {
        e = cudaFuncSetAttribute(gauss_sketch_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e == cudaSuccess) { gauss_sketch_kernel<true><<<grid, GM_THREADS, smem, st>>>(p); e = cudaGetLastError(); note_launch(); }
    } else {
        e = cudaFuncSetAttribute(gauss_sketch_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e == cudaSuccess) { gauss_sketch_kernel<false><<<grid, GM_THREADS, smem, st>>>(p); e = cudaGetLastError(); note_launch(); }
    }
    if (e != cudaSuccess) { set_error("sketch_gauss: launch failed: %s", cudaGetErrorString(e)); return (int)e; }
    if (zs > 1) {
        long long total = p.M * p.N;
        int nb = (int)((total + 255) / 256);
        if (nb > 4 * num_sms()) nb = 4 * num_sms();
        gemm_splitk_reduce<<<nb, 256, 0, st>>>(p.part, zs, p.M, p.N, p.alpha, p.beta, p.C, p.ldc);
        e = cudaGetLastError();
        note_launch();
        if (e != cudaSuccess) { set_error("sketch_gauss: reduce launch failed: %s", cudaGetErrorString(e)); return (int)e; }
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
    p.ctas_in_flight = 0;
    // 16-byte copies need 16-byte aligned rows (even leading dimension); an odd extent along the contiguous axis
    // ends in a half-filled slot (cp.async src-size 8), so the d x (n + 1) sketch with its even pitch qualifies
    const bool even = ((transa ? M : K) % 2 == 0) && ((transb ? K : N) % 2 == 0);
    const int vec16 = (aligned16(A, lda) && aligned16(B, ldb)) ? (even ? 1 : 2) : 0;
    cudaStream_t st = (cudaStream_t)stream;
    if (!transa && !transb) return launch_gemm<0, 0>(p, ws, ws_bytes, vec16, st);
    if (transa && !transb) return launch_gemm<1, 0>(p, ws, ws_bytes, vec16, st);
    if (!transa && transb) return launch_gemm<0, 1>(p, ws, ws_bytes, vec16, st);
    return launch_gemm<1, 1>(p, ws, ws_bytes, vec16, st);
}

extern "C" size_t pla_sketch_gauss_workspace_bytes(int64_t d, int64_t n, int64_t m) {
    size_t best = (size_t)d * 8 * 64;                 // k-split partials of the separate rhs sketch
    for (int64_t nn = n; nn <= n + 1; ++nn) {
        const long long tiles = ((d + GS_BM - 1) / GS_BM) * ((nn + GS_BN - 1) / GS_BN);
        const int sp = choose_splits_tiles(tiles, m);
        const size_t need = sp > 1 ? (size_t)sp * d * nn * sizeof(double) : 0;
        if (need > best) best = need;
    }
    return best;
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
    cudaStream_t st = (cudaStream_t)stream;
    // The rhs rides as an extra logical column of A unless that would open a whole new 128-column
    // tile (n a multiple of 128): then S @ b is a separate generate-and-dot kernel (~1 % of the work).
    const bool separate_b = (bvec != nullptr) && (n % GS_BN == 0) && (m % 4 == 0);
    GemmParams p;
    p.A = nullptr; p.lda = 0; p.B = A; p.ldb = lda; p.C = out; p.ldc = ldo; p.M = d; p.K = m;
    p.N = separate_b ? n : ncols;
    p.Nb = n; p.xcol = separate_b ? nullptr : bvec;
    p.alpha = scale; p.beta = beta; p.seed = seed; p.col_offset = col_offset; p.part = nullptr; p.ctas_in_flight = 0;
    const bool vec16 = aligned16(A, lda) && (n % 2 == 0);
    int rc = launch_gauss(p, ws, ws_bytes, vec16, st);
    if (rc != 0 || !separate_b) return rc;
    // workspace is free again once the GEMM (and its split-K reduce) are enqueued on the same stream
    int splits = (int)((4LL * num_sms() * 8 + d - 1) / d);
    if (splits < 1) splits = 1;
    const long long max_splits = (long long)(ws_bytes / ((size_t)d * 8));
    if (splits > max_splits) splits = (int)max_splits;
    if (splits > 4096) splits = 4096;
    if (splits < 1) { set_error("pla_sketch_gauss_f64: workspace too small for the rhs sketch"); return -13; }
    long long kps = ((m + splits - 1) / splits + 127) / 128 * 128;
    splits = (int)((m + kps - 1) / kps);
    dim3 grid((unsigned)((d + 7) / 8), (unsigned)splits);
    gauss_matvec_kernel<<<grid, 256, 0, st>>>(bvec, m, d, seed, col_offset, kps, (double*)ws);
    PLA_LAUNCH_CHECK();
    gauss_matvec_reduce<<<(unsigned)((d + 255) / 256), 256, 0, st>>>((const double*)ws, splits, d, scale, beta,
                                                                      out + n, ldo);
    PLA_LAUNCH_CHECK();
    return 0;
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
