// K3a / K6: blocked Householder QR of a row-major M x N matrix (LAPACK dgeqrf conventions) and
// explicit formation of Q (dorgqr).
//
// Replaces scipy.linalg.qr(., mode='economic') on the hot path: the sketch factorisation
// (parla/drivers/least_squares.py:311), the stabiliser `orth` (parla/utils/linalg_wrappers.py:6-7),
// rangefinders.py:187 and qb.py:470.
//
// Panel (16 columns) factorisation: one cooperative kernel, rows distributed over the CTAs, ONE
// grid barrier per column (the dot products needed by column j+1 are accumulated while column j's
// reflector is applied).  Trailing matrix: W = V^T C (row-split partials, fixed-order reduction),
// C -= V (T^T W) with T the compact-WY factor.  Every reduction has a fixed order => deterministic.
#include <cooperative_groups.h>
#include <type_traits>
#include "common.cuh"
#include "../../include/parla_b200.h"

namespace cg = cooperative_groups;

namespace pla {

constexpr int QR_NB = 16;
constexpr int QR_THREADS = 256;
constexpr int QR_NP = QR_NB + 1;         // partial sums per column step: sigma + 16 dots

__device__ __forceinline__ void grid_barrier(unsigned int* counter, unsigned int target) {
    __syncthreads();
    if (threadIdx.x == 0) {
        __threadfence();
        atomicAdd(counter, 1u);
        while (*((volatile unsigned int*)counter) < target) {}
        __threadfence();
    }
    __syncthreads();
}

struct PanelParams {
    double* A; long long lda;
    long long M;            // total rows
    long long j0;           // first column (and first row) of the panel
    int jb;                 // panel width (<= 16)
    double* tau;            // tau + j0
    double* gpart;          // [2][grid][QR_NP]
    double* drowbuf;        // [2][QR_NB]: snapshot of the diagonal row of the current column step
    unsigned int* counter;  // zeroed before launch
    long long rows_per_cta;
};

// Factor A[j0:M, j0:j0+jb] in place.
__global__ void __launch_bounds__(QR_THREADS) qr_panel_kernel(const PanelParams p) {
    __shared__ double red[QR_NP];
    __shared__ double wsum[QR_THREADS / 32][QR_NP];
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    const int G = gridDim.x, jb = p.jb;
    const long long r_begin = p.j0 + (long long)blockIdx.x * p.rows_per_cta;
    long long r_end = r_begin + p.rows_per_cta;
    if (r_end > p.M) r_end = p.M;
    double* Ap = p.A + p.j0;                     // column offset of the panel

    double part[QR_NP];
    // ---- dots for column 0
#pragma unroll
    for (int c = 0; c < QR_NP; ++c) part[c] = 0.0;
    for (long long i = r_begin + tid; i < r_end; i += QR_THREADS) {
        if (i == p.j0) {
            // The diagonal row of step j is rewritten by its owner DURING step j, so every other
            // thread must read a snapshot taken before the barrier, never the matrix itself.
            const double* row = Ap + i * p.lda;
#pragma unroll
            for (int c = 0; c < QR_NB; ++c)
                if (c < jb) p.drowbuf[c] = row[c];
        }
        if (i > p.j0) {
            const double* row = Ap + i * p.lda;
            const double x = row[0];
            part[0] = fma(x, x, part[0]);
#pragma unroll
            for (int c = 1; c < QR_NB; ++c)
                if (c < jb) part[1 + c] = fma(x, row[c], part[1 + c]);
        }
    }

    for (int j = 0; j < jb; ++j) {
        // ---- publish this CTA's partial sums for column j, then meet at the barrier
        {
            double v32[32];
#pragma unroll
            for (int c = 0; c < QR_NP; ++c) v32[c] = part[c];
            warp_multi_sum<QR_NP>(v32);
            if (lane < QR_NP) wsum[wid][lane] = v32[0];
        }
        __syncthreads();
        double* mypart = p.gpart + ((size_t)(j & 1) * G + blockIdx.x) * QR_NP;
        if (tid < QR_NP) {
            double acc = 0.0;
#pragma unroll
            for (int q = 0; q < QR_THREADS / 32; ++q) acc += wsum[q][tid];
            mypart[tid] = acc;
        }
        grid_barrier(p.counter, (unsigned int)(j + 1) * G);
        {   // fixed-order sum of the G per-CTA partials: lane <-> CTA, warp <-> column (loads in parallel)
            const double* gp = p.gpart + (size_t)(j & 1) * G * QR_NP;
            for (int c = wid; c < QR_NP; c += QR_THREADS / 32) {
                double acc = 0.0;
                for (int b = lane; b < G; b += 32) acc += __ldcg(gp + (size_t)b * QR_NP + c);
                acc = warp_sum(acc);
                if (lane == 0) red[c] = acc;
            }
        }
        __syncthreads();

        // ---- reflector j (LAPACK dlarfg): beta = -sign(alpha)|x|, tau = (beta-alpha)/beta, v = x/(alpha-beta)
        const long long dj = p.j0 + j;
        const double* drow = p.drowbuf + (j & 1) * QR_NB;
        const double alpha = __ldcg(drow + j);
        const double sigma = red[0];
        double beta = alpha, tau = 0.0, scale = 0.0;
        if (sigma != 0.0) {
            const double nrm = sqrt(fma(alpha, alpha, sigma));
            beta = alpha >= 0.0 ? -nrm : nrm;
            tau = (beta - alpha) / beta;
            scale = 1.0 / (alpha - beta);
        }
        double wv[QR_NB];      // tau * (v^T A[:, c]) for c > j
#pragma unroll
        for (int c = 0; c < QR_NB; ++c) {
            wv[c] = 0.0;
            if (c > j && c < jb) wv[c] = tau * fma(scale, red[1 + c], __ldcg(drow + c));
        }
        if (blockIdx.x == 0 && tid == 0) p.tau[j] = tau;

        // ---- apply to my rows; gather the dots of column j+1 on the fly
#pragma unroll
        for (int c = 0; c < QR_NP; ++c) part[c] = 0.0;
        for (long long i = r_begin + tid; i < r_end; i += QR_THREADS) {
            if (i < dj) continue;
            double* row = Ap + i * p.lda;
            if (i == dj) {
                row[j] = beta;
#pragma unroll
                for (int c = 0; c < QR_NB; ++c)
                    if (c > j && c < jb) row[c] -= wv[c];
            } else {
                const double v = row[j] * scale;
                row[j] = v;
                double a[QR_NB];
#pragma unroll
                for (int c = 0; c < QR_NB; ++c) {
                    a[c] = 0.0;
                    if (c > j && c < jb) { a[c] = fma(-v, wv[c], row[c]); row[c] = a[c]; }
                }
                if (i == dj + 1 && j + 1 < jb) {          // next step's diagonal row: publish its snapshot
                    double* nxt = p.drowbuf + ((j + 1) & 1) * QR_NB;
#pragma unroll
                    for (int c = 0; c < QR_NB; ++c)
                        if (c > j && c < jb) nxt[c] = a[c];
                }
                if (i > dj + 1 && j + 1 < jb) {
                    double x = 0.0;
#pragma unroll
                    for (int c = 0; c < QR_NB; ++c) if (c == j + 1) x = a[c];
                    part[0] = fma(x, x, part[0]);
#pragma unroll
                    for (int c = 0; c < QR_NB; ++c)
                        if (c > j + 1 && c < jb) part[1 + c] = fma(x, a[c], part[1 + c]);
                }
            }
        }
    }
}

// ---- Panel factorisation inside ONE thread-block cluster (rows <= cluster_size * QRC_MAXROWS) -----
// The 16-column panel lives in the shared memory of the cluster's CTAs for the whole factorisation;
// the per-column partial sums are exchanged through distributed shared memory and ordered by the
// hardware cluster barrier, so a column step costs ~1 us instead of several L2 round trips.
constexpr int QRC_THREADS = 512;
constexpr int QRC_RPT = 3;                               // rows per thread
constexpr int QRC_MAXROWS = QRC_THREADS * QRC_RPT;       // rows per CTA
constexpr int QRC_MAXCS = 16;
constexpr int QRC_XW = QR_NP + QR_NB;                    // exchange record: 17 partial sums + diagonal-row snapshot

__global__ void __launch_bounds__(QRC_THREADS, 1) qr_panel_cluster_kernel(const PanelParams p) {
    cg::cluster_group cluster = cg::this_cluster();
    const int rank = (int)cluster.block_rank(), CS = (int)cluster.num_blocks();
    extern __shared__ double qc_smem[];
    double* P = qc_smem;                                           // [QR_NB][QRC_MAXROWS] column-major panel slice
    double* xch = P + QR_NB * QRC_MAXROWS;                         // [2][QRC_MAXCS][QRC_XW]
    double* wsum = xch + 2 * QRC_MAXCS * QRC_XW;                   // [16 warps][QR_NP]
    double* red = wsum + (QRC_THREADS / 32) * QR_NP;               // [QR_NP] cluster-wide totals of the step
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    const int jb = p.jb;
    const long long r_begin = p.j0 + (long long)rank * p.rows_per_cta;
    long long r_end = r_begin + p.rows_per_cta;
    if (r_end > p.M) r_end = p.M;
    const int nloc = (int)max(0LL, r_end - r_begin);
    double* Ap = p.A + p.j0;

    // load my rows (coalesced over the 16 columns of a row: consecutive threads take consecutive elements)
    for (int idx = tid; idx < nloc * QR_NB; idx += QRC_THREADS) {
        const int r = idx / QR_NB, c = idx % QR_NB;
        P[c * QRC_MAXROWS + r] = (c < jb) ? Ap[(r_begin + r) * p.lda + c] : 0.0;
    }
    __syncthreads();

    double part[QR_NP];
#pragma unroll
    for (int c = 0; c < QR_NP; ++c) part[c] = 0.0;
    for (int r = tid; r < nloc; r += QRC_THREADS) {
        if (r_begin + r > p.j0) {
            const double x = P[r];
            part[0] = fma(x, x, part[0]);
#pragma unroll
            for (int c = 1; c < QR_NB; ++c) part[1 + c] = fma(x, P[c * QRC_MAXROWS + r], part[1 + c]);
        }
    }

    for (int j = 0; j < jb; ++j) {
        const long long dj = p.j0 + j;
        // ---- CTA-level sums
        {
            double v32[32];
#pragma unroll
            for (int c = 0; c < QR_NP; ++c) v32[c] = part[c];
            warp_multi_sum<QR_NP>(v32);
            if (lane < QR_NP) wsum[wid * QR_NP + lane] = v32[0];
        }
        __syncthreads();
        // ---- publish to every CTA of the cluster (DSMEM): thread (dst, c) writes one double
        double* rec = xch + ((size_t)(j & 1) * QRC_MAXCS + rank) * QRC_XW;
        for (int idx = tid; idx < CS * QRC_XW; idx += QRC_THREADS) {
            const int dst = idx / QRC_XW, c = idx % QRC_XW;
            double v;
            if (c < QR_NP) {
                v = 0.0;
#pragma unroll
                for (int q = 0; q < QRC_THREADS / 32; ++q) v += wsum[q * QR_NP + c];
            } else {
                // snapshot of the diagonal row (owned by rank 0, local row j), taken before anyone rewrites it
                v = (rank == 0) ? P[(c - QR_NP) * QRC_MAXROWS + j] : 0.0;
            }
            double* remote = cluster.map_shared_rank(rec, dst);
            remote[c] = v;
        }
        cluster.sync();
        // ---- every CTA reduces the CS records in rank order (identical results everywhere); one warp
        //      does the sums (lane <-> quantity), the rest of the CTA reads the 17 totals
        const double* recs = xch + (size_t)(j & 1) * QRC_MAXCS * QRC_XW;
        if (wid == 0 && lane < QR_NP) {
            double acc = 0.0;
            for (int b = 0; b < CS; ++b) acc += recs[b * QRC_XW + lane];
            red[lane] = acc;
        }
        __syncthreads();
        const double sigma = red[0];
        const double alpha = recs[QR_NP + j];                       // rank 0's record holds the diagonal row
        double beta = alpha, tau = 0.0, scale = 0.0;
        if (sigma != 0.0) {
            const double nrm = sqrt(fma(alpha, alpha, sigma));
            beta = alpha >= 0.0 ? -nrm : nrm;
            tau = (beta - alpha) / beta;
            scale = 1.0 / (alpha - beta);
        }
        double wv[QR_NB];
#pragma unroll
        for (int c = 0; c < QR_NB; ++c) {
            wv[c] = 0.0;
            if (c > j && c < jb) wv[c] = tau * fma(scale, red[1 + c], recs[QR_NP + c]);
        }
        if (rank == 0 && tid == 0) p.tau[j] = tau;
        // ---- apply to my rows (shared memory), gathering the dots of column j + 1
#pragma unroll
        for (int c = 0; c < QR_NP; ++c) part[c] = 0.0;
        for (int r = tid; r < nloc; r += QRC_THREADS) {
            const long long i = r_begin + r;
            if (i < dj) continue;
            if (i == dj) {
                P[j * QRC_MAXROWS + r] = beta;
#pragma unroll
                for (int c = 0; c < QR_NB; ++c)
                    if (c > j && c < jb) P[c * QRC_MAXROWS + r] -= wv[c];
            } else {
                const double v = P[j * QRC_MAXROWS + r] * scale;
                P[j * QRC_MAXROWS + r] = v;
                double a[QR_NB];
#pragma unroll
                for (int c = 0; c < QR_NB; ++c) {
                    a[c] = 0.0;
                    if (c > j && c < jb) { a[c] = fma(-v, wv[c], P[c * QRC_MAXROWS + r]); P[c * QRC_MAXROWS + r] = a[c]; }
                }
                if (i > dj + 1 && j + 1 < jb) {
                    double x = 0.0;
#pragma unroll
                    for (int c = 0; c < QR_NB; ++c) if (c == j + 1) x = a[c];
                    part[0] = fma(x, x, part[0]);
#pragma unroll
                    for (int c = 0; c < QR_NB; ++c)
                        if (c > j + 1 && c < jb) part[1 + c] = fma(x, a[c], part[1 + c]);
                }
            }
        }
        __syncthreads();          // row dj+1 (next diagonal) is final before rank 0 snapshots it
    }
    // ---- write the factored panel back
    for (int idx = tid; idx < nloc * QR_NB; idx += QRC_THREADS) {
        const int r = idx / QR_NB, c = idx % QR_NB;
        if (c < jb) Ap[(r_begin + r) * p.lda + c] = P[c * QRC_MAXROWS + r];
    }
    cluster.sync();               // no CTA may exit while others can still write into its shared memory
}

// ---- Gram of the panel's reflectors: Gpart[cta][16][16] = sum_rows V[i][a] V[i][b]
struct VView {
    const double* A; long long lda; long long j0; int jb;
    // V[i][a] for global row i (>= j0): implicit unit diagonal / zeros above it
    __device__ __forceinline__ double at(long long i, int a) const {
        if (a >= jb) return 0.0;
        const long long d = j0 + a;
        if (i > d) return A[i * lda + d];
        return i == d ? 1.0 : 0.0;
    }
};

constexpr int QR_TR = 64;   // rows per staged V tile

__device__ __forceinline__ void stage_v(double (*vt)[QR_NB], const VView& V, long long row0, long long M) {
    for (int idx = threadIdx.x; idx < QR_TR * QR_NB; idx += blockDim.x) {
        const int r = idx / QR_NB, a = idx % QR_NB;
        const long long i = row0 + r;
        vt[r][a] = i < M ? V.at(i, a) : 0.0;
    }
}

__global__ void __launch_bounds__(QR_THREADS) qr_gram_kernel(VView V, long long M, long long rows_per_cta,
                                                             double* gpart) {
    __shared__ __align__(16) double vt[QR_TR][QR_NB];
    const int a = threadIdx.x / QR_NB, b = threadIdx.x % QR_NB;
    const long long r_begin = V.j0 + (long long)blockIdx.x * rows_per_cta;
    long long r_end = r_begin + rows_per_cta;
    if (r_end > M) r_end = M;
    double acc = 0.0;
    for (long long row0 = r_begin; row0 < r_end; row0 += QR_TR) {
        __syncthreads();
        stage_v(vt, V, row0, r_end);
        __syncthreads();
#pragma unroll 8
        for (int r = 0; r < QR_TR; ++r) acc = fma(vt[r][a], vt[r][b], acc);
    }
    gpart[(size_t)blockIdx.x * QR_NB * QR_NB + threadIdx.x] = acc;
}

// T (upper triangular, jb x jb, stored 16 x 16 row-major): T[j][j] = tau_j,
// T[0:j, j] = -tau_j * T[0:j, 0:j] * (V[:, 0:j]^T v_j)      (LAPACK dlarft, forward / columnwise)
__global__ void __launch_bounds__(QR_THREADS) qr_build_t_kernel(const double* gpart, int nparts, const double* tau,
                                                                int jb, double* T) {
    __shared__ double G[QR_NB][QR_NB];
    __shared__ double Ts[QR_NB][QR_NB];
    double acc = 0.0;
    for (int b = 0; b < nparts; ++b) acc += gpart[(size_t)b * QR_NB * QR_NB + threadIdx.x];
    G[threadIdx.x / QR_NB][threadIdx.x % QR_NB] = acc;
    Ts[threadIdx.x / QR_NB][threadIdx.x % QR_NB] = 0.0;
    __syncthreads();
    // column j of T needs columns < j: sequential over j, rows of a column in parallel (thread r <-> row r)
    for (int j = 0; j < jb; ++j) {
        const double tj = tau[j];
        const int r = threadIdx.x;
        if (r < j) {
            double sacc = 0.0;
            for (int l = r; l < j; ++l) sacc = fma(Ts[r][l], G[l][j], sacc);
            Ts[r][j] = -tj * sacc;
        } else if (r == j) {
            Ts[j][j] = tj;
        }
        __syncthreads();
    }
    T[threadIdx.x] = Ts[threadIdx.x / QR_NB][threadIdx.x % QR_NB];
}

// ---- trailing update, step 1: Wpart[split][k][c] = sum_{rows in split} V[i][k] * C[i][c]
struct TrailParams {
    VView V; long long M;
    double* C; long long ldc; long long nc;          // C = columns to update (pointer already offset)
    long long rows_per_split; int splits;
    double* wpart;                                    // [splits][16][nc]
    double* w2;                                       // [16][nc] = op(T) * sum_splits wpart
    const double* T; int t_transpose;                 // apply T^T (QR) or T (forming Q)
};

__global__ void __launch_bounds__(QR_THREADS) qr_trail_w_kernel(const TrailParams p) {
    __shared__ __align__(16) double vt[QR_TR][QR_NB];
    const long long c = (long long)blockIdx.x * QR_THREADS + threadIdx.x;
    const long long r_begin = p.V.j0 + (long long)blockIdx.y * p.rows_per_split;
    long long r_end = r_begin + p.rows_per_split;
    if (r_end > p.M) r_end = p.M;
    double acc[QR_NB];
#pragma unroll
    for (int k = 0; k < QR_NB; ++k) acc[k] = 0.0;
    for (long long row0 = r_begin; row0 < r_end; row0 += QR_TR) {
        __syncthreads();
        stage_v(vt, p.V, row0, r_end);
        __syncthreads();
        if (c < p.nc) {
            const int lim = (int)min((long long)QR_TR, r_end - row0);
            for (int r = 0; r < lim; r += 8) {
                double x[8];
#pragma unroll
                for (int q = 0; q < 8; ++q) x[q] = (r + q < lim) ? p.C[(row0 + r + q) * p.ldc + c] : 0.0;   // 8 loads in flight
#pragma unroll
                for (int q = 0; q < 8; ++q) {
                    const double2* v2 = reinterpret_cast<const double2*>(vt[r + q]);
#pragma unroll
                    for (int k = 0; k < QR_NB / 2; ++k) {
                        const double2 vv = v2[k];
                        acc[2 * k] = fma(vv.x, x[q], acc[2 * k]);
                        acc[2 * k + 1] = fma(vv.y, x[q], acc[2 * k + 1]);
                    }
                }
            }
        }
    }
    if (c < p.nc) {
#pragma unroll
        for (int k = 0; k < QR_NB; ++k) p.wpart[((size_t)blockIdx.y * QR_NB + k) * p.nc + c] = acc[k];
    }
}

// ---- step 2a: W2 = op(T) * (sum over row splits of Wpart), once per column (fixed summation order)
// block = 32 columns x 8 split-groups; partial sums over split-groups are combined in group order.
__global__ void __launch_bounds__(QR_THREADS) qr_trail_reduce_kernel(const TrailParams p) {
    __shared__ double acc_s[8][QR_NB][33];
    __shared__ double Ts[QR_NB][QR_NB];
    const int cl = threadIdx.x & 31, grp = threadIdx.x >> 5;
    const long long c = (long long)blockIdx.x * 32 + cl;
    Ts[threadIdx.x / QR_NB][threadIdx.x % QR_NB] = p.T[threadIdx.x];
    double acc[QR_NB];
#pragma unroll
    for (int k = 0; k < QR_NB; ++k) acc[k] = 0.0;
    if (c < p.nc) {
        // four independent batches of loads in flight per thread (fixed association order)
        for (int sp = grp; sp < p.splits; sp += 32) {
            double t4[4][QR_NB];
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                const int s2 = sp + 8 * u;
                const double* src = p.wpart + (size_t)(s2 < p.splits ? s2 : 0) * QR_NB * p.nc + c;
#pragma unroll
                for (int k = 0; k < QR_NB; ++k) t4[u][k] = (s2 < p.splits) ? src[(size_t)k * p.nc] : 0.0;
            }
#pragma unroll
            for (int u = 0; u < 4; ++u)
#pragma unroll
                for (int k = 0; k < QR_NB; ++k) acc[k] += t4[u][k];
        }
    }
#pragma unroll
    for (int k = 0; k < QR_NB; ++k) acc_s[grp][k][cl] = acc[k];
    __syncthreads();
    // thread (k = grp*2 + h, column cl): w[l] = sum_groups acc_s[g][l][cl];  W2[k] = sum_l op(T)[k][l] w[l]
    if (c < p.nc) {
#pragma unroll
        for (int h = 0; h < 2; ++h) {
            const int k = grp * 2 + h;
            double sacc = 0.0;
            for (int l = 0; l < QR_NB; ++l) {
                double wl = 0.0;
#pragma unroll
                for (int g = 0; g < 8; ++g) wl += acc_s[g][l][cl];
                sacc = fma(p.t_transpose ? Ts[l][k] : Ts[k][l], wl, sacc);
            }
            p.w2[(size_t)k * p.nc + c] = sacc;
        }
    }
}

// ---- step 2b: C -= V W2
__global__ void __launch_bounds__(QR_THREADS) qr_trail_apply_kernel(const TrailParams p) {
    __shared__ __align__(16) double vt[QR_TR][QR_NB];
    const long long c = (long long)blockIdx.x * QR_THREADS + threadIdx.x;
    const long long r_begin = p.V.j0 + (long long)blockIdx.y * p.rows_per_split;
    long long r_end = r_begin + p.rows_per_split;
    if (r_end > p.M) r_end = p.M;
    double w2[QR_NB];
#pragma unroll
    for (int k = 0; k < QR_NB; ++k) w2[k] = (c < p.nc) ? p.w2[(size_t)k * p.nc + c] : 0.0;
    for (long long row0 = r_begin; row0 < r_end; row0 += QR_TR) {
        __syncthreads();
        stage_v(vt, p.V, row0, r_end);
        __syncthreads();
        if (c < p.nc) {
            const int lim = (int)min((long long)QR_TR, r_end - row0);
            for (int r = 0; r < lim; r += 8) {
                double x[8];
#pragma unroll
                for (int q = 0; q < 8; ++q) x[q] = (r + q < lim) ? p.C[(row0 + r + q) * p.ldc + c] : 0.0;
#pragma unroll
                for (int q = 0; q < 8; ++q) {
                    const double2* v2 = reinterpret_cast<const double2*>(vt[r + q]);
                    double s = 0.0;
#pragma unroll
                    for (int k = 0; k < QR_NB / 2; ++k) {
                        const double2 vv = v2[k];
                        s = fma(vv.x, w2[2 * k], s);
                        s = fma(vv.y, w2[2 * k + 1], s);
                    }
                    if (r + q < lim) p.C[(row0 + r + q) * p.ldc + c] = x[q] - s;
                }
            }
        }
    }
}

__global__ void __launch_bounds__(256) qr_eye_kernel(double* Q, long long M, long long K, long long ldq) {
    const long long total = M * K;
    for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < total;
         idx += (long long)gridDim.x * blockDim.x) {
        const long long r = idx / K, c = idx - r * K;
        Q[r * ldq + c] = (r == c) ? 1.0 : 0.0;
    }
}

// ---- outer (BLAS-3) level: 128-column block reflector applied with the DMMA GEMM -------------
constexpr int QR_NBO = 128;

// Vx[i - J0][a] = V[i][a] for the JB reflectors starting at column/row J0 (explicit unit diagonal and zeros)
__global__ void __launch_bounds__(256) qr_form_v_kernel(const double* __restrict__ A, long long lda, long long M,
                                                        long long J0, int JB, double* Vx) {
    const long long rows = M - J0, total = rows * QR_NBO;
    for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < total;
         idx += (long long)gridDim.x * blockDim.x) {
        const long long r = idx / QR_NBO;
        const int a = (int)(idx - r * QR_NBO);
        double v = 0.0;
        if (a < JB) {
            if (r > a) v = A[(J0 + r) * lda + J0 + a];
            else if (r == a) v = 1.0;
        }
        Vx[idx] = v;
    }
}

// ================================================================================================
// Cooperative block factorisation: ONE launch factors a whole 128-column block.
//
// The block's rows are spread over G co-resident CTAs (cooperative launch, one per SM) and stay in shared
// memory for the whole factorisation.  Cross-CTA sums travel as NCCL-"LL"-style 16-byte lines
// {lo, flag, hi, flag} in global memory (8-byte-atomic halves carrying their own epoch flag), so an
// all-to-all exchange costs one L2 write + one L2 read and needs no barrier, fence or atomic:
//   * per column: every CTA publishes <= 16 partial sums (|x|^2 of the column and its dots with the rest of
//     the 16-column sub-panel), polls the G x 16 lines of the others and reduces them in a fixed order --
//     every CTA computes bit-identical beta / tau / scale;
//   * per 16-column sub-panel: W = V^T C for the rest of the block and the sub-panel's Gram matrix V^T V are
//     reduce-scattered (plain partials + flag barrier, each CTA sums one slice over all CTAs in a fixed
//     order) and all-gathered as LL lines; W2 = T^T W comes from the forward substitution
//     w2_k = tau_k (w_k - sum_{l<k} G_lk w2_l)   (T^{-1} = striu(V^T V) + diag(1/tau); no division, tau = 0 ok),
//     so T is never formed; C -= V W2 is local.
// The kernel also emits the explicit reflector block Vx for the tensor-core update of the columns to the
// right of the block.  Same arithmetic (LAPACK dlarfg/dlarfb conventions) as the panel kernels above.
constexpr int QB_THREADS = 256;
constexpr int QB_LDP = QR_NBO + 1;           // 129 doubles: conflict-free column walks
constexpr int QB_MAXG = 160;                 // >= number of SMs
constexpr int QB_RPC_MAX = 160;              // rows per CTA (shared-memory limit)
constexpr int QB_WLD = QR_NBO - QR_NB;       // 112: leading dimension of the W totals
constexpr int QB_NV = QR_NB * QB_WLD + QR_NB * QR_NB;   // 2048 values per reduce-scatter
constexpr int QB_NL = 12;                    // fan-in records per thread (>= QB_MAXG / 16, multiple of 4)
constexpr int QP_NV = 2 * QR_NB;             // pair steps: 32 values per exchange (dots of TWO columns with the sub-panel)
constexpr int QP_NL = 20;                    // pair steps: fan-in records per warp (>= QB_MAXG / 8, multiple of 4)
constexpr double QP_THETA = 0.05;            // look-ahead accepted while |x'_{j+1}|^2 >= theta |x_{j+1}|^2 (see the kernel)

struct __align__(16) LLLine { uint32_t lo, f1, hi, f2; };

__device__ __forceinline__ void ll_store(LLLine* p, double v, uint32_t flag, bool pred = true) {
    const uint32_t lo = (uint32_t)__double2loint(v), hi = (uint32_t)__double2hiint(v);
    // predicated in PTX (never a branch): the column step must stay free of intra-warp divergence
    asm volatile("{\n\t.reg .pred pq;\n\tsetp.ne.u32 pq, %5, 0;\n\t"
                 "@pq st.relaxed.gpu.global.v4.u32 [%0], {%1, %2, %3, %4};\n\t}"
                 ::"l"(p), "r"(lo), "r"(flag), "r"(hi), "r"(flag), "r"((uint32_t)pred) : "memory");
}
__device__ __forceinline__ bool ll_try_load(const LLLine* p, uint32_t flag, double& v) {
    uint32_t lo, f1, hi, f2;
    asm volatile("ld.relaxed.gpu.global.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(lo), "=r"(f1), "=r"(hi), "=r"(f2) : "l"(p)
                 : "memory");
    v = __hiloint2double((int)hi, (int)lo);
    return f1 == flag && f2 == flag;
}
__device__ __forceinline__ double ll_load(const LLLine* p, uint32_t flag) {
    double v;
    while (!ll_try_load(p, flag, v)) {}
    return v;
}
// Four LL lines at p + i * stride_bytes (i = 0..3), line i loaded iff bit i of `mask`: ONE asm block, so the four
// 16-byte loads issue back to back and their L2 round trips overlap (separate asm statements interleaved with the
// flag tests were serialised: ~750 cycles per line and poll).  w0 = {lo, flag}, w1 = {hi, flag} of each line.
template <int STRIDE>
__device__ __forceinline__ void ll_load4(const void* p, unsigned mask, unsigned long long (&w0)[4],
                                         unsigned long long (&w1)[4]) {
    asm volatile(
        "{\n\t.reg .pred p0, p1, p2, p3;\n\t.reg .b32 t;\n\t"
        "and.b32 t, %9, 1;\n\tsetp.ne.u32 p0, t, 0;\n\t"
        "and.b32 t, %9, 2;\n\tsetp.ne.u32 p1, t, 0;\n\t"
        "and.b32 t, %9, 4;\n\tsetp.ne.u32 p2, t, 0;\n\t"
        "and.b32 t, %9, 8;\n\tsetp.ne.u32 p3, t, 0;\n\t"
        "@p0 ld.relaxed.gpu.global.v2.u64 {%0, %1}, [%8];\n\t"
        "@p1 ld.relaxed.gpu.global.v2.u64 {%2, %3}, [%8 + %10];\n\t"
        "@p2 ld.relaxed.gpu.global.v2.u64 {%4, %5}, [%8 + %11];\n\t"
        "@p3 ld.relaxed.gpu.global.v2.u64 {%6, %7}, [%8 + %12];\n\t}"
        : "+l"(w0[0]), "+l"(w1[0]), "+l"(w0[1]), "+l"(w1[1]), "+l"(w0[2]), "+l"(w1[2]), "+l"(w0[3]), "+l"(w1[3])
        : "l"(p), "r"(mask), "n"(STRIDE), "n"(2 * STRIDE), "n"(3 * STRIDE)
        : "memory");
}
// line received (both halves carry `flag`)?  -> value
__device__ __forceinline__ bool ll_ready(unsigned long long w0, unsigned long long w1, uint32_t flag, double& v) {
    v = __hiloint2double((int)(uint32_t)w1, (int)(uint32_t)w0);
    return (uint32_t)(w0 >> 32) == flag && (uint32_t)(w1 >> 32) == flag;
}
__device__ __forceinline__ void st_shared_pred(double* p, double v, bool pred) {
    asm volatile("{\n\t.reg .pred pq;\n\tsetp.ne.u32 pq, %2, 0;\n\t@pq st.shared.f64 [%0], %1;\n\t}"
                 ::"r"(smem_u32(p)), "d"(v), "r"((uint32_t)pred) : "memory");
}
__device__ __forceinline__ void st_global_pred(double* p, double v, bool pred) {
    asm volatile("{\n\t.reg .pred pq;\n\tsetp.ne.u32 pq, %2, 0;\n\t@pq st.global.f64 [%0], %1;\n\t}"
                 ::"l"(p), "d"(v), "r"((uint32_t)pred) : "memory");
}

struct BlockParams {
    double* A; long long lda; long long M;
    long long J0; int JB;          // block = columns [C0, C0 + JB), rows [J0, M)  (C0 == J0 inside pla_geqrf_f64)
    long long C0;
    double* tau;                   // the block's JB scalar factors
    double* Vx;                    // out: (M - J0) x 128 explicit reflectors (unit diagonal, zeros above / right of JB)
    int rpc;                       // rows per CTA (multiple of 16)
    // (sizes for the pair steps, 32 values per exchange; the one-column steps use the first half of each record)
    LLLine* ll_step;               // [2][32][QB_MAXG] partials, ONE line per 32-byte sector (stride 2 lines)
    LLLine* ll_diag;               // [2][32] diagonal-row snapshots (owner -> everybody when G <= 16, -> reducers otherwise)
    LLLine* ll_tot;                // [2][QB_MAXG][32] per-reader copies of the diagonal rows (G > 16)
    LLLine* ll_fan;                // [2][QB_MAXG readers][QB_MAXG writers][32] pushed partials (G > 16)
    int pair;                      // 1: two columns per exchange (look-ahead through Gram identities), 0: one
    double* gram_part;             // out (or null): [G][36][256] per-CTA partials of striu(Vx^T Vx), see qr_gram_reduce_kernel
    LLLine* ll_bar;                // [2][QB_MAXG]
    LLLine* ll_res;                // [2][QB_NV]
    double* wpart;                 // [G][QB_NV]
    uint32_t epoch_base;           // unique per launch within one factorisation, > 0
    long long* prof;               // debug (PLA_QR_PROF=1): [G][QB_NPROF] cycles per phase of thread 0, or null
};
constexpr int QB_NPROF = 12;
#define QB_T(i) do { if (p.prof != nullptr && threadIdx.x == 0) { const long long now_ = clock64(); \
                     pacc[i] += now_ - plast; plast = now_; } } while (0)

__global__ void __launch_bounds__(QB_THREADS, 1) qr_block_coop_kernel(const BlockParams p) {
    extern __shared__ __align__(16) double qb_smem[];
    double* P = qb_smem;                                   // [rpc][QB_LDP]  block slice, row-major
    double* Vs = P + (size_t)p.rpc * QB_LDP;               // [rpc][16]      reflectors of the current sub-panel
    double* Ws = Vs + (size_t)p.rpc * QR_NB;               // [16][QB_WLD]   W totals, then W2 in place
    double* Gs = Ws + QR_NB * QB_WLD;                      // [16][16]       Gram of the sub-panel's reflectors
    double* red = Gs + QR_NB * QR_NB;                      // [16]           totals of the column step
    double* drow = red + QR_NB;                            // [16]           snapshot of the diagonal row
    double* wsum = drow + QR_NB;                           // [8][16]
    double* taus = wsum + (QB_THREADS / 32) * QR_NB;       // [16]
    double* tot16 = taus + QR_NB;                          // [32] this CTA's 16 partial sums + its diagonal-row snapshot
    double* pr = tot16 + 2 * QR_NB;                        // pair steps: [32] totals S0_q | S1_q
    double* pd = pr + QP_NV;                               //             [32] diagonal rows D0_q | D1_q
    double* pt = pd + QP_NV;                               //             [64] this CTA's totals | its diagonal-row snapshots
    double* pw = pt + 2 * QP_NV;                           //             [8][32] per-warp partials
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    const int G = gridDim.x, b = blockIdx.x, JB = p.JB, rpc = p.rpc;
    const long long rows = p.M - p.J0;
    const long long l0 = (long long)b * rpc;               // block-local index of my first row
    const int nloc = (int)max(0LL, min((long long)rpc, rows - l0));
    const int nit = rpc / QR_NB;                           // row sweeps of the 16 x 16 thread grid
    double* Ablk = p.A + p.J0 * p.lda + p.C0;
    long long pacc[QB_NPROF], plast = clock64();
#pragma unroll
    for (int i = 0; i < QB_NPROF; ++i) pacc[i] = 0;

    for (int idx = tid; idx < nloc * QR_NBO; idx += QB_THREADS) {
        const int r = idx >> 7, c = idx & (QR_NBO - 1);
        P[r * QB_LDP + c] = (c < JB) ? Ablk[(l0 + r) * p.lda + c] : 0.0;
    }
    __syncthreads();
    QB_T(0);

    const int c = tid & 15, g = tid >> 4;                  // column of the sub-panel / row group
    uint32_t xs = 0;                                       // pair steps: exchanges so far in this block (<= JB)
    for (int c0 = 0; c0 < JB; c0 += QR_NB) {
        const int jbp = min(QR_NB, JB - c0);
        if (p.pair) {
        // ================= pair steps: ONE exchange serves TWO columns =================================================
        // Before step j every CTA contributes, over its rows BELOW row gc + 1,
        //     S0_q = x_j . x_q (q >= j)   and   S1_q = x_{j+1} . x_q (q >= j + 1)
        // and the owners of rows gc and gc + 1 publish those rows (D0_q, D1_q).  Reflector j follows exactly as in the
        // one-column step (sigma_j = S0_j + D1_j^2, x_j . x_q = S0_q + D1_j D1_q).  Because H_j is the rank-one update
        // x_q <- x_q - v wv0_q with v = scale_j x_j below the diagonal, the inner products that reflector j + 1 needs
        // after H_j follow from the SAME sums:
        //     T_q = x'_{j+1} . x'_q = S1_q - sc (wv0_q S0_{j+1} + wv0_{j+1} S0_q) + wv0_{j+1} wv0_q sc^2 S0_j,
        // and row gc + 1 after H_j is D1_q - (sc D1_j) wv0_q.  So both reflectors are known after one exchange and are
        // applied in one sweep over the rows (x_q -= a wv0_q + a' wv1_q), which also gathers the sums of the next step.
        // T_{j+1} = |x'_{j+1}|^2 is a difference: if it keeps less than QP_THETA of |x_{j+1}|^2 (column j + 1 almost
        // parallel to column j) the look-ahead is dropped for this step -- every CTA sees the same numbers and takes
        // the same decision -- and column j + 1 gets its own exchange with exactly summed norms, as in LAPACK.
        double part0 = 0.0, part1 = 0.0;
        for (int it = 0; it < nit; ++it) {
            const int r = g + QR_NB * it;
            if (r < nloc && l0 + r > c0 + 1 && c < jbp) {
                const double xc = P[r * QB_LDP + c0 + c];
                part0 = fma(P[r * QB_LDP + c0], xc, part0);
                if (c >= 1) part1 = fma(P[r * QB_LDP + c0 + 1], xc, part1);
            }
        }
        int j = 0;
        while (j < jbp) {
            const int gc = c0 + j;                         // block-local column == block-local diagonal row
            const uint32_t epoch = p.epoch_base + xs;
            const int par = (int)(xs & 1u);
            ++xs;
            const bool has1 = j + 1 < jbp;
            const bool row1 = (long long)gc + 1 < rows;
            const bool own0 = gc >= l0 && gc < l0 + nloc;
            const bool own1 = gc + 1 >= l0 && gc + 1 < l0 + nloc;
            // (1) CTA totals of the 32 quantities (lane v: v < 16 -> S0_v / D0_v, v >= 16 -> S1_{v-16} / D1_{v-16})
            part0 += __shfl_xor_sync(0xffffffffu, part0, 16);
            part1 += __shfl_xor_sync(0xffffffffu, part1, 16);
            st_shared_pred(pw + wid * QP_NV + (lane & 15), part0, lane < QR_NB);
            st_shared_pred(pw + wid * QP_NV + QR_NB + (lane & 15), part1, lane < QR_NB);
            __syncthreads();
            if (wid == 0) {                                      // warp-uniform
                const int q = lane & 15, hi = lane >> 4;
                const bool inq = q >= j && q < jbp;
                const bool actS = hi == 0 ? inq : (has1 && q >= j + 1 && q < jbp);
                const bool actD = inq && (hi == 0 || row1);
                const bool ownd = hi == 0 ? own0 : own1;
                double tot = 0.0;
#pragma unroll
                for (int w = 0; w < QB_THREADS / 32; ++w) tot += pw[w * QP_NV + lane];
                const double dv = P[(ownd ? (int)(gc + hi - l0) : 0) * QB_LDP + c0 + q];
                if (G <= QR_NB) {
                    ll_store(p.ll_step + 2 * (((size_t)(par * QP_NV + lane)) * QB_MAXG + b), tot, epoch, actS);
                    ll_store(p.ll_diag + par * QP_NV + lane, dv, epoch, actD && ownd);
                } else {
                    pt[lane] = tot;
                    pt[QP_NV + lane] = dv;
                }
            }
            if (G > QR_NB) {
                __syncthreads();
                const int q = lane & 15, hi = lane >> 4;         // a warp writes one 512-byte record per reader
                const bool inq = q >= j && q < jbp;
                const bool actS = hi == 0 ? inq : (has1 && q >= j + 1 && q < jbp);
                const bool actD = inq && (hi == 0 || row1) && (hi == 0 ? own0 : own1);
                const double tv = pt[lane], dv = pt[QP_NV + lane];
                for (int r = wid; r < G; r += QB_THREADS / 32) {
                    ll_store(p.ll_fan + (((size_t)(par * QB_MAXG + r)) * QB_MAXG + b) * QP_NV + lane, tv, epoch, actS);
                    ll_store(p.ll_tot + ((size_t)(par * QB_MAXG + r)) * QP_NV + lane, dv, epoch, actD);
                }
            }
            QB_T(1);
            // (2) gather
            if (G <= QR_NB) {
#pragma unroll
                for (int h2 = 0; h2 < 2; ++h2) {
                    const int vq = (tid >> 4) + QR_NB * h2, slot = tid & 15, q = vq & 15;
                    const bool actS = h2 == 0 ? (q >= j && q < jbp) : (has1 && q >= j + 1 && q < jbp);
                    const bool act = actS && slot < G;
                    const LLLine* line = p.ll_step + 2 * (((size_t)(par * QP_NV + vq)) * QB_MAXG + (slot < G ? slot : 0));
                    double acc = 0.0;
                    bool ok;
                    do {
                        ok = ll_try_load(line, epoch, acc) || !act;
                    } while (!__all_sync(0xffffffffu, ok));
                    if (!act) acc = 0.0;
#pragma unroll
                    for (int o = 8; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
                    st_shared_pred(pr + vq, acc, slot == 0);
                }
            } else {
                // warp w reads the records of CTAs w, w + 8, ...; its 32 lanes are the 32 values of a record
                const int q = lane & 15, hi = lane >> 4;
                const bool actS = hi == 0 ? (q >= j && q < jbp) : (has1 && q >= j + 1 && q < jbp);
                const LLLine* base = p.ll_fan + (((size_t)(par * QB_MAXG + b)) * QB_MAXG + wid) * QP_NV + lane;
                double v[QP_NL];
                unsigned pending = 0;
#pragma unroll
                for (int i = 0; i < QP_NL; ++i) {
                    v[i] = 0.0;
                    if (actS && wid + (QB_THREADS / 32) * i < G) pending |= 1u << i;
                }
                do {
#pragma unroll
                    for (int g4 = 0; g4 < QP_NL / 4; ++g4) {
                        unsigned long long w0[4] = {0, 0, 0, 0}, w1[4] = {0, 0, 0, 0};
                        const unsigned m4 = (pending >> (4 * g4)) & 15u;
                        ll_load4<(QB_THREADS / 32) * QP_NV * (int)sizeof(LLLine)>(
                            base + (size_t)4 * (QB_THREADS / 32) * QP_NV * g4, m4, w0, w1);
#pragma unroll
                        for (int i = 0; i < 4; ++i) {
                            double val;
                            const bool got = ll_ready(w0[i], w1[i], epoch, val) && ((m4 >> i) & 1u);
                            if (got) { v[4 * g4 + i] = val; pending &= ~(1u << (4 * g4 + i)); }
                        }
                    }
                } while (__any_sync(0xffffffffu, pending != 0));
                double acc = 0.0;
#pragma unroll
                for (int i = 0; i < QP_NL; ++i) acc += v[i];
                pw[wid * QP_NV + lane] = acc;                                  // (the step's pw has been consumed)
                __syncthreads();
                if (wid == 1) {                                                // warp-uniform: fixed-order final sums
                    double tot = 0.0;
#pragma unroll
                    for (int w = 0; w < QB_THREADS / 32; ++w) tot += pw[w * QP_NV + lane];
                    pr[lane] = tot;
                }
            }
            if (wid == 0) {
                // warp-uniform: the two diagonal-row snapshots (my private copies when G > 16)
                const int q = lane & 15, hi = lane >> 4;
                const bool need = q >= j && q < jbp && (hi == 0 || row1);
                const LLLine* line = G > QR_NB ? p.ll_tot + ((size_t)(par * QB_MAXG + b)) * QP_NV + lane
                                               : p.ll_diag + par * QP_NV + lane;
                double dv = 0.0;
                bool ok;
                do {
                    ok = ll_try_load(line, epoch, dv) || !need;
                } while (!__all_sync(0xffffffffu, ok));
                pd[lane] = need ? dv : 0.0;
            }
            __syncthreads();
            QB_T(2);
            // (3) reflector j (LAPACK dlarfg) and, through the identities above, reflector j + 1 -- computed redundantly
            //     (and bit-identically) by every thread
            const double S0j = pr[j], D1j = pd[QR_NB + j], alpha0 = pd[j];
            const double sigma0 = fma(D1j, D1j, S0j);
            double beta0 = alpha0, tau0 = 0.0, sc0 = 0.0;
            if (sigma0 != 0.0) {
                const double nrm = sqrt(fma(alpha0, alpha0, sigma0));
                beta0 = alpha0 >= 0.0 ? -nrm : nrm;
                tau0 = (beta0 - alpha0) / beta0;
                sc0 = 1.0 / (alpha0 - beta0);
            }
            // wv0_q = tau_j (v . x_q), q > j
            auto wv0 = [&](int q) { return (q > j && q < jbp) ? tau0 * fma(sc0, fma(D1j, pd[QR_NB + q], pr[q]), pd[q]) : 0.0; };
            const double w0j1 = wv0(j + 1);
            const double S0j1 = has1 ? pr[j + 1] : 0.0;
            const double a1 = D1j * sc0;                                       // entry of v_j in row gc + 1
            // T_q = x'_{j+1} . x'_q over the rows below gc + 1 (q >= j + 1)
            auto tq = [&](int q, double w0q) {
                return fma(w0j1 * w0q, sc0 * sc0 * S0j, pr[QR_NB + q] - sc0 * fma(w0q, S0j1, w0j1 * pr[q]));
            };
            int adv = 1;
            double beta1 = 0.0, tau1 = 0.0, sc1 = 0.0, alpha1 = 0.0;
            if (has1) {
                alpha1 = fma(-a1, w0j1, pd[QR_NB + j + 1]);
                const double sigma1 = tq(j + 1, w0j1);
                if (sigma1 >= QP_THETA * pr[QR_NB + j + 1] && sigma1 >= 0.0) {
                    adv = 2;
                    beta1 = alpha1;
                    if (sigma1 != 0.0) {
                        const double nrm = sqrt(fma(alpha1, alpha1, sigma1));
                        beta1 = alpha1 >= 0.0 ? -nrm : nrm;
                        tau1 = (beta1 - alpha1) / beta1;
                        sc1 = 1.0 / (alpha1 - beta1);
                    }
                }
            }
            // wv1_q = tau_{j+1} (v' . x'_q), q > j + 1 (zero when the look-ahead is not taken)
            auto wv1 = [&](int q, double w0q) {
                return (adv == 2 && q > j + 1 && q < jbp) ? tau1 * fma(sc1, tq(q, w0q), fma(-a1, w0q, pd[QR_NB + q])) : 0.0;
            };
            const int n0 = j + adv, n1 = n0 + 1;                               // the columns of the next step
            const double w0c = wv0(c), w1c = wv1(c, w0c);
            const double w0n0 = wv0(n0), w1n0 = wv1(n0, w0n0);
            const double w0n1 = wv0(n1), w1n1 = wv1(n1, w0n1);
            st_shared_pred(taus + j, tau0, tid == 0);
            st_global_pred(p.tau + gc, tau0, tid == 0 && b == 0);
            st_shared_pred(taus + (has1 ? j + 1 : j), tau1, tid == 0 && adv == 2);
            st_global_pred(p.tau + gc + 1, tau1, tid == 0 && b == 0 && adv == 2);
            QB_T(10);
            // (4) apply to my rows; gather the sums of the next step on the fly
            part0 = 0.0;
            part1 = 0.0;
            const bool two = adv == 2;
            const bool cn0 = c >= n0 && c < jbp && n0 < jbp, cn1 = c >= n1 && c < jbp && n1 < jbp;
            // row sweeps in batches of 4 (20 shared-memory loads in flight), then 2, then 1: no padded iterations
            const int l0i = (int)l0;
            auto sweep = [&](auto nu_tag, int it0) {
                constexpr int NU = decltype(nu_tag)::value;
                double xj[NU], xj1[NU], xc[NU], xn0[NU], xn1[NU];
#pragma unroll
                for (int u = 0; u < NU; ++u) {
                    const int r = g + QR_NB * (it0 + u);
                    const bool valid = r < nloc;
                    const double* row = P + (valid ? r : 0) * QB_LDP + c0;
                    xj[u] = valid ? row[j] : 0.0;
                    xj1[u] = (valid && has1) ? row[j + 1] : 0.0;
                    xc[u] = valid ? row[c] : 0.0;
                    xn0[u] = (valid && n0 < jbp) ? row[n0] : 0.0;
                    xn1[u] = (valid && n1 < jbp) ? row[n1] : 0.0;
                }
                __syncwarp();                               // every lane has read the old rows before anyone writes
#pragma unroll
                for (int u = 0; u < NU; ++u) {
                    // branch-free (selects + predicated store), as in the one-column step
                    const int r = g + QR_NB * (it0 + u);
                    const bool valid = r < nloc;
                    const int li = l0i + r;
                    const bool d0 = li == gc, d1 = two && li == gc + 1;
                    const double a = d0 ? 1.0 : xj[u] * sc0;                   // entry of v_j in this row (rows >= gc)
                    const double x1p = fma(-a, w0j1, xj1[u]);                  // column j + 1 after H_j
                    const double a2 = d1 ? 1.0 : ((two && li > gc + 1) ? x1p * sc1 : 0.0);   // entry of v_{j+1}
                    const double yc = fma(-a2, w1c, fma(-a, w0c, xc[u]));
                    const double yn0 = fma(-a2, w1n0, fma(-a, w0n0, xn0[u]));
                    const double yn1 = fma(-a2, w1n1, fma(-a, w0n1, xn1[u]));
                    double out = yc;
                    if (c == j) out = d0 ? beta0 : a;
                    if (two && c == j + 1 && li > gc) out = d1 ? beta1 : a2;
                    if (valid && li >= gc && c >= j && c < jbp) P[r * QB_LDP + c0 + c] = out;
                    const bool cnt = valid && li > gc + adv + 1;
                    part0 = fma((cnt && cn0) ? yn0 : 0.0, yc, part0);
                    part1 = fma((cnt && cn1) ? yn1 : 0.0, yc, part1);
                }
            };
            {
                int it0 = 0;
                for (; it0 + 4 <= nit; it0 += 4) sweep(std::integral_constant<int, 4>{}, it0);
                if (it0 + 2 <= nit) { sweep(std::integral_constant<int, 2>{}, it0); it0 += 2; }
                if (it0 < nit) sweep(std::integral_constant<int, 1>{}, it0);
            }
            QB_T(11);
            __syncthreads();
            QB_T(3);
            j += adv;
        }
        } else {
            // ---- partial sums for the first column of the sub-panel
            double part = 0.0;
            for (int it = 0; it < nit; ++it) {
                const int r = g + QR_NB * it;
                if (r < nloc && l0 + r > c0 && c < jbp) part = fma(P[r * QB_LDP + c0], P[r * QB_LDP + c0 + c], part);
            }
            for (int j = 0; j < jbp; ++j) {
                const int gc = c0 + j;                         // block-local column == block-local diagonal row
                const uint32_t epoch = p.epoch_base + (uint32_t)gc;
                const int par = (int)(epoch & 1u);
                // The column step is written without intra-warp divergence (uniform loops closed by warp votes, PTX-
                // predicated stores): a divergent spin in warp 0 made the whole CTA wait thousands of cycles per column.
                // Exchange = ONE hop through L2 (measured on B200: a line written by one SM is seen by a polling SM after
                // ~2000 cycles once both dies take part; lines polled by many SMs, or reductions that need a second hop,
                // cost 4000-8000, see scripts/ll_hop_probe.cu):
                //   G <= 16 CTAs: each CTA publishes 16 lines, everybody polls all of them (few readers per line);
                //   G  > 16     : each CTA PUSHES its 16 partials into a private 256-byte record of every reader
                //                 (fan[par][reader][writer][16]: one writer and one reader per line), a reader polls its
                //                 own G records with coalesced loads and sums them in CTA order.
                // Every CTA adds the same numbers in the same order, so beta / tau / scale are bit-identical everywhere.
                // (1) CTA totals of the 16 quantities
                part += __shfl_xor_sync(0xffffffffu, part, 16);
                st_shared_pred(wsum + wid * QR_NB + (lane & 15), part, lane < QR_NB);
                __syncthreads();
                const bool own = gc >= l0 && gc < l0 + nloc;         // I own the diagonal row: I publish its snapshot
                if (wid == 0) {                                      // warp-uniform
                    const int q = lane & 15;
                    const bool act = lane < QR_NB && q >= j && q < jbp;
                    double tot = 0.0;
    #pragma unroll
                    for (int w = 0; w < QB_THREADS / 32; ++w) tot += wsum[w * QR_NB + q];
                    const double dv = P[(own ? (int)(gc - l0) : 0) * QB_LDP + c0 + q];
                    if (G <= QR_NB) {
                        ll_store(p.ll_step + 2 * (((size_t)(par * QR_NB + q)) * QB_MAXG + b), tot, epoch, act);
                        ll_store(p.ll_diag + par * QR_NB + q, dv, epoch, act && own);
                    } else {
                        st_shared_pred(tot16 + q, tot, lane < QR_NB);
                        st_shared_pred(tot16 + QR_NB + q, dv, lane < QR_NB);
                    }
                }
                if (G > QR_NB) {
                    __syncthreads();
                    const int q = tid & 15, r0 = tid >> 4;
                    const bool actq = q >= j && q < jbp;
                    const double tv = tot16[q], dv = tot16[QR_NB + q];
                    for (int r = r0; r < G; r += 16) {               // (CTA-uniform trip count up to the predicate)
                        ll_store(p.ll_fan + (((size_t)(par * QB_MAXG + r)) * QB_MAXG + b) * QR_NB + q, tv, epoch, actq);
                        ll_store(p.ll_tot + ((size_t)(par * QB_MAXG + r)) * QR_NB + q, dv, epoch, actq && own);
                    }
                }
                QB_T(1);
                // (2) gather
                if (G <= QR_NB) {
                    const int q = tid >> 4, slot = tid & 15;
                    const bool act = q >= j && q < jbp && slot < G;
                    const LLLine* line = p.ll_step + 2 * (((size_t)(par * QR_NB + q)) * QB_MAXG + (slot < G ? slot : 0));
                    double acc = 0.0;
                    bool ok;
                    do {
                        ok = ll_try_load(line, epoch, acc) || !act;
                    } while (!__all_sync(0xffffffffu, ok));
                    if (!act) acc = 0.0;
    #pragma unroll
                    for (int o = 8; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
                    st_shared_pred(red + q, acc, slot == 0);
                } else {
                    // thread (q, bg): records of CTAs bg, bg + 16, ... ; a warp reads two consecutive 256-byte records
                    const int q = tid & 15, bg = tid >> 4;
                    const bool actq = q >= j && q < jbp;
                    const LLLine* base = p.ll_fan + (((size_t)(par * QB_MAXG + b)) * QB_MAXG + bg) * QR_NB + q;
                    double v[QB_NL];
                    unsigned pending = 0;
    #pragma unroll
                    for (int i = 0; i < QB_NL; ++i) {
                        v[i] = 0.0;
                        if (actq && bg + 16 * i < G) pending |= 1u << i;
                    }
                    do {
    #pragma unroll
                        for (int g4 = 0; g4 < QB_NL / 4; ++g4) {
                            unsigned long long w0[4] = {0, 0, 0, 0}, w1[4] = {0, 0, 0, 0};
                            const unsigned m4 = (pending >> (4 * g4)) & 15u;
                            ll_load4<16 * QR_NB * (int)sizeof(LLLine)>(base + (size_t)64 * QR_NB * g4, m4, w0, w1);
    #pragma unroll
                            for (int i = 0; i < 4; ++i) {
                                double val;
                                const bool got = ll_ready(w0[i], w1[i], epoch, val) && ((m4 >> i) & 1u);
                                if (got) { v[4 * g4 + i] = val; pending &= ~(1u << (4 * g4 + i)); }
                            }
                        }
                    } while (__any_sync(0xffffffffu, pending != 0));
                    double acc = 0.0;
    #pragma unroll
                    for (int i = 0; i < QB_NL; ++i) acc += v[i];
                    acc += __shfl_xor_sync(0xffffffffu, acc, 16);
                    st_shared_pred(wsum + wid * QR_NB + q, acc, lane < QR_NB);     // (the step's wsum has been consumed)
                    __syncthreads();
                    if (wid == 1) {                                                // warp-uniform: fixed-order final sums
                        double tot = 0.0;
    #pragma unroll
                        for (int w = 0; w < QB_THREADS / 32; ++w) tot += wsum[w * QR_NB + (lane & 15)];
                        st_shared_pred(red + (lane & 15), tot, lane < QR_NB);
                    }
                }
                if (wid == 0) {
                    // warp-uniform: the diagonal-row snapshot (my private copy when G > 16)
                    const int qd = lane & 15;
                    const bool need = lane < QR_NB && qd >= j && qd < jbp;
                    const LLLine* line = G > QR_NB ? p.ll_tot + ((size_t)(par * QB_MAXG + b)) * QR_NB + qd
                                                   : p.ll_diag + par * QR_NB + qd;
                    double dv = 0.0;
                    bool ok;
                    do {
                        ok = ll_try_load(line, epoch, dv) || !need;
                    } while (!__all_sync(0xffffffffu, ok));
                    st_shared_pred(drow + qd, dv, need);
                }
                __syncthreads();
                QB_T(2);
                // (3) reflector j (LAPACK dlarfg), computed redundantly (and bit-identically) by every thread
                const double sigma = red[j], alpha = drow[j];
                double beta = alpha, tau = 0.0, scale = 0.0;
                if (sigma != 0.0) {
                    const double nrm = sqrt(fma(alpha, alpha, sigma));
                    beta = alpha >= 0.0 ? -nrm : nrm;
                    tau = (beta - alpha) / beta;
                    scale = 1.0 / (alpha - beta);
                }
                st_shared_pred(taus + j, tau, tid == 0);
                st_global_pred(p.tau + gc, tau, tid == 0 && b == 0);
                const double wv_c = (c > j && c < jbp) ? tau * fma(scale, red[c], drow[c]) : 0.0;
                const double wv_n = (j + 1 < jbp) ? tau * fma(scale, red[j + 1], drow[j + 1]) : 0.0;
                QB_T(10);
                // (4) apply to my rows; gather the partial sums of column j + 1 on the fly
                part = 0.0;
                for (int it0 = 0; it0 < nit; it0 += 4) {           // 4 row sweeps per batch: 12 shared-memory loads in flight
                    double aj[4], an[4], ac[4];
    #pragma unroll
                    for (int u = 0; u < 4; ++u) {
                        const int r = g + QR_NB * (it0 + u);
                        const bool valid = it0 + u < nit && r < nloc;
                        const double* row = P + (valid ? r : 0) * QB_LDP + c0;
                        aj[u] = valid ? row[j] : 0.0;
                        an[u] = (valid && j + 1 < jbp) ? row[j + 1] : 0.0;
                        ac[u] = valid ? row[c] : 0.0;
                    }
                    __syncwarp();                               // every lane has read the old rows before anyone writes
    #pragma unroll
                    for (int u = 0; u < 4; ++u) {
                        // branch-free (selects + predicated store): lanes c < j, c == j, c > j and the diagonal row would
                        // otherwise serialise as separate divergent paths
                        const int r = g + QR_NB * (it0 + u);
                        const bool valid = it0 + u < nit && r < nloc;
                        const long long li = l0 + r;
                        const bool diag = li == gc;
                        const double vv = diag ? 1.0 : aj[u] * scale;              // reflector entry of this row
                        const double anew = fma(-vv, wv_c, ac[u]);                 // (wv_c = 0 for c <= j)
                        const double out = (c == j) ? (diag ? beta : vv) : anew;
                        if (valid && li >= gc && c >= j && c < jbp) P[r * QB_LDP + c0 + c] = out;
                        const double x = (c == j + 1) ? anew : fma(-vv, wv_n, an[u]);
                        const bool cnt = valid && li > gc + 1 && c > j && c < jbp;
                        part = fma(cnt ? x : 0.0, anew, part);
                    }
                }
                QB_T(11);
                __syncthreads();
                QB_T(3);
            }
        }
        if (tid >= jbp && tid < QR_NB) taus[tid] = 0.0;

        // ---- the rest of the block: C <- (I - V T^T V^T) C
        const int cs = c0 + jbp;                           // first remaining column
        const int ncb = JB - cs;
        if (ncb <= 0) break;
        const uint32_t epoch = p.epoch_base + (uint32_t)(c0 / QR_NB) + 1u;   // sub-panel index (own buffers)
        const int par = (int)(epoch & 1u);
        for (int idx = tid; idx < rpc * QR_NB; idx += QB_THREADS) {
            const int r = idx >> 4, k = idx & 15;
            const long long li = l0 + r;
            const int dk = c0 + k;
            double v = 0.0;
            if (r < nloc && k < jbp) v = li > dk ? P[r * QB_LDP + dk] : (li == dk ? 1.0 : 0.0);
            Vs[idx] = v;
        }
        __syncthreads();
        double* mypart = p.wpart + (size_t)b * QB_NV;
        {   // W partial: thread <-> (column pair, 4 of the 16 reflectors), all my rows
            const int cp = tid & 63, kq = tid >> 6;
            const int ca = cp < QB_WLD / 2 ? cp : QB_WLD, cb = ca + QB_WLD / 2;     // columns cp and cp + 56
            double acc[4][2];
#pragma unroll
            for (int k = 0; k < 4; ++k) acc[k][0] = acc[k][1] = 0.0;
            if (ca < ncb && l0 + nloc > c0) {
                const bool hb = cb < ncb;
                for (int r = 0; r < nloc; ++r) {
                    const double xa = P[r * QB_LDP + cs + ca];
                    const double xb = hb ? P[r * QB_LDP + cs + cb] : 0.0;
                    const double2 v01 = *reinterpret_cast<const double2*>(Vs + r * QR_NB + 4 * kq);
                    const double2 v23 = *reinterpret_cast<const double2*>(Vs + r * QR_NB + 4 * kq + 2);
                    acc[0][0] = fma(v01.x, xa, acc[0][0]); acc[0][1] = fma(v01.x, xb, acc[0][1]);
                    acc[1][0] = fma(v01.y, xa, acc[1][0]); acc[1][1] = fma(v01.y, xb, acc[1][1]);
                    acc[2][0] = fma(v23.x, xa, acc[2][0]); acc[2][1] = fma(v23.x, xb, acc[2][1]);
                    acc[3][0] = fma(v23.y, xa, acc[3][0]); acc[3][1] = fma(v23.y, xb, acc[3][1]);
                }
            }
            if (ca < QB_WLD) {
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                    mypart[(4 * kq + k) * QB_WLD + ca] = acc[k][0];
                    mypart[(4 * kq + k) * QB_WLD + cb] = acc[k][1];
                }
            }
            // Gram partial: thread <-> (i, jj), i < jj
            const int gi = tid >> 4, gj = tid & 15;
            double gacc = 0.0;
            if (gi < gj && gj < jbp && l0 + nloc > c0)
                for (int r = 0; r < nloc; ++r) gacc = fma(Vs[r * QR_NB + gi], Vs[r * QR_NB + gj], gacc);
            mypart[QR_NB * QB_WLD + tid] = gacc;
        }
        QB_T(4);
        // flag barrier: my partial is visible before my flag
        __threadfence();
        __syncthreads();
        if (tid == 0) ll_store(p.ll_bar + par * QB_MAXG + b, 0.0, epoch);
        if (tid < G) (void)ll_load(p.ll_bar + par * QB_MAXG + tid, epoch);
        __syncthreads();
        __threadfence();
        QB_T(5);
        {   // my slice of the sums over CTAs: 8 threads per value, fixed order
            const int S = (QB_NV + G - 1) / G;
            const int e = tid >> 3, p8 = tid & 7;
            for (int e0 = 0; e0 < S; e0 += QB_THREADS / 8) {
                const int idx = b * S + e0 + e;
                const bool ok = (e0 + e < S) && idx < QB_NV;
                double acc = 0.0;
                if (ok) {
                    for (int bb = p8; bb < G; bb += 32) {
                        double t[4];
#pragma unroll
                        for (int u = 0; u < 4; ++u) {
                            const int b2 = bb + 8 * u;
                            t[u] = b2 < G ? __ldcg(p.wpart + (size_t)b2 * QB_NV + idx) : 0.0;
                        }
                        acc += (t[0] + t[1]) + (t[2] + t[3]);
                    }
                }
#pragma unroll
                for (int o = 4; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
                if (ok && p8 == 0) ll_store(p.ll_res + (size_t)par * QB_NV + idx, acc, epoch);
            }
        }
        QB_T(6);
        // all-gather of the totals (only the entries in use are polled; all of a thread's lines in flight at once)
        {
            const LLLine* resl = p.ll_res + (size_t)par * QB_NV;
            double v[QB_NV / QB_THREADS];
            unsigned pending = 0;
#pragma unroll
            for (int i = 0; i < QB_NV / QB_THREADS; ++i) {
                const int idx = tid + QB_THREADS * i;
                bool need;
                if (idx < QR_NB * QB_WLD) {
                    const int k = idx / QB_WLD, cc = idx - k * QB_WLD;
                    need = k < jbp && cc < ncb;
                } else {
                    const int e = idx - QR_NB * QB_WLD, gi = e >> 4, gj = e & 15;
                    need = gi < gj && gj < jbp;
                }
                v[i] = 0.0;
                if (need) pending |= 1u << i;
            }
            do {
#pragma unroll
                for (int g4 = 0; g4 < QB_NV / QB_THREADS / 4; ++g4) {
                    unsigned long long w0[4] = {0, 0, 0, 0}, w1[4] = {0, 0, 0, 0};
                    const unsigned m4 = (pending >> (4 * g4)) & 15u;
                    ll_load4<QB_THREADS * (int)sizeof(LLLine)>(resl + tid + 4 * QB_THREADS * g4, m4, w0, w1);
#pragma unroll
                    for (int i = 0; i < 4; ++i) {
                        double val;
                        const bool got = ll_ready(w0[i], w1[i], epoch, val) && ((m4 >> i) & 1u);
                        if (got) { v[4 * g4 + i] = val; pending &= ~(1u << (4 * g4 + i)); }
                    }
                }
            } while (__any_sync(0xffffffffu, pending != 0));
#pragma unroll
            for (int i = 0; i < QB_NV / QB_THREADS; ++i) {
                const int idx = tid + QB_THREADS * i;
                if (idx < QR_NB * QB_WLD) Ws[idx] = v[i];
                else Gs[idx - QR_NB * QB_WLD] = v[i];
            }
        }
        __syncthreads();
        QB_T(7);
        if (tid < ncb) {                                   // W2 = T^T W by forward substitution, one column per thread
            double w2[QR_NB];
#pragma unroll
            for (int k = 0; k < QR_NB; ++k) {
                double sacc = (k < jbp) ? Ws[k * QB_WLD + tid] : 0.0;
#pragma unroll
                for (int l = 0; l < k; ++l) sacc = fma(-Gs[l * QR_NB + k], w2[l], sacc);
                w2[k] = taus[k] * sacc;
            }
#pragma unroll
            for (int k = 0; k < QR_NB; ++k) Ws[k * QB_WLD + tid] = w2[k];
        }
        __syncthreads();
        {   // C -= V W2: thread <-> (column pair, every 4th row)
            const int cp = tid & 63, rq = tid >> 6;
            const int ca = cp < QB_WLD / 2 ? cp : QB_WLD, cb = ca + QB_WLD / 2;
            if (ca < ncb && l0 + nloc > c0) {
                const bool hb = cb < ncb;
                double wa[QR_NB], wb[QR_NB];
#pragma unroll
                for (int k = 0; k < QR_NB; ++k) { wa[k] = Ws[k * QB_WLD + ca]; wb[k] = hb ? Ws[k * QB_WLD + cb] : 0.0; }
                for (int r = rq; r < nloc; r += 4) {
                    if (l0 + r < c0) continue;
                    const double2* v2 = reinterpret_cast<const double2*>(Vs + r * QR_NB);
                    double sa = 0.0, sb = 0.0;
#pragma unroll
                    for (int k = 0; k < QR_NB / 2; ++k) {
                        const double2 vv = v2[k];
                        sa = fma(vv.x, wa[2 * k], sa); sb = fma(vv.x, wb[2 * k], sb);
                        sa = fma(vv.y, wa[2 * k + 1], sa); sb = fma(vv.y, wb[2 * k + 1], sb);
                    }
                    P[r * QB_LDP + cs + ca] -= sa;
                    if (hb) P[r * QB_LDP + cs + cb] -= sb;
                }
            }
        }
        __syncthreads();
        QB_T(8);
    }
    __syncthreads();
    // ---- write the factored block back, and the explicit reflectors for the tensor-core update
    for (int idx = tid; idx < nloc * QR_NBO; idx += QB_THREADS) {
        const int r = idx >> 7, cc = idx & (QR_NBO - 1);
        const long long li = l0 + r;
        const double v = P[r * QB_LDP + cc];
        const double vx = (cc < JB) ? (li > cc ? v : (li == cc ? 1.0 : 0.0)) : 0.0;
        if (cc < JB) Ablk[li * p.lda + cc] = v;
        if (p.Vx != nullptr) p.Vx[li * QR_NBO + cc] = vx;
        P[r * QB_LDP + cc] = vx;                           // the slice now holds the explicit reflectors
    }
    // ---- my rows' share of the Gram matrix Vx^T Vx (the substitution kernel of the outer level needs its strict
    // upper triangle): the 36 blocks (bi <= bj) of 16 x 16, thread (ti, tj) owns entry (ti, tj) of every block.
    // A separate 128 x 128 x rows GEMM cost 55-65 us per block (one output tile, split-K + reduce); here the flops
    // ride on SMs that hold the reflectors in shared memory anyway (~4 us), a small kernel adds the G partials.
    if (p.gram_part != nullptr) {
        __syncthreads();
        const int ti = tid >> 4, tj = tid & 15;
        double acc[36];
#pragma unroll
        for (int e = 0; e < 36; ++e) acc[e] = 0.0;
        for (int r = 0; r < nloc; ++r) {
            const double* row = P + r * QB_LDP;
            double va[8], vb[8];
#pragma unroll
            for (int k = 0; k < 8; ++k) { va[k] = row[16 * k + ti]; vb[k] = row[16 * k + tj]; }
            int e = 0;
#pragma unroll
            for (int bi = 0; bi < 8; ++bi)
#pragma unroll
                for (int bj = bi; bj < 8; ++bj) { acc[e] = fma(va[bi], vb[bj], acc[e]); ++e; }
        }
        double* gp = p.gram_part + (size_t)b * 36 * QB_THREADS + tid;
#pragma unroll
        for (int e = 0; e < 36; ++e) gp[(size_t)e * QB_THREADS] = acc[e];
    }
    QB_T(9);
    if (p.prof != nullptr && tid == 0)
        for (int i = 0; i < QB_NPROF; ++i) p.prof[(size_t)b * QB_NPROF + i] = pacc[i];
}

// Gm[16 bi + ti][16 bj + tj] = sum over the G CTAs (fixed order) of the partials qr_block_coop_kernel left, for the 36
// blocks bi <= bj; the strictly lower blocks are never read (qr_w2_kernel uses striu(Gm)).  36 CTAs x 256 threads.
__global__ void __launch_bounds__(QB_THREADS) qr_gram_reduce_kernel(const double* __restrict__ part, int G, double* __restrict__ Gm) {
    const int e = blockIdx.x, tid = threadIdx.x;
    int bi = 0, rem = e;
    while (rem >= 8 - bi) { rem -= 8 - bi; ++bi; }
    const int bj = bi + rem;
    const double* src = part + (size_t)e * QB_THREADS + tid;
    double acc = 0.0;
    int g = 0;
    for (; g + 4 <= G; g += 4) {
        const double t0 = src[(size_t)g * 36 * QB_THREADS], t1 = src[(size_t)(g + 1) * 36 * QB_THREADS];
        const double t2 = src[(size_t)(g + 2) * 36 * QB_THREADS], t3 = src[(size_t)(g + 3) * 36 * QB_THREADS];
        acc += (t0 + t1) + (t2 + t3);
    }
    for (; g < G; ++g) acc += src[(size_t)g * 36 * QB_THREADS];
    Gm[(size_t)(16 * bi + (tid >> 4)) * QR_NBO + 16 * bj + (tid & 15)] = acc;
}

// W2 = op(T) W for the compact-WY factor of a block of <= 128 reflectors, T given implicitly by the Gram matrix
// Gm = V^T V and tau:  T^{-1} = striu(Gm) + diag(1 / tau).  Forward substitution gives T^T W (QR: apply H_k ... H_1),
// backward substitution T W (forming Q).  One thread per column of W, blocked by 16 rows; the 1/tau of the diagonal
// turns into a multiplication, so tau = 0 (H = I) needs no special case.  Replaces forming T (dlarft) + a GEMM.
constexpr int QW_THREADS = 64;
__global__ void __launch_bounds__(QW_THREADS) qr_w2_kernel(const double* __restrict__ Gm, const double* __restrict__ tau,
                                                           int JB, const double* __restrict__ W, long long ldw,
                                                           long long nc, double* __restrict__ W2, long long ldw2,
                                                           int transpose) {
    extern __shared__ __align__(16) double qw_smem[];
    double* Hs = qw_smem;                          // [128][128]: Hs[a][e] = coefficient of w2_a in equation e
    double* w2s = Hs + QR_NBO * QR_NBO;            // [128][QW_THREADS]
    double* ts = w2s + QR_NBO * QW_THREADS;        // [128]
    const int tid = threadIdx.x;
    for (int idx = tid; idx < QR_NBO * QR_NBO; idx += QW_THREADS) {
        const int a = idx >> 7, e = idx & (QR_NBO - 1);
        double v = 0.0;
        if (a < JB && e < JB) v = transpose ? (a < e ? Gm[idx] : 0.0) : (a > e ? Gm[e * QR_NBO + a] : 0.0);
        Hs[idx] = v;
    }
    for (int idx = tid; idx < QR_NBO; idx += QW_THREADS) ts[idx] = idx < JB ? tau[idx] : 0.0;
    __syncthreads();
    const long long col = (long long)blockIdx.x * QW_THREADS + tid;
    if (col >= nc) return;
    const int nblk = (JB + QR_NB - 1) / QR_NB;
    for (int qi = 0; qi < nblk; ++qi) {
        const int q = transpose ? qi : nblk - 1 - qi;      // equations of block q
        double acc[QR_NB];
#pragma unroll
        for (int kk = 0; kk < QR_NB; ++kk) {
            const int k = QR_NB * q + kk;
            acc[kk] = k < JB ? W[(long long)k * ldw + col] : 0.0;
        }
        for (int pi = 0; pi < qi; ++pi) {                  // blocks already solved
            const int pb = transpose ? pi : nblk - 1 - pi;
            for (int l = 0; l < QR_NB; ++l) {
                const double w = w2s[(QR_NB * pb + l) * QW_THREADS + tid];
                const double2* h2 = reinterpret_cast<const double2*>(Hs + (QR_NB * pb + l) * QR_NBO + QR_NB * q);
#pragma unroll
                for (int kk = 0; kk < QR_NB / 2; ++kk) {
                    const double2 h = h2[kk];
                    acc[2 * kk] = fma(-h.x, w, acc[2 * kk]);
                    acc[2 * kk + 1] = fma(-h.y, w, acc[2 * kk + 1]);
                }
            }
        }
        double w2[QR_NB];
        if (transpose) {
#pragma unroll
            for (int kk = 0; kk < QR_NB; ++kk) {
                double sacc = acc[kk];
#pragma unroll
                for (int l = 0; l < kk; ++l) sacc = fma(-Hs[(QR_NB * q + l) * QR_NBO + QR_NB * q + kk], w2[l], sacc);
                w2[kk] = ts[QR_NB * q + kk] * sacc;
            }
        } else {
#pragma unroll
            for (int kk = QR_NB - 1; kk >= 0; --kk) {
                double sacc = acc[kk];
#pragma unroll
                for (int l = QR_NB - 1; l > kk; --l) sacc = fma(-Hs[(QR_NB * q + l) * QR_NBO + QR_NB * q + kk], w2[l], sacc);
                w2[kk] = ts[QR_NB * q + kk] * sacc;
            }
        }
#pragma unroll
        for (int kk = 0; kk < QR_NB; ++kk) {
            const int k = QR_NB * q + kk;
            w2s[k * QW_THREADS + tid] = w2[kk];
            W2[(long long)k * ldw2 + col] = w2[kk];
        }
    }
}

// ------------------------------------------------------------------------------------------ host side
struct QrWs {
    unsigned int* counter;   // 64 B
    double* gpart;           // panel partials: 2 * sms * QR_NP
    double* drowbuf;         // 2 * QR_NB
    double* gram;            // sms * 256
    double* T;               // 256
    double* wpart;           // splits * 16 * ncols
    double* w2;              // 16 * ncols
    int max_splits;
    // outer (BLAS-3) level
    double* Vx;              // (M) x 128   explicit block reflector
    double* Gbig;            // 128 x 128
    LLLine* ll;              // LL lines of the cooperative block kernel: step | diag | bar | res
    size_t ll_bytes;
    double* cpart;           // [QB_MAXG][QB_NV] reduce-scatter partials
    double* gram_part;       // [QB_MAXG][36][256] per-CTA partials of the block's Gram matrix
    long long* prof;         // [QB_MAXG][QB_NPROF] debug cycle counters
    double* Wbig;            // 128 x (N + 1)
    double* W2big;           // 128 x (N + 1)
    void* gemm_ws; size_t gemm_ws_bytes;
};

static size_t qr_ws_layout(long long M, long long N, void* base, QrWs* out) {
    const int sms = num_sms();
    size_t off = 0;
    auto take = [&](size_t bytes) { size_t o = off; off += (bytes + 255) & ~(size_t)255; return o; };
    const size_t o_counter = take(256);
    const size_t o_gpart = take((size_t)2 * sms * QR_NP * 8);
    const size_t o_drow = take((size_t)2 * QR_NB * 8);
    const size_t o_gram = take((size_t)sms * QR_NB * QR_NB * 8);
    const size_t o_T = take(QR_NB * QR_NB * 8);
    int max_splits = (int)((M + QR_TR - 1) / QR_TR);
    if (max_splits > 2 * sms) max_splits = 2 * sms;
    if (max_splits < 1) max_splits = 1;
    const size_t o_w = take((size_t)max_splits * QR_NB * (size_t)(N > 0 ? N : 1) * 8);
    const size_t o_w2s = take((size_t)QR_NB * (size_t)(N > 0 ? N : 1) * 8);
    const size_t o_vx = take((size_t)M * QR_NBO * 8);
    const size_t o_gb = take((size_t)QR_NBO * QR_NBO * 8);
    const size_t ll_lines = (size_t)4 * QP_NV * QB_MAXG + 2 * QP_NV + (size_t)2 * QB_MAXG * QP_NV + 2 * QB_MAXG + 2 * QB_NV +
                            (size_t)2 * QB_MAXG * QB_MAXG * QP_NV;
    const size_t o_ll = take(ll_lines * sizeof(LLLine));
    const size_t o_cp = take((size_t)QB_MAXG * QB_NV * 8);
    const size_t o_gp = take((size_t)QB_MAXG * 36 * QB_THREADS * 8);
    const size_t o_pf = take((size_t)QB_MAXG * QB_NPROF * 8);
    const size_t o_wb = take((size_t)QR_NBO * (size_t)(N + 1) * 8);        // even pitch: nc + (nc & 1) <= N + 1
    const size_t o_w2 = take((size_t)QR_NBO * (size_t)(N + 1) * 8);
    // split-K partials of the block-reflector GEMMs: at most 64 splits of a 128 x nc tile row (nc <= N)
    const size_t gws = (size_t)64 * QR_NBO * (size_t)(N > QR_NBO ? N : QR_NBO) * 8;
    const size_t o_gw = take(gws + 256);
    if (out) {
        char* b = (char*)base;
        out->counter = (unsigned int*)(b + o_counter);
        out->gpart = (double*)(b + o_gpart);
        out->drowbuf = (double*)(b + o_drow);
        out->gram = (double*)(b + o_gram);
        out->T = (double*)(b + o_T);
        out->wpart = (double*)(b + o_w);
        out->w2 = (double*)(b + o_w2s);
        out->max_splits = max_splits;
        out->Vx = (double*)(b + o_vx);
        out->Gbig = (double*)(b + o_gb);
        out->ll = (LLLine*)(b + o_ll);
        out->ll_bytes = ll_lines * sizeof(LLLine);
        out->cpart = (double*)(b + o_cp);
        out->gram_part = (double*)(b + o_gp);
        out->prof = (long long*)(b + o_pf);
        out->Wbig = (double*)(b + o_wb);
        out->W2big = (double*)(b + o_w2);
        out->gemm_ws = (void*)(b + o_gw);
        out->gemm_ws_bytes = gws + 256;
    }
    return off;
}

static int run_panel(double* A, long long lda, long long M, long long j0, int jb, double* tau, const QrWs& w,
                     cudaStream_t st) {
    const long long rows = M - j0;
    int G = (int)((rows + QR_THREADS - 1) / QR_THREADS);
    const int sms = num_sms();
    if (G > sms) G = sms;
    if (G < 1) G = 1;
    PanelParams pp;
    pp.A = A; pp.lda = lda; pp.M = M; pp.j0 = j0; pp.jb = jb; pp.tau = tau + j0; pp.gpart = w.gpart;
    pp.drowbuf = w.drowbuf; pp.counter = w.counter; pp.rows_per_cta = (rows + G - 1) / G;
    bool launched = false;
    static const bool use_cluster = [] { const char* e = getenv("PLA_QR_CLUSTER"); return !(e && e[0] == '0'); }();
    if (use_cluster && rows >= 64 && rows <= (long long)QRC_MAXCS * QRC_MAXROWS) {
        // smallest cluster (1, 2, 4, 8, 16 CTAs) whose shared memory holds the panel
        int cs = 1;
        while ((long long)cs * QRC_MAXROWS < rows) cs *= 2;
        pp.rows_per_cta = (rows + cs - 1) / cs;
        const size_t smem = (size_t)(QR_NB * QRC_MAXROWS + 2 * QRC_MAXCS * QRC_XW + (QRC_THREADS / 32) * QR_NP + QR_NP + 7) * 8;
        PLA_CUDA(cudaFuncSetAttribute(qr_panel_cluster_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        if (cs > 8)
            PLA_CUDA(cudaFuncSetAttribute(qr_panel_cluster_kernel, cudaFuncAttributeNonPortableClusterSizeAllowed, 1));
        cudaLaunchConfig_t cfg = {};
        cfg.gridDim = dim3(cs); cfg.blockDim = dim3(QRC_THREADS); cfg.dynamicSmemBytes = smem; cfg.stream = st;
        cudaLaunchAttribute attr[1];
        attr[0].id = cudaLaunchAttributeClusterDimension;
        attr[0].val.clusterDim.x = cs; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
        cfg.attrs = attr; cfg.numAttrs = 1;
        launched = (cudaLaunchKernelEx(&cfg, qr_panel_cluster_kernel, pp) == cudaSuccess);
        if (launched) note_launch();
        else (void)cudaGetLastError();      // cluster not schedulable on this part: use the grid-barrier kernel
    }
    if (!launched) {
        pp.rows_per_cta = (rows + G - 1) / G;
        PLA_CUDA(cudaMemsetAsync(w.counter, 0, sizeof(unsigned int), st));
        void* args[] = {(void*)&pp};
        PLA_CUDA(cudaLaunchCooperativeKernel((const void*)qr_panel_kernel, dim3(G), dim3(QR_THREADS), args, 0, st));
        note_launch();
    }
    // compact-WY factor T
    int GG = (int)((rows + 4 * QR_TR - 1) / (4 * QR_TR));
    if (GG > sms) GG = sms;
    if (GG < 1) GG = 1;
    VView V{A, lda, j0, jb};
    qr_gram_kernel<<<GG, QR_THREADS, 0, st>>>(V, M, (rows + GG - 1) / GG, w.gram);
    PLA_LAUNCH_CHECK();
    qr_build_t_kernel<<<1, QR_THREADS, 0, st>>>(w.gram, GG, tau + j0, jb, w.T);
    PLA_LAUNCH_CHECK();
    return 0;
}

// C[j0:M, 0:nc] <- (I - V op(T) V^T) C   with V the reflectors of panel (j0, jb)
static int run_trailing(const double* Afac, long long lda, long long M, long long j0, int jb, double* C,
                        long long ldc, long long nc, int t_transpose, const QrWs& w, cudaStream_t st) {
    if (nc <= 0) return 0;
    const long long rows = M - j0;
    const int sms = num_sms();
    const int col_blocks = (int)((nc + QR_THREADS - 1) / QR_THREADS);
    long long want = (2LL * sms + col_blocks - 1) / col_blocks;       // ~2 CTAs per SM
    long long by_rows = (rows + QR_TR - 1) / QR_TR;                   // >= 64 rows (one staged tile) per split
    int splits = (int)(want < by_rows ? want : by_rows);
    if (splits > w.max_splits) splits = w.max_splits;
    if (splits < 1) splits = 1;
    TrailParams tp;
    tp.V = VView{Afac, lda, j0, jb}; tp.M = M; tp.C = C; tp.ldc = ldc; tp.nc = nc;
    tp.rows_per_split = (rows + splits - 1) / splits;
    tp.splits = (int)((rows + tp.rows_per_split - 1) / tp.rows_per_split);
    tp.wpart = w.wpart; tp.w2 = w.w2; tp.T = w.T; tp.t_transpose = t_transpose;
    dim3 grid(col_blocks, tp.splits);
    qr_trail_w_kernel<<<grid, QR_THREADS, 0, st>>>(tp);
    PLA_LAUNCH_CHECK();
    qr_trail_reduce_kernel<<<(unsigned)((nc + 31) / 32), QR_THREADS, 0, st>>>(tp);
    PLA_LAUNCH_CHECK();
    qr_trail_apply_kernel<<<grid, QR_THREADS, 0, st>>>(tp);
    PLA_LAUNCH_CHECK();
    return 0;
}

}  // namespace pla

using namespace pla;

extern "C" size_t pla_qr_workspace_bytes(int64_t M, int64_t N) { return qr_ws_layout(M, N, nullptr, nullptr); }

// Cooperative factorisation of the block (J0, JB); returns 1 when the block does not fit (caller falls back).
// *gram_ctas = number of CTAs whose Gram partials are in w.gram_part afterwards (0: none, use the GEMM).
static int run_block_coop(double* A, long long lda, long long M, long long J0, long long C0, int JB, double* tau_blk,
                          const QrWs& w, uint32_t epoch_base, cudaStream_t st, int* gram_ctas) {
    *gram_ctas = 0;
    static const int target = [] { const char* e = getenv("PLA_QR_RPC"); int v = e ? atoi(e) : 64; return v < 16 ? 64 : v; }();
    const long long rows = M - J0;
    const int sms = num_sms() < QB_MAXG ? num_sms() : QB_MAXG;
    long long G = (rows + target - 1) / target;
    if (G > sms) G = sms;
    if (G < 1) G = 1;
    long long rpc = (rows + G - 1) / G;
    rpc = (rpc + QR_NB - 1) / QR_NB * QR_NB;
    if (rpc > QB_RPC_MAX) return 1;
    G = (rows + rpc - 1) / rpc;
    BlockParams bp;
    bp.A = A; bp.lda = lda; bp.M = M; bp.J0 = J0; bp.C0 = C0; bp.JB = JB; bp.tau = tau_blk; bp.Vx = w.Vx; bp.rpc = (int)rpc;
    bp.ll_step = w.ll;
    bp.ll_diag = bp.ll_step + (size_t)4 * QP_NV * QB_MAXG;
    bp.ll_tot = bp.ll_diag + 2 * QP_NV;
    bp.ll_bar = bp.ll_tot + (size_t)2 * QB_MAXG * QP_NV;
    bp.ll_res = bp.ll_bar + 2 * QB_MAXG;
    bp.ll_fan = bp.ll_res + 2 * QB_NV;
    bp.wpart = w.cpart;
    bp.epoch_base = epoch_base;
    static const bool pair = [] { const char* e = getenv("PLA_QR_PAIR"); return !(e && e[0] == '0'); }();
    bp.pair = pair ? 1 : 0;
    static const bool gram_in = [] { const char* e = getenv("PLA_QR_GRAM"); return !(e && e[0] == '0'); }();
    bp.gram_part = (gram_in && w.Vx != nullptr) ? w.gram_part : nullptr;
    *gram_ctas = bp.gram_part != nullptr ? (int)G : 0;
    static const bool prof = [] { const char* e = getenv("PLA_QR_PROF"); return e && e[0] == '1'; }();
    bp.prof = prof ? w.prof : nullptr;
    const size_t smem = ((size_t)rpc * QB_LDP + (size_t)rpc * QR_NB + QR_NB * QB_WLD + QR_NB * QR_NB + 3 * QR_NB +
                         (QB_THREADS / 32) * QR_NB + 2 * QR_NB + 4 * QP_NV + (QB_THREADS / 32) * QP_NV) * sizeof(double);
    PLA_CUDA(cudaFuncSetAttribute(qr_block_coop_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    void* args[] = {(void*)&bp};
    PLA_CUDA(cudaLaunchCooperativeKernel((const void*)qr_block_coop_kernel, dim3((unsigned)G), dim3(QB_THREADS), args,
                                         smem, st));
    note_launch();
    if (prof) {      // debug only: synchronises
        static const char* names[QB_NPROF] = {"load", "publish", "gather", "apply", "Wpart", "barrier", "slice", "allgather",
                                              "W2+update", "store", "scalars", "applyloop"};
        long long h[QB_MAXG * QB_NPROF];
        PLA_CUDA(cudaStreamSynchronize(st));
        PLA_CUDA(cudaMemcpy(h, w.prof, (size_t)G * QB_NPROF * 8, cudaMemcpyDeviceToHost));
        fprintf(stderr, "qr_block_coop J0=%lld rows=%lld G=%lld rpc=%lld | kcycles (CTA0 / max):", J0, rows, G, rpc);
        for (int i = 0; i < QB_NPROF; ++i) {
            long long mx = 0;
            for (int g2 = 0; g2 < G; ++g2) mx = h[g2 * QB_NPROF + i] > mx ? h[g2 * QB_NPROF + i] : mx;
            fprintf(stderr, " %s %.1f/%.1f", names[i], h[i] / 1e3, mx / 1e3);
        }
        fprintf(stderr, "\n");
    }
    return 0;
}

// Explicit reflectors Vx of block (J0, JB) (unless the cooperative kernel already wrote them) and their Gram matrix.
static int build_block_reflector(const double* A, long long lda, long long M, long long J0, int JB, bool have_vx,
                                 const QrWs& w, cudaStream_t st, int gram_ctas = 0) {
    const long long rows = M - J0;
    if (have_vx && gram_ctas > 0) {          // the cooperative kernel left per-CTA partials of the Gram matrix
        qr_gram_reduce_kernel<<<36, QB_THREADS, 0, st>>>(w.gram_part, gram_ctas, w.Gbig);
        PLA_LAUNCH_CHECK();
        return 0;
    }
    if (!have_vx) {
        int nb = (int)((rows * QR_NBO + 255) / 256);
        if (nb > 8 * num_sms()) nb = 8 * num_sms();
        qr_form_v_kernel<<<nb, 256, 0, st>>>(A, lda, M, J0, JB, w.Vx);
        PLA_LAUNCH_CHECK();
    }
    return pla_gemm_f64(1, 0, QR_NBO, QR_NBO, rows, 1.0, w.Vx, QR_NBO, w.Vx, QR_NBO, 0.0, w.Gbig, QR_NBO,
                        w.gemm_ws, w.gemm_ws_bytes, st);
}

// C[J0:M, 0:nc] <- (I - Vx op(T) Vx^T) C: two DMMA GEMMs around the substitution kernel (T never formed)
static int apply_block_reflector(long long M, long long J0, int JB, const double* tau_blk, double* C, long long ldc,
                                 long long nc, int t_transpose, const QrWs& w, cudaStream_t st) {
    if (nc <= 0) return 0;
    const long long rows = M - J0;
    double* Crows = C + J0 * ldc;
    const long long ldw = nc + (nc & 1);     // 16-byte aligned rows: the second product reads W2 with vector copies
    int rc = pla_gemm_f64(1, 0, QR_NBO, nc, rows, 1.0, w.Vx, QR_NBO, Crows, ldc, 0.0, w.Wbig, ldw, w.gemm_ws,
                          w.gemm_ws_bytes, st);                                   // W = Vx^T C
    if (rc) return rc;
    const size_t smem = ((size_t)QR_NBO * QR_NBO + (size_t)QR_NBO * QW_THREADS + QR_NBO) * sizeof(double);
    PLA_CUDA(cudaFuncSetAttribute(qr_w2_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    qr_w2_kernel<<<(unsigned)((nc + QW_THREADS - 1) / QW_THREADS), QW_THREADS, smem, st>>>(
        w.Gbig, tau_blk, JB, w.Wbig, ldw, nc, w.W2big, ldw, t_transpose);         // W2 = op(T) W
    PLA_LAUNCH_CHECK();
    return pla_gemm_f64(0, 0, rows, nc, QR_NBO, -1.0, w.Vx, QR_NBO, w.W2big, ldw, 1.0, Crows, ldc, w.gemm_ws,
                        w.gemm_ws_bytes, st);                                     // C -= Vx W2
}

extern "C" int pla_geqrf_f64(double* A, int64_t M, int64_t N, int64_t lda, int64_t ncols_factor, double* tau,
                             void* ws, size_t ws_bytes, void* stream) {
    PLA_CHECK_ARG(A != nullptr, 1, "A is null");
    PLA_CHECK_ARG(M >= 1 && N >= 1, 2, "empty matrix");
    PLA_CHECK_ARG(lda >= N, 4, "lda < N");
    PLA_CHECK_ARG(ncols_factor >= 1 && ncols_factor <= N && ncols_factor <= M, 5, "ncols_factor out of range");
    PLA_CHECK_ARG(tau != nullptr, 6, "tau is null");
    PLA_CHECK_ARG(ws != nullptr && ws_bytes >= pla_qr_workspace_bytes(M, N), 8, "workspace too small");
    QrWs w;
    qr_ws_layout(M, N, ws, &w);
    cudaStream_t st = (cudaStream_t)stream;
    static const bool use_coop = [] { const char* e = getenv("PLA_QR_COOP"); return !(e && e[0] == '0'); }();
    bool ll_clean = false;
    uint32_t blk = 0;
    for (long long J0 = 0; J0 < ncols_factor; J0 += QR_NBO, ++blk) {
        const int JB = (int)((ncols_factor - J0) < QR_NBO ? (ncols_factor - J0) : QR_NBO);
        const long long nc = N - (J0 + JB);
        bool done = false;
        int gram_ctas = 0;
        if (use_coop) {
            if (!ll_clean) {                 // epochs restart with every factorisation: stale lines must not match
                PLA_CUDA(cudaMemsetAsync(w.ll, 0, w.ll_bytes, st));
                ll_clean = true;
            }
            const int rc = run_block_coop(A, lda, M, J0, J0, JB, tau + J0, w, (blk + 1u) * 256u, st, &gram_ctas);
            if (rc < 0 || rc > 1) return rc;
            done = (rc == 0);
        }
        if (!done) {
            // inner level: 16-column panels; their reflectors are applied only inside this block
            for (long long j0 = J0; j0 < J0 + JB; j0 += QR_NB) {
                const int jb = (int)((J0 + JB - j0) < QR_NB ? (J0 + JB - j0) : QR_NB);
                int rc = run_panel(A, lda, M, j0, jb, tau, w, st);
                if (rc) return rc;
                rc = run_trailing(A, lda, M, j0, jb, A + j0 + jb, lda, (J0 + JB) - (j0 + jb), /*T^T*/ 1, w, st);
                if (rc) return rc;
            }
        }
        // outer level: the whole block reflector hits the remaining columns through the tensor cores
        if (nc > 0) {
            int rc = build_block_reflector(A, lda, M, J0, JB, done, w, st, done ? gram_ctas : 0);
            if (rc) return rc;
            rc = apply_block_reflector(M, J0, JB, tau + J0, A + J0 + JB, lda, nc, /*T^T*/ 1, w, st);
            if (rc) return rc;
        }
    }
    return 0;
}

// ---- block-level entry points: a QR whose trailing columns are spread over several GPUs (distla.geqrf_distributed)
extern "C" int pla_qr_factor_block_f64(double* A, int64_t M, int64_t lda, int64_t r0, int64_t c0, int64_t jb, double* tau_blk,
                                       int block_index, int64_t n_layout, void* ws, size_t ws_bytes, void* stream) {
    PLA_CHECK_ARG(A != nullptr, 1, "A is null");
    PLA_CHECK_ARG(M >= 1 && r0 >= 0 && r0 < M, 4, "row offset out of range");
    PLA_CHECK_ARG(c0 >= 0 && jb >= 1 && jb <= QR_NBO && lda >= c0 + jb, 6, "bad block columns (jb <= 128, c0 + jb <= lda)");
    PLA_CHECK_ARG(M - r0 >= jb, 6, "block has fewer rows than columns");
    PLA_CHECK_ARG(tau_blk != nullptr && block_index >= 0, 7, "bad tau / block index");
    PLA_CHECK_ARG(n_layout >= QR_NBO && ws != nullptr && ws_bytes >= pla_qr_workspace_bytes(M, n_layout), 10,
                  "n_layout < 128 or workspace too small");
    QrWs w;
    qr_ws_layout(M, n_layout, ws, &w);
    cudaStream_t st = (cudaStream_t)stream;
    if (block_index == 0) PLA_CUDA(cudaMemsetAsync(w.ll, 0, w.ll_bytes, st));
    int gram_ctas = 0;
    int rc = run_block_coop(A, lda, M, r0, c0, (int)jb, tau_blk, w, ((uint32_t)block_index + 1u) * 256u, st, &gram_ctas);
    if (rc == 1) { set_error("pla_qr_factor_block_f64: %lld rows do not fit the cooperative block kernel", (long long)(M - r0)); return -2; }
    if (rc) return rc;
    return build_block_reflector(A, lda, M, r0, (int)jb, true, w, st, gram_ctas);
}

extern "C" int pla_qr_apply_block_f64(int64_t M, int64_t r0, int64_t jb, const double* tau_blk, double* C, int64_t ldc,
                                      int64_t nc, int64_t n_layout, void* ws, size_t ws_bytes, void* stream) {
    PLA_CHECK_ARG(M >= 1 && r0 >= 0 && r0 < M && jb >= 1 && jb <= QR_NBO, 2, "bad block");
    PLA_CHECK_ARG(tau_blk != nullptr, 4, "tau is null");
    PLA_CHECK_ARG(nc == 0 || (C != nullptr && ldc >= nc), 5, "bad C / ldc");
    PLA_CHECK_ARG(n_layout >= QR_NBO && n_layout >= nc && ws != nullptr && ws_bytes >= pla_qr_workspace_bytes(M, n_layout), 8,
                  "n_layout too small or workspace too small");
    QrWs w;
    qr_ws_layout(M, n_layout, ws, &w);           // same layout as the factor call that left Vx and the Gram matrix here
    return apply_block_reflector(M, r0, (int)jb, tau_blk, C, ldc, nc, /*T^T*/ 1, w, (cudaStream_t)stream);
}

extern "C" int pla_orgqr_f64(const double* A, int64_t M, int64_t K, int64_t lda, const double* tau, double* Q,
                             int64_t ldq, void* ws, size_t ws_bytes, void* stream) {
    PLA_CHECK_ARG(A != nullptr, 1, "A is null");
    PLA_CHECK_ARG(M >= 1 && K >= 1 && K <= M, 3, "need 1 <= K <= M");
    PLA_CHECK_ARG(lda >= K, 4, "lda < K");
    PLA_CHECK_ARG(tau != nullptr && Q != nullptr && ldq >= K, 6, "bad tau / Q / ldq");
    PLA_CHECK_ARG(ws != nullptr && ws_bytes >= pla_qr_workspace_bytes(M, K), 9, "workspace too small");
    QrWs w;
    qr_ws_layout(M, K, ws, &w);
    cudaStream_t st = (cudaStream_t)stream;
    const int sms = num_sms();
    qr_eye_kernel<<<4 * sms, 256, 0, st>>>(Q, M, K, ldq);
    PLA_LAUNCH_CHECK();
    // Q = H_1 H_2 ... H_K [I; 0]: apply the 128-column block reflectors last to first; block J0 only
    // touches rows >= J0 and columns >= J0
    const long long last = ((K - 1) / QR_NBO) * QR_NBO;
    for (long long J0 = last; J0 >= 0; J0 -= QR_NBO) {
        const int JB = (int)((K - J0) < QR_NBO ? (K - J0) : QR_NBO);
        int rc = build_block_reflector(A, lda, M, J0, JB, false, w, st);
        if (rc) return rc;
        rc = apply_block_reflector(M, J0, JB, tau + J0, Q + J0, ldq, K - J0, /*T*/ 0, w, st);
        if (rc) return rc;
    }
    return 0;
}
