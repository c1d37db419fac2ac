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

// T (JB x JB upper, ld = QR_NBO) from the Gram matrix G = Vx^T Vx and tau (LAPACK dlarft, forward columnwise).
__global__ void __launch_bounds__(QR_NBO) qr_build_tbig_kernel(const double* __restrict__ G, const double* __restrict__ tau,
                                                               int JB, double* T) {
    extern __shared__ double tb_smem[];
    double* Ts = tb_smem;                       // [QR_NBO][QR_NBO + 1]
    double* gcol = Ts + QR_NBO * (QR_NBO + 1);  // [QR_NBO]
    const int r = threadIdx.x;
    for (int c = 0; c < QR_NBO; ++c) Ts[r * (QR_NBO + 1) + c] = 0.0;
    __syncthreads();
    for (int j = 0; j < JB; ++j) {
        gcol[r] = (r < j) ? G[(size_t)r * QR_NBO + j] : 0.0;
        __syncthreads();
        const double tj = tau[j];
        if (r < j) {
            double sacc = 0.0;
            for (int l = r; l < j; ++l) sacc = fma(Ts[r * (QR_NBO + 1) + l], gcol[l], sacc);
            Ts[r * (QR_NBO + 1) + j] = -tj * sacc;
        } else if (r == j) {
            Ts[j * (QR_NBO + 1) + j] = tj;
        }
        __syncthreads();
    }
    for (int c = 0; c < QR_NBO; ++c) T[(size_t)r * QR_NBO + c] = Ts[r * (QR_NBO + 1) + c];
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
    double* Tbig;            // 128 x 128
    double* Wbig;            // 128 x N
    double* W2big;           // 128 x N
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
    const size_t o_tb = take((size_t)QR_NBO * QR_NBO * 8);
    const size_t o_wb = take((size_t)QR_NBO * (size_t)N * 8);
    const size_t o_w2 = take((size_t)QR_NBO * (size_t)N * 8);
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
        out->Tbig = (double*)(b + o_tb);
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

// Build Vx (explicit reflectors J0 .. J0+JB-1) and the compact-WY factor Tbig of the whole block.
static int build_block_reflector(const double* A, long long lda, long long M, long long J0, int JB,
                                 const double* tau, const QrWs& w, cudaStream_t st) {
    const long long rows = M - J0;
    int nb = (int)((rows * QR_NBO + 255) / 256);
    if (nb > 8 * num_sms()) nb = 8 * num_sms();
    qr_form_v_kernel<<<nb, 256, 0, st>>>(A, lda, M, J0, JB, w.Vx);
    PLA_LAUNCH_CHECK();
    int rc = pla_gemm_f64(1, 0, QR_NBO, QR_NBO, rows, 1.0, w.Vx, QR_NBO, w.Vx, QR_NBO, 0.0, w.Gbig, QR_NBO,
                          w.gemm_ws, w.gemm_ws_bytes, st);
    if (rc) return rc;
    const size_t smem = (size_t)(QR_NBO * (QR_NBO + 1) + QR_NBO) * sizeof(double);
    PLA_CUDA(cudaFuncSetAttribute(qr_build_tbig_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    qr_build_tbig_kernel<<<1, QR_NBO, smem, st>>>(w.Gbig, tau + J0, JB, w.Tbig);
    PLA_LAUNCH_CHECK();
    return 0;
}

// C[J0:M, 0:nc] <- (I - Vx op(Tbig) Vx^T) C  with three DMMA GEMMs
static int apply_block_reflector(long long M, long long J0, double* C, long long ldc, long long nc, int t_transpose,
                                 const QrWs& w, cudaStream_t st) {
    if (nc <= 0) return 0;
    const long long rows = M - J0;
    double* Crows = C + J0 * ldc;
    int rc = pla_gemm_f64(1, 0, QR_NBO, nc, rows, 1.0, w.Vx, QR_NBO, Crows, ldc, 0.0, w.Wbig, nc, w.gemm_ws,
                          w.gemm_ws_bytes, st);                                   // W = Vx^T C
    if (rc) return rc;
    rc = pla_gemm_f64(t_transpose ? 1 : 0, 0, QR_NBO, nc, QR_NBO, 1.0, w.Tbig, QR_NBO, w.Wbig, nc, 0.0, w.W2big, nc,
                      w.gemm_ws, w.gemm_ws_bytes, st);                            // W2 = op(T) W
    if (rc) return rc;
    return pla_gemm_f64(0, 0, rows, nc, QR_NBO, -1.0, w.Vx, QR_NBO, w.W2big, nc, 1.0, Crows, ldc, w.gemm_ws,
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
    for (long long J0 = 0; J0 < ncols_factor; J0 += QR_NBO) {
        const int JB = (int)((ncols_factor - J0) < QR_NBO ? (ncols_factor - J0) : QR_NBO);
        // inner level: 16-column panels; their reflectors are applied only inside this block
        for (long long j0 = J0; j0 < J0 + JB; j0 += QR_NB) {
            const int jb = (int)((J0 + JB - j0) < QR_NB ? (J0 + JB - j0) : QR_NB);
            int rc = run_panel(A, lda, M, j0, jb, tau, w, st);
            if (rc) return rc;
            rc = run_trailing(A, lda, M, j0, jb, A + j0 + jb, lda, (J0 + JB) - (j0 + jb), /*T^T*/ 1, w, st);
            if (rc) return rc;
        }
        // outer level: the whole block reflector hits the remaining columns through the tensor cores
        const long long nc = N - (J0 + JB);
        if (nc > 0) {
            int rc = build_block_reflector(A, lda, M, J0, JB, tau, w, st);
            if (rc) return rc;
            rc = apply_block_reflector(M, J0, A + J0 + JB, lda, nc, /*T^T*/ 1, w, st);
            if (rc) return rc;
        }
    }
    return 0;
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
        int rc = build_block_reflector(A, lda, M, J0, JB, tau, w, st);
        if (rc) return rc;
        rc = apply_block_reflector(M, J0, Q + J0, ldq, K - J0, /*T*/ 0, w, st);
        if (rc) return rc;
    }
    return 0;
}
