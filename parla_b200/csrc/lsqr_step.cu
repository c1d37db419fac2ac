// Device-resident LSQR recurrences (single CTA; n-sized vectors only).
//
// Follows the Golub-Kahan / Paige-Saunders recurrences as PARLA runs them
// (parla/comps/determiter/lsqr.py:342-395 initialisation, :412-526 iteration, damp == 0).
// The m-sized work (A v, A^T u, |u|) is done by pla_stream_pass_f64 on an UNNORMALISED u~:
//   lsqr.py:421-425  u = A v - alfa u ; beta = |u| ; u /= beta      ->  pass with (sa, su) = (1, -alfa/beta_prev)
//   lsqr.py:427      v = A^T u - beta v                             ->  t / beta - beta v   (linearity)
// so the only quantities needed here are t = M^T A^T u~ and |u~|^2.
#include <cooperative_groups.h>
#include "common.cuh"
#include "../../include/parla_b200.h"

namespace cg = cooperative_groups;

namespace pla {

constexpr int LS_THREADS = 1024;

__device__ __forceinline__ double sgn(double a) { return (a > 0.0) - (a < 0.0); }

// lsqr.py:63-95
__device__ __forceinline__ void sym_ortho(double a, double b, double& c, double& s, double& r) {
    if (b == 0.0) { c = sgn(a); s = 0.0; r = fabs(a); }
    else if (a == 0.0) { c = 0.0; s = sgn(b); r = fabs(b); }
    else if (fabs(b) > fabs(a)) {
        const double t = a / b;
        s = sgn(b) / sqrt(1.0 + t * t);
        c = s * t;
        r = b / s;
    } else {
        const double t = b / a;
        c = sgn(a) / sqrt(1.0 + t * t);
        s = c * t;
        r = a / c;
    }
}

__global__ void __launch_bounds__(LS_THREADS) lsqr_init_kernel(long long n, const double* __restrict__ t,
                                                               const double* __restrict__ zss,
                                                               const double* __restrict__ bsq, double atol,
                                                               double btol, double ctol, int iter_lim,
                                                               const double* __restrict__ x0, double* x, double* v,
                                                               double* w, double* ds, int* is) {
    __shared__ double scratch[33];
    const double beta = sqrt(zss[n]);
    for (long long i = threadIdx.x; i < n; i += blockDim.x) x[i] = x0 ? x0[i] : 0.0;
    __syncthreads();
    double alfa = 0.0;
    if (beta > 0.0) {                                   // lsqr.py:372-375
        double acc = 0.0;
        for (long long i = threadIdx.x; i < n; i += blockDim.x) {
            const double vi = t[i] / beta;
            v[i] = vi;
            acc = fma(vi, vi, acc);
        }
        alfa = sqrt(block_sum(acc, scratch));
    } else {                                            // lsqr.py:376-378
        for (long long i = threadIdx.x; i < n; i += blockDim.x) v[i] = x[i];
    }
    __syncthreads();
    for (long long i = threadIdx.x; i < n; i += blockDim.x) {
        double vi = v[i];
        if (alfa > 0.0) { vi = vi / alfa; v[i] = vi; }  // lsqr.py:380-381
        w[i] = vi;                                      // :382
    }
    if (threadIdx.x == 0) {
        for (int i = 0; i < PLA_LSQR_NDOUBLE; ++i) ds[i] = 0.0;
        ds[PLA_LSQR_ALFA] = alfa;
        ds[PLA_LSQR_BETA] = beta;
        ds[PLA_LSQR_RHOBAR] = alfa;
        ds[PLA_LSQR_PHIBAR] = beta;
        ds[PLA_LSQR_CS2] = -1.0;
        ds[PLA_LSQR_BNORM] = sqrt(bsq[0]);
        ds[PLA_LSQR_ARNORM] = alfa * beta;
        ds[PLA_LSQR_RNORM] = beta;
        ds[PLA_LSQR_SA] = 1.0;
        ds[PLA_LSQR_SU] = beta > 0.0 ? -alfa / beta : 0.0;
        ds[PLA_LSQR_ATOL] = atol;
        ds[PLA_LSQR_BTOL] = btol;
        ds[PLA_LSQR_CTOL] = ctol;
        for (int i = 0; i < PLA_LSQR_NINT; ++i) is[i] = 0;
        is[PLA_LSQR_ITERLIM] = iter_lim;
        is[PLA_LSQR_ISTOP] = (alfa * beta == 0.0) ? 100 : 0;   // lsqr.py:392-395 early return
    }
}

// Scalar part of one LSQR step (lsqr.py:458-526): second plane rotation, norm estimates, stopping tests, state update.
// ONE thread calls it, after every thread has taken its snapshot of the state.
__device__ __forceinline__ void lsqr_step_tail(double* ds, int* is, double* hist, double alfa, double beta, double anorm,
                                               double rho, double theta, double rhobar, double phi, double phibar,
                                               double tau, double ww) {
    const double eps = 2.220446049250313e-16;
    const double ddnorm = ds[PLA_LSQR_DDNORM] + ww / (rho * rho);
    const double cs2 = ds[PLA_LSQR_CS2], sn2 = ds[PLA_LSQR_SN2], zprev = ds[PLA_LSQR_Z];
    double xxnorm = ds[PLA_LSQR_XXNORM];
    const double delta = sn2 * rho;                        // :462-471
    const double gambar = -cs2 * rho;
    const double rhs = phi - delta * zprev;
    const double zbar = rhs / gambar;
    const double xnorm = sqrt(xxnorm + zbar * zbar);
    const double gamma = sqrt(gambar * gambar + theta * theta);
    const double z = rhs / gamma;
    xxnorm += z * z;
    const double acond = anorm * sqrt(ddnorm);             // :476-483
    const double rnorm = sqrt(phibar * phibar);
    const double arnorm = alfa * fabs(tau);
    const double bnorm = ds[PLA_LSQR_BNORM];
    const double atol = ds[PLA_LSQR_ATOL], btol = ds[PLA_LSQR_BTOL], ctol = ds[PLA_LSQR_CTOL];
    const double test1 = rnorm / bnorm;                    // :500-526
    const double test2 = arnorm / (anorm * rnorm + eps);
    const double test3 = 1.0 / (acond + eps);
    const double tt1 = test1 / (1.0 + anorm * xnorm / bnorm);
    const double rtol = btol + atol * anorm * xnorm / bnorm;

    const int itn_prev = is[PLA_LSQR_ITN];
    hist[itn_prev] = ds[PLA_LSQR_ARNORM];                  // :413 (value BEFORE this step)
    const int itn = itn_prev + 1;
    int istop = 0;
    if (itn >= is[PLA_LSQR_ITERLIM]) istop = 7;
    if (1.0 + test3 <= 1.0) istop = 6;
    if (1.0 + test2 <= 1.0) istop = 5;
    if (1.0 + tt1 <= 1.0) istop = 4;
    if (test3 <= ctol) istop = 3;
    if (test2 <= atol) istop = 2;
    if (test1 <= rtol) istop = 1;

    ds[PLA_LSQR_ALFA] = alfa;
    ds[PLA_LSQR_BETA] = beta;
    ds[PLA_LSQR_RHOBAR] = rhobar;
    ds[PLA_LSQR_PHIBAR] = phibar;
    ds[PLA_LSQR_ANORM] = anorm;
    ds[PLA_LSQR_DDNORM] = ddnorm;
    ds[PLA_LSQR_XXNORM] = xxnorm;
    ds[PLA_LSQR_Z] = z;
    ds[PLA_LSQR_CS2] = gambar / gamma;
    ds[PLA_LSQR_SN2] = theta / gamma;
    ds[PLA_LSQR_ARNORM] = arnorm;
    ds[PLA_LSQR_XNORM] = xnorm;
    ds[PLA_LSQR_ACOND] = acond;
    ds[PLA_LSQR_RNORM] = rnorm;
    ds[PLA_LSQR_SA] = 1.0;
    ds[PLA_LSQR_SU] = beta > 0.0 ? -alfa / beta : 0.0;
    is[PLA_LSQR_ITN] = itn;
    is[PLA_LSQR_ISTOP] = istop;
}

__global__ void __launch_bounds__(LS_THREADS) lsqr_step_kernel(long long n, const double* __restrict__ t,
                                                               const double* __restrict__ zss, double* x, double* v,
                                                               double* w, double* ds, int* is, double* hist) {
    if (is[PLA_LSQR_ISTOP] != 0) return;
    __shared__ double scratch[33];
    // every thread snapshots the scalars it needs BEFORE the first barrier; thread 0 rewrites the
    // state only after the last one.
    double alfa = ds[PLA_LSQR_ALFA];
    double anorm = ds[PLA_LSQR_ANORM];
    const double rhobar0 = ds[PLA_LSQR_RHOBAR], phibar0 = ds[PLA_LSQR_PHIBAR];
    const double beta = sqrt(zss[n]);                    // lsqr.py:422
    const double alfa_prev = alfa;

    // ---- v = A^T u - beta v ; alfa = |v| ; v /= alfa          (lsqr.py:424-430)
    if (beta > 0.0) {
        double acc = 0.0;
        for (long long i = threadIdx.x; i < n; i += blockDim.x) {
            const double vi = t[i] / beta - beta * v[i];
            v[i] = vi;
            acc = fma(vi, vi, acc);
        }
        alfa = sqrt(block_sum(acc, scratch));
        anorm = sqrt(anorm * anorm + alfa_prev * alfa_prev + beta * beta);
    }
    // ---- |w|^2 for ddnorm (lsqr.py:453-457) while w is still the old direction
    double wacc = 0.0;
    for (long long i = threadIdx.x; i < n; i += blockDim.x) wacc = fma(w[i], w[i], wacc);
    const double ww = block_sum(wacc, scratch);

    // ---- scalar recurrences, computed redundantly by every thread (identical results)
    double cs, sn, rho;
    sym_ortho(rhobar0, beta, cs, sn, rho);    // rhobar1 == rhobar (damp = 0)
    const double theta = sn * alfa;
    const double rhobar = -cs * alfa;
    const double phi = cs * phibar0;
    const double phibar = sn * phibar0;
    const double tau = sn * phi;
    const double t1 = phi / rho, t2 = -theta / rho;
    const bool scale_v = (beta > 0.0) && (alfa > 0.0);
    for (long long i = threadIdx.x; i < n; i += blockDim.x) {
        double vi = v[i];
        if (scale_v) { vi = vi / alfa; v[i] = vi; }
        const double wi = w[i];
        x[i] = fma(t1, wi, x[i]);                          // :455
        w[i] = fma(t2, wi, vi);                            // :456
    }
    if (threadIdx.x != 0) return;
    lsqr_step_tail(ds, is, hist, alfa, beta, anorm, rho, theta, rhobar, phi, phibar, tau, ww);
}

// ------------------------------------------------------------------------------------------------
// Fused vector phase of one LSQR iteration for SMALL preconditioners (single GPU, delta == 0, dense M = R^{-1} or
// V / sigma, n_in x r row-major):
//     [z | |u~|^2] = sum of the streaming pass's per-CTA partials        (stream_pass_reduce_kernel)
//     t = M^T z                                                          (preconditioning.py:37, pass over M)
//     Golub-Kahan / plane-rotation step on v, w, x, stopping tests       (lsqr_step_kernel)
//     xw = M v_new                                                       (preconditioning.py:28 of the NEXT iteration)
// = six launches of the unfused chain, which at 2^16 x 500 cost more than the pass over A itself.  ONE thread-block
// cluster of LF_CS CTAs does all of it: CTA c owns rows [i0, i1) of M -- the same slice of z it has just reduced, so
// the partial of t over its rows needs no exchange -- the partials of t are summed through distributed shared memory
// after one cluster barrier, every CTA then performs the (cheap) n-vector step redundantly with identical arithmetic,
// so each holds the new v for its rows of M v; CTA 0 alone writes x, v, w and the state back.
constexpr int LF_THREADS = 512;
constexpr int LF_MAXK = 4;             // columns of M per thread: r <= LF_MAXK * LF_THREADS
constexpr int LF_CS = 8;               // CTAs per cluster (portable maximum)
constexpr int LF_MAX_R = LF_MAXK * LF_THREADS;
constexpr int LF_MAX_NIN = 4096;

struct FusedStepParams {
    long long n_in, r, ldm;
    const double* M;
    const double* zpart;               // [nparts][n_in] partials of z = A^T u~ ;  nparts == 0: read zss instead
    const double* sspart;              // [nparts] partials of |u~|^2
    int nparts;
    double* zss;                       // [n_in + 1]  written when nparts > 0, read otherwise
    double* t;                         // [r]   (written for callers that look at it; not read)
    double* x; double* v; double* w;   // [r]
    double* xw;                        // [n_in] M v_new
    double* ds; int* is; double* hist;
};

__global__ void __launch_bounds__(LF_THREADS, 1) lsqr_fused_step_kernel(const FusedStepParams p) {
    cg::cluster_group cluster = cg::this_cluster();
    const int rank = (int)cluster.block_rank(), CS = (int)cluster.num_blocks();
    // cluster-uniform: CTA 0 rewrites the flag only after the cluster barrier below, which every CTA reaches after this read
    if (p.is[PLA_LSQR_ISTOP] != 0) return;
    extern __shared__ double lf_sm[];
    const long long n_in = p.n_in, r = p.r;
    const long long rp = (n_in + CS - 1) / CS;
    const long long i0 = (long long)rank * rp < n_in ? (long long)rank * rp : n_in;
    const long long i1 = i0 + rp < n_in ? i0 + rp : n_in;
    const int nloc = (int)(i1 - i0);
    double* zs = lf_sm;                     // [rp + 1] my slice of z, then |u~|^2 at index nloc
    double* tp = zs + rp + 1;               // [r] partial of t over my rows of M (read by the whole cluster)
    double* ts = tp + r;                    // [r] t
    double* vs = ts + r;                    // [r] v (old, then new)
    double* wsm = vs + r;                   // [r] w (old)
    double* red = wsm + r;                  // [2][8][33]
    double* scratch = red + 2 * 8 * 33;     // [33]
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;

    // snapshot of everything CTA 0 will overwrite, taken before the cluster barrier
    double alfa = p.ds[PLA_LSQR_ALFA];
    double anorm = p.ds[PLA_LSQR_ANORM];
    const double rhobar0 = p.ds[PLA_LSQR_RHOBAR], phibar0 = p.ds[PLA_LSQR_PHIBAR];
    const double alfa_prev = alfa;
    for (long long j = tid; j < r; j += LF_THREADS) { vs[j] = p.v[j]; wsm[j] = p.w[j]; }

    // ---- (1) my slice of z and |u~|^2: the summation order of stream_pass_reduce_kernel (8 interleaved groups of
    //          partials, 8-way trees, groups added in order)
    const int nvirt = nloc + 1;
    if (p.nparts > 0) {
        const int half = tid >> 8, cl = lane, grp = wid & 7;
        for (int cb = 0; cb < nvirt; cb += 64) {
            const int lc = cb + half * 32 + cl;
            double acc = 0.0;
            if (lc < nvirt) {
                const double* src = lc < nloc ? p.zpart + (i0 + lc) : p.sspart;
                const long long stride = lc < nloc ? n_in : 1;
                for (int b0 = grp; b0 < p.nparts; b0 += 64) {
                    double tt[8];
#pragma unroll
                    for (int u = 0; u < 8; ++u) {
                        const int b = b0 + 8 * u;
                        tt[u] = b < p.nparts ? src[(size_t)b * stride] : 0.0;
                    }
                    acc += ((tt[0] + tt[1]) + (tt[2] + tt[3])) + ((tt[4] + tt[5]) + (tt[6] + tt[7]));
                }
            }
            red[(half * 8 + grp) * 33 + cl] = acc;
            __syncthreads();
            if (grp == 0 && lc < nvirt) {
                double tot = 0.0;
#pragma unroll
                for (int g2 = 0; g2 < 8; ++g2) tot += red[(half * 8 + g2) * 33 + cl];
                zs[lc] = tot;
                if (lc < nloc) p.zss[i0 + lc] = tot;
                else if (rank == 0) p.zss[n_in] = tot;
            }
            __syncthreads();
        }
    } else {
        for (int lc = tid; lc < nvirt; lc += LF_THREADS) zs[lc] = lc < nloc ? p.zss[i0 + lc] : p.zss[n_in];
        __syncthreads();
    }
    const double beta = sqrt(zs[nloc]);                  // lsqr.py:422

    // ---- (2) partial of t = M^T z over my rows: thread <-> columns tid + 512 k
    {
        double acc[LF_MAXK];
#pragma unroll
        for (int k = 0; k < LF_MAXK; ++k) acc[k] = 0.0;
        const int kc = (int)((r + LF_THREADS - 1) / LF_THREADS);
#pragma unroll 8
        for (int i = 0; i < nloc; ++i) {
            const double zi = zs[i];
            const double* __restrict__ row = p.M + (i0 + i) * p.ldm;
#pragma unroll
            for (int k = 0; k < LF_MAXK; ++k) {
                const long long j = tid + (long long)k * LF_THREADS;
                if (k < kc && j < r) acc[k] = fma(__ldg(row + j), zi, acc[k]);
            }
        }
#pragma unroll
        for (int k = 0; k < LF_MAXK; ++k) {
            const long long j = tid + (long long)k * LF_THREADS;
            if (j < r) tp[j] = acc[k];
        }
    }
    cluster.sync();
    // ---- (3) t = sum of the cluster's partials, in CTA order (identical in every CTA)
    for (long long j = tid; j < r; j += LF_THREADS) {
        double part[LF_CS];
#pragma unroll
        for (int c = 0; c < LF_CS; ++c) part[c] = c < CS ? cluster.map_shared_rank(tp, c)[j] : 0.0;
        double tot = 0.0;
#pragma unroll
        for (int c = 0; c < LF_CS; ++c) tot += part[c];
        ts[j] = tot;
        if (rank == 0) p.t[j] = tot;
    }

    // ---- (4) the step of lsqr_step_kernel on the shared-memory copies (thread j touches only its own entries)
    if (beta > 0.0) {
        double acc = 0.0;
        for (long long i = tid; i < r; i += LF_THREADS) {
            const double vi = ts[i] / beta - beta * vs[i];
            vs[i] = vi;
            acc = fma(vi, vi, acc);
        }
        alfa = sqrt(block_sum(acc, scratch));
        anorm = sqrt(anorm * anorm + alfa_prev * alfa_prev + beta * beta);
    }
    double wacc = 0.0;
    for (long long i = tid; i < r; i += LF_THREADS) wacc = fma(wsm[i], wsm[i], wacc);
    const double ww = block_sum(wacc, scratch);
    double cs, sn, rho;
    sym_ortho(rhobar0, beta, cs, sn, rho);
    const double theta = sn * alfa;
    const double rhobar = -cs * alfa;
    const double phi = cs * phibar0;
    const double phibar = sn * phibar0;
    const double tau = sn * phi;
    const double t1 = phi / rho, t2 = -theta / rho;
    const bool scale_v = (beta > 0.0) && (alfa > 0.0);
    for (long long i = tid; i < r; i += LF_THREADS) {
        double vi = vs[i];
        if (scale_v) { vi = vi / alfa; vs[i] = vi; }
        if (rank == 0) {
            const double wi = wsm[i];
            p.v[i] = vi;
            p.x[i] = fma(t1, wi, p.x[i]);
            p.w[i] = fma(t2, wi, vi);
        }
    }
    if (rank == 0 && tid == 0)
        lsqr_step_tail(p.ds, p.is, p.hist, alfa, beta, anorm, rho, theta, rhobar, phi, phibar, tau, ww);
    __syncthreads();                                     // vs complete

    // ---- (5) xw = M v_new on my rows: warp <-> row
    for (int i = wid; i < nloc; i += LF_THREADS / 32) {
        const double* __restrict__ row = p.M + (i0 + i) * p.ldm;
        double acc = 0.0;
#pragma unroll 8
        for (long long j = lane; j < r; j += 32) acc = fma(__ldg(row + j), vs[j], acc);
        acc = warp_sum(acc);
        if (lane == 0) p.xw[i0 + i] = acc;
    }
    cluster.sync();                                      // no CTA may exit while another can still read its tp
}


__global__ void __launch_bounds__(LS_THREADS) lsqr_ridge_kernel(long long n, double sd, const double* __restrict__ xw,
                                                                double* ub, const double* sc, double sa, double su,
                                                                double* zss, const int* istop) {
    if (istop != nullptr && *istop != 0) return;
    __shared__ double scratch[33];
    if (sc != nullptr) { sa = sc[0]; su = sc[1]; }
    double acc = 0.0;
    for (long long i = threadIdx.x; i < n; i += blockDim.x) {
        const double xi = xw ? xw[i] : 0.0;
        const double ui = fma(sa * sd, xi, su * ub[i]);
        ub[i] = ui;
        zss[i] = fma(sd, ui, zss[i]);
        acc = fma(ui, ui, acc);
    }
    acc = block_sum(acc, scratch);
    if (threadIdx.x == 0) zss[n] += acc;
}

// ------------------------------------------------------------------------------------------------
// LSQR on the ADJOINT operator B = A_pc^T (under-determined problems, saddle.py:203-214): the short
// vector is u (rank of the preconditioner), the long ones are v, w, x (rows of A).  With the
// unnormalised v~ kept by the streaming pass (v = v~ / alfa):
//   head : u <- (t / alfa - alfa u) normalised, beta = its norm        (lsqr.py:421-425 with roles swapped)
//   pass : v~ <- A (M u) - (beta / alfa) v~ ,  z = A^T v~ ,  |v~|^2     (lsqr.py:427-430)
//   tail : alfa = |v~|, plane rotations, stopping tests                (:434-526)
//   long : x += t1 w ;  w <- v~ / alfa + t2 w ;  |w|^2                  (:449-457)
constexpr int LSU_T1 = 20, LSU_T2 = 21, LSU_INVA = 22, LSU_WW = 23;
constexpr int LSU_LONG_ITN = 3;          // istate: iteration whose long-vector update is due

__global__ void __launch_bounds__(LS_THREADS) lsqr_under_init_kernel(long long r, const double* __restrict__ cpc,
                                                                     double* u, double atol, double btol, double ctol,
                                                                     int iter_lim, double* ds, int* is) {
    __shared__ double scratch[33];
    double acc = 0.0;
    for (long long i = threadIdx.x; i < r; i += blockDim.x) acc = fma(cpc[i], cpc[i], acc);
    const double beta = sqrt(block_sum(acc, scratch));
    for (long long i = threadIdx.x; i < r; i += blockDim.x) u[i] = beta > 0.0 ? cpc[i] / beta : cpc[i];
    if (threadIdx.x == 0) {
        for (int i = 0; i < PLA_LSQR_NDOUBLE; ++i) ds[i] = 0.0;
        ds[PLA_LSQR_BETA] = beta;
        ds[PLA_LSQR_PHIBAR] = beta;
        ds[PLA_LSQR_BNORM] = beta;
        ds[PLA_LSQR_RNORM] = beta;
        ds[PLA_LSQR_CS2] = -1.0;
        ds[PLA_LSQR_SA] = 1.0;
        ds[PLA_LSQR_SU] = 0.0;
        ds[PLA_LSQR_ATOL] = atol; ds[PLA_LSQR_BTOL] = btol; ds[PLA_LSQR_CTOL] = ctol;
        for (int i = 0; i < PLA_LSQR_NINT; ++i) is[i] = 0;
        is[PLA_LSQR_ITERLIM] = iter_lim;
    }
}

// after the first pass (v~_0 = A_pc u_0): alfa_0 = |v~_0|, w = v = v~_0 / alfa_0 (done by the long kernel, t1 = t2 = 0)
__global__ void lsqr_under_init2_kernel(long long nz, const double* __restrict__ zss, double* ds, int* is) {
    const double beta = ds[PLA_LSQR_BETA];
    const double alfa = beta > 0.0 ? sqrt(zss[nz]) : 0.0;
    ds[PLA_LSQR_ALFA] = alfa;
    ds[PLA_LSQR_RHOBAR] = alfa;
    ds[PLA_LSQR_ARNORM] = alfa * beta;
    ds[LSU_T1] = 0.0; ds[LSU_T2] = 0.0;
    ds[LSU_INVA] = alfa > 0.0 ? 1.0 / alfa : 0.0;
    ds[LSU_WW] = alfa > 0.0 ? 1.0 : 0.0;
    is[LSU_LONG_ITN] = 0;
    is[PLA_LSQR_ISTOP] = (alfa * beta == 0.0) ? 100 : 0;
}

__global__ void __launch_bounds__(LS_THREADS) lsqr_under_head_kernel(long long r, const double* __restrict__ t, double* u,
                                                                     double* ds, const int* is) {
    if (is[PLA_LSQR_ISTOP] != 0) return;
    __shared__ double scratch[33];
    const double alfa = ds[PLA_LSQR_ALFA], anorm = ds[PLA_LSQR_ANORM];
    double acc = 0.0;
    for (long long i = threadIdx.x; i < r; i += blockDim.x) {
        const double ui = t[i] / alfa - alfa * u[i];
        u[i] = ui;
        acc = fma(ui, ui, acc);
    }
    const double beta = sqrt(block_sum(acc, scratch));
    if (beta > 0.0)
        for (long long i = threadIdx.x; i < r; i += blockDim.x) u[i] = u[i] / beta;
    if (threadIdx.x == 0) {
        ds[PLA_LSQR_BETA] = beta;
        if (beta > 0.0) ds[PLA_LSQR_ANORM] = sqrt(anorm * anorm + alfa * alfa + beta * beta);
        ds[PLA_LSQR_SA] = beta > 0.0 ? 1.0 : 0.0;           // beta == 0: v~ is left as it is (lsqr.py:424)
        ds[PLA_LSQR_SU] = beta > 0.0 ? -beta / alfa : 1.0;
    }
}

__global__ void lsqr_under_tail_kernel(long long nz, const double* __restrict__ zss, double* ds, int* is, double* hist) {
    if (is[PLA_LSQR_ISTOP] != 0) return;
    const double eps = 2.220446049250313e-16;
    const double beta = ds[PLA_LSQR_BETA];
    double alfa = ds[PLA_LSQR_ALFA];
    const double alfa_old = alfa;
    if (beta > 0.0) alfa = sqrt(zss[nz]);
    double cs, sn, rho;
    sym_ortho(ds[PLA_LSQR_RHOBAR], beta, cs, sn, rho);
    const double theta = sn * alfa;
    const double rhobar = -cs * alfa;
    const double phi = cs * ds[PLA_LSQR_PHIBAR];
    const double phibar = sn * ds[PLA_LSQR_PHIBAR];
    const double tau = sn * phi;
    const double ddnorm = ds[PLA_LSQR_DDNORM] + ds[LSU_WW] / (rho * rho);
    const double cs2 = ds[PLA_LSQR_CS2], sn2 = ds[PLA_LSQR_SN2], zprev = ds[PLA_LSQR_Z];
    double xxnorm = ds[PLA_LSQR_XXNORM];
    const double delta = sn2 * rho, gambar = -cs2 * rho, rhs = phi - delta * zprev;
    const double zbar = rhs / gambar;
    const double xnorm = sqrt(xxnorm + zbar * zbar);
    const double gamma = sqrt(gambar * gambar + theta * theta);
    const double z = rhs / gamma;
    xxnorm += z * z;
    const double anorm = ds[PLA_LSQR_ANORM];
    const double acond = anorm * sqrt(ddnorm);
    const double rnorm = sqrt(phibar * phibar);
    const double arnorm = alfa * fabs(tau);
    const double bnorm = ds[PLA_LSQR_BNORM];
    const double atol = ds[PLA_LSQR_ATOL], btol = ds[PLA_LSQR_BTOL], ctol = ds[PLA_LSQR_CTOL];
    const double test1 = rnorm / bnorm, test2 = arnorm / (anorm * rnorm + eps), test3 = 1.0 / (acond + eps);
    const double tt1 = test1 / (1.0 + anorm * xnorm / bnorm), rtol = btol + atol * anorm * xnorm / bnorm;
    const int itn_prev = is[PLA_LSQR_ITN];
    hist[itn_prev] = ds[PLA_LSQR_ARNORM];
    const int itn = itn_prev + 1;
    int istop = 0;
    if (itn >= is[PLA_LSQR_ITERLIM]) istop = 7;
    if (1.0 + test3 <= 1.0) istop = 6;
    if (1.0 + test2 <= 1.0) istop = 5;
    if (1.0 + tt1 <= 1.0) istop = 4;
    if (test3 <= ctol) istop = 3;
    if (test2 <= atol) istop = 2;
    if (test1 <= rtol) istop = 1;
    ds[PLA_LSQR_ALFA] = alfa; ds[PLA_LSQR_RHOBAR] = rhobar; ds[PLA_LSQR_PHIBAR] = phibar;
    ds[PLA_LSQR_DDNORM] = ddnorm; ds[PLA_LSQR_XXNORM] = xxnorm; ds[PLA_LSQR_Z] = z;
    ds[PLA_LSQR_CS2] = gambar / gamma; ds[PLA_LSQR_SN2] = theta / gamma;
    ds[PLA_LSQR_ARNORM] = arnorm; ds[PLA_LSQR_XNORM] = xnorm; ds[PLA_LSQR_ACOND] = acond; ds[PLA_LSQR_RNORM] = rnorm;
    ds[LSU_T1] = phi / rho;
    ds[LSU_T2] = -theta / rho;
    // w <- v + t2 w with v = v~/alfa when the step produced a new v~ (beta > 0, alfa > 0); otherwise v is unchanged,
    // i.e. v~ / alfa_old
    ds[LSU_INVA] = (beta > 0.0 && alfa > 0.0) ? 1.0 / alfa : (alfa_old > 0.0 ? 1.0 / alfa_old : 0.0);
    is[PLA_LSQR_ITN] = itn;
    is[LSU_LONG_ITN] = itn;
    is[PLA_LSQR_ISTOP] = istop;
}

// x += t1 w ; w <- v~ * inva + t2 w ; part[block] = sum w_new^2.   Runs iff istate says iteration `itn` is due.
__global__ void __launch_bounds__(256) lsqr_under_long_kernel(long long len, const double* __restrict__ vt, double* x,
                                                              double* w, const double* ds, const int* is, int itn,
                                                              double* part) {
    __shared__ double scratch[33];
    if (is[LSU_LONG_ITN] != itn || (itn > 0 && is[PLA_LSQR_ITN] != itn)) { if (threadIdx.x == 0) part[blockIdx.x] = -1.0; return; }
    const double t1 = ds[LSU_T1], t2 = ds[LSU_T2], inva = ds[LSU_INVA];
    double acc = 0.0;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < len; i += (long long)gridDim.x * blockDim.x) {
        const double wi = w[i];
        x[i] = fma(t1, wi, x[i]);
        const double wn = fma(vt[i], inva, t2 * wi);
        w[i] = wn;
        acc = fma(wn, wn, acc);
    }
    acc = block_sum(acc, scratch);
    if (threadIdx.x == 0) part[blockIdx.x] = acc;
}
// ww_out[0] (+)= sum of the partials of one long-vector update (skipped launches wrote -1)
__global__ void __launch_bounds__(256) lsqr_under_long_reduce_kernel(const double* part, int nb, int add, double* ww_out) {
    __shared__ double scratch[33];
    if (part[0] < 0.0) return;
    double acc = 0.0;
    for (int i = threadIdx.x; i < nb; i += blockDim.x) acc += part[i];
    acc = block_sum(acc, scratch);
    if (threadIdx.x == 0) ww_out[0] = add ? ww_out[0] + acc : acc;
}

}  // namespace pla

using namespace pla;

extern "C" int pla_lsqr_under_init_f64(int64_t r, const double* cpc, double* u, double atol, double btol, double conlim,
                                       int iter_lim, double* dstate, int* istate, void* stream) {
    PLA_CHECK_ARG(r >= 1, 1, "r < 1");
    PLA_CHECK_ARG(cpc && u && dstate && istate, 2, "null argument");
    PLA_CHECK_ARG(iter_lim >= 1, 7, "iter_lim < 1");
    lsqr_under_init_kernel<<<1, LS_THREADS, 0, (cudaStream_t)stream>>>(r, cpc, u, atol, btol, conlim > 0 ? 1.0 / conlim : 0.0,
                                                                        iter_lim, dstate, istate);
    PLA_LAUNCH_CHECK();
    return 0;
}
extern "C" int pla_lsqr_under_init2_f64(int64_t nz, const double* zss, double* dstate, int* istate, void* stream) {
    PLA_CHECK_ARG(zss && dstate && istate, 2, "null argument");
    lsqr_under_init2_kernel<<<1, 1, 0, (cudaStream_t)stream>>>(nz, zss, dstate, istate);
    PLA_LAUNCH_CHECK();
    return 0;
}
extern "C" int pla_lsqr_under_head_f64(int64_t r, const double* t, double* u, double* dstate, const int* istate,
                                       void* stream) {
    PLA_CHECK_ARG(r >= 1 && t && u && dstate && istate, 1, "bad argument");
    lsqr_under_head_kernel<<<1, LS_THREADS, 0, (cudaStream_t)stream>>>(r, t, u, dstate, istate);
    PLA_LAUNCH_CHECK();
    return 0;
}
extern "C" int pla_lsqr_under_tail_f64(int64_t nz, const double* zss, double* dstate, int* istate, double* arnorm_hist,
                                       void* stream) {
    PLA_CHECK_ARG(zss && dstate && istate && arnorm_hist, 2, "null argument");
    lsqr_under_tail_kernel<<<1, 1, 0, (cudaStream_t)stream>>>(nz, zss, dstate, istate, arnorm_hist);
    PLA_LAUNCH_CHECK();
    return 0;
}
extern "C" int pla_lsqr_under_long_f64(int64_t len, const double* vt, double* x, double* w, double* dstate,
                                       const int* istate, int itn, int add_to_ww, void* ws, size_t ws_bytes, void* stream) {
    PLA_CHECK_ARG(len >= 1 && vt && x && w && dstate && istate, 1, "bad argument");
    int nb = (int)((len + 255) / 256);
    if (nb > 4 * num_sms()) nb = 4 * num_sms();
    PLA_CHECK_ARG(ws != nullptr && ws_bytes >= (size_t)nb * 8, 9, "workspace too small");
    lsqr_under_long_kernel<<<nb, 256, 0, (cudaStream_t)stream>>>(len, vt, x, w, dstate, istate, itn, (double*)ws);
    PLA_LAUNCH_CHECK();
    lsqr_under_long_reduce_kernel<<<1, 256, 0, (cudaStream_t)stream>>>((const double*)ws, nb, add_to_ww, dstate + LSU_WW);
    PLA_LAUNCH_CHECK();
    return 0;
}


extern "C" int pla_lsqr_init_f64(int64_t n, const double* t, const double* zss, const double* bsq_dev, double atol,
                                 double btol, double conlim, int iter_lim, const double* x0, double* x, double* v,
                                 double* w, double* dstate, int* istate, void* stream) {
    PLA_CHECK_ARG(n >= 1, 1, "n < 1");
    PLA_CHECK_ARG(t && zss && bsq_dev, 2, "null input");
    PLA_CHECK_ARG(iter_lim >= 1, 8, "iter_lim < 1");
    PLA_CHECK_ARG(x && v && w && dstate && istate, 10, "null state");
    const double ctol = conlim > 0 ? 1.0 / conlim : 0.0;
    lsqr_init_kernel<<<1, LS_THREADS, 0, (cudaStream_t)stream>>>(n, t, zss, bsq_dev, atol, btol, ctol, iter_lim, x0,
                                                                  x, v, w, dstate, istate);
    PLA_LAUNCH_CHECK();
    return 0;
}

extern "C" int pla_lsqr_step_f64(int64_t n, const double* t, const double* zss, double* x, double* v, double* w,
                                 double* dstate, int* istate, double* arnorm_hist, void* stream) {
    PLA_CHECK_ARG(n >= 1, 1, "n < 1");
    PLA_CHECK_ARG(t && zss, 2, "null input");
    PLA_CHECK_ARG(x && v && w && dstate && istate && arnorm_hist, 4, "null state");
    lsqr_step_kernel<<<1, LS_THREADS, 0, (cudaStream_t)stream>>>(n, t, zss, x, v, w, dstate, istate, arnorm_hist);
    PLA_LAUNCH_CHECK();
    return 0;
}

extern "C" int pla_lsqr_fused_step_f64(int64_t n_in, int64_t r, const double* M, int64_t ldm, const double* pass_ws,
                                       int64_t nparts, int64_t ss_offset, double* zss, double* t, double* x, double* v,
                                       double* w, double* xw, double* dstate, int* istate, double* arnorm_hist,
                                       void* stream) {
    PLA_CHECK_ARG(n_in >= 1 && n_in <= LF_MAX_NIN, 1, "n_in out of range (1..4096)");
    PLA_CHECK_ARG(r >= 1 && r <= LF_MAX_R, 2, "r out of range (1..2048)");
    PLA_CHECK_ARG(M != nullptr && ldm >= r, 4, "M is null or ldm < r");
    PLA_CHECK_ARG(nparts >= 0 && nparts <= (1 << 20) && (nparts == 0 || (pass_ws != nullptr && ss_offset >= nparts * n_in)), 6,
                  "bad partials (nparts, ss_offset)");
    PLA_CHECK_ARG(zss && t && x && v && w && xw, 8, "null vector");
    PLA_CHECK_ARG(dstate && istate && arnorm_hist, 14, "null state");
    FusedStepParams p;
    p.n_in = n_in; p.r = r; p.ldm = ldm; p.M = M;
    p.zpart = pass_ws; p.sspart = pass_ws ? pass_ws + ss_offset : nullptr; p.nparts = (int)nparts;
    p.zss = zss; p.t = t; p.x = x; p.v = v; p.w = w; p.xw = xw; p.ds = dstate; p.is = istate; p.hist = arnorm_hist;
    const long long rp = (n_in + LF_CS - 1) / LF_CS;
    const size_t smem = ((size_t)rp + 1 + 4 * (size_t)r + 2 * 8 * 33 + 33) * sizeof(double);
    PLA_CUDA(cudaFuncSetAttribute(lsqr_fused_step_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(LF_CS); cfg.blockDim = dim3(LF_THREADS); cfg.dynamicSmemBytes = smem; cfg.stream = (cudaStream_t)stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = LF_CS; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr; cfg.numAttrs = 1;
    PLA_CUDA(cudaLaunchKernelEx(&cfg, lsqr_fused_step_kernel, p));
    note_launch();
    return 0;
}

extern "C" int pla_lsqr_ridge_f64(int64_t n, double sd, const double* xw, double* ub, const double* sc_dev,
                                  double sa, double su, double* zss, const int* istop_dev, void* stream) {
    PLA_CHECK_ARG(n >= 1, 1, "n < 1");
    PLA_CHECK_ARG(ub && zss, 4, "null vector");
    lsqr_ridge_kernel<<<1, LS_THREADS, 0, (cudaStream_t)stream>>>(n, sd, xw, ub, sc_dev, sa, su, zss, istop_dev);
    PLA_LAUNCH_CHECK();
    return 0;
}
