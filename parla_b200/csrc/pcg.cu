// Device-resident preconditioned conjugate gradients on the normal equations
//   (A^T A + delta I) x = A^T b - c
// as PARLA runs it (parla/comps/determiter/pcg.py:5-47, called from PcSS1, saddle.py:144-160).
//
// The O(m n) work -- the Gram product A^T (A p) -- is ONE fused pla_stream_pass_f64 (u <- A p,
// z = A^T u in a single read of A; the reference reads A twice, saddle.py:144-148).  The kernels
// here are the n-sized recurrences; all scalars stay on the device, and a stop flag turns
// run-ahead launches into no-ops exactly as in lsqr_step.cu, so the host never synchronises
// inside an iteration.
//
//   pcg.py:16-23  r = rhs - mat x0 ; d = pre(r) ; delta1 = r.d ; rel_tol = tol |r|   -> residual(init) + direction(init)
//   pcg.py:28-33  hist[i] = |r| ; q = mat d ; alpha = delta1 / d.q ; x += alpha d     -> update
//   pcg.py:34-37  r = rhs - mat x (i % 10 == 0)  |  r -= alpha q                      -> residual | update
//   pcg.py:38-43  s = pre(r) ; beta = r.s / delta1 ; d = s + beta d ; i += 1          -> direction
#include "common.cuh"
#include "../../include/parla_b200.h"

namespace pla {

constexpr int PCG_THREADS = 1024;

// r = rhs - (gx + delta x)  (gx == nullptr: x is the zero vector, r = rhs);  err = |r|
__global__ void __launch_bounds__(PCG_THREADS) pcg_residual_kernel(long long n, const double* __restrict__ rhs,
                                                                   const double* __restrict__ gx, double delta,
                                                                   const double* __restrict__ x, double* r,
                                                                   double* ds, int* is, int init, double tol) {
    if (!init && is[PLA_PCG_ISTOP] != 0) return;
    __shared__ double scratch[33];
    double acc = 0.0;
    for (long long i = threadIdx.x; i < n; i += blockDim.x) {
        const double ri = gx ? rhs[i] - fma(delta, x[i], gx[i]) : rhs[i];
        r[i] = ri;
        acc = fma(ri, ri, acc);
    }
    const double err = sqrt(block_sum(acc, scratch));
    if (threadIdx.x == 0) {
        ds[PLA_PCG_ERR] = err;
        if (init) ds[PLA_PCG_STOP_AT] = tol * err;             // pcg.py:23
    }
}

// init: delta1 = r.s ; p = s                       (pcg.py:18-20)
// else: beta = r.s / delta1 ; p = s + beta p ; i++ (pcg.py:39-43), then the loop test of pcg.py:26
__global__ void __launch_bounds__(PCG_THREADS) pcg_direction_kernel(long long n, const double* __restrict__ r,
                                                                    const double* __restrict__ s, double* p,
                                                                    double* ds, int* is, int init, int iter_lim) {
    if (!init && is[PLA_PCG_ISTOP] != 0) return;
    __shared__ double scratch[33];
    const double rz_old = ds[PLA_PCG_RZ];
    double acc = 0.0;
    for (long long i = threadIdx.x; i < n; i += blockDim.x) acc = fma(r[i], s[i], acc);
    const double rz = block_sum(acc, scratch);
    const double beta = init ? 0.0 : rz / rz_old;
    for (long long i = threadIdx.x; i < n; i += blockDim.x) p[i] = init ? s[i] : fma(beta, p[i], s[i]);
    if (threadIdx.x != 0) return;
    ds[PLA_PCG_RZ] = rz;
    int itn = 0;
    if (init) {
        is[PLA_PCG_ITERLIM] = iter_lim;
    } else {
        itn = is[PLA_PCG_ITN] + 1;
        iter_lim = is[PLA_PCG_ITERLIM];
    }
    is[PLA_PCG_ITN] = itn;
    const bool go_on = (itn < iter_lim) && (ds[PLA_PCG_ERR] > ds[PLA_PCG_STOP_AT]);
    is[PLA_PCG_ISTOP] = go_on ? 0 : (itn < iter_lim ? 1 : 7);
}

// hist[i] = |r| ; q = gp + delta p ; alpha = delta1 / p.q ; x += alpha p ; (r -= alpha q ; err = |r|)
__global__ void __launch_bounds__(PCG_THREADS) pcg_update_kernel(long long n, const double* __restrict__ gp,
                                                                 double delta, const double* __restrict__ p, double* x,
                                                                 double* r, double* ds, const int* is, double* hist,
                                                                 int recompute) {
    if (is[PLA_PCG_ISTOP] != 0) return;
    __shared__ double scratch[33];
    const double rz = ds[PLA_PCG_RZ];
    const double err_in = ds[PLA_PCG_ERR];
    double acc = 0.0;
    for (long long i = threadIdx.x; i < n; i += blockDim.x) acc = fma(p[i], fma(delta, p[i], gp[i]), acc);
    const double den = block_sum(acc, scratch);
    const double alpha = rz / den;
    double racc = 0.0;
    for (long long i = threadIdx.x; i < n; i += blockDim.x) {
        const double pi = p[i];
        x[i] = fma(alpha, pi, x[i]);
        if (!recompute) {
            const double ri = r[i] - alpha * fma(delta, pi, gp[i]);
            r[i] = ri;
            racc = fma(ri, ri, racc);
        }
    }
    double err = 0.0;
    if (!recompute) err = sqrt(block_sum(racc, scratch));
    if (threadIdx.x == 0) {
        hist[is[PLA_PCG_ITN]] = err_in;                        // pcg.py:28
        ds[PLA_PCG_ALPHA] = alpha;
        if (!recompute) ds[PLA_PCG_ERR] = err;
    }
}

}  // namespace pla

using namespace pla;

extern "C" int pla_pcg_residual_f64(int64_t n, const double* rhs, const double* gx, double delta, const double* x,
                                    double* r, double* dstate, int* istate, int init, double tol, void* stream) {
    PLA_CHECK_ARG(n >= 1, 1, "n < 1");
    PLA_CHECK_ARG(rhs && r && dstate && istate, 2, "null pointer");
    PLA_CHECK_ARG(gx == nullptr || x != nullptr, 5, "x is required with gx");
    pcg_residual_kernel<<<1, PCG_THREADS, 0, (cudaStream_t)stream>>>(n, rhs, gx, delta, x, r, dstate, istate, init, tol);
    PLA_LAUNCH_CHECK();
    return 0;
}

extern "C" int pla_pcg_direction_f64(int64_t n, const double* r, const double* s, double* p, double* dstate,
                                     int* istate, int init, int iter_lim, void* stream) {
    PLA_CHECK_ARG(n >= 1, 1, "n < 1");
    PLA_CHECK_ARG(r && s && p && dstate && istate, 2, "null pointer");
    pcg_direction_kernel<<<1, PCG_THREADS, 0, (cudaStream_t)stream>>>(n, r, s, p, dstate, istate, init, iter_lim);
    PLA_LAUNCH_CHECK();
    return 0;
}

extern "C" int pla_pcg_update_f64(int64_t n, const double* gp, double delta, const double* p, double* x, double* r,
                                  double* dstate, const int* istate, double* hist, int recompute, void* stream) {
    PLA_CHECK_ARG(n >= 1, 1, "n < 1");
    PLA_CHECK_ARG(gp && p && x && r && dstate && istate && hist, 2, "null pointer");
    pcg_update_kernel<<<1, PCG_THREADS, 0, (cudaStream_t)stream>>>(n, gp, delta, p, x, r, dstate, istate, hist,
                                                                   recompute);
    PLA_LAUNCH_CHECK();
    return 0;
}
