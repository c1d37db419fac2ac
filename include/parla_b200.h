/* libparla_b200 -- C ABI of the B200 (sm_100a) kernels behind PARLA's randomized-sketching hot path.
 *
 * Conventions
 *   - Every matrix is ROW-MAJOR (numpy C order) fp64; `ld*` are leading dimensions in ELEMENTS.
 *   - All data pointers are DEVICE pointers unless the name ends in `_host`.
 *   - `stream` is a cudaStream_t passed as void*; every call only ENQUEUES work on it (no hidden
 *     device synchronisation) unless stated otherwise.
 *   - Return value: 0 = ok, <0 = -(index of the offending argument, 1-based), >0 = cudaError_t.
 *     pla_last_error() returns a thread-local description.  Nothing throws or exits.
 *   - The library never allocates device memory: scratch comes in through `ws` / `ws_bytes`
 *     (query the matching *_workspace_bytes()).
 *   - `istop_dev` (may be NULL): device int; when non-NULL and *istop_dev != 0 the call is a no-op.
 *     This lets a host loop enqueue LSQR iterations ahead of the convergence test.
 *
 * Each entry point names the reference call site (BallisticLA/parla v0.1.4, file:line) it replaces.
 */
#ifndef PARLA_B200_H
#define PARLA_B200_H
#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

int pla_version(void);
const char* pla_last_error(void);
int pla_num_sms(void);
/* Number of kernels this library has launched in this process so far (bench.py's gpu_launches). */
long long pla_launch_count(void);
/* Account for kernels launched on the library's behalf by a CUDA-graph replay (the graph was captured from
 * this library's own launches; a replay does not pass through the entry points that count).           */
void pla_note_launches(long long n);
/* Measurement hook: `iters` x 8 independent DMMA.8x8x4 per warp, 8 warps per CTA, ctas_per_sm CTAs per SM.
 * Used once to find the FP64 tensor-pipe peak the Gaussian sketch is graded against.               */
int pla_dmma_probe(int iters, int ctas_per_sm, double* sink, void* stream);

/* ---------------------------------------------------------------- K4: streaming GEMV pass
 * Replaces  parla/comps/preconditioning.py:30  `out = A_lift @ work`            (DOT)
 *           parla/comps/preconditioning.py:34  `np.dot(A.T, arg[:m], out=work)` (AXPY)
 *           parla/drivers/least_squares.py:344 `r = A @ x_ske - b`, :361 `ar0 = A.T @ b`
 *           parla/comps/determiter/saddle.py:199 `y = b[:m] - A @ x`
 * in ONE read of A:
 *     DOT :  u_i <- sa * (A[i,:] . w) + su * u_i            (u updated in place)
 *     AXPY:  z   <- sum_i A[i,:]^T q_i,  q = g if AXPY_G else (new) u
 *     zss[0..n) = z (zeros without AXPY),  zss[n] = sum_i u_i^2  (new u; 0 if u == NULL)
 * sa/su come from sc_dev[0], sc_dev[1] when sc_dev != NULL (device scalars), else the immediates.
 * 1 <= n <= PLA_PASS_MAX_N (odd n: <= 4096).                                                       */
#define PLA_PASS_DOT 1
#define PLA_PASS_AXPY 2
#define PLA_PASS_AXPY_G 4
#define PLA_PASS_MAX_N 8192
size_t pla_stream_pass_workspace_bytes(int64_t m, int64_t n);
int pla_stream_pass_f64(const double* A, int64_t m, int64_t n, int64_t lda, const double* w, double* u,
                        const double* g, const double* sc_dev, double sa, double su, double* zss, int flags,
                        const int* istop_dev, void* ws, size_t ws_bytes, void* stream);

/* The same pass WITHOUT its reduce launch: the per-(CTA, group) partials of [z | |u|^2] stay in `ws`
 * (parts_out[0] of them; partial b of z is ws[b * n .. b * n + n), the |u|^2 partials start at double index
 * parts_out[1]) for pla_lsqr_fused_step_f64, which sums them in the same order.  parts_out is HOST memory [2]. */
int pla_stream_pass_parts_f64(const double* A, int64_t m, int64_t n, int64_t lda, const double* w, double* u,
                              const double* g, const double* sc_dev, double sa, double su, int flags,
                              const int* istop_dev, void* ws, size_t ws_bytes, int64_t* parts_out, void* stream);
/* Row-sharded A (one process per GPU of a node): the same pass, with the sum over the ranks of
 * zss = [z | |u|^2] (the reference has no counterpart: its `np.dot(A.T, u)` at preconditioning.py:34 sees all rows)
 * fused into the reduce kernel as a one-shot all-reduce over NVLink peer memory instead of a separate NCCL call:
 * the thread that finishes column c stores {value, epoch} lines into every rank's exchange buffer and adds the
 * `world` lines of its own buffer in rank order (bit-identical on every rank).
 *   peer_recv : HOST array of `world` device pointers, peer_recv[r] = rank r's exchange buffer (pla_peer_alloc on
 *               rank r, mapped here with pla_peer_import; peer_recv[rank] is the local allocation), each of
 *               pla_peer_exchange_bytes(world, slot_lines) bytes, slot_lines >= n + 1
 *   epoch     : call counter, identical on all ranks, starting at 1 and incremented by every call that uses the
 *               buffers (flag 0 = never written).  All ranks must make the same sequence of calls.
 * 1 <= world <= 8 (world = 1 exchanges with itself).  A rank whose peers do not answer within 30 s gets NaN in
 * zss instead of hanging the GPU.                                                                             */
int pla_stream_pass_peer_f64(const double* A, int64_t m, int64_t n, int64_t lda, const double* w, double* u,
                             const double* g, const double* sc_dev, double sa, double su, double* zss, int flags,
                             const int* istop_dev, void* ws, size_t ws_bytes, void* const* peer_recv, int rank,
                             int world, int64_t slot_lines, uint32_t epoch, void* stream);
size_t pla_peer_exchange_bytes(int world, int64_t slot_lines);
/* Exchange buffers: a zero-filled cudaMalloc block (not torch's allocator: CUDA IPC exports whole allocations),
 * its 64-byte CUDA IPC handle, and the mapping of another rank's handle into this process.                   */
int pla_peer_alloc(size_t bytes, void** dev_ptr);
int pla_peer_free(void* dev_ptr);
int pla_peer_export(const void* dev_ptr, unsigned char* handle64);
int pla_peer_import(const unsigned char* handle64, void** peer_ptr);
int pla_peer_close(void* peer_ptr);

/* ---------------------------------------------------------------- K3b: triangular solve, one rhs
 * Replaces scipy.linalg.solve_triangular(R, x, trans, lower=False) at
 *   parla/comps/preconditioning.py:28,37,40,41 and parla/drivers/least_squares.py:316,363.
 * R is n x n upper triangular (strict lower part ignored).  trans = 0: R x = b;  1: R^T x = b.
 * x may alias b.                                                                                  */
int pla_trsv_upper_f64(const double* R, int64_t n, int64_t ldr, int trans, const double* b, double* x,
                       const int* istop_dev, void* stream);

/* Base case of the blocked inversion of an upper-triangular R: X[32-block diag] = inverse of the
 * matching diagonal block of R (other entries of X untouched).  The host side completes
 * X = R^{-1} level by level with pla_gemm_f64 (X12 = -X11 (R12 X22)); the explicit inverse is then
 * applied with pla_stream_pass_f64, which turns the two sequential solves per LSQR iteration
 * (preconditioning.py:28,37) into two bandwidth-bound matvecs over an L2-resident matrix.        */
int pla_trtri_diag_f64(const double* R, int64_t n, int64_t ldr, double* X, int64_t ldx, void* stream);
/* One recursion level of the triangular inverse for small blocks, every pair of the level in one launch: with the
 * s x s diagonal blocks of X already inverted (s = 32 after pla_trtri_diag_f64, then 64), X12 = -X11 (R12 X22).
 * Larger levels are DMMA GEMMs (pla_gemm_f64).                                                        */
int pla_trtri_merge_f64(const double* R, int64_t n, int64_t ldr, double* X, int64_t ldx, int64_t s, void* stream);

/* ---------------------------------------------------------------- LSQR recurrences on device
 * Replaces the scalar/vector part of parla/comps/determiter/lsqr.py:342-395 (init) and :412-526
 * (one iteration) so that the host never synchronises inside the loop.
 * State layout: dstate = PLA_LSQR_NDOUBLE doubles, istate = PLA_LSQR_NINT ints (see lsqr_step.cu).
 *   t   : M^T (A^T u~)   (preconditioned adjoint product of the UNNORMALISED u~ from the pass)
 *   zss : output of pla_stream_pass_f64 (only zss[n] = |u~|^2 is read here)
 *   bsq_dev : device scalar |b|^2 (pla_sumsq_f64);  x0 : initial iterate in preconditioned
 *             coordinates or NULL (lsqr.py:363-370)
 * After the call dstate[PLA_LSQR_SA..SU] hold the (sa, su) of the next pass.                      */
#define PLA_LSQR_NDOUBLE 32
#define PLA_LSQR_NINT 8
#define PLA_LSQR_ALFA 0
#define PLA_LSQR_BETA 1
#define PLA_LSQR_RHOBAR 2
#define PLA_LSQR_PHIBAR 3
#define PLA_LSQR_ANORM 4
#define PLA_LSQR_DDNORM 5
#define PLA_LSQR_XXNORM 6
#define PLA_LSQR_Z 7
#define PLA_LSQR_CS2 8
#define PLA_LSQR_SN2 9
#define PLA_LSQR_BNORM 10
#define PLA_LSQR_ARNORM 11
#define PLA_LSQR_XNORM 12
#define PLA_LSQR_ACOND 13
#define PLA_LSQR_RNORM 14
#define PLA_LSQR_SA 15
#define PLA_LSQR_SU 16
#define PLA_LSQR_ATOL 17
#define PLA_LSQR_BTOL 18
#define PLA_LSQR_CTOL 19
#define PLA_LSQR_ISTOP 0   /* istate: 0 running, 1..7 as lsqr.py:500-526, 100 = alfa*beta == 0 at init */
#define PLA_LSQR_ITN 1
#define PLA_LSQR_ITERLIM 2
int pla_lsqr_init_f64(int64_t n, const double* t, const double* zss, const double* bsq_dev, double atol,
                      double btol, double conlim, int iter_lim, const double* x0, double* x, double* v,
                      double* w, double* dstate, int* istate, void* stream);
int pla_lsqr_step_f64(int64_t n, const double* t, const double* zss, double* x, double* v, double* w,
                      double* dstate, int* istate, double* arnorm_hist, void* stream);
/* One launch for everything of an LSQR iteration that is not the pass over A, for small dense preconditioners
 * M (n_in x r, row-major, r <= 2048, n_in <= 4096; single GPU, delta == 0) -- parla/comps/preconditioning.py:28,37
 * and determiter/lsqr.py:421-526 of one iteration:
 *   [z | |u~|^2] = sum of the partials pla_stream_pass_parts_f64 left in pass_ws (nparts == 0: read from zss)
 *   t = M^T z ;  the step of pla_lsqr_step_f64 on (x, v, w, dstate, istate) ;  xw = M v_new  (the next pass's `w`)
 * ONE thread-block cluster of 8 CTAs: each owns a row slice of M, partials of t are exchanged through distributed
 * shared memory, the n-vector step is done redundantly by every CTA (identical arithmetic), CTA 0 writes back.
 * A no-op once istate[PLA_LSQR_ISTOP] != 0.                                                                   */
int pla_lsqr_fused_step_f64(int64_t n_in, int64_t r, const double* M, int64_t ldm, const double* pass_ws,
                            int64_t nparts, int64_t ss_offset, double* zss, double* t, double* x, double* v,
                            double* w, double* xw, double* dstate, int* istate, double* arnorm_hist, void* stream);
/* LSQR on the adjoint operator A_pc^T (under-determined branch of PcSS2, saddle.py:203-214; SPU1,
 * least_squares.py:425-494): short vector u (length r), long vectors v~, w, x (length of A's rows).
 *   init  : u = c_pc / |c_pc|
 *   init2 : after the first pass (v~ = A_pc u): alfa = |v~|
 *   head  : u <- normalised (t / alfa - alfa u), t = M^T A^T v~;  sets (sa, su) of the next pass
 *   tail  : alfa = |v~| from zss[nz], rotations, stopping tests, arnorm history
 *   long  : x += t1 w ; w <- v~/alfa + t2 w ; |w|^2 -> state (call once per long block: top rows, ridge rows)
 * `itn` is the iteration the long update belongs to (0 = initialisation); it is skipped on the device
 * when LSQR had already stopped before that iteration.                                            */
int pla_lsqr_under_init_f64(int64_t r, const double* cpc, double* u, double atol, double btol, double conlim,
                            int iter_lim, double* dstate, int* istate, void* stream);
int pla_lsqr_under_init2_f64(int64_t nz, const double* zss, double* dstate, int* istate, void* stream);
int pla_lsqr_under_head_f64(int64_t r, const double* t, double* u, double* dstate, const int* istate, void* stream);
int pla_lsqr_under_tail_f64(int64_t nz, const double* zss, double* dstate, int* istate, double* arnorm_hist,
                            void* stream);
int pla_lsqr_under_long_f64(int64_t len, const double* vt, double* x, double* w, double* dstate, const int* istate,
                            int itn, int add_to_ww, void* ws, size_t ws_bytes, void* stream);

/* ridge (delta > 0) rows of [A; sqrt(delta) I], applied implicitly (the reference materialises
 * them: parla/comps/preconditioning.py:6-13,35-36):
 *   ub <- sa * sd * xw + su * ub ;  zss[0..n) += sd * ub ;  zss[n] += |ub|^2                      */
int pla_lsqr_ridge_f64(int64_t n, double sd, const double* xw, double* ub, const double* sc_dev, double sa,
                       double su, double* zss, const int* istop_dev, void* stream);

/* ---------------------------------------------------------------- PCG recurrences on device
 * Replaces parla/comps/determiter/pcg.py:5-47 as called by PcSS1 (parla/comps/determiter/saddle.py:144-160)
 * on the normal equations (A^T A + delta I) x = A^T b - c.  The Gram product is ONE pla_stream_pass_f64
 * (flags DOT|AXPY, su = 0): gp = A^T (A p); delta * p is added here.  Single CTA, n-sized vectors;
 * dstate = PLA_LSQR_NDOUBLE doubles, istate = PLA_LSQR_NINT ints (same buffers/polling as LSQR).
 *   residual : r = rhs - (gx + delta x)  [gx == NULL: r = rhs], err = |r|; init != 0 also sets the
 *              stopping threshold tol * err (pcg.py:16,22-23)
 *   direction: init: delta1 = r.s, p = s (pcg.py:18-20); else beta = r.s / delta1, p = s + beta p, itn += 1
 *              (pcg.py:39-43); then istop = 0 while (itn < iter_lim && err > threshold) (pcg.py:26), else 1 / 7
 *   update   : hist[itn] = err; alpha = delta1 / p.(gp + delta p); x += alpha p; unless recompute:
 *              r -= alpha (gp + delta p), err = |r|   (pcg.py:28-37; recompute on iterations 0, 10, 20, ...)
 * Every kernel but the init forms is a no-op once istate[PLA_PCG_ISTOP] != 0.                          */
#define PLA_PCG_RZ 0
#define PLA_PCG_ERR 1
#define PLA_PCG_STOP_AT 2
#define PLA_PCG_ALPHA 3
#define PLA_PCG_ISTOP 0
#define PLA_PCG_ITN 1
#define PLA_PCG_ITERLIM 2
int pla_pcg_residual_f64(int64_t n, const double* rhs, const double* gx, double delta, const double* x, double* r,
                         double* dstate, int* istate, int init, double tol, void* stream);
int pla_pcg_direction_f64(int64_t n, const double* r, const double* s, double* p, double* dstate, int* istate,
                          int init, int iter_lim, void* stream);
int pla_pcg_update_f64(int64_t n, const double* gp, double delta, const double* p, double* x, double* r,
                       double* dstate, const int* istate, double* hist, int recompute, void* stream);

/* ---------------------------------------------------------------- K2: SJLT sketch
 * Replaces `S @ A` for the scipy CSC operator of parla/utils/sketching.py:51-74 applied at
 * parla/drivers/least_squares.py:303,314 (scipy.sparse csc_matvecs).
 * S is d x m with exactly k nonzeros per column, values sign/sqrt(k).
 *   pla_sjlt_plan_f64: builds the destination-major plan from the index form
 *       rows[m*k] (int32, row index of the q-th nonzero of column i at rows[i*k+q]),
 *       signs[m*k] (int8, +1/-1).  plan: header, bucket offsets (int64) and m*k packed entries (int32:
 *       source row * 2 + (sign < 0)); bucket (r, w) = nonzeros of destination r with source row in
 *       window w (<= 2^16 rows), each bucket ordered by source row (deterministic summation order).
 *   pla_sjlt_apply_f64: out[d x n] (+)= scale * S @ A,  accumulate = 0 overwrites; when bvec != NULL
 *       also out_b[r * ldob] (+)= scale * (S @ bvec)[r] (the `S @ b` of least_squares.py:314) in the same launch.
 *   Buckets longer than 8192 entries (only possible for extreme m*k/d) are left in arrival order: the
 *   sum is then correct but its rounding is not run-to-run reproducible.                           */
size_t pla_sjlt_plan_bytes(int64_t d, int64_t m, int64_t k);
size_t pla_sjlt_plan_workspace_bytes(int64_t d, int64_t m, int64_t k);
int pla_sjlt_plan_f64(const int32_t* rows, const int8_t* signs, int64_t m, int64_t k, int64_t d, void* plan,
                      void* ws, size_t ws_bytes, void* stream);
int pla_sjlt_apply_f64(const void* plan, int64_t d, int64_t m, int64_t k, const double* A, int64_t n,
                       int64_t lda, const double* bvec, double scale, double* out, int64_t ldo, double* out_b,
                       int64_t ldob, int accumulate, void* ws, size_t ws_bytes, void* stream);
/* Optional scratch (0 bytes when d*n already fills the GPU): lets small outputs split every destination
 * list over several warps; the segment sums are added in a fixed order (still deterministic).     */
size_t pla_sjlt_apply_workspace_bytes(int64_t d, int64_t n);
/* Validation hook (SYNCHRONISES): number of out-of-range row indices seen while planning.          */
int pla_sjlt_plan_status(const void* plan, int64_t* bad_index_count_host);
/* Native (throughput) operator: index form generated on device from Philox4x32-10; column i of S
 * depends only on (seed, col_offset + i), so row-sharded ranks generate consistent slices.         */
int pla_sjlt_generate(int64_t d, int64_t m, int64_t k, uint64_t seed, int64_t col_offset, int32_t* rows,
                      int8_t* signs, void* stream);

/* Adjoint of the sketching operators on one vector, out[m] = scale * S^T v  (v has d entries):
 * replaces `S.T @ v[:d]` of parla/drivers/saddlesys.py:291-292 (SPS2 folds c into b).
 *   pla_sjlt_rmatvec_f64 : S in index form (rows/signs as for pla_sjlt_plan_f64)
 *   pla_gauss_rmatvec_f64: virtual Gaussian operator G(seed)[0:d, col_offset:col_offset+m]       */
int pla_sjlt_rmatvec_f64(const int32_t* rows, const int8_t* signs, int64_t m, int64_t k, int64_t d, const double* v,
                         double scale, double* out, void* stream);
int pla_gauss_rmatvec_f64(int64_t d, int64_t m, uint64_t seed, int64_t col_offset, double scale, const double* v,
                          double* out, void* stream);

/* ---------------------------------------------------------------- SRCT sketch (subsampled cosine transform)
 * Replaces `apply_srct` (parla/utils/sketching.py:118-176: mat[perm] * e -> scipy.fft.dct(axis=0, 'ortho') -> rows r)
 * as used by `srct_operator` / `SkOpTC` (:179-201, comps/sketchers/oblivious.py:68-72).  Only the d sampled rows
 * are computed: a pruned two-level DCT whose two levels are pla_gemm_f64 calls (see csrc/srct.cu).
 *   pla_srct_weights_f64: out[i, c] = f(k_i) cos(pi (2j+1) k_i / 2m) e[j],  j = jmap ? jmap[j0+c] : j0+c,
 *       and with_sin: out[i, ncols+c] = -sgn_i f(k_i) sin(...) e[j];  f = sqrt(1/m) (k = 0) or sqrt(2/m)
 *       (jmap/e/sgn may be NULL).  Angles reduced exactly in 64-bit integers.
 *   pla_gather_rows_scale_f64: out[t, c] = e[t] * A[perm[t], c0+c]   (`mat[perm, :] * e[:, None]`, :153-155) */
int pla_srct_weights_f64(const int64_t* k, int64_t g, int64_t m, int64_t j0, int64_t ncols, const int64_t* jmap,
                         const double* e, const double* sgn, int with_sin, double* out, int64_t ldo, void* stream);
int pla_gather_rows_scale_f64(const double* A, int64_t lda, const int64_t* perm, const double* e, int64_t rows,
                              int64_t c0, int64_t nb, double* out, int64_t ldo, void* stream);

/* ---------------------------------------------------------------- FP64 tensor-core GEMM (DMMA)
 * C[M x N] = alpha * op(A) * op(B) + beta * C, op(X) = X or X^T (transa/transb = 0/1).
 * Replaces the OpenBLAS dgemm behind `S @ A` (dense S; least_squares.py:303), `A @ S`, `A.T @ S`
 * (sketchers/aware.py:166,175,179), `Y = A @ S` (rangefinders.py:186), `B = Q.T @ A` (qb.py:351,471),
 * `A -= Qi @ Bi` (qb.py:475), project_out (qb.py:612), `U = Q @ U` (svd.py:175), `C = B @ Q`,
 * `V = Q @ U` (evd.py:278,288).  Split-K partials are reduced in a fixed order (deterministic).    */
size_t pla_gemm_workspace_bytes(int64_t M, int64_t N, int64_t K);
int pla_gemm_f64(int transa, int transb, int64_t M, int64_t N, int64_t K, double alpha, const double* A,
                 int64_t lda, const double* B, int64_t ldb, double beta, double* C, int64_t ldc, void* ws,
                 size_t ws_bytes, void* stream);

/* ---------------------------------------------------------------- K1: Gaussian sketch, S never in HBM
 * Replaces `S = rng.normal(0, 1/sqrt(d), (d, m))` (parla/utils/sketching.py:23-24) followed by
 * `A_ske = S @ A` (least_squares.py:303):  out[d x n] = beta*out + scale * G(seed)[0:d, off:off+m] @ A
 * where G is the virtual Philox4x32-10 / Box-Muller operator defined in oracle/philox_ref.py.
 * When bvec != NULL the rhs is sketched in the same launch: out[:, n] = beta*out[:, n] + scale * G @ bvec
 * (`b_ske = S @ b`, least_squares.py:314), so out must have n+1 columns (ldo >= n+1).
 * pla_philox_normal_fill_f64 materialises the same operator (test hook / small dense operators).  */
size_t pla_sketch_gauss_workspace_bytes(int64_t d, int64_t n, int64_t m);
int pla_sketch_gauss_f64(const double* A, int64_t m, int64_t n, int64_t lda, const double* bvec, int64_t d,
                         uint64_t seed, int64_t col_offset, double scale, double beta, double* out, int64_t ldo,
                         void* ws, size_t ws_bytes, void* stream);
int pla_philox_normal_fill_f64(double* out, int64_t rows, int64_t cols, int64_t ldo, uint64_t seed,
                               int64_t row_offset, int64_t col_offset, double scale, void* stream);

/* ---------------------------------------------------------------- K3a/K6: Householder QR
 * Replaces scipy.linalg.qr(A_ske, mode='economic') (least_squares.py:311; LAPACK dgeqrf+dorgqr),
 * utils/linalg_wrappers.py:6-7 (orth), rangefinders.py:187, qb.py:470.
 *   pla_geqrf_f64: in-place blocked Householder QR of the leading `ncols_factor` columns of the
 *       M x N matrix A; the reflectors are applied to ALL N columns (append rhs columns to get
 *       Q^T b for free).  On exit R is in the upper triangle, the reflector vectors below the
 *       diagonal (unit diagonal implied, LAPACK layout), tau[ncols_factor].
 *   pla_orgqr_f64: Q[M x K] (K = ncols_factor) with orthonormal columns from (A, tau).
 * The panel kernel is a cooperative launch (grid barrier): grid <= number of SMs.                 */
size_t pla_qr_workspace_bytes(int64_t M, int64_t N);
int pla_geqrf_f64(double* A, int64_t M, int64_t N, int64_t lda, int64_t ncols_factor, double* tau, void* ws,
                  size_t ws_bytes, void* stream);
/* Block-level pieces of pla_geqrf_f64 for a QR whose trailing columns are spread over several GPUs (the replicated
 * d x (n+1) sketch of the row-sharded solvers, least_squares.py:311): every rank factors the same 128-column panel,
 * each rank updates only the columns it owns.
 *   factor: Householder QR of A[r0:M, c0:c0+jb] in place (jb <= 128; LAPACK conventions), tau_blk[0:jb]; leaves the
 *           explicit reflector block and its Gram matrix in `ws`.  block_index = 0 resets the workspace's exchange lines;
 *           it must increase by one per call within one factorisation.
 *   apply : C[r0:M, 0:nc] <- H_jb ... H_1 C with the reflectors left in `ws` by the last factor call.
 * Both calls must pass the same M, `n_layout` (>= 128 and >= every nc) and workspace (pla_qr_workspace_bytes(M, n_layout)). */
int pla_qr_factor_block_f64(double* A, int64_t M, int64_t lda, int64_t r0, int64_t c0, int64_t jb, double* tau_blk,
                            int block_index, int64_t n_layout, void* ws, size_t ws_bytes, void* stream);
int pla_qr_apply_block_f64(int64_t M, int64_t r0, int64_t jb, const double* tau_blk, double* C, int64_t ldc, int64_t nc,
                           int64_t n_layout, void* ws, size_t ws_bytes, void* stream);
int pla_orgqr_f64(const double* A, int64_t M, int64_t K, int64_t lda, const double* tau, double* Q, int64_t ldq,
                  void* ws, size_t ws_bytes, void* stream);

/* ---------------------------------------------------------------- small vector helpers
 * out[0] = sum x_i^2 (deterministic two-stage).  Used for |b| (least_squares.py:347).             */
int pla_sumsq_f64(const double* x, int64_t n, double* out, void* ws, size_t ws_bytes, void* stream);
size_t pla_sumsq_workspace_bytes(int64_t n);

#ifdef __cplusplus
}
#endif
#endif /* PARLA_B200_H */
