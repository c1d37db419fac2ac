// K4: single-pass fused GEMV-N / GEMV-T over a tall row-major matrix.
//
// Replaces the two BLAS dgemv calls per LSQR iteration of the reference
// (parla/comps/preconditioning.py:30 `A_lift @ work` and :34 `np.dot(A.T, arg[:m])`, called from
// parla/comps/determiter/lsqr.py:421,427) by ONE streaming read of A:
//     u_i <- sa * (A[i,:] . w) + su * u_i ,   z <- sum_i A[i,:]^T q_i ,   ss <- sum_i u_i^2
// A row tile (R full rows = one contiguous chunk of memory) is brought into shared memory by
// the TMA engine (cp.async.bulk + mbarrier ring) and consumed twice on chip (row dots, then the
// transposed accumulation), so HBM traffic per LSQR iteration is m*n*8 bytes instead of 2*m*n*8.
// Reductions are fixed-order (thread -> warp tree -> warp list -> CTA list): run-to-run deterministic.
#include "common.cuh"
#include "../../include/parla_b200.h"

namespace pla {

constexpr int SP_CONS = 256;          // consumer threads (8 warps)
constexpr int SP_THREADS = SP_CONS + 32;
constexpr int SP_RMAX = 8;            // max rows per tile
constexpr int SP_MAX_STAGES = 8;

struct StreamPassParams {
    const double* A;
    long long m, n, lda;
    const double* w;
    double* u;
    const double* g;
    const double* sc;       // device {sa, su} or null
    double sa, su;
    double* zpart;          // [grid][n]
    double* sspart;         // [grid]
    const int* istop;
    int flags;
    int R;                  // rows per tile
    int stages;
    int use_tma;            // 1: contiguous + 16B aligned rows tiles
    long long ntiles;
};

template <int VEC, int J>
__global__ void __launch_bounds__(SP_THREADS, 1) stream_pass_kernel(const StreamPassParams p) {
    if (p.istop != nullptr && *p.istop != 0) return;
    extern __shared__ __align__(128) unsigned char sp_smem[];
    const int tid = threadIdx.x;
    const long long n = p.n;
    const int R = p.R;
    const int S = p.stages;
    const size_t stage_elems = (size_t)R * n;
    double* tiles = reinterpret_cast<double*>(sp_smem);
    size_t off = ((size_t)S * stage_elems * 8 + 15) & ~(size_t)15;
    uint64_t* full = reinterpret_cast<uint64_t*>(sp_smem + off);
    uint64_t* empty = full + SP_MAX_STAGES;
    double* red = reinterpret_cast<double*>(empty + SP_MAX_STAGES);     // [2][8 warps][RMAX]

    if (tid == 0) {
        for (int s = 0; s < S; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], SP_CONS / 32); }
        fence_mbar_init();
    }
    __syncthreads();

    const bool do_dot = (p.flags & PLA_PASS_DOT) != 0;
    const bool do_axpy = (p.flags & PLA_PASS_AXPY) != 0;
    const bool axpy_g = (p.flags & PLA_PASS_AXPY_G) != 0;

    if (tid >= SP_CONS) {
        // ------------------------------------------------------------- producer warp
        const int lane = tid - SP_CONS;
        const uint64_t pol = l2_policy_evict_first();
        int s = 0; uint32_t ph = 0;
        for (long long t = blockIdx.x; t < p.ntiles; t += gridDim.x) {
            const long long r0 = t * R;
            const int rows = (int)min((long long)R, p.m - r0);
            const size_t elems = (size_t)rows * n;
            if (lane == 0) mbar_wait(&empty[s], ph ^ 1);
            __syncwarp();
            double* dst = tiles + (size_t)s * stage_elems;
            const bool tma_ok = p.use_tma && ((elems & 1) == 0);
            if (tma_ok) {
                if (lane == 0) {
                    mbar_arrive_expect_tx(&full[s], (uint32_t)(elems * 8));
                    bulk_g2s(dst, p.A + r0 * p.lda, (uint32_t)(elems * 8), &full[s], pol);
                }
            } else {
                // generic path: strided / unaligned / odd-sized tail tile
                for (int r = 0; r < rows; ++r) {
                    const double* src = p.A + (r0 + r) * p.lda;
                    for (long long c = lane; c < n; c += 32) dst[(size_t)r * n + c] = __ldg(src + c);
                }
                __syncwarp();
                if (lane == 0) mbar_arrive(&full[s]);
            }
            if (++s == S) { s = 0; ph ^= 1; }
        }
        return;
    }

    // ----------------------------------------------------------------- consumer warps
    const int lane = tid & 31, wid = tid >> 5;
    double sa = p.sa, su = p.su;
    if (p.sc != nullptr) { sa = p.sc[0]; su = p.sc[1]; }

    // column ownership: VEC consecutive columns at c0(j) = VEC * (tid + SP_CONS * j)
    double wreg[J * VEC], zacc[J * VEC];
#pragma unroll
    for (int j = 0; j < J; ++j)
#pragma unroll
        for (int v = 0; v < VEC; ++v) {
            const long long c = (long long)VEC * (tid + SP_CONS * j) + v;
            wreg[j * VEC + v] = (do_dot && c < n) ? p.w[c] : 0.0;
            zacc[j * VEC + v] = 0.0;
        }
    double ss = 0.0;     // thread r (< RMAX) accumulates u_r^2 over its tiles

    int s = 0; uint32_t ph = 0; int flip = 0;
    for (long long t = blockIdx.x; t < p.ntiles; t += gridDim.x) {
        const long long r0 = t * R;
        const int rows = (int)min((long long)R, p.m - r0);
        // prefetch the per-row scalars of this tile (broadcast loads) before blocking on the tile
        double uo[SP_RMAX], gq[SP_RMAX];
#pragma unroll
        for (int r = 0; r < SP_RMAX; ++r) {
            uo[r] = (r < rows && p.u != nullptr) ? p.u[r0 + r] : 0.0;
            gq[r] = (r < rows && axpy_g) ? p.g[r0 + r] : 0.0;
        }
        mbar_wait(&full[s], ph);
        const double* tile = tiles + (size_t)s * stage_elems;

        double unew[SP_RMAX];
        if (do_dot) {
            double dot[SP_RMAX];
#pragma unroll
            for (int r = 0; r < SP_RMAX; ++r) dot[r] = 0.0;
#pragma unroll
            for (int r = 0; r < SP_RMAX; ++r) {
                if (r < rows) {
                    const double* row = tile + (size_t)r * n;
#pragma unroll
                    for (int j = 0; j < J; ++j) {
                        const long long c = (long long)VEC * (tid + SP_CONS * j);
                        if (c < n) {
                            if (VEC == 2) {
                                const double2 a = *reinterpret_cast<const double2*>(row + c);
                                dot[r] = fma(a.x, wreg[j * VEC], dot[r]);
                                dot[r] = fma(a.y, wreg[j * VEC + VEC - 1], dot[r]);
                            } else {
                                dot[r] = fma(row[c], wreg[j * VEC], dot[r]);
                            }
                        }
                    }
                }
            }
            double* myred = red + (size_t)flip * (SP_CONS / 32) * SP_RMAX;
#pragma unroll
            for (int r = 0; r < SP_RMAX; ++r) {
                if (r < rows) {
                    const double v = warp_sum(dot[r]);
                    if (lane == 0) myred[wid * SP_RMAX + r] = v;
                }
            }
            asm volatile("bar.sync 1, %0;" ::"n"(SP_CONS) : "memory");
#pragma unroll
            for (int r = 0; r < SP_RMAX; ++r) {
                double tot = 0.0;
                if (r < rows) {
#pragma unroll
                    for (int q = 0; q < SP_CONS / 32; ++q) tot += myred[q * SP_RMAX + r];
                }
                unew[r] = fma(sa, tot, su * uo[r]);
            }
            flip ^= 1;
            if (tid < rows) {
                // thread r owns row r of the tile
                double mine = 0.0;
#pragma unroll
                for (int r = 0; r < SP_RMAX; ++r) if (r == tid) mine = unew[r];
                p.u[r0 + tid] = mine;
                ss = fma(mine, mine, ss);
            }
        } else {
#pragma unroll
            for (int r = 0; r < SP_RMAX; ++r) unew[r] = uo[r];
            if (tid < rows) {
                double mine = 0.0;
#pragma unroll
                for (int r = 0; r < SP_RMAX; ++r) if (r == tid) mine = unew[r];
                ss = fma(mine, mine, ss);
            }
        }

        if (do_axpy) {
#pragma unroll
            for (int r = 0; r < SP_RMAX; ++r) {
                if (r < rows) {
                    const double q = axpy_g ? gq[r] : unew[r];
                    const double* row = tile + (size_t)r * n;
#pragma unroll
                    for (int j = 0; j < J; ++j) {
                        const long long c = (long long)VEC * (tid + SP_CONS * j);
                        if (c < n) {
                            if (VEC == 2) {
                                const double2 a = *reinterpret_cast<const double2*>(row + c);
                                zacc[j * VEC] = fma(a.x, q, zacc[j * VEC]);
                                zacc[j * VEC + VEC - 1] = fma(a.y, q, zacc[j * VEC + VEC - 1]);
                            } else {
                                zacc[j * VEC] = fma(row[c], q, zacc[j * VEC]);
                            }
                        }
                    }
                }
            }
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(&empty[s]);
        if (++s == S) { s = 0; ph ^= 1; }
    }

    // ----------------------------------------------------------------- per-CTA partial results
    if (do_axpy) {
        double* zp = p.zpart + (size_t)blockIdx.x * n;
#pragma unroll
        for (int j = 0; j < J; ++j)
#pragma unroll
            for (int v = 0; v < VEC; ++v) {
                const long long c = (long long)VEC * (tid + SP_CONS * j) + v;
                if (c < n) zp[c] = zacc[j * VEC + v];
            }
    }
    if (wid == 0) {
        const double tot = warp_sum(ss);     // only lanes < RMAX hold non-zero
        if (lane == 0) p.sspart[blockIdx.x] = tot;
    }
}

// zss[c] = sum_b zpart[b][c] (fixed order), zss[n] = sum_b sspart[b]
__global__ void __launch_bounds__(256) stream_pass_reduce_kernel(const double* __restrict__ zpart,
                                                                 const double* __restrict__ sspart, int nblk,
                                                                 long long n, double* __restrict__ zss, int do_axpy,
                                                                 const int* istop) {
    if (istop != nullptr && *istop != 0) return;
    const long long c = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (c < n) {
        double acc = 0.0;
        if (do_axpy)
            for (int b = 0; b < nblk; ++b) acc += zpart[(size_t)b * n + c];
        zss[c] = acc;
    }
    if (c == n) {
        double acc = 0.0;
        for (int b = 0; b < nblk; ++b) acc += sspart[b];
        zss[n] = acc;
    }
}

static int pick_tile_rows(long long n, size_t tile_bytes_target) {
    long long R = (long long)(tile_bytes_target / (size_t)(8 * n));
    if (R < 1) R = 1;
    if (R > SP_RMAX) R = SP_RMAX;
    if ((n & 1) && (R & 1)) R = (R < SP_RMAX) ? R + 1 : R - 1;   // keep R*n even -> 16-byte tiles
    return (int)R;
}

template <int VEC, int J>
static cudaError_t launch_pass(const StreamPassParams& p, int grid, size_t smem, cudaStream_t st) {
    cudaError_t e = cudaFuncSetAttribute(stream_pass_kernel<VEC, J>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         (int)smem);
    if (e != cudaSuccess) return e;
    stream_pass_kernel<VEC, J><<<grid, SP_THREADS, smem, st>>>(p);
    return cudaGetLastError();
}

static size_t sp_tile_bytes_target() {
    static size_t v = 0;
    if (v == 0) {
        const char* e = getenv("PLA_PASS_TILE_BYTES");
        v = e ? (size_t)atoll(e) : (size_t)32768;
        if (v < 1024) v = 1024;
    }
    return v;
}

}  // namespace pla

using namespace pla;

extern "C" size_t pla_stream_pass_workspace_bytes(int64_t m, int64_t n) {
    (void)m;
    return (size_t)num_sms() * (size_t)(n + 1) * sizeof(double) + 256;
}

extern "C" int pla_stream_pass_f64(const double* A, int64_t m, int64_t n, int64_t lda, const double* w, double* u,
                                   const double* g, const double* sc_dev, double sa, double su, double* zss,
                                   int flags, const int* istop_dev, void* ws, size_t ws_bytes, void* stream) {
    PLA_CHECK_ARG(A != nullptr, 1, "A is null");
    PLA_CHECK_ARG(m >= 1, 2, "m < 1");
    PLA_CHECK_ARG(n >= 1 && n <= PLA_PASS_MAX_N, 3, "n out of range for the streaming pass (1..8192)");
    PLA_CHECK_ARG(lda >= n, 4, "lda < n");
    const bool do_dot = flags & PLA_PASS_DOT, do_axpy = flags & PLA_PASS_AXPY;
    PLA_CHECK_ARG(!do_dot || (w != nullptr && u != nullptr), 5, "DOT needs w and u");
    PLA_CHECK_ARG(do_dot || u != nullptr || (flags & PLA_PASS_AXPY_G), 6, "u is null");
    PLA_CHECK_ARG(!(flags & PLA_PASS_AXPY_G) || g != nullptr, 7, "AXPY_G needs g");
    PLA_CHECK_ARG(zss != nullptr, 11, "zss is null");
    PLA_CHECK_ARG(ws != nullptr && ws_bytes >= pla_stream_pass_workspace_bytes(m, n), 15, "workspace too small");
    cudaStream_t st = (cudaStream_t)stream;

    StreamPassParams p;
    p.A = A; p.m = m; p.n = n; p.lda = lda; p.w = w; p.u = u; p.g = g; p.sc = sc_dev; p.sa = sa; p.su = su;
    p.istop = istop_dev; p.flags = flags;
    const int vec = (n % 2 == 0) ? 2 : 1;
    const long long groups = (n + (long long)vec * SP_CONS - 1) / ((long long)vec * SP_CONS);
    PLA_CHECK_ARG(groups <= 16, 3, "n too large for this vector width (odd n must be <= 4096)");
    p.R = pick_tile_rows(n, sp_tile_bytes_target());
    const size_t stage_bytes = (size_t)p.R * n * 8;
    const size_t budget = 200 * 1024;
    int stages = (int)(budget / stage_bytes);
    if (stages > SP_MAX_STAGES) stages = SP_MAX_STAGES;
    PLA_CHECK_ARG(stages >= 2, 3, "row tile does not fit twice in shared memory");
    p.stages = stages;
    p.ntiles = (m + p.R - 1) / p.R;
    p.use_tma = (lda == n) && ((reinterpret_cast<uintptr_t>(A) & 15) == 0) && (((size_t)p.R * n) % 2 == 0);
    int grid = num_sms();
    if ((long long)grid > p.ntiles) grid = (int)p.ntiles;
    p.zpart = reinterpret_cast<double*>(ws);
    p.sspart = p.zpart + (size_t)num_sms() * n;
    const size_t smem = (((size_t)stages * stage_bytes + 15) & ~(size_t)15) + 2 * SP_MAX_STAGES * 8 +
                        2 * (SP_CONS / 32) * SP_RMAX * 8;

    cudaError_t e;
#define PLA_SP_CASE(V, JJ) e = launch_pass<V, JJ>(p, grid, smem, st)
    if (vec == 2) {
        if (groups <= 1) PLA_SP_CASE(2, 1);
        else if (groups <= 2) PLA_SP_CASE(2, 2);
        else if (groups <= 4) PLA_SP_CASE(2, 4);
        else if (groups <= 8) PLA_SP_CASE(2, 8);
        else PLA_SP_CASE(2, 16);
    } else {
        if (groups <= 1) PLA_SP_CASE(1, 1);
        else if (groups <= 2) PLA_SP_CASE(1, 2);
        else if (groups <= 4) PLA_SP_CASE(1, 4);
        else if (groups <= 8) PLA_SP_CASE(1, 8);
        else PLA_SP_CASE(1, 16);
    }
#undef PLA_SP_CASE
    if (e != cudaSuccess) { set_error("pla_stream_pass_f64: launch failed: %s", cudaGetErrorString(e)); return (int)e; }
    const int rb = (int)((n + 1 + 255) / 256);
    stream_pass_reduce_kernel<<<rb, 256, 0, st>>>(p.zpart, p.sspart, grid, n, zss, do_axpy ? 1 : 0, istop_dev);
    PLA_LAUNCH_CHECK();
    return 0;
}
