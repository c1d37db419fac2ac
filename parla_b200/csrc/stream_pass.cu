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
#include <cstring>
#include "common.cuh"
#include "../../include/parla_b200.h"

namespace pla {

constexpr int SP_GROUP_MAX = 512;     // consumer threads per group: 256 (8 warps); 128 for narrow A; 512 for the widest rows
constexpr int SP_RMAX = 16;           // max rows per tile (16 for matrices of <= 256 columns: 32 KB tiles)
constexpr int SP_MAX_STAGES = 8;
constexpr int SP_MAX_GROUPS = 4;

// rows per tile as a function of the column-ownership shape: ~32 KB tiles, R*n even for odd n
template <int VEC, int J, int GS> struct SpRows {
    static constexpr int cols = VEC * J * GS;                 // widest matrix this shape serves
    static constexpr int raw = 4096 / cols;                   // rows of a ~32 KB tile
    static constexpr int capped = raw < 1 ? 1 : (raw > SP_RMAX ? SP_RMAX : raw);
    static constexpr int value = (VEC == 1 && capped == 1) ? 2 : capped;   // odd n needs an even row count
};

struct StreamPassParams {
    const double* A;
    long long m, n, lda;
    const double* w;
    double* u;
    const double* g;
    const double* sc;       // device {sa, su} or null
    double sa, su;
    double* zpart;          // [grid * groups][n]
    double* sspart;         // [grid * groups]
    const int* istop;
    int flags;
    int stages;
    int use_tag;            // 1: consumers check the stage's tile tag after the wait (ring length not a multiple of the groups)
    int use_tma;            // 1: 16B-aligned row tiles -- one bulk copy per tile (lda == n) or per row (lda > n)
    long long ntiles;
};

// NG independent consumer groups per CTA take alternate tiles of the CTA's sequence, so the
// dot -> reduce -> barrier -> axpy latency chain of one tile overlaps with the other group's.
template <int VEC, int J, int NG, int GS>
__global__ void __launch_bounds__(NG * GS + 32, 1) stream_pass_kernel(const StreamPassParams p) {
    if (p.istop != nullptr && *p.istop != 0) return;
    constexpr int SP_GROUP = GS;
    constexpr int RT = SpRows<VEC, J, GS>::value;
    constexpr int NCONS = NG * SP_GROUP;
    constexpr int NW = SP_GROUP / 32;
    extern __shared__ __align__(128) unsigned char sp_smem[];
    const int tid = threadIdx.x;
    const long long n = p.n;
    const int S = p.stages;
    const size_t stage_elems = (size_t)RT * n;
    double* tiles = reinterpret_cast<double*>(sp_smem);
    size_t off = ((size_t)S * stage_elems * 8 + 15) & ~(size_t)15;
    uint64_t* full = reinterpret_cast<uint64_t*>(sp_smem + off);
    uint64_t* empty = full + SP_MAX_STAGES;
    double* red = reinterpret_cast<double*>(empty + SP_MAX_STAGES);     // [groups][2][8 warps][RMAX]
    double* ustage = red + SP_MAX_GROUPS * 2 * (SP_GROUP_MAX / 32) * SP_RMAX;   // [stages][RMAX]  old u of the tile rows
    double* gstage = ustage + SP_MAX_STAGES * SP_RMAX;                  // [stages][RMAX]  g of the tile rows
    long long* tag = reinterpret_cast<long long*>(gstage + SP_MAX_STAGES * SP_RMAX);   // [stages] tile held / being loaded

    if (tid == 0) {
        for (int s = 0; s < S; ++s) { mbar_init(&full[s], 1 + 32); mbar_init(&empty[s], NW); }
        fence_mbar_init();
    }
    __syncthreads();

    const bool do_dot = (p.flags & PLA_PASS_DOT) != 0;
    const bool do_axpy = (p.flags & PLA_PASS_AXPY) != 0;
    const bool axpy_g = (p.flags & PLA_PASS_AXPY_G) != 0;

    if (tid >= NCONS) {
        // ------------------------------------------------------------- producer warp
        const int lane = tid - NCONS;
        const uint64_t pol = l2_policy_evict_first();
        int s = 0; uint32_t ph = 0;
        for (long long t = blockIdx.x; t < p.ntiles; t += gridDim.x) {
            const long long r0 = t * RT;
            const int rows = (int)min((long long)RT, p.m - r0);
            const size_t elems = (size_t)rows * n;
            if (lane == 0) mbar_wait(&empty[s], ph ^ 1);
            __syncwarp();
            // per-row scalars (old u, g) ride the same pipeline: their HBM latency is hidden by the ring
            if (lane < rows) {
                if (p.u != nullptr) cp_async8(ustage + s * SP_RMAX + lane, p.u + r0 + lane, true);
                if (axpy_g) cp_async8(gstage + s * SP_RMAX + lane, p.g + r0 + lane, true);
            }
            asm volatile("cp.async.mbarrier.arrive.noinc.shared::cta.b64 [%0];" ::"r"(smem_u32(&full[s])) : "memory");
            double* dst = tiles + (size_t)s * stage_elems;
            const bool tma_ok = p.use_tma && ((elems & 1) == 0);
            if (tma_ok) {
                if (p.lda == n) {
                    if (lane == 0) {
                        *reinterpret_cast<volatile long long*>(tag + s) = t;      // (released by the arrive below)
                        mbar_arrive_expect_tx(&full[s], (uint32_t)(elems * 8));
                        bulk_g2s(dst, p.A + r0 * p.lda, (uint32_t)(elems * 8), &full[s], pol);
                    }
                } else {
                    // a column block / padded matrix: rows are 16-byte aligned pieces, one bulk copy each
                    if (lane == 0) {
                        *reinterpret_cast<volatile long long*>(tag + s) = t;
                        mbar_arrive_expect_tx(&full[s], (uint32_t)(elems * 8));
                    }
                    __syncwarp();
                    if (lane < rows)
                        bulk_g2s(dst + (size_t)lane * n, p.A + (r0 + lane) * p.lda, (uint32_t)(n * 8), &full[s], pol);
                }
            } else {
                // generic path: strided / unaligned / odd-sized tail tile
                for (int r = 0; r < rows; ++r) {
                    const double* src = p.A + (r0 + r) * p.lda;
                    for (long long c = lane; c < n; c += 32) dst[(size_t)r * n + c] = __ldg(src + c);
                }
                __syncwarp();
                if (lane == 0) {
                    *reinterpret_cast<volatile long long*>(tag + s) = t;
                    mbar_arrive(&full[s]);
                }
            }
            if (++s == S) { s = 0; ph ^= 1; }
        }
        return;
    }

    // ----------------------------------------------------------------- consumer warps
    const int grp = tid / SP_GROUP, gt = tid % SP_GROUP;     // group, thread within group
    const int lane = gt & 31, wid = gt >> 5;
    double sa = p.sa, su = p.su;
    if (p.sc != nullptr) { sa = p.sc[0]; su = p.sc[1]; }

    // column ownership: VEC consecutive columns at c0(j) = VEC * (gt + SP_GROUP * j)
    double wreg[J * VEC], zacc[J * VEC];
#pragma unroll
    for (int j = 0; j < J; ++j)
#pragma unroll
        for (int v = 0; v < VEC; ++v) {
            const long long c = (long long)VEC * (gt + SP_GROUP * j) + v;
            wreg[j * VEC + v] = (do_dot && c < n) ? p.w[c] : 0.0;
            zacc[j * VEC + v] = 0.0;
        }
    double ss = 0.0;     // thread r (< RT) of each group accumulates u_r^2 over its tiles
    double* gred = red + (size_t)grp * 2 * NW * SP_RMAX;

    // this group consumes tiles q = grp, grp + NG, ... of the CTA's sequence; tile q lives in stage q % S
    int flip = 0;
    long long q = grp;
    for (long long t = blockIdx.x + (long long)grp * gridDim.x; t < p.ntiles; t += (long long)NG * gridDim.x, q += NG) {
        const int s = (int)(q % S);
        const uint32_t ph = (uint32_t)((q / S) & 1);
        const long long r0 = t * RT;
        const int rows = (int)min((long long)RT, p.m - r0);
        mbar_wait(&full[s], ph);
        // S % NG != 0: this group meets stage s only on every NG / gcd(S, NG)-th fill and always waits on the same
        // parity, so the completion of an OLDER fill can satisfy the wait (a parity wait must see every phase).  The
        // producer tags the stage with the tile it is loading before it arrives on `full`; once the tag is ours the
        // stage is in the phase we wait for and the wait is exact.  Until then: spin (at most one fill time).
        if (p.use_tag)
            while (*reinterpret_cast<volatile long long*>(tag + s) != t) mbar_wait(&full[s], ph);
        const double* tile = tiles + (size_t)s * stage_elems;
        const double* uo = ustage + s * SP_RMAX;           // old u / g of the tile rows (broadcast reads)
        const double* gq = gstage + s * SP_RMAX;
        const bool have_u = p.u != nullptr;

        double unew[RT];
        if (do_dot) {
            double dot[32];
#pragma unroll
            for (int r = 0; r < RT; ++r) dot[r] = 0.0;
#pragma unroll
            for (int j = 0; j < J; ++j) {
                const long long c = (long long)VEC * (gt + SP_GROUP * j);
                if (c < n) {
#pragma unroll
                    for (int r = 0; r < RT; ++r) {
                        if (r < rows) {
                            const double* row = tile + (size_t)r * n;
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
            double* myred = gred + (size_t)flip * NW * SP_RMAX;
            if (RT >= 4) {
                // many short rows: one multi-value butterfly (2 RT - 1 + ... shuffles instead of 5 RT); lane r ends
                // up with the warp total of row r
                warp_multi_sum<RT>(dot);
                if (lane < RT) myred[wid * SP_RMAX + lane] = dot[0];
            } else {
                // all rows reduced together: RT independent shuffle chains in flight
#pragma unroll
                for (int o = 16; o > 0; o >>= 1)
#pragma unroll
                    for (int r = 0; r < RT; ++r) dot[r] += __shfl_xor_sync(0xffffffffu, dot[r], o);
                if (lane == 0) {
#pragma unroll
                    for (int r = 0; r < RT; ++r) myred[wid * SP_RMAX + r] = dot[r];
                }
            }
            asm volatile("bar.sync %0, %1;" ::"r"(1 + grp), "n"(GS) : "memory");
#pragma unroll
            for (int r = 0; r < RT; ++r) {
                double tot = 0.0;
#pragma unroll
                for (int k = 0; k < NW; ++k) tot += myred[k * SP_RMAX + r];
                const double uold = (r < rows && have_u && su != 0.0) ? uo[r] : 0.0;
                unew[r] = (su != 0.0) ? fma(sa, tot, su * uold) : sa * tot;     // su == 0: u is write-only (may be uninitialised)
            }
            flip ^= 1;
            if (gt < rows) {
                // thread r owns row r of the tile
                double mine = 0.0;
#pragma unroll
                for (int r = 0; r < RT; ++r) if (r == gt) mine = unew[r];
                p.u[r0 + gt] = mine;
                ss = fma(mine, mine, ss);
            }
        } else {
#pragma unroll
            for (int r = 0; r < RT; ++r) unew[r] = (r < rows && have_u) ? uo[r] : 0.0;
            if (gt < rows) {
                double mine = 0.0;
#pragma unroll
                for (int r = 0; r < RT; ++r) if (r == gt) mine = unew[r];
                ss = fma(mine, mine, ss);
            }
        }

        if (do_axpy) {
#pragma unroll
            for (int j = 0; j < J; ++j) {
                const long long c = (long long)VEC * (gt + SP_GROUP * j);
                if (c < n) {
#pragma unroll
                    for (int r = 0; r < RT; ++r) {
                        if (r < rows) {
                            const double qv = axpy_g ? gq[r] : unew[r];     // (gq: broadcast read of the staged g)
                            const double* row = tile + (size_t)r * n;
                            if (VEC == 2) {
                                const double2 a = *reinterpret_cast<const double2*>(row + c);
                                zacc[j * VEC] = fma(a.x, qv, zacc[j * VEC]);
                                zacc[j * VEC + VEC - 1] = fma(a.y, qv, zacc[j * VEC + VEC - 1]);
                            } else {
                                zacc[j * VEC] = fma(row[c], qv, zacc[j * VEC]);
                            }
                        }
                    }
                }
            }
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(&empty[s]);
    }

    // ----------------------------------------------------------------- per-(CTA, group) partial results
    const size_t pidx = (size_t)blockIdx.x * NG + grp;
    if (do_axpy) {
        double* zp = p.zpart + pidx * n;
#pragma unroll
        for (int j = 0; j < J; ++j)
#pragma unroll
            for (int v = 0; v < VEC; ++v) {
                const long long c = (long long)VEC * (gt + SP_GROUP * j) + v;
                if (c < n) zp[c] = zacc[j * VEC + v];
            }
    }
    if (wid == 0) {
        const double tot = warp_sum(ss);     // only lanes < RT hold non-zero
        if (lane == 0) p.sspart[pidx] = tot;
    }
}

// zss[c] = sum_b zpart[b][c], zss[n] = sum_b sspart[b], both in a fixed order (bit-reproducible).
// A CTA owns 32 columns; its 8 warps each sum every 8th partial (8 independent loads in flight per thread), the 8
// warp results are added in warp order.  (One thread per column walking all ~600 partials serially cost 30-40 us,
// several times the streaming pass itself on a 2^16 x 500 matrix.)
constexpr int SPR_COLS = 32, SPR_WARPS = 8;

// ---- cross-GPU sum fused into the reduce: one-shot all-reduce over NVLink peer memory ---------------------------
// When A is row-sharded over the GPUs of a node, [z | |u|^2] must be summed over the ranks before the LSQR step
// (collective (2): n + 1 doubles per iteration, latency-critical).  Instead of a separate NCCL all-reduce,
// the thread that finishes column c PUSHES its total into slot [epoch parity][my rank][c] of every rank's exchange
// buffer (plain stores into peer memory mapped with CUDA IPC) and then polls the `world` slots of its OWN buffer,
// adding them in rank order -- every rank obtains the same bits.  A slot is a 16-byte line {lo, flag, hi, flag}
// (the NCCL "LL" layout: each 8-byte half carries the epoch, so a line is valid exactly when both flags match and no
// fence or barrier is needed); flag = epoch of the call, counted identically on every rank.  Two parities suffice:
// a rank can only reach call e + 2 after it has received every peer's data of call e + 1, which a peer sends after
// its kernel of call e has finished reading.  A poll gives up after 30 s (a missing peer) and returns NaN.
constexpr int SP_MAX_PEERS = 8;
struct PeerLine { uint32_t lo, f0, hi, f1; };
struct PeerExchange {
    PeerLine* recv[SP_MAX_PEERS];       // exchange buffer of every rank (recv[rank] is local memory)
    int rank, world;                    // world == 0: no exchange (world == 1 exchanges with itself: test hook)
    long long stride;                   // lines per (parity, source rank) slot, >= n + 1
    uint32_t epoch;
};
__device__ __forceinline__ void peer_store(PeerLine* line, double v, uint32_t flag) {
    const unsigned long long b = (unsigned long long)__double_as_longlong(v);
    asm volatile("st.relaxed.sys.global.v4.u32 [%0], {%1, %2, %3, %4};" ::"l"(line), "r"((uint32_t)b), "r"(flag),
                 "r"((uint32_t)(b >> 32)), "r"(flag)
                 : "memory");
}
__device__ __forceinline__ bool peer_try_load(const PeerLine* line, uint32_t flag, double& v) {
    uint32_t a, f0, c, f1;
    asm volatile("ld.relaxed.sys.global.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(a), "=r"(f0), "=r"(c), "=r"(f1) : "l"(line) : "memory");
    if (f0 != flag || f1 != flag) return false;
    v = __longlong_as_double((long long)(((unsigned long long)c << 32) | a));
    return true;
}
__device__ __forceinline__ unsigned long long global_timer_ns() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}

__global__ void __launch_bounds__(SPR_COLS * SPR_WARPS) stream_pass_reduce_kernel(const double* __restrict__ zpart,
                                                                                   const double* __restrict__ sspart,
                                                                                   int nblk, long long n,
                                                                                   double* __restrict__ zss, int do_axpy,
                                                                                   const int* istop, const PeerExchange px) {
    if (istop != nullptr && *istop != 0) return;
    __shared__ double acc_s[SPR_WARPS][SPR_COLS + 1];
    const int cl = threadIdx.x & 31, grp = threadIdx.x >> 5;
    const long long c = (long long)blockIdx.x * SPR_COLS + cl;      // column n is the |u|^2 slot
    const bool live = c < n ? do_axpy != 0 : c == n;
    double acc = 0.0;
    if (live) {
        const double* src = c < n ? zpart + c : sspart;
        const long long stride = c < n ? n : 1;
        for (int b0 = grp; b0 < nblk; b0 += SPR_WARPS * 8) {
            double t[8];
#pragma unroll
            for (int u = 0; u < 8; ++u) {
                const int b = b0 + SPR_WARPS * u;
                t[u] = b < nblk ? src[(size_t)b * stride] : 0.0;
            }
            acc += ((t[0] + t[1]) + (t[2] + t[3])) + ((t[4] + t[5]) + (t[6] + t[7]));
        }
    }
    acc_s[grp][cl] = acc;
    __syncthreads();
    if (grp == 0 && c <= n) {
        double tot = 0.0;
#pragma unroll
        for (int g2 = 0; g2 < SPR_WARPS; ++g2) tot += acc_s[g2][cl];
        if (px.world > 0 && live) {
            const size_t par = (size_t)(px.epoch & 1u) * px.world;
            const size_t mine = (par + px.rank) * px.stride + (size_t)c;
#pragma unroll
            for (int r = 0; r < SP_MAX_PEERS; ++r)
                if (r < px.world) peer_store(px.recv[r] + mine, tot, px.epoch);
            const PeerLine* in = px.recv[px.rank] + par * px.stride + (size_t)c;
            double v[SP_MAX_PEERS];
            unsigned pending = (1u << px.world) - 1u;
            const unsigned long long t0 = global_timer_ns();
            while (pending != 0) {
#pragma unroll
                for (int r = 0; r < SP_MAX_PEERS; ++r)
                    if (((pending >> r) & 1u) && peer_try_load(in + (size_t)r * px.stride, px.epoch, v[r])) pending &= ~(1u << r);
                if (pending != 0 && global_timer_ns() - t0 > 30000000000ull) {
#pragma unroll
                    for (int r = 0; r < SP_MAX_PEERS; ++r)
                        if ((pending >> r) & 1u) v[r] = __longlong_as_double(0x7ff8000000000000ll);
                    pending = 0;
                }
            }
            tot = 0.0;
#pragma unroll
            for (int r = 0; r < SP_MAX_PEERS; ++r)
                if (r < px.world) tot += v[r];
        }
        zss[c] = tot;
    }
}

template <int VEC, int J, int NG, int GS>
static cudaError_t launch_pass(StreamPassParams& p, cudaStream_t st, int* nparts) {
    constexpr int RT = SpRows<VEC, J, GS>::value;
    const size_t stage_bytes = (size_t)RT * p.n * 8;
    const size_t budget = 200 * 1024;
    int stages = (int)(budget / stage_bytes);
    if (stages > SP_MAX_STAGES) stages = SP_MAX_STAGES;
    // Tile q lives in stage q % S and belongs to group q % NG.  Unless S is a multiple of NG a group meets a stage
    // only on every second (or fourth) fill and always waits on the same parity, so it can take the completion of an
    // OLDER fill for its own, read a stale tile and arrive twice on the stage's `empty` barrier (A^T u alone at
    // 65536 x 500 -- 4 groups, 6 stages, fast consumers -- died with a launch failure).  Default: keep the deepest
    // ring and let the consumers verify the stage's tile tag (use_tag); PLA_PASS_TAGS=0: round S down instead.
    static const bool tags = [] { const char* e = getenv("PLA_PASS_TAGS"); return !(e && e[0] == '0'); }();
    if (!tags) stages = stages / NG * NG;
    if (stages < 2 || stages < NG) return cudaErrorInvalidValue;
    p.use_tag = (stages % NG != 0) ? 1 : 0;
    p.stages = stages;
    p.ntiles = (p.m + RT - 1) / RT;
    p.use_tma = ((reinterpret_cast<uintptr_t>(p.A) & 15) == 0) &&
                (p.lda == p.n ? (((size_t)RT * p.n) % 2 == 0) : (p.n % 2 == 0 && p.lda % 2 == 0));
    int grid = num_sms();
    if ((long long)grid > p.ntiles) grid = (int)p.ntiles;
    *nparts = grid * NG;
    const size_t smem = (((size_t)stages * stage_bytes + 15) & ~(size_t)15) + 2 * SP_MAX_STAGES * 8 +
                        SP_MAX_GROUPS * 2 * (SP_GROUP_MAX / 32) * SP_RMAX * 8 + 2 * SP_MAX_STAGES * SP_RMAX * 8 + SP_MAX_STAGES * 8;
    cudaError_t e = cudaFuncSetAttribute(stream_pass_kernel<VEC, J, NG, GS>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         (int)smem);
    if (e != cudaSuccess) return e;
    stream_pass_kernel<VEC, J, NG, GS><<<grid, NG * GS + 32, smem, st>>>(p);
    note_launch();
    return cudaGetLastError();
}

static int sp_groups() {
    static int v = 0;
    if (v == 0) {
        const char* e = getenv("PLA_PASS_GROUPS");
        v = e ? atoi(e) : 2;
        if (v < 1 || v > 2) v = 2;
    }
    return v;
}

}  // namespace pla

using namespace pla;

extern "C" size_t pla_stream_pass_workspace_bytes(int64_t m, int64_t n) {
    (void)m;
    return (size_t)num_sms() * SP_MAX_GROUPS * (size_t)(n + 1) * sizeof(double) + 256;
}

// parts_out != nullptr: the reduce kernel is NOT launched; the per-(CTA, group) partials stay in `ws` and
// parts_out = {number of partials, offset (in doubles) of the |u|^2 partials inside ws} (pla_stream_pass_parts_f64).
static int stream_pass_impl(const double* A, int64_t m, int64_t n, int64_t lda, const double* w, double* u,
                            const double* g, const double* sc_dev, double sa, double su, double* zss,
                            int flags, const int* istop_dev, void* ws, size_t ws_bytes, void* stream, const PeerExchange& px,
                            int64_t* parts_out = nullptr) {
    PLA_CHECK_ARG(A != nullptr, 1, "A is null");
    PLA_CHECK_ARG(m >= 1, 2, "m < 1");
    PLA_CHECK_ARG(n >= 1 && n <= PLA_PASS_MAX_N, 3, "n out of range for the streaming pass (1..8192)");
    PLA_CHECK_ARG(lda >= n, 4, "lda < n");
    const bool do_dot = flags & PLA_PASS_DOT, do_axpy = flags & PLA_PASS_AXPY;
    PLA_CHECK_ARG(!do_dot || (w != nullptr && u != nullptr), 5, "DOT needs w and u");
    PLA_CHECK_ARG(do_dot || u != nullptr || (flags & PLA_PASS_AXPY_G), 6, "u is null");
    PLA_CHECK_ARG(!(flags & PLA_PASS_AXPY_G) || g != nullptr, 7, "AXPY_G needs g");
    PLA_CHECK_ARG(zss != nullptr || parts_out != nullptr, 11, "zss is null");
    PLA_CHECK_ARG(ws != nullptr && ws_bytes >= pla_stream_pass_workspace_bytes(m, n), 15, "workspace too small");
    cudaStream_t st = (cudaStream_t)stream;

    StreamPassParams p;
    p.A = A; p.m = m; p.n = n; p.lda = lda; p.w = w; p.u = u; p.g = g; p.sc = sc_dev; p.sa = sa; p.su = su;
    p.istop = istop_dev; p.flags = flags;
    const int vec = (n % 2 == 0) ? 2 : 1;
    // Narrow matrices (<= 1024 columns when even, <= 512 when odd): four groups of 128 threads, so four
    // tiles are in flight per SM and the per-tile reduce/barrier latency stays hidden.  Wider: groups of 256.
    const bool narrow = n <= (long long)vec * 4 * 128 && sp_groups() == 2;
    const int gs = narrow ? 128 : 256;
    const long long groups = (n + (long long)vec * gs - 1) / ((long long)vec * gs);
    PLA_CHECK_ARG(groups <= 16, 3, "n too large for this vector width (odd n must be <= 4096)");
    p.zpart = reinterpret_cast<double*>(ws);
    p.sspart = p.zpart + (size_t)num_sms() * SP_MAX_GROUPS * n;
    const bool two = sp_groups() == 2;

    cudaError_t e;
    int nparts = 0;
#define PLA_SP_CASE(V, JJ, NGG, GSS) e = launch_pass<V, JJ, NGG, GSS>(p, st, &nparts)
    if (narrow) {
        if (vec == 2) {
            if (groups <= 1) PLA_SP_CASE(2, 1, 4, 128);
            else if (groups <= 2) PLA_SP_CASE(2, 2, 4, 128);
            else PLA_SP_CASE(2, 4, 4, 128);
        } else {
            if (groups <= 1) PLA_SP_CASE(1, 1, 4, 128);
            else if (groups <= 2) PLA_SP_CASE(1, 2, 4, 128);
            else PLA_SP_CASE(1, 4, 4, 128);
        }
    } else if (vec == 2) {
        if (groups <= 1) { if (two) PLA_SP_CASE(2, 1, 2, 256); else PLA_SP_CASE(2, 1, 1, 256); }
        else if (groups <= 2) { if (two) PLA_SP_CASE(2, 2, 2, 256); else PLA_SP_CASE(2, 2, 1, 256); }
        else if (groups <= 4) { if (two) PLA_SP_CASE(2, 4, 2, 256); else PLA_SP_CASE(2, 4, 1, 256); }
        else if (groups <= 8) PLA_SP_CASE(2, 8, 1, 256);
        else PLA_SP_CASE(2, 8, 1, 512);              // 4096 < n <= 8192: one group of 16 warps, 16 columns per thread
    } else {
        if (groups <= 1) { if (two) PLA_SP_CASE(1, 1, 2, 256); else PLA_SP_CASE(1, 1, 1, 256); }
        else if (groups <= 2) { if (two) PLA_SP_CASE(1, 2, 2, 256); else PLA_SP_CASE(1, 2, 1, 256); }
        else if (groups <= 4) { if (two) PLA_SP_CASE(1, 4, 2, 256); else PLA_SP_CASE(1, 4, 1, 256); }
        else if (groups <= 8) PLA_SP_CASE(1, 8, 1, 256);
        else PLA_SP_CASE(1, 8, 1, 512);
    }
#undef PLA_SP_CASE
    if (e != cudaSuccess) { set_error("pla_stream_pass_f64: launch failed: %s", cudaGetErrorString(e)); return (int)e; }
    if (parts_out != nullptr) {
        parts_out[0] = nparts;
        parts_out[1] = (int64_t)(p.sspart - p.zpart);
        return 0;
    }
    const int rb = (int)((n + 1 + SPR_COLS - 1) / SPR_COLS);
    stream_pass_reduce_kernel<<<rb, SPR_COLS * SPR_WARPS, 0, st>>>(p.zpart, p.sspart, nparts, n, zss, do_axpy ? 1 : 0,
                                                                  istop_dev, px);
    PLA_LAUNCH_CHECK();
    return 0;
}

extern "C" int pla_stream_pass_f64(const double* A, int64_t m, int64_t n, int64_t lda, const double* w, double* u,
                                   const double* g, const double* sc_dev, double sa, double su, double* zss,
                                   int flags, const int* istop_dev, void* ws, size_t ws_bytes, void* stream) {
    PeerExchange px;
    memset(&px, 0, sizeof(px));
    return stream_pass_impl(A, m, n, lda, w, u, g, sc_dev, sa, su, zss, flags, istop_dev, ws, ws_bytes, stream, px);
}

extern "C" int pla_stream_pass_parts_f64(const double* A, int64_t m, int64_t n, int64_t lda, const double* w, double* u,
                                         const double* g, const double* sc_dev, double sa, double su, int flags,
                                         const int* istop_dev, void* ws, size_t ws_bytes, int64_t* parts_out, void* stream) {
    PLA_CHECK_ARG(parts_out != nullptr, 15, "parts_out is null");
    PeerExchange px;
    memset(&px, 0, sizeof(px));
    return stream_pass_impl(A, m, n, lda, w, u, g, sc_dev, sa, su, nullptr, flags, istop_dev, ws, ws_bytes, stream, px,
                            parts_out);
}

extern "C" int pla_stream_pass_peer_f64(const double* A, int64_t m, int64_t n, int64_t lda, const double* w, double* u,
                                        const double* g, const double* sc_dev, double sa, double su, double* zss,
                                        int flags, const int* istop_dev, void* ws, size_t ws_bytes,
                                        void* const* peer_recv, int rank, int world, int64_t slot_lines, uint32_t epoch,
                                        void* stream) {
    PLA_CHECK_ARG(world >= 1 && world <= SP_MAX_PEERS && rank >= 0 && rank < world, 18, "need 0 <= rank < world <= 8");
    PLA_CHECK_ARG(peer_recv != nullptr && slot_lines >= n + 1 && epoch != 0, 17, "bad exchange buffers / slot size / epoch");
    PeerExchange px;
    memset(&px, 0, sizeof(px));
    for (int r = 0; r < world; ++r) {
        PLA_CHECK_ARG(peer_recv[r] != nullptr, 17, "null peer buffer");
        px.recv[r] = (PeerLine*)peer_recv[r];
    }
    px.rank = rank; px.world = world; px.stride = slot_lines; px.epoch = epoch;
    return stream_pass_impl(A, m, n, lda, w, u, g, sc_dev, sa, su, zss, flags, istop_dev, ws, ws_bytes, stream, px);
}

extern "C" size_t pla_peer_exchange_bytes(int world, int64_t slot_lines) {
    return (size_t)2 * (size_t)world * (size_t)slot_lines * sizeof(PeerLine);
}
