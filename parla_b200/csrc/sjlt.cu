// K2: SJLT sketch  out[d x n] = scale * S @ A  for a sparse sign operator with k nonzeros per column.
//
// Replaces scipy.sparse csc_matvecs behind `S @ A` / `S @ b` (parla/drivers/least_squares.py:303,314)
// for the operator built by parla/utils/sketching.py:51-74.
//
// Formulation: destination-major gather.  A one-off plan turns the column-wise index form into
// buckets (destination row, window of 2^16 source rows), each sorted by source row.  One warp owns
// (destination row, 256-column chunk): it walks its list, reads 2 KB contiguous pieces of the source
// rows (coalesced), accumulates in registers and adds into the output once per launch -- no atomics,
// no shared-memory read-modify-write, fixed summation order (deterministic).  The apply is issued as
// one launch per source window, so at any time all warps read the same ~1 GB slice of A: the k readers
// of a source-row piece then mostly hit it in the 126 MB L2 / TLB instead of re-reading HBM.
#include "common.cuh"
#include "philox.cuh"
#include "../../include/parla_b200.h"

namespace pla {

constexpr long long SJ_MAGIC = 0x504C41534A4C5431LL;   // "PLASJLT1"
constexpr int SJ_SORT_MAX = 8192;                      // per-destination segment sorted in smem up to this

// Plan layout: header | boff[d * nwin + 1] (int64) | entries[nnz] (int32, (src << 1) | negative).
// Bucket (r, w) = nonzeros of destination row r whose source row lies in window w (win_rows = 1 << win_shift
// consecutive rows of A); buckets are stored r-major, so destination r's full list is the concatenation of
// its windows in source order, and every bucket is sorted by source row (deterministic summation order).
struct SjltPlanHeader { long long magic, nnz, bad_index_count, nwin, win_shift, pad0, pad1, pad2; };

static int sjlt_win_shift(long long d, long long k) {
    // windows of <= 2^16 source rows (~1 GB of A at n = 2048: bounded TLB / L2 footprint per launch;
    // measured on B200: 72 ms un-windowed -> 49 ms at 2^22 x 2048, flat below 2^16) and
    // <= ~4096 expected entries per bucket (sorted in shared memory)
    long long lim = 4096 * d / (k > 0 ? k : 1);
    static int max_shift = 0;
    if (max_shift == 0) {
        const char* e = getenv("PLA_SJLT_WIN_SHIFT");
        max_shift = e ? atoi(e) : 16;
        if (max_shift < 10 || max_shift > 24) max_shift = 16;
    }
    int sh = max_shift;
    while (sh > 10 && (1LL << sh) > lim) --sh;
    return sh;
}

__global__ void __launch_bounds__(256) sjlt_count_kernel(const int32_t* __restrict__ rows, long long nnz, long long k,
                                                         long long d, long long nwin, int win_shift,
                                                         unsigned long long* counts, SjltPlanHeader* hdr) {
    for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < nnz; e += (long long)gridDim.x * blockDim.x) {
        const long long r = rows[e];
        if (r < 0 || r >= d) atomicAdd((unsigned long long*)&hdr->bad_index_count, 1ULL);
        else atomicAdd(&counts[r * nwin + ((e / k) >> win_shift)], 1ULL);
    }
}

// exclusive scan of counts[d] -> offsets[d+1] (single CTA, chunked)
__global__ void __launch_bounds__(1024) sjlt_scan_kernel(const unsigned long long* __restrict__ counts, long long d,
                                                         long long* offsets) {
    __shared__ long long wsum[32];
    __shared__ long long carry;
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    if (threadIdx.x == 0) carry = 0;
    __syncthreads();
    for (long long base = 0; base < d; base += 1024) {
        const long long i = base + threadIdx.x;
        const long long v = i < d ? (long long)counts[i] : 0;
        long long incl = v;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const long long t = __shfl_up_sync(0xffffffffu, incl, o);
            if (lane >= o) incl += t;
        }
        if (lane == 31) wsum[wid] = incl;
        __syncthreads();
        if (wid == 0) {
            long long w = wsum[lane];
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const long long t = __shfl_up_sync(0xffffffffu, w, o);
                if (lane >= o) w += t;
            }
            wsum[lane] = w;     // inclusive over warps
        }
        __syncthreads();
        const long long before = carry + (wid > 0 ? wsum[wid - 1] : 0) + incl - v;
        if (i < d) offsets[i] = before;
        __syncthreads();
        if (threadIdx.x == 1023) carry = before + v;
        __syncthreads();
    }
    if (threadIdx.x == 0) offsets[d] = carry;
}

__global__ void __launch_bounds__(256) sjlt_fill_kernel(const int32_t* __restrict__ rows, const int8_t* __restrict__ signs,
                                                        long long nnz, long long k, long long d, long long nwin,
                                                        int win_shift, const long long* __restrict__ boff,
                                                        unsigned long long* cursor, int32_t* entries) {
    for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < nnz; e += (long long)gridDim.x * blockDim.x) {
        const long long r = rows[e];
        if (r < 0 || r >= d) continue;
        const long long src = e / k;
        const long long bkt = r * nwin + (src >> win_shift);
        const unsigned long long slot = atomicAdd(&cursor[bkt], 1ULL);
        entries[boff[bkt] + (long long)slot] = (int32_t)((src << 1) | (signs[e] < 0 ? 1 : 0));
    }
}

// sort each bucket by packed entry (= by source row): bitonic sort in shared memory; a bucket longer than the
// shared-memory buffer (only a contrived operator can produce one: a bucket holds at most one entry per source row
// of its window) is sorted in place in global memory by the same network in its all-ascending form, where positions
// past the end act as +infinity and their comparisons are skipped -- never left in atomic (run-dependent) order.
__global__ void __launch_bounds__(256) sjlt_segsort_kernel(const long long* __restrict__ offsets, int32_t* entries) {
    extern __shared__ int32_t seg[];
    const long long beg = offsets[blockIdx.x], end = offsets[blockIdx.x + 1];
    const long long len64 = end - beg;
    if (len64 <= 1) return;
    if (len64 > SJ_SORT_MAX) {
        int32_t* a = entries + beg;
        long long np2 = 1;
        while (np2 < len64) np2 <<= 1;
        for (long long size = 2; size <= np2; size <<= 1) {
            const long long half = size >> 1;
            for (long long i = threadIdx.x; i < np2 / 2; i += blockDim.x) {          // flip stage
                const long long blk = i / half, off = i - blk * half;
                const long long lo = blk * size + off, hi = blk * size + size - 1 - off;
                if (hi < len64) {
                    const int32_t x = a[lo], y = a[hi];
                    if (x > y) { a[lo] = y; a[hi] = x; }
                }
            }
            __syncthreads();
            for (long long stride = half >> 1; stride > 0; stride >>= 1) {           // half-cleaners
                for (long long i = threadIdx.x; i < np2 / 2; i += blockDim.x) {
                    const long long lo = 2 * i - (i & (stride - 1)), hi = lo + stride;
                    if (hi < len64) {
                        const int32_t x = a[lo], y = a[hi];
                        if (x > y) { a[lo] = y; a[hi] = x; }
                    }
                }
                __syncthreads();
            }
        }
        return;
    }
    const int len = (int)len64;
    int np2 = 1;
    while (np2 < len) np2 <<= 1;
    for (int i = threadIdx.x; i < np2; i += blockDim.x) seg[i] = i < len ? entries[beg + i] : 0x7fffffff;
    __syncthreads();
    for (int size = 2; size <= np2; size <<= 1) {
        for (int stride = size >> 1; stride > 0; stride >>= 1) {
            for (int i = threadIdx.x; i < np2 / 2; i += blockDim.x) {
                const int lo = 2 * i - (i & (stride - 1));
                const int hi = lo + stride;
                const bool up = (lo & size) == 0;
                const int32_t a = seg[lo], b = seg[hi];
                if ((a > b) == up) { seg[lo] = b; seg[hi] = a; }
            }
            __syncthreads();
        }
    }
    for (int i = threadIdx.x; i < len; i += blockDim.x) entries[beg + i] = seg[i];
}

// ---- apply: warp <-> (DPW consecutive destination rows, column chunk of 32*VEC*J columns)
// All destination rows of a chunk should be co-resident on the GPU: their (source-sorted) lists are
// then swept in step, so the k readers of a source-row piece hit it in L2 and A is read from HBM about
// once.  DPW > 1 (several lists interleaved per warp, batch by batch) and narrower chunks trade
// accumulator registers for that co-residency when d is large.
template <int VEC, int J, int DPW>
__global__ void __launch_bounds__(256, 4) sjlt_apply_kernel(const long long* __restrict__ offsets,
                                                            const int32_t* __restrict__ entries, long long d,
                                                            const double* __restrict__ A, long long n, long long lda,
                                                            const double* __restrict__ bvec, double scale, double* out,
                                                            long long ldo, double* out_b, long long ldob,
                                                            int accumulate, long long nchunks, int splits,
                                                            double* part, long long nwin, long long w_lo,
                                                            long long w_hi) {
    // splits > 1: every destination list is cut into `splits` segments handled by different warps
    // (more parallelism when d * n is small); segment sums go to part[seg][d][n+1] and are added in
    // segment order by sjlt_reduce_kernel.
    constexpr int CW = 32 * VEC * J;                    // chunk width in columns
    const int lane = threadIdx.x & 31;
    const long long dgroups = (d + DPW - 1) / DPW;
    const long long task = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (task >= nchunks * dgroups * splits) return;
    const int seg = (int)(task % splits);
    const long long cr = task / splits;
    const long long chunk = cr / dgroups, r0 = (cr - chunk * dgroups) * DPW;
    const long long c0 = chunk * CW;
    if (splits > 1) {
        out = part + (size_t)seg * d * (n + 1);
        ldo = n + 1;
        out_b = (bvec != nullptr) ? out + n : nullptr;
        ldob = n + 1;
        accumulate = 0;
        scale = 1.0;
    }
    long long pos[DPW], end[DPW];
    double acc[DPW][J * VEC], bacc[DPW];
#pragma unroll
    for (int p = 0; p < DPW; ++p) {
        const long long r = r0 + p;
        long long b0 = 0, e0 = 0;
        if (r < d) {
            b0 = offsets[r * nwin + w_lo];          // buckets (r, w_lo) .. (r, w_hi - 1) are contiguous
            e0 = offsets[r * nwin + w_hi];
            if (splits > 1) {
                const long long len = e0 - b0;
                const long long b2 = b0 + len * seg / splits, e2 = b0 + len * (seg + 1) / splits;
                b0 = b2; e0 = e2;
            }
        }
        pos[p] = b0; end[p] = e0; bacc[p] = 0.0;
#pragma unroll
        for (int j = 0; j < J * VEC; ++j) acc[p][j] = 0.0;
    }
    const bool do_b = (bvec != nullptr) && (chunk == 0);

    bool any = true;
    while (any) {
        any = false;
#pragma unroll
        for (int p = 0; p < DPW; ++p) {
            if (pos[p] >= end[p]) continue;            // warp-uniform
            const long long me = pos[p] + lane;
            const int32_t ent = me < end[p] ? entries[me] : 0;
            if (do_b && me < end[p]) {
                const double bv = bvec[ent >> 1];
                bacc[p] += (ent & 1) ? -bv : bv;
            }
            const int cnt = (int)min(32LL, end[p] - pos[p]);
#pragma unroll 2
            for (int t = 0; t < cnt; ++t) {
                const int32_t en = __shfl_sync(0xffffffffu, ent, t);
                const double* src = A + (long long)(en >> 1) * lda + c0;
                const double sg = (en & 1) ? -1.0 : 1.0;
#pragma unroll
                for (int j = 0; j < J; ++j) {
                    const long long c = (long long)VEC * (lane + 32 * j);
                    if (c0 + c < n) {
                        if (VEC == 2) {
                            const double2 a = __ldg(reinterpret_cast<const double2*>(src + c));
                            acc[p][j * VEC] = fma(sg, a.x, acc[p][j * VEC]);
                            acc[p][j * VEC + VEC - 1] = fma(sg, a.y, acc[p][j * VEC + VEC - 1]);
                        } else {
                            acc[p][j * VEC] = fma(sg, __ldg(src + c), acc[p][j * VEC]);
                        }
                    }
                }
            }
            pos[p] += 32;
            any = any || (pos[p] < end[p]);
        }
    }
#pragma unroll
    for (int p = 0; p < DPW; ++p) {
        const long long r = r0 + p;
        if (r >= d) continue;
        double* dst = out + r * ldo + c0;
#pragma unroll
        for (int j = 0; j < J; ++j)
#pragma unroll
            for (int v = 0; v < VEC; ++v) {
                const long long c = (long long)VEC * (lane + 32 * j) + v;
                if (c0 + c < n) {
                    const double val = scale * acc[p][j * VEC + v];
                    dst[c] = accumulate ? dst[c] + val : val;
                }
            }
        if (do_b && out_b != nullptr) {
            const double tot = warp_sum(bacc[p]);
            if (lane == 0) out_b[r * ldob] = accumulate ? out_b[r * ldob] + scale * tot : scale * tot;
        }
    }
}

__global__ void __launch_bounds__(256) sjlt_reduce_kernel(const double* __restrict__ part, int splits, long long d,
                                                          long long n, int with_b, double scale, double* out,
                                                          long long ldo, double* out_b, long long ldob, int accumulate) {
    const long long total = d * (n + 1);
    for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < total;
         idx += (long long)gridDim.x * blockDim.x) {
        const long long r = idx / (n + 1), c = idx - r * (n + 1);
        if (c == n && !with_b) continue;
        double acc = 0.0;
        for (int sgm = 0; sgm < splits; ++sgm) acc += part[(size_t)sgm * total + idx];
        double* dst = (c == n) ? out_b + r * ldob : out + r * ldo + c;
        *dst = accumulate ? *dst + scale * acc : scale * acc;
    }
}

// ---- native operator: k distinct rows per column from Philox (see oracle/philox_ref.py: sjlt_columns)
__global__ void __launch_bounds__(256) sjlt_generate_kernel(long long d, long long m, int k, uint64_t seed,
                                                            long long col_offset, int32_t* rows, int8_t* signs) {
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < m; i += (long long)gridDim.x * blockDim.x) {
        const uint64_t gi = (uint64_t)(col_offset + i);
        uint32_t call = 0;
        Philox4 o = philox4x32_10((uint32_t)gi, (uint32_t)(gi >> 32), call, 0x534A4C54u, (uint32_t)seed,
                                  (uint32_t)(seed >> 32));
        uint32_t words[4] = {o.x, o.y, o.z, o.w};
        int wi = 1;                                   // words[0] = sign bits
        const uint32_t sbits = words[0];
        int32_t picked[32];
        for (int q = 0; q < k; ++q) {
            int32_t cand;
            bool dup;
            do {
                if (wi == 4) {
                    ++call;
                    o = philox4x32_10((uint32_t)gi, (uint32_t)(gi >> 32), call, 0x534A4C54u, (uint32_t)seed,
                                      (uint32_t)(seed >> 32));
                    words[0] = o.x; words[1] = o.y; words[2] = o.z; words[3] = o.w;
                    wi = 0;
                }
                cand = (int32_t)(((uint64_t)words[wi++] * (uint64_t)d) >> 32);
                dup = false;
                if (d >= k)
                    for (int t = 0; t < q; ++t) dup |= (picked[t] == cand);
            } while (dup);
            picked[q] = cand;
            rows[i * k + q] = cand;
            signs[i * k + q] = ((sbits >> q) & 1u) ? (int8_t)-1 : (int8_t)1;
        }
    }
}

}  // namespace pla

using namespace pla;

static size_t align256(size_t v) { return (v + 255) & ~(size_t)255; }

static long long sjlt_nwin(long long d, long long m, long long k) {
    const int sh = sjlt_win_shift(d, k);
    return (m + (1LL << sh) - 1) >> sh;
}

extern "C" size_t pla_sjlt_plan_bytes(int64_t d, int64_t m, int64_t k) {
    return align256(sizeof(SjltPlanHeader)) + align256((size_t)(d * sjlt_nwin(d, m, k) + 1) * 8) +
           align256((size_t)m * k * 4);
}
extern "C" size_t pla_sjlt_plan_workspace_bytes(int64_t d, int64_t m, int64_t k) {
    return 2 * align256((size_t)d * sjlt_nwin(d, m, k) * 8);
}

extern "C" int pla_sjlt_plan_f64(const int32_t* rows, const int8_t* signs, int64_t m, int64_t k, int64_t d,
                                 void* plan, void* ws, size_t ws_bytes, void* stream) {
    PLA_CHECK_ARG(rows != nullptr && signs != nullptr, 1, "null index arrays");
    PLA_CHECK_ARG(m >= 1 && m < (1LL << 30), 3, "m out of range (1 .. 2^30-1)");
    PLA_CHECK_ARG(k >= 1 && k <= 32, 4, "k out of range (1..32)");
    PLA_CHECK_ARG(d >= 1, 5, "d < 1");
    PLA_CHECK_ARG(plan != nullptr, 6, "plan is null");
    PLA_CHECK_ARG(ws != nullptr && ws_bytes >= pla_sjlt_plan_workspace_bytes(d, m, k), 8, "workspace too small");
    cudaStream_t st = (cudaStream_t)stream;
    const int win_shift = sjlt_win_shift(d, k);
    const long long nwin = sjlt_nwin(d, m, k), nb_total = d * nwin;
    char* pb = (char*)plan;
    SjltPlanHeader* hdr = (SjltPlanHeader*)pb;
    long long* boff = (long long*)(pb + align256(sizeof(SjltPlanHeader)));
    int32_t* entries = (int32_t*)(pb + align256(sizeof(SjltPlanHeader)) + align256((size_t)(nb_total + 1) * 8));
    unsigned long long* counts = (unsigned long long*)ws;
    unsigned long long* cursor = (unsigned long long*)((char*)ws + align256((size_t)nb_total * 8));
    const long long nnz = m * k;
    SjltPlanHeader h{SJ_MAGIC, nnz, 0, nwin, win_shift, 0, 0, 0};
    PLA_CUDA(cudaMemcpyAsync(hdr, &h, sizeof(h), cudaMemcpyHostToDevice, st));
    PLA_CUDA(cudaMemsetAsync(ws, 0, 2 * align256((size_t)nb_total * 8), st));
    int nb = (int)((nnz + 255) / 256);
    if (nb > 16 * num_sms()) nb = 16 * num_sms();
    sjlt_count_kernel<<<nb, 256, 0, st>>>(rows, nnz, k, d, nwin, win_shift, counts, hdr);
    PLA_LAUNCH_CHECK();
    sjlt_scan_kernel<<<1, 1024, 0, st>>>(counts, nb_total, boff);
    PLA_LAUNCH_CHECK();
    sjlt_fill_kernel<<<nb, 256, 0, st>>>(rows, signs, nnz, k, d, nwin, win_shift, boff, cursor, entries);
    PLA_LAUNCH_CHECK();
    PLA_CUDA(cudaFuncSetAttribute(sjlt_segsort_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, SJ_SORT_MAX * 4));
    sjlt_segsort_kernel<<<(unsigned)nb_total, 256, SJ_SORT_MAX * 4, st>>>(boff, entries);
    PLA_LAUNCH_CHECK();
    return 0;
}

// Launch shape: chunk width (32*VEC*J columns) and destinations per warp (DPW) chosen so that one
// chunk's d / DPW warps fit among the ~32 resident warps per SM; few warps overall -> split the lists.
struct SjltShape { int J, DPW; long long cw, nchunks, dgroups; int splits; };

static SjltShape sjlt_shape(long long d, long long n, int vec, size_t ws_bytes, bool have_ws) {
    const long long resident = (long long)num_sms() * 32;
    SjltShape sh;
    sh.J = 4; sh.DPW = 1;                       // measured best on B200; (J=2,DPW=2) / (J=1,DPW=4) via env
    if (const char* e = getenv("PLA_SJLT_DPW")) {
        const int v = atoi(e);
        if (v == 2) { sh.J = 2; sh.DPW = 2; } else if (v == 4) { sh.J = 1; sh.DPW = 4; }
    }
    (void)resident;
    sh.cw = 32LL * vec * sh.J;
    sh.nchunks = (n + sh.cw - 1) / sh.cw;
    sh.dgroups = (d + sh.DPW - 1) / sh.DPW;
    const long long warps = sh.dgroups * sh.nchunks, want = resident;
    long long sp = (want + warps - 1) / warps;
    if (sp > 16) sp = 16;
    const long long fit = have_ws ? (long long)(ws_bytes / ((size_t)d * (size_t)(n + 1) * 8)) : 0;
    if (sp > fit) sp = fit;
    sh.splits = sp < 2 ? 1 : (int)sp;
    return sh;
}

extern "C" size_t pla_sjlt_apply_workspace_bytes(int64_t d, int64_t n) {
    const SjltShape sh = sjlt_shape(d, n, (n % 2 == 0) ? 2 : 1, ~(size_t)0, true);
    return sh.splits < 2 ? 0 : (size_t)sh.splits * d * (n + 1) * 8;
}

extern "C" int pla_sjlt_apply_f64(const void* plan, int64_t d, int64_t m, int64_t k, const double* A, int64_t n,
                                  int64_t lda, const double* bvec, double scale, double* out, int64_t ldo,
                                  double* out_b, int64_t ldob, int accumulate, void* ws, size_t ws_bytes,
                                  void* stream) {
    PLA_CHECK_ARG(plan != nullptr, 1, "plan is null");
    PLA_CHECK_ARG(d >= 1 && m >= 1 && k >= 1, 2, "bad dims");
    PLA_CHECK_ARG(A != nullptr && n >= 1 && lda >= n, 5, "bad A / n / lda");
    PLA_CHECK_ARG(out != nullptr && ldo >= n, 10, "bad out / ldo");
    PLA_CHECK_ARG(bvec == nullptr || (out_b != nullptr && ldob >= 1), 12, "out_b is null / ldob < 1");
    const long long nwin = sjlt_nwin(d, m, k), nb_total = d * nwin;
    const char* pb = (const char*)plan;
    const long long* offsets = (const long long*)(pb + align256(sizeof(SjltPlanHeader)));
    const int32_t* entries = (const int32_t*)(pb + align256(sizeof(SjltPlanHeader)) + align256((size_t)(nb_total + 1) * 8));
    cudaStream_t st = (cudaStream_t)stream;
    const bool vec2 = (n % 2 == 0) && (lda % 2 == 0) && ((reinterpret_cast<uintptr_t>(A) & 15) == 0);
    const SjltShape sh = sjlt_shape(d, n, vec2 ? 2 : 1, ws_bytes, ws != nullptr);
    const int warps_per_cta = 8;
    const long long ctas = (sh.nchunks * sh.dgroups * sh.splits + warps_per_cta - 1) / warps_per_cta;
    double* part = sh.splits > 1 ? (double*)ws : nullptr;
    // One launch per group of source windows: every launch only touches a few GB of A, and the output
    // is accumulated across launches.  wstep windows per launch keeps >= ~64 list entries per warp.
    long long wstep = 1;
    {
        const int win_shift = sjlt_win_shift(d, k);
        const double per_win = (double)(1LL << win_shift) * (double)k / (double)d / (double)sh.splits;
        while (wstep < nwin && per_win * (double)wstep < 64.0) wstep *= 2;
        const char* e = getenv("PLA_SJLT_WINDOWS_PER_LAUNCH");
        if (e && atoll(e) > 0) wstep = atoll(e);
    }
    for (long long w_lo = 0; w_lo < nwin; w_lo += wstep) {
        const long long w_hi = (w_lo + wstep < nwin) ? w_lo + wstep : nwin;
        const int acc_flag = (w_lo == 0) ? accumulate : 1;
#define PLA_SJ_LAUNCH(V, JJ, DD)                                                                                     \
    sjlt_apply_kernel<V, JJ, DD><<<(unsigned)ctas, 256, 0, st>>>(offsets, entries, d, A, n, lda, bvec, scale, out, ldo, \
                                                                 out_b, ldob, acc_flag, sh.nchunks, sh.splits, part,  \
                                                                 nwin, w_lo, w_hi)
        if (vec2) {
            if (sh.DPW == 1) PLA_SJ_LAUNCH(2, 4, 1);
            else if (sh.DPW == 2) PLA_SJ_LAUNCH(2, 2, 2);
            else PLA_SJ_LAUNCH(2, 1, 4);
        } else {
            if (sh.DPW == 1) PLA_SJ_LAUNCH(1, 4, 1);
            else if (sh.DPW == 2) PLA_SJ_LAUNCH(1, 2, 2);
            else PLA_SJ_LAUNCH(1, 1, 4);
        }
#undef PLA_SJ_LAUNCH
        PLA_LAUNCH_CHECK();
        if (sh.splits > 1) {
            long long total = d * (n + 1);
            int nb = (int)((total + 255) / 256);
            if (nb > 8 * num_sms()) nb = 8 * num_sms();
            sjlt_reduce_kernel<<<nb, 256, 0, st>>>(part, sh.splits, d, n, bvec != nullptr ? 1 : 0, scale, out, ldo,
                                                   out_b, ldob, acc_flag);
            PLA_LAUNCH_CHECK();
        }
    }
    return 0;
}

extern "C" int pla_sjlt_plan_status(const void* plan, int64_t* bad_index_count_host) {
    PLA_CHECK_ARG(plan != nullptr && bad_index_count_host != nullptr, 1, "null argument");
    SjltPlanHeader h;
    PLA_CUDA(cudaMemcpy(&h, plan, sizeof(h), cudaMemcpyDeviceToHost));   // synchronising: debug / validation only
    if (h.magic != SJ_MAGIC) { set_error("pla_sjlt_plan_status: not a plan"); return -1; }
    *bad_index_count_host = h.bad_index_count;
    return 0;
}

extern "C" int pla_sjlt_generate(int64_t d, int64_t m, int64_t k, uint64_t seed, int64_t col_offset, int32_t* rows,
                                 int8_t* signs, void* stream) {
    PLA_CHECK_ARG(d >= 1 && d < (1LL << 31), 1, "d out of range");
    PLA_CHECK_ARG(m >= 1, 2, "m < 1");
    PLA_CHECK_ARG(k >= 1 && k <= 32, 3, "k out of range (1..32)");
    PLA_CHECK_ARG(col_offset >= 0, 5, "col_offset < 0");
    PLA_CHECK_ARG(rows != nullptr && signs != nullptr, 6, "null outputs");
    int nb = (int)((m + 255) / 256);
    if (nb > 16 * num_sms()) nb = 16 * num_sms();
    sjlt_generate_kernel<<<nb, 256, 0, (cudaStream_t)stream>>>(d, m, (int)k, seed, col_offset, rows, signs);
    PLA_LAUNCH_CHECK();
    return 0;
}
