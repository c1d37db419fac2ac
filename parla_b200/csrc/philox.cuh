// Counter-based Philox4x32-10 (Salmon et al., SC'11) and the virtual Gaussian / SJLT operators
// built on it.  The CPU statement of exactly these definitions is oracle/philox_ref.py.
#pragma once
#include <stdint.h>

namespace pla {

struct Philox4 { uint32_t x, y, z, w; };

__host__ __device__ __forceinline__ Philox4 philox4x32_10(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3,
                                                          uint32_t k0, uint32_t k1) {
#pragma unroll
    for (int r = 0; r < 10; ++r) {
        const uint64_t p0 = (uint64_t)0xD2511F53u * c0;
        const uint64_t p1 = (uint64_t)0xCD9E8D57u * c2;
        const uint32_t n0 = (uint32_t)(p1 >> 32) ^ c1 ^ k0;
        const uint32_t n2 = (uint32_t)(p0 >> 32) ^ c3 ^ k1;
        c1 = (uint32_t)p1;
        c3 = (uint32_t)p0;
        c0 = n0;
        c2 = n2;
        k0 += 0x9E3779B9u;
        k1 += 0xBB67AE85u;
    }
    return Philox4{c0, c1, c2, c3};
}

#ifdef __CUDACC__
// Box-Muller on the four words of one Philox block (second half of philox_normal4, split out so the
// integer rounds and the transcendental part can be issued at different points of a DMMA loop).
__device__ __forceinline__ void philox_words_to_normal4(const Philox4& o, double out[4]) {
    const float s = 2.3283064365386963e-10f;   // 2^-32
    const float ua = fmaf((float)o.x, s, 0.5f * s);
    const float ub = fmaf((float)o.y, s, 0.5f * s);
    const float uc = fmaf((float)o.z, s, 0.5f * s);
    const float ud = fmaf((float)o.w, s, 0.5f * s);
    const float ra = sqrtf(-2.0f * __logf(fminf(ua, 1.0f)));
    const float rc = sqrtf(-2.0f * __logf(fminf(uc, 1.0f)));
    float sa, ca, sc, cc;
    __sincosf(3.14159265358979f * (2.0f * ub - 1.0f), &sa, &ca);
    __sincosf(3.14159265358979f * (2.0f * ud - 1.0f), &sc, &cc);
    out[0] = (double)(ra * ca);
    out[1] = (double)(ra * sa);
    out[2] = (double)(rc * cc);
    out[3] = (double)(rc * sc);
}

// Four standard normals for operator row r, column block q (columns 4q .. 4q+3).
// u = (o + 0.5) 2^-32 in (0,1); Box-Muller with fp32 fast intrinsics, widened to fp64.
__device__ __forceinline__ void philox_normal4(uint64_t seed, uint32_t r, uint64_t q, double out[4]) {
    const Philox4 o = philox4x32_10((uint32_t)q, (uint32_t)(q >> 32), r, 0u, (uint32_t)seed, (uint32_t)(seed >> 32));
    const float s = 2.3283064365386963e-10f;   // 2^-32
    const float ua = fmaf((float)o.x, s, 0.5f * s);
    const float ub = fmaf((float)o.y, s, 0.5f * s);
    const float uc = fmaf((float)o.z, s, 0.5f * s);
    const float ud = fmaf((float)o.w, s, 0.5f * s);
    // (float)o can round up to 2^32 -> u == 1 + tiny: clamp so the log stays <= 0
    const float ra = sqrtf(-2.0f * __logf(fminf(ua, 1.0f)));
    const float rc = sqrtf(-2.0f * __logf(fminf(uc, 1.0f)));
    float sa, ca, sc, cc;
    __sincosf(3.14159265358979f * (2.0f * ub - 1.0f), &sa, &ca);
    __sincosf(3.14159265358979f * (2.0f * ud - 1.0f), &sc, &cc);
    out[0] = (double)(ra * ca);
    out[1] = (double)(ra * sa);
    out[2] = (double)(rc * cc);
    out[3] = (double)(rc * sc);
}
#endif

}  // namespace pla
