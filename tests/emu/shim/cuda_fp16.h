// stand-in for <cuda_fp16.h> (CPU emulator build only): storage type + exact conversions
#pragma once
#include <cuda_runtime.h>
struct __half { unsigned short x; };
struct __half2 { __half x, y; };
inline float __half2float(__half h) {
    const uint32_t s = (uint32_t)(h.x & 0x8000u) << 16, e = (h.x >> 10) & 0x1f, m = h.x & 0x3ffu;
    uint32_t u;
    if (e == 0) {
        if (m == 0) u = s;
        else {  // subnormal: normalise
            int sh = 0;
            uint32_t mm = m;
            while (!(mm & 0x400u)) { mm <<= 1; sh++; }
            u = s | ((uint32_t)(127 - 15 - sh + 1) << 23) | ((mm & 0x3ffu) << 13);
        }
    } else if (e == 31) u = s | 0x7f800000u | (m << 13);
    else u = s | ((e + 112u) << 23) | (m << 13);
    float f;
    memcpy(&f, &u, 4);
    return f;
}
// round to nearest even (IEEE), like cvt.rn.f16.f32
inline __half __float2half_rn(float f) {
    uint32_t u;
    memcpy(&u, &f, 4);
    const uint32_t s = (u >> 16) & 0x8000u;
    const uint32_t a = u & 0x7fffffffu;
    __half h;
    if (a >= 0x7f800000u) { h.x = (unsigned short)(s | 0x7c00u | ((a > 0x7f800000u) ? 0x200u | ((a >> 13) & 0x3ffu) : 0)); return h; }
    if (a >= 0x477ff000u) { h.x = (unsigned short)(s | 0x7c00u); return h; }  // rounds to inf (>= 65520)
    if (a < 0x33000001u) { h.x = (unsigned short)s; return h; }               // rounds to zero (<= 2^-25)
    int e = (int)(a >> 23) - 127;
    uint32_t m = (a & 0x7fffffu) | 0x800000u;
    int shift = (e < -14) ? (13 + (-14 - e)) : 13;  // bits dropped from the 24-bit significand
    uint32_t q = m >> shift, rem = m & ((1u << shift) - 1u), halfway = 1u << (shift - 1);
    if (rem > halfway || (rem == halfway && (q & 1u))) q++;
    uint32_t out;
    if (e < -14) out = q;  // subnormal (q may carry into the normal range: still correct)
    else out = ((uint32_t)(e + 15) << 10) + (q - 0x400u);  // q in [0x400, 0x800]: a carry bumps the exponent
    h.x = (unsigned short)(s | out);
    return h;
}
inline __half __float2half(float f) { return __float2half_rn(f); }
inline float2 __half22float2(__half2 h) { return float2{__half2float(h.x), __half2float(h.y)}; }
