// Minimal stand-in for <cuda_runtime.h> used ONLY by the CPU lane-level emulator (tests/emu/):
// lets g++ compile the device headers of distributed_faiss_b200/csrc unchanged.  Every CUDA
// thread of a CTA runs as a fiber (simt.h); warp collectives and __syncthreads are barriers
// between fibers, so full-mask warp-synchronous code behaves as on hardware.  Test infrastructure.
#pragma once
// system headers first: the qualifier macros below must not leak into them
#include <stdint.h>
#include <stddef.h>
#include <stdlib.h>
#include <string.h>
#include <math.h>
#include <float.h>
#include <assert.h>
#include <string>
#include <atomic>
#include <vector>
#include <mutex>
#include <memory>
#include <algorithm>
#include <functional>
#include "../simt.h"

#define __global__
#define __device__
#define __host__
#define __forceinline__ inline __attribute__((always_inline))
#define __noinline__ __attribute__((noinline))
#define __launch_bounds__(...)
#define __align__(x) __attribute__((aligned(x)))
#define __shared__ static

#define threadIdx (simt::cur_tid())
#define blockIdx (simt::g_block)
#define blockDim (simt::g_bdim)
#define gridDim (simt::g_gdim)

struct alignas(16) uint4 { unsigned x, y, z, w; };
struct alignas(16) float4 { float x, y, z, w; };
struct dim3 {
    unsigned x, y, z;
    dim3(unsigned x_ = 1, unsigned y_ = 1, unsigned z_ = 1) : x(x_), y(y_), z(z_) {}
};

// ---- runtime API stubs (only what the inline host helpers of dfx_common.cuh mention)
typedef int cudaError_t;
enum { cudaSuccess = 0 };
typedef void* cudaStream_t;
typedef void* cudaEvent_t;
enum cudaMemcpyKind { cudaMemcpyHostToHost, cudaMemcpyHostToDevice, cudaMemcpyDeviceToHost, cudaMemcpyDeviceToDevice };
inline const char* cudaGetErrorString(cudaError_t) { return "emu"; }
inline cudaError_t cudaMalloc(void** p, size_t n) { *p = malloc(n); return 0; }
inline cudaError_t cudaFree(void* p) { free(p); return 0; }
inline cudaError_t cudaMemcpyAsync(void* d, const void* s, size_t n, cudaMemcpyKind, cudaStream_t) { memcpy(d, s, n); return 0; }
inline cudaError_t cudaStreamSynchronize(cudaStream_t) { return 0; }
inline cudaError_t cudaGetLastError() { return 0; }

// ---- intrinsics
using std::min;
using std::max;
inline int __popc(unsigned x) { return __builtin_popcount(x); }
inline float __int_as_float(int x) { float f; memcpy(&f, &x, 4); return f; }
inline int __float_as_int(float f) { int x; memcpy(&x, &f, 4); return x; }
inline float __uint_as_float(unsigned x) { float f; memcpy(&f, &x, 4); return f; }
inline unsigned __float_as_uint(float f) { unsigned x; memcpy(&x, &f, 4); return x; }
inline float __fmaf_rn(float a, float b, float c) { return fmaf(a, b, c); }
template <class T> inline T __ldg(const T* p) { return *p; }
inline unsigned __byte_perm(unsigned x, unsigned y, unsigned s) {
    const uint64_t pool = ((uint64_t)y << 32) | x;
    unsigned r = 0;
    for (int i = 0; i < 4; i++) {
        const unsigned sel = (s >> (4 * i)) & 0xf;
        unsigned b = (unsigned)(pool >> (8 * (sel & 7))) & 0xff;
        if (sel & 8) b = (b & 0x80) ? 0xff : 0x00;  // sign-replicate mode
        r |= b << (8 * i);
    }
    return r;
}
inline unsigned atomicMin(unsigned* p, unsigned v) { unsigned o = *p; if (v < o) *p = v; return o; }
inline unsigned long long atomicMin(unsigned long long* p, unsigned long long v) { unsigned long long o = *p; if (v < o) *p = v; return o; }
inline unsigned atomicAdd(unsigned* p, unsigned v) { unsigned o = *p; *p = o + v; return o; }
inline int atomicAdd(int* p, int v) { int o = *p; *p = o + v; return o; }

inline void __syncthreads() { simt::cta_barrier(); }
inline void __syncwarp(unsigned = 0xffffffffu) { simt::warp_barrier(); }
template <class T> inline T __shfl_sync(unsigned mask, T v, int src, int = 32) {
    assert(mask == 0xffffffffu);
    return simt::warp_exchange(v, (unsigned)src & 31u);
}
template <class T> inline T __shfl_xor_sync(unsigned mask, T v, int lanemask, int = 32) {
    assert(mask == 0xffffffffu);
    return simt::warp_exchange(v, (simt::lane() ^ (unsigned)lanemask) & 31u);
}
inline unsigned __ballot_sync(unsigned mask, int pred) {
    assert(mask == 0xffffffffu);
    return simt::warp_ballot(pred != 0);
}
inline int __any_sync(unsigned mask, int pred) { return __ballot_sync(mask, pred) != 0; }
inline int __all_sync(unsigned mask, int pred) { return __ballot_sync(mask, pred) == 0xffffffffu; }
