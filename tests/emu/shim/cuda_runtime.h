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
struct alignas(8) uint2 { unsigned x, y; };
struct alignas(8) float2 { float x, y; };
inline float4 make_float4(float x, float y, float z, float w) { return float4{x, y, z, w}; }
inline float2 make_float2(float x, float y) { return float2{x, y}; }
inline uint4 make_uint4(unsigned x, unsigned y, unsigned z, unsigned w) { return uint4{x, y, z, w}; }
inline uint2 make_uint2(unsigned x, unsigned y) { return uint2{x, y}; }
struct dim3 {
    unsigned x, y, z;
    dim3(unsigned x_ = 1, unsigned y_ = 1, unsigned z_ = 1) : x(x_), y(y_), z(z_) {}
};

// ---- runtime API: "device memory" is host memory, streams and events are inert, every launch is
// synchronous (simt::launch3)
typedef int cudaError_t;
enum { cudaSuccess = 0 };
typedef void* cudaStream_t;
typedef void* cudaEvent_t;
enum cudaMemcpyKind { cudaMemcpyHostToHost, cudaMemcpyHostToDevice, cudaMemcpyDeviceToHost, cudaMemcpyDeviceToDevice, cudaMemcpyDefault };
enum cudaFuncAttribute { cudaFuncAttributeMaxDynamicSharedMemorySize = 8, cudaFuncAttributePreferredSharedMemoryCarveout = 9 };
enum { cudaStreamNonBlocking = 1 };
struct cudaDeviceProp {
    char name[256];
    int major, minor, multiProcessorCount;
    size_t totalGlobalMem, sharedMemPerBlockOptin;
};
inline const char* cudaGetErrorString(cudaError_t) { return "emulated"; }
inline cudaError_t cudaMalloc(void** p, size_t n) { *p = malloc(n ? n : 1); return *p ? 0 : 2; }
template <class T> inline cudaError_t cudaMalloc(T** p, size_t n) { return cudaMalloc(reinterpret_cast<void**>(p), n); }
inline cudaError_t cudaFree(void* p) { free(p); return 0; }
inline cudaError_t cudaMemcpyAsync(void* d, const void* s, size_t n, cudaMemcpyKind, cudaStream_t = nullptr) { if (n) memmove(d, s, n); return 0; }
inline cudaError_t cudaMemcpy(void* d, const void* s, size_t n, cudaMemcpyKind) { if (n) memmove(d, s, n); return 0; }
inline cudaError_t cudaMemsetAsync(void* d, int v, size_t n, cudaStream_t = nullptr) { if (n) memset(d, v, n); return 0; }
inline cudaError_t cudaMemset(void* d, int v, size_t n) { if (n) memset(d, v, n); return 0; }
inline cudaError_t cudaStreamSynchronize(cudaStream_t) { return 0; }
inline cudaError_t cudaDeviceSynchronize() { return 0; }
inline cudaError_t cudaStreamCreateWithFlags(cudaStream_t* s, unsigned) { *s = nullptr; return 0; }
inline cudaError_t cudaStreamDestroy(cudaStream_t) { return 0; }
inline cudaError_t cudaGetLastError() { return 0; }
inline cudaError_t cudaGetDeviceCount(int* n) { *n = 1; return 0; }
inline cudaError_t cudaGetDevice(int* d) { *d = 0; return 0; }
inline cudaError_t cudaSetDevice(int) { return 0; }
inline cudaError_t cudaGetDeviceProperties(cudaDeviceProp* p, int) {
    memset(p, 0, sizeof(*p));
    strcpy(p->name, "fiber SIMT emulator");
    p->major = 10; p->minor = 0; p->multiProcessorCount = 148;
    p->totalGlobalMem = (size_t)180 << 30; p->sharedMemPerBlockOptin = 227 * 1024;
    return 0;
}
template <class F> inline cudaError_t cudaFuncSetAttribute(F, cudaFuncAttribute, int) { return 0; }
inline cudaError_t cudaMallocHost(void** p, size_t n) { *p = malloc(n ? n : 1); return *p ? 0 : 2; }
inline cudaError_t cudaFreeHost(void* p) { free(p); return 0; }
inline cudaError_t cudaEventQuery(cudaEvent_t) { return 0; }
inline cudaError_t cudaEventCreateWithFlags(cudaEvent_t* e, unsigned) { *e = nullptr; return 0; }
enum cudaStreamCaptureStatus { cudaStreamCaptureStatusNone = 0, cudaStreamCaptureStatusActive = 1 };
inline cudaError_t cudaStreamIsCapturing(cudaStream_t, cudaStreamCaptureStatus* st) { *st = cudaStreamCaptureStatusNone; return 0; }
#define cudaEventDisableTiming 2u
inline cudaError_t cudaEventCreate(cudaEvent_t* e) { *e = nullptr; return 0; }
inline cudaError_t cudaEventDestroy(cudaEvent_t) { return 0; }
inline cudaError_t cudaEventRecord(cudaEvent_t, cudaStream_t = nullptr) { return 0; }
inline cudaError_t cudaEventSynchronize(cudaEvent_t) { return 0; }
inline cudaError_t cudaEventElapsedTime(float* ms, cudaEvent_t, cudaEvent_t) { *ms = 0.f; return 0; }

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
inline float atomicAdd(float* p, float v) { float o = *p; *p = o + v; return o; }
inline unsigned long long atomicAdd(unsigned long long* p, unsigned long long v) { unsigned long long o = *p; *p = o + v; return o; }
inline int atomicMax(int* p, int v) { int o = *p; if (v > o) *p = v; return o; }
inline unsigned atomicMax(unsigned* p, unsigned v) { unsigned o = *p; if (v > o) *p = v; return o; }
inline int atomicMin(int* p, int v) { int o = *p; if (v < o) *p = v; return o; }
inline unsigned atomicCAS(unsigned* p, unsigned cmp, unsigned v) { unsigned o = *p; if (o == cmp) *p = v; return o; }
inline unsigned long long atomicCAS(unsigned long long* p, unsigned long long cmp, unsigned long long v) { unsigned long long o = *p; if (o == cmp) *p = v; return o; }
inline unsigned atomicExch(unsigned* p, unsigned v) { unsigned o = *p; *p = v; return o; }
inline int __clz(int x) { return x ? __builtin_clz((unsigned)x) : 32; }
inline int __ffs(int x) { return __builtin_ffs(x); }
inline float __fadd_rn(float a, float b) { return a + b; }
inline float __fmul_rn(float a, float b) { return a * b; }
inline float __fsub_rn(float a, float b) { return a - b; }
inline float __fdividef(float a, float b) { return a / b; }
inline float rsqrtf(float x) { return 1.0f / sqrtf(x); }
inline unsigned __activemask() { return 0xffffffffu; }
// fast-math intrinsics (only the synthetic data generator uses them: values need not match the GPU's)
#define __logf(x) logf(x)
#define __cosf(x) cosf(x)
#define __sinf(x) sinf(x)
#define __expf(x) expf(x)
inline void __threadfence() {}
inline void __threadfence_block() {}
inline unsigned long long __umul64hi(unsigned long long a, unsigned long long b) { return (unsigned long long)(((unsigned __int128)a * b) >> 64); }
inline unsigned __umulhi(unsigned a, unsigned b) { return (unsigned)(((unsigned long long)a * b) >> 32); }

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
