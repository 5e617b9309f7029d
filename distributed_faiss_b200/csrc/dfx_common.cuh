// dfx_common.cuh -- shared device/host helpers for libdfx (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <cuda_fp16.h>
#include <stdint.h>
#include <float.h>
#include <string>
#include <atomic>

#include "../../include/dfx.h"

// ---------------------------------------------------------------- errors
void dfx_set_error(const std::string& msg);
extern std::atomic<long long> g_dfx_launches;

struct DfxError {
    std::string msg;
};

#define DFX_CUDA(expr)                                                                       \
    do {                                                                                     \
        cudaError_t _e = (expr);                                                             \
        if (_e != cudaSuccess) {                                                             \
            throw DfxError{std::string(#expr) + " failed: " + cudaGetErrorString(_e) + " (" + \
                           __FILE__ + ":" + std::to_string(__LINE__) + ")"};                 \
        }                                                                                    \
    } while (0)

#define DFX_REQUIRE(cond, msg)                  \
    do {                                        \
        if (!(cond)) throw DfxError{(msg)};     \
    } while (0)

// every kernel launch goes through this so that dfx_launch_count() is honest
#ifdef DFX_EMU
// CPU emulator build (tests/emu/, never a product build): the kernel runs on the fiber SIMT
// runtime, one CTA after the other, synchronously
#define DFX_LAUNCH(kernel, grid, block, smem, stream, ...)                                      \
    do {                                                                                        \
        const dim3 _g(grid), _b(block);                                                         \
        simt::launch3(_g.x, _g.y, _b.x, (size_t)(smem), [&]() { kernel(__VA_ARGS__); });         \
        g_dfx_launches.fetch_add(1, std::memory_order_relaxed);                                 \
    } while (0)
#else
#define DFX_LAUNCH(kernel, grid, block, smem, stream, ...)                                      \
    do {                                                                                        \
        const dim3 _g(grid), _b(block);                                                         \
        kernel<<<_g, _b, (smem), (stream)>>>(__VA_ARGS__);                                      \
        g_dfx_launches.fetch_add(1, std::memory_order_relaxed);                                 \
        cudaError_t _le = cudaGetLastError();                                                   \
        if (_le != cudaSuccess)                                                                 \
            throw DfxError{std::string("launch of " #kernel " failed: ") + cudaGetErrorString(_le) + \
                           " grid=(" + std::to_string(_g.x) + "," + std::to_string(_g.y) + ") block=" + \
                           std::to_string(_b.x) + " smem=" + std::to_string((size_t)(smem)) + " (" +  \
                           __FILE__ + ":" + std::to_string(__LINE__) + ")"};                    \
    } while (0)
#endif

// grow-only device buffer
struct DevBuf {
    void* p = nullptr;
    size_t cap = 0;
    DevBuf() = default;
    DevBuf(const DevBuf&) = delete;
    DevBuf& operator=(const DevBuf&) = delete;
    DevBuf(DevBuf&& o) noexcept : p(o.p), cap(o.cap) {
        o.p = nullptr;
        o.cap = 0;
    }
    DevBuf& operator=(DevBuf&& o) noexcept {
        if (this != &o) {
            release();
            p = o.p;
            cap = o.cap;
            o.p = nullptr;
            o.cap = 0;
        }
        return *this;
    }
    ~DevBuf() { release(); }
    void release() {
        if (p) cudaFree(p);
        p = nullptr;
        cap = 0;
    }
    // ensure capacity; contents are NOT preserved unless keep_bytes > 0
    void reserve(size_t bytes, size_t keep_bytes = 0, cudaStream_t st = 0) {
        if (bytes <= cap) return;
        void* np_ = nullptr;
        DFX_CUDA(cudaMalloc(&np_, bytes));
        if (p && keep_bytes) {
            DFX_CUDA(cudaMemcpyAsync(np_, p, keep_bytes, cudaMemcpyDeviceToDevice, st));
            DFX_CUDA(cudaStreamSynchronize(st));
        }
        if (p) cudaFree(p);
        p = np_;
        cap = bytes;
    }
    template <typename T>
    T* as() const {
        return reinterpret_cast<T*>(p);
    }
};

// ---------------------------------------------------------------- ordering keys
// Everything on device minimises a float "value" (distance, or -inner_product) and breaks
// ties on a 32-bit secondary key (local id / column / position): composite = key<<32 | sec.
#define DFX_SEC_NONE 0xffffffffu
#define DFX_COMP_NONE 0xffffffffffffffffull

__host__ __device__ __forceinline__ uint32_t dfx_f2key(float v) {
    v = v + 0.0f;  // -0.0 -> +0.0 so that equal floats give equal keys
#ifdef __CUDA_ARCH__
    uint32_t u = __float_as_uint(v);
#else
    uint32_t u;
    memcpy(&u, &v, 4);
#endif
    return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}
__host__ __device__ __forceinline__ float dfx_key2f(uint32_t k) {
    uint32_t u = (k & 0x80000000u) ? (k & 0x7fffffffu) : ~k;
#ifdef __CUDA_ARCH__
    return __uint_as_float(u);
#else
    float f;
    memcpy(&f, &u, 4);
    return f;
#endif
}
__device__ __forceinline__ uint64_t dfx_comp(float v, uint32_t sec) {
    return ((uint64_t)dfx_f2key(v) << 32) | (uint64_t)sec;
}

// ---------------------------------------------------------------- canonical arithmetic (device)
// warp-dot: lane j owns k = 128*i + 4*j + t; butterfly xor 16,8,4,2,1
__device__ __forceinline__ float dfx_warp_butterfly(float acc) {
#pragma unroll
    for (int off = 16; off >= 1; off >>= 1) acc = acc + __shfl_xor_sync(0xffffffffu, acc, off);
    return acc;
}

static inline int64_t dfx_ceil_div(int64_t a, int64_t b) { return (a + b - 1) / b; }
static inline int dfx_next_pow2(int x) {
    int p = 1;
    while (p < x) p <<= 1;
    return p;
}
