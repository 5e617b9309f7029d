// simt.h -- a tiny SIMT runtime for the CPU emulator: one CTA at a time, every CUDA thread is a
// fiber on ONE OS thread; fibers run until they reach a collective (warp exchange,
// ballot, __syncwarp, __syncthreads) or an explicit yield, where the next fiber is scheduled.
// Collectives are full-warp (all 32 lanes must arrive), which is how the scan kernels use them.
// Dynamic shared memory is filled with a poison pattern before each CTA; shared-memory reads by
// address are bounds-checked.  Test infrastructure only.
#pragma once
#include <stdint.h>
#include <string.h>
#include <stdio.h>
#include <stdlib.h>
#include <functional>
#include <mutex>
#include <vector>

// Context switch.  glibc's swapcontext saves/restores the signal mask with a system call on every
// switch, which dominates the run time of an emulated kernel (millions of switches); on x86-64 a
// dozen instructions that swap the callee-saved registers and the stack pointer do the same job.
#if defined(__x86_64__)
extern "C" void simt_switch(void** save_sp, void* load_sp);
#ifdef SIMT_IMPLEMENTATION  // exactly one translation unit of a binary defines it
asm(R"(
.text
.globl simt_switch
.type simt_switch,@function
simt_switch:
    pushq %rbp
    pushq %rbx
    pushq %r12
    pushq %r13
    pushq %r14
    pushq %r15
    movq %rsp, (%rdi)
    movq %rsi, %rsp
    popq %r15
    popq %r14
    popq %r13
    popq %r12
    popq %rbx
    popq %rbp
    ret
.size simt_switch,.-simt_switch
)");
#endif
#else
#error "tests/emu/simt.h: only x86-64 is supported (tests/test_emu_kernels.py skips elsewhere)"
#endif

#if defined(__SANITIZE_ADDRESS__)
#include <sanitizer/common_interface_defs.h>
#define SIMT_ASAN_START(save, bottom, size) __sanitizer_start_switch_fiber(save, bottom, size)
#define SIMT_ASAN_FINISH(save, bottom, size) __sanitizer_finish_switch_fiber(save, bottom, size)
#include <sanitizer/asan_interface.h>
#define SIMT_ASAN_POISON(p, n) __asan_poison_memory_region(p, n)
#define SIMT_ASAN_UNPOISON(p, n) __asan_unpoison_memory_region(p, n)
#else
#define SIMT_ASAN_POISON(p, n) ((void)0)
#define SIMT_ASAN_UNPOISON(p, n) ((void)0)
#define SIMT_ASAN_START(save, bottom, size) ((void)0)
#define SIMT_ASAN_FINISH(save, bottom, size) ((void)0)
#endif

namespace simt {

struct Tid { unsigned x, y, z; };
struct Fiber {
    void* sp;  // saved stack pointer while the fiber is not running
    Tid tid;
    bool done;
};
struct Warp {
    int arrive, gen;
    uint64_t slot[32];
};
struct State {
    std::vector<Fiber> fibers;
    std::vector<char> stacks;
    std::vector<Warp> warps;
    int nthreads = 0, cur = -1;
    int bar_arrive = 0, bar_gen = 0;
    void* sched_sp = nullptr;
    std::vector<unsigned char> dyn;
    std::function<void()> kernel;
    uint64_t rng = 0;  // != 0: fibers are resumed in a pseudo-random order
    long switches = 0;
    long events = 0;  // barrier completions + thread exits (deadlock watchdog)
};
inline State& S() {
    static State s;
    return s;
}
inline Tid g_block, g_bdim, g_gdim;

inline Tid& cur_tid() { return S().fibers[S().cur].tid; }
inline unsigned lane() { return (unsigned)S().cur & 31u; }
inline unsigned char* dyn_smem() { return S().dyn.data(); }
inline uint32_t smem_addr(const void* p) {
    const unsigned char* c = static_cast<const unsigned char*>(p);
    if (c < S().dyn.data() || c >= S().dyn.data() + S().dyn.size()) return 0xdead0000u;  // static __shared__: opaque
    return (uint32_t)(c - S().dyn.data());
}
inline unsigned char* smem_ptr(uint32_t addr) {
    if ((size_t)addr + 4 > S().dyn.size()) {
        fprintf(stderr, "simt: shared-memory access at %u outside the %zu dynamic bytes (thread %d)\n", addr,
                S().dyn.size(), S().cur);
        abort();
    }
    return S().dyn.data() + addr;
}
inline const void* g_sched_bottom = nullptr;  // the scheduler's stack, as AddressSanitizer sees it
inline size_t g_sched_size = 0;
inline void yield() {
    S().switches++;
    void* fake = nullptr;
    (void)fake;
    SIMT_ASAN_START(&fake, g_sched_bottom, g_sched_size);
    simt_switch(&S().fibers[S().cur].sp, S().sched_sp);
    SIMT_ASAN_FINISH(fake, &g_sched_bottom, &g_sched_size);
}
inline void warp_barrier() {
    Warp& w = S().warps[S().cur >> 5];
    const int g = w.gen;
    if (++w.arrive == 32) {
        w.arrive = 0;
        w.gen++;
        S().events++;
    } else {
        while (w.gen == g) yield();
    }
}
inline void cta_barrier() {
    State& s = S();
    const int g = s.bar_gen;
    if (++s.bar_arrive == s.nthreads) {
        s.bar_arrive = 0;
        s.bar_gen++;
        s.events++;
    } else {
        while (s.bar_gen == g) yield();
    }
}
// every lane deposits v; returns the value deposited by lane `src`
template <class T>
inline T warp_exchange(T v, unsigned src) {
    static_assert(sizeof(T) <= 8, "warp_exchange: at most 64 bits");
    Warp& w = S().warps[S().cur >> 5];
    uint64_t raw = 0;
    memcpy(&raw, &v, sizeof(T));
    w.slot[lane()] = raw;
    warp_barrier();
    const uint64_t got = w.slot[src];
    warp_barrier();  // everyone has read before the slots are reused
    T out;
    memcpy(&out, &got, sizeof(T));
    return out;
}
inline unsigned warp_ballot(bool pred) {
    Warp& w = S().warps[S().cur >> 5];
    w.slot[lane()] = pred ? 1 : 0;
    warp_barrier();
    unsigned m = 0;
    for (int i = 0; i < 32; i++) m |= (unsigned)(w.slot[i] & 1) << i;
    warp_barrier();
    return m;
}

inline void fiber_entry() {
    State& s = S();
    SIMT_ASAN_FINISH(nullptr, &g_sched_bottom, &g_sched_size);
    s.kernel();
    s.fibers[s.cur].done = true;
    SIMT_ASAN_START(nullptr, g_sched_bottom, g_sched_size);  // this fiber never resumes
    simt_switch(&s.fibers[s.cur].sp, s.sched_sp);
    abort();
}

// run `kernel` (a closure over the kernel arguments) on grid x block threads, dyn_bytes of
// dynamic shared memory per CTA.  CTAs run one after the other.
inline uint64_t g_default_seed = 0;  // scheduling order of launches that do not pass a seed
inline void launch3(unsigned gx, unsigned gy, unsigned block, size_t dyn_bytes, std::function<void()> kernel);
inline void launch(unsigned grid, unsigned block, size_t dyn_bytes, std::function<void()> kernel, uint64_t seed = 0) {
    const uint64_t keep = g_default_seed;
    g_default_seed = seed;
    launch3(grid, 1, block, dyn_bytes, kernel);
    g_default_seed = keep;
}
inline std::mutex& launch_mutex() {
    static std::mutex m;
    return m;
}
inline void launch3(unsigned gx, unsigned gy, unsigned block, size_t dyn_bytes, std::function<void()> kernel) {
    // one emulated device: launches from different host threads run one after the other
    std::lock_guard<std::mutex> lk(launch_mutex());
    State& s = S();
    const unsigned grid = gx * gy;
    const uint64_t seed = g_default_seed;
    if (block % 32 != 0 || block == 0 || block > 1024) {
        fprintf(stderr, "simt: block size %u must be a multiple of 32\n", block);
        abort();
    }
    constexpr size_t STACK = 512 * 1024;
    s.kernel = kernel;
    s.nthreads = (int)block;
    s.rng = seed;
    g_bdim = {block, 1, 1};
    g_gdim = {gx, gy, 1};
    s.stacks.resize((size_t)block * STACK);
    for (unsigned b = 0; b < grid; b++) {
        g_block = {b % gx, b / gx, 0};
        // filled with a pattern (reads of unwritten shared memory show up as garbage); the slack
        // behind the requested size is poisoned for AddressSanitizer, so that any access past the
        // end of the dynamic shared memory aborts (the buffer itself is reused: no reallocation)
        SIMT_ASAN_UNPOISON(s.dyn.data(), s.dyn.capacity());
        s.dyn.assign(dyn_bytes + 256, 0xCD);
        s.dyn.resize(dyn_bytes);
        SIMT_ASAN_POISON(s.dyn.data() + dyn_bytes, s.dyn.capacity() - dyn_bytes);
        s.fibers.assign(block, Fiber());
        s.warps.assign(block / 32, Warp());
        for (auto& w : s.warps) w.arrive = w.gen = 0;
        s.bar_arrive = 0;
        s.bar_gen = 0;
        for (unsigned t = 0; t < block; t++) {
            Fiber& f = s.fibers[t];
            f.tid = {t, 0, 0};
            f.done = false;
            // initial frame: six callee-saved registers, then fiber_entry as the return target;
            // at entry the stack pointer is 8 below a 16-byte boundary, as after a call
            uintptr_t top = reinterpret_cast<uintptr_t>(s.stacks.data() + (size_t)(t + 1) * STACK) & ~(uintptr_t)15;
            void** frame = reinterpret_cast<void**>(top - 64);
            for (int r = 0; r < 6; r++) frame[r] = nullptr;
            frame[6] = reinterpret_cast<void*>(&fiber_entry);
            frame[7] = nullptr;
            f.sp = frame;
        }
        int live = (int)block;
        long stale_sweeps = 0;
        while (live > 0) {
            const long before = s.events;
            for (unsigned i = 0; i < block; i++) {
                unsigned t = i;
                if (s.rng) {  // xorshift: visit threads in a scrambled order
                    s.rng ^= s.rng << 13; s.rng ^= s.rng >> 7; s.rng ^= s.rng << 17;
                    t = (unsigned)(s.rng % block);
                }
                if (s.fibers[t].done) continue;
                s.cur = (int)t;
                {
                    void* fake = nullptr;
                    (void)fake;
                    SIMT_ASAN_START(&fake, s.stacks.data() + (size_t)t * STACK, STACK);
                    simt_switch(&s.sched_sp, s.fibers[t].sp);
                    SIMT_ASAN_FINISH(fake, nullptr, nullptr);
                }
                if (s.fibers[t].done) {
                    live--;
                    s.events++;
                }
            }
            stale_sweeps = (s.events == before) ? stale_sweeps + 1 : 0;
            if (stale_sweeps > 100000) {
                fprintf(stderr, "simt: CTA %u makes no progress (a collective some threads never reach?)\n", b);
                abort();
            }
        }
        s.cur = -1;
        SIMT_ASAN_UNPOISON(s.dyn.data(), s.dyn.capacity());
    }
}

}  // namespace simt
