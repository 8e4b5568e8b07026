/*
 * cuda_emu.h  --  a tiny single-process CUDA emulator.  TEST INFRASTRUCTURE ONLY.
 *
 * The build container has nvcc but no GPU, so a kernel bug would otherwise be found only on
 * a (budgeted) B200 box.  tests/emu/build_emu.py compiles py-debruijn_b200/csrc/dbg_b200.cu
 * with g++ and this header (-DDBG_EMU) into tests/emu/_build/libdbg_b200_emu.so; the CPU
 * tests (-m "not gpu") then run the host logic AND the kernel bodies, serially, against the
 * oracle.  The product never loads that library: py-debruijn_b200/_native.py only ever
 * opens the nvcc-built libdbg_b200.so and refuses to work without a CUDA device.
 *
 * Model: blocks run one after another; the threads of a block are ucontext fibers scheduled
 * round-robin, switching only at __syncthreads / warp collectives, so atomics are trivially
 * atomic and "device memory" is the host heap.  No tensor cores, TMA, clusters or streams
 * (everything is synchronous).  Timing APIs return wall-clock time.
 *
 * Chaos mode (DBG_EMU_CHAOS=seed[:permille], read at every launch): a fiber ALSO gives way, with that probability,
 * in front of every atomic, every __ldcg / __ldcv / __stcg and the 128-bit CAS, and each scheduler pass starts at a
 * random thread.  The threads of a block then interleave at exactly the points where the lock-free protocols of the
 * product (CAS on the first key half, plain store of the second, retry rings, min / add on a slot another lane is
 * still publishing) can observe each other half-way, which the serial schedule never shows.  Single operations
 * stay atomic.  tests/test_emu_chaos.py runs the parity cases under several seeds.
 */
#pragma once
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>
#include <ucontext.h>

#include <functional>
#include <vector>

#define __global__
#define __device__
#define __host__
#define __forceinline__ inline
#define __restrict__
#define __launch_bounds__(...)
#define __align__(n) __attribute__((aligned(n)))
#define __shared__ static

struct uint3 { unsigned x, y, z; };
struct dim3 {
    unsigned x, y, z;
    dim3(unsigned a = 1, unsigned b = 1, unsigned c = 1) : x(a), y(b), z(c) {}
};
struct ulonglong2 { unsigned long long x, y; } __attribute__((aligned(16)));
struct uint4 { unsigned x, y, z, w; } __attribute__((aligned(16)));
struct uint2 { unsigned x, y; } __attribute__((aligned(8)));
static inline uint4 make_uint4(unsigned a, unsigned b, unsigned c, unsigned d) { uint4 r = {a, b, c, d}; return r; }
static inline ulonglong2 make_ulonglong2(unsigned long long a, unsigned long long b) { ulonglong2 r = {a, b}; return r; }

namespace emu {
struct Warp {
    uint32_t arrived = 0;
    uint32_t gen = 0;
    uint64_t val[32];
    uint64_t snap[2][32];
    uint32_t snap_mask[2];
};
/* Fiber switch.  glibc's swapcontext makes a sigprocmask system call on every switch (a third of the CPU suite's time);
 * on x86-64 a fiber only has to keep the callee-saved registers and its stack pointer, so the switch is ten
 * instructions.  Other targets (or -DEMU_UCONTEXT) keep ucontext. */
#if defined(__x86_64__) && !defined(EMU_UCONTEXT)
#define EMU_FAST_SWITCH 1
extern "C" void emu_switch(void **save_sp, void *load_sp);
__asm__(".text\n"
        ".globl emu_switch\n"
        ".type emu_switch,@function\n"
        "emu_switch:\n"
        "    pushq %rbp\n    pushq %rbx\n    pushq %r12\n    pushq %r13\n    pushq %r14\n    pushq %r15\n"
        "    movq %rsp, (%rdi)\n"
        "    movq %rsi, %rsp\n"
        "    popq %r15\n    popq %r14\n    popq %r13\n    popq %r12\n    popq %rbx\n    popq %rbp\n"
        "    ret\n"
        ".size emu_switch,.-emu_switch\n");
#endif
struct Fiber {
#ifdef EMU_FAST_SWITCH
    void *sp = nullptr;
#else
    ucontext_t ctx;
#endif
    bool done = false;
    bool at_barrier = false;
};
struct State {
    uint3 threadIdx, blockIdx;
    dim3 blockDim, gridDim;
#ifdef EMU_FAST_SWITCH
    void *sched_sp = nullptr;
#else
    ucontext_t sched;
#endif
    std::vector<Fiber> fibers;
    std::vector<Warp> warps;
    std::vector<char> stacks;
    int cur = -1;
    int n_threads = 0;
    const std::function<void()> *body = nullptr;
    char *dyn_smem = nullptr;
    uint64_t chaos_rng = 0;                         /* 0 = chaos mode off */
    uint32_t chaos_permille = 0;
    uint64_t chaos_yields = 0;                      /* how often chaos mode switched fibers (emu_chaos_yields()) */
};
inline State &S() { static State s; return s; }
static const size_t STACK = 256 * 1024;

inline void yield() {
    State &s = S();
    int me = s.cur;
#ifdef EMU_FAST_SWITCH
    emu_switch(&s.fibers[me].sp, s.sched_sp);
#else
    swapcontext(&s.fibers[me].ctx, &s.sched);
#endif
}
/* chaos mode: give way here with probability chaos_permille / 1000 (xorshift64*, seeded per launch) */
inline void chaos_point() {
    State &s = S();
    if (!s.chaos_rng || s.cur < 0) return;
    uint64_t x = s.chaos_rng;
    x ^= x >> 12; x ^= x << 25; x ^= x >> 27;
    s.chaos_rng = x;
    if ((uint32_t)((x * 0x2545F4914F6CDD1DULL) >> 40) % 1000u < s.chaos_permille) { s.chaos_yields++; yield(); }
}
inline void chaos_setup() {
    State &s = S();
    static uint64_t launches = 0;
    const char *e = getenv("DBG_EMU_CHAOS");
    s.chaos_rng = 0; s.chaos_permille = 0;
    if (!e || !*e) return;
    char *end = nullptr;
    unsigned long long seed = strtoull(e, &end, 10);
    s.chaos_permille = (end && *end == ':') ? (uint32_t)strtoul(end + 1, nullptr, 10) : 250u;
    s.chaos_rng = (seed * 0x9E3779B97F4A7C15ULL + (++launches) * 0xD1B54A32D192ED03ULL) | 1ULL;
}
inline void trampoline() {
    State &s = S();
    (*s.body)();
    s.fibers[s.cur].done = true;
#ifdef EMU_FAST_SWITCH
    emu_switch(&s.fibers[s.cur].sp, s.sched_sp);
    abort();                                        /* a finished fiber is never resumed */
#else
    swapcontext(&s.fibers[s.cur].ctx, &s.sched);
#endif
}
inline void launch(dim3 grid, dim3 block, size_t smem, const std::function<void()> &body) {
    State &s = S();
    int nt = (int)(block.x * block.y * block.z);
    s.n_threads = nt;
    s.blockDim = block; s.gridDim = grid; s.body = &body;
    chaos_setup();
    if (s.stacks.size() < (size_t)nt * STACK) s.stacks.resize((size_t)nt * STACK);
    std::vector<char> dyn(smem ? smem : 1);
    s.dyn_smem = dyn.data();
    s.fibers.resize(nt);
    s.warps.assign((nt + 31) / 32, Warp());
    for (unsigned bz = 0; bz < grid.z; bz++)
    for (unsigned by = 0; by < grid.y; by++)
    for (unsigned bx = 0; bx < grid.x; bx++) {
        s.blockIdx = {bx, by, bz};
        for (int t = 0; t < nt; t++) {
            Fiber &f = s.fibers[t];
            f.done = false; f.at_barrier = false;
#ifdef EMU_FAST_SWITCH
            {   /* a stack that emu_switch "returns" from into trampoline(): 6 saved registers, the entry address, and
                   padding so that the stack pointer is 8 mod 16 at entry, as after a call */
                uintptr_t top = ((uintptr_t)(s.stacks.data() + (size_t)(t + 1) * STACK)) & ~(uintptr_t)15;
                void **sp = (void **)(top - 8);
                *--sp = (void *)(void (*)())trampoline;
                for (int i = 0; i < 6; i++) *--sp = nullptr;
                f.sp = (void *)sp;
            }
#else
            getcontext(&f.ctx);
            f.ctx.uc_stack.ss_sp = s.stacks.data() + (size_t)t * STACK;
            f.ctx.uc_stack.ss_size = STACK;
            f.ctx.uc_link = &s.sched;
            makecontext(&f.ctx, (void (*)())trampoline, 0);
#endif
        }
        for (auto &w : s.warps) { w.arrived = 0; w.gen = 0; }
        int remaining = nt;
        while (remaining > 0) {
            int progressed = 0, waiting = 0;
            int t0 = 0;
            if (s.chaos_rng) { s.chaos_rng ^= s.chaos_rng << 13; s.chaos_rng ^= s.chaos_rng >> 7; s.chaos_rng ^= s.chaos_rng << 17; t0 = (int)((s.chaos_rng >> 20) % (uint64_t)nt); }
            for (int tt = 0; tt < nt; tt++) {
                const int t = tt + t0 < nt ? tt + t0 : tt + t0 - nt;
                Fiber &f = s.fibers[t];
                if (f.done) continue;
                if (f.at_barrier) { waiting++; continue; }
                s.cur = t;
                s.threadIdx = { (unsigned)(t % block.x), (unsigned)((t / block.x) % block.y),
                                (unsigned)(t / (block.x * block.y)) };
#ifdef EMU_FAST_SWITCH
                emu_switch(&s.sched_sp, f.sp);
#else
                swapcontext(&s.sched, &f.ctx);
#endif
                progressed++;
                if (f.done) remaining--;
            }
            if (remaining > 0 && waiting == remaining)          /* everyone alive is at the barrier */
                for (int t = 0; t < nt; t++) s.fibers[t].at_barrier = false;
            else if (!progressed && remaining > 0) { fprintf(stderr, "emu: deadlock\n"); abort(); }
        }
    }
    s.cur = -1;
}
inline int lane() { State &s = S(); return s.cur & 31; }
inline Warp &warp() { State &s = S(); return s.warps[s.cur >> 5]; }
/* all lanes in `mask` deposit v; returns the snapshot buffer index */
inline const uint64_t *collective(uint32_t mask, uint64_t v, uint32_t *out_mask) {
    Warp &w = warp();
    int l = lane();
    uint32_t g = w.gen;
    w.val[l] = v;
    w.arrived |= 1u << l;
    if ((w.arrived & mask) == mask) {
        memcpy(w.snap[g & 1], w.val, sizeof(w.val));
        w.snap_mask[g & 1] = w.arrived;
        w.arrived = 0;
        w.gen = g + 1;
    } else {
        while (w.gen == g) yield();
    }
    if (out_mask) *out_mask = w.snap_mask[g & 1];
    return w.snap[g & 1];
}
}  // namespace emu

extern "C" __attribute__((used, visibility("default"))) unsigned long long emu_chaos_yields() { return emu::S().chaos_yields; }
#define threadIdx (emu::S().threadIdx)
#define blockIdx (emu::S().blockIdx)
#define blockDim (emu::S().blockDim)
#define gridDim (emu::S().gridDim)
#define warpSize 32

static inline void __syncthreads() {
    emu::State &s = emu::S();
    s.fibers[s.cur].at_barrier = true;
    emu::yield();
}
static inline uint32_t emu_warp_full_mask() {
    emu::State &s = emu::S();
    int w = s.cur >> 5;
    int n = s.n_threads - w * 32;
    return n >= 32 ? 0xFFFFFFFFu : ((1u << n) - 1);
}
static inline void __syncwarp(uint32_t mask = 0xFFFFFFFFu) { emu::collective(mask & emu_warp_full_mask(), 0, nullptr); }
static inline uint32_t __activemask() { return emu_warp_full_mask(); }
template <typename T> static inline T __shfl_sync(uint32_t mask, T v, int src, int width = 32) {
    uint64_t raw = 0; memcpy(&raw, &v, sizeof(T) <= 8 ? sizeof(T) : 8);
    const uint64_t *s = emu::collective(mask & emu_warp_full_mask(), raw, nullptr);
    int l = emu::lane(); int base = l & ~(width - 1);
    T r; memcpy(&r, &s[base + (src & (width - 1))], sizeof(T) <= 8 ? sizeof(T) : 8); return r;
}
template <typename T> static inline T __shfl_xor_sync(uint32_t mask, T v, int x, int width = 32) {
    uint64_t raw = 0; memcpy(&raw, &v, sizeof(T));
    const uint64_t *s = emu::collective(mask & emu_warp_full_mask(), raw, nullptr);
    int l = emu::lane(); int p = l ^ x; if ((p & ~(width - 1)) != (l & ~(width - 1))) p = l;
    T r; memcpy(&r, &s[p], sizeof(T)); return r;
}
template <typename T> static inline T __shfl_down_sync(uint32_t mask, T v, unsigned d, int width = 32) {
    uint64_t raw = 0; memcpy(&raw, &v, sizeof(T));
    const uint64_t *s = emu::collective(mask & emu_warp_full_mask(), raw, nullptr);
    int l = emu::lane(); int p = l + (int)d; if ((p & ~(width - 1)) != (l & ~(width - 1))) p = l;
    T r; memcpy(&r, &s[p], sizeof(T)); return r;
}
template <typename T> static inline T __shfl_up_sync(uint32_t mask, T v, unsigned d, int width = 32) {
    uint64_t raw = 0; memcpy(&raw, &v, sizeof(T));
    const uint64_t *s = emu::collective(mask & emu_warp_full_mask(), raw, nullptr);
    int l = emu::lane(); int p = l - (int)d; if (p < (l & ~(width - 1))) p = l;
    T r; memcpy(&r, &s[p], sizeof(T)); return r;
}
static inline uint32_t __ballot_sync(uint32_t mask, int pred) {
    uint32_t m; const uint64_t *s = emu::collective(mask & emu_warp_full_mask(), pred ? 1 : 0, &m);
    uint32_t r = 0; for (int i = 0; i < 32; i++) if (((m >> i) & 1) && s[i]) r |= 1u << i; return r;
}
static inline int __any_sync(uint32_t mask, int pred) { return __ballot_sync(mask, pred) != 0; }
static inline int __all_sync(uint32_t mask, int pred) {
    uint32_t fm = mask & emu_warp_full_mask(); return (__ballot_sync(mask, pred) & fm) == fm;
}
template <typename T> static inline uint32_t __match_any_sync(uint32_t mask, T v) {
    uint64_t raw = 0; memcpy(&raw, &v, sizeof(T));
    uint32_t m; const uint64_t *s = emu::collective(mask & emu_warp_full_mask(), raw, &m);
    uint32_t r = 0; for (int i = 0; i < 32; i++) if (((m >> i) & 1) && s[i] == raw) r |= 1u << i; return r;
}

/* reductions over the lanes named in `mask` (REDUX): lanes of different groups may call it together with their own
   group's mask (cooperative-groups labeled partitions do); the emulator needs the WHOLE warp to make the call. */
static inline unsigned emu_reduce(uint32_t mask, unsigned v, int op) {
    uint32_t m; const uint64_t *s = emu::collective(emu_warp_full_mask(), v, &m);
    unsigned r = op == 0 ? 0u : 0xFFFFFFFFu;
    for (int i = 0; i < 32; i++) if ((mask >> i) & (m >> i) & 1) { unsigned x = (unsigned)s[i]; r = op == 0 ? r + x : (x < r ? x : r); }
    return r;
}
static inline unsigned __reduce_add_sync(uint32_t mask, unsigned v) { return emu_reduce(mask, v, 0); }
static inline unsigned __reduce_min_sync(uint32_t mask, unsigned v) { return emu_reduce(mask, v, 1); }

/* ---- atomics (fibers never preempt, so plain code is atomic; chaos mode gives way BEFORE the operation) ---- */
template <typename T> static inline T atomicAdd(T *p, T v) { emu::chaos_point(); T o = *p; *p = o + v; return o; }
template <typename T> static inline T atomicSub(T *p, T v) { emu::chaos_point(); T o = *p; *p = o - v; return o; }
template <typename T> static inline T atomicMin(T *p, T v) { emu::chaos_point(); T o = *p; if (v < o) *p = v; return o; }
template <typename T> static inline T atomicMax(T *p, T v) { emu::chaos_point(); T o = *p; if (v > o) *p = v; return o; }
template <typename T> static inline T atomicOr(T *p, T v) { emu::chaos_point(); T o = *p; *p = o | v; return o; }
template <typename T> static inline T atomicAnd(T *p, T v) { emu::chaos_point(); T o = *p; *p = o & v; return o; }
template <typename T> static inline T atomicExch(T *p, T v) { emu::chaos_point(); T o = *p; *p = v; return o; }
template <typename T> static inline T atomicCAS(T *p, T c, T v) { emu::chaos_point(); T o = *p; if (o == c) *p = v; return o; }

/* ---- intrinsics ---- */
static inline int __popc(unsigned x) { return __builtin_popcount(x); }
static inline int __popcll(unsigned long long x) { return __builtin_popcountll(x); }
static inline int __clz(int x) { return x ? __builtin_clz((unsigned)x) : 32; }
static inline int __clzll(long long x) { return x ? __builtin_clzll((unsigned long long)x) : 64; }
static inline int __ffs(int x) { return __builtin_ffs(x); }
static inline int __ffsll(long long x) { return __builtin_ffsll(x); }
static inline unsigned __brev(unsigned x) { unsigned r = 0; for (int i = 0; i < 32; i++) if (x >> i & 1) r |= 1u << (31 - i); return r; }
static inline unsigned long long __umul64hi(unsigned long long a, unsigned long long b) {
    return (unsigned long long)(((unsigned __int128)a * b) >> 64);
}
static inline unsigned __byte_perm(unsigned a, unsigned b, unsigned s) {
    unsigned long long src = ((unsigned long long)b << 32) | a; unsigned r = 0;
    for (int i = 0; i < 4; i++) { unsigned sel = (s >> (4 * i)) & 0xF; unsigned byte = (src >> (8 * (sel & 7))) & 0xFF;
        if (sel & 8) byte = (byte & 0x80) ? 0xFF : 0; r |= byte << (8 * i); }
    return r;
}
static inline unsigned __vcmpgeu4(unsigned a, unsigned b) {
    unsigned r = 0; for (int i = 0; i < 4; i++) if (((a >> (8 * i)) & 0xFF) >= ((b >> (8 * i)) & 0xFF)) r |= 0xFFu << (8 * i); return r;
}
static inline unsigned __vcmpleu4(unsigned a, unsigned b) {
    unsigned r = 0; for (int i = 0; i < 4; i++) if (((a >> (8 * i)) & 0xFF) <= ((b >> (8 * i)) & 0xFF)) r |= 0xFFu << (8 * i); return r;
}
static inline unsigned __funnelshift_l(unsigned lo, unsigned hi, unsigned s) {
    s &= 31; return s ? (hi << s) | (lo >> (32 - s)) : hi;
}
template <typename T> static inline T __ldg(const T *p) { return *p; }
template <typename T> static inline T __ldcg(const T *p) { emu::chaos_point(); return *p; }
template <typename T> static inline T __ldcs(const T *p) { return *p; }
template <typename T> static inline T __ldcv(const T *p) { emu::chaos_point(); return *p; }
template <typename T> static inline void __stcg(T *p, T v) { emu::chaos_point(); *p = v; }
template <typename T> static inline void __stcs(T *p, T v) { *p = v; }
static inline void __threadfence() {}
static inline void __threadfence_block() {}
static inline double __ddiv_rn(double a, double b) { return a / b; }
static inline unsigned long long clock64() { return 0; }

/* ---- runtime API ---- */
typedef int cudaError_t;
enum { cudaSuccess = 0, cudaErrorMemoryAllocation = 2, cudaErrorInvalidValue = 1 };
enum cudaMemcpyKind { cudaMemcpyHostToHost, cudaMemcpyHostToDevice, cudaMemcpyDeviceToHost, cudaMemcpyDeviceToDevice, cudaMemcpyDefault };
typedef struct emu_stream *cudaStream_t;
struct emu_event { double t; };
typedef emu_event *cudaEvent_t;
static inline const char *cudaGetErrorString(cudaError_t e) { return e ? "emu error" : "no error"; }
static inline cudaError_t cudaGetLastError() { return cudaSuccess; }
static inline cudaError_t cudaPeekAtLastError() { return cudaSuccess; }
static inline cudaError_t cudaSetDevice(int) { return cudaSuccess; }
static inline cudaError_t cudaGetDevice(int *d) { *d = 0; return cudaSuccess; }
static inline cudaError_t cudaGetDeviceCount(int *n) { *n = 1; return cudaSuccess; }
static inline cudaError_t cudaDeviceSynchronize() { return cudaSuccess; }
/* device memory comes back full of garbage, as on a real GPU, so reads of never-written memory show up */
template <typename T> static inline cudaError_t cudaMalloc(T **p, size_t n) { size_t m = (n + 255) / 256 * 256 + 256; *p = (T *)aligned_alloc(256, m); if (*p) memset((void *)*p, 0xA5, m); return *p ? cudaSuccess : cudaErrorMemoryAllocation; }
template <typename T> static inline cudaError_t cudaMallocHost(T **p, size_t n) { *p = (T *)aligned_alloc(256, (n + 255) / 256 * 256 + 256); return *p ? cudaSuccess : cudaErrorMemoryAllocation; }
static inline cudaError_t cudaFree(void *p) { free(p); return cudaSuccess; }
static inline cudaError_t cudaFreeHost(void *p) { free(p); return cudaSuccess; }
static inline cudaError_t cudaMemcpy(void *d, const void *s, size_t n, cudaMemcpyKind) { if (n) memmove(d, s, n); return cudaSuccess; }
static inline cudaError_t cudaMemcpyAsync(void *d, const void *s, size_t n, cudaMemcpyKind, cudaStream_t = 0) { if (n) memmove(d, s, n); return cudaSuccess; }
static inline cudaError_t cudaMemset(void *d, int v, size_t n) { if (n) memset(d, v, n); return cudaSuccess; }
static inline cudaError_t cudaMemsetAsync(void *d, int v, size_t n, cudaStream_t = 0) { if (n) memset(d, v, n); return cudaSuccess; }
static inline cudaError_t cudaStreamCreate(cudaStream_t *s) { *s = 0; return cudaSuccess; }
static inline cudaError_t cudaStreamCreateWithFlags(cudaStream_t *s, unsigned) { *s = 0; return cudaSuccess; }
static inline cudaError_t cudaStreamDestroy(cudaStream_t) { return cudaSuccess; }
static inline cudaError_t cudaStreamSynchronize(cudaStream_t) { return cudaSuccess; }
struct emu_event;
static inline cudaError_t cudaStreamWaitEvent(cudaStream_t, emu_event *, unsigned = 0) { return cudaSuccess; }
#define cudaStreamNonBlocking 1
static inline double emu_now() { struct timespec ts; clock_gettime(CLOCK_MONOTONIC, &ts); return ts.tv_sec * 1e3 + ts.tv_nsec * 1e-6; }
static inline cudaError_t cudaEventCreate(cudaEvent_t *e) { *e = new emu_event{0}; return cudaSuccess; }
static inline cudaError_t cudaEventDestroy(cudaEvent_t e) { delete e; return cudaSuccess; }
static inline cudaError_t cudaEventRecord(cudaEvent_t e, cudaStream_t = 0) { e->t = emu_now(); return cudaSuccess; }
static inline cudaError_t cudaEventSynchronize(cudaEvent_t) { return cudaSuccess; }
static inline cudaError_t cudaEventElapsedTime(float *ms, cudaEvent_t a, cudaEvent_t b) { *ms = (float)(b->t - a->t); return cudaSuccess; }
enum cudaDeviceAttr { cudaDevAttrMultiProcessorCount = 16 };
static inline cudaError_t cudaDeviceGetAttribute(int *v, cudaDeviceAttr, int) { *v = 4; return cudaSuccess; }
