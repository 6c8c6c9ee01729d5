// cuda_emu.h — a single-threaded warp emulator (TEST INFRASTRUCTURE).
// Runs the 32 lanes of one warp as cooperative fibers (ucontext); every warp collective (__shfl*_sync,
// __ballot_sync, __syncwarp, __reduce_max_sync) is a rendezvous of all live lanes. This lets the device code in
// badread_b200/csrc/*.cuh be compiled with g++ and checked against the CPU oracle in the CPU-only test tier,
// before any GPU time is spent. Divergent code that calls a collective would deadlock here, exactly as it
// would be undefined on the GPU.
#pragma once
#include <ucontext.h>

#include <algorithm>
#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <functional>

#define __device__
#define __global__
#ifndef __align__
#define __align__(n) alignas(n)
#endif
#define __host__
#define __forceinline__ inline
#define __restrict__
#define __launch_bounds__(...)
#define __shared__ static
#define __constant__

#define BB_EMULATOR 1
struct uint2 { uint32_t x, y; };
struct int2 { int x, y; };
static inline int2 make_int2(int x, int y) { return int2{x, y}; }
struct uint4 { uint32_t x, y, z, w; };
static inline uint2 make_uint2(uint32_t x, uint32_t y) { return uint2{x, y}; }
static inline uint4 make_uint4(uint32_t x, uint32_t y, uint32_t z, uint32_t w) { return uint4{x, y, z, w}; }
struct EmuDim { unsigned x, y, z; };

namespace emu {
constexpr int MAXT = 512;
constexpr size_t STACK = 1 << 20;
struct Block {
    ucontext_t main_ctx, fib[MAXT];
    char *stacks[MAXT];
    bool done[MAXT];
    int nthreads, cur, alive;
    int w_alive[MAXT / 32], w_arrived[MAXT / 32];
    uint64_t w_gen[MAXT / 32];
    int b_arrived;
    uint64_t b_gen;
    uint64_t xchg[MAXT];
    std::function<void()> body;
};
inline Block &blk() { static Block b; return b; }
}  // namespace emu
inline EmuDim threadIdx = {0, 0, 0}, blockIdx = {0, 0, 0}, blockDim = {32, 1, 1}, gridDim = {1, 1, 1};

namespace emu {
inline int next_alive(int from) {
    Block &w = blk();
    for (int k = 1; k <= w.nthreads; k++) {
        const int c = (from + k) % w.nthreads;
        if (!w.done[c]) return c;
    }
    return -1;
}
inline void yield() {
    Block &w = blk();
    const int me = w.cur;
    const int nxt = next_alive(me);
    if (nxt < 0 || nxt == me) return;
    w.cur = nxt;
    swapcontext(&w.fib[me], &w.fib[nxt]);
    threadIdx.x = (unsigned)w.cur;
}
inline void barrier() {  // warp-level rendezvous of the live lanes of the caller's warp
    Block &w = blk();
    const int wi = (int)threadIdx.x >> 5;
    const uint64_t g = w.w_gen[wi];
    w.w_arrived[wi]++;
    if (w.w_arrived[wi] >= w.w_alive[wi]) { w.w_arrived[wi] = 0; w.w_gen[wi]++; return; }
    while (w.w_gen[wi] == g) yield();
}
inline void block_barrier() {
    Block &w = blk();
    const uint64_t g = w.b_gen;
    w.b_arrived++;
    if (w.b_arrived >= w.alive) { w.b_arrived = 0; w.b_gen++; return; }
    while (w.b_gen == g) yield();
}
inline void trampoline() {
    Block &w = blk();
    threadIdx.x = (unsigned)w.cur;
    w.body();
    const int me = w.cur;
    const int wi = me >> 5;
    w.done[me] = true;
    w.alive--;
    w.w_alive[wi]--;
    if (w.w_alive[wi] > 0 && w.w_arrived[wi] >= w.w_alive[wi]) { w.w_arrived[wi] = 0; w.w_gen[wi]++; }
    if (w.alive > 0 && w.b_arrived >= w.alive) { w.b_arrived = 0; w.b_gen++; }
    const int nxt = next_alive(me);
    if (nxt < 0) { setcontext(&w.main_ctx); }
    w.cur = nxt;
    setcontext(&w.fib[nxt]);
}
inline void run_block(int nthreads, std::function<void()> body) {
    Block &w = blk();
    w.body = body;
    w.nthreads = nthreads; w.alive = nthreads; w.b_arrived = 0; w.b_gen = 0;
    blockDim.x = (unsigned)nthreads;
    for (int i = 0; i < MAXT / 32; i++) { w.w_alive[i] = 0; w.w_arrived[i] = 0; w.w_gen[i] = 0; }
    for (int i = 0; i < nthreads; i++) {
        w.done[i] = false;
        w.w_alive[i >> 5]++;
        if (!w.stacks[i]) w.stacks[i] = (char *)malloc(STACK);
        getcontext(&w.fib[i]);
        w.fib[i].uc_stack.ss_sp = w.stacks[i];
        w.fib[i].uc_stack.ss_size = STACK;
        w.fib[i].uc_link = nullptr;
        makecontext(&w.fib[i], (void (*)())trampoline, 0);
    }
    w.cur = 0;
    swapcontext(&w.main_ctx, &w.fib[0]);
}
inline void run_warp(std::function<void()> body) { run_block(32, body); }
// bar.sync id, count: rendezvous of `count` threads on barrier `id` (ids 1..15)
inline void named_barrier(int id, int count) {
    static int arrived[16];
    static uint64_t gen[16];
    const uint64_t g = gen[id];
    arrived[id]++;
    if (arrived[id] >= count) { arrived[id] = 0; gen[id]++; return; }
    while (gen[id] == g) yield();
}
template <typename T>
inline T exchange(T v, int src) {
    Block &w = blk();
    uint64_t bits = 0;
    std::memcpy(&bits, &v, sizeof(T));
    w.xchg[threadIdx.x] = bits;
    barrier();
    const uint64_t r = w.xchg[((int)threadIdx.x & ~31) | (src & 31)];
    barrier();
    T out;
    std::memcpy(&out, &r, sizeof(T));
    return out;
}
inline int lane_id() { return (int)threadIdx.x & 31; }
inline int warp_base() { return (int)threadIdx.x & ~31; }
}  // namespace emu

template <typename T> inline T __shfl_sync(unsigned, T v, int src) { return emu::exchange(v, src); }
template <typename T> inline T __shfl_up_sync(unsigned, T v, unsigned d) {
    const int me = emu::lane_id();
    return emu::exchange(v, me - (int)d >= 0 ? me - (int)d : me);
}
template <typename T> inline T __shfl_xor_sync(unsigned, T v, int m) { return emu::exchange(v, emu::lane_id() ^ m); }
inline unsigned __ballot_sync(unsigned, bool p) {
    emu::Block &w = emu::blk();
    const int b0 = emu::warp_base();
    w.xchg[threadIdx.x] = p ? 1 : 0;
    emu::barrier();
    unsigned r = 0;
    for (int i = 0; i < 32; i++) if (b0 + i < w.nthreads && !w.done[b0 + i] && w.xchg[b0 + i]) r |= 1u << i;
    emu::barrier();
    return r;
}
inline int __reduce_max_sync(unsigned, int v) {
    emu::Block &w = emu::blk();
    const int b0 = emu::warp_base();
    w.xchg[threadIdx.x] = (uint64_t)(int64_t)v;
    emu::barrier();
    int r = INT32_MIN;
    for (int i = 0; i < 32; i++) if (b0 + i < w.nthreads && !w.done[b0 + i]) r = std::max(r, (int)(int64_t)w.xchg[b0 + i]);
    emu::barrier();
    return r;
}
inline unsigned __reduce_or_sync(unsigned, unsigned v) {
    emu::Block &w = emu::blk();
    const int b0 = emu::warp_base();
    w.xchg[threadIdx.x] = v;
    emu::barrier();
    unsigned r = 0;
    for (int i = 0; i < 32; i++) if (b0 + i < w.nthreads && !w.done[b0 + i]) r |= (unsigned)w.xchg[b0 + i];
    emu::barrier();
    return r;
}
inline void __syncwarp(unsigned = 0xffffffffu) { emu::barrier(); }
inline int __popc(unsigned x) { return __builtin_popcount(x); }
inline int __ffs(unsigned x) { return __builtin_ffs((int)x); }
inline int __clz(unsigned x) { return x ? __builtin_clz(x) : 32; }
inline unsigned __brev(unsigned x) {
    x = ((x >> 1) & 0x55555555u) | ((x & 0x55555555u) << 1);
    x = ((x >> 2) & 0x33333333u) | ((x & 0x33333333u) << 2);
    x = ((x >> 4) & 0x0f0f0f0fu) | ((x & 0x0f0f0f0fu) << 4);
    return __builtin_bswap32(x);
}
inline unsigned __funnelshift_r(unsigned lo, unsigned hi, unsigned sh) {
    const uint64_t v = ((uint64_t)hi << 32) | lo;
    return (unsigned)(v >> (sh & 31));
}
inline unsigned __funnelshift_l(unsigned lo, unsigned hi, unsigned sh) {
    const uint64_t v = ((uint64_t)hi << 32) | lo;
    return (unsigned)((v << (sh & 31)) >> 32);
}
inline unsigned __umulhi(unsigned a, unsigned b) { return (unsigned)(((uint64_t)a * b) >> 32); }
inline double __dmul_rn(double a, double b) { return a * b; }
inline double __dadd_rn(double a, double b) { return a + b; }
inline double __dsub_rn(double a, double b) { return a - b; }
inline double __ddiv_rn(double a, double b) { return a / b; }
inline double __dsqrt_rn(double a) { return __builtin_sqrt(a); }
template <typename T> inline T __ldg(const T *p) { return *p; }
inline int atomicAdd(int *p, int v) { const int o = *p; *p = o + v; return o; }
inline unsigned atomicAdd(unsigned *p, unsigned v) { const unsigned o = *p; *p = o + v; return o; }
inline int atomicOr(int *p, int v) { const int o = *p; *p = o | v; return o; }
inline int atomicExch(int *p, int v) { const int o = *p; *p = v; return o; }
inline int atomicMax(int *p, int v) { const int o = *p; if (v > o) *p = v; return o; }
inline unsigned long long atomicAdd(unsigned long long *p, unsigned long long v) { const unsigned long long o = *p; *p = o + v; return o; }
inline unsigned long long atomicMin(unsigned long long *p, unsigned long long v) { const unsigned long long o = *p; if (v < o) *p = v; return o; }
inline unsigned long long atomicCAS(unsigned long long *p, unsigned long long cmp, unsigned long long v) {
    const unsigned long long o = *p;
    if (o == cmp) *p = v;
    return o;
}
inline long long clock64() { return 0; }
inline void __syncthreads() { emu::block_barrier(); }
inline void __threadfence_block() {}
inline bool __all_sync(unsigned m, bool p) { return __ballot_sync(m, !p) == 0u; }
inline bool __any_sync(unsigned m, bool p) { return __ballot_sync(m, p) != 0u; }
using std::max;
using std::min;
