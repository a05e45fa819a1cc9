// simt_emu.h — TEST INFRASTRUCTURE ONLY: a small SIMT emulator that lets the CUDA sources under
// texture_gs_b200/csrc compile with g++ and run on the host, so that `-m "not gpu"` tests can execute the REAL kernel
// source (warp collectives, block barriers, mbarrier / bulk-copy pipelines, atomics) against the oracle on a box
// without a GPU. Nothing in the product imports or links this; tests/simt/emu.py builds it into tests/simt/_build/.
//
// Execution model: the blocks of a launch run one after the other on the calling OS thread; every CUDA thread of a
// block is a fiber (ucontext) with its own stack. A fiber runs until it blocks in a warp collective, a block barrier
// or an mbarrier wait; the scheduler resumes fibers whose wait condition has changed. Deliberately adversarial where
// real hardware is permissive only by luck:
//   * a warp collective whose mask names a lane that has already exited is an error;
//   * bulk copies (cp.async.bulk) are DEFERRED until no thread of the block can make progress without them, so shared
//     memory read before the mbarrier wait holds stale data, and a block that ends with copies in flight is an error;
//   * dynamic shared memory is filled with NaN patterns before every block;
//   * the warps of a block can be interleaved at random, and the blocks of a launch run in random order
//     (simt_set_schedule_seed), instead of block after block, warp after warp;
//   * fast-math intrinsics can return results perturbed by a repeatable error of a few ulp (simt_set_fastmath_noise);
//   * no runnable fiber while some are unfinished (divergent barrier, missing arrival) is reported as a deadlock.
#pragma once
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <ucontext.h>

#include <algorithm>
#include <deque>
#include <string>
#include <unordered_map>
#include <vector>

#define TEXGS_HOST_EMU 1

#define __global__
#define __device__
#define __host__
#define __forceinline__ inline
#define __noinline__ __attribute__((noinline))
#define __launch_bounds__(...)
#define __constant__
#define __shared__ thread_local        // block scope: implicitly static; `extern __shared__ T x[];` stays a declaration
#define __align__(n) __attribute__((aligned(n)))

// ---------------------------------------------------------------------------------------------
// vector types
// ---------------------------------------------------------------------------------------------
struct alignas(8) float2 { float x, y; };
struct float3 { float x, y, z; };
struct alignas(16) float4 { float x, y, z, w; };
struct alignas(8) int2 { int x, y; };
struct int3 { int x, y, z; };
struct alignas(16) int4 { int x, y, z, w; };
struct alignas(8) uint2 { unsigned x, y; };
struct uint3 { unsigned x, y, z; };
struct alignas(16) uint4 { unsigned x, y, z, w; };
struct alignas(16) double2 { double x, y; };
inline float2 make_float2(float x, float y) { return float2{x, y}; }
inline float3 make_float3(float x, float y, float z) { return float3{x, y, z}; }
inline float4 make_float4(float x, float y, float z, float w) { return float4{x, y, z, w}; }
inline int2 make_int2(int x, int y) { return int2{x, y}; }
inline int3 make_int3(int x, int y, int z) { return int3{x, y, z}; }
inline int4 make_int4(int x, int y, int z, int w) { return int4{x, y, z, w}; }
inline uint2 make_uint2(unsigned x, unsigned y) { return uint2{x, y}; }
inline uint3 make_uint3(unsigned x, unsigned y, unsigned z) { return uint3{x, y, z}; }
inline uint4 make_uint4(unsigned x, unsigned y, unsigned z, unsigned w) { return uint4{x, y, z, w}; }
inline double2 make_double2(double x, double y) { return double2{x, y}; }

struct dim3 {
    unsigned x, y, z;
    dim3(unsigned x_ = 1, unsigned y_ = 1, unsigned z_ = 1) : x(x_), y(y_), z(z_) {}
};

// ---------------------------------------------------------------------------------------------
// scheduler
// ---------------------------------------------------------------------------------------------
namespace simt {

// Context switch between fibers. glibc's swapcontext saves the signal mask with a system call on every switch; a
// warp collective costs 64 switches, so on x86-64 a minimal switch (callee-saved registers + stack pointer; nobody in
// the emulated code changes the floating-point control words) is used instead. SIMT_USE_UCONTEXT=1 forces ucontext.
#if defined(__x86_64__) && !defined(SIMT_USE_UCONTEXT)
#define SIMT_FAST_SWITCH 1
struct Context { void* sp = nullptr; };
extern "C" void simt_switch_asm(void** save_sp, void* load_sp);
__asm__(
    ".text\n"
    ".globl simt_switch_asm\n"
    ".type simt_switch_asm,@function\n"
    "simt_switch_asm:\n"
    "    pushq %rbp\n    pushq %rbx\n    pushq %r12\n    pushq %r13\n    pushq %r14\n    pushq %r15\n"
    "    movq %rsp, (%rdi)\n"
    "    movq %rsi, %rsp\n"
    "    popq %r15\n    popq %r14\n    popq %r13\n    popq %r12\n    popq %rbx\n    popq %rbp\n"
    "    ret\n"
    ".size simt_switch_asm,.-simt_switch_asm\n");
inline void switch_context(Context* from, Context* to) { simt_switch_asm(&from->sp, to->sp); }
inline void make_context(Context* c, char* stack, size_t bytes, void (*entry)()) {
    uintptr_t top = ((uintptr_t)stack + bytes) & ~(uintptr_t)15;
    void** sp = (void**)top;
    *--sp = nullptr;                 // return address of ``entry`` (it never returns)
    *--sp = (void*)entry;            // popped by the ``ret`` of the first switch: rsp is then 8 mod 16, as after a call
    for (int i = 0; i < 6; ++i) *--sp = nullptr;      // rbp rbx r12 r13 r14 r15
    c->sp = (void*)sp;
}
#else
#define SIMT_FAST_SWITCH 0
struct Context { ucontext_t uc; };
inline void switch_context(Context* from, Context* to) { swapcontext(&from->uc, &to->uc); }
inline void make_context(Context* c, char* stack, size_t bytes, void (*entry)()) {
    getcontext(&c->uc);
    c->uc.uc_stack.ss_sp = stack;
    c->uc.uc_stack.ss_size = bytes;
    c->uc.uc_link = nullptr;
    makecontext(&c->uc, entry, 0);
}
#endif

constexpr size_t kStackBytes = 160 * 1024;
constexpr size_t kDynSmemMax = 232448;      // 227 KB, the sm_100 opt-in maximum

struct PendingCopy { void* dst; const void* src; unsigned bytes; int op = 0; };   // op 0: copy, 1: add.f32 (bulk reduction)

struct Fiber {
    Context ctx;
    uint3 tid;
    int lin = 0, lane = 0, warp = 0;
    bool done = false;
    // blocked while *wait_addr == wait_val (wait_addr == nullptr: runnable)
    const volatile unsigned* wait_addr = nullptr;
    unsigned wait_val = 0;
    const char* wait_what = "";
    // cp.async.bulk shared -> global (bulk_group completion): copies of the group being built and the committed groups
    // still in flight, oldest first
    std::vector<PendingCopy> bulk_open;
    std::deque<std::vector<PendingCopy>> bulk_groups;
};

struct Warp {
    unsigned exist = 0;      // lanes that exist in this warp
    unsigned exited = 0;     // lanes that have returned from the kernel
    unsigned arrived = 0;
    volatile unsigned gen = 0;
    uint64_t in[32], out[32];
    unsigned mask_in = 0;
};

struct MBar {
    volatile unsigned phase = 0;   // parity of the phase in progress
    int count = 0, pending = 0;
    long long tx = 0;
};

struct Global {
    dim3 bid, bdim, gdim;
    std::vector<Fiber> fibers;
    std::vector<char*> stacks;
    std::vector<Warp> warps;
    Context sched;
    Fiber* cur = nullptr;
    void (*entry)(void*) = nullptr;
    void* entry_arg = nullptr;
    // block barrier
    int live = 0, bar_arrived = 0;
    volatile unsigned bar_gen = 0;
    // mbarriers of the running block, keyed by shared-memory address
    std::unordered_map<const void*, MBar> mbars;
    std::unordered_map<const void*, std::deque<PendingCopy>> copies;
    // error state (sticky until read by cudaGetLastError)
    int err = 0;
    std::string err_msg;
    bool abort_block = false;
    unsigned long long launches = 0, collectives = 0;
    // bulk copies: false = land as LATE as legal (exposes reads before the mbarrier wait), true = land at issue, i.e. as
    // EARLY as legal (exposes a stage overwritten while other lanes still read its previous content)
    bool eager_copies = false;
    // fast-math intrinsics (__expf, __logf, __fdividef, rsqrtf): 0 = exact; n > 0 = results perturbed by a pseudo-random
    // (but repeatable: a hash of the value) relative error of up to n * 2^-23 — the GPU's ex2.approx / rcp.approx /
    // lg2.approx paths are not correctly rounded
    // scheduling: 0 = warp after warp, each as far as it can go (deterministic); otherwise the seed of a random
    // interleaving of the warps of a block (random order, random time slices, random first lane): intra-block races
    unsigned sched_seed = 0, sched_state = 1;
    // vote profiler: per call site of __ballot/__any/__all_sync, how many lanes voted true (histogram 0..32)
    bool profile_votes = false;
    struct VoteStat { unsigned long long calls = 0, hist[33] = {0}; };
    std::unordered_map<const void*, VoteStat> votes;
    // with the profiler on, every __ballot_sync also leaves {launch number, block << 8 | warp, lanes true} here, in
    // execution order: the per-warp sequence of survivor counts the list walk of the render kernels is made of
    std::vector<uint32_t> ballot_trace;
    unsigned fastmath_noise_ulps = 0;
    uint32_t noise_state = 0x9e3779b9u;
};
inline Global& G() { static Global g; return g; }

inline void fail(const std::string& msg) {
    Global& g = G();
    if (!g.err) {
        g.err = 700;
        char where[96];
        snprintf(where, sizeof(where), " [block (%u,%u,%u)]", g.bid.x, g.bid.y, g.bid.z);
        g.err_msg = "simt emulator: " + msg + where;
    }
    g.abort_block = true;
}

inline void yield_to_scheduler() { Global& g = G(); switch_context(&g.cur->ctx, &g.sched); }

// block the running fiber while *addr == val
inline void wait_while_equal(const volatile unsigned* addr, unsigned val, const char* what) {
    Global& g = G();
    while (*addr == val && !g.abort_block) {
        g.cur->wait_addr = addr; g.cur->wait_val = val; g.cur->wait_what = what;
        yield_to_scheduler();
    }
    g.cur->wait_addr = nullptr;
    if (g.abort_block) {           // unwind: park this fiber for good
        g.cur->done = true;
        for (;;) yield_to_scheduler();
    }
}

inline void fiber_main() {
    Global& g = G();
    g.entry(g.entry_arg);
    Fiber* f = g.cur;
    f->done = true;
    if (!f->bulk_open.empty() || !f->bulk_groups.empty()) fail("a thread exited with bulk stores in flight (no cp.async.bulk.wait_group covered them): their shared-memory source dies with the block");
    Warp& w = g.warps[f->warp];
    w.exited |= 1u << f->lane;
    if (w.arrived && (w.mask_in & (1u << f->lane))) fail("a lane exited while its warp waits for it in a *_sync collective");
    g.live -= 1;
    if (g.live > 0 && g.bar_arrived == g.live) {   // the remaining threads all wait at __syncthreads
        g.bar_arrived = 0;
        g.bar_gen = g.bar_gen + 1;
    }
    for (;;) yield_to_scheduler();
}

inline uint64_t* warp_collective(unsigned mask, uint64_t v) {
    Global& g = G();
    Fiber* f = g.cur;
    Warp& w = g.warps[f->warp];
    mask &= w.exist;
    if (!(mask & (1u << f->lane))) { fail("lane calls a *_sync collective without being in its mask"); wait_while_equal(&w.gen, w.gen, "abort"); }
    if (mask & w.exited) { fail("*_sync collective names a lane that has already exited"); wait_while_equal(&w.gen, w.gen, "abort"); }
    if (w.arrived == 0) w.mask_in = mask;
    else if (w.mask_in != mask) { fail("lanes of one warp meet in *_sync collectives with different masks"); wait_while_equal(&w.gen, w.gen, "abort"); }
    w.in[f->lane] = v;
    w.arrived |= 1u << f->lane;
    const unsigned gen = w.gen;
    if ((w.arrived & mask) == mask) {
        memcpy(w.out, w.in, sizeof(w.out));
        w.arrived = 0;
        w.gen = gen + 1;
        g.collectives++;
    } else {
        wait_while_equal(&w.gen, gen, "warp collective");
    }
    return w.out;
}

inline void block_barrier() {
    Global& g = G();
    const unsigned gen = g.bar_gen;
    g.bar_arrived += 1;
    if (g.bar_arrived == g.live) {
        g.bar_arrived = 0;
        g.bar_gen = gen + 1;
    } else {
        wait_while_equal(&g.bar_gen, gen, "__syncthreads");
    }
}

void run_block();   // defined below

template <class F>
inline void launch(dim3 grid, dim3 block, size_t smem_bytes, F&& body);

// `extern __shared__ T name[];` in a kernel is a declaration of a thread_local array here; the emulated translation
// unit defines each such array once with SIMT_DEFINE_DYN_SMEM(T, name), which also registers it for poisoning
struct DynSmem { void* p; size_t bytes; };
inline std::vector<DynSmem>& dyn_smem_arrays() { static std::vector<DynSmem> v; return v; }
struct DynSmemRegistrar { DynSmemRegistrar(void* p, size_t n) { dyn_smem_arrays().push_back(DynSmem{p, n}); } };

}  // namespace simt

#define threadIdx (simt::G().cur->tid)
#define blockIdx (simt::G().bid)
#define blockDim (simt::G().bdim)
#define gridDim (simt::G().gdim)

// ---------------------------------------------------------------------------------------------
// warp / block intrinsics
// ---------------------------------------------------------------------------------------------
namespace simt {
template <class T> inline uint64_t to_bits(T v) { static_assert(sizeof(T) <= 8, "shuffle payload"); uint64_t b = 0; memcpy(&b, &v, sizeof(T)); return b; }
template <class T> inline T from_bits(uint64_t b) { T v; memcpy(&v, &b, sizeof(T)); return v; }
inline int my_lane() { return G().cur->lane; }
}  // namespace simt

template <class T> inline T __shfl_sync(unsigned mask, T v, int src, int width = 32) {
    const int lane = simt::my_lane();
    uint64_t* out = simt::warp_collective(mask, simt::to_bits(v));
    const int s = (lane & ~(width - 1)) | (src & (width - 1));
    return simt::from_bits<T>(out[s]);
}
template <class T> inline T __shfl_xor_sync(unsigned mask, T v, int lanemask, int width = 32) {
    const int lane = simt::my_lane();
    uint64_t* out = simt::warp_collective(mask, simt::to_bits(v));
    const int s = lane ^ lanemask;
    return simt::from_bits<T>(out[(s / width == lane / width) ? s : lane]);
}
template <class T> inline T __shfl_up_sync(unsigned mask, T v, unsigned delta, int width = 32) {
    const int lane = simt::my_lane();
    uint64_t* out = simt::warp_collective(mask, simt::to_bits(v));
    const int s = lane - (int)delta;
    return simt::from_bits<T>(out[(s >= (lane & ~(width - 1))) ? s : lane]);
}
template <class T> inline T __shfl_down_sync(unsigned mask, T v, unsigned delta, int width = 32) {
    const int lane = simt::my_lane();
    uint64_t* out = simt::warp_collective(mask, simt::to_bits(v));
    const int s = lane + (int)delta;
    return simt::from_bits<T>(out[(s <= (lane | (width - 1))) ? s : lane]);
}
namespace simt {
// noipa + the empty asm: the compiler must not treat two calls as the same value and merge them
__attribute__((noinline, noipa)) inline const void* call_site() {
    const void* p = __builtin_return_address(0);
    __asm__ volatile("" : "+r"(p));
    return p;
}
inline unsigned ballot_at(const void* site, unsigned mask, int pred, bool trace = false) {
    Global& g = G();
    const unsigned eff = mask & g.warps[g.cur->warp].exist;
    const int lane = g.cur->lane;
    uint64_t* out = warp_collective(mask, pred ? 1u : 0u);
    unsigned r = 0;
    for (int l = 0; l < 32; ++l) if ((eff >> l) & 1u) r |= (unsigned)(out[l] & 1u) << l;
    if (g.profile_votes && lane == __builtin_ctz(eff)) {       // once per collective: the lowest participating lane records
        Global::VoteStat& v = g.votes[site];
        v.calls++;
        v.hist[__builtin_popcount(r)]++;
        if (trace) {
            g.ballot_trace.push_back((uint32_t)g.launches);
            g.ballot_trace.push_back((uint32_t)(g.bid.x << 8) | (uint32_t)g.cur->warp);
            g.ballot_trace.push_back((uint32_t)__builtin_popcount(r));
        }
    }
    return r;
}
}  // namespace simt
// always inlined into the kernel, so that call_site() names the line of the kernel that votes
__attribute__((always_inline)) inline unsigned __ballot_sync(unsigned mask, int pred) { return simt::ballot_at(simt::call_site(), mask, pred, true); }
__attribute__((always_inline)) inline int __any_sync(unsigned mask, int pred) { return simt::ballot_at(simt::call_site(), mask, pred) != 0u; }
__attribute__((always_inline)) inline int __all_sync(unsigned mask, int pred) {
    simt::Global& g = simt::G();
    const unsigned eff = mask & g.warps[g.cur->warp].exist;
    return simt::ballot_at(simt::call_site(), mask, pred) == eff;
}
inline void __syncwarp(unsigned mask = 0xffffffffu) { simt::warp_collective(mask, 0); }
inline void __syncthreads() { simt::block_barrier(); }
inline void __threadfence() {}
inline void __threadfence_block() {}

// ---------------------------------------------------------------------------------------------
// scalar intrinsics
// ---------------------------------------------------------------------------------------------
template <class T> inline T __ldg(const T* p) { return *p; }
template <class T> inline T __ldcs(const T* p) { return *p; }
namespace simt {
inline float approx(float exact) {
    Global& g = G();
    if (g.fastmath_noise_ulps == 0 || !(exact == exact)) return exact;
    // a deterministic function of the value, like the hardware units: the same input gives the same output in every
    // kernel (the count and the scatter kernel of the binning rely on that, as do forward and backward)
    uint32_t h;
    memcpy(&h, &exact, 4);
    h = (h ^ g.noise_state) * 0x9E3779B1u; h ^= h >> 15; h *= 0x85EBCA77u; h ^= h >> 13;
    const float r = (float)((int32_t)h) * (1.0f / 2147483648.0f);                      // [-1, 1)
    return exact * (1.0f + r * (float)g.fastmath_noise_ulps * 1.1920929e-7f);
}
}  // namespace simt
inline float __expf(float x) { return simt::approx(expf(x)); }
inline float __logf(float x) { return simt::approx(logf(x)); }
inline float __fdividef(float a, float b) { return simt::approx(a / b); }
inline float __frcp_rn(float a) { return 1.0f / a; }
inline float rsqrtf(float x) { return simt::approx(1.0f / sqrtf(x)); }
inline float __saturatef(float x) { return fminf(fmaxf(x, 0.f), 1.f); }
inline float __fmaf_rn(float a, float b, float c) { return fmaf(a, b, c); }
inline int __float_as_int(float f) { int i; memcpy(&i, &f, 4); return i; }
inline unsigned __float_as_uint(float f) { unsigned i; memcpy(&i, &f, 4); return i; }
inline float __int_as_float(int i) { float f; memcpy(&f, &i, 4); return f; }
inline float __uint_as_float(unsigned i) { float f; memcpy(&f, &i, 4); return f; }
inline int __popc(unsigned x) { return __builtin_popcount(x); }
inline int __ffs(int x) { return __builtin_ffs(x); }
inline int __clz(int x) { return x ? __builtin_clz((unsigned)x) : 32; }
inline unsigned __brev(unsigned x) { unsigned r = 0; for (int i = 0; i < 32; ++i) r |= ((x >> i) & 1u) << (31 - i); return r; }

inline int min(int a, int b) { return a < b ? a : b; }
inline int max(int a, int b) { return a > b ? a : b; }
inline unsigned min(unsigned a, unsigned b) { return a < b ? a : b; }
inline unsigned max(unsigned a, unsigned b) { return a > b ? a : b; }
inline unsigned min(unsigned a, int b) { return min(a, (unsigned)b); }
inline unsigned max(unsigned a, int b) { return max(a, (unsigned)b); }
inline unsigned min(int a, unsigned b) { return min((unsigned)a, b); }
inline unsigned max(int a, unsigned b) { return max((unsigned)a, b); }
inline long long min(long long a, long long b) { return a < b ? a : b; }
inline long long max(long long a, long long b) { return a > b ? a : b; }
inline unsigned long min(unsigned long a, unsigned long b) { return a < b ? a : b; }
inline unsigned long max(unsigned long a, unsigned long b) { return a > b ? a : b; }
inline unsigned long long min(unsigned long long a, unsigned long long b) { return a < b ? a : b; }
inline unsigned long long max(unsigned long long a, unsigned long long b) { return a > b ? a : b; }
inline float min(float a, float b) { return fminf(a, b); }
inline float max(float a, float b) { return fmaxf(a, b); }
inline double min(double a, double b) { return fmin(a, b); }
inline double max(double a, double b) { return fmax(a, b); }

// atomics: fibers of one OS thread never interleave inside these
template <class T> inline T simt_atomic_add(T* a, T v) { T old = *a; *a = old + v; return old; }
inline float atomicAdd(float* a, float v) { return simt_atomic_add(a, v); }
inline double atomicAdd(double* a, double v) { return simt_atomic_add(a, v); }
inline int atomicAdd(int* a, int v) { return simt_atomic_add(a, v); }
inline unsigned atomicAdd(unsigned* a, unsigned v) { return simt_atomic_add(a, v); }
inline unsigned long long atomicAdd(unsigned long long* a, unsigned long long v) { return simt_atomic_add(a, v); }
inline int atomicMax(int* a, int v) { int o = *a; if (v > o) *a = v; return o; }
inline unsigned atomicMax(unsigned* a, unsigned v) { unsigned o = *a; if (v > o) *a = v; return o; }
inline int atomicMin(int* a, int v) { int o = *a; if (v < o) *a = v; return o; }
inline unsigned atomicMin(unsigned* a, unsigned v) { unsigned o = *a; if (v < o) *a = v; return o; }
inline unsigned atomicOr(unsigned* a, unsigned v) { unsigned o = *a; *a = o | v; return o; }
inline unsigned atomicExch(unsigned* a, unsigned v) { unsigned o = *a; *a = v; return o; }
inline unsigned atomicCAS(unsigned* a, unsigned cmp, unsigned v) { unsigned o = *a; if (o == cmp) *a = v; return o; }

// ---------------------------------------------------------------------------------------------
// PTX wrappers of texgs_common.cuh (mbarrier, 1-D bulk copy, vector reduction)
// ---------------------------------------------------------------------------------------------
namespace simt {
inline void mbar_check(MBar& b) {
    if (b.pending == 0 && b.tx == 0) {
        b.phase = b.phase ^ 1u;
        b.pending = b.count;
    }
}
inline void mbar_flush(const void* bar) {
    Global& g = G();
    auto it = g.copies.find(bar);
    if (it == g.copies.end()) return;
    MBar& b = g.mbars[bar];
    while (!it->second.empty()) {
        const PendingCopy c = it->second.front();
        it->second.pop_front();
        memcpy(c.dst, c.src, c.bytes);
        b.tx -= c.bytes;
        mbar_check(b);
    }
}
}  // namespace simt

inline void mbar_init(uint64_t* bar, uint32_t count) {
    simt::MBar& b = simt::G().mbars[bar];
    b.phase = 0; b.count = (int)count; b.pending = (int)count; b.tx = 0;
    *bar = 0x6d626172u;   // marks the slot as initialised for debugging
}
inline void mbar_fence_init() {}
inline void mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes) {
    simt::Global& g = simt::G();
    auto it = g.mbars.find(bar);
    if (it == g.mbars.end()) { simt::fail("mbarrier.arrive on an uninitialised barrier"); return; }
    simt::MBar& b = it->second;
    b.tx += bytes;
    if (b.pending <= 0) { simt::fail("mbarrier: more arrivals than the barrier was initialised for"); return; }
    b.pending -= 1;
    simt::mbar_check(b);
}
inline void mbar_wait(uint64_t* bar, uint32_t parity) {
    simt::Global& g = simt::G();
    auto it = g.mbars.find(bar);
    if (it == g.mbars.end()) { simt::fail("mbarrier wait on an uninitialised barrier"); parity = 2; }
    // deferred bulk copies land only when no thread of the block can run any more (run_block): as late as legal
    if (it != g.mbars.end()) simt::wait_while_equal(&it->second.phase, parity, "mbarrier wait");
}
inline void bulk_g2s(void* smem_dst, const void* gmem_src, uint32_t bytes, uint64_t* bar) {
    if ((bytes & 15u) || ((uintptr_t)smem_dst & 15u) || ((uintptr_t)gmem_src & 15u)) simt::fail("cp.async.bulk: size / addresses must be multiples of 16");
    simt::G().copies[bar].push_back(simt::PendingCopy{smem_dst, gmem_src, bytes});
    if (simt::G().eager_copies) simt::mbar_flush(bar);
}
// cp.async.bulk shared::cta -> global and its reducing form (cp.reduce.async.bulk ... .add.f32), bulk_group completion.
// Late as legal: the data leaves shared memory at the wait that covers the group; eager: at issue.
inline void simt_apply_bulk(const simt::PendingCopy& c) {
    if (c.op == 0) { memcpy(c.dst, c.src, c.bytes); return; }
    float* d = static_cast<float*>(c.dst);
    const float* s = static_cast<const float*>(c.src);
    for (unsigned i = 0; i < c.bytes / 4; ++i) d[i] += s[i];
}
inline void simt_bulk_issue(void* gdst, const void* ssrc, unsigned bytes, int op) {
    if ((bytes & 15u) || ((uintptr_t)gdst & 15u) || ((uintptr_t)ssrc & 15u)) simt::fail("cp.async.bulk (shared -> global): size / addresses must be multiples of 16");
    simt::Global& g = simt::G();
    const simt::PendingCopy c{gdst, ssrc, bytes, op};
    if (g.eager_copies) simt_apply_bulk(c);
    else g.cur->bulk_open.push_back(c);
}
inline void bulk_s2g(void* gdst, const void* ssrc, uint32_t bytes) { simt_bulk_issue(gdst, ssrc, bytes, 0); }
inline void bulk_reduce_add_f32(float* gdst, const float* ssrc, uint32_t bytes) { simt_bulk_issue(gdst, ssrc, bytes, 1); }
inline void bulk_commit() {
    simt::Fiber* f = simt::G().cur;
    f->bulk_groups.push_back(std::move(f->bulk_open));
    f->bulk_open.clear();
}
template <int N> inline void bulk_wait() {
    simt::Fiber* f = simt::G().cur;
    while ((int)f->bulk_groups.size() > N) {
        for (const simt::PendingCopy& c : f->bulk_groups.front()) simt_apply_bulk(c);
        f->bulk_groups.pop_front();
    }
}
inline void fence_proxy_async_smem() {}
inline void red_add_v4(float* addr, float a, float b, float c, float d) {
    if ((uintptr_t)addr & 15u) simt::fail("red.global.add.v4.f32: address must be 16-byte aligned");
    addr[0] += a; addr[1] += b; addr[2] += c; addr[3] += d;
}

// ---------------------------------------------------------------------------------------------
// the few CUDA runtime calls the host side of libtexgs makes
// ---------------------------------------------------------------------------------------------
typedef int cudaError_t;
enum { cudaSuccess = 0 };
typedef void* cudaStream_t;
typedef void* cudaEvent_t;
enum cudaMemcpyKind { cudaMemcpyHostToHost, cudaMemcpyHostToDevice, cudaMemcpyDeviceToHost, cudaMemcpyDeviceToDevice };
enum { cudaFuncAttributeMaxDynamicSharedMemorySize = 8, cudaDevAttrMultiProcessorCount = 16 };
inline const char* cudaGetErrorString(cudaError_t) { return simt::G().err_msg.c_str(); }
inline cudaError_t cudaGetLastError() { simt::Global& g = simt::G(); const int e = g.err; g.err = 0; return e; }
inline cudaError_t cudaStreamSynchronize(cudaStream_t) { return cudaSuccess; }
inline cudaError_t cudaEventRecord(cudaEvent_t, cudaStream_t) { return cudaSuccess; }
inline cudaError_t cudaStreamWaitEvent(cudaStream_t, cudaEvent_t, unsigned) { return cudaSuccess; }
inline cudaError_t cudaMemsetAsync(void* p, int v, size_t n, cudaStream_t) { memset(p, v, n); return cudaSuccess; }
inline cudaError_t cudaMemcpyAsync(void* d, const void* s, size_t n, cudaMemcpyKind, cudaStream_t) { memcpy(d, s, n); return cudaSuccess; }
inline cudaError_t cudaGetDevice(int* d) { *d = 0; return cudaSuccess; }
template <class K> inline cudaError_t cudaFuncSetAttribute(K, int, int bytes) {
    return (size_t)bytes <= simt::kDynSmemMax ? cudaSuccess : 1;
}
inline cudaError_t cudaDeviceGetAttribute(int* v, int, int) { *v = 148; return cudaSuccess; }

// ---------------------------------------------------------------------------------------------
// launch
// ---------------------------------------------------------------------------------------------
namespace simt {

inline void run_block() {
    Global& g = G();
    const unsigned nthreads = g.bdim.x * g.bdim.y * g.bdim.z;
    const unsigned nwarps = (nthreads + 31) / 32;
    if (g.fibers.size() < nthreads) g.fibers.resize(nthreads);
    while (g.stacks.size() < nthreads) g.stacks.push_back((char*)aligned_alloc(64, kStackBytes));
    g.warps.assign(nwarps, Warp());
    g.live = (int)nthreads; g.bar_arrived = 0; g.bar_gen = 0;
    g.mbars.clear(); g.copies.clear();
    g.abort_block = false;
    for (unsigned t = 0; t < nthreads; ++t) {
        Fiber& f = g.fibers[t];
        f.lin = (int)t; f.lane = (int)(t & 31); f.warp = (int)(t >> 5);
        f.tid = uint3{t % g.bdim.x, (t / g.bdim.x) % g.bdim.y, t / (g.bdim.x * g.bdim.y)};
        f.done = false; f.wait_addr = nullptr; f.bulk_open.clear(); f.bulk_groups.clear();
        g.warps[f.warp].exist |= 1u << f.lane;
        make_context(&f.ctx, g.stacks[t], kStackBytes, &fiber_main);
    }
    unsigned remaining = nthreads;
    std::vector<unsigned> order(nwarps);
    for (unsigned w = 0; w < nwarps; ++w) order[w] = w;
    auto rnd = [&g]() { g.sched_state = g.sched_state * 1664525u + 1013904223u; return g.sched_state >> 8; };
    while (remaining > 0 && !g.abort_block) {
        bool ran_any = false;
        if (g.sched_seed)                           // random warp order, random time slices, random first lane
            for (unsigned i = nwarps; i > 1; --i) std::swap(order[i - 1], order[rnd() % i]);
        for (unsigned wi = 0; wi < nwarps && !g.abort_block; ++wi) {
            const unsigned w = order[wi];
            bool warp_ran = true;
            unsigned budget = g.sched_seed ? 1 + rnd() % 6 : ~0u;
            const unsigned l0 = g.sched_seed ? rnd() % 32 : 0;
            while (warp_ran && budget-- > 0 && !g.abort_block) {
                warp_ran = false;
                for (unsigned li = 0; li < 32; ++li) {
                    const unsigned l = (li + l0) & 31u;
                    const unsigned t = w * 32 + l;
                    if (t >= nthreads) continue;
                    Fiber& f = g.fibers[t];
                    if (f.done) continue;
                    if (f.wait_addr && *f.wait_addr == f.wait_val) continue;
                    g.cur = &f;
                    switch_context(&g.sched, &f.ctx);
                    warp_ran = ran_any = true;
                    if (f.done) --remaining;
                    if (g.abort_block) break;
                }
            }
        }
        if (!ran_any && remaining > 0 && !g.abort_block) {
            bool landed = false;                    // everybody is stuck: now the outstanding bulk copies complete
            for (auto& kv : g.copies)
                if (!kv.second.empty()) { mbar_flush(kv.first); landed = true; }
            if (landed) continue;
            std::string what = "deadlock: no runnable thread;";
            int shown = 0;
            for (unsigned t = 0; t < nthreads && shown < 4; ++t)
                if (!g.fibers[t].done) { what += " thread " + std::to_string(t) + " waits in " + g.fibers[t].wait_what + ";"; ++shown; }
            fail(what);
        }
    }
    if (!g.abort_block)
        for (auto& kv : g.copies)
            if (!kv.second.empty()) { fail("block ended with bulk copies in flight (no thread waited on their mbarrier)"); break; }
    g.cur = nullptr;
}

template <class F> inline void entry_thunk(void* p) { (*static_cast<F*>(p))(); }

template <class F>
inline void launch(dim3 grid, dim3 block, size_t smem_bytes, F&& body) {
    Global& g = G();
    if (g.err) return;                                   // sticky error, like a CUDA context after a fault
    const unsigned long long nthreads = (unsigned long long)block.x * block.y * block.z;
    if (nthreads == 0 || nthreads > 1024 || smem_bytes > kDynSmemMax || grid.x == 0 || grid.y == 0 || grid.z == 0 ||
        grid.y > 65535 || grid.z > 65535) {
        g.err = 9; g.err_msg = "invalid configuration argument"; return;
    }
    typedef typename std::remove_reference<F>::type Fn;
    g.entry = &entry_thunk<Fn>;
    g.entry_arg = (void*)&body;
    g.bdim = block; g.gdim = grid;
    g.launches++;
    static const bool trace = getenv("SIMT_TRACE") != nullptr;
    if (trace) fprintf(stderr, "simt launch #%llu grid (%u,%u,%u) block (%u,%u,%u) smem %zu\n", g.launches, grid.x, grid.y, grid.z, block.x, block.y, block.z, smem_bytes);
    const unsigned long long nblocks = (unsigned long long)grid.x * grid.y * grid.z;
    std::vector<unsigned> border;
    if (g.sched_seed && nblocks <= (1u << 24)) {          // blocks in random order too: nothing may depend on block i running before block j
        border.resize((size_t)nblocks);
        for (unsigned i = 0; i < nblocks; ++i) border[i] = i;
        for (unsigned i = (unsigned)nblocks; i > 1; --i) {
            g.sched_state = g.sched_state * 1664525u + 1013904223u;
            std::swap(border[i - 1], border[(g.sched_state >> 8) % i]);
        }
    }
    for (unsigned long long b = 0; b < nblocks; ++b) {
        const unsigned long long lin = border.empty() ? b : border[(size_t)b];
        g.bid = dim3((unsigned)(lin % grid.x), (unsigned)((lin / grid.x) % grid.y), (unsigned)(lin / ((unsigned long long)grid.x * grid.y)));
        for (const DynSmem& d : dyn_smem_arrays()) memset(d.p, 0xff, std::min(d.bytes, smem_bytes));   // NaN pattern: shared memory starts undefined
        run_block();
        if (g.err) return;
    }
}

}  // namespace simt

#define SIMT_DEFINE_DYN_SMEM(type, name)                                             \
    alignas(1024) thread_local type name[simt::kDynSmemMax / sizeof(type)];          \
    static simt::DynSmemRegistrar simt_reg_##name(name, sizeof(name));

#define SIMT_LAUNCH(kernel, grid, block, smem, stream, ...) \
    simt::launch(dim3(grid), dim3(block), (size_t)(smem), [&]() { kernel(__VA_ARGS__); })
