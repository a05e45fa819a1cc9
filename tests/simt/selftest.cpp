// selftest.cpp — the SIMT emulator checked against kernels with known answers and kernels with known BUGS
// (tests/test_simt_emulator_cpu.py compiles and runs this; exit code 0 and "ok" lines expected).
#include "simt_emu.h"

SIMT_DEFINE_DYN_SMEM(unsigned char, dyn_raw)

static int failures = 0;
#define EXPECT(cond, what)                                                  \
    do {                                                                    \
        if (cond) printf("ok   %s\n", what);                                \
        else { printf("FAIL %s (%s)\n", what, simt::G().err_msg.c_str()); ++failures; } \
    } while (0)

static bool take_error(const char* needle) {
    simt::Global& g = simt::G();
    const bool hit = g.err != 0 && g.err_msg.find(needle) != std::string::npos;
    g.err = 0;
    return hit;
}

// ---- known answers -------------------------------------------------------------------------------------------------
__global__ void k_collectives(int* out) {
    __shared__ int s[96];
    const int tid = threadIdx.x, lane = tid & 31;
    s[tid] = tid;
    __syncthreads();
    int v = s[(tid + 1) % 96];
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);          // warp sum
    const unsigned b = __ballot_sync(0xffffffffu, lane % 3 == 0);
    const int up = __shfl_up_sync(0xffffffffu, lane, 1), dn = __shfl_down_sync(0xffffffffu, lane, 2);
    const int bc = __shfl_sync(0xffffffffu, lane * 7, 5);
    const bool any = __any_sync(0xffffffffu, lane == 31), all = __all_sync(0xffffffffu, lane < 31);
    const double dsum = __shfl_xor_sync(0xffffffffu, (double)lane + 0.5, 1);
    if (lane == 9) {
        int* o = out + 8 * (tid >> 5);
        o[0] = v; o[1] = (int)b; o[2] = up; o[3] = dn; o[4] = bc; o[5] = any; o[6] = all; o[7] = (int)(2.0 * dsum);
    }
}

__global__ void k_blockidx(unsigned* out) {
    const unsigned id = (blockIdx.z * gridDim.y + blockIdx.y) * gridDim.x + blockIdx.x;
    const unsigned t = (threadIdx.z * blockDim.y + threadIdx.y) * blockDim.x + threadIdx.x;
    atomicAdd(out + id, t + 1);
    if (t == 0) atomicMax(out + gridDim.x * gridDim.y * gridDim.z, id);
}

// two-stage bulk-copy ring driven by one warp, the pattern of texgs_render.cuh
__global__ void k_ring(const float4* src, float* out, int nchunks) {
    extern __shared__ __align__(128) unsigned char dyn_raw[];
    float4* stage = reinterpret_cast<float4*>(dyn_raw);            // 2 x 32 float4
    uint64_t* bar = reinterpret_cast<uint64_t*>(dyn_raw + 2 * 32 * sizeof(float4));
    const int lane = threadIdx.x;
    if (lane == 0) { mbar_init(&bar[0], 1); mbar_init(&bar[1], 1); mbar_fence_init(); }
    __syncwarp();
    auto issue = [&](int c) {
        const int s = c & 1;
        if (lane == 0) mbar_arrive_expect_tx(&bar[s], 32 * sizeof(float4));
        bulk_g2s(&stage[s * 32 + lane], src + c * 32 + lane, sizeof(float4), &bar[s]);
    };
    issue(0);
    if (nchunks > 1) issue(1);
    float acc = 0.f;
    for (int c = 0; c < nchunks; ++c) {
        const int s = c & 1;
        mbar_wait(&bar[s], (unsigned)(c >> 1) & 1u);
        __syncwarp();
        const float4 v = stage[s * 32 + ((lane + 1) & 31)];
        acc += v.x + v.y + v.z + v.w;
        __syncwarp();
        if (c + 2 < nchunks) issue(c + 2);
    }
    out[lane] = acc;
}

// ---- known bugs ----------------------------------------------------------------------------------------------------
__global__ void bug_divergent_barrier() {
    if (threadIdx.x < 40) __syncthreads();
    else if (threadIdx.x >= 64) __syncthreads();
    // threads 40..63 never arrive, but they do not exit either: they spin in a second barrier nobody else reaches
    if (threadIdx.x >= 40 && threadIdx.x < 64) { __shfl_xor_sync(0xffffffffu, 1, 1); }
}
__global__ void bug_shuffle_after_exit(int* out) {
    if (threadIdx.x == 3) return;
    out[threadIdx.x] = __shfl_xor_sync(0xffffffffu, (int)threadIdx.x, 1);
}
__global__ void bug_read_before_wait(const float4* src, float* out) {
    extern __shared__ __align__(128) unsigned char dyn_raw[];
    float4* stage = reinterpret_cast<float4*>(dyn_raw);
    uint64_t* bar = reinterpret_cast<uint64_t*>(dyn_raw + 32 * sizeof(float4));
    const int lane = threadIdx.x;
    if (lane == 0) { mbar_init(bar, 1); mbar_fence_init(); }
    __syncwarp();
    if (lane == 0) mbar_arrive_expect_tx(bar, 32 * sizeof(float4));
    bulk_g2s(&stage[lane], src + lane, sizeof(float4), bar);
    __syncwarp();
    out[lane] = stage[lane].x;        // BUG: no mbar_wait — real hardware may or may not have the data yet
    mbar_wait(bar, 0);
}
__global__ void bug_exit_with_copy_in_flight(const float4* src) {
    extern __shared__ __align__(128) unsigned char dyn_raw[];
    float4* stage = reinterpret_cast<float4*>(dyn_raw);
    uint64_t* bar = reinterpret_cast<uint64_t*>(dyn_raw + 32 * sizeof(float4));
    const int lane = threadIdx.x;
    if (lane == 0) { mbar_init(bar, 1); mbar_fence_init(); }
    __syncwarp();
    if (lane == 0) mbar_arrive_expect_tx(bar, 32 * sizeof(float4));
    bulk_g2s(&stage[lane], src + lane, sizeof(float4), bar);
}
__global__ void bug_wait_wrong_byte_count(const float4* src) {
    extern __shared__ __align__(128) unsigned char dyn_raw[];
    float4* stage = reinterpret_cast<float4*>(dyn_raw);
    uint64_t* bar = reinterpret_cast<uint64_t*>(dyn_raw + 32 * sizeof(float4));
    const int lane = threadIdx.x;
    if (lane == 0) { mbar_init(bar, 1); mbar_fence_init(); }
    __syncwarp();
    if (lane == 0) mbar_arrive_expect_tx(bar, 33 * sizeof(float4));       // BUG: one record more than is copied
    bulk_g2s(&stage[lane], src + lane, sizeof(float4), bar);
    mbar_wait(bar, 0);
}

// stage re-filled while other lanes may still read its previous content (missing __syncwarp before the re-issue)
__global__ void bug_overwrite_while_reading(const float4* src, float* out) {
    extern __shared__ __align__(128) unsigned char dyn_raw[];
    float4* stage = reinterpret_cast<float4*>(dyn_raw);
    uint64_t* bar = reinterpret_cast<uint64_t*>(dyn_raw + 32 * sizeof(float4));
    const int lane = threadIdx.x;
    if (lane == 0) { mbar_init(bar, 1); mbar_fence_init(); }
    __syncwarp();
    if (lane == 0) mbar_arrive_expect_tx(bar, 32 * sizeof(float4));
    bulk_g2s(&stage[lane], src + lane, sizeof(float4), bar);
    mbar_wait(bar, 0);
    __syncwarp();
    const float mine = stage[(lane + 1) & 31].x;
    // BUG: no __syncwarp() here — the neighbour's slot may be overwritten before this lane has read it
    if (lane == 0) mbar_arrive_expect_tx(bar, 32 * sizeof(float4));
    bulk_g2s(&stage[lane], src + 32 + lane, sizeof(float4), bar);
    out[lane] = mine;
    mbar_wait(bar, 1);
}

// inter-warp race: warp 1 reads what warp 0 writes without a block barrier in between
__global__ void bug_missing_syncthreads(int* out) {
    __shared__ int s[64];
    const int tid = threadIdx.x;
    s[tid] = 0;
    __syncthreads();
    for (int it = 0; it < 8; ++it) {
        s[tid] = it + 1;
        __syncwarp();                       // keeps every warp in step with itself, not with the other warp
        // BUG: no __syncthreads() before reading the other warp's slot
        if (s[tid ^ 32] != it + 1) atomicAdd(out, 1);
        __syncthreads();
    }
}

int main() {
    {
        int out[24] = {0};
        SIMT_LAUNCH((k_collectives), 1, 96, 0, 0, out);
        bool ok = simt::G().err == 0;
        for (int w = 0; w < 3; ++w) {
            int sum = 0;
            for (int l = 0; l < 32; ++l) sum += (w * 32 + l + 1) % 96;
            const int* o = out + 8 * w;
            ok = ok && o[0] == sum && (unsigned)o[1] == 0x49249249u && o[2] == 8 && o[3] == 11 && o[4] == 35 && o[5] == 1 && o[6] == 0 &&
                 o[7] == 2 * 8 + 1;
        }
        EXPECT(ok, "shuffles, ballot, any/all, 64-bit payloads, block barrier over three warps");
    }
    {
        unsigned out[2 * 3 * 4 + 1] = {0};
        SIMT_LAUNCH((k_blockidx), dim3(2, 3, 4), dim3(4, 2, 3), 0, 0, out);
        bool ok = simt::G().err == 0 && out[24] == 23;
        for (int i = 0; i < 24; ++i) ok = ok && out[i] == 24 * 25 / 2;
        EXPECT(ok, "3-D grid and block indices, atomics");
    }
    alignas(16) static float4 src[5 * 32];
    for (int i = 0; i < 5 * 32; ++i) src[i] = make_float4((float)i, 1.f, 2.f, 3.f);
    {
        float out[32];
        SIMT_LAUNCH((k_ring), 1, 32, 2 * 32 * sizeof(float4) + 16, 0, src, out, 5);
        bool ok = simt::G().err == 0;
        for (int l = 0; l < 32; ++l) {
            float e = 0.f;
            for (int c = 0; c < 5; ++c) e += (float)(c * 32 + ((l + 1) & 31)) + 6.f;
            ok = ok && out[l] == e;
        }
        EXPECT(ok, "two-stage mbarrier / bulk-copy ring with phase parity over five chunks");
    }
    SIMT_LAUNCH((bug_divergent_barrier), 1, 96, 0, 0);
    EXPECT(take_error("deadlock"), "divergent __syncthreads is reported as a deadlock");
    {
        int out[32] = {0};
        SIMT_LAUNCH((bug_shuffle_after_exit), 1, 32, 0, 0, out);
        EXPECT(take_error("exited"), "full-mask shuffle after a lane returned is reported");
    }
    {
        float out[32];
        SIMT_LAUNCH((bug_read_before_wait), 1, 32, 32 * sizeof(float4) + 8, 0, src, out);
        bool stale = simt::G().err == 0;
        for (int l = 0; l < 32; ++l) stale = stale && out[l] != src[l].x;          // NaN pattern, not the data
        if (!stale) printf("     err=%d out[0]=%f out[31]=%f\n", simt::G().err, out[0], out[31]);
        EXPECT(stale, "shared memory read before the mbarrier wait sees stale (poisoned) data");
    }
    SIMT_LAUNCH((bug_exit_with_copy_in_flight), 1, 32, 32 * sizeof(float4) + 8, 0, src);
    EXPECT(take_error("in flight"), "block exit with a bulk copy in flight is reported");
    SIMT_LAUNCH((bug_wait_wrong_byte_count), 1, 32, 32 * sizeof(float4) + 8, 0, src);
    EXPECT(take_error("deadlock"), "expect_tx byte count that the copies never deliver is reported as a deadlock");
    for (int eager = 0; eager < 2; ++eager) {
        float out[32];
        simt::G().eager_copies = eager != 0;
        SIMT_LAUNCH((bug_overwrite_while_reading), 1, 32, 32 * sizeof(float4) + 8, 0, src, out);
        int wrong = 0;
        for (int l = 0; l < 32; ++l) wrong += out[l] != src[(l + 1) & 31].x;
        if (!eager) EXPECT(simt::G().err == 0 && wrong == 0, "late-landing copies hide an overwrite-while-reading hazard ...");
        else EXPECT(simt::G().err == 0 && wrong > 0, "... early-landing copies (eager mode) expose it");
        simt::G().eager_copies = false;
    }
    {
        int stale_det = 0, stale_rnd = 0;
        SIMT_LAUNCH((bug_missing_syncthreads), 1, 64, 0, 0, &stale_det);
        for (unsigned seed = 1; seed <= 8; ++seed) {
            simt::G().sched_seed = seed; simt::G().sched_state = seed * 2654435761u + 1u;
            SIMT_LAUNCH((bug_missing_syncthreads), 1, 64, 0, 0, &stale_rnd);
        }
        simt::G().sched_seed = 0;
        EXPECT(simt::G().err == 0 && stale_det > 0 && stale_rnd > 0, "a missing __syncthreads between two warps is observed (fixed and random schedules)");
    }
    SIMT_LAUNCH((k_blockidx), 1, 2048, 0, 0, (unsigned*)nullptr);
    EXPECT(take_error("invalid configuration"), "more than 1024 threads per block is refused");
    printf("%s\n", failures ? "SELFTEST FAILED" : "SELFTEST PASSED");
    return failures ? 1 : 0;
}
