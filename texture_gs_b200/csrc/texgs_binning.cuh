// texgs_binning.cuh — tile binning without a global sort (spec E4).
//
// The reference lineage sorts all K (tile | depth) 64-bit keys with one device-wide radix sort
// (6-8 passes over 12 B*K). Here tiles are bins from the start:
//   1. preprocess counts pairs per tile (atomics on tile_count[T])
//   2. one CTA scans tile_count -> tile_offset (T ~ 8k..32k: one block is enough), publishes K
//   3. scatter: every visible Gaussian claims a slot in each of its tiles' segments and writes
//      {id, depth bits} there (unordered inside the segment)
//   4. one CTA per tile sorts its segment by the 64-bit key (depth bits << 32 | id) in shared
//      memory (bitonic network) and emits the sorted id list.
// Total traffic: 8 B*K written + read once, 4 B*K written — vs ~150 B*K for the radix sort.
// Order is exactly (tile, depth, Gaussian index) ascending, i.e. what a stable radix sort of
// index-ordered duplicates produces.
#pragma once
#include "texgs_common.cuh"

#define TEXGS_SCAN_THREADS 1024

// One CTA, ONE block-wide scan: every thread owns ceil(T / 1024) consecutive tiles, scans them serially, the 1024 thread
// totals are scanned with shuffles (two barriers in all; the round-1 version ran 8 block-wide scans with 4 barriers each
// for the 8160 tiles of a 1080p view: 13 us of pure latency).
__global__ void __launch_bounds__(TEXGS_SCAN_THREADS) texgs_scan_tiles(const RasterParams p) {
    __shared__ unsigned warp_tot[32];
    __shared__ unsigned smax[32];
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    const int per = (p.num_tiles + TEXGS_SCAN_THREADS - 1) / TEXGS_SCAN_THREADS;
    const int i0 = min(tid * per, p.num_tiles), i1 = min(i0 + per, p.num_tiles);
    unsigned sum = 0, vmax = 0;
    for (int b = i0; b < i1; b += 8) {                 // 8 independent loads in flight, not a chain of dependent ones
        unsigned c[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) c[j] = (b + j < i1) ? p.tile_count[b + j] : 0u;
#pragma unroll
        for (int j = 0; j < 8; ++j) { sum += c[j]; vmax = max(vmax, c[j]); }
    }
    unsigned x = sum;                                  // inclusive scan of the thread totals inside the warp
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const unsigned y = __shfl_up_sync(0xffffffffu, x, o);
        if (lane >= o) x += y;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) vmax = max(vmax, __shfl_xor_sync(0xffffffffu, vmax, o));
    if (lane == 31) warp_tot[wid] = x;
    if (lane == 0) smax[wid] = vmax;
    __syncthreads();
    if (wid == 0) {
        unsigned w = warp_tot[lane];
        unsigned m = smax[lane];
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const unsigned y = __shfl_up_sync(0xffffffffu, w, o);
            if (lane >= o) w += y;
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) m = max(m, __shfl_xor_sync(0xffffffffu, m, o));
        warp_tot[lane] = w;                            // inclusive over warps
        if (lane == 31) {
            const unsigned K = w;
            p.tile_offset[p.num_tiles] = K;
            p.counters->num_pairs = K;
            p.counters->max_tile_len = m;
            p.counters->overflow = ((unsigned long long)K > p.pair_capacity) ? 1u : 0u;
        }
    }
    __syncthreads();
    unsigned run = ((wid == 0) ? 0u : warp_tot[wid - 1]) + x - sum;     // exclusive prefix of this thread's first tile
    for (int b = i0; b < i1; b += 8) {
        unsigned c[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) c[j] = (b + j < i1) ? p.tile_count[b + j] : 0u;
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            if (b + j < i1) p.tile_offset[b + j] = run;
            run += c[j];
        }
    }
}

__global__ void __launch_bounds__(256) texgs_scatter_pairs(const RasterParams p) {
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= p.P) return;
    if (p.counters->overflow) return;
    const uint2 rc = p.rects[idx];
    const int x0 = rc.x & 0xffff, x1 = rc.x >> 16, y0 = rc.y & 0xffff, y1 = rc.y >> 16;
    if (x1 <= x0 || y1 <= y0) return;
    const float4 r0 = p.recs[idx].q[0], r1 = p.recs[idx].q[1];
    const unsigned depth_bits = __float_as_uint(r1.z);
    for (int ty = y0; ty < y1; ++ty)
        for (int tx = x0; tx < x1; ++tx) {
            if (!splat_hits_tile(r0.x, r0.y, r0.z, r0.w, r1.x, r1.y, tx, ty)) continue;   // same test as the count
            const int t = ty * p.grid_x + tx;
            const unsigned slot = p.tile_offset[t] + atomicAdd(&p.tile_cursor[t], 1u);
            p.pairs[slot] = make_uint2((unsigned)idx, depth_bits);
        }
}

// Per-tile sort. Keys are 64-bit (depth bits << 32 | id); +inf padding = all ones.
#define TEXGS_SORT_THREADS 256
#define TEXGS_SORT_SMEM_ELEMS 4096   // 32 KB of u64

// Short lists (the common case: a few hundred entries) are sorted by 64-thread CTAs with 4 KB of
// shared memory — many more of them are resident per SM and their barriers cost two warps, not eight.
#define TEXGS_SORT_SMALL 512
#define TEXGS_SORT_SMALL_THREADS 64

__global__ void __launch_bounds__(TEXGS_SORT_SMALL_THREADS) texgs_sort_tiles_small(const RasterParams p) {
    __shared__ unsigned long long keys[TEXGS_SORT_SMALL];
    if (p.counters->overflow) return;
    const int tile = blockIdx.x;
    const unsigned start = p.tile_offset[tile];
    const unsigned n = p.tile_offset[tile + 1] - start;
    if (n > TEXGS_SORT_SMALL) {                             // long list: queued for texgs_sort_tiles (the cursors of the
        if (threadIdx.x == 0)                               // scatter kernel are dead by now: they hold the queue)
            p.tile_cursor[atomicAdd(&p.counters->num_long_tiles, 1u)] = (unsigned)tile;
        return;
    }
    if (n == 0) return;
    unsigned long long* seg = reinterpret_cast<unsigned long long*>(p.pairs + start);
    const unsigned tid = threadIdx.x;
    unsigned npad = 2;
    while (npad < n) npad <<= 1;
    for (unsigned i = tid; i < npad; i += TEXGS_SORT_SMALL_THREADS) keys[i] = (i < n) ? seg[i] : ~0ull;
    __syncthreads();
    // k, j are powers of two: index arithmetic by shifts / masks (lk = log2 k, lj = log2 j)
    for (unsigned k = 2, lk = 1; k <= npad; k <<= 1, ++lk) {
        for (unsigned i = tid; i < npad / 2; i += TEXGS_SORT_SMALL_THREADS) {
            const unsigned blk = i >> (lk - 1), off = i & ((k >> 1) - 1);
            const unsigned lo = (blk << lk) + off, hi = (blk << lk) + (k - 1 - off);
            const unsigned long long a = keys[lo], b = keys[hi];
            if (a > b) { keys[lo] = b; keys[hi] = a; }
        }
        __syncthreads();
        for (unsigned j = k >> 2, lj = lk - 2; j > 0; j >>= 1, --lj) {
            for (unsigned i = tid; i < npad / 2; i += TEXGS_SORT_SMALL_THREADS) {
                const unsigned lo = ((i >> lj) << (lj + 1)) + (i & (j - 1)), hi = lo + j;
                const unsigned long long a = keys[lo], b = keys[hi];
                if (a > b) { keys[lo] = b; keys[hi] = a; }
            }
            __syncthreads();
        }
    }
    for (unsigned i = tid; i < n; i += TEXGS_SORT_SMALL_THREADS) {
        const unsigned long long kv = keys[i];
        seg[i] = kv;
        p.sorted_ids[start + i] = (unsigned)(kv & 0xffffffffull);
    }
}

// Long lists (> 512 entries; a handful of tiles per view): a small fixed grid walks the queue the small-list kernel filled.
// (Round 1 launched one 256-thread CTA per tile here as well: 8160 CTAs of which a few had work, 26 us.)
#define TEXGS_SORT_LONG_CTAS 296
__global__ void __launch_bounds__(TEXGS_SORT_THREADS) texgs_sort_tiles(const RasterParams p) {
    __shared__ unsigned long long keys[TEXGS_SORT_SMEM_ELEMS];
    if (p.counters->overflow) return;
    const unsigned nlong = p.counters->num_long_tiles;
    const int tid = threadIdx.x;
    for (unsigned q = blockIdx.x; q < nlong; q += gridDim.x) {
    const int tile = (int)p.tile_cursor[q];
    const unsigned start = p.tile_offset[tile];
    const unsigned n = p.tile_offset[tile + 1] - start;
    unsigned long long* seg = reinterpret_cast<unsigned long long*>(p.pairs + start);
    unsigned npad = 2;
    while (npad < n) npad <<= 1;
    if (npad <= TEXGS_SORT_SMEM_ELEMS) {
        for (unsigned i = tid; i < npad; i += TEXGS_SORT_THREADS) keys[i] = (i < n) ? seg[i] : ~0ull;
        __syncthreads();
        // bitonic network, all compare-exchanges ascending (flip-merge formulation)
        for (unsigned k = 2; k <= npad; k <<= 1) {
            // first step of the merge: partner = i ^ (k - 1)
            for (unsigned i = tid; i < npad / 2; i += TEXGS_SORT_THREADS) {
                const unsigned blk = i / (k / 2), off = i % (k / 2);
                const unsigned lo = blk * k + off, hi = blk * k + (k - 1 - off);
                unsigned long long a = keys[lo], b = keys[hi];
                if (a > b) { keys[lo] = b; keys[hi] = a; }
            }
            __syncthreads();
            for (unsigned j = k / 4; j > 0; j >>= 1) {
                for (unsigned i = tid; i < npad / 2; i += TEXGS_SORT_THREADS) {
                    const unsigned lo = 2 * j * (i / j) + (i % j), hi = lo + j;
                    unsigned long long a = keys[lo], b = keys[hi];
                    if (a > b) { keys[lo] = b; keys[hi] = a; }
                }
                __syncthreads();
            }
        }
        for (unsigned i = tid; i < n; i += TEXGS_SORT_THREADS) {
            const unsigned long long kv = keys[i];
            seg[i] = kv;
            p.sorted_ids[start + i] = (unsigned)(kv & 0xffffffffull);
        }
    } else {
        // rare: list longer than the shared-memory capacity -> same network on global memory with
        // virtual +inf padding (exchanges touching an index >= n are no-ops)
        for (unsigned k = 2; k <= npad; k <<= 1) {
            for (unsigned i = tid; i < npad / 2; i += TEXGS_SORT_THREADS) {
                const unsigned blk = i / (k / 2), off = i % (k / 2);
                const unsigned lo = blk * k + off, hi = blk * k + (k - 1 - off);
                if (hi < n) {
                    unsigned long long a = seg[lo], b = seg[hi];
                    if (a > b) { seg[lo] = b; seg[hi] = a; }
                }
            }
            __syncthreads();
            for (unsigned j = k / 4; j > 0; j >>= 1) {
                for (unsigned i = tid; i < npad / 2; i += TEXGS_SORT_THREADS) {
                    const unsigned lo = 2 * j * (i / j) + (i % j), hi = lo + j;
                    if (hi < n) {
                        unsigned long long a = seg[lo], b = seg[hi];
                        if (a > b) { seg[lo] = b; seg[hi] = a; }
                    }
                }
                __syncthreads();
            }
        }
        for (unsigned i = tid; i < n; i += TEXGS_SORT_THREADS) p.sorted_ids[start + i] = (unsigned)(seg[i] & 0xffffffffull);
    }
    __syncthreads();                                        // ``keys`` is reused by the next queue entry
    }
}
