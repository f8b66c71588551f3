// hm.cu -- histogram matching on B200: 256-bin per-channel histogram of the whole batch (H3),
// bit-exact LUT construction (H1/H2) and the streaming LUT remap (H4).
//
// Reference semantics: src/stainx/backends/torch_backend.py:L139-141, L194-301 (torch CPU oracle);
// the kernels it replaces: csrc/histogram_matching.cu:L21-226 and the ATen glue in
// src/stainx_cuda_torch/csrc/histogram_matching.cu:L25-169.
//
// Data path (uint8 NCHW, the BASELINE config): 3 B/px read for the histogram, 3 B/px read +
// 3 B/px written for the remap = 9 algorithmic bytes per pixel, all 128-bit coalesced.
#include <type_traits>

#include "common.cuh"

namespace sx {
namespace hm {

constexpr int kThreads = 256;
constexpr int kWarps = kThreads / 32;
constexpr int kUnroll = 4;                            // 128-bit loads in flight per thread
constexpr int kTileVecs = kThreads * kUnroll;         // uint4 / float4 vectors per tile
constexpr int kFlushTiles = 3;                        // 3 tiles * 4 loads * 16 B = 192 <= 255 per counter

// ------------------------------------------------------------------------------------------------
// Histogram, planar uint8.  grid = (blocks, 3 channels); a CTA only ever sees one channel.
//
// Counting scheme: every THREAD owns a private 256-bin histogram of 8-bit counters in shared
// memory, so counting is a plain byte load / add / store -- no atomics and no bank conflicts:
// the counter of (bin, lane) lives in word (bin >> 2) * 32 + lane, byte (bin & 3), i.e. lane l only
// ever touches bank l.  Byte counters overflow after 255 hits, so each warp folds its 8 KB region
// into the CTA's 32-bit histogram every kFlushTiles tiles (<= 192 values per thread).
// ------------------------------------------------------------------------------------------------
struct ByteCounters {
    // warp-private region: 64 rows (bin >> 2) x 32 lanes x 4 byte-counters
    static constexpr int kBytesPerWarp = 64 * 32 * 4;

    unsigned char *mine;   // &region[lane * 4]
    unsigned int *region;  // warp region as words
    unsigned int *hist32;  // CTA histogram (256 x u32)
    int lane;

    __device__ __forceinline__ void init(unsigned char *smem_counters, unsigned int *h32) {
        int warp = threadIdx.x >> 5;
        lane = threadIdx.x & 31;
        region = reinterpret_cast<unsigned int *>(smem_counters + warp * kBytesPerWarp);
        mine = reinterpret_cast<unsigned char *>(region) + lane * 4;
        hist32 = h32;
        for (int i = lane; i < 64 * 32; i += 32) region[i] = 0u;
        __syncwarp();
    }
    __device__ __forceinline__ void add(unsigned b) {  // b in [0,255]
        unsigned off = ((b & 0xfcu) << 5) | (b & 3u);
        mine[off] = (unsigned char)(mine[off] + 1);
    }
    __device__ __forceinline__ void add4(unsigned w) {
        add(w & 0xffu);
        add((w >> 8) & 0xffu);
        add((w >> 16) & 0xffu);
        add(w >> 24);
    }
    // Fold the warp's byte counters into hist32 and clear them.  Lane j sums rows j and j + 32,
    // walking the 32 words of a row with a lane-dependent rotation (bank = (k + lane) & 31).
    __device__ __forceinline__ void flush() {
        __syncwarp();
#pragma unroll
        for (int rr = 0; rr < 2; ++rr) {
            int q = lane + 32 * rr;
            unsigned lo = 0, hi = 0;  // two 16-bit partial sums each
#pragma unroll 8
            for (int k = 0; k < 32; ++k) {
                int col = (k + lane) & 31;
                unsigned w = region[q * 32 + col];
                region[q * 32 + col] = 0u;
                lo += w & 0x00ff00ffu;
                hi += (w >> 8) & 0x00ff00ffu;
            }
            if (lo & 0xffffu) atomicAdd(&hist32[4 * q + 0], lo & 0xffffu);
            if (hi & 0xffffu) atomicAdd(&hist32[4 * q + 1], hi & 0xffffu);
            if (lo >> 16) atomicAdd(&hist32[4 * q + 2], lo >> 16);
            if (hi >> 16) atomicAdd(&hist32[4 * q + 3], hi >> 16);
        }
        __syncwarp();
    }
};

// Counting scheme C: the same lane-private packed layout, but the update is a fire-and-forget
// shared-memory reduction (RED.ADD of 1 << 8*(bin & 3)) instead of a byte load/add/store.  Lane l
// only touches bank l, so the warp-wide RED is conflict-free, and nothing waits on its result.
// A byte lane holds at most 255 hits between flushes, so a carry can never cross into its
// neighbour.
struct PackedRed {
    static constexpr int kBytesPerWarp = 64 * 32 * 4;
    unsigned int *mine;    // &region[lane]
    unsigned int *region;
    unsigned int *hist32;
    int lane;
    __device__ __forceinline__ void init(unsigned char *smem_counters, unsigned int *h32) {
        int warp = threadIdx.x >> 5;
        lane = threadIdx.x & 31;
        region = reinterpret_cast<unsigned int *>(smem_counters + warp * kBytesPerWarp);
        mine = region + lane;
        hist32 = h32;
        for (int i = lane; i < 64 * 32; i += 32) region[i] = 0u;
        __syncwarp();
    }
    __device__ __forceinline__ void add(unsigned b) {
        atomicAdd(mine + ((b & 0xfcu) << 3), 1u << ((b & 3u) << 3));  // word (b>>2)*32 + lane
    }
    __device__ __forceinline__ void add4(unsigned w) {
        add(w & 0xffu);
        add((w >> 8) & 0xffu);
        add((w >> 16) & 0xffu);
        add(w >> 24);
    }
    __device__ __forceinline__ void flush() {
        __syncwarp();
#pragma unroll
        for (int rr = 0; rr < 2; ++rr) {
            int q = lane + 32 * rr;
            unsigned lo = 0, hi = 0;
#pragma unroll 8
            for (int k = 0; k < 32; ++k) {
                int col = (k + lane) & 31;
                unsigned w = region[q * 32 + col];
                region[q * 32 + col] = 0u;
                lo += w & 0x00ff00ffu;
                hi += (w >> 8) & 0x00ff00ffu;
            }
            if (lo & 0xffffu) atomicAdd(&hist32[4 * q + 0], lo & 0xffffu);
            if (hi & 0xffffu) atomicAdd(&hist32[4 * q + 1], hi & 0xffffu);
            if (lo >> 16) atomicAdd(&hist32[4 * q + 2], lo >> 16);
            if (hi >> 16) atomicAdd(&hist32[4 * q + 3], hi >> 16);
        }
        __syncwarp();
    }
};

// Alternative counting scheme (selectable for A/B measurements): warp-private 32-bit histograms
// updated with shared-memory atomics.
struct WarpAtomics {
    unsigned int *wh;  // warp histogram (256 x u32)
    __device__ __forceinline__ void init(unsigned int *smem_hist) {
        wh = smem_hist + (threadIdx.x >> 5) * 256;
        for (int i = threadIdx.x & 31; i < 256; i += 32) wh[i] = 0u;
        __syncwarp();
    }
    __device__ __forceinline__ void add(unsigned b) { atomicAdd(&wh[b], 1u); }
    __device__ __forceinline__ void add4(unsigned w) {
        add(w & 0xffu);
        add((w >> 8) & 0xffu);
        add((w >> 16) & 0xffu);
        add(w >> 24);
    }
};

// Work decomposition shared by the planar kernels: a plane of `len` elements starting at `p` is
// split into an unaligned head (< 16 B), a body of 16-byte vectors and a tail.
struct PlaneSplit {
    int64_t head;   // elements before the first aligned vector
    int64_t nvec;   // aligned vectors
    int64_t tail0;  // first element after the body
};
template <typename T>
__device__ __forceinline__ PlaneSplit split_plane(const T *p, int64_t len) {
    constexpr int kPerVec = 16 / sizeof(T);
    uintptr_t a = reinterpret_cast<uintptr_t>(p);
    int64_t head = (int64_t)(((16 - (a & 15)) & 15) / sizeof(T));
    if (head > len) head = len;
    PlaneSplit s;
    s.head = head;
    s.nvec = (len - head) / kPerVec;
    s.tail0 = head + s.nvec * kPerVec;
    return s;
}

template <int MODE>  // 0: warp-private shared atomics, 1: byte counters, 2: packed lane-private RED
__global__ void __launch_bounds__(kThreads) hist_u8_planar_kernel(const uint8_t *__restrict__ img, int64_t hw, int64_t n_img, int64_t tiles_per_plane, unsigned long long *__restrict__ counts) {
    extern __shared__ __align__(16) unsigned char smem[];
    unsigned int *hist32 = reinterpret_cast<unsigned int *>(smem);  // 256 words
    unsigned char *scratch = smem + 256 * sizeof(unsigned int);
    const int c = blockIdx.y;
    for (int i = threadIdx.x; i < 256; i += kThreads) hist32[i] = 0u;

    constexpr bool BYTE_COUNTERS = MODE != 0;  // modes 1 and 2 need the periodic flush
    using Packed = typename std::conditional<MODE == 2, PackedRed, ByteCounters>::type;
    Packed bc;
    WarpAtomics wa;
    if (BYTE_COUNTERS) bc.init(scratch, hist32);
    else wa.init(reinterpret_cast<unsigned int *>(scratch));
    __syncthreads();

    const int64_t items = n_img * tiles_per_plane;
    int since_flush = 0;
    for (int64_t item = blockIdx.x; item < items; item += gridDim.x) {
        const int64_t n = item / tiles_per_plane;
        const int64_t t = item - n * tiles_per_plane;
        const uint8_t *plane = img + (n * 3 + c) * hw;
        const PlaneSplit sp = split_plane(plane, hw);
        const uint4 *body = reinterpret_cast<const uint4 *>(plane + sp.head);
        const int64_t v0 = t * kTileVecs;
        uint4 v[kUnroll];
        bool ok[kUnroll];
#pragma unroll
        for (int u = 0; u < kUnroll; ++u) {
            int64_t vi = v0 + u * kThreads + threadIdx.x;
            ok[u] = vi < sp.nvec;
            if (ok[u]) v[u] = ld_stream(body + vi);
        }
#pragma unroll
        for (int u = 0; u < kUnroll; ++u) {
            if (ok[u]) {
                if (BYTE_COUNTERS) { bc.add4(v[u].x); bc.add4(v[u].y); bc.add4(v[u].z); bc.add4(v[u].w); }
                else { wa.add4(v[u].x); wa.add4(v[u].y); wa.add4(v[u].z); wa.add4(v[u].w); }
            }
        }
        if (t == 0) {  // ragged ends of the plane: < 32 bytes, one thread each
            int64_t ragged = sp.head + (hw - sp.tail0);
            if ((int64_t)threadIdx.x < ragged) {
                int64_t idx = (int64_t)threadIdx.x < sp.head ? (int64_t)threadIdx.x : sp.tail0 + ((int64_t)threadIdx.x - sp.head);
                if (BYTE_COUNTERS) bc.add(plane[idx]);
                else wa.add(plane[idx]);
            }
        }
        if (BYTE_COUNTERS && ++since_flush == kFlushTiles) {
            bc.flush();
            since_flush = 0;
        }
    }
    if (BYTE_COUNTERS) {
        bc.flush();
    } else {
        __syncwarp();
        for (int i = threadIdx.x & 31; i < 256; i += 32)
            if (wa.wh[i]) atomicAdd(&hist32[i], wa.wh[i]);
    }
    __syncthreads();
    for (int i = threadIdx.x; i < 256; i += kThreads)
        if (hist32[i]) atomicAdd(&counts[c * 256 + i], (unsigned long long)hist32[i]);
}

// ------------------------------------------------------------------------------------------------
// Histogram, planar uint8, scheme D ("lane32"): every LANE owns a full 256-bin histogram of 32-bit
// counters, laid out [bin][lane] inside a 32 KB warp region, so counter (bin, lane) sits in bank
// `lane`.  Counting one byte is SHF + LOP3 + a conflict-free fire-and-forget ATOMS.ADD; there is
// no overflow, hence no periodic flush.  7 warps x 32 KB fill the SM's shared memory, so the CTA
// is persistent (one per SM) and walks a CONTIGUOUS range of the (channel, image, tile) list,
// folding its histogram into the global counts whenever the channel changes (at most twice).
// Memory-level parallelism comes from 8 independent 128-bit loads per thread.
// ------------------------------------------------------------------------------------------------
constexpr int kL32Warps = 7;
constexpr int kL32Threads = kL32Warps * 32;
constexpr int kL32Unroll = 8;
constexpr int kL32TileVecs = kL32Threads * kL32Unroll;
constexpr int kL32SmemBytes = kL32Warps * 256 * 32 * 4 + 256 * 4;

__device__ __forceinline__ void l32_count4(unsigned w, unsigned lane_off, uint32_t region_addr) {
    // byte offset of counter (bin, lane) = bin * 128 + lane * 4
    unsigned o0 = ((w << 7) & 0x7f80u) | lane_off;
    unsigned o1 = ((w >> 1) & 0x7f80u) | lane_off;
    unsigned o2 = ((w >> 9) & 0x7f80u) | lane_off;
    unsigned o3 = ((w >> 17) & 0x7f80u) | lane_off;
    asm volatile("red.shared.add.u32 [%0], 1;" ::"r"(region_addr + o0) : "memory");
    asm volatile("red.shared.add.u32 [%0], 1;" ::"r"(region_addr + o1) : "memory");
    asm volatile("red.shared.add.u32 [%0], 1;" ::"r"(region_addr + o2) : "memory");
    asm volatile("red.shared.add.u32 [%0], 1;" ::"r"(region_addr + o3) : "memory");
}

__global__ void __launch_bounds__(kL32Threads, 1) hist_u8_planar_lane32_kernel(const uint8_t *__restrict__ img, int64_t hw, int64_t n_img, int64_t tiles_per_plane, unsigned long long *__restrict__ counts) {
    extern __shared__ __align__(16) unsigned char smem[];
    unsigned int *regions = reinterpret_cast<unsigned int *>(smem);                                  // [warp][bin][lane]
    unsigned int *hist32 = reinterpret_cast<unsigned int *>(smem + kL32Warps * 256 * 32 * 4);        // [bin]
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    unsigned int *region = regions + warp * 256 * 32;
    const uint32_t region_addr = (uint32_t)__cvta_generic_to_shared(region);
    const unsigned lane_off = (unsigned)lane * 4u;

    for (int i = threadIdx.x; i < kL32Warps * 256 * 32; i += kL32Threads) regions[i] = 0u;
    for (int i = threadIdx.x; i < 256; i += kL32Threads) hist32[i] = 0u;
    __syncthreads();

    // fold the lane-private counters of every warp into the global counts of channel c, re-zero
    auto fold = [&](int c) {
        __syncthreads();
        for (int bin = lane; bin < 256; bin += 32) {  // lane j: bins j, j+32, ...; rotated walk = no conflicts
            unsigned sum = 0;
#pragma unroll 8
            for (int k = 0; k < 32; ++k) {
                const int col = (k + lane) & 31;
                sum += region[bin * 32 + col];
                region[bin * 32 + col] = 0u;
            }
            if (sum) atomicAdd(&hist32[bin], sum);
        }
        __syncthreads();
        for (int i = threadIdx.x; i < 256; i += kL32Threads) {
            const unsigned v = hist32[i];
            if (v) atomicAdd(&counts[c * 256 + i], (unsigned long long)v);
            hist32[i] = 0u;
        }
        __syncthreads();
    };

    const int64_t per_channel = n_img * tiles_per_plane;
    const int64_t items = 3 * per_channel;
    const int64_t per_cta = (items + gridDim.x - 1) / gridDim.x;
    const int64_t first = (int64_t)blockIdx.x * per_cta;
    const int64_t last = first + per_cta < items ? first + per_cta : items;

    // Two register buffers: the 8 loads of tile i+1 are in flight while tile i is counted.
    uint4 va[kL32Unroll], vb[kL32Unroll];
    unsigned oka = 0, okb = 0;
    auto issue = [&](int64_t item, uint4(&v)[kL32Unroll], unsigned &okmask) {
        const int c = (int)(item / per_channel);
        const int64_t rem = item - (int64_t)c * per_channel;
        const int64_t n = rem / tiles_per_plane;
        const int64_t t = rem - n * tiles_per_plane;
        const uint8_t *plane = img + (n * 3 + c) * hw;
        const PlaneSplit sp = split_plane(plane, hw);
        const uint4 *body = reinterpret_cast<const uint4 *>(plane + sp.head);
        const int64_t v0 = t * kL32TileVecs;
        okmask = 0;
#pragma unroll
        for (int u = 0; u < kL32Unroll; ++u) {
            const int64_t vi = v0 + u * kL32Threads + threadIdx.x;
            if (vi < sp.nvec) {
                v[u] = ld_stream(body + vi);
                okmask |= 1u << u;
            }
        }
    };
    auto count = [&](int64_t item, const uint4(&v)[kL32Unroll], unsigned okmask) {
#pragma unroll
        for (int u = 0; u < kL32Unroll; ++u) {
            if (okmask & (1u << u)) {
                l32_count4(v[u].x, lane_off, region_addr);
                l32_count4(v[u].y, lane_off, region_addr);
                l32_count4(v[u].z, lane_off, region_addr);
                l32_count4(v[u].w, lane_off, region_addr);
            }
        }
        const int c = (int)(item / per_channel);
        const int64_t rem = item - (int64_t)c * per_channel;
        const int64_t n = rem / tiles_per_plane;
        if (rem - n * tiles_per_plane == 0) {  // ragged ends of the plane
            const uint8_t *plane = img + (n * 3 + c) * hw;
            const PlaneSplit sp = split_plane(plane, hw);
            const int64_t ragged = sp.head + (hw - sp.tail0);
            if ((int64_t)threadIdx.x < ragged) {
                const int64_t idx = (int64_t)threadIdx.x < sp.head ? (int64_t)threadIdx.x : sp.tail0 + ((int64_t)threadIdx.x - sp.head);
                atomicAdd(&region[(unsigned)plane[idx] * 32 + lane], 1u);
            }
        }
    };

    int cur_c = -1;
    if (first < last) issue(first, va, oka);
    for (int64_t item = first; item < last; item += 2) {
        if (item + 1 < last) issue(item + 1, vb, okb);
        int c = (int)(item / per_channel);
        if (c != cur_c) {
            if (cur_c >= 0) fold(cur_c);
            cur_c = c;
        }
        count(item, va, oka);
        if (item + 1 < last) {
            if (item + 2 < last) issue(item + 2, va, oka);
            c = (int)((item + 1) / per_channel);
            if (c != cur_c) {
                fold(cur_c);
                cur_c = c;
            }
            count(item + 1, vb, okb);
        }
    }
    if (cur_c >= 0) fold(cur_c);
}

// ------------------------------------------------------------------------------------------------
// Histogram, planar uint8, scheme P ("pairs"): the shared-memory atomic unit retires a fixed number
// of lane-updates per clock whatever the address pattern (measured: ~6.5 per clk per SM), so the
// way to count faster is to count MORE PER UPDATE.  Two neighbouring bytes (a, b) of a plane are
// one 16-bit value h = a | b << 8; a single RED.ADD into a 65 536-cell table counts the pair, and
// the 256-bin histogram is the sum of the table's two marginals:
//     hist[v] = sum_b cell[v][b] + sum_a cell[a][v].
// Cells are 16-bit counters packed two per word (128 KB, one CTA of 1024 threads per SM): cell h
// lives in word h & 0x7fff, half h >> 15.  A cell overflows after 65 535 hits -- only possible when
// one byte pair makes up a large share of a CTA's segment (flat images).  A carry changes the sum
// over all cells, so the fold compares that sum with the number of pairs counted; on a mismatch the
// CTA discards the segment's table and recounts the segment with 32-bit warp-private atomics.
// ------------------------------------------------------------------------------------------------
constexpr int kPairThreads = 1024;
constexpr int kPairWarps = kPairThreads / 32;
constexpr int kPairUnroll = 2;                                  // 128-bit loads per thread per tile
constexpr int kPairTileVecs = kPairThreads * kPairUnroll;       // 32 KB of a plane per tile
constexpr int kPairTableWords = 32768;
constexpr int kPairSmemBytes = kPairTableWords * 4 + 256 * 4 + 64;

__device__ __forceinline__ void pair_count_word(unsigned w, uint32_t table_addr) {
    // low halfword: byte offset of its word = (h & 0x7fff) * 4, increment 1 or 1 << 16
    const unsigned o0 = (w << 2) & 0x1fffcu;
    const unsigned i0 = ((w >> 15) & 1u) * 0xffffu + 1u;
    const unsigned o1 = (w >> 14) & 0x1fffcu;
    const unsigned i1 = (w >> 31) * 0xffffu + 1u;
    asm volatile("red.shared.add.u32 [%0], %1;" ::"r"(table_addr + o0), "r"(i0) : "memory");
    asm volatile("red.shared.add.u32 [%0], %1;" ::"r"(table_addr + o1), "r"(i1) : "memory");
}

__global__ void __launch_bounds__(kPairThreads, 1) hist_u8_planar_pairs_kernel(const uint8_t *__restrict__ img, int64_t hw, int64_t n_img, int64_t tiles_per_plane, unsigned long long *__restrict__ counts) {
    extern __shared__ __align__(16) unsigned char smem[];
    unsigned int *table = reinterpret_cast<unsigned int *>(smem);              // [32768] packed 16-bit cells
    unsigned int *hist32 = table + kPairTableWords;                            // [256] segment histogram
    unsigned int *s_misc = hist32 + 256;                                       // [0] pairs counted, [1] sum of cells
    const uint32_t table_addr = (uint32_t)__cvta_generic_to_shared(table);
    const int lane = threadIdx.x & 31;

    for (int i = threadIdx.x; i < kPairTableWords; i += kPairThreads) table[i] = 0u;
    if (threadIdx.x < 256) hist32[threadIdx.x] = 0u;
    if (threadIdx.x < 2) s_misc[threadIdx.x] = 0u;
    __syncthreads();

    const int64_t per_channel = n_img * tiles_per_plane;
    const int64_t items = 3 * per_channel;
    const int64_t per_cta = (items + gridDim.x - 1) / gridDim.x;
    const int64_t first = (int64_t)blockIdx.x * per_cta;
    const int64_t last = first + per_cta < items ? first + per_cta : items;

    struct Item {
        const uint4 *body;
        const uint8_t *plane;
        int64_t v0, nvec, head, tail0;
        bool first_tile;
    };
    auto locate = [&](int64_t item) {
        const int c = (int)(item / per_channel);
        const int64_t rem = item - (int64_t)c * per_channel;
        const int64_t n = rem / tiles_per_plane;
        const int64_t t = rem - n * tiles_per_plane;
        Item it;
        it.plane = img + (n * 3 + c) * hw;
        const PlaneSplit sp = split_plane(it.plane, hw);
        it.body = reinterpret_cast<const uint4 *>(it.plane + sp.head);
        it.v0 = t * kPairTileVecs;
        it.nvec = sp.nvec;
        it.head = sp.head;
        it.tail0 = sp.tail0;
        it.first_tile = t == 0;
        return it;
    };
    auto issue = [&](int64_t item, uint4(&v)[kPairUnroll], unsigned &okmask) {
        const Item it = locate(item);
        okmask = 0;
#pragma unroll
        for (int u = 0; u < kPairUnroll; ++u) {
            const int64_t vi = it.v0 + u * kPairThreads + threadIdx.x;
            if (vi < it.nvec) {
                v[u] = ld_stream(it.body + vi);
                okmask |= 1u << u;
            }
        }
    };
    // ragged ends of a plane (< 32 bytes): single values, straight into the segment histogram
    auto stragglers = [&](int64_t item) {
        const Item it = locate(item);
        if (!it.first_tile) return;
        const int64_t ragged = it.head + (hw - it.tail0);
        if ((int64_t)threadIdx.x < ragged) {
            const int64_t idx = (int64_t)threadIdx.x < it.head ? (int64_t)threadIdx.x : it.tail0 + ((int64_t)threadIdx.x - it.head);
            atomicAdd(&hist32[it.plane[idx]], 1u);
        }
    };

    unsigned my_pairs = 0;  // pairs this thread has counted since the last fold
    auto count_pairs = [&](const uint4(&v)[kPairUnroll], unsigned okmask) {
#pragma unroll
        for (int u = 0; u < kPairUnroll; ++u) {
            if (okmask & (1u << u)) {
                pair_count_word(v[u].x, table_addr);
                pair_count_word(v[u].y, table_addr);
                pair_count_word(v[u].z, table_addr);
                pair_count_word(v[u].w, table_addr);
                my_pairs += 8;
            }
        }
    };

    // Fold the pair table into hist32 (both marginals), re-zero it, and check the cell sum.
    // Thread t walks words t, t + 1024, ...: a = t & 255 is fixed per thread, b & 127 = (t >> 8) + 4k
    // is uniform over a warp.  Returns true when no cell overflowed.
    auto fold_table = [&]() -> bool {
        __syncthreads();
        unsigned acc_a = 0;
#pragma unroll 4
        for (int k = 0; k < kPairTableWords / kPairThreads; ++k) {
            const int i = threadIdx.x + k * kPairThreads;
            const unsigned w = table[i];
            table[i] = 0u;
            const unsigned lo = w & 0xffffu, hi = w >> 16;
            acc_a += lo + hi;
            const unsigned slo = __reduce_add_sync(0xffffffffu, lo), shi = __reduce_add_sync(0xffffffffu, hi);
            if (lane == 0) {
                const int b7 = i >> 8;
                if (slo) atomicAdd(&hist32[b7], slo);
                if (shi) atomicAdd(&hist32[b7 + 128], shi);
            }
        }
        if (acc_a) atomicAdd(&hist32[threadIdx.x & 255], acc_a);
        const unsigned cells = __reduce_add_sync(0xffffffffu, acc_a), pairs = __reduce_add_sync(0xffffffffu, my_pairs);
        if (lane == 0) {
            atomicAdd(&s_misc[0], pairs);
            atomicAdd(&s_misc[1], cells);
        }
        my_pairs = 0;
        __syncthreads();
        const bool ok = s_misc[0] == s_misc[1];
        __syncthreads();
        if (threadIdx.x < 2) s_misc[threadIdx.x] = 0u;
        return ok;
    };
    // hist32 -> global counts of channel c, re-zero.
    auto flush_hist = [&](int c) {
        __syncthreads();
        if (threadIdx.x < 256) {
            const unsigned v = hist32[threadIdx.x];
            if (v) atomicAdd(&counts[c * 256 + threadIdx.x], (unsigned long long)v);
            hist32[threadIdx.x] = 0u;
        }
        __syncthreads();
    };
    // Safe recount of items [a, b) (one channel) with warp-private 32-bit histograms held in the
    // (already re-zeroed) table memory; hist32 keeps only the stragglers until it is rebuilt.
    auto recount = [&](int64_t a, int64_t b) {
        unsigned int *wh = table + (threadIdx.x >> 5) * 256;
        for (int64_t item = a; item < b; ++item) {
            const Item it = locate(item);
#pragma unroll
            for (int u = 0; u < kPairUnroll; ++u) {
                const int64_t vi = it.v0 + u * kPairThreads + threadIdx.x;
                if (vi < it.nvec) {
                    const uint4 v = ld_stream(it.body + vi);
                    const unsigned w[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
                    for (int j = 0; j < 16; ++j) atomicAdd(&wh[(w[j >> 2] >> (8 * (j & 3))) & 0xffu], 1u);
                }
            }
        }
        __syncthreads();
        if (threadIdx.x < 256) {
            unsigned s = 0;
            for (int wgt = 0; wgt < kPairWarps; ++wgt) s += table[wgt * 256 + threadIdx.x];
            hist32[threadIdx.x] += s;
        }
        __syncthreads();
        for (int i = threadIdx.x; i < kPairWarps * 256; i += kPairThreads) table[i] = 0u;
        __syncthreads();
    };

    // Walk the CTA's contiguous item range one channel segment at a time.
    int64_t seg = first;
    while (seg < last) {
        const int c = (int)(seg / per_channel);
        const int64_t chan_end = (int64_t)(c + 1) * per_channel;
        const int64_t seg_end = chan_end < last ? chan_end : last;
        uint4 va[kPairUnroll], vb[kPairUnroll];
        unsigned oka = 0, okb = 0;
        issue(seg, va, oka);
        for (int64_t item = seg; item < seg_end; item += 2) {
            if (item + 1 < seg_end) issue(item + 1, vb, okb);
            count_pairs(va, oka);
            stragglers(item);
            if (item + 1 < seg_end) {
                if (item + 2 < seg_end) issue(item + 2, va, oka);
                count_pairs(vb, okb);
                stragglers(item + 1);
            }
        }
        if (!fold_table()) {  // a 16-bit cell overflowed: drop the pair counts, keep the stragglers
            __syncthreads();
            // hist32 currently holds marginals (garbage) + stragglers; rebuild it from scratch
            if (threadIdx.x < 256) hist32[threadIdx.x] = 0u;
            __syncthreads();
            for (int64_t item = seg; item < seg_end; ++item) stragglers(item);
            recount(seg, seg_end);
        }
        flush_hist(c);
        seg = seg_end;
    }
}

// Histogram, planar float32: quantise, then warp-private shared atomics (12 B/px of traffic per
// 3 values, so the atomic rate is 4x lower than in the uint8 kernel).
__global__ void __launch_bounds__(kThreads) hist_f32_planar_kernel(const float *__restrict__ img, int64_t hw, int64_t n_img, int64_t tiles_per_plane, unsigned long long *__restrict__ counts) {
    __shared__ unsigned int hist32[256];
    __shared__ unsigned int whist[kWarps * 256];
    const int c = blockIdx.y;
    for (int i = threadIdx.x; i < 256; i += kThreads) hist32[i] = 0u;
    WarpAtomics wa;
    wa.init(whist);
    __syncthreads();
    const int64_t items = n_img * tiles_per_plane;
    for (int64_t item = blockIdx.x; item < items; item += gridDim.x) {
        const int64_t n = item / tiles_per_plane;
        const int64_t t = item - n * tiles_per_plane;
        const float *plane = img + (n * 3 + c) * hw;
        const PlaneSplit sp = split_plane(plane, hw);
        const float4 *body = reinterpret_cast<const float4 *>(plane + sp.head);
        const int64_t v0 = t * kTileVecs;
        float4 v[kUnroll];
        bool ok[kUnroll];
#pragma unroll
        for (int u = 0; u < kUnroll; ++u) {
            int64_t vi = v0 + u * kThreads + threadIdx.x;
            ok[u] = vi < sp.nvec;
            if (ok[u]) v[u] = ld_stream(body + vi);
        }
#pragma unroll
        for (int u = 0; u < kUnroll; ++u) {
            if (ok[u]) {
                wa.add(quantize_u8(v[u].x));
                wa.add(quantize_u8(v[u].y));
                wa.add(quantize_u8(v[u].z));
                wa.add(quantize_u8(v[u].w));
            }
        }
        if (t == 0) {
            int64_t ragged = sp.head + (hw - sp.tail0);
            if ((int64_t)threadIdx.x < ragged) {
                int64_t idx = (int64_t)threadIdx.x < sp.head ? (int64_t)threadIdx.x : sp.tail0 + ((int64_t)threadIdx.x - sp.head);
                wa.add(quantize_u8(plane[idx]));
            }
        }
    }
    __syncwarp();
    for (int i = threadIdx.x & 31; i < 256; i += 32)
        if (wa.wh[i]) atomicAdd(&hist32[i], wa.wh[i]);
    __syncthreads();
    for (int i = threadIdx.x; i < 256; i += kThreads)
        if (hist32[i]) atomicAdd(&counts[c * 256 + i], (unsigned long long)hist32[i]);
}

// Histogram, interleaved (NHWC): the batch is one flat array of 3*npix elements, element i
// belongs to channel i % 3.  A thread takes 3 consecutive vectors = 48 B (uint8: 16 px) or 12
// floats (4 px), so the channel of every register lane is a compile-time constant.
template <typename T>
__global__ void __launch_bounds__(kThreads) hist_nhwc_kernel(const T *__restrict__ img, int64_t total, unsigned long long *__restrict__ counts) {
    constexpr int kPerVec = 16 / sizeof(T);
    constexpr int kGroup = 3 * kPerVec;  // elements per thread-iteration, multiple of 3
    __shared__ unsigned int whist[kWarps * 3 * 256];
    unsigned int *wh = whist + (threadIdx.x >> 5) * 768;
    for (int i = threadIdx.x & 31; i < 768; i += 32) wh[i] = 0u;
    __syncwarp();

    const bool vec_ok = (reinterpret_cast<uintptr_t>(img) & 15) == 0;
    const int64_t groups = vec_ok ? total / kGroup : 0;
    for (int64_t g = (int64_t)blockIdx.x * kThreads + threadIdx.x; g < groups; g += (int64_t)gridDim.x * kThreads) {
        if constexpr (sizeof(T) == 1) {
            const uint4 *p = reinterpret_cast<const uint4 *>(img + g * kGroup);
            uint4 a = ld_stream(p), b = ld_stream(p + 1), d = ld_stream(p + 2);
            unsigned w[12] = {a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w, d.x, d.y, d.z, d.w};
#pragma unroll
            for (int j = 0; j < 48; ++j) {
                unsigned v = (w[j >> 2] >> (8 * (j & 3))) & 0xffu;
                atomicAdd(&wh[(j % 3) * 256 + v], 1u);
            }
        } else {
            const float4 *p = reinterpret_cast<const float4 *>(img + g * kGroup);
            float4 a = ld_stream(p), b = ld_stream(p + 1), d = ld_stream(p + 2);
            float f[12] = {a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w, d.x, d.y, d.z, d.w};
#pragma unroll
            for (int j = 0; j < 12; ++j) atomicAdd(&wh[(j % 3) * 256 + quantize_u8(f[j])], 1u);
        }
    }
    // scalar remainder (everything when the base pointer is not 16-byte aligned)
    for (int64_t i = groups * kGroup + (int64_t)blockIdx.x * kThreads + threadIdx.x; i < total; i += (int64_t)gridDim.x * kThreads) {
        unsigned v;
        if constexpr (sizeof(T) == 1) v = img[i];
        else v = quantize_u8(img[i]);
        atomicAdd(&wh[(int)(i % 3) * 256 + v], 1u);
    }
    __syncthreads();
    for (int i = threadIdx.x; i < 768; i += kThreads) {
        unsigned long long s = 0;
#pragma unroll
        for (int wgt = 0; wgt < kWarps; ++wgt) s += whist[wgt * 768 + i];
        if (s) atomicAdd(&counts[i], s);
    }
}

// ------------------------------------------------------------------------------------------------
// LUT construction.  Tiny, single CTA per channel; every float32 operation is pinned with
// round-to-nearest intrinsics so the result is bit-identical to the torch CPU backend of the reference.
// ------------------------------------------------------------------------------------------------

// torch.sum of a 256-vector on CPU: 8 lanes x 4 interleaved accumulators (probed against torch; DESIGN.md section "bit-exact LUT").
__device__ float torch_sum_256(const float *a) {
    float acc[4][8];
#pragma unroll
    for (int k = 0; k < 4; ++k)
#pragma unroll
        for (int l = 0; l < 8; ++l) acc[k][l] = 0.0f;
    for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int k = 0; k < 4; ++k)
#pragma unroll
            for (int l = 0; l < 8; ++l) acc[k][l] = __fadd_rn(acc[k][l], a[(i * 4 + k) * 8 + l]);
#pragma unroll
    for (int k = 1; k < 4; ++k)
#pragma unroll
        for (int l = 0; l < 8; ++l) acc[0][l] = __fadd_rn(acc[0][l], acc[k][l]);
    float f = 0.0f;
#pragma unroll
    for (int l = 0; l < 8; ++l) f = __fadd_rn(f, acc[0][l]);
    return f;
}

// H1: torch_backend.py:L139-141.
__global__ void ref_hist_kernel(const unsigned long long *__restrict__ counts, float *__restrict__ ref_hist) {
    __shared__ float cf[256];
    __shared__ float denom;
    const int c = blockIdx.x, b = threadIdx.x;
    cf[b] = __ull2float_rn(counts[c * 256 + b]);
    __syncthreads();
    if (b == 0) denom = __fadd_rn(torch_sum_256(cf), 1e-8f);
    __syncthreads();
    ref_hist[c * 256 + b] = __fdiv_rn(cf[b], denom);
}

// torch.cumsum(float32) on CPU: one double accumulator, rounded to float32 per element.
// Serial form (one thread): the 32 loads + conversions of a chunk are issued together so that only
// the DADD chain is on the critical path (still ~10 us for 256 elements: FP64 adds are slow here).
__device__ __forceinline__ void serial_cumsum_256(const float *__restrict__ in, double *__restrict__ out) {
    double acc = 0.0;
    for (int chunk = 0; chunk < 8; ++chunk) {
        double v[32];
#pragma unroll
        for (int j = 0; j < 32; ++j) v[j] = (double)in[chunk * 32 + j];
#pragma unroll
        for (int j = 0; j < 32; ++j) {
            acc = __dadd_rn(acc, v[j]);
            out[chunk * 32 + j] = acc;
        }
    }
}

// The same running sums by a 256-thread scan, used whenever double addition is EXACT for the data
// and therefore independent of the order: every term is a non-negative float32 >= 2^-29 (or 0),
// i.e. a multiple of 2^-52, and the total stays below 2, so each partial sum is a multiple of
// 2^-52 below 2 and fits the 53-bit significand.  (Histogram fractions count / npix satisfy this up
// to 2^29 pixels per batch.)  Anything else takes the serial loop.  out[b] = sum_{i <= b} in[i].
__device__ __forceinline__ void cumsum_256(const float *__restrict__ in, double *__restrict__ out) {
    __shared__ double s_warp[8];
    const int b = threadIdx.x, lane = b & 31, warp = b >> 5;
    const float v = in[b];
    const bool exact_term = v == 0.0f || (v >= 1.862645149230957e-09f && v < 2.0f);  // 2^-29
    double x = (double)v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const double y = __shfl_up_sync(0xffffffffu, x, o);
        if (lane >= o) x = __dadd_rn(x, y);
    }
    if (lane == 31) s_warp[warp] = x;
    __syncthreads();
    double base = 0.0;
    for (int k = 0; k < warp; ++k) base = __dadd_rn(base, s_warp[k]);
    x = __dadd_rn(x, base);
    const bool fine = exact_term && (b != 255 || x < 2.0);
    if (__syncthreads_and(fine)) {
        out[b] = x;
    } else if (b == 0) {
        serial_cumsum_256(in, out);
    }
    __syncthreads();
}

// Reference CDF of channel c into rq[256] (shared): H2a, torch_backend.py:L221-223.
__device__ __forceinline__ void ref_cdf_to_smem(const float *__restrict__ ref_hist_c, float *h, double *dacc, float *rq) {
    __shared__ float s_denom;
    const int b = threadIdx.x;
    h[b] = ref_hist_c[b];
    __syncthreads();
    if (b == 0) s_denom = __fadd_rn(torch_sum_256(h), 1e-8f);
    __syncthreads();
    h[b] = __fdiv_rn(h[b], s_denom);
    __syncthreads();
    cumsum_256(h, dacc);
    rq[b] = __double2float_rn(dacc[b]);
    __syncthreads();
}

__global__ void ref_cdf_kernel(const float *__restrict__ ref_hist, float *__restrict__ ref_cdf) {
    __shared__ float h[256];
    __shared__ double dacc[256];
    __shared__ float rq[256];
    ref_cdf_to_smem(ref_hist + blockIdx.x * 256, h, dacc, rq);
    ref_cdf[blockIdx.x * 256 + threadIdx.x] = rq[threadIdx.x];
}

// H2b: torch_backend.py:L234-281.  npix < 0: derive the pixel count from the counts themselves
// (sum over the 256 bins of the channel), which keeps a sharded run free of host round trips.
// FROM_HIST: the reference CDF is rebuilt from ref_hist inside the same kernel (fused transform).
template <bool FROM_HIST>
__global__ void build_lut_kernel(const unsigned long long *__restrict__ counts, long long npix, const float *__restrict__ ref, float *__restrict__ lut) {
    __shared__ float rq[256];
    __shared__ float sq[256];
    __shared__ double dacc[256];
    __shared__ float s_npix_f;
    const int c = blockIdx.x, b = threadIdx.x;
    if (FROM_HIST) ref_cdf_to_smem(ref + c * 256, sq, dacc, rq);
    else rq[b] = ref[c * 256 + b];
    const unsigned long long my_count = counts[c * 256 + b];
    {
        __shared__ unsigned long long s_part[8];
        unsigned long long t = my_count;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) t += __shfl_xor_sync(0xffffffffu, t, o);
        if ((b & 31) == 0) s_part[b >> 5] = t;
        __syncthreads();
        if (b == 0) {
            unsigned long long total = 0;
            for (int k = 0; k < 8; ++k) total += s_part[k];
            if (npix >= 0) total = (unsigned long long)npix;
            // L235: python float (num_pixels + 1e-8), cast to float32 for the division
            s_npix_f = __double2float_rn(__dadd_rn((double)total, 1e-8));
        }
    }
    __syncthreads();
    sq[b] = __fdiv_rn(__ull2float_rn(my_count), s_npix_f);  // L234-235
    __syncthreads();
    cumsum_256(sq, dacc);  // L236: cumsum, double accumulator rounded per element
    sq[b] = __double2float_rn(dacc[b]);
    __syncthreads();
    const float q = sq[b];
    int lo = 0, hi = 256;  // L260: searchsorted(right=False)
    while (lo < hi) {
        int mid = (lo + hi) >> 1;
        if (rq[mid] < q) lo = mid + 1;
        else hi = mid;
    }
    const int idx = min(max(lo, 1), 255);                                   // L261
    const float ql = rq[idx - 1], qr = rq[idx];                             // L264-265
    const float d = __fsub_rn(qr, ql);                                      // L272
    const float alpha = d > 1e-10f ? __fdiv_rn(__fsub_rn(q, ql), d) : 0.f;  // L273
    float v = __fadd_rn((float)(idx - 1), alpha);                           // L276 (ref_values step is exactly 1)
    if (q <= rq[0]) v = 0.0f;                                               // L268, L279
    if (q >= rq[255]) v = 255.0f;                                           // L269, L280
    lut[c * 256 + b] = fminf(fmaxf(v, 0.0f), 255.0f);                       // L281
}

// ------------------------------------------------------------------------------------------------
// LUT remap (H4).
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ unsigned remap4(const unsigned char *l, unsigned w) {
    return (unsigned)l[w & 0xffu] | ((unsigned)l[(w >> 8) & 0xffu] << 8) | ((unsigned)l[(w >> 16) & 0xffu] << 16) | ((unsigned)l[w >> 24] << 24);
}

// uint8 planar.  Items are walked from the END of the batch: the histogram pass has just
// streamed the batch front to back, so the tail is what is still resident in L2.
__global__ void __launch_bounds__(kThreads) apply_u8_planar_kernel(const uint8_t *__restrict__ img, uint8_t *__restrict__ out, int64_t hw, int64_t planes, int64_t tiles_per_plane, const float *__restrict__ lut) {
    __shared__ unsigned char lut8[3 * 256];
    for (int i = threadIdx.x; i < 768; i += kThreads) lut8[i] = (unsigned char)__float2int_rz(lut[i]);  // trunc, L296-298
    __syncthreads();
    const int64_t items = planes * tiles_per_plane;
    for (int64_t it = blockIdx.x; it < items; it += gridDim.x) {
        const int64_t item = items - 1 - it;
        const int64_t pl = item / tiles_per_plane;
        const int64_t t = item - pl * tiles_per_plane;
        const unsigned char *l = lut8 + (int)(pl % 3) * 256;
        const uint8_t *src = img + pl * hw;
        uint8_t *dst = out + pl * hw;
        const PlaneSplit sp = split_plane(src, hw);
        const bool dst_vec = ((reinterpret_cast<uintptr_t>(dst + sp.head)) & 15) == 0;
        const uint4 *body = reinterpret_cast<const uint4 *>(src + sp.head);
        const int64_t v0 = t * kTileVecs;
        uint4 v[kUnroll];
        bool ok[kUnroll];
#pragma unroll
        for (int u = 0; u < kUnroll; ++u) {
            int64_t vi = v0 + u * kThreads + threadIdx.x;
            ok[u] = vi < sp.nvec;
            if (ok[u]) v[u] = ld_stream(body + vi);
        }
#pragma unroll
        for (int u = 0; u < kUnroll; ++u) {
            if (!ok[u]) continue;
            int64_t vi = v0 + u * kThreads + threadIdx.x;
            uint4 r;
            r.x = remap4(l, v[u].x); r.y = remap4(l, v[u].y); r.z = remap4(l, v[u].z); r.w = remap4(l, v[u].w);
            if (dst_vec) {
                st_stream(reinterpret_cast<uint4 *>(dst + sp.head) + vi, r);
            } else {
                unsigned w[4] = {r.x, r.y, r.z, r.w};
                uint8_t *d = dst + sp.head + vi * 16;
#pragma unroll
                for (int j = 0; j < 16; ++j) d[j] = (uint8_t)(w[j >> 2] >> (8 * (j & 3)));
            }
        }
        if (t == 0) {
            int64_t ragged = sp.head + (hw - sp.tail0);
            if ((int64_t)threadIdx.x < ragged) {
                int64_t idx = (int64_t)threadIdx.x < sp.head ? (int64_t)threadIdx.x : sp.tail0 + ((int64_t)threadIdx.x - sp.head);
                dst[idx] = l[src[idx]];
            }
        }
    }
}

// float32 planar: out = clamp(lut[trunc(clamp(255 x))] / 255, 0, 1)   (L290-296).
__global__ void __launch_bounds__(kThreads) apply_f32_planar_kernel(const float *__restrict__ img, float *__restrict__ out, int64_t hw, int64_t planes, int64_t tiles_per_plane, const float *__restrict__ lut) {
    __shared__ float lutf[3 * 256];
    for (int i = threadIdx.x; i < 768; i += kThreads) lutf[i] = fminf(fmaxf(__fdiv_rn(lut[i], 255.0f), 0.0f), 1.0f);
    __syncthreads();
    const int64_t items = planes * tiles_per_plane;
    for (int64_t it = blockIdx.x; it < items; it += gridDim.x) {
        const int64_t item = items - 1 - it;
        const int64_t pl = item / tiles_per_plane;
        const int64_t t = item - pl * tiles_per_plane;
        const float *l = lutf + (int)(pl % 3) * 256;
        const float *src = img + pl * hw;
        float *dst = out + pl * hw;
        const PlaneSplit sp = split_plane(src, hw);
        const bool dst_vec = ((reinterpret_cast<uintptr_t>(dst + sp.head)) & 15) == 0;
        const float4 *body = reinterpret_cast<const float4 *>(src + sp.head);
        const int64_t v0 = t * kTileVecs;
        float4 v[kUnroll];
        bool ok[kUnroll];
#pragma unroll
        for (int u = 0; u < kUnroll; ++u) {
            int64_t vi = v0 + u * kThreads + threadIdx.x;
            ok[u] = vi < sp.nvec;
            if (ok[u]) v[u] = ld_stream(body + vi);
        }
#pragma unroll
        for (int u = 0; u < kUnroll; ++u) {
            if (!ok[u]) continue;
            int64_t vi = v0 + u * kThreads + threadIdx.x;
            float4 r = make_float4(l[quantize_u8(v[u].x)], l[quantize_u8(v[u].y)], l[quantize_u8(v[u].z)], l[quantize_u8(v[u].w)]);
            if (dst_vec) {
                st_stream(reinterpret_cast<float4 *>(dst + sp.head) + vi, r);
            } else {
                float *d = dst + sp.head + vi * 4;
                d[0] = r.x; d[1] = r.y; d[2] = r.z; d[3] = r.w;
            }
        }
        if (t == 0) {
            int64_t ragged = sp.head + (hw - sp.tail0);
            if ((int64_t)threadIdx.x < ragged) {
                int64_t idx = (int64_t)threadIdx.x < sp.head ? (int64_t)threadIdx.x : sp.tail0 + ((int64_t)threadIdx.x - sp.head);
                dst[idx] = l[quantize_u8(src[idx])];
            }
        }
    }
}

// Interleaved (NHWC) remap, uint8 or float32.
template <typename T>
__global__ void __launch_bounds__(kThreads) apply_nhwc_kernel(const T *__restrict__ img, T *__restrict__ out, int64_t total, const float *__restrict__ lut) {
    constexpr int kPerVec = 16 / sizeof(T);
    constexpr int kGroup = 3 * kPerVec;
    __shared__ float lutf[3 * 256];
    __shared__ unsigned char lut8[3 * 256];
    for (int i = threadIdx.x; i < 768; i += kThreads) {
        lut8[i] = (unsigned char)__float2int_rz(lut[i]);
        lutf[i] = fminf(fmaxf(__fdiv_rn(lut[i], 255.0f), 0.0f), 1.0f);
    }
    __syncthreads();
    const bool vec_ok = ((reinterpret_cast<uintptr_t>(img) | reinterpret_cast<uintptr_t>(out)) & 15) == 0;
    const int64_t groups = vec_ok ? total / kGroup : 0;
    for (int64_t g = (int64_t)blockIdx.x * kThreads + threadIdx.x; g < groups; g += (int64_t)gridDim.x * kThreads) {
        if constexpr (sizeof(T) == 1) {
            const uint4 *p = reinterpret_cast<const uint4 *>(img + g * kGroup);
            uint4 a = ld_stream(p), b = ld_stream(p + 1), d = ld_stream(p + 2);
            unsigned w[12] = {a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w, d.x, d.y, d.z, d.w};
            unsigned r[12];
#pragma unroll
            for (int k = 0; k < 12; ++k) r[k] = 0u;
#pragma unroll
            for (int j = 0; j < 48; ++j) {
                unsigned v = (w[j >> 2] >> (8 * (j & 3))) & 0xffu;
                r[j >> 2] |= (unsigned)lut8[(j % 3) * 256 + v] << (8 * (j & 3));
            }
            uint4 *q = reinterpret_cast<uint4 *>(out + g * kGroup);
            st_stream(q, make_uint4(r[0], r[1], r[2], r[3]));
            st_stream(q + 1, make_uint4(r[4], r[5], r[6], r[7]));
            st_stream(q + 2, make_uint4(r[8], r[9], r[10], r[11]));
        } else {
            const float4 *p = reinterpret_cast<const float4 *>(img + g * kGroup);
            float4 a = ld_stream(p), b = ld_stream(p + 1), d = ld_stream(p + 2);
            float f[12] = {a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w, d.x, d.y, d.z, d.w};
#pragma unroll
            for (int j = 0; j < 12; ++j) f[j] = lutf[(j % 3) * 256 + quantize_u8(f[j])];
            float4 *q = reinterpret_cast<float4 *>(out + g * kGroup);
            st_stream(q, make_float4(f[0], f[1], f[2], f[3]));
            st_stream(q + 1, make_float4(f[4], f[5], f[6], f[7]));
            st_stream(q + 2, make_float4(f[8], f[9], f[10], f[11]));
        }
    }
    for (int64_t i = groups * kGroup + (int64_t)blockIdx.x * kThreads + threadIdx.x; i < total; i += (int64_t)gridDim.x * kThreads) {
        if constexpr (sizeof(T) == 1) out[i] = lut8[(int)(i % 3) * 256 + img[i]];
        else out[i] = lutf[(int)(i % 3) * 256 + quantize_u8(img[i])];
    }
}

// ---- tuning knobs (A/B measurements; defaults are the measured winners) -----------------------
static int g_hist_byte_counters = 0;  // counting scheme: 0 warp atomics, 1 byte counters, 2 packed RED, 3 lane32, 4 pairs
static int g_hist_ctas_per_sm = 8;
static int g_apply_ctas_per_sm = 8;

}  // namespace hm
}  // namespace sx

using namespace sx;
using namespace sx::hm;

extern "C" {

// Undocumented-in-header tuning hook used by bench/profiling scripts.
int sx_hm_set_tuning(int hist_byte_counters, int hist_ctas_per_sm, int apply_ctas_per_sm) {
    if (hist_byte_counters >= 0) g_hist_byte_counters = hist_byte_counters;
    if (hist_ctas_per_sm > 0) g_hist_ctas_per_sm = hist_ctas_per_sm;
    if (apply_ctas_per_sm > 0) g_apply_ctas_per_sm = apply_ctas_per_sm;
    return SX_OK;
}

int sx_hm_hist(const void *images, int dtype, int layout, int64_t n, int64_t h, int64_t w, uint64_t *counts, sx_stream_t stream_) {
    if (int rc = check_images(images, dtype, n, h, w)) return rc;
    SX_REQUIRE(counts != nullptr, "counts is NULL");
    SX_REQUIRE(layout == SX_NCHW || layout == SX_NHWC, "layout must be SX_NCHW or SX_NHWC, got %d", layout);
    cudaStream_t stream = static_cast<cudaStream_t>(stream_);
    const int64_t hw = h * w;
    if (n == 0 || hw == 0) return SX_OK;
    auto *cnt = reinterpret_cast<unsigned long long *>(counts);
    if (layout == SX_NHWC) {
        const int64_t total = n * hw * 3;
        if (dtype == SX_U8) {
            unsigned grid = stream_grid((total / 48 + kThreads - 1) / kThreads + 1, 8);
            hist_nhwc_kernel<uint8_t><<<grid, kThreads, 0, stream>>>(static_cast<const uint8_t *>(images), total, cnt);
        } else {
            unsigned grid = stream_grid((total / 12 + kThreads - 1) / kThreads + 1, 8);
            hist_nhwc_kernel<float><<<grid, kThreads, 0, stream>>>(static_cast<const float *>(images), total, cnt);
        }
        SX_LAUNCHED("hist_nhwc_kernel");
        return SX_OK;
    }
    if (dtype == SX_U8) {
        const int64_t tiles = max_i64(1, (hw / 16 + kTileVecs - 1) / kTileVecs);
        const int64_t items = n * tiles;
        if (g_hist_byte_counters == 4) {
            static bool attr4_set = false;
            if (!attr4_set) {
                SX_CUDA(cudaFuncSetAttribute(hist_u8_planar_pairs_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kPairSmemBytes));
                attr4_set = true;
            }
            const int64_t tiles_p = max_i64(1, (hw / 16 + kPairTileVecs - 1) / kPairTileVecs);
            const unsigned grid_p = stream_grid(3 * n * tiles_p, 1);
            hist_u8_planar_pairs_kernel<<<grid_p, kPairThreads, kPairSmemBytes, stream>>>(static_cast<const uint8_t *>(images), hw, n, tiles_p, cnt);
        } else if (g_hist_byte_counters == 3) {
            static bool attr3_set = false;
            if (!attr3_set) {
                SX_CUDA(cudaFuncSetAttribute(hist_u8_planar_lane32_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kL32SmemBytes));
                attr3_set = true;
            }
            const int64_t tiles32 = max_i64(1, (hw / 16 + kL32TileVecs - 1) / kL32TileVecs);
            const unsigned grid32 = stream_grid(3 * n * tiles32, 1);
            hist_u8_planar_lane32_kernel<<<grid32, kL32Threads, kL32SmemBytes, stream>>>(static_cast<const uint8_t *>(images), hw, n, tiles32, cnt);
        } else if (g_hist_byte_counters) {
            const size_t smem = 256 * sizeof(unsigned) + (size_t)kWarps * ByteCounters::kBytesPerWarp;
            static bool attr_set = false;
            if (!attr_set) {
                SX_CUDA(cudaFuncSetAttribute(hist_u8_planar_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
                SX_CUDA(cudaFuncSetAttribute(hist_u8_planar_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
                attr_set = true;
            }
            dim3 grid(stream_grid(items, g_hist_ctas_per_sm), 3);
            grid.x = (grid.x + 2) / 3 > 0 ? (grid.x + 2) / 3 : 1;
            if (g_hist_byte_counters == 2) hist_u8_planar_kernel<2><<<grid, kThreads, smem, stream>>>(static_cast<const uint8_t *>(images), hw, n, tiles, cnt);
            else hist_u8_planar_kernel<1><<<grid, kThreads, smem, stream>>>(static_cast<const uint8_t *>(images), hw, n, tiles, cnt);
        } else {
            const size_t smem = 256 * sizeof(unsigned) + (size_t)kWarps * 256 * sizeof(unsigned);
            dim3 grid(stream_grid(items, g_hist_ctas_per_sm), 3);
            grid.x = (grid.x + 2) / 3 > 0 ? (grid.x + 2) / 3 : 1;
            hist_u8_planar_kernel<0><<<grid, kThreads, smem, stream>>>(static_cast<const uint8_t *>(images), hw, n, tiles, cnt);
        }
        SX_LAUNCHED("hist_u8_planar_kernel");
    } else {
        const int64_t tiles = max_i64(1, (hw / 4 + kTileVecs - 1) / kTileVecs);
        dim3 grid(stream_grid(n * tiles, 6), 3);
        grid.x = (grid.x + 2) / 3 > 0 ? (grid.x + 2) / 3 : 1;
        hist_f32_planar_kernel<<<grid, kThreads, 0, stream>>>(static_cast<const float *>(images), hw, n, tiles, cnt);
        SX_LAUNCHED("hist_f32_planar_kernel");
    }
    return SX_OK;
}

int sx_hm_ref_hist(const uint64_t *counts, float *ref_hist, sx_stream_t stream) {
    SX_REQUIRE(counts && ref_hist, "NULL argument");
    ref_hist_kernel<<<3, 256, 0, static_cast<cudaStream_t>(stream)>>>(reinterpret_cast<const unsigned long long *>(counts), ref_hist);
    SX_LAUNCHED("ref_hist_kernel");
    return SX_OK;
}

int sx_hm_ref_cdf(const float *ref_hist, float *ref_cdf, sx_stream_t stream) {
    SX_REQUIRE(ref_hist && ref_cdf, "NULL argument");
    ref_cdf_kernel<<<3, 256, 0, static_cast<cudaStream_t>(stream)>>>(ref_hist, ref_cdf);
    SX_LAUNCHED("ref_cdf_kernel");
    return SX_OK;
}

int sx_hm_build_lut(const uint64_t *counts, int64_t npix, const float *ref_cdf, float *lut, sx_stream_t stream) {
    SX_REQUIRE(counts && ref_cdf && lut, "NULL argument");
    build_lut_kernel<false><<<3, 256, 0, static_cast<cudaStream_t>(stream)>>>(reinterpret_cast<const unsigned long long *>(counts), (long long)npix, ref_cdf, lut);
    SX_LAUNCHED("build_lut_kernel");
    return SX_OK;
}

int sx_hm_apply(const void *images, int dtype, int layout, int64_t n, int64_t h, int64_t w, const float *lut, void *out, sx_stream_t stream_) {
    if (int rc = check_images(images, dtype, n, h, w)) return rc;
    SX_REQUIRE(layout == SX_NCHW || layout == SX_NHWC, "layout must be SX_NCHW or SX_NHWC, got %d", layout);
    cudaStream_t stream = static_cast<cudaStream_t>(stream_);
    const int64_t hw = h * w;
    if (n == 0 || hw == 0) return SX_OK;
    SX_REQUIRE(lut && out, "NULL argument");
    if (layout == SX_NHWC) {
        const int64_t total = n * hw * 3;
        if (dtype == SX_U8) {
            unsigned grid = stream_grid((total / 48 + kThreads - 1) / kThreads + 1, 8);
            apply_nhwc_kernel<uint8_t><<<grid, kThreads, 0, stream>>>(static_cast<const uint8_t *>(images), static_cast<uint8_t *>(out), total, lut);
        } else {
            unsigned grid = stream_grid((total / 12 + kThreads - 1) / kThreads + 1, 8);
            apply_nhwc_kernel<float><<<grid, kThreads, 0, stream>>>(static_cast<const float *>(images), static_cast<float *>(out), total, lut);
        }
        SX_LAUNCHED("apply_nhwc_kernel");
        return SX_OK;
    }
    const int64_t planes = n * 3;
    if (dtype == SX_U8) {
        const int64_t tiles = max_i64(1, (hw / 16 + kTileVecs - 1) / kTileVecs);
        unsigned grid = stream_grid(planes * tiles, g_apply_ctas_per_sm);
        apply_u8_planar_kernel<<<grid, kThreads, 0, stream>>>(static_cast<const uint8_t *>(images), static_cast<uint8_t *>(out), hw, planes, tiles, lut);
        SX_LAUNCHED("apply_u8_planar_kernel");
    } else {
        const int64_t tiles = max_i64(1, (hw / 4 + kTileVecs - 1) / kTileVecs);
        unsigned grid = stream_grid(planes * tiles, g_apply_ctas_per_sm);
        apply_f32_planar_kernel<<<grid, kThreads, 0, stream>>>(static_cast<const float *>(images), static_cast<float *>(out), hw, planes, tiles, lut);
        SX_LAUNCHED("apply_f32_planar_kernel");
    }
    return SX_OK;
}

// workspace: counts u64[768] | ref_cdf f32[768] | lut f32[768]
int64_t sx_hm_workspace_bytes(void) { return 768 * 8 + 768 * 4 + 768 * 4; }

int sx_hm_transform(const void *images, int dtype, int layout, int64_t n, int64_t h, int64_t w, const float *ref_hist, void *out, void *workspace, int64_t workspace_bytes, sx_stream_t stream) {
    SX_REQUIRE(workspace && workspace_bytes >= sx_hm_workspace_bytes(), "workspace too small (%lld < %lld)", (long long)workspace_bytes, (long long)sx_hm_workspace_bytes());
    SX_REQUIRE(ref_hist != nullptr, "ref_hist is NULL");
    auto *counts = static_cast<uint64_t *>(workspace);
    auto *ref_cdf = reinterpret_cast<float *>(counts + 768);
    auto *lut = ref_cdf + 768;
    SX_CUDA(cudaMemsetAsync(counts, 0, 768 * 8, static_cast<cudaStream_t>(stream)));
    if (int rc = sx_hm_hist(images, dtype, layout, n, h, w, counts, stream)) return rc;
    (void)ref_cdf;
    build_lut_kernel<true><<<3, 256, 0, static_cast<cudaStream_t>(stream)>>>(reinterpret_cast<const unsigned long long *>(counts), (long long)(n * h * w), ref_hist, lut);
    SX_LAUNCHED("build_lut_kernel<fused>");
    return sx_hm_apply(images, dtype, layout, n, h, w, lut, out, stream);
}

int sx_hm_fit(const void *images, int dtype, int layout, int64_t n, int64_t h, int64_t w, float *ref_hist, void *workspace, int64_t workspace_bytes, sx_stream_t stream) {
    SX_REQUIRE(workspace && workspace_bytes >= sx_hm_workspace_bytes(), "workspace too small (%lld < %lld)", (long long)workspace_bytes, (long long)sx_hm_workspace_bytes());
    auto *counts = static_cast<uint64_t *>(workspace);
    SX_CUDA(cudaMemsetAsync(counts, 0, 768 * 8, static_cast<cudaStream_t>(stream)));
    if (int rc = sx_hm_hist(images, dtype, layout, n, h, w, counts, stream)) return rc;
    return sx_hm_ref_hist(counts, ref_hist, stream);
}

}  // extern "C"
