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

// ------------------------------------------------------------------------------------------------
// Histogram, planar uint8, general kernel.  grid = (blocks, 3 channels); a CTA only ever sees one
// channel.  Warp-private 256-bin tables updated with shared-memory atomics (ATOMS.POPC.INC merges
// equal bins of a warp, so flat image regions count faster than noise).  Handles any alignment;
// aligned large batches take the lane-private TMA-fed kernel below.  Counting schemes that were
// built, measured and removed (201 M noise values): thread-private byte counters with plain
// LDS/ADD/STS 120 us; lane-private packed byte counters with RED 109 us; lane-private 32-bit counters
// fed from registers 95 us; this kernel 90 us; the same counting behind a TMA ring 91 us.
// ------------------------------------------------------------------------------------------------
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

// One 128-bit vector of a float-like image (float32: 4 values, float16 / bfloat16: 8), widened to float32.
template <typename T>
struct FloatVec {
    static constexpr int kPer = 16 / sizeof(T);
    float f[kPer];
    __device__ __forceinline__ void load(const T *p) {
        if constexpr (sizeof(T) == 4) {
            const float4 v = ld_stream(reinterpret_cast<const float4 *>(p));
            f[0] = v.x; f[1] = v.y; f[2] = v.z; f[3] = v.w;
        } else {
            const uint4 v = ld_stream(reinterpret_cast<const uint4 *>(p));
            const unsigned w[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                const float2 x = Half2IO<T>::unpack(w[k]);
                f[2 * k] = x.x; f[2 * k + 1] = x.y;
            }
        }
    }
    __device__ __forceinline__ void store(T *p) const {
        if constexpr (sizeof(T) == 4) st_stream(reinterpret_cast<float4 *>(p), make_float4(f[0], f[1], f[2], f[3]));
        else st_stream(reinterpret_cast<uint4 *>(p), make_uint4(Half2IO<T>::pack(f[0], f[1]), Half2IO<T>::pack(f[2], f[3]), Half2IO<T>::pack(f[4], f[5]), Half2IO<T>::pack(f[6], f[7])));
    }
    static __device__ __forceinline__ float widen(T v) {
        if constexpr (sizeof(T) == 4) return v;
        else return Half2IO<T>::widen(v);
    }
    static __device__ __forceinline__ T narrow(float v) {
        if constexpr (sizeof(T) == 4) return v;
        else return Half2IO<T>::narrow(v);
    }
};

__global__ void __launch_bounds__(kThreads) hist_u8_planar_kernel(const uint8_t *__restrict__ img, int64_t hw, int64_t n_img, int64_t tiles_per_plane, unsigned long long *__restrict__ counts) {
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
        for (int u = 0; u < kUnroll; ++u)
            if (ok[u]) { wa.add4(v[u].x); wa.add4(v[u].y); wa.add4(v[u].z); wa.add4(v[u].w); }
        if (t == 0) {  // ragged ends of the plane: < 32 bytes, one thread each
            int64_t ragged = sp.head + (hw - sp.tail0);
            if ((int64_t)threadIdx.x < ragged) {
                int64_t idx = (int64_t)threadIdx.x < sp.head ? (int64_t)threadIdx.x : sp.tail0 + ((int64_t)threadIdx.x - sp.head);
                wa.add(plane[idx]);
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

// ------------------------------------------------------------------------------------------------
// TMA-fed kernels: tile list and shared-memory ring.
// ------------------------------------------------------------------------------------------------
// Position in the channel-major tile list (channel, image, tile) of an ALIGNED planar uint8 batch
// (16-byte aligned base, H*W a multiple of 16: every plane is a whole number of 128-bit vectors).
// Advancing costs an add and a counter: consecutive items are consecutive tiles of a plane, then
// the same channel of the next image.  All items of a cursor share a channel.
struct TileCursor {
    int64_t off;        // vector index (from the batch base) of the first vector of the current tile
    int t;              // tile within the plane
    int tiles;          // tiles per plane
    int last_vecs;      // vectors in the last tile of a plane (1 .. tile size)
    int64_t plane_gap;  // vectors from the start of a plane's last tile to the next image's plane (same channel)
    __device__ __forceinline__ void seek(int64_t hw, int tiles_per_plane, int tile_vecs, int64_t per_channel, int64_t item) {
        const int64_t c = item / per_channel;
        const int64_t rem = item - c * per_channel;
        const int64_t n = rem / tiles_per_plane;
        const int64_t vecs = hw / 16;
        t = (int)(rem - n * tiles_per_plane);
        tiles = tiles_per_plane;
        last_vecs = (int)(vecs - (int64_t)(tiles_per_plane - 1) * tile_vecs);
        plane_gap = 3 * vecs - (int64_t)(tiles_per_plane - 1) * tile_vecs;
        off = (n * 3 + c) * vecs + (int64_t)t * tile_vecs;
    }
    __device__ __forceinline__ int vecs(int tile_vecs) const { return t == tiles - 1 ? last_vecs : tile_vecs; }
    __device__ __forceinline__ void next(int tile_vecs) {
        if (++t == tiles) { t = 0; off += plane_gap; } else off += tile_vecs;
    }
};


// Shared-memory ring of tiles filled by TMA bulk copies.  kStages is a power of two.
template <int kStages, int kTileVecs, int kWarpsInCta>
struct TileRing {
    uint4 *ring;             // [kStages][kTileVecs]
    uint64_t *full, *empty;  // [kStages] each
    unsigned produced, consumed;  // running tile counters (thread 0 / every thread)
    __device__ __forceinline__ void init(unsigned char *ring_mem, uint64_t *bars) {
        ring = reinterpret_cast<uint4 *>(ring_mem);
        full = bars;
        empty = bars + kStages;
        produced = consumed = 0;
        if (threadIdx.x == 0) {
            for (int i = 0; i < kStages; ++i) { mbar_init(full + i, 1); mbar_init(empty + i, kWarpsInCta); }
            mbar_fence_init();
        }
    }
    // thread 0: start the copy of `bytes` (> 0, multiple of 16) at `src` into the next stage
    __device__ __forceinline__ void produce(const uint4 *src, unsigned bytes) {
        const unsigned st = produced & (kStages - 1), use = produced / kStages;
        if (use > 0) mbar_wait(empty + st, (use - 1) & 1);  // every warp has read the previous tenant
        mbar_arrive_expect_tx(full + st, bytes);
        tma_load_1d(ring + (size_t)st * kTileVecs, src, bytes, full + st);
        ++produced;
    }
    // every thread: wait for the next tile; returns its stage.  release() after the reads.
    __device__ __forceinline__ const uint4 *acquire() {
        const unsigned st = consumed & (kStages - 1), use = consumed / kStages;
        mbar_wait(full + st, use & 1);
        return ring + (size_t)st * kTileVecs;
    }
    __device__ __forceinline__ void release() {
        const unsigned st = consumed & (kStages - 1);
        __syncwarp();
        if ((threadIdx.x & 31) == 0) mbar_arrive(empty + st);
        ++consumed;
    }
    // the same for the tile with running index `g` (consumers that take every n-th tile)
    __device__ __forceinline__ const uint4 *acquire_at(unsigned g) {
        const unsigned st = g & (kStages - 1), use = g / kStages;
        mbar_wait(full + st, use & 1);
        return ring + (size_t)st * kTileVecs;
    }
    __device__ __forceinline__ void release_at(unsigned g) {
        __syncwarp();
        if ((threadIdx.x & 31) == 0) mbar_arrive(empty + (g & (kStages - 1)));
    }
};

// ------------------------------------------------------------------------------------------------
// Histogram, planar uint8, scheme L ("lane-private, TMA-fed").  Random bytes make ~3.2-way bank
// conflicts in a warp-private 256-bin table, and the shared-memory atomic unit then retires only
// ~7.7 values per clock per SM (schemes 0 and S both land there).  With one full histogram PER
// LANE, laid out [bin][lane] so that counter (bin, lane) sits in bank `lane`, every update is
// conflict-free and the unit retires ~15 per clock (tools/atomsbench.cu, "lane32").  The price is
// 32 KB of shared memory per warp, i.e. only five warps per SM -- far too few to hide HBM latency
// with their own loads (scheme 3 tried: 95 us).  Here the five warps never touch HBM: one thread
// streams 16 KB tiles into a 4-stage shared-memory ring with TMA bulk copies (48 KB in flight per
// SM), the warps read their vectors from the ring and only count.  No overflow handling: 32-bit
// counters.  Per byte: PRMT + IMAD + ATOMS.POPC.INC.
// ------------------------------------------------------------------------------------------------
// Geometry variants (warps, vectors per tile, ring stages): 5 x 16 KB x 4 (the first cut) and
// 6 warps with smaller tiles, which just fit the 227 KB of an SM.
template <int W, int TV, int ST, bool COUNT = true, bool PRODUCER = false>
struct LaneCfg {
    static constexpr bool kCount = COUNT;        // false: ring streaming only (feed-rate measurement)
    static constexpr bool kProducer = PRODUCER;  // an extra warp that only issues the TMA copies
    static_assert((ST & (ST - 1)) == 0, "ring stages must be a power of two");
    static constexpr int kWarps = W, kCountThreads = W * 32, kThreads = (W + (PRODUCER ? 1 : 0)) * 32, kTileVecs = TV, kStages = ST;
    static constexpr int kPerThread = (TV + W * 32 - 1) / (W * 32);
    static constexpr int kSmem = W * 32768 + ST * TV * 16 + 2 * ST * 8 + 256 * 4;
};

__device__ __forceinline__ void lane_count4(unsigned w, unsigned base) {  // base = &region[0][lane]
    const unsigned a0 = (w & 0xffu) * 128u + base;
    const unsigned a1 = __byte_perm(w, 0, 0x4441) * 128u + base;
    const unsigned a2 = __byte_perm(w, 0, 0x4442) * 128u + base;
    const unsigned a3 = (w >> 24) * 128u + base;
    asm volatile("red.shared.add.u32 [%0], 1;" ::"r"(a0) : "memory");
    asm volatile("red.shared.add.u32 [%0], 1;" ::"r"(a1) : "memory");
    asm volatile("red.shared.add.u32 [%0], 1;" ::"r"(a2) : "memory");
    asm volatile("red.shared.add.u32 [%0], 1;" ::"r"(a3) : "memory");
}

// (Run-length merging inside the word -- a run of equal neighbouring bytes becomes ONE update carrying the run's
// length, `red.shared.add.u32 [a], len` -- was built and measured in round 2 on 64 x 3 x 1024^2: noise 217 us, a
// flat batch 197 us, tiled real H&E crops 232 us, against 60 us for the kernel below on all three
// (profiles/r02_hm_hist_runlength_merge_experiment.log).  The +1 form compiles to ATOMS.POPC.INC, which the unit
// retires at 2.1 clk per warp; an update with a register addend is ATOMS.ADD and is ~13x slower per update, so even
// a batch that needs 4x fewer updates loses.  Merging only pays if the addend is the constant 1.)
template <typename Cfg>
__global__ void __launch_bounds__(Cfg::kThreads, 1) hist_u8_planar_lane_tma_kernel(const uint8_t *__restrict__ img, int64_t hw, int64_t n_img, int64_t tiles_per_plane, unsigned long long *__restrict__ counts) {
    constexpr int kLaneWarps = Cfg::kWarps, kLaneThreads = Cfg::kThreads, kCountThreads = Cfg::kCountThreads, kLaneTileVecs = Cfg::kTileVecs, kLanePerThread = Cfg::kPerThread, kLaneStages = Cfg::kStages;
    extern __shared__ __align__(16) unsigned char smem_lane[];
    unsigned int *regions = reinterpret_cast<unsigned int *>(smem_lane);  // [warp][bin][lane]
    unsigned char *ring_mem = smem_lane + kLaneWarps * 32768;
    unsigned char *tail = ring_mem + kLaneStages * kLaneTileVecs * 16;
    unsigned int *hist32 = reinterpret_cast<unsigned int *>(tail + 2 * kLaneStages * 8);  // [bin]
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    TileRing<kLaneStages, kLaneTileVecs, kLaneWarps> tr;
    tr.init(ring_mem, reinterpret_cast<uint64_t *>(tail));
    for (int i = threadIdx.x; i < kLaneWarps * 8192; i += kLaneThreads) regions[i] = 0u;
    for (int i = threadIdx.x; i < 256; i += kLaneThreads) hist32[i] = 0u;
    __syncthreads();
    const bool counting = warp < kLaneWarps;  // false on the producer warp
    unsigned int *region = regions + (counting ? warp : 0) * 8192;
    const unsigned cbase = smem_u32(region) + (unsigned)lane * 4u;

    const int64_t per_channel = n_img * tiles_per_plane;
    const int64_t items = 3 * per_channel;
    const int64_t per_cta = (items + gridDim.x - 1) / gridDim.x;
    const int64_t first = (int64_t)blockIdx.x * per_cta;
    const int64_t last = first + per_cta < items ? first + per_cta : items;
    const uint4 *base = reinterpret_cast<const uint4 *>(img);

    // lane-private counters of every warp -> global counts of channel c, re-zero
    auto flush = [&](int c) {
        __syncthreads();
        for (int bin = lane; counting && bin < 256; bin += 32) {  // lane j: bins j, j + 32, ...; rotated walk = no conflicts
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
        for (int i = threadIdx.x; i < 256; i += kLaneThreads) {
            const unsigned v = hist32[i];
            if (v) atomicAdd(&counts[c * 256 + i], (unsigned long long)v);
            hist32[i] = 0u;
        }
        __syncthreads();
    };

    int64_t seg = first;
    while (seg < last) {  // one channel segment at a time (at most three per CTA)
        const int c = (int)(seg / per_channel);
        const int64_t chan_end = (int64_t)(c + 1) * per_channel;
        const int64_t seg_end = chan_end < last ? chan_end : last;
        const int n_items = (int)(seg_end - seg);
        TileCursor pc, cc;
        pc.seek(hw, (int)tiles_per_plane, kLaneTileVecs, per_channel, seg);
        cc = pc;
        if (Cfg::kProducer && !counting) {
            // producer warp: keep every free stage of the ring in flight (produce() waits for the
            // consumers to release the stage's previous tenant)
            if (lane == 0) {
                for (int item = 0; item < n_items; ++item) {
                    tr.produce(base + pc.off, (unsigned)pc.vecs(kLaneTileVecs) * 16u);
                    pc.next(kLaneTileVecs);
                }
            }
            __syncwarp();
        } else {
            int issued = 0;
            for (int item = 0; item < n_items; ++item) {
                if (!Cfg::kProducer && threadIdx.x == 0) {
                    while (issued < n_items && issued - item < kLaneStages - 1) {
                        tr.produce(base + pc.off, (unsigned)pc.vecs(kLaneTileVecs) * 16u);
                        pc.next(kLaneTileVecs);
                        ++issued;
                    }
                }
                const int nv = cc.vecs(kLaneTileVecs);
                cc.next(kLaneTileVecs);
                const uint4 *tile = tr.acquire();
                uint4 v[kLanePerThread];
                bool ok[kLanePerThread];
#pragma unroll
                for (int u = 0; u < kLanePerThread; ++u) {
                    const int idx = (int)threadIdx.x + u * kCountThreads;
                    ok[u] = idx < nv;
                    if (ok[u]) v[u] = tile[idx];
                }
                tr.release();
#pragma unroll
                for (int u = 0; u < kLanePerThread; ++u)
                    if (ok[u]) {
                        if constexpr (Cfg::kCount) { lane_count4(v[u].x, cbase); lane_count4(v[u].y, cbase); lane_count4(v[u].z, cbase); lane_count4(v[u].w, cbase); }
                        else if ((v[u].x ^ v[u].y ^ v[u].z ^ v[u].w) == 0x12345678u) lane_count4(v[u].x, cbase);
                    }
            }
        }
        flush(c);
        seg = seg_end;
    }
}

// The same kernel with WARP-GRANULAR tiles: tile g of the CTA's list goes to ring stage g % stages
// and is consumed by counting warp g % warps alone (the stage's "empty" barrier expects one
// arrival).  With CTA-wide tiles all counting warps run out of data at the same moment and the
// shared-memory atomic unit idles until the next tile lands; here the warps drift apart, so while
// one waits the others keep the unit busy.  A warp reads its tile straight from the ring, four
// 128-bit vectors per lane at a time.
template <typename Cfg>
__global__ void __launch_bounds__(Cfg::kThreads, 1) hist_u8_planar_lane_pw_kernel(const uint8_t *__restrict__ img, int64_t hw, int64_t n_img, int64_t tiles_per_plane, unsigned long long *__restrict__ counts) {
    constexpr int kW = Cfg::kWarps, kThreadsAll = Cfg::kThreads, kTV = Cfg::kTileVecs, kStages = Cfg::kStages;
    static_assert(Cfg::kProducer, "the per-warp kernel needs the producer warp");
    extern __shared__ __align__(16) unsigned char smem_lane[];
    unsigned int *regions = reinterpret_cast<unsigned int *>(smem_lane);  // [warp][bin][lane]
    unsigned char *ring_mem = smem_lane + kW * 32768;
    unsigned char *tail = ring_mem + kStages * kTV * 16;
    unsigned int *hist32 = reinterpret_cast<unsigned int *>(tail + 2 * kStages * 8);  // [bin]
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    pdl_trigger();  // the LUT kernel behind this one may become resident; it waits for our completion
    TileRing<kStages, kTV, 1> tr;
    tr.init(ring_mem, reinterpret_cast<uint64_t *>(tail));
    __syncthreads();  // barriers initialised: the producer starts streaming while the counters are zeroed
    const bool counting = warp < kW;
    unsigned int *region = regions + (counting ? warp : 0) * 8192;
    const unsigned cbase = smem_u32(region) + (unsigned)lane * 4u;
    if (counting) {  // a region is private to its warp until the first flush (which starts with a CTA barrier)
        uint4 *r4 = reinterpret_cast<uint4 *>(region);
        for (int i = lane; i < 2048; i += 32) r4[i] = make_uint4(0u, 0u, 0u, 0u);
        if (warp == 0)
            for (int i = lane; i < 256; i += 32) hist32[i] = 0u;
        __syncwarp();
    }

    const int64_t per_channel = n_img * tiles_per_plane;
    const int64_t items = 3 * per_channel;
    const int64_t per_cta = (items + gridDim.x - 1) / gridDim.x;
    const int64_t first = (int64_t)blockIdx.x * per_cta;
    const int64_t last = first + per_cta < items ? first + per_cta : items;
    const uint4 *base = reinterpret_cast<const uint4 *>(img);

    auto flush = [&](int c) {  // lane-private counters of every warp -> global counts of channel c, re-zero
        __syncthreads();
        for (int bin = lane; counting && bin < 256; bin += 32) {
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
        pdl_wait();  // `counts` is zeroed by the kernel in front of us when we run inside sx_hm_transform / sx_hm_fit
        for (int i = threadIdx.x; i < 256; i += kThreadsAll) {
            const unsigned v = hist32[i];
            if (v) atomicAdd(&counts[c * 256 + i], (unsigned long long)v);
            hist32[i] = 0u;
        }
        __syncthreads();
    };

    unsigned g0 = 0;  // running index of the segment's first tile
    int64_t seg = first;
    while (seg < last) {  // one channel segment at a time (at most three per CTA)
        const int c = (int)(seg / per_channel);
        const int64_t chan_end = (int64_t)(c + 1) * per_channel;
        const int64_t seg_end = chan_end < last ? chan_end : last;
        const int n_items = (int)(seg_end - seg);
        TileCursor cur;
        cur.seek(hw, (int)tiles_per_plane, kTV, per_channel, seg);
        if (warp == kW) {  // producer warp
            if (lane == 0) {
                for (int item = 0; item < n_items; ++item) {
                    tr.produce(base + cur.off, (unsigned)cur.vecs(kTV) * 16u);
                    cur.next(kTV);
                }
            }
            __syncwarp();
        } else if (counting) {
            for (int k = 0; k < warp; ++k) cur.next(kTV);
            for (int item = warp; item < n_items; item += kW) {
                const int nv = cur.vecs(kTV);
                const uint4 *tile = tr.acquire_at(g0 + (unsigned)item);
                for (int v0 = lane; v0 < nv; v0 += 128) {
                    uint4 v[4];
                    bool ok[4];
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
                        ok[j] = v0 + 32 * j < nv;
                        if (ok[j]) v[j] = tile[v0 + 32 * j];
                    }
#pragma unroll
                    for (int j = 0; j < 4; ++j)
                        if (ok[j]) { lane_count4(v[j].x, cbase); lane_count4(v[j].y, cbase); lane_count4(v[j].z, cbase); lane_count4(v[j].w, cbase); }
                }
                tr.release_at(g0 + (unsigned)item);
#pragma unroll
                for (int k = 0; k < kW; ++k) cur.next(kTV);
            }
        }
        g0 += (unsigned)n_items;
        flush(c);
        seg = seg_end;
    }
}

// Histogram, planar float32 / float16 / bfloat16: quantise, then warp-private shared atomics (12 or 6 B/px of
// traffic per 3 values, so the atomic rate is 4x / 2x lower than in the uint8 kernel).
template <typename T>
__global__ void __launch_bounds__(kThreads) hist_f32_planar_kernel(const T *__restrict__ img, int64_t hw, int64_t n_img, int64_t tiles_per_plane, unsigned long long *__restrict__ counts) {
    constexpr int kPer = FloatVec<T>::kPer;
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
        const T *plane = img + (n * 3 + c) * hw;
        const PlaneSplit sp = split_plane(plane, hw);
        const T *body = plane + sp.head;
        const int64_t v0 = t * kTileVecs;
        FloatVec<T> v[kUnroll];
        bool ok[kUnroll];
#pragma unroll
        for (int u = 0; u < kUnroll; ++u) {
            int64_t vi = v0 + u * kThreads + threadIdx.x;
            ok[u] = vi < sp.nvec;
            if (ok[u]) v[u].load(body + vi * kPer);
        }
#pragma unroll
        for (int u = 0; u < kUnroll; ++u) {
            if (ok[u]) {
#pragma unroll
                for (int k = 0; k < kPer; ++k) wa.add(quantize_u8(v[u].f[k]));
            }
        }
        if (t == 0) {
            int64_t ragged = sp.head + (hw - sp.tail0);
            if ((int64_t)threadIdx.x < ragged) {
                int64_t idx = (int64_t)threadIdx.x < sp.head ? (int64_t)threadIdx.x : sp.tail0 + ((int64_t)threadIdx.x - sp.head);
                wa.add(quantize_u8(FloatVec<T>::widen(plane[idx])));
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
            // L1-allocating loads: the three 128-bit loads of a warp touch the same 48 sectors (each lane's
            // 48 bytes start on a 16-byte boundary), so the second and third find their halves in L1
            // (96 -> 87 us per 64 x 1024^2 batch against L1::no_allocate loads)
            uint4 a = __ldg(p), b = __ldg(p + 1), d = __ldg(p + 2);
            unsigned w[12] = {a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w, d.x, d.y, d.z, d.w};
#pragma unroll
            for (int j = 0; j < 48; ++j) {
                unsigned v = (w[j >> 2] >> (8 * (j & 3))) & 0xffu;
                atomicAdd(&wh[(j % 3) * 256 + v], 1u);
            }
        } else {
            FloatVec<T> a, b, d;
            a.load(img + g * kGroup); b.load(img + g * kGroup + kPerVec); d.load(img + g * kGroup + 2 * kPerVec);
#pragma unroll
            for (int j = 0; j < kPerVec; ++j) {  // element e of the group has channel e % 3 (kGroup is a multiple of 3)
                atomicAdd(&wh[(j % 3) * 256 + quantize_u8(a.f[j])], 1u);
                atomicAdd(&wh[((kPerVec + j) % 3) * 256 + quantize_u8(b.f[j])], 1u);
                atomicAdd(&wh[((2 * kPerVec + j) % 3) * 256 + quantize_u8(d.f[j])], 1u);
            }
        }
    }
    // scalar remainder (everything when the base pointer is not 16-byte aligned)
    for (int64_t i = groups * kGroup + (int64_t)blockIdx.x * kThreads + threadIdx.x; i < total; i += (int64_t)gridDim.x * kThreads) {
        unsigned v;
        if constexpr (sizeof(T) == 1) v = img[i];
        else v = quantize_u8(FloatVec<T>::widen(img[i]));
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
    pdl_wait();  // launched with launch_pdl() behind the histogram kernel
    cf[b] = __ull2float_rn(__ldcg(counts + c * 256 + b));  // not an invariant load: see build_lut_kernel
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
struct LutSmem {  // shared scratch of one channel's LUT build
    double dacc[256];
    float rq[256];
    float sq[256];
};

// One channel (CTA of 256 threads): thread b holds the count of bin b.
template <bool FROM_HIST>
__device__ __forceinline__ void build_lut_channel(const int c, const unsigned long long my_count, long long npix, const float *__restrict__ ref, float *__restrict__ lut, LutSmem *ls) {
    float *rq = ls->rq, *sq = ls->sq;
    double *dacc = ls->dacc;
    __shared__ float s_npix_f;
    const int b = threadIdx.x;
    if (!FROM_HIST) rq[b] = __ldcg(ref + c * 256 + b);  // FROM_HIST: the caller has filled rq (ref_cdf_to_smem)
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

// Launched with launch_pdl().  FROM_HIST (only inside sx_hm_transform, whose first kernel is started in
// normal stream order, so `ref` is visible to every kernel of the chain): the reference CDF is built
// while the histogram kernel in front of us is still running; the counts are read after pdl_wait().
template <bool FROM_HIST>
__global__ void build_lut_kernel(const unsigned long long *__restrict__ counts, long long npix, const float *__restrict__ ref, float *__restrict__ lut) {
    __shared__ LutSmem ls;
    pdl_trigger();
    if (FROM_HIST) ref_cdf_to_smem(ref + blockIdx.x * 256, ls.sq, ls.dacc, ls.rq);
    pdl_wait();
    // __ldcg (a volatile, coherent load): a plain load through a const __restrict__ pointer is an
    // invariant (ld.global.nc) load to the compiler, which may hoist it above pdl_wait()
    build_lut_channel<FROM_HIST>(blockIdx.x, __ldcg(counts + blockIdx.x * 256 + threadIdx.x), npix, ref, lut, &ls);
}

// Zeroes the counts at the head of sx_hm_transform / sx_hm_fit (a kernel instead of a memset node so
// that the histogram kernel behind it can be a programmatic dependent launch).
__global__ void zero_counts_kernel(unsigned long long *__restrict__ counts) {
    pdl_trigger();
    counts[threadIdx.x] = 0ull;
}

// ------------------------------------------------------------------------------------------------
// Sharded batches: the all-reduce of the counts FUSED into the LUT build, over NVLink peer memory.
//
// Every rank keeps its counts in a buffer that all ranks of the node have mapped (symmetric
// memory); bufs[p] is rank p's buffer as seen from this GPU:
//     uint64 counts[2][3][256]   (two parities: a rank may start the next step's histogram while a
//                                 slower peer still reads this step's counts)
//     uint32 flags[64]           flags[q] = last epoch whose counts rank q has published to us
// One kernel per rank and step: (1) publish: a system-scope fence, then store-release the epoch
// into flags[rank] of every peer; (2) wait until every peer's epoch has arrived here (acquire);
// (3) every thread adds its bin over all ranks with peer loads in rank order -- integers, so every
// rank gets bit-identical sums; (4) the bit-exact LUT build.  No NCCL call, no extra launch: the
// exchange costs one NVLink round trip inside a kernel that had to run anyway (~25 us of NCCL
// all-reduce + launch per step saved at 8 GPUs).
// Ranks run on different GPUs, so the spin in (2) waits for work that is already running or queued
// on its own device; it never waits for a kernel behind it on the same device.
// ------------------------------------------------------------------------------------------------
constexpr int kPeerCountsBytes = 2 * 768 * 8;
constexpr int kPeerMaxWorld = 64;

__global__ void __launch_bounds__(256) build_lut_peers_kernel(unsigned char *const *__restrict__ bufs, int world, int rank, unsigned epoch, const float *__restrict__ ref_cdf, float *__restrict__ lut, unsigned long long *__restrict__ counts_out, unsigned long long budget_ns, unsigned *status) {
    const int c = blockIdx.x, b = threadIdx.x;
    const int parity = (int)(epoch & 1u);
    pdl_trigger();
    pdl_wait();  // launched with launch_pdl(): this rank's counts come from the kernel in front of us
    if (c == 0 && b < world) {  // (1) publish my counts (written by earlier kernels of this stream)
        __threadfence_system();
        unsigned *flag = reinterpret_cast<unsigned *>(bufs[b] + kPeerCountsBytes) + rank;
        asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(flag), "r"(epoch) : "memory");
    }
    if (b < world) {  // (2) wait for every rank's counts of this epoch (bounded: a dead rank must not hang the node)
        const unsigned *flag = reinterpret_cast<const unsigned *>(bufs[rank] + kPeerCountsBytes) + b;
        wait_peer_flag(flag, epoch, budget_ns, status, b, rank);
    }
    __syncthreads();
    unsigned long long total = 0;  // (3) all-reduce of bin (c, b)
    for (int p = 0; p < world; ++p) {
        const unsigned long long *src = reinterpret_cast<const unsigned long long *>(bufs[p]) + parity * 768 + c * 256 + b;
        unsigned long long v;
        asm volatile("ld.relaxed.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(src) : "memory");
        total += v;
    }
    if (counts_out != nullptr) counts_out[c * 256 + b] = total;
    __shared__ LutSmem ls;
    build_lut_channel<false>(c, total, -1, ref_cdf, lut, &ls);  // (4); pixel count = sum of the channel's counts
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

// uint8 planar with whole 128-bit vectors per plane (16-byte aligned base, H*W % 16 == 0): the same
// remap with a third of the instructions.  The general kernel above spends ~85 instructions per
// 128-bit vector (64-bit plane / channel arithmetic per item, shift + mask + multiply-add + load per
// byte, an alignment branch with a byte-store fall-back) and runs at 64 % issue utilisation; here
// the channel's 256-byte table is 256-byte aligned in shared memory, so ONE PRMT builds the full
// lookup address (table address with its low byte replaced by the pixel byte), and three PRMTs
// reassemble a word: 11 instructions per four pixels.
__device__ __forceinline__ unsigned remap4_prmt(unsigned w, unsigned tab_addr) {  // tab_addr: shared-space address, multiple of 256
    unsigned t0, t1, t2, t3;
    asm volatile("ld.shared.u8 %0, [%1];" : "=r"(t0) : "r"(__byte_perm(w, tab_addr, 0x7650)));
    asm volatile("ld.shared.u8 %0, [%1];" : "=r"(t1) : "r"(__byte_perm(w, tab_addr, 0x7651)));
    asm volatile("ld.shared.u8 %0, [%1];" : "=r"(t2) : "r"(__byte_perm(w, tab_addr, 0x7652)));
    asm volatile("ld.shared.u8 %0, [%1];" : "=r"(t3) : "r"(__byte_perm(w, tab_addr, 0x7653)));
    return __byte_perm(__byte_perm(t0, t1, 0x0040), __byte_perm(t2, t3, 0x0040), 0x5410);
}

// Launched with launch_pdl().  `chain` != 0 (inside sx_hm_transform: the images were visible before the
// first kernel of the chain started): the CTA's first tile is loaded BEFORE pdl_wait(), i.e. while the
// LUT kernel -- and the tail of the histogram kernel -- in front of us are still running.
__global__ void __launch_bounds__(kThreads) apply_u8_planar_vec_kernel(const uint4 *__restrict__ src, uint4 *__restrict__ dst, unsigned vecs, unsigned planes, unsigned tiles_per_plane, const float *__restrict__ lut, int chain) {
    __shared__ __align__(256) unsigned char lut8[3 * 256];
    const unsigned items = planes * tiles_per_plane;  // < 2^31 (checked by the caller)
    pdl_trigger();
    if (!chain) pdl_wait();
    uint4 v[kUnroll];
    auto load_item = [&](unsigned it) {
        const unsigned item = items - 1 - it;
        const unsigned pl = item / tiles_per_plane, t = item - pl * tiles_per_plane;
        const size_t base = (size_t)pl * vecs;
        const unsigned v0 = t * kTileVecs + threadIdx.x;
#pragma unroll
        for (int u = 0; u < kUnroll; ++u)
            if (v0 + u * kThreads < vecs) v[u] = ld_stream(src + base + v0 + u * kThreads);
    };
    if (blockIdx.x < items) load_item(blockIdx.x);
    if (chain) pdl_wait();
    for (int i = threadIdx.x; i < 768; i += kThreads) lut8[i] = (unsigned char)__float2int_rz(__ldcg(lut + i));  // trunc, L296-298; not an invariant load (see build_lut_kernel)
    __syncthreads();
    const unsigned tab0 = smem_u32(lut8);
    for (unsigned it = blockIdx.x; it < items; it += gridDim.x) {
        const unsigned item = items - 1 - it;  // from the end of the batch (what a front-to-back histogram left in L2)
        const unsigned pl = item / tiles_per_plane, t = item - pl * tiles_per_plane;
        const unsigned tab = tab0 + (pl % 3u) * 256u;
        const size_t base = (size_t)pl * vecs;
        const unsigned v0 = t * kTileVecs + threadIdx.x;
        if (it != blockIdx.x) load_item(it);
#pragma unroll
        for (int u = 0; u < kUnroll; ++u) {
            if (v0 + u * kThreads < vecs) {
                uint4 o;
                o.x = remap4_prmt(v[u].x, tab); o.y = remap4_prmt(v[u].y, tab); o.z = remap4_prmt(v[u].z, tab); o.w = remap4_prmt(v[u].w, tab);
                st_stream(dst + base + v0 + u * kThreads, o);
            }
        }
    }
}

// (A 65 536-entry two-byte LUT in shared memory behind a TMA ring was built and measured for the uint8
// planar remap: 80 us against 72 us for the kernel above; removed.)
// float32 / float16 / bfloat16 planar: out = clamp(lut[trunc(clamp(255 x))] / 255, 0, 1)   (L290-296), stored in
// the input's dtype (round to nearest even for the 16-bit types, the reference's cast-back of L131).
template <typename T>
__global__ void __launch_bounds__(kThreads) apply_f32_planar_kernel(const T *__restrict__ img, T *__restrict__ out, int64_t hw, int64_t planes, int64_t tiles_per_plane, const float *__restrict__ lut) {
    constexpr int kPer = FloatVec<T>::kPer;
    __shared__ float lutf[3 * 256];
    for (int i = threadIdx.x; i < 768; i += kThreads) lutf[i] = fminf(fmaxf(__fdiv_rn(lut[i], 255.0f), 0.0f), 1.0f);
    __syncthreads();
    const int64_t items = planes * tiles_per_plane;
    for (int64_t it = blockIdx.x; it < items; it += gridDim.x) {
        const int64_t item = items - 1 - it;
        const int64_t pl = item / tiles_per_plane;
        const int64_t t = item - pl * tiles_per_plane;
        const float *l = lutf + (int)(pl % 3) * 256;
        const T *src = img + pl * hw;
        T *dst = out + pl * hw;
        const PlaneSplit sp = split_plane(src, hw);
        const bool dst_vec = ((reinterpret_cast<uintptr_t>(dst + sp.head)) & 15) == 0;
        const T *body = src + sp.head;
        const int64_t v0 = t * kTileVecs;
        FloatVec<T> v[kUnroll];
        bool ok[kUnroll];
#pragma unroll
        for (int u = 0; u < kUnroll; ++u) {
            int64_t vi = v0 + u * kThreads + threadIdx.x;
            ok[u] = vi < sp.nvec;
            if (ok[u]) v[u].load(body + vi * kPer);
        }
#pragma unroll
        for (int u = 0; u < kUnroll; ++u) {
            if (!ok[u]) continue;
            int64_t vi = v0 + u * kThreads + threadIdx.x;
#pragma unroll
            for (int k = 0; k < kPer; ++k) v[u].f[k] = l[quantize_u8(v[u].f[k])];
            T *d = dst + sp.head + vi * kPer;
            if (dst_vec) {
                v[u].store(d);
            } else {
#pragma unroll
                for (int k = 0; k < kPer; ++k) d[k] = FloatVec<T>::narrow(v[u].f[k]);
            }
        }
        if (t == 0) {
            int64_t ragged = sp.head + (hw - sp.tail0);
            if ((int64_t)threadIdx.x < ragged) {
                int64_t idx = (int64_t)threadIdx.x < sp.head ? (int64_t)threadIdx.x : sp.tail0 + ((int64_t)threadIdx.x - sp.head);
                dst[idx] = FloatVec<T>::narrow(l[quantize_u8(FloatVec<T>::widen(src[idx]))]);
            }
        }
    }
}

// Interleaved (NHWC) uint8 remap over whole 128-bit vectors of the flat array: a warp's load / store is one
// contiguous 512-byte run (the general kernel below gives every thread 48 consecutive bytes, so that a
// byte's channel is a compile-time constant -- but then each 128-bit access of a warp touches 32 half-used
// sectors: 125 us per 64 x 1024^2 batch whatever the instruction count).  Here byte b of vector v belongs to
// channel (v + b) % 3 (16 = 1 mod 3): the three table addresses are rotated once per vector and the
// lookups are the PRMT-addressed ones of the planar kernel.
__global__ void __launch_bounds__(kThreads) apply_u8_nhwc_vec_kernel(const uint4 *__restrict__ src, uint4 *__restrict__ dst, unsigned nvec, unsigned tiles, const float *__restrict__ lut) {
    __shared__ __align__(256) unsigned char lut8[3 * 256];
    for (int i = threadIdx.x; i < 768; i += kThreads) lut8[i] = (unsigned char)__float2int_rz(lut[i]);  // trunc, L296-298
    __syncthreads();
    const unsigned tab0 = smem_u32(lut8);
    static_assert(kTileVecs % 3 == 1 && kThreads % 3 == 1, "the channel rotation below assumes 1024 = 256 = 1 (mod 3)");
    for (unsigned t = blockIdx.x; t < tiles; t += gridDim.x) {
        const unsigned v0 = t * kTileVecs + threadIdx.x;
        // channel of byte 0 of vector v0 = v0 % 3 = (t + threadIdx.x) % 3; vector v0 + u * 256 is u channels further
        const unsigned c0 = (t + threadIdx.x) % 3u;
        const unsigned c1 = c0 == 2u ? 0u : c0 + 1u, c2 = c0 == 0u ? 2u : c0 - 1u;
        const unsigned tab[3] = {tab0 + c0 * 256u, tab0 + c1 * 256u, tab0 + c2 * 256u};
        uint4 v[kUnroll];
#pragma unroll
        for (int u = 0; u < kUnroll; ++u)
            if (v0 + u * kThreads < nvec) v[u] = ld_stream(src + v0 + u * kThreads);
#pragma unroll
        for (int u = 0; u < kUnroll; ++u) {
            const unsigned vi = v0 + u * kThreads;
            if (vi < nvec) {
                const unsigned w[4] = {v[u].x, v[u].y, v[u].z, v[u].w};
                unsigned r[4];
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                    unsigned tt[4];
#pragma unroll
                    for (int i = 0; i < 4; ++i) asm volatile("ld.shared.u8 %0, [%1];" : "=r"(tt[i]) : "r"(__byte_perm(w[k], tab[(u + 4 * k + i) % 3], 0x7650 + i)));
                    r[k] = __byte_perm(__byte_perm(tt[0], tt[1], 0x0040), __byte_perm(tt[2], tt[3], 0x0040), 0x5410);
                }
                st_stream(dst + vi, make_uint4(r[0], r[1], r[2], r[3]));
            }
        }
    }
}

// The < 16 bytes behind the last whole vector of an interleaved uint8 array; byte i has channel (c0 + i) % 3.
__global__ void apply_nhwc_tail_kernel(const uint8_t *__restrict__ img, uint8_t *__restrict__ out, int n, int c0, const float *__restrict__ lut) {
    const int i = threadIdx.x;
    if (i < n) out[i] = (uint8_t)__float2int_rz(lut[((c0 + i) % 3) * 256 + img[i]]);
}

// Interleaved (NHWC) remap, uint8 or float32.
template <typename T>
__global__ void __launch_bounds__(kThreads) apply_nhwc_kernel(const T *__restrict__ img, T *__restrict__ out, int64_t total, const float *__restrict__ lut) {
    constexpr int kPerVec = 16 / sizeof(T);
    constexpr int kGroup = 3 * kPerVec;
    __shared__ float lutf[3 * 256];
    __shared__ __align__(256) unsigned char lut8[3 * 256];  // 256-byte aligned tables: PRMT-built lookup addresses (remap4_prmt)
    for (int i = threadIdx.x; i < 768; i += kThreads) {
        lut8[i] = (unsigned char)__float2int_rz(lut[i]);
        lutf[i] = fminf(fmaxf(__fdiv_rn(lut[i], 255.0f), 0.0f), 1.0f);
    }
    __syncthreads();
    const unsigned tab0 = smem_u32(lut8);
    const bool vec_ok = ((reinterpret_cast<uintptr_t>(img) | reinterpret_cast<uintptr_t>(out)) & 15) == 0;
    const int64_t groups = vec_ok ? total / kGroup : 0;
    for (int64_t g = (int64_t)blockIdx.x * kThreads + threadIdx.x; g < groups; g += (int64_t)gridDim.x * kThreads) {
        if constexpr (sizeof(T) == 1) {
            const uint4 *p = reinterpret_cast<const uint4 *>(img + g * kGroup);
            uint4 a = __ldg(p), b = __ldg(p + 1), d = __ldg(p + 2);  // L1-allocating: see hist_nhwc_kernel
            unsigned w[12] = {a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w, d.x, d.y, d.z, d.w};
            unsigned r[12];
#pragma unroll
            for (int k = 0; k < 12; ++k) {
                // byte j of the 48-byte group belongs to channel j % 3: one PRMT per byte builds the address of
                // its entry in that channel's table, three more reassemble the word
                unsigned t[4];
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    const unsigned tab = tab0 + (unsigned)((4 * k + i) % 3) * 256u;
                    asm volatile("ld.shared.u8 %0, [%1];" : "=r"(t[i]) : "r"(__byte_perm(w[k], tab, 0x7650 + i)));
                }
                r[k] = __byte_perm(__byte_perm(t[0], t[1], 0x0040), __byte_perm(t[2], t[3], 0x0040), 0x5410);
            }
            uint4 *q = reinterpret_cast<uint4 *>(out + g * kGroup);
            st_stream(q, make_uint4(r[0], r[1], r[2], r[3]));
            st_stream(q + 1, make_uint4(r[4], r[5], r[6], r[7]));
            st_stream(q + 2, make_uint4(r[8], r[9], r[10], r[11]));
        } else {
            FloatVec<T> a, b, d;
            a.load(img + g * kGroup); b.load(img + g * kGroup + kPerVec); d.load(img + g * kGroup + 2 * kPerVec);
#pragma unroll
            for (int j = 0; j < kPerVec; ++j) {
                a.f[j] = lutf[(j % 3) * 256 + quantize_u8(a.f[j])];
                b.f[j] = lutf[((kPerVec + j) % 3) * 256 + quantize_u8(b.f[j])];
                d.f[j] = lutf[((2 * kPerVec + j) % 3) * 256 + quantize_u8(d.f[j])];
            }
            a.store(out + g * kGroup); b.store(out + g * kGroup + kPerVec); d.store(out + g * kGroup + 2 * kPerVec);
        }
    }
    for (int64_t i = groups * kGroup + (int64_t)blockIdx.x * kThreads + threadIdx.x; i < total; i += (int64_t)gridDim.x * kThreads) {
        if constexpr (sizeof(T) == 1) out[i] = lut8[(int)(i % 3) * 256 + img[i]];
        else out[i] = FloatVec<T>::narrow(lutf[(int)(i % 3) * 256 + quantize_u8(FloatVec<T>::widen(img[i]))]);
    }
}

// ---- tuning knobs (A/B measurements; defaults are the measured winners) -----------------------
static int g_hist_byte_counters = 5;  // uint8 planar histogram: 5 lane-private counters fed by a TMA ring, warp-granular tiles (default
                                      // for 16-byte aligned planes from 8 MB on: data-independent), 9 the same with CTA-wide tiles,
                                      // 0 warp-private atomics (any alignment; faster on constant images, slower on noise),
                                      // 6 ring feed without counting (measurement only: wrong counts)
static int g_hist_ctas_per_sm = 8;
static int g_apply_ctas_per_sm = 16;

}  // namespace hm
}  // namespace sx

using namespace sx;
using namespace sx::hm;

template <typename Cfg>
static int launch_lane_tma_cfg(const uint8_t *images, int64_t hw, int64_t n, unsigned long long *cnt, cudaStream_t stream) {
    if (int rc = allow_big_smem(hist_u8_planar_lane_tma_kernel<Cfg>, Cfg::kSmem)) return rc;
    const int64_t tiles = max_i64(1, (hw / 16 + Cfg::kTileVecs - 1) / Cfg::kTileVecs);
    const unsigned grid = stream_grid(3 * n * tiles, 1);
    hist_u8_planar_lane_tma_kernel<Cfg><<<grid, Cfg::kThreads, Cfg::kSmem, stream>>>(images, hw, n, tiles, cnt);
    return SX_OK;
}
template <typename Cfg>
static int launch_lane_pw_cfg(const uint8_t *images, int64_t hw, int64_t n, unsigned long long *cnt, cudaStream_t stream, bool pdl) {
    if (int rc = allow_big_smem(hist_u8_planar_lane_pw_kernel<Cfg>, Cfg::kSmem)) return rc;
    const int64_t tiles = max_i64(1, (hw / 16 + Cfg::kTileVecs - 1) / Cfg::kTileVecs);
    const unsigned grid = stream_grid(3 * n * tiles, 1);
    // (Register-fed lane-private counters were measured again in this form: five counting warps, no ring, every
    // lane keeping a batch of sixteen 128-bit loads in flight while it counts the previous batch -- 40 KB in
    // flight per SM, 164 KB carve-out.  Correct, but 75.4 us against 59.7 us: the loads share the LSU queue with
    // the shared-memory atomics, the TMA ring does not.)
    // (L2 eviction priorities were measured for this chain -- the tail of every CTA's range loaded with
    // evict_last, 32-100 MB in total, the rest and all remap traffic with evict_first: 127.1-127.4 us against
    // 127.1 us without any hint.  The remap runs at the speed of a device copy whether or not part of its
    // input is still in L2.)
    if (pdl) SX_CUDA(launch_pdl(hist_u8_planar_lane_pw_kernel<Cfg>, dim3(grid), dim3(Cfg::kThreads), (size_t)Cfg::kSmem, stream, images, hw, n, tiles, cnt));
    else hist_u8_planar_lane_pw_kernel<Cfg><<<grid, Cfg::kThreads, Cfg::kSmem, stream>>>(images, hw, n, tiles, cnt);
    return SX_OK;
}
static int launch_lane_tma(int mode, const uint8_t *images, int64_t hw, int64_t n, unsigned long long *cnt, cudaStream_t stream, bool pdl) {
    // measured on B200, 64 x 3 x 1024^2 noise.  CTA-wide tiles: <5 warps, 16 KB x 4, no producer> 80 us; <4, 24 KB x 4> 69 us;
    // <4 + producer, 24 KB x 4> 62 us; <4 + producer, 12 KB x 8> 67 us; <3 + producer, 32 KB x 4> 75 us.
    // Warp-granular tiles: <4 + producer, 12 KB x 8> 59.7 us (default); <4 + producer, 6 KB x 16> 62.8 us.
    // Floor of this design: ATOMS (2.1 clk per warp instruction) + TMA writes + LDS reads share the SM's
    // one-wavefront-per-clock shared-memory pipe: ~110 K clk per SM = 56 us.
    if (mode == 6) return launch_lane_tma_cfg<LaneCfg<4, 1536, 4, false, true>>(images, hw, n, cnt, stream);  // feed rate only: 40 us
    if (mode == 9) return launch_lane_tma_cfg<LaneCfg<4, 1536, 4, true, true>>(images, hw, n, cnt, stream);   // CTA-wide tiles
    // (Building the LUT in the tail of this kernel -- last CTA, CTA padded to 256 threads -- was measured: the
    // padded kernel is 6 us slower and the serial three-channel build costs more than the saved launch.)
    // Smaller rings that leave shared memory for early-resident CTAs of the remap kernel behind this one
    // (programmatic dependent launch) were measured: 640 / 576 / 448 / 384 vectors x 8 stages make the
    // transform 4-8 us SLOWER -- a remap CTA that becomes resident next to this kernel pins the SM's
    // 196-228 KB carve-out for the whole remap, whose loads in flight then do not fit the remaining L1.
    if (mode == 8) return launch_lane_pw_cfg<LaneCfg<4, 512, 8, true, true>>(images, hw, n, cnt, stream, pdl);  // 64 KB ring (196 KB carve-out): 61.8 us
    return launch_lane_pw_cfg<LaneCfg<4, 768, 8, true, true>>(images, hw, n, cnt, stream, pdl);
}

extern "C" {

// Development hook (declared in the header under "development hooks"): process-global, not
// thread-safe, and inert unless SX_ENABLE_TUNING=1 is set in the environment.
int sx_hm_set_tuning(int hist_byte_counters, int hist_ctas_per_sm, int apply_ctas_per_sm) {
    if (!tuning_enabled()) return sx::fail(SX_ERR_UNSUPPORTED, "tuning hooks are disabled (set SX_ENABLE_TUNING=1 before loading the library)");
    if (hist_byte_counters >= 0) g_hist_byte_counters = hist_byte_counters;
    if (hist_ctas_per_sm > 0) g_hist_ctas_per_sm = hist_ctas_per_sm;
    if (apply_ctas_per_sm > 0) g_apply_ctas_per_sm = apply_ctas_per_sm;
    return SX_OK;
}

// `chain`: called from sx_hm_transform / sx_hm_fit right behind zero_counts_kernel (which was started in
// normal stream order): the persistent histogram kernel is a programmatic dependent launch.
static int hist_impl(const void *images, int dtype, int layout, int64_t n, int64_t h, int64_t w, uint64_t *counts, sx_stream_t stream_, bool chain) {
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
            prefer_l1(hist_nhwc_kernel<uint8_t>, kThreads);
            hist_nhwc_kernel<uint8_t><<<grid, kThreads, 0, stream>>>(static_cast<const uint8_t *>(images), total, cnt);
        } else {
            unsigned grid = stream_grid((total / 12 + kThreads - 1) / kThreads + 1, 8);
            if (dtype == SX_F32) {
                prefer_l1(hist_nhwc_kernel<float>, kThreads);
                hist_nhwc_kernel<float><<<grid, kThreads, 0, stream>>>(static_cast<const float *>(images), total, cnt);
            } else if (dtype == SX_F16) {
                prefer_l1(hist_nhwc_kernel<__half>, kThreads);
                hist_nhwc_kernel<__half><<<grid, kThreads, 0, stream>>>(static_cast<const __half *>(images), total, cnt);
            } else {
                prefer_l1(hist_nhwc_kernel<__nv_bfloat16>, kThreads);
                hist_nhwc_kernel<__nv_bfloat16><<<grid, kThreads, 0, stream>>>(static_cast<const __nv_bfloat16 *>(images), total, cnt);
            }
        }
        SX_LAUNCHED("hist_nhwc_kernel");
        return SX_OK;
    }
    if (dtype == SX_U8) {
        const int64_t tiles = max_i64(1, (hw / 16 + kTileVecs - 1) / kTileVecs);
        const int64_t items = n * tiles;
        // the TMA-fed kernel needs whole 128-bit vectors per plane; anything else takes the general kernel
        const bool planes_aligned = aligned16(images) && hw % 16 == 0;
        // the persistent lane-private kernel zeroes 128 KB of counters per SM: worth it from ~8 MB of pixels on
        if (g_hist_byte_counters >= 5 && planes_aligned && n * hw * 3 >= ((int64_t)8 << 20)) {
            if (int rc = launch_lane_tma(g_hist_byte_counters, static_cast<const uint8_t *>(images), hw, n, cnt, stream, chain)) return rc;
        } else {
            dim3 grid(stream_grid(items, g_hist_ctas_per_sm), 3);
            grid.x = (grid.x + 2) / 3 > 0 ? (grid.x + 2) / 3 : 1;
            prefer_l1(hist_u8_planar_kernel, kThreads);
            hist_u8_planar_kernel<<<grid, kThreads, 0, stream>>>(static_cast<const uint8_t *>(images), hw, n, tiles, cnt);
        }
        SX_LAUNCHED("hist_u8_planar_kernel");
    } else {
        const int64_t per = 16 / dtype_bytes(dtype);
        const int64_t tiles = max_i64(1, (hw / per + kTileVecs - 1) / kTileVecs);
        dim3 grid(stream_grid(n * tiles, 6), 3);
        grid.x = (grid.x + 2) / 3 > 0 ? (grid.x + 2) / 3 : 1;
        if (dtype == SX_F32) {
            prefer_l1(hist_f32_planar_kernel<float>, kThreads);
            hist_f32_planar_kernel<float><<<grid, kThreads, 0, stream>>>(static_cast<const float *>(images), hw, n, tiles, cnt);
        } else if (dtype == SX_F16) {
            prefer_l1(hist_f32_planar_kernel<__half>, kThreads);
            hist_f32_planar_kernel<__half><<<grid, kThreads, 0, stream>>>(static_cast<const __half *>(images), hw, n, tiles, cnt);
        } else {
            prefer_l1(hist_f32_planar_kernel<__nv_bfloat16>, kThreads);
            hist_f32_planar_kernel<__nv_bfloat16><<<grid, kThreads, 0, stream>>>(static_cast<const __nv_bfloat16 *>(images), hw, n, tiles, cnt);
        }
        SX_LAUNCHED("hist_f32_planar_kernel");
    }
    return SX_OK;
}

int sx_hm_hist(const void *images, int dtype, int layout, int64_t n, int64_t h, int64_t w, uint64_t *counts, sx_stream_t stream) {
    return hist_impl(images, dtype, layout, n, h, w, counts, stream, false);
}

int sx_hm_ref_hist(const uint64_t *counts, float *ref_hist, sx_stream_t stream) {
    SX_REQUIRE(counts && ref_hist, "NULL argument");
    SX_CUDA(launch_pdl(ref_hist_kernel, dim3(3), dim3(256), 0, static_cast<cudaStream_t>(stream), reinterpret_cast<const unsigned long long *>(counts), ref_hist));
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
    SX_CUDA(launch_pdl(build_lut_kernel<false>, dim3(3), dim3(256), 0, static_cast<cudaStream_t>(stream), reinterpret_cast<const unsigned long long *>(counts), (long long)npix, ref_cdf, lut));
    SX_LAUNCHED("build_lut_kernel");
    return SX_OK;
}

static int apply_impl(const void *images, int dtype, int layout, int64_t n, int64_t h, int64_t w, const float *lut, void *out, sx_stream_t stream_, bool chain) {
    if (int rc = check_images(images, dtype, n, h, w)) return rc;
    SX_REQUIRE(layout == SX_NCHW || layout == SX_NHWC, "layout must be SX_NCHW or SX_NHWC, got %d", layout);
    cudaStream_t stream = static_cast<cudaStream_t>(stream_);
    const int64_t hw = h * w;
    if (n == 0 || hw == 0) return SX_OK;
    SX_REQUIRE(lut && out, "NULL argument");
    if (layout == SX_NHWC) {
        const int64_t total = n * hw * 3;
        if (dtype == SX_U8 && aligned16(images) && aligned16(out) && total / 16 < ((int64_t)1 << 31) && total >= ((int64_t)1 << 20)) {
            // whole vectors with the coalesced kernel; the < 16 trailing bytes (total = 3 N H W is a multiple of
            // 48 whenever H W is a multiple of 16) with the general kernel, whose channel is the byte index mod 3
            const int64_t nvec = total / 16, tiles = (nvec + kTileVecs - 1) / kTileVecs;
            prefer_l1(apply_u8_nhwc_vec_kernel, kThreads);
            apply_u8_nhwc_vec_kernel<<<stream_grid(tiles, g_apply_ctas_per_sm), kThreads, 0, stream>>>(static_cast<const uint4 *>(images), static_cast<uint4 *>(out), (unsigned)nvec, (unsigned)tiles, lut);
            const int64_t tail = total - nvec * 16;
            if (tail > 0) {
                SX_LAUNCHED("apply_u8_nhwc_vec_kernel");
                // the tail starts at byte nvec * 16, whose channel is (nvec * 16) % 3 = nvec % 3: shift the LUT rows accordingly
                apply_nhwc_tail_kernel<<<1, 32, 0, stream>>>(static_cast<const uint8_t *>(images) + nvec * 16, static_cast<uint8_t *>(out) + nvec * 16, (int)tail, (int)(nvec % 3), lut);
            }
        } else if (dtype == SX_U8) {
            unsigned grid = stream_grid((total / 48 + kThreads - 1) / kThreads + 1, 8);
            prefer_l1(apply_nhwc_kernel<uint8_t>, kThreads);
            apply_nhwc_kernel<uint8_t><<<grid, kThreads, 0, stream>>>(static_cast<const uint8_t *>(images), static_cast<uint8_t *>(out), total, lut);
        } else {
            unsigned grid = stream_grid((total / 12 + kThreads - 1) / kThreads + 1, 8);
            if (dtype == SX_F32) {
                prefer_l1(apply_nhwc_kernel<float>, kThreads);
                apply_nhwc_kernel<float><<<grid, kThreads, 0, stream>>>(static_cast<const float *>(images), static_cast<float *>(out), total, lut);
            } else if (dtype == SX_F16) {
                prefer_l1(apply_nhwc_kernel<__half>, kThreads);
                apply_nhwc_kernel<__half><<<grid, kThreads, 0, stream>>>(static_cast<const __half *>(images), static_cast<__half *>(out), total, lut);
            } else {
                prefer_l1(apply_nhwc_kernel<__nv_bfloat16>, kThreads);
                apply_nhwc_kernel<__nv_bfloat16><<<grid, kThreads, 0, stream>>>(static_cast<const __nv_bfloat16 *>(images), static_cast<__nv_bfloat16 *>(out), total, lut);
            }
        }
        SX_LAUNCHED("apply_nhwc_kernel");
        return SX_OK;
    }
    const int64_t planes = n * 3;
    const int64_t tiles_v = max_i64(1, (hw / 16 + kTileVecs - 1) / kTileVecs);
    if (dtype == SX_U8 && aligned16(images) && aligned16(out) && hw % 16 == 0 && planes * tiles_v < ((int64_t)1 << 31) && hw / 16 < ((int64_t)1 << 31)) {
        unsigned grid = stream_grid(planes * tiles_v, g_apply_ctas_per_sm);
        prefer_l1(apply_u8_planar_vec_kernel, kThreads);
        SX_CUDA(launch_pdl(apply_u8_planar_vec_kernel, dim3(grid), dim3(kThreads), 0, stream, static_cast<const uint4 *>(images), static_cast<uint4 *>(out), (unsigned)(hw / 16), (unsigned)planes, (unsigned)tiles_v, lut, chain ? 1 : 0));
        SX_LAUNCHED("apply_u8_planar_vec_kernel");
    } else if (dtype == SX_U8) {
        const int64_t tiles = max_i64(1, (hw / 16 + kTileVecs - 1) / kTileVecs);
        unsigned grid = stream_grid(planes * tiles, g_apply_ctas_per_sm);
        prefer_l1(apply_u8_planar_kernel, kThreads);
        apply_u8_planar_kernel<<<grid, kThreads, 0, stream>>>(static_cast<const uint8_t *>(images), static_cast<uint8_t *>(out), hw, planes, tiles, lut);
        SX_LAUNCHED("apply_u8_planar_kernel");
    } else {
        const int64_t per = 16 / dtype_bytes(dtype);
        const int64_t tiles = max_i64(1, (hw / per + kTileVecs - 1) / kTileVecs);
        unsigned grid = stream_grid(planes * tiles, g_apply_ctas_per_sm);
        if (dtype == SX_F32) {
            prefer_l1(apply_f32_planar_kernel<float>, kThreads);
            apply_f32_planar_kernel<float><<<grid, kThreads, 0, stream>>>(static_cast<const float *>(images), static_cast<float *>(out), hw, planes, tiles, lut);
        } else if (dtype == SX_F16) {
            prefer_l1(apply_f32_planar_kernel<__half>, kThreads);
            apply_f32_planar_kernel<__half><<<grid, kThreads, 0, stream>>>(static_cast<const __half *>(images), static_cast<__half *>(out), hw, planes, tiles, lut);
        } else {
            prefer_l1(apply_f32_planar_kernel<__nv_bfloat16>, kThreads);
            apply_f32_planar_kernel<__nv_bfloat16><<<grid, kThreads, 0, stream>>>(static_cast<const __nv_bfloat16 *>(images), static_cast<__nv_bfloat16 *>(out), hw, planes, tiles, lut);
        }
        SX_LAUNCHED("apply_f32_planar_kernel");
    }
    return SX_OK;
}

int sx_hm_apply(const void *images, int dtype, int layout, int64_t n, int64_t h, int64_t w, const float *lut, void *out, sx_stream_t stream) {
    return apply_impl(images, dtype, layout, n, h, w, lut, out, stream, false);
}

int64_t sx_hm_peer_buffer_bytes(void) { return kPeerCountsBytes + kPeerMaxWorld * 4; }

int sx_hm_build_lut_peers(const void *peer_buffers_dev, int world, int rank, uint32_t epoch, const float *ref_cdf, float *lut, uint64_t *counts_out, sx_stream_t stream) {
    SX_REQUIRE(peer_buffers_dev && ref_cdf && lut, "NULL argument");
    SX_REQUIRE(world >= 1 && world <= kPeerMaxWorld && rank >= 0 && rank < world, "bad rank/world (%d, %d)", rank, world);
    SX_REQUIRE(epoch != 0, "epoch must start at 1 (flags are zero-initialised)");
    if (int rc = peer_status_check("sx_hm_build_lut_peers")) return rc;
    SX_CUDA(launch_pdl(build_lut_peers_kernel, dim3(3), dim3(256), 0, static_cast<cudaStream_t>(stream), static_cast<unsigned char *const *>(peer_buffers_dev), world, rank, (unsigned)epoch, ref_cdf, lut, reinterpret_cast<unsigned long long *>(counts_out),
                       peer_timeout_ns(), peer_status_device_ptr()));
    SX_LAUNCHED("build_lut_peers_kernel");
    return SX_OK;
}

// workspace: counts u64[768] | ref_cdf f32[768] | lut f32[768]
int64_t sx_hm_workspace_bytes(void) { return 768 * 8 + 768 * 4 + 768 * 4; }

int sx_hm_transform(const void *images, int dtype, int layout, int64_t n, int64_t h, int64_t w, const float *ref_hist, void *out, void *workspace, int64_t workspace_bytes, sx_stream_t stream) {
    SX_REQUIRE(workspace && workspace_bytes >= sx_hm_workspace_bytes(), "workspace too small (%lld < %lld)", (long long)workspace_bytes, (long long)sx_hm_workspace_bytes());
    SX_REQUIRE(ref_hist != nullptr, "ref_hist is NULL");
    auto *counts = static_cast<uint64_t *>(workspace);
    auto *lut = reinterpret_cast<float *>(counts + 768) + 768;
    // One chain of programmatic dependent launches: only the first kernel waits for the stream in the
    // ordinary way; each later one is resident (prologue done, first loads in flight where they do not
    // depend on the chain) by the time the kernel in front of it drains.
    zero_counts_kernel<<<1, 768, 0, static_cast<cudaStream_t>(stream)>>>(reinterpret_cast<unsigned long long *>(counts));
    SX_LAUNCHED("zero_counts_kernel");
    if (int rc = hist_impl(images, dtype, layout, n, h, w, counts, stream, true)) return rc;
    // H2 with the reference CDF rebuilt from ref_hist inside the same kernel
    SX_CUDA(launch_pdl(build_lut_kernel<true>, dim3(3), dim3(256), 0, static_cast<cudaStream_t>(stream), reinterpret_cast<const unsigned long long *>(counts), (long long)(n * h * w), ref_hist, lut));
    SX_LAUNCHED("build_lut_kernel<fused>");
    return apply_impl(images, dtype, layout, n, h, w, lut, out, stream, true);
}

// The sharded transform as ONE chain of dependent launches: zero this rank's counts of the epoch in its
// peer-mapped buffer, histogram into them, LUT with the all-reduce over NVLink peer memory, remap.
int sx_hm_transform_peers(const void *images, int dtype, int layout, int64_t n, int64_t h, int64_t w, const void *peer_buffers_dev, void *own_buffer, int world, int rank, uint32_t epoch,
                          const float *ref_cdf, void *out, void *workspace, int64_t workspace_bytes, sx_stream_t stream) {
    SX_REQUIRE(workspace && workspace_bytes >= 768 * 4, "workspace too small (%lld < %d)", (long long)workspace_bytes, 768 * 4);
    SX_REQUIRE(peer_buffers_dev && own_buffer && ref_cdf, "NULL argument");
    SX_REQUIRE(world >= 1 && world <= kPeerMaxWorld && rank >= 0 && rank < world, "bad rank/world (%d, %d)", rank, world);
    SX_REQUIRE(epoch != 0, "epoch must start at 1 (flags are zero-initialised)");
    auto *counts = reinterpret_cast<uint64_t *>(static_cast<unsigned char *>(own_buffer) + (size_t)(epoch & 1u) * 768 * 8);
    auto *lut = static_cast<float *>(workspace);
    zero_counts_kernel<<<1, 768, 0, static_cast<cudaStream_t>(stream)>>>(reinterpret_cast<unsigned long long *>(counts));
    SX_LAUNCHED("zero_counts_kernel");
    if (int rc = hist_impl(images, dtype, layout, n, h, w, counts, stream, true)) return rc;
    if (int rc = sx_hm_build_lut_peers(peer_buffers_dev, world, rank, epoch, ref_cdf, lut, nullptr, stream)) return rc;
    return apply_impl(images, dtype, layout, n, h, w, lut, out, stream, true);
}

int sx_hm_fit(const void *images, int dtype, int layout, int64_t n, int64_t h, int64_t w, float *ref_hist, void *workspace, int64_t workspace_bytes, sx_stream_t stream) {
    SX_REQUIRE(workspace && workspace_bytes >= sx_hm_workspace_bytes(), "workspace too small (%lld < %lld)", (long long)workspace_bytes, (long long)sx_hm_workspace_bytes());
    auto *counts = static_cast<uint64_t *>(workspace);
    zero_counts_kernel<<<1, 768, 0, static_cast<cudaStream_t>(stream)>>>(reinterpret_cast<unsigned long long *>(counts));
    SX_LAUNCHED("zero_counts_kernel");
    if (int rc = hist_impl(images, dtype, layout, n, h, w, counts, stream, true)) return rc;
    return sx_hm_ref_hist(counts, ref_hist, stream);
}

}  // extern "C"
