// common.cuh -- shared host/device helpers for libstainx_b200 (sm_100a only).
#pragma once
#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include <atomic>

#include "../../include/stainx_b200.h"

namespace sx {

// ---- host side ------------------------------------------------------------------------------
int fail(int code, const char *fmt, ...);
void note_launch(int n = 1);
int sm_count();
int64_t l2_bytes();
// Peer rendezvous (kernels that wait for other ranks over NVLink): device pointer of this device's
// status record (NULL when it could not be mapped), the time budget of one wait, and the host-side
// check every *_peers entry point makes first (lib.cu).
unsigned *peer_status_device_ptr();
unsigned long long peer_timeout_ns();
int peer_status_check(const char *what);
// Development hooks (sx_*_set_tuning) only act when SX_ENABLE_TUNING=1 was set before the first call.
bool tuning_enabled();

#define SX_CUDA(expr)                                                                              \
    do {                                                                                           \
        cudaError_t _e = (expr);                                                                   \
        if (_e != cudaSuccess) return ::sx::fail(SX_ERR_CUDA, "%s: %s", #expr, cudaGetErrorString(_e)); \
    } while (0)

// After a <<<>>> launch: count it and surface launch errors (reference: cudaGetLastError after
// every launch, src/stainx_cuda_torch/csrc/reinhard.cu:L76-77).
#define SX_LAUNCHED(name)                                                                          \
    do {                                                                                           \
        ::sx::note_launch();                                                                       \
        cudaError_t _e = cudaGetLastError();                                                       \
        if (_e != cudaSuccess) return ::sx::fail(SX_ERR_CUDA, "launch of %s: %s", name, cudaGetErrorString(_e)); \
    } while (0)

#define SX_REQUIRE(cond, ...)                                                                      \
    do {                                                                                           \
        if (!(cond)) return ::sx::fail(SX_ERR_INVALID, __VA_ARGS__);                               \
    } while (0)

inline int64_t max_i64(int64_t a, int64_t b) { return a > b ? a : b; }
inline int dtype_bytes(int dtype) { return dtype == SX_U8 ? 1 : (dtype == SX_F32 ? 4 : 2); }
inline bool aligned16(const void *p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }

inline int check_images(const void *images, int dtype, int64_t n, int64_t h, int64_t w) {
    SX_REQUIRE(images != nullptr || n == 0 || h == 0 || w == 0, "images is NULL");
    SX_REQUIRE(dtype == SX_U8 || dtype == SX_F32 || dtype == SX_F16 || dtype == SX_BF16, "dtype must be SX_U8, SX_F32, SX_F16 or SX_BF16, got %d", dtype);
    SX_REQUIRE(n >= 0 && h >= 0 && w >= 0, "negative image extent (%lld, %lld, %lld)", (long long)n, (long long)h, (long long)w);
    return SX_OK;
}

// Grid for a grid-stride streaming kernel: enough CTAs to fill every SM `per_sm` times, never
// more than the work items.
inline unsigned stream_grid(int64_t items, int per_sm) {
    int64_t cap = (int64_t)sm_count() * per_sm;
    int64_t g = items < cap ? items : cap;
    return (unsigned)(g < 1 ? 1 : g);
}

// ---- L1 / shared-memory split --------------------------------------------------------------------
// prefer_l1(kernel, threads, dynamic smem): once per kernel and device, ask for the SMALLEST
// shared-memory carve-out that still holds the kernel's resident CTAs, i.e. the largest L1.  A
// streaming kernel's loads in flight live in L1 lines (also with L1::no_allocate), so its bandwidth
// depends on the split: the uint8 LUT remap runs at 66 us with a 16 KB carve-out and at 76 us with
// >= 132 KB.  Without a preference the driver keeps whatever split the previous kernel left when the
// new kernel fits it -- behind the 226 KB histogram kernel that cost the remap 6 us per step.
// SX_L1_PREF=0 in the environment disables the calls (A/B measurements).
void prefer_l1_impl(const void *kernel, int block_threads, size_t dyn_smem);
template <typename K>
inline void prefer_l1(K kernel, int block_threads, size_t dyn_smem = 0) {
    prefer_l1_impl(reinterpret_cast<const void *>(kernel), block_threads, dyn_smem);
}

// ---- opt-in to more than 48 KB of dynamic shared memory ---------------------------------------------
// cudaFuncAttributeMaxDynamicSharedMemorySize is a per-DEVICE function attribute: it is set once per kernel
// instantiation and device (a process that drives two GPUs must set it on both), not on every launch (a
// driver call of ~1 us on paths that are launch-bound for small batches).
template <typename K>
inline int allow_big_smem(K kernel, int bytes) {
    static std::atomic<bool> done[64];  // one array per kernel instantiation, indexed by the current device
    int d = 0;
    SX_CUDA(cudaGetDevice(&d));
    if (d < 0 || d >= 64) d = 0;
    if (!done[d].load(std::memory_order_acquire)) {
        SX_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes));
        done[d].store(true, std::memory_order_release);
    }
    return SX_OK;
}

// ---- programmatic dependent launch -------------------------------------------------------------
// launch_pdl() starts `kernel` with cudaLaunchAttributeProgrammaticStreamSerialization: the grid may
// become resident as soon as every CTA of the kernel in front of it on the stream has executed
// pdl_trigger() (or exited), instead of after that kernel has drained and the launch has been
// processed (~2 us per edge on B200).  Contract inside this library: a kernel started this way
// executes pdl_wait() before its first access to anything the kernels in front of it in the SAME
// library call produce, and -- unless the caller of the launch knows that the first kernel of the
// chain was started with normal stream order -- before its first global access of any kind.
// pdl_wait() in a kernel that was launched normally is a no-op.
#ifdef __CUDACC__
template <typename... KArgs, typename... Args>
inline cudaError_t launch_pdl(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t stream, Args... args) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = grid;
    cfg.blockDim = block;
    cfg.dynamicSmemBytes = smem;
    cfg.stream = stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    return cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...);
}
#endif

// ---- device side ----------------------------------------------------------------------------
#ifdef __CUDACC__

__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_trigger() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }

// 128-bit streaming loads/stores.  .nc + L1::no_allocate: every byte is used once per pass, so
// keep it out of L1; L2 allocation stays default so that later passes over the same image can
// hit the 126 MB L2.
__device__ __forceinline__ uint4 ld_stream(const uint4 *p) {
    uint4 r;
    asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "l"(p));
    return r;
}
__device__ __forceinline__ float4 ld_stream(const float4 *p) {
    float4 r;
    asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w) : "l"(p));
    return r;
}
__device__ __forceinline__ void st_stream(uint4 *p, const uint4 &v) {
    asm volatile("st.global.L1::no_allocate.v4.u32 [%0], {%1,%2,%3,%4};" ::"l"(p), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
}
__device__ __forceinline__ void st_stream(float4 *p, const float4 &v) {
    asm volatile("st.global.L1::no_allocate.v4.f32 [%0], {%1,%2,%3,%4};" ::"l"(p), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
}

// ---- 16-bit float storage (SX_F16 / SX_BF16 images) ------------------------------------------------
// A 32-bit word holds two consecutive values (low half first).  Widening is exact; narrowing rounds
// to nearest even (what torch's `.to(dtype)` does, torch_backend.py:L131).
template <typename T>
struct Half2IO;
template <>
struct Half2IO<__half> {
    static __device__ __forceinline__ float2 unpack(unsigned w) { return __half22float2(*reinterpret_cast<const __half2 *>(&w)); }
    static __device__ __forceinline__ unsigned pack(float a, float b) {
        const __half2 h = __floats2half2_rn(a, b);
        return *reinterpret_cast<const unsigned *>(&h);
    }
    static __device__ __forceinline__ float widen(__half v) { return __half2float(v); }
    static __device__ __forceinline__ __half narrow(float v) { return __float2half_rn(v); }
};
template <>
struct Half2IO<__nv_bfloat16> {
    static __device__ __forceinline__ float2 unpack(unsigned w) { return make_float2(__uint_as_float(w << 16), __uint_as_float(w & 0xffff0000u)); }
    static __device__ __forceinline__ unsigned pack(float a, float b) {
        const __nv_bfloat162 h = __floats2bfloat162_rn(a, b);
        return *reinterpret_cast<const unsigned *>(&h);
    }
    static __device__ __forceinline__ float widen(__nv_bfloat16 v) { return __bfloat162float(v); }
    static __device__ __forceinline__ __nv_bfloat16 narrow(float v) { return __float2bfloat16_rn(v); }
};
// Round a float to the storage type and back (the value the reference holds after its cast-back).
template <typename T>
__device__ __forceinline__ float round_trip(float v) { return Half2IO<T>::widen(Half2IO<T>::narrow(v)); }

// ---- TMA bulk copies (global -> shared) completed on an mbarrier --------------------------------
// One thread arms the barrier with the byte count and issues cp.async.bulk; the copy engine moves
// the bytes without occupying registers, scoreboard slots or LSU issue slots of the consumers, so
// an SM can keep ~100 KB of HBM reads in flight (what its share of the bandwidth needs at 1.5-2 us
// loaded latency) from a single CTA.  Addresses and sizes must be multiples of 16 bytes.
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void st_shared_v2(uint32_t addr, unsigned a, unsigned b) {
    asm volatile("st.shared.v2.u32 [%0], {%1, %2};" ::"r"(addr), "r"(a), "r"(b) : "memory");
}
__device__ __forceinline__ void mbar_init(uint64_t *bar, unsigned count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_fence_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void mbar_arrive(uint64_t *bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t *bar, unsigned bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, unsigned parity) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "WAIT_LOOP:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra WAIT_DONE;\n"
        "bra WAIT_LOOP;\n"
        "WAIT_DONE:\n"
        "}\n" ::"r"(smem_u32(bar)),
        "r"(parity)
        : "memory");
}
__device__ __forceinline__ void tma_load_1d(void *smem_dst, const void *gmem_src, unsigned bytes, uint64_t *bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(smem_dst)), "l"(gmem_src), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}

// ---- bounded wait for a peer's flag ----------------------------------------------------------------
// Spins (load-acquire, system scope) until *flag has reached `epoch`; gives up after `budget_ns` of
// the global timer, records {1, peer, epoch, own rank} in the device's status record (host-mapped) and
// returns false: a rank that died or skipped a step must not hang every GPU of the node.
__device__ __forceinline__ unsigned long long global_timer_ns() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
    return t;
}
__device__ __forceinline__ bool wait_peer_flag(const unsigned *flag, unsigned epoch, unsigned long long budget_ns, unsigned *status, int peer, int rank) {
    const unsigned long long t0 = global_timer_ns();
    unsigned polls = 0;
    for (;;) {
        unsigned seen;
        asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(seen) : "l"(flag) : "memory");
        if ((int)(seen - epoch) >= 0) return true;
        if ((++polls & 255u) == 0u && global_timer_ns() - t0 > budget_ns) {
            if (status != nullptr) {  // plain stores to mapped host memory (several timed-out threads may race: any one record will do)
                volatile unsigned *st = status;
                st[1] = (unsigned)peer;
                st[2] = epoch;
                st[3] = (unsigned)rank;
                __threadfence_system();
                st[0] = 1u;
            }
            return false;
        }
    }
}

// Raw SFU operations.  __log2f / exp2f / __fdividef wrap the MUFU instruction in denormal handling
// (FSETP + FMUL 2^24 + FADD -24: three extra instructions per call); the pixel passes are bound by
// instruction issue, their arguments are never denormal (255 x + 1 >= 1, outputs clamped), and
// flush-to-zero is the right answer where a result underflows.
__device__ __forceinline__ float fast_lg2(float x) {
    float y;
    asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}
__device__ __forceinline__ float fast_ex2(float x) {
    float y;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}
__device__ __forceinline__ float fast_rcp(float x) {
    float y;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}

// float atomic max / min on plain float storage (works for mixed signs; NaN is ignored).
__device__ __forceinline__ void atomic_max_f32(float *addr, float v) {
    v += 0.0f;  // -0.0 -> +0.0: as an int, -0.0 is INT_MIN and would lose against every negative float
    if (v >= 0.0f) atomicMax(reinterpret_cast<int *>(addr), __float_as_int(v));
    else atomicMin(reinterpret_cast<unsigned int *>(addr), __float_as_uint(v));
}
__device__ __forceinline__ void atomic_min_f32(float *addr, float v) {
    v += 0.0f;
    if (v >= 0.0f) atomicMin(reinterpret_cast<int *>(addr), __float_as_int(v));
    else atomicMax(reinterpret_cast<unsigned int *>(addr), __float_as_uint(v));
}

// float32 -> grey level exactly as torch_backend.py:L115-120: trunc(clamp(x*255, 0, 255)).
__device__ __forceinline__ unsigned quantize_u8(float x) {
    float v = __fmul_rn(x, 255.0f);
    v = fminf(fmaxf(v, 0.0f), 255.0f);
    return (unsigned)__float2int_rz(v);
}

#endif  // __CUDACC__
}  // namespace sx
