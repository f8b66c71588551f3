// atomsbench.cu -- development microbenchmark: throughput laws of shared-memory counting on B200.
// Every thread draws pseudo-random bins (LCG, ALU only) and counts them with one of several
// schemes; no global traffic in the loop.  Reports counted values per clock per SM against the
// number of resident warps.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o gpurun_out/atomsbench tools/atomsbench.cu
#include <cuda_runtime.h>

#include <cstdio>
#include <cstdlib>

#define CK(x)                                                          \
    do {                                                               \
        cudaError_t e = (x);                                           \
        if (e != cudaSuccess) {                                        \
            printf("%s: %s\n", #x, cudaGetErrorString(e));             \
            exit(1);                                                   \
        }                                                              \
    } while (0)

enum Mode {
    WARP_POPC = 0,     // warp-private 256 x u32, atomicAdd(.., 1)  (ATOMS.POPC.INC), random conflicts
    LANE32 = 1,        // lane-private u32 [bin][lane], red.add 1 (conflict-free), 32 KB / warp
    PAIR16 = 2,        // 2 lanes share a 16-bit histogram, banks {2g, 2g+1}: 8 KB / warp
    LANE8 = 3,         // lane-private packed 8-bit, red.add: 8 KB / warp (no flush here)
    SAME_ADDR = 4,     // all lanes the same bin (POPC.INC aggregation)
    LDS_STS = 5,       // lane-private packed 8-bit, plain load/add/store: 8 KB / warp
    LANE16 = 6,        // lane-private packed 16-bit, red.add: 16 KB / warp
    QUAD16 = 7,        // 4 lanes share a 16-bit histogram, banks {4g..4g+3}: 4 KB / warp
    WARP_RED = 8,      // warp-private 256 x u32, red.add with a variable increment (no POPC.INC)
    MIX = 9,           // even warps: WARP_POPC, odd warps: LDS_STS (8 KB / warp for simplicity)
    MIX31 = 10,        // warps with id % 4 == 3: LDS_STS, others WARP_POPC
};

__device__ __forceinline__ unsigned smem_bytes_per_warp(int mode) {
    switch (mode) {
        case WARP_POPC: case SAME_ADDR: case WARP_RED: return 1024;
        case LANE32: return 32768;
        case PAIR16: case LANE8: case LDS_STS: case MIX: case MIX31: return 8192;
        case LANE16: return 16384;
        case QUAD16: return 4096;
    }
    return 0;
}
static unsigned host_bytes_per_warp(int mode) {
    switch (mode) {
        case WARP_POPC: case SAME_ADDR: case WARP_RED: return 1024;
        case LANE32: return 32768;
        case PAIR16: case LANE8: case LDS_STS: case MIX: case MIX31: return 8192;
        case LANE16: return 16384;
        case QUAD16: return 4096;
    }
    return 0;
}

template <int MODE>
__global__ void count_kernel(int iters, unsigned *sink) {
    extern __shared__ __align__(16) unsigned smem[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const unsigned words = smem_bytes_per_warp(MODE) / 4;
    unsigned *region = smem + warp * words;
    for (unsigned i = lane; i < words; i += 32) region[i] = 0u;
    __syncwarp();
    const unsigned base = (unsigned)__cvta_generic_to_shared(region);
    unsigned x = (blockIdx.x * blockDim.x + threadIdx.x) * 2654435761u + 12345u;
    for (int it = 0; it < iters; ++it) {
        x = x * 1664525u + 1013904223u;
        unsigned w = x ^ (x >> 15);  // 4 pseudo-random bytes
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const unsigned b = (w >> (8 * j)) & 0xffu;
            if (MODE == WARP_POPC) {
                atomicAdd(&region[b], 1u);
            } else if (MODE == WARP_RED) {
                asm volatile("red.shared.add.u32 [%0], %1;" ::"r"(base + b * 4), "r"(1u + (w >> 31)) : "memory");
            } else if (MODE == SAME_ADDR) {
                atomicAdd(&region[(it + j) & 0xff], 1u);
            } else if (MODE == LANE32) {
                asm volatile("red.shared.add.u32 [%0], 1;" ::"r"(base + b * 128 + lane * 4) : "memory");
            } else if (MODE == PAIR16) {
                // word k = b >> 1 of group g = lane >> 1 at row k >> 1, column 2g + (k & 1)
                const unsigned k = b >> 1;
                const unsigned addr = base + ((k >> 1) * 32 + (lane & ~1) + (k & 1)) * 4;
                asm volatile("red.shared.add.u32 [%0], %1;" ::"r"(addr), "r"(1u << ((b & 1) * 16)) : "memory");
            } else if (MODE == QUAD16) {
                // word k = b >> 1 (128 words) of group g = lane >> 2 at row k >> 2, column 4g + (k & 3)
                const unsigned k = b >> 1;
                const unsigned addr = base + ((k >> 2) * 32 + (lane & ~3) + (k & 3)) * 4;
                asm volatile("red.shared.add.u32 [%0], %1;" ::"r"(addr), "r"(1u << ((b & 1) * 16)) : "memory");
            } else if (MODE == LANE16) {
                const unsigned addr = base + ((b >> 1) * 32 + lane) * 4;
                asm volatile("red.shared.add.u32 [%0], %1;" ::"r"(addr), "r"(1u << ((b & 1) * 16)) : "memory");
            } else if (MODE == LANE8) {
                const unsigned addr = base + ((b >> 2) * 32 + lane) * 4;
                asm volatile("red.shared.add.u32 [%0], %1;" ::"r"(addr), "r"(1u << ((b & 3) * 8)) : "memory");
            } else if (MODE == MIX || MODE == MIX31) {
                const bool ldsts = MODE == MIX ? (warp & 1) : ((warp & 3) == 3);
                if (ldsts) {
                    unsigned char *p = reinterpret_cast<unsigned char *>(region) + ((b >> 2) * 32 + lane) * 4 + (b & 3);
                    *p = (unsigned char)(*p + 1);
                } else {
                    atomicAdd(&region[b], 1u);
                }
            } else if (MODE == LDS_STS) {
                unsigned char *p = reinterpret_cast<unsigned char *>(region) + ((b >> 2) * 32 + lane) * 4 + (b & 3);
                *p = (unsigned char)(*p + 1);
            }
        }
    }
    __syncwarp();
    unsigned s = 0;
    for (unsigned i = lane; i < words; i += 32) s += region[i];
    if (s == 0xdeadbeefu) *sink = s;
}

template <int MODE>
static void run(const char *name, int sms, double ghz) {
    unsigned *sink;
    CK(cudaMalloc(&sink, 4));
    cudaEvent_t e0, e1;
    CK(cudaEventCreate(&e0));
    CK(cudaEventCreate(&e1));
    const unsigned bpw = host_bytes_per_warp(MODE);
    printf("%-12s", name);
    for (int warps : {4, 8, 12, 16, 24, 32, 48, 64}) {
        // one CTA per SM with `warps` warps when it fits, else split into CTAs of <= 32 warps
        int ctas_per_sm = warps > 32 ? 2 : 1;
        int wpc = warps / ctas_per_sm;
        size_t smem = (size_t)wpc * bpw;
        if (smem * ctas_per_sm > 220 * 1024) {
            printf("      -  ");
            continue;
        }
        CK(cudaFuncSetAttribute(count_kernel<MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        const int iters = 4096;
        count_kernel<MODE><<<sms * ctas_per_sm, wpc * 32, smem>>>(64, sink);
        CK(cudaEventRecord(e0));
        count_kernel<MODE><<<sms * ctas_per_sm, wpc * 32, smem>>>(iters, sink);
        CK(cudaEventRecord(e1));
        CK(cudaEventSynchronize(e1));
        float ms;
        CK(cudaEventElapsedTime(&ms, e0, e1));
        const double values = (double)sms * warps * 32 * iters * 4;
        printf(" %7.2f ", values / (ms * 1e-3) / (ghz * 1e9) / sms);
    }
    printf("\n");
}

int main() {
    cudaDeviceProp prop;
    CK(cudaGetDeviceProperties(&prop, 0));
    const int sms = prop.multiProcessorCount;
    const double ghz = prop.clockRate / 1e6;
    printf("values counted per clock per SM (at %.3f GHz nominal) vs resident warps per SM\n", ghz);
    printf("%-12s", "warps/SM");
    for (int warps : {4, 8, 12, 16, 24, 32, 48, 64}) printf(" %7d ", warps);
    printf("\n");
    run<WARP_POPC>("warp_popc", sms, ghz);
    run<WARP_RED>("warp_red", sms, ghz);
    run<SAME_ADDR>("same_addr", sms, ghz);
    run<LANE32>("lane32", sms, ghz);
    run<LANE16>("lane16", sms, ghz);
    run<PAIR16>("pair16", sms, ghz);
    run<QUAD16>("quad16", sms, ghz);
    run<LANE8>("lane8_red", sms, ghz);
    run<LDS_STS>("lane8_ldsts", sms, ghz);
    run<MIX>("mix 1:1", sms, ghz);
    run<MIX31>("mix 3:1", sms, ghz);
    return 0;
}
