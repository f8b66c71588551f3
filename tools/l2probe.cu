// l2probe.cu -- development microbenchmarks that size the Macenko pipeline (run on the GPU box):
//   1. read bandwidth of a working set re-read many times, vs its size (L2 capacity / bandwidth)
//   2. the same with a concurrent streaming write (the output stream of the apply phase)
//   3. latency of a software barrier among co-resident CTAs (cooperative launch)
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o gpurun_out/l2probe tools/l2probe.cu
#include <cooperative_groups.h>
#include <cuda_runtime.h>

#include <cstdio>
#include <cstdlib>
#include <vector>

#define CK(x)                                                                       \
    do {                                                                            \
        cudaError_t e = (x);                                                        \
        if (e != cudaSuccess) {                                                     \
            printf("%s: %s\n", #x, cudaGetErrorString(e));                          \
            exit(1);                                                                \
        }                                                                           \
    } while (0)

__device__ __forceinline__ uint4 ldg_nc(const uint4 *p) {
    uint4 r;
    asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "l"(p));
    return r;
}
__device__ __forceinline__ void stg_na(uint4 *p, uint4 v) {
    asm volatile("st.global.L1::no_allocate.v4.u32 [%0], {%1,%2,%3,%4};" ::"l"(p), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
}

// Every CTA walks the whole working set `reps` times (grid-stride), 4 loads in flight per thread.
__global__ void __launch_bounds__(256) read_kernel(const uint4 *__restrict__ buf, size_t nvec, int reps, unsigned *sink) {
    unsigned acc = 0;
    const size_t stride = (size_t)gridDim.x * 256;
    for (int r = 0; r < reps; ++r) {
        size_t i = (size_t)blockIdx.x * 256 + threadIdx.x;
        for (; i + 3 * stride < nvec; i += 4 * stride) {
            uint4 a = ldg_nc(buf + i), b = ldg_nc(buf + i + stride), c = ldg_nc(buf + i + 2 * stride), d = ldg_nc(buf + i + 3 * stride);
            acc += a.x ^ b.y ^ c.z ^ d.w;
        }
        for (; i < nvec; i += stride) acc += ldg_nc(buf + i).x;
    }
    if (acc == 0x12345678u) *sink = acc;
}

// Re-read the working set while streaming `wvec` vectors of output per repetition.
__global__ void __launch_bounds__(256) read_write_kernel(const uint4 *__restrict__ buf, size_t nvec, uint4 *__restrict__ out, size_t out_vec, int reps, unsigned *sink) {
    unsigned acc = 0;
    const size_t stride = (size_t)gridDim.x * 256;
    size_t o = (size_t)blockIdx.x * 256 + threadIdx.x;
    for (int r = 0; r < reps; ++r) {
        for (size_t i = (size_t)blockIdx.x * 256 + threadIdx.x; i < nvec; i += stride) {
            uint4 a = ldg_nc(buf + i);
            acc += a.x;
            stg_na(out + o, a);
            o += stride;
            if (o >= out_vec) o -= out_vec;
        }
    }
    if (acc == 0x12345678u) *sink = acc;
}

// Software barrier: one arrival counter, monotonically increasing target.
__global__ void __launch_bounds__(256) barrier_kernel(unsigned *counter, int rounds, long long *cycles) {
    const unsigned n = gridDim.x;
    long long t0 = clock64();
    for (int r = 1; r <= rounds; ++r) {
        __syncthreads();
        if (threadIdx.x == 0) {
            asm volatile("red.release.gpu.global.add.u32 [%0], 1;" ::"l"(counter) : "memory");
            unsigned v;
            do {
                asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(counter) : "memory");
            } while (v < n * (unsigned)r);
        }
        __syncthreads();
    }
    if (threadIdx.x == 0 && blockIdx.x == 0) *cycles = clock64() - t0;
}

__global__ void __launch_bounds__(256) gridsync_kernel(int rounds, long long *cycles) {
    namespace cg = cooperative_groups;
    cg::grid_group grid = cg::this_grid();
    long long t0 = clock64();
    for (int r = 0; r < rounds; ++r) grid.sync();
    if (threadIdx.x == 0 && blockIdx.x == 0) *cycles = clock64() - t0;
}

int main() {
    cudaDeviceProp prop;
    CK(cudaGetDeviceProperties(&prop, 0));
    printf("%s: %d SMs, L2 %.1f MB, persistingL2CacheMaxSize %.1f MB, clock %d kHz\n", prop.name, prop.multiProcessorCount, prop.l2CacheSize / 1048576.0, prop.persistingL2CacheMaxSize / 1048576.0, prop.clockRate);
    const int sms = prop.multiProcessorCount;
    const size_t max_bytes = (size_t)1 << 30;
    uint4 *buf, *out;
    unsigned *sink;
    CK(cudaMalloc(&buf, max_bytes));
    CK(cudaMalloc(&out, max_bytes));
    CK(cudaMalloc(&sink, 256));
    CK(cudaMemset(buf, 1, max_bytes));
    CK(cudaMemset(out, 0, max_bytes));
    cudaEvent_t e0, e1;
    CK(cudaEventCreate(&e0));
    CK(cudaEventCreate(&e1));

    printf("\n-- re-read bandwidth vs working set (grid = %d x 8 CTAs) --\n", sms);
    const double sizes_mb[] = {4, 8, 16, 24, 32, 48, 64, 80, 96, 112, 128, 160, 256, 1024};
    for (double mb : sizes_mb) {
        const size_t bytes = (size_t)(mb * 1048576.0);
        const size_t nvec = bytes / 16;
        int reps = (int)(4096.0 / mb);
        if (reps < 4) reps = 4;
        read_kernel<<<sms * 8, 256>>>(buf, nvec, 2, sink);  // warm
        CK(cudaEventRecord(e0));
        read_kernel<<<sms * 8, 256>>>(buf, nvec, reps, sink);
        CK(cudaEventRecord(e1));
        CK(cudaEventSynchronize(e1));
        float ms;
        CK(cudaEventElapsedTime(&ms, e0, e1));
        printf("  %7.0f MB x %4d reps: %8.1f GB/s\n", mb, reps, (double)bytes * reps / (ms * 1e-3) / 1e9);
    }

    printf("\n-- re-read + streaming write of the same volume (read GB/s; total is 2x) --\n");
    for (double mb : {8.0, 16.0, 32.0, 48.0, 64.0, 96.0, 128.0}) {
        const size_t bytes = (size_t)(mb * 1048576.0);
        const size_t nvec = bytes / 16;
        int reps = (int)(4096.0 / mb);
        read_write_kernel<<<sms * 8, 256>>>(buf, nvec, out, max_bytes / 16, 2, sink);
        CK(cudaEventRecord(e0));
        read_write_kernel<<<sms * 8, 256>>>(buf, nvec, out, max_bytes / 16, reps, sink);
        CK(cudaEventRecord(e1));
        CK(cudaEventSynchronize(e1));
        float ms;
        CK(cudaEventElapsedTime(&ms, e0, e1));
        printf("  %7.0f MB x %4d reps: read %8.1f GB/s (+ equal write)\n", mb, reps, (double)bytes * reps / (ms * 1e-3) / 1e9);
    }

    printf("\n-- barrier latency among co-resident CTAs --\n");
    unsigned *counter;
    long long *cycles;
    CK(cudaMalloc(&counter, 256));
    CK(cudaMallocManaged(&cycles, 8));
    for (int per_sm : {1, 2, 4, 6}) {
        const int grid = sms * per_sm;
        const int rounds = 2000;
        CK(cudaMemset(counter, 0, 4));
        void *args[] = {&counter, (void *)&rounds, &cycles};
        CK(cudaEventRecord(e0));
        CK(cudaLaunchCooperativeKernel((void *)barrier_kernel, dim3(grid), dim3(256), args, 0, 0));
        CK(cudaEventRecord(e1));
        CK(cudaEventSynchronize(e1));
        float ms;
        CK(cudaEventElapsedTime(&ms, e0, e1));
        printf("  atomic barrier, %4d CTAs: %.3f us per barrier (%lld cycles)\n", grid, ms * 1e3 / rounds, *cycles / rounds);
        void *args2[] = {(void *)&rounds, &cycles};
        CK(cudaEventRecord(e0));
        CK(cudaLaunchCooperativeKernel((void *)gridsync_kernel, dim3(grid), dim3(256), args2, 0, 0));
        CK(cudaEventRecord(e1));
        CK(cudaEventSynchronize(e1));
        CK(cudaEventElapsedTime(&ms, e0, e1));
        printf("  cg grid.sync,   %4d CTAs: %.3f us per barrier (%lld cycles)\n", grid, ms * 1e3 / rounds, *cycles / rounds);
    }
    return 0;
}
