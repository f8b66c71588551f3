// lib.cu -- library-level entry points: error reporting, device info, launch counter.
#include <atomic>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <utility>
#include <vector>

#include "common.cuh"

namespace sx {

static thread_local char g_error[512] = "";
static std::atomic<int64_t> g_launches{0};

int fail(int code, const char *fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_error, sizeof(g_error), fmt, ap);
    va_end(ap);
    return code;
}

void note_launch(int n) { g_launches.fetch_add(n, std::memory_order_relaxed); }

// Per-device cached attributes (the only global mutable state besides the launch counter).
struct DevCache {
    int sms = 0;
    int64_t l2 = 0;
};
static DevCache g_dev[64];

static DevCache &dev() {
    int d = 0;
    cudaGetDevice(&d);
    if (d < 0 || d >= 64) d = 0;
    DevCache &c = g_dev[d];
    if (c.sms == 0) {
        int sms = 0, l2 = 0;
        cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, d);
        cudaDeviceGetAttribute(&l2, cudaDevAttrL2CacheSize, d);
        c.l2 = l2;
        c.sms = sms > 0 ? sms : 148;
    }
    return c;
}

void prefer_l1_impl(const void *kernel, int block_threads, size_t dyn_smem) {
    static std::mutex mu;
    static std::vector<std::pair<int, const void *>> seen;
    static const bool enabled = [] {
        const char *e = getenv("SX_L1_PREF");
        return !(e && e[0] == '0');
    }();
    if (!enabled) return;
    int d = 0;
    cudaGetDevice(&d);
    std::lock_guard<std::mutex> lock(mu);
    for (const auto &k : seen)
        if (k.first == d && k.second == kernel) return;
    seen.emplace_back(d, kernel);
    cudaFuncAttributes fa;
    int ctas = 0, max_smem = 0;
    if (cudaFuncGetAttributes(&fa, kernel) != cudaSuccess || cudaOccupancyMaxActiveBlocksPerMultiprocessor(&ctas, kernel, block_threads, dyn_smem) != cudaSuccess ||
        cudaDeviceGetAttribute(&max_smem, cudaDevAttrMaxSharedMemoryPerMultiprocessor, d) != cudaSuccess || max_smem <= 0) {
        cudaGetLastError();
        return;
    }
    // every resident CTA also holds 1 KB of system shared memory
    const size_t need = (size_t)(ctas < 1 ? 1 : ctas) * (fa.sharedSizeBytes + dyn_smem + 1024);
    int pct = (int)((need * 100 + (size_t)max_smem - 1) / (size_t)max_smem);
    if (pct > 100) pct = 100;
    static const int forced = [] {  // SX_L1_PCT=<0..100>: the same carve-out for every kernel (A/B measurements)
        const char *e = getenv("SX_L1_PCT");
        return e && e[0] ? atoi(e) : -1;
    }();
    if (forced >= 0 && forced <= 100) pct = forced;
    if (cudaFuncSetAttribute(kernel, cudaFuncAttributePreferredSharedMemoryCarveout, pct) != cudaSuccess) cudaGetLastError();
}

// ---- peer rendezvous status ----------------------------------------------------------------------
// One 16-byte record per device in pinned, device-mapped HOST memory (allocated on first use, never
// freed; the only allocation the library makes, and it is host memory).  A peer kernel whose wait for
// another rank exceeds the time budget writes {1, rank waited for, epoch, own rank} here and carries on
// (its results are then meaningless); the host reads the record without synchronising: every later
// *_peers call on that device fails with SX_ERR_CUDA until sx_peer_status_clear().
static unsigned *g_peer_status_host = nullptr;
static std::mutex g_peer_status_mu;

static unsigned *peer_status_host() {
    std::lock_guard<std::mutex> lock(g_peer_status_mu);
    if (g_peer_status_host == nullptr) {
        void *p = nullptr;
        if (cudaHostAlloc(&p, 64 * 16, cudaHostAllocMapped | cudaHostAllocPortable) != cudaSuccess) {
            cudaGetLastError();
            return nullptr;
        }
        memset(p, 0, 64 * 16);
        g_peer_status_host = static_cast<unsigned *>(p);
    }
    return g_peer_status_host;
}

unsigned *peer_status_device_ptr() {
    unsigned *h = peer_status_host();
    if (h == nullptr) return nullptr;
    int d = 0;
    cudaGetDevice(&d);
    if (d < 0 || d >= 64) d = 0;
    void *dp = nullptr;
    if (cudaHostGetDevicePointer(&dp, h + d * 4, 0) != cudaSuccess) {
        cudaGetLastError();
        return nullptr;
    }
    return static_cast<unsigned *>(dp);
}

int peer_status_check(const char *what) {
    unsigned *h = peer_status_host();
    if (h == nullptr) return SX_OK;
    int d = 0;
    cudaGetDevice(&d);
    if (d < 0 || d >= 64) d = 0;
    const volatile unsigned *r = h + d * 4;
    if (r[0] != 0u) return fail(SX_ERR_CUDA, "%s: an earlier peer exchange on this device timed out (rank %u waited for rank %u at epoch %u); results since then are invalid -- see sx_peer_status()", what, r[3], r[1], r[2]);
    return SX_OK;
}

unsigned long long peer_timeout_ns() {
    static const unsigned long long ns = [] {
        const char *e = getenv("SX_PEER_TIMEOUT_MS");
        const long long ms = e ? atoll(e) : 0;
        return (unsigned long long)(ms > 0 ? ms : 20000) * 1000000ull;  // default: 20 s
    }();
    return ns;
}

bool tuning_enabled() {
    static const bool on = [] {
        const char *e = getenv("SX_ENABLE_TUNING");
        return e && e[0] == '1';
    }();
    return on;
}

int sm_count() { return dev().sms; }
int64_t l2_bytes() { return dev().l2; }

}  // namespace sx

extern "C" {

int sx_abi_version(void) { return SX_ABI_VERSION; }

const char *sx_last_error(void) { return sx::g_error; }

int64_t sx_kernel_launches(void) { return sx::g_launches.load(std::memory_order_relaxed); }

int sx_device_info(int *sm_count, int *cc_major, int *cc_minor, int64_t *l2_bytes) {
    int d = 0;
    SX_CUDA(cudaGetDevice(&d));
    int sms = 0, maj = 0, min = 0, l2 = 0;
    SX_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, d));
    SX_CUDA(cudaDeviceGetAttribute(&maj, cudaDevAttrComputeCapabilityMajor, d));
    SX_CUDA(cudaDeviceGetAttribute(&min, cudaDevAttrComputeCapabilityMinor, d));
    SX_CUDA(cudaDeviceGetAttribute(&l2, cudaDevAttrL2CacheSize, d));
    if (sm_count) *sm_count = sms;
    if (cc_major) *cc_major = maj;
    if (cc_minor) *cc_minor = min;
    if (l2_bytes) *l2_bytes = l2;
    return SX_OK;
}

int sx_peer_status(int *timed_out, int *waited_for_rank, uint32_t *epoch) {
    unsigned *h = sx::peer_status_host();
    int d = 0;
    SX_CUDA(cudaGetDevice(&d));
    if (d < 0 || d >= 64) d = 0;
    const volatile unsigned *r = h ? h + d * 4 : nullptr;
    if (timed_out) *timed_out = r ? (int)r[0] : 0;
    if (waited_for_rank) *waited_for_rank = r ? (int)r[1] : 0;
    if (epoch) *epoch = r ? r[2] : 0u;
    return SX_OK;
}

int sx_peer_status_clear(void) {
    unsigned *h = sx::peer_status_host();
    int d = 0;
    SX_CUDA(cudaGetDevice(&d));
    if (d < 0 || d >= 64) d = 0;
    if (h) memset(h + d * 4, 0, 16);
    return SX_OK;
}

}  // extern "C"
