// reinhard.cu -- Reinhard colour transfer on B200.
//
// Pass 1 (sx_reinhard_stats): RGB -> LAB fused with the whole-batch per-channel sums (R1 + R2).
// Pass 2 (sx_reinhard_apply): RGB -> LAB -> per-channel affine -> LAB -> RGB (R1 + R3 + R4).
// The LAB image is never materialised: 2 reads + 1 write of the RGB batch = 36 B/px (float32) or
// 9 B/px (uint8) instead of the reference's ~130 B/px.
//
// Reference semantics: src/stainx/backends/torch_backend.py:L16-101 (colour), L308-355 (Reinhard);
// kernels replaced: csrc/reinhard.cu:L45-139 and the ATen glue (NHWC copy, mean/std, host sync) in
// src/stainx_cuda_torch/csrc/reinhard.cu:L25-121.
#include "common.cuh"

namespace sx {
namespace reinhard {

constexpr int kThreads = 256;

// sRGB -> linear (torch_backend.py:L28-29).  pow(t, 2.4) = 2^(2.4 log2 t) on the SFU.
__device__ __forceinline__ float srgb_to_linear(float x) {
    float t = __fmaf_rn(x, 1.0f / 1.055f, 0.055f / 1.055f);
    float p = fast_ex2(2.4f * fast_lg2(t));
    return x > 0.04045f ? p : x * (1.0f / 12.92f);
}
// linear -> sRGB (L93-94), clamped to [0,1] (L96).
__device__ __forceinline__ float linear_to_srgb(float v) {
    float p = __fmaf_rn(1.055f, fast_ex2((1.0f / 2.4f) * fast_lg2(v)), -0.055f);
    float r = v > 0.0031308f ? p : 12.92f * v;
    return fminf(fmaxf(r, 0.0f), 1.0f);
}
// f(t) of XYZ -> LAB (L41-42).
__device__ __forceinline__ float lab_f(float t) {
    float c = fast_ex2((1.0f / 3.0f) * fast_lg2(t));
    return t > 0.008856f ? c : __fmaf_rn(7.787f, t, 16.0f / 116.0f);
}
// inverse (L78-80).
__device__ __forceinline__ float lab_finv(float t) {
    return t > 0.2068966f ? t * t * t : (t - 16.0f / 116.0f) * (1.0f / 7.787f);
}

// ---- interpolated transfer-curve tables ---------------------------------------------------------
// Both passes are bound by the SFU (each pow costs MUFU.LG2 + MUFU.EX2; pass 2 needs nine pows and
// three cube roots per pixel: 18 MUFU/px against a budget of ~12 at HBM speed).  The two sRGB
// transfer curves are smooth on [0, 1], so a CTA builds them once in shared memory (accurate powf)
// and evaluates them by linear interpolation on the FMA pipe: index and fraction come from one
// FFMA.RZ against 2^23 (no F2I / I2F, which run on the SFU as well), value and slope from one
// LDS.64.  Interpolation error: 1024 intervals for sRGB -> linear (|f''| <= 3.1): 4e-7; 4096
// intervals for linear -> sRGB (|g''| <= 2.4e3 at the knee 0.0031): 1.8e-5 on [0, 1] outputs.
constexpr int kFwdN = 1024, kInvN = 4096;
constexpr int kTableBytes = (kFwdN + 1 + kInvN + 1) * 8;

__device__ __forceinline__ float srgb_to_linear_exact(float x) {
    return x > 0.04045f ? powf((x + 0.055f) / 1.055f, 2.4f) : x / 12.92f;
}
__device__ __forceinline__ float linear_to_srgb_exact(float v) {
    return v > 0.0031308f ? 1.055f * powf(v, 1.0f / 2.4f) - 0.055f : 12.92f * v;
}
// table[i] = (f(i / n), f((i + 1) / n) - f(i / n)), i = 0 .. n (the last slope is 0)
template <typename F>
__device__ __forceinline__ void build_curve(float2 *table, int n, F f) {
    for (int i = threadIdx.x; i <= n; i += blockDim.x) {
        const float a = f((float)i / (float)n);
        const float b = i < n ? f((float)(i + 1) / (float)n) : a;
        table[i] = make_float2(a, b - a);
    }
}
// x in [0, 1]
template <int N>
__device__ __forceinline__ float curve(const float2 *table, float x) {
    const float y = __fmaf_rz(x, (float)N, 8388608.0f);       // 2^23 + floor(x N)
    const float frac = __fmaf_rn(x, (float)N, 8388608.0f - y);  // x N - floor(x N), exact difference
    const float2 e = table[__float_as_uint(y) & 0x1fffu];
    return __fmaf_rn(frac, e.y, e.x);
}

// linear RGB -> (L, a, b) in the reference's 0..255 scaling (L32-53); the white point is folded
// into the matrix rows.
template <bool SHIFTED = false>
__device__ __forceinline__ void linear_to_lab(float r, float g, float b, float &L, float &A, float &B) {
    constexpr float wx = 1.0f / 0.95047f, wz = 1.0f / 1.08883f;
    float x = 0.412453f * wx * r + 0.357580f * wx * g + 0.180423f * wx * b;
    float y = 0.212671f * r + 0.715160f * g + 0.072169f * b;
    float z = 0.019334f * wz * r + 0.119193f * wz * g + 0.950227f * wz * b;
    float fx = lab_f(x), fy = lab_f(y), fz = lab_f(z);
    constexpr float off = SHIFTED ? 128.0f : 0.0f;  // pass 1 accumulates lab - 128
    L = __fmaf_rn(116.0f * 2.55f, fy, -16.0f * 2.55f - off);
    A = __fmaf_rn(500.0f, fx - fy, 128.0f - off);
    B = __fmaf_rn(200.0f, fy - fz, 128.0f - off);
}

// (L, a, b) -> linear RGB (L70-90); the white point is folded into the matrix columns.
__device__ __forceinline__ void lab_to_linear(float L, float A, float B, float &lr, float &lg, float &lb) {
    float fy = __fmaf_rn(L, 1.0f / (2.55f * 116.0f), 16.0f / 116.0f);
    float fx = __fmaf_rn(A - 128.0f, 1.0f / 500.0f, fy);
    float fz = __fmaf_rn(B - 128.0f, -1.0f / 200.0f, fy);
    float x = lab_finv(fx) * 0.95047f, y = lab_finv(fy), z = lab_finv(fz) * 1.08883f;
    lr = 3.2404542f * x - 1.5371385f * y - 0.4985314f * z;
    lg = -0.9692660f * x + 1.8760108f * y + 0.0415560f * z;
    lb = 0.0556434f * x - 0.2040259f * y + 1.0572252f * z;
}
// (L, a, b) -> sRGB in [0,1] (L70-96).
__device__ __forceinline__ void lab_to_srgb(float L, float A, float B, float &r, float &g, float &b) {
    float lr, lg, lb;
    lab_to_linear(L, A, B, lr, lg, lb);
    r = linear_to_srgb(lr);
    g = linear_to_srgb(lg);
    b = linear_to_srgb(lb);
}

// uint8 input: sRGB -> linear is a 256-entry table (built once per CTA with the accurate powf).
__device__ __forceinline__ void build_linear_lut(float *lut) {
    for (int i = threadIdx.x; i < 256; i += blockDim.x) {
        float x = __fdiv_rn((float)i, 255.0f);  // torch_backend.py:L19
        lut[i] = x > 0.04045f ? powf((x + 0.055f) / 1.055f, 2.4f) : x / 12.92f;
    }
}

struct Affine {
    float a[3], b[3];
};

// Pixel groups: a thread owns kPix consecutive pixels of one image, loaded as one 128-bit vector
// per colour plane (float32: 4 px, uint8: 16 px).  Images whose planes are not 16-byte aligned
// (H*W not a multiple of kPix, or an odd base pointer) take the scalar kernels below.
template <typename T>
struct Px;
template <>
struct Px<float> {
    static constexpr int kPix = 4;
    using Vec = float4;
};
template <>
struct Px<uint8_t> {
    static constexpr int kPix = 16;
    using Vec = uint4;
};

struct Acc {
    double s[6];
    double n;
};

__device__ __forceinline__ void block_reduce_and_add(double *v, int count, double *global) {
    __shared__ double red[kThreads / 32][8];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (int i = 0; i < count; ++i) {
        double r = warp_sum(v[i]);
        if (lane == 0) red[warp][i] = r;
    }
    __syncthreads();
    if (threadIdx.x < count) {
        double r = 0.0;
        for (int k = 0; k < kThreads / 32; ++k) r += red[k][threadIdx.x];
        atomicAdd(&global[threadIdx.x], r);
    }
}

// One pixel group as linear RGB: r/g/b[k] of pixel k.  uint8 goes through the 256-entry table,
// float32 through the interpolated curve (TAB) or the SFU.  A float32 group with a value outside
// [0, 1] (the reference does not clamp its input) takes the exact formula.
template <typename T, bool VEC, bool TAB>
__device__ __forceinline__ void load_linear(const T *__restrict__ base, int64_t hw, const float *lin_lut, const float2 *fwd, float (&r)[VEC ? Px<T>::kPix : 1], float (&g)[VEC ? Px<T>::kPix : 1], float (&b)[VEC ? Px<T>::kPix : 1]) {
    constexpr int kPix = VEC ? Px<T>::kPix : 1;
    if constexpr (sizeof(T) == 1) {
        if constexpr (VEC) {
            const uint4 vr = ld_stream(reinterpret_cast<const uint4 *>(base));
            const uint4 vg = ld_stream(reinterpret_cast<const uint4 *>(base + hw));
            const uint4 vb = ld_stream(reinterpret_cast<const uint4 *>(base + 2 * hw));
            const unsigned wr[4] = {vr.x, vr.y, vr.z, vr.w}, wg[4] = {vg.x, vg.y, vg.z, vg.w}, wb[4] = {vb.x, vb.y, vb.z, vb.w};
#pragma unroll
            for (int k = 0; k < kPix; ++k) {
                r[k] = lin_lut[(wr[k >> 2] >> (8 * (k & 3))) & 0xffu];
                g[k] = lin_lut[(wg[k >> 2] >> (8 * (k & 3))) & 0xffu];
                b[k] = lin_lut[(wb[k >> 2] >> (8 * (k & 3))) & 0xffu];
            }
        } else {
            r[0] = lin_lut[base[0]]; g[0] = lin_lut[base[hw]]; b[0] = lin_lut[base[2 * hw]];
        }
    } else {
        if constexpr (VEC) {
            const float4 vr = ld_stream(reinterpret_cast<const float4 *>(base));
            const float4 vg = ld_stream(reinterpret_cast<const float4 *>(base + hw));
            const float4 vb = ld_stream(reinterpret_cast<const float4 *>(base + 2 * hw));
            r[0] = vr.x; r[1] = vr.y; r[2] = vr.z; r[3] = vr.w;
            g[0] = vg.x; g[1] = vg.y; g[2] = vg.z; g[3] = vg.w;
            b[0] = vb.x; b[1] = vb.y; b[2] = vb.z; b[3] = vb.w;
        } else {
            r[0] = base[0]; g[0] = base[hw]; b[0] = base[2 * hw];
        }
        bool in_range = TAB;
        if constexpr (TAB) {
            // as unsigned integers, floats in [0, 1] are <= 0x3f800000; negatives and NaN are larger
            unsigned top = 0u;
#pragma unroll
            for (int k = 0; k < kPix; ++k) top = max(top, max(__float_as_uint(r[k]), max(__float_as_uint(g[k]), __float_as_uint(b[k]))));
            in_range = top <= 0x3f800000u;
        }
        if (in_range) {
#pragma unroll
            for (int k = 0; k < kPix; ++k) { r[k] = curve<kFwdN>(fwd, r[k]); g[k] = curve<kFwdN>(fwd, g[k]); b[k] = curve<kFwdN>(fwd, b[k]); }
        } else {
#pragma unroll
            for (int k = 0; k < kPix; ++k) { r[k] = srgb_to_linear(r[k]); g[k] = srgb_to_linear(g[k]); b[k] = srgb_to_linear(b[k]); }
        }
    }
}

// Shared memory of both passes: [fwd curve | inv curve] (TAB only), then the uint8 table.
template <typename T, bool TAB>
__device__ __forceinline__ void setup_tables(unsigned char *smem, const float2 *&fwd, const float2 *&inv, float *lin_lut, bool need_inv) {
    float2 *f = reinterpret_cast<float2 *>(smem);
    float2 *i = f + kFwdN + 1;
    fwd = f;
    inv = i;
    if constexpr (TAB) {
        if (sizeof(T) == 4) build_curve(f, kFwdN, [](float x) { return srgb_to_linear_exact(x); });
        // out-of-gamut values (v > 1) are clamped to 1 after the curve in the reference (L96): the
        // saturated argument must map to exactly 1 (1.055f - 0.055f is 1 - 2^-24 in float32)
        if (need_inv) build_curve(i, kInvN, [](float v) { return v >= 1.0f ? 1.0f : linear_to_srgb_exact(v); });
    }
    if constexpr (sizeof(T) == 1) build_linear_lut(lin_lut);
    __syncthreads();
}

// ---- pass 1: statistics ---------------------------------------------------------------------
template <typename T, bool VEC, bool TAB>
__global__ void __launch_bounds__(kThreads) stats_kernel(const T *__restrict__ img, int64_t n_img, int64_t hw, double *__restrict__ sums) {
    constexpr int kPix = VEC ? Px<T>::kPix : 1;
    extern __shared__ __align__(16) unsigned char dyn_smem[];
    __shared__ float lin_lut[256];
    const float2 *fwd, *inv;
    setup_tables<T, TAB>(dyn_smem, fwd, inv, lin_lut, false);
    const int64_t groups_per_img = hw / kPix;  // VEC: hw % kPix == 0
    const int64_t groups = n_img * groups_per_img;
    double acc[7] = {0, 0, 0, 0, 0, 0, 0};
    for (int64_t g = (int64_t)blockIdx.x * kThreads + threadIdx.x; g < groups; g += (int64_t)gridDim.x * kThreads) {
        const int64_t n = g / groups_per_img;
        const int64_t p0 = (g - n * groups_per_img) * kPix;
        float r[kPix], gr[kPix], b[kPix];
        load_linear<T, VEC, TAB>(img + n * 3 * hw + p0, hw, lin_lut, fwd, r, gr, b);
        // float32 partial sums over this thread's <= 16 pixels, shifted by 128 to keep the
        // second moments small; folded into double accumulators once per group.
        float s[6] = {0, 0, 0, 0, 0, 0};
#pragma unroll
        for (int k = 0; k < kPix; ++k) {
            float L, A, B;
            linear_to_lab<true>(r[k], gr[k], b[k], L, A, B);  // shifted by -128
            s[0] += L; s[1] += A; s[2] += B;
            s[3] = __fmaf_rn(L, L, s[3]); s[4] = __fmaf_rn(A, A, s[4]); s[5] = __fmaf_rn(B, B, s[5]);
        }
#pragma unroll
        for (int i = 0; i < 6; ++i) acc[i] += (double)s[i];
        acc[6] += (double)kPix;
    }
    block_reduce_and_add(acc, 7, sums);
}

// mean / unbiased std from the shifted sums (torch_backend.py:L320-321).
__global__ void finalize_kernel(const double *__restrict__ sums, float *__restrict__ mean, float *__restrict__ std) {
    const int c = threadIdx.x;
    if (c < 3) {
        const double n = sums[6];
        const double m = sums[c] / n;
        const double var = (sums[3 + c] - sums[c] * m) / (n - 1.0);
        mean[c] = (float)(m + 128.0);
        std[c] = (float)sqrt(var > 0.0 ? var : 0.0);
    }
}

// Sharded batches: the SUM all-reduce of the shifted sums fused into the finalize kernel, over
// NVLink peer memory (the same protocol as hm.cu: build_lut_peers_kernel).  bufs[p] = rank p's
// buffer as mapped here: double sums[2][8] (two parities), then uint32 flags[64].  Publish the own
// sums of this epoch, wait for every rank's, add them in rank order (identical on every rank),
// finalize.
constexpr int kPeerSumsBytes = 2 * 8 * 8;
constexpr int kPeerMaxWorld = 64;

__global__ void finalize_peers_kernel(unsigned char *const *__restrict__ bufs, int world, int rank, unsigned epoch, float *__restrict__ mean, float *__restrict__ std) {
    __shared__ double tot[8];
    const int t = threadIdx.x;
    const int parity = (int)(epoch & 1u);
    if (t < world) {
        __threadfence_system();
        unsigned *flag = reinterpret_cast<unsigned *>(bufs[t] + kPeerSumsBytes) + rank;
        asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(flag), "r"(epoch) : "memory");
        const unsigned *mine = reinterpret_cast<const unsigned *>(bufs[rank] + kPeerSumsBytes) + t;
        unsigned seen;
        do {
            asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(seen) : "l"(mine) : "memory");
        } while ((int)(seen - epoch) < 0);
    }
    __syncthreads();
    if (t < 7) {
        double acc = 0.0;
        for (int p = 0; p < world; ++p) {
            const double *src = reinterpret_cast<const double *>(bufs[p]) + parity * 8 + t;
            double v;
            asm volatile("ld.relaxed.sys.global.f64 %0, [%1];" : "=d"(v) : "l"(src) : "memory");
            acc += v;
        }
        tot[t] = acc;
    }
    __syncthreads();
    if (t < 3) {
        const double n = tot[6];
        const double m = tot[t] / n;
        const double var = (tot[3 + t] - tot[t] * m) / (n - 1.0);
        mean[t] = (float)(m + 128.0);
        std[t] = (float)sqrt(var > 0.0 ? var : 0.0);
    }
}

// ---- pass 2: transform ----------------------------------------------------------------------
template <typename T, bool VEC, bool TAB>
__global__ void __launch_bounds__(kThreads) apply_kernel(const T *__restrict__ img, T *__restrict__ out, int64_t n_img, int64_t hw, const float *__restrict__ src_mean, const float *__restrict__ src_std, const float *__restrict__ ref_mean, const float *__restrict__ ref_std) {
    constexpr int kPix = VEC ? Px<T>::kPix : 1;
    extern __shared__ __align__(16) unsigned char dyn_smem[];
    __shared__ float lin_lut[256];
    const float2 *fwd, *inv;
    setup_tables<T, TAB>(dyn_smem, fwd, inv, lin_lut, true);
    // L349: ((lab - mu_s) / (sigma_s + 1e-8)) * sigma_r + mu_r  ==  a * lab + b
    Affine af;
#pragma unroll
    for (int c = 0; c < 3; ++c) {
        af.a[c] = __fdiv_rn(ref_std[c], __fadd_rn(src_std[c], 1e-8f));
        af.b[c] = __fmaf_rn(-src_mean[c], af.a[c], ref_mean[c]);
    }
    const int64_t groups_per_img = hw / kPix;
    const int64_t groups = n_img * groups_per_img;
    for (int64_t g = (int64_t)blockIdx.x * kThreads + threadIdx.x; g < groups; g += (int64_t)gridDim.x * kThreads) {
        const int64_t n = g / groups_per_img;
        const int64_t p0 = (g - n * groups_per_img) * kPix;
        T *obase = out + n * 3 * hw + p0;
        float r[kPix], gr[kPix], b[kPix];
        load_linear<T, VEC, TAB>(img + n * 3 * hw + p0, hw, lin_lut, fwd, r, gr, b);
#pragma unroll
        for (int k = 0; k < kPix; ++k) {
            float L, A, B;
            linear_to_lab(r[k], gr[k], b[k], L, A, B);
            L = __fmaf_rn(af.a[0], L, af.b[0]);
            A = __fmaf_rn(af.a[1], A, af.b[1]);
            B = __fmaf_rn(af.a[2], B, af.b[2]);
            if constexpr (TAB) {
                // the clamp of the output (L96) commutes with the monotone transfer curve
                float lr, lg, lb;
                lab_to_linear(L, A, B, lr, lg, lb);
                r[k] = curve<kInvN>(inv, __saturatef(lr));
                gr[k] = curve<kInvN>(inv, __saturatef(lg));
                b[k] = curve<kInvN>(inv, __saturatef(lb));
            } else {
                lab_to_srgb(L, A, B, r[k], gr[k], b[k]);
            }
        }
        if constexpr (sizeof(T) == 4) {
            if constexpr (VEC) {
                st_stream(reinterpret_cast<float4 *>(obase), make_float4(r[0], r[1], r[2], r[3]));
                st_stream(reinterpret_cast<float4 *>(obase + hw), make_float4(gr[0], gr[1], gr[2], gr[3]));
                st_stream(reinterpret_cast<float4 *>(obase + 2 * hw), make_float4(b[0], b[1], b[2], b[3]));
            } else {
                obase[0] = r[0]; obase[hw] = gr[0]; obase[2 * hw] = b[0];
            }
        } else {
            // uint8 out: trunc(clamp(rgb * 255, 0, 255))  (torch_backend.py:L122-131)
            if constexpr (VEC) {
                unsigned wr[4] = {0, 0, 0, 0}, wg[4] = {0, 0, 0, 0}, wb[4] = {0, 0, 0, 0};
#pragma unroll
                for (int k = 0; k < kPix; ++k) {
                    wr[k >> 2] |= quantize_u8(r[k]) << (8 * (k & 3));
                    wg[k >> 2] |= quantize_u8(gr[k]) << (8 * (k & 3));
                    wb[k >> 2] |= quantize_u8(b[k]) << (8 * (k & 3));
                }
                st_stream(reinterpret_cast<uint4 *>(obase), make_uint4(wr[0], wr[1], wr[2], wr[3]));
                st_stream(reinterpret_cast<uint4 *>(obase + hw), make_uint4(wg[0], wg[1], wg[2], wg[3]));
                st_stream(reinterpret_cast<uint4 *>(obase + 2 * hw), make_uint4(wb[0], wb[1], wb[2], wb[3]));
            } else {
                obase[0] = (uint8_t)quantize_u8(r[0]); obase[hw] = (uint8_t)quantize_u8(gr[0]); obase[2 * hw] = (uint8_t)quantize_u8(b[0]);
            }
        }
    }
}

static int g_ctas_per_sm = 5;  // 5 x 41 KB of curve tables fill the shared memory of an SM
static int g_tables = 1;  // interpolated transfer curves instead of SFU pows (large batches)

template <typename T>
static bool can_vectorize(const void *a, const void *b, int64_t hw) {
    return hw % Px<T>::kPix == 0 && aligned16(a) && (b == nullptr || aligned16(b));
}

}  // namespace reinhard
}  // namespace sx

using namespace sx;
using namespace sx::reinhard;

// The interpolated curves pay for their construction (a few microseconds per CTA) from ~2 MP on.
static bool use_tables(int64_t n, int64_t hw) { return g_tables && n * hw >= (int64_t)1 << 21; }

template <typename T, bool VEC, bool TAB>
static int launch_stats(const T *p, int64_t n, int64_t hw, double *sums, cudaStream_t stream) {
    constexpr int kPix = VEC ? Px<T>::kPix : 1;
    const size_t smem = TAB ? kTableBytes : 0;
    if (TAB) SX_CUDA(cudaFuncSetAttribute(stats_kernel<T, VEC, TAB>, cudaFuncAttributeMaxDynamicSharedMemorySize, kTableBytes));
    const unsigned grid = stream_grid((n * hw / kPix + kThreads - 1) / kThreads, g_ctas_per_sm);
    prefer_l1(stats_kernel<T, VEC, TAB>, kThreads, smem);
    stats_kernel<T, VEC, TAB><<<grid, kThreads, smem, stream>>>(p, n, hw, sums);
    return SX_OK;
}
template <typename T, bool VEC, bool TAB>
static int launch_apply(const T *p, T *o, int64_t n, int64_t hw, const float *src_mean, const float *src_std, const float *ref_mean, const float *ref_std, cudaStream_t stream) {
    constexpr int kPix = VEC ? Px<T>::kPix : 1;
    const size_t smem = TAB ? kTableBytes : 0;
    if (TAB) SX_CUDA(cudaFuncSetAttribute(apply_kernel<T, VEC, TAB>, cudaFuncAttributeMaxDynamicSharedMemorySize, kTableBytes));
    const unsigned grid = stream_grid((n * hw / kPix + kThreads - 1) / kThreads, g_ctas_per_sm);
    prefer_l1(apply_kernel<T, VEC, TAB>, kThreads, smem);
    apply_kernel<T, VEC, TAB><<<grid, kThreads, smem, stream>>>(p, o, n, hw, src_mean, src_std, ref_mean, ref_std);
    return SX_OK;
}

#define SX_REINHARD_DISPATCH(T, vec, tab, CALL)                         \
    do {                                                                \
        if (vec) { if (tab) rc = CALL(T, true, true); else rc = CALL(T, true, false); }   \
        else { if (tab) rc = CALL(T, false, true); else rc = CALL(T, false, false); }     \
    } while (0)

extern "C" {

int sx_reinhard_set_tuning(int ctas_per_sm) {
    if (ctas_per_sm > 0 && ctas_per_sm < 100) g_ctas_per_sm = ctas_per_sm;
    if (ctas_per_sm == 100) g_tables = 0;
    if (ctas_per_sm == 101) g_tables = 1;
    return SX_OK;
}

int sx_reinhard_stats(const void *images, int dtype, int64_t n, int64_t h, int64_t w, double *sums, sx_stream_t stream_) {
    if (int rc = check_images(images, dtype, n, h, w)) return rc;
    SX_REQUIRE(sums != nullptr, "sums is NULL");
    cudaStream_t stream = static_cast<cudaStream_t>(stream_);
    const int64_t hw = h * w;
    if (n == 0 || hw == 0) return SX_OK;
    const bool tab = use_tables(n, hw) && dtype == SX_F32;  // pass 1 only needs the forward curve
    int rc = SX_OK;
    if (dtype == SX_F32) {
        const float *p = static_cast<const float *>(images);
#define SX_CALL(T, V, B) launch_stats<T, V, B>(p, n, hw, sums, stream)
        SX_REINHARD_DISPATCH(float, can_vectorize<float>(p, nullptr, hw), tab, SX_CALL);
#undef SX_CALL
    } else {
        const uint8_t *p = static_cast<const uint8_t *>(images);
#define SX_CALL(T, V, B) launch_stats<T, V, B>(p, n, hw, sums, stream)
        SX_REINHARD_DISPATCH(uint8_t, can_vectorize<uint8_t>(p, nullptr, hw), tab, SX_CALL);
#undef SX_CALL
    }
    if (rc) return rc;
    SX_LAUNCHED("reinhard::stats_kernel");
    return SX_OK;
}

int sx_reinhard_finalize(const double *sums, float *mean, float *std, sx_stream_t stream) {
    SX_REQUIRE(sums && mean && std, "NULL argument");
    finalize_kernel<<<1, 32, 0, static_cast<cudaStream_t>(stream)>>>(sums, mean, std);
    SX_LAUNCHED("reinhard::finalize_kernel");
    return SX_OK;
}

int sx_reinhard_apply(const void *images, int dtype, int64_t n, int64_t h, int64_t w, const float *src_mean, const float *src_std, const float *ref_mean, const float *ref_std, void *out, sx_stream_t stream_) {
    if (int rc = check_images(images, dtype, n, h, w)) return rc;
    cudaStream_t stream = static_cast<cudaStream_t>(stream_);
    const int64_t hw = h * w;
    if (n == 0 || hw == 0) return SX_OK;
    SX_REQUIRE(src_mean && src_std && ref_mean && ref_std && out, "NULL argument");
    const bool tab = use_tables(n, hw);
    int rc = SX_OK;
    if (dtype == SX_F32) {
        const float *p = static_cast<const float *>(images);
        float *o = static_cast<float *>(out);
#define SX_CALL(T, V, B) launch_apply<T, V, B>(p, o, n, hw, src_mean, src_std, ref_mean, ref_std, stream)
        SX_REINHARD_DISPATCH(float, can_vectorize<float>(p, o, hw), tab, SX_CALL);
#undef SX_CALL
    } else {
        const uint8_t *p = static_cast<const uint8_t *>(images);
        uint8_t *o = static_cast<uint8_t *>(out);
#define SX_CALL(T, V, B) launch_apply<T, V, B>(p, o, n, hw, src_mean, src_std, ref_mean, ref_std, stream)
        SX_REINHARD_DISPATCH(uint8_t, can_vectorize<uint8_t>(p, o, hw), tab, SX_CALL);
#undef SX_CALL
    }
    if (rc) return rc;
    SX_LAUNCHED("reinhard::apply_kernel");
    return SX_OK;
}

int64_t sx_reinhard_peer_buffer_bytes(void) { return kPeerSumsBytes + kPeerMaxWorld * 4; }

int sx_reinhard_finalize_peers(const void *peer_buffers_dev, int world, int rank, uint32_t epoch, float *mean, float *std, sx_stream_t stream) {
    SX_REQUIRE(peer_buffers_dev && mean && std, "NULL argument");
    SX_REQUIRE(world >= 1 && world <= kPeerMaxWorld && rank >= 0 && rank < world, "bad rank/world (%d, %d)", rank, world);
    SX_REQUIRE(epoch != 0, "epoch must start at 1 (flags are zero-initialised)");
    finalize_peers_kernel<<<1, 64, 0, static_cast<cudaStream_t>(stream)>>>(static_cast<unsigned char *const *>(peer_buffers_dev), world, rank, epoch, mean, std);
    SX_LAUNCHED("reinhard::finalize_peers_kernel");
    return SX_OK;
}

// workspace: sums f64[8] | src_mean f32[4] | src_std f32[4]
int64_t sx_reinhard_workspace_bytes(void) { return 8 * 8 + 4 * 4 + 4 * 4; }

int sx_reinhard_transform(const void *images, int dtype, int64_t n, int64_t h, int64_t w, const float *ref_mean, const float *ref_std, void *out, void *workspace, int64_t workspace_bytes, sx_stream_t stream) {
    SX_REQUIRE(workspace && workspace_bytes >= sx_reinhard_workspace_bytes(), "workspace too small");
    auto *sums = static_cast<double *>(workspace);
    auto *mean = reinterpret_cast<float *>(sums + 8);
    auto *std = mean + 4;
    SX_CUDA(cudaMemsetAsync(sums, 0, 8 * 8, static_cast<cudaStream_t>(stream)));
    if (int rc = sx_reinhard_stats(images, dtype, n, h, w, sums, stream)) return rc;
    if (int rc = sx_reinhard_finalize(sums, mean, std, stream)) return rc;
    return sx_reinhard_apply(images, dtype, n, h, w, mean, std, ref_mean, ref_std, out, stream);
}

int sx_reinhard_fit(const void *images, int dtype, int64_t n, int64_t h, int64_t w, float *mean, float *std, void *workspace, int64_t workspace_bytes, sx_stream_t stream) {
    SX_REQUIRE(workspace && workspace_bytes >= sx_reinhard_workspace_bytes(), "workspace too small");
    auto *sums = static_cast<double *>(workspace);
    SX_CUDA(cudaMemsetAsync(sums, 0, 8 * 8, static_cast<cudaStream_t>(stream)));
    if (int rc = sx_reinhard_stats(images, dtype, n, h, w, sums, stream)) return rc;
    return sx_reinhard_finalize(sums, mean, std, stream);
}

}  // extern "C"
