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
    float p = exp2f(2.4f * __log2f(t));
    return x > 0.04045f ? p : x * (1.0f / 12.92f);
}
// linear -> sRGB (L93-94), clamped to [0,1] (L96).
__device__ __forceinline__ float linear_to_srgb(float v) {
    float p = __fmaf_rn(1.055f, exp2f((1.0f / 2.4f) * __log2f(v)), -0.055f);
    float r = v > 0.0031308f ? p : 12.92f * v;
    return fminf(fmaxf(r, 0.0f), 1.0f);
}
// f(t) of XYZ -> LAB (L41-42).
__device__ __forceinline__ float lab_f(float t) {
    float c = exp2f((1.0f / 3.0f) * __log2f(t));
    return t > 0.008856f ? c : __fmaf_rn(7.787f, t, 16.0f / 116.0f);
}
// inverse (L78-80).
__device__ __forceinline__ float lab_finv(float t) {
    return t > 0.2068966f ? t * t * t : (t - 16.0f / 116.0f) * (1.0f / 7.787f);
}

// linear RGB -> (L, a, b) in the reference's 0..255 scaling (L32-53); the white point is folded
// into the matrix rows.
__device__ __forceinline__ void linear_to_lab(float r, float g, float b, float &L, float &A, float &B) {
    constexpr float wx = 1.0f / 0.95047f, wz = 1.0f / 1.08883f;
    float x = 0.412453f * wx * r + 0.357580f * wx * g + 0.180423f * wx * b;
    float y = 0.212671f * r + 0.715160f * g + 0.072169f * b;
    float z = 0.019334f * wz * r + 0.119193f * wz * g + 0.950227f * wz * b;
    float fx = lab_f(x), fy = lab_f(y), fz = lab_f(z);
    L = __fmaf_rn(116.0f * 2.55f, fy, -16.0f * 2.55f);
    A = __fmaf_rn(500.0f, fx - fy, 128.0f);
    B = __fmaf_rn(200.0f, fy - fz, 128.0f);
}

// (L, a, b) -> sRGB in [0,1] (L70-96); the white point is folded into the matrix columns.
__device__ __forceinline__ void lab_to_srgb(float L, float A, float B, float &r, float &g, float &b) {
    float fy = __fmaf_rn(L, 1.0f / (2.55f * 116.0f), 16.0f / 116.0f);
    float fx = __fmaf_rn(A - 128.0f, 1.0f / 500.0f, fy);
    float fz = __fmaf_rn(B - 128.0f, -1.0f / 200.0f, fy);
    float x = lab_finv(fx) * 0.95047f, y = lab_finv(fy), z = lab_finv(fz) * 1.08883f;
    float lr = 3.2404542f * x - 1.5371385f * y - 0.4985314f * z;
    float lg = -0.9692660f * x + 1.8760108f * y + 0.0415560f * z;
    float lb = 0.0556434f * x - 0.2040259f * y + 1.0572252f * z;
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

// ---- pass 1: statistics ---------------------------------------------------------------------
template <typename T, bool VEC>
__global__ void __launch_bounds__(kThreads) stats_kernel(const T *__restrict__ img, int64_t n_img, int64_t hw, double *__restrict__ sums) {
    constexpr int kPix = VEC ? Px<T>::kPix : 1;
    __shared__ float lin_lut[256];
    if constexpr (sizeof(T) == 1) {
        build_linear_lut(lin_lut);
        __syncthreads();
    }
    const int64_t groups_per_img = hw / kPix;  // VEC: hw % kPix == 0
    const int64_t groups = n_img * groups_per_img;
    double acc[7] = {0, 0, 0, 0, 0, 0, 0};
    for (int64_t g = (int64_t)blockIdx.x * kThreads + threadIdx.x; g < groups; g += (int64_t)gridDim.x * kThreads) {
        const int64_t n = g / groups_per_img;
        const int64_t p0 = (g - n * groups_per_img) * kPix;
        const T *base = img + n * 3 * hw + p0;
        float r[kPix], gr[kPix], b[kPix];
        if constexpr (VEC && sizeof(T) == 4) {
            float4 vr = ld_stream(reinterpret_cast<const float4 *>(base));
            float4 vg = ld_stream(reinterpret_cast<const float4 *>(base + hw));
            float4 vb = ld_stream(reinterpret_cast<const float4 *>(base + 2 * hw));
            r[0] = vr.x; r[1] = vr.y; r[2] = vr.z; r[3] = vr.w;
            gr[0] = vg.x; gr[1] = vg.y; gr[2] = vg.z; gr[3] = vg.w;
            b[0] = vb.x; b[1] = vb.y; b[2] = vb.z; b[3] = vb.w;
#pragma unroll
            for (int k = 0; k < kPix; ++k) { r[k] = srgb_to_linear(r[k]); gr[k] = srgb_to_linear(gr[k]); b[k] = srgb_to_linear(b[k]); }
        } else if constexpr (VEC && sizeof(T) == 1) {
            uint4 vr = ld_stream(reinterpret_cast<const uint4 *>(base));
            uint4 vg = ld_stream(reinterpret_cast<const uint4 *>(base + hw));
            uint4 vb = ld_stream(reinterpret_cast<const uint4 *>(base + 2 * hw));
            unsigned wr[4] = {vr.x, vr.y, vr.z, vr.w}, wg[4] = {vg.x, vg.y, vg.z, vg.w}, wb[4] = {vb.x, vb.y, vb.z, vb.w};
#pragma unroll
            for (int k = 0; k < kPix; ++k) {
                r[k] = lin_lut[(wr[k >> 2] >> (8 * (k & 3))) & 0xffu];
                gr[k] = lin_lut[(wg[k >> 2] >> (8 * (k & 3))) & 0xffu];
                b[k] = lin_lut[(wb[k >> 2] >> (8 * (k & 3))) & 0xffu];
            }
        } else if constexpr (sizeof(T) == 4) {
            r[0] = srgb_to_linear(base[0]); gr[0] = srgb_to_linear(base[hw]); b[0] = srgb_to_linear(base[2 * hw]);
        } else {
            r[0] = lin_lut[base[0]]; gr[0] = lin_lut[base[hw]]; b[0] = lin_lut[base[2 * hw]];
        }
        // float32 partial sums over this thread's <= 16 pixels, shifted by 128 to keep the
        // second moments small; folded into double accumulators once per group.
        float s[6] = {0, 0, 0, 0, 0, 0};
#pragma unroll
        for (int k = 0; k < kPix; ++k) {
            float L, A, B;
            linear_to_lab(r[k], gr[k], b[k], L, A, B);
            L -= 128.0f; A -= 128.0f; B -= 128.0f;
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

// ---- pass 2: transform ----------------------------------------------------------------------
template <typename T, bool VEC>
__global__ void __launch_bounds__(kThreads) apply_kernel(const T *__restrict__ img, T *__restrict__ out, int64_t n_img, int64_t hw, const float *__restrict__ src_mean, const float *__restrict__ src_std, const float *__restrict__ ref_mean, const float *__restrict__ ref_std) {
    constexpr int kPix = VEC ? Px<T>::kPix : 1;
    __shared__ float lin_lut[256];
    if constexpr (sizeof(T) == 1) {
        build_linear_lut(lin_lut);
        __syncthreads();
    }
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
        const T *base = img + n * 3 * hw + p0;
        T *obase = out + n * 3 * hw + p0;
        float r[kPix], gr[kPix], b[kPix];
        if constexpr (VEC && sizeof(T) == 4) {
            float4 vr = ld_stream(reinterpret_cast<const float4 *>(base));
            float4 vg = ld_stream(reinterpret_cast<const float4 *>(base + hw));
            float4 vb = ld_stream(reinterpret_cast<const float4 *>(base + 2 * hw));
            r[0] = vr.x; r[1] = vr.y; r[2] = vr.z; r[3] = vr.w;
            gr[0] = vg.x; gr[1] = vg.y; gr[2] = vg.z; gr[3] = vg.w;
            b[0] = vb.x; b[1] = vb.y; b[2] = vb.z; b[3] = vb.w;
#pragma unroll
            for (int k = 0; k < kPix; ++k) { r[k] = srgb_to_linear(r[k]); gr[k] = srgb_to_linear(gr[k]); b[k] = srgb_to_linear(b[k]); }
        } else if constexpr (VEC && sizeof(T) == 1) {
            uint4 vr = ld_stream(reinterpret_cast<const uint4 *>(base));
            uint4 vg = ld_stream(reinterpret_cast<const uint4 *>(base + hw));
            uint4 vb = ld_stream(reinterpret_cast<const uint4 *>(base + 2 * hw));
            unsigned wr[4] = {vr.x, vr.y, vr.z, vr.w}, wg[4] = {vg.x, vg.y, vg.z, vg.w}, wb[4] = {vb.x, vb.y, vb.z, vb.w};
#pragma unroll
            for (int k = 0; k < kPix; ++k) {
                r[k] = lin_lut[(wr[k >> 2] >> (8 * (k & 3))) & 0xffu];
                gr[k] = lin_lut[(wg[k >> 2] >> (8 * (k & 3))) & 0xffu];
                b[k] = lin_lut[(wb[k >> 2] >> (8 * (k & 3))) & 0xffu];
            }
        } else if constexpr (sizeof(T) == 4) {
            r[0] = srgb_to_linear(base[0]); gr[0] = srgb_to_linear(base[hw]); b[0] = srgb_to_linear(base[2 * hw]);
        } else {
            r[0] = lin_lut[base[0]]; gr[0] = lin_lut[base[hw]]; b[0] = lin_lut[base[2 * hw]];
        }
#pragma unroll
        for (int k = 0; k < kPix; ++k) {
            float L, A, B;
            linear_to_lab(r[k], gr[k], b[k], L, A, B);
            L = __fmaf_rn(af.a[0], L, af.b[0]);
            A = __fmaf_rn(af.a[1], A, af.b[1]);
            B = __fmaf_rn(af.a[2], B, af.b[2]);
            lab_to_srgb(L, A, B, r[k], gr[k], b[k]);
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

static int g_ctas_per_sm = 4;

template <typename T>
static bool can_vectorize(const void *a, const void *b, int64_t hw) {
    return hw % Px<T>::kPix == 0 && aligned16(a) && (b == nullptr || aligned16(b));
}

}  // namespace reinhard
}  // namespace sx

using namespace sx;
using namespace sx::reinhard;

extern "C" {

int sx_reinhard_set_tuning(int ctas_per_sm) {
    if (ctas_per_sm > 0) g_ctas_per_sm = ctas_per_sm;
    return SX_OK;
}

int sx_reinhard_stats(const void *images, int dtype, int64_t n, int64_t h, int64_t w, double *sums, sx_stream_t stream_) {
    if (int rc = check_images(images, dtype, n, h, w)) return rc;
    SX_REQUIRE(sums != nullptr, "sums is NULL");
    cudaStream_t stream = static_cast<cudaStream_t>(stream_);
    const int64_t hw = h * w;
    if (n == 0 || hw == 0) return SX_OK;
    if (dtype == SX_F32) {
        const float *p = static_cast<const float *>(images);
        if (can_vectorize<float>(p, nullptr, hw)) {
            unsigned grid = stream_grid((n * hw / 4 + kThreads - 1) / kThreads, g_ctas_per_sm);
            stats_kernel<float, true><<<grid, kThreads, 0, stream>>>(p, n, hw, sums);
        } else {
            unsigned grid = stream_grid((n * hw + kThreads - 1) / kThreads, g_ctas_per_sm);
            stats_kernel<float, false><<<grid, kThreads, 0, stream>>>(p, n, hw, sums);
        }
    } else {
        const uint8_t *p = static_cast<const uint8_t *>(images);
        if (can_vectorize<uint8_t>(p, nullptr, hw)) {
            unsigned grid = stream_grid((n * hw / 16 + kThreads - 1) / kThreads, g_ctas_per_sm);
            stats_kernel<uint8_t, true><<<grid, kThreads, 0, stream>>>(p, n, hw, sums);
        } else {
            unsigned grid = stream_grid((n * hw + kThreads - 1) / kThreads, g_ctas_per_sm);
            stats_kernel<uint8_t, false><<<grid, kThreads, 0, stream>>>(p, n, hw, sums);
        }
    }
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
    if (dtype == SX_F32) {
        const float *p = static_cast<const float *>(images);
        float *o = static_cast<float *>(out);
        if (can_vectorize<float>(p, o, hw)) {
            unsigned grid = stream_grid((n * hw / 4 + kThreads - 1) / kThreads, g_ctas_per_sm);
            apply_kernel<float, true><<<grid, kThreads, 0, stream>>>(p, o, n, hw, src_mean, src_std, ref_mean, ref_std);
        } else {
            unsigned grid = stream_grid((n * hw + kThreads - 1) / kThreads, g_ctas_per_sm);
            apply_kernel<float, false><<<grid, kThreads, 0, stream>>>(p, o, n, hw, src_mean, src_std, ref_mean, ref_std);
        }
    } else {
        const uint8_t *p = static_cast<const uint8_t *>(images);
        uint8_t *o = static_cast<uint8_t *>(out);
        if (can_vectorize<uint8_t>(p, o, hw)) {
            unsigned grid = stream_grid((n * hw / 16 + kThreads - 1) / kThreads, g_ctas_per_sm);
            apply_kernel<uint8_t, true><<<grid, kThreads, 0, stream>>>(p, o, n, hw, src_mean, src_std, ref_mean, ref_std);
        } else {
            unsigned grid = stream_grid((n * hw + kThreads - 1) / kThreads, g_ctas_per_sm);
            apply_kernel<uint8_t, false><<<grid, kThreads, 0, stream>>>(p, o, n, hw, src_mean, src_std, ref_mean, ref_std);
        }
    }
    SX_LAUNCHED("reinhard::apply_kernel");
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
