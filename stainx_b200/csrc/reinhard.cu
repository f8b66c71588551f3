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
#include <type_traits>

#include "common.cuh"

namespace sx {
namespace reinhard {

constexpr int kThreads = 256;
// How many of the three colour channels evaluate the linear -> sRGB curve of pass 2 on the SFU (lg2 + ex2)
// instead of the interpolated table in shared memory.  ncu on the all-table pass 2 (float32): issue 57 %,
// SFU 28 %, top stall short_scoreboard 4.5 warps per issue -- the warps wait for the 32 KB table, whose
// lookups by 32 unrelated lanes replay ~4.5 times.  On the SFU: pass 2 342 -> 281 us (float32).  uint8
// pass 2 already spends more of the SFU per byte moved: one channel is its optimum (226 us; none 237,
// all 240).  sRGB -> linear stays on its 8 KB table in both passes: 1 / 2 / 3 channels on the SFU were
// measured at 509 / 558 / 600 us per float32 transform against 460 us (pass 1 is at 64 % SFU already).
constexpr int kInvSfuF32 = 3, kInvSfuU8 = 1;
// Pixel groups whose loads are in flight ahead of the one being processed (measured per float32
// transform: 0: 546 us, 1: 520 us, 2: 531 us, 3: 560 us -- registers against latency).
constexpr int kAhead = 1;

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

// ---- interpolated transfer-curve tables ---------------------------------------------------------
// Both passes are bound by the SFU (each pow costs MUFU.LG2 + MUFU.EX2; pass 2 needs nine pows and
// three cube roots per pixel: 18 MUFU/px against a budget of ~12 at HBM speed).  The two sRGB
// transfer curves are smooth on [0, 1], so a CTA builds them once in shared memory (accurate powf)
// and evaluates them by linear interpolation on the FMA pipe: index and fraction come from one
// FFMA.RZ against 2^23 (no F2I / I2F, which run on the SFU as well), value and slope from one
// LDS.64.  Interpolation error: 1024 intervals for sRGB -> linear (|f''| <= 3.1): 4e-7; 4096
// intervals for linear -> sRGB (|g''| <= 2.4e3 at the knee 0.0031): 1.8e-5 on [0, 1] outputs.
constexpr int kFwdN = 1024, kInvN = 4096;
constexpr int kFwdTableBytes = (kFwdN + 1) * 8;
constexpr int kTableBytes = (kFwdN + 1 + kInvN + 1) * 8;

__device__ __forceinline__ float srgb_to_linear_exact(float x) {
    return x > 0.04045f ? powf((x + 0.055f) / 1.055f, 2.4f) : x / 12.92f;
}
__device__ __forceinline__ float linear_to_srgb_exact(float v) {
    return v > 0.0031308f ? 1.055f * powf(v, 1.0f / 2.4f) - 0.055f : 12.92f * v;
}
// table[i] = (c_i, m_i), i = 0 .. n: the chord of f over [i / n, (i + 1) / n] as a function of x itself,
// f(x) ~ c_i + m_i x  (m_i = n (f((i+1)/n) - f(i/n)), c_i = f(i/n) - m_i i/n; composed in double; the
// last slope is 0).  A lookup is then FFMA.RZ (byte offset), LOP, LDS.64, FFMA -- no fraction to extract.
template <typename F>
__device__ __forceinline__ void build_curve(float2 *table, int n, F f) {
    for (int i = threadIdx.x; i <= n; i += blockDim.x) {
        const double a = (double)f((float)i / (float)n);
        const double b = i < n ? (double)f((float)(i + 1) / (float)n) : a;
        const double m = (b - a) * (double)n;
        table[i] = make_float2((float)(a - m * ((double)i / (double)n)), (float)m);
    }
}
// x in [0, 1]
template <int N>
__device__ __forceinline__ float curve(const float2 *table, float x) {
    // 2^20 + floor(8 x N) / 8: the mantissa holds floor(x N) from bit 3 up, i.e. the BYTE offset of entry
    // floor(x N) after one mask (no shift)
    const float y = __fmaf_rz(x, (float)N, 1048576.0f);
    const float2 e = *reinterpret_cast<const float2 *>(reinterpret_cast<const unsigned char *>(table) + (__float_as_uint(y) & 0xfff8u));
    return __fmaf_rn(e.y, x, e.x);
}

// ---- pass 2 in f-space ------------------------------------------------------------------------------
// LAB is affine in (fx, fy, fz) (L45-47: L = 295.8 fy - 40.8, a = 500 (fx - fy) + 128, b = 200 (fy - fz) + 128),
// the Reinhard map is affine per LAB channel (L349) and LAB -> (fx, fy, fz) is affine again (L70-73),
// so pass 2 never forms LAB: with a_c = sigma_r / (sigma_s + 1e-8), b_c = mu_r - a_c mu_s,
//     fy' = a0 fy + cy,   fx' = a1 fx + (a0 - a1) fy + (cy + cx),   fz' = a2 fz + (a0 - a2) fy + (cy - cz)
//     cy = (b0 - 40.8 a0) / 295.8 + 16/116,  cx = (128 a1 + b1 - 128) / 500,  cz = (128 a2 + b2 - 128) / 200
// (five FMAs per pixel instead of twelve operations, and without the cancellation of 500 (fx - fy)).
struct FMap {
    float ay, cy;       // fy' = ay fy + cy
    float ax, bx, cx;   // fx' = ax fx + bx fy + cx
    float az, bz, cz;   // fz' = az fz + bz fy + cz
};
__device__ __forceinline__ FMap make_fmap(const float *src_mean, const float *src_std, const float *ref_mean, const float *ref_std) {
    double a[3], b[3];
#pragma unroll
    for (int c = 0; c < 3; ++c) {
        // the reference's float32 coefficients (L349), then exact composition in double
        const float af = __fdiv_rn(__ldcg(ref_std + c), __fadd_rn(__ldcg(src_std + c), 1e-8f));
        a[c] = (double)af;
        b[c] = (double)__ldcg(ref_mean + c) - (double)__ldcg(src_mean + c) * a[c];
    }
    const double kL = 116.0 * 2.55, k0 = 16.0 * 2.55;
    const double cy = (b[0] - k0 * a[0]) / kL + 16.0 / 116.0;
    const double cx = (128.0 * a[1] + b[1] - 128.0) / 500.0;
    const double cz = (128.0 * a[2] + b[2] - 128.0) / 200.0;
    FMap m;
    m.ay = (float)a[0]; m.cy = (float)cy;
    m.ax = (float)a[1]; m.bx = (float)(a[0] - a[1]); m.cx = (float)(cy + cx);
    m.az = (float)a[2]; m.bz = (float)(a[0] - a[2]); m.cz = (float)(cy - cz);
    return m;
}

__device__ __forceinline__ float cbrt_sfu(float t) { return fast_ex2((1.0f / 3.0f) * fast_lg2(t)); }

// linear RGB -> white-normalised XYZ, in place (r -> x, g -> y, b -> z).
__device__ __forceinline__ void linear_to_xyz(float &r, float &g, float &b) {
    constexpr float wx = 1.0f / 0.95047f, wz = 1.0f / 1.08883f;
    const float x = 0.412453f * wx * r + 0.357580f * wx * g + 0.180423f * wx * b;
    const float y = 0.212671f * r + 0.715160f * g + 0.072169f * b;
    const float z = 0.019334f * wz * r + 0.119193f * wz * g + 0.950227f * wz * b;
    r = x; g = y; b = z;
}
// f(t) over a pixel chunk, in place.  Branch-free on purpose: deciding the knee of L41-42 once per
// chunk ("every argument is above it: cube roots only") was built and measured -- on noise 2 % of the
// pixels are below the knee, which is 8 % of the 4-pixel chunks and 93 % of the warps, so nearly every
// warp ran both paths (float32 statistics pass 219 -> 290 us).
template <int K>
__device__ __forceinline__ void lab_f_group(float (&x)[K], float (&y)[K], float (&z)[K]) {
#pragma unroll
    for (int k = 0; k < K; ++k) { x[k] = lab_f(x[k]); y[k] = lab_f(y[k]); z[k] = lab_f(z[k]); }
}
// (fx, fy, fz) -> linear RGB (L78-90), white point folded into the matrix columns.
__device__ __forceinline__ void xyz_to_linear(float x, float y, float z, float &lr, float &lg, float &lb) {
    constexpr float wx = 0.95047f, wz = 1.08883f;
    lr = 3.2404542f * wx * x - 1.5371385f * y - 0.4985314f * wz * z;
    lg = -0.9692660f * wx * x + 1.8760108f * y + 0.0415560f * wz * z;
    lb = 0.0556434f * wx * x - 0.2040259f * y + 1.0572252f * wz * z;
}
__device__ __forceinline__ float lab_finv_general(float t) {
    return t > 0.2068966f ? t * t * t : __fmaf_rn(t, 1.0f / 7.787f, -(16.0f / 116.0f) / 7.787f);
}

// uint8 input: sRGB -> linear is a 256-entry table (built once per CTA with the accurate powf).
__device__ __forceinline__ void build_linear_lut(float *lut) {
    for (int i = threadIdx.x; i < 256; i += blockDim.x) {
        float x = __fdiv_rn((float)i, 255.0f);  // torch_backend.py:L19
        lut[i] = x > 0.04045f ? powf((x + 0.055f) / 1.055f, 2.4f) : x / 12.92f;
    }
}

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

template <>
struct Px<__half> {
    static constexpr int kPix = 8;  // float16 / bfloat16 storage: 8 pixels per 128-bit vector, float32 arithmetic
    using Vec = uint4;
};
template <>
struct Px<__nv_bfloat16> {
    static constexpr int kPix = 8;
    using Vec = uint4;
};

struct Acc {
    double s[6];
    double n;
};

// (image, group within the image) of the grid-stride walk over all pixel groups of the batch, advanced
// incrementally: one 64-bit division per thread instead of one per group (the division was ~25 of the
// ~490 instructions a float32 group cost in pass 2, plus an I2F and a MUFU.RCP on the SFU).
struct GroupCursor {
    int64_t n, q, dn, dq, per_img;
    __device__ __forceinline__ GroupCursor(int64_t g, int64_t stride, int64_t groups_per_img) : per_img(groups_per_img) {
        n = g / per_img;
        q = g - n * per_img;
        dn = stride / per_img;
        dq = stride - dn * per_img;
    }
    __device__ __forceinline__ void next() {
        q += dq;
        n += dn;
        if (q >= per_img) { q -= per_img; ++n; }
    }
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

// The raw loads of one pixel group (one 128-bit vector per colour plane, or one scalar), converted to
// linear RGB a CHUNK of kChunk pixels at a time so that a uint8 group (16 pixels) never holds more
// than 12 pixel values in registers.  uint8 goes through the 256-entry table, float32 through the
// interpolated curve (TAB) or the SFU.  A float32 group with a value outside [0, 1] (the reference
// does not clamp its input) takes the exact formula.
template <typename T, bool VEC, bool TAB>
struct RawPx {
    static constexpr int kPix = VEC ? Px<T>::kPix : 1;
    static constexpr int kChunk = VEC ? 4 : 1;
    static constexpr int kChunks = kPix / kChunk;
    using Vec = typename std::conditional<VEC, typename Px<T>::Vec, T>::type;
    Vec v[3];
    __device__ __forceinline__ void load(const T *__restrict__ base, int64_t hw) {
#pragma unroll
        for (int c = 0; c < 3; ++c) {
            if constexpr (VEC) v[c] = ld_stream(reinterpret_cast<const Vec *>(base + c * hw));
            else v[c] = base[c * hw];
        }
    }
    __device__ __forceinline__ void linear(int chunk, const float *lin_lut, const float2 *fwd, float (&r)[kChunk], float (&g)[kChunk], float (&b)[kChunk]) const {
        if constexpr (sizeof(T) == 1) {
            if constexpr (VEC) {
                const unsigned wr[4] = {v[0].x, v[0].y, v[0].z, v[0].w}, wg[4] = {v[1].x, v[1].y, v[1].z, v[1].w}, wb[4] = {v[2].x, v[2].y, v[2].z, v[2].w};
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                    r[k] = lin_lut[(wr[chunk] >> (8 * k)) & 0xffu];
                    g[k] = lin_lut[(wg[chunk] >> (8 * k)) & 0xffu];
                    b[k] = lin_lut[(wb[chunk] >> (8 * k)) & 0xffu];
                }
            } else {
                r[0] = lin_lut[v[0]]; g[0] = lin_lut[v[1]]; b[0] = lin_lut[v[2]];
            }
        } else {
            if constexpr (sizeof(T) == 2) {  // two words of each plane's vector per chunk of four pixels
                if constexpr (VEC) {
                    const unsigned wr[4] = {v[0].x, v[0].y, v[0].z, v[0].w}, wg[4] = {v[1].x, v[1].y, v[1].z, v[1].w}, wb[4] = {v[2].x, v[2].y, v[2].z, v[2].w};
#pragma unroll
                    for (int k = 0; k < 2; ++k) {
                        const float2 fr = Half2IO<T>::unpack(wr[2 * chunk + k]), fg = Half2IO<T>::unpack(wg[2 * chunk + k]), fb = Half2IO<T>::unpack(wb[2 * chunk + k]);
                        r[2 * k] = fr.x; r[2 * k + 1] = fr.y;
                        g[2 * k] = fg.x; g[2 * k + 1] = fg.y;
                        b[2 * k] = fb.x; b[2 * k + 1] = fb.y;
                    }
                } else {
                    r[0] = Half2IO<T>::widen(v[0]); g[0] = Half2IO<T>::widen(v[1]); b[0] = Half2IO<T>::widen(v[2]);
                }
            } else if constexpr (VEC) {
                r[0] = v[0].x; r[1] = v[0].y; r[2] = v[0].z; r[3] = v[0].w;
                g[0] = v[1].x; g[1] = v[1].y; g[2] = v[1].z; g[3] = v[1].w;
                b[0] = v[2].x; b[1] = v[2].y; b[2] = v[2].z; b[3] = v[2].w;
            } else {
                r[0] = v[0]; g[0] = v[1]; b[0] = v[2];
            }
            bool in_range = TAB;
            if constexpr (TAB) {
                // as unsigned integers, floats in [0, 1] are <= 0x3f800000; negatives and NaN are larger
                unsigned top = 0u;
#pragma unroll
                for (int k = 0; k < kChunk; ++k) top = max(top, max(__float_as_uint(r[k]), max(__float_as_uint(g[k]), __float_as_uint(b[k]))));
                in_range = top <= 0x3f800000u;
            }
            if (in_range) {
#pragma unroll
                for (int k = 0; k < kChunk; ++k) {
                    r[k] = curve<kFwdN>(fwd, r[k]);
                    g[k] = curve<kFwdN>(fwd, g[k]);
                    b[k] = curve<kFwdN>(fwd, b[k]);
                }
            } else {
#pragma unroll
                for (int k = 0; k < kChunk; ++k) { r[k] = srgb_to_linear(r[k]); g[k] = srgb_to_linear(g[k]); b[k] = srgb_to_linear(b[k]); }
            }
        }
    }
};

// Shared memory of both passes: [fwd curve | inv curve] (TAB only), then the uint8 table.
template <typename T, bool TAB>
__device__ __forceinline__ void setup_tables(unsigned char *smem, const float2 *&fwd, const float2 *&inv, float *lin_lut, bool need_inv) {
    float2 *f = reinterpret_cast<float2 *>(smem);
    float2 *i = f + kFwdN + 1;
    fwd = f;
    inv = i;
    if constexpr (TAB) {
        if (sizeof(T) != 1) build_curve(f, kFwdN, [](float x) { return srgb_to_linear_exact(x); });
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
    pdl_trigger();  // the finalize kernel behind this one may become resident (it waits for our completion)
    setup_tables<T, TAB>(dyn_smem, fwd, inv, lin_lut, false);
    const int64_t groups_per_img = hw / kPix;  // VEC: hw % kPix == 0
    const int64_t groups = n_img * groups_per_img;
    double acc[7] = {0, 0, 0, 0, 0, 0, 0};
    unsigned ngroups = 0u;  // groups this thread has summed (< 2^32: a thread sees groups / (grid x 256) of them)
    const int64_t g0 = (int64_t)blockIdx.x * kThreads + threadIdx.x, stride = (int64_t)gridDim.x * kThreads;
    GroupCursor cur(g0, stride, groups_per_img);
    using Raw = RawPx<T, VEC, TAB>;
    Raw raw, ahead[kAhead > 0 ? kAhead : 1];  // the loads of the next group(s) are in flight while this one is processed
#pragma unroll
    for (int a = 0; a < kAhead; ++a) {
        if (g0 + a * stride < groups) ahead[a].load(img + cur.n * 3 * hw + cur.q * kPix, hw);
        cur.next();
    }
    for (int64_t g = g0; g < groups; g += stride) {
        if (kAhead) {
            raw = ahead[0];
#pragma unroll
            for (int a = 0; a + 1 < kAhead; ++a) ahead[a] = ahead[a + 1];
            if (g + kAhead * stride < groups) ahead[kAhead > 0 ? kAhead - 1 : 0].load(img + cur.n * 3 * hw + cur.q * kPix, hw);
        } else {
            raw.load(img + cur.n * 3 * hw + cur.q * kPix, hw);
        }
        cur.next();
        // float32 partial sums over this thread's <= 16 pixels, shifted by 128 to keep the
        // second moments small; folded into double accumulators once per group.
        float s[6] = {0, 0, 0, 0, 0, 0};
#pragma unroll
        for (int ch = 0; ch < Raw::kChunks; ++ch) {
            float r[Raw::kChunk], gr[Raw::kChunk], b[Raw::kChunk];
            raw.linear(ch, lin_lut, fwd, r, gr, b);
#pragma unroll
            for (int k = 0; k < Raw::kChunk; ++k) linear_to_xyz(r[k], gr[k], b[k]);
            lab_f_group<Raw::kChunk>(r, gr, b);  // (fx, fy, fz)
#pragma unroll
            for (int k = 0; k < Raw::kChunk; ++k) {
                const float L = __fmaf_rn(116.0f * 2.55f, gr[k], -16.0f * 2.55f - 128.0f);  // L45-47, shifted by -128
                const float A = 500.0f * (r[k] - gr[k]);
                const float B = 200.0f * (gr[k] - b[k]);
                s[0] += L; s[1] += A; s[2] += B;
                s[3] = __fmaf_rn(L, L, s[3]); s[4] = __fmaf_rn(A, A, s[4]); s[5] = __fmaf_rn(B, B, s[5]);
            }
        }
#pragma unroll
        for (int i = 0; i < 6; ++i) acc[i] += (double)s[i];
        ++ngroups;
    }
    acc[6] = (double)ngroups * (double)kPix;
    block_reduce_and_add(acc, 7, sums);
}

// mean / unbiased std from the shifted sums (torch_backend.py:L320-321).
__global__ void finalize_kernel(const double *__restrict__ sums, float *__restrict__ mean, float *__restrict__ std) {
    const int c = threadIdx.x;
    pdl_trigger();
    pdl_wait();  // launched with launch_pdl() behind the statistics kernel
    if (c < 3) {
        const double n = __ldcg(sums + 6), sc = __ldcg(sums + c), sq = __ldcg(sums + 3 + c);
        const double m = sc / n;
        const double var = (sq - sc * m) / (n - 1.0);
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

__global__ void finalize_peers_kernel(unsigned char *const *__restrict__ bufs, int world, int rank, unsigned epoch, float *__restrict__ mean, float *__restrict__ std, unsigned long long budget_ns, unsigned *status) {
    __shared__ double tot[8];
    const int t = threadIdx.x;
    const int parity = (int)(epoch & 1u);
    pdl_trigger();
    pdl_wait();  // launched with launch_pdl() behind the statistics kernel
    if (t < world) {
        __threadfence_system();
        unsigned *flag = reinterpret_cast<unsigned *>(bufs[t] + kPeerSumsBytes) + rank;
        asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(flag), "r"(epoch) : "memory");
        const unsigned *mine = reinterpret_cast<const unsigned *>(bufs[rank] + kPeerSumsBytes) + t;
        wait_peer_flag(mine, epoch, budget_ns, status, t, rank);  // bounded: a dead rank must not hang the node
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

// Four values in [0, 1] -> four grey levels in one word, exactly trunc(x * 255) of torch_backend.py:L122-131
// (the product rounded to nearest as there, then truncated) without F2I, which runs on the SFU like
// the cube roots: adding 2^23 with round-towards-zero leaves floor(p) in the low mantissa byte, and
// three PRMTs gather the four bytes.  The inputs are already clamped (transfer curve / L96).
__device__ __forceinline__ unsigned pack_u8x4(const float (&x)[4]) {
    unsigned q[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) q[k] = __float_as_uint(__fadd_rz(__fmul_rn(x[k], 255.0f), 8388608.0f));
    return __byte_perm(__byte_perm(q[0], q[1], 0x0040), __byte_perm(q[2], q[3], 0x0040), 0x5410);
}

// ---- pass 2: transform ----------------------------------------------------------------------
template <typename T, bool VEC, bool TAB>
__global__ void __launch_bounds__(kThreads) apply_kernel(const T *__restrict__ img, T *__restrict__ out, int64_t n_img, int64_t hw, const float *__restrict__ src_mean, const float *__restrict__ src_std, const float *__restrict__ ref_mean, const float *__restrict__ ref_std) {
    constexpr int kPix = VEC ? Px<T>::kPix : 1;
    extern __shared__ __align__(16) unsigned char dyn_smem[];
    __shared__ float lin_lut[256];
    const float2 *fwd, *inv;
    setup_tables<T, TAB>(dyn_smem, fwd, inv, lin_lut, (sizeof(T) != 1 ? kInvSfuF32 : kInvSfuU8) < 3);
    // L349: ((lab - mu_s) / (sigma_s + 1e-8)) * sigma_r + mu_r, composed with LAB <-> (fx, fy, fz)
    const FMap fm = make_fmap(src_mean, src_std, ref_mean, ref_std);
    const int64_t groups_per_img = hw / kPix;
    const int64_t groups = n_img * groups_per_img;
    const int64_t g0 = (int64_t)blockIdx.x * kThreads + threadIdx.x, stride = (int64_t)gridDim.x * kThreads;
    GroupCursor cur(g0, stride, groups_per_img);
    using Raw = RawPx<T, VEC, TAB>;
    GroupCursor pre = cur;  // position of the next load
    Raw raw, ahead[kAhead > 0 ? kAhead : 1];  // the loads of the next group(s) are in flight while this one is processed
#pragma unroll
    for (int a = 0; a < kAhead; ++a) {
        if (g0 + a * stride < groups) ahead[a].load(img + pre.n * 3 * hw + pre.q * kPix, hw);
        pre.next();
    }
    for (int64_t g = g0; g < groups; g += stride) {
        T *obase = out + cur.n * 3 * hw + cur.q * kPix;
        if (kAhead) {
            raw = ahead[0];
#pragma unroll
            for (int a = 0; a + 1 < kAhead; ++a) ahead[a] = ahead[a + 1];
            if (g + kAhead * stride < groups) ahead[kAhead > 0 ? kAhead - 1 : 0].load(img + pre.n * 3 * hw + pre.q * kPix, hw);
            pre.next();
        } else {
            raw.load(img + cur.n * 3 * hw + cur.q * kPix, hw);
        }
        cur.next();
        unsigned wr[4] = {0, 0, 0, 0}, wg[4] = {0, 0, 0, 0}, wb[4] = {0, 0, 0, 0};  // uint8 output words
#pragma unroll
        for (int ch = 0; ch < Raw::kChunks; ++ch) {
            constexpr int K = Raw::kChunk;
            float r[K], gr[K], b[K];
            raw.linear(ch, lin_lut, fwd, r, gr, b);
#pragma unroll
            for (int k = 0; k < K; ++k) linear_to_xyz(r[k], gr[k], b[k]);
            lab_f_group<K>(r, gr, b);  // (fx, fy, fz)
#pragma unroll
            for (int k = 0; k < K; ++k) {  // the Reinhard map in f-space, then f^-1 (L78-80)
                const float fy = gr[k];
                r[k] = lab_finv_general(__fmaf_rn(fm.ax, r[k], __fmaf_rn(fm.bx, fy, fm.cx)));
                b[k] = lab_finv_general(__fmaf_rn(fm.az, b[k], __fmaf_rn(fm.bz, fy, fm.cz)));
                gr[k] = lab_finv_general(__fmaf_rn(fm.ay, fy, fm.cy));
            }
#pragma unroll
            for (int k = 0; k < K; ++k) {
                float lr, lg, lb;
                xyz_to_linear(r[k], gr[k], b[k], lr, lg, lb);
                if constexpr (TAB) {
                    // the clamp of the output (L96) commutes with the monotone transfer curve
                    constexpr int kSfu = sizeof(T) != 1 ? kInvSfuF32 : kInvSfuU8;
                    r[k] = kSfu >= 1 ? linear_to_srgb(lr) : curve<kInvN>(inv, __saturatef(lr));
                    gr[k] = kSfu >= 2 ? linear_to_srgb(lg) : curve<kInvN>(inv, __saturatef(lg));
                    b[k] = kSfu >= 3 ? linear_to_srgb(lb) : curve<kInvN>(inv, __saturatef(lb));
                } else {
                    r[k] = linear_to_srgb(lr);
                    gr[k] = linear_to_srgb(lg);
                    b[k] = linear_to_srgb(lb);
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
            } else if constexpr (sizeof(T) == 2) {
                // float16 / bfloat16 out: the float32 result rounded to the input's dtype (torch_backend.py:L131)
                if constexpr (VEC) {
                    wr[2 * ch] = Half2IO<T>::pack(r[0], r[1]); wr[2 * ch + 1] = Half2IO<T>::pack(r[2], r[3]);
                    wg[2 * ch] = Half2IO<T>::pack(gr[0], gr[1]); wg[2 * ch + 1] = Half2IO<T>::pack(gr[2], gr[3]);
                    wb[2 * ch] = Half2IO<T>::pack(b[0], b[1]); wb[2 * ch + 1] = Half2IO<T>::pack(b[2], b[3]);
                } else {
                    obase[0] = Half2IO<T>::narrow(r[0]); obase[hw] = Half2IO<T>::narrow(gr[0]); obase[2 * hw] = Half2IO<T>::narrow(b[0]);
                }
            } else {
                // uint8 out: trunc(clamp(rgb * 255, 0, 255))  (torch_backend.py:L122-131)
                if constexpr (VEC) {
                    wr[ch] = pack_u8x4(r);
                    wg[ch] = pack_u8x4(gr);
                    wb[ch] = pack_u8x4(b);
                } else {
                    obase[0] = (uint8_t)quantize_u8(r[0]); obase[hw] = (uint8_t)quantize_u8(gr[0]); obase[2 * hw] = (uint8_t)quantize_u8(b[0]);
                }
            }
        }
        if constexpr (sizeof(T) != 4 && VEC) {
            st_stream(reinterpret_cast<uint4 *>(obase), make_uint4(wr[0], wr[1], wr[2], wr[3]));
            st_stream(reinterpret_cast<uint4 *>(obase + hw), make_uint4(wg[0], wg[1], wg[2], wg[3]));
            st_stream(reinterpret_cast<uint4 *>(obase + 2 * hw), make_uint4(wb[0], wb[1], wb[2], wb[3]));
        }
    }
}

static int g_ctas_per_sm = 3;  // measured (float32 64 x 1024^2 transform, loads one group ahead): 3 CTAs per SM 522 us, 4: 530, 5: 539
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
    const size_t smem = TAB ? kFwdTableBytes : 0;  // pass 1 only uses the sRGB -> linear curve
    const unsigned grid = stream_grid((n * hw / kPix + kThreads - 1) / kThreads, g_ctas_per_sm);
    prefer_l1(stats_kernel<T, VEC, TAB>, kThreads, smem);
    stats_kernel<T, VEC, TAB><<<grid, kThreads, smem, stream>>>(p, n, hw, sums);
    return SX_OK;
}
template <typename T, bool VEC, bool TAB>
static int launch_apply(const T *p, T *o, int64_t n, int64_t hw, const float *src_mean, const float *src_std, const float *ref_mean, const float *ref_std, cudaStream_t stream) {
    constexpr int kPix = VEC ? Px<T>::kPix : 1;
    const size_t smem = TAB ? ((sizeof(T) != 1 ? kInvSfuF32 : kInvSfuU8) >= 3 ? kFwdTableBytes : kTableBytes) : 0;
    if (TAB)
        if (int rc = allow_big_smem(apply_kernel<T, VEC, TAB>, kTableBytes)) return rc;
    const unsigned grid = stream_grid((n * hw / kPix + kThreads - 1) / kThreads, g_ctas_per_sm);
    prefer_l1(apply_kernel<T, VEC, TAB>, kThreads, smem);
    // A plain launch on purpose.  As a programmatic dependent launch this kernel was measured 80 us SLOWER
    // (627 against 546 us per float32 transform): its CTAs are placed while the statistics kernel drains,
    // some SMs end up with five of them and others with three, and every CTA of this grid-stride kernel
    // carries the same share of the batch -- the slowest SM sets the time.
    apply_kernel<T, VEC, TAB><<<grid, kThreads, smem, stream>>>(p, o, n, hw, src_mean, src_std, ref_mean, ref_std);
    return SX_OK;
}

#define SX_REINHARD_DISPATCH(T, vec, tab, CALL)                         \
    do {                                                                \
        if (vec) { if (tab) rc = CALL(T, true, true); else rc = CALL(T, true, false); }   \
        else { if (tab) rc = CALL(T, false, true); else rc = CALL(T, false, false); }     \
    } while (0)

extern "C" {

// Development hook: process-global, not thread-safe, inert unless SX_ENABLE_TUNING=1.
int sx_reinhard_set_tuning(int ctas_per_sm) {
    if (!tuning_enabled()) return sx::fail(SX_ERR_UNSUPPORTED, "tuning hooks are disabled (set SX_ENABLE_TUNING=1 before loading the library)");
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
    const bool tab = use_tables(n, hw) && dtype != SX_U8;  // pass 1 only needs the forward curve
    int rc = SX_OK;
    if (dtype == SX_F32) {
        const float *p = static_cast<const float *>(images);
#define SX_CALL(T, V, B) launch_stats<T, V, B>(p, n, hw, sums, stream)
        SX_REINHARD_DISPATCH(float, can_vectorize<float>(p, nullptr, hw), tab, SX_CALL);
#undef SX_CALL
    } else if (dtype == SX_F16) {
        const __half *p = static_cast<const __half *>(images);
#define SX_CALL(T, V, B) launch_stats<T, V, B>(p, n, hw, sums, stream)
        SX_REINHARD_DISPATCH(__half, can_vectorize<__half>(p, nullptr, hw), tab, SX_CALL);
#undef SX_CALL
    } else if (dtype == SX_BF16) {
        const __nv_bfloat16 *p = static_cast<const __nv_bfloat16 *>(images);
#define SX_CALL(T, V, B) launch_stats<T, V, B>(p, n, hw, sums, stream)
        SX_REINHARD_DISPATCH(__nv_bfloat16, can_vectorize<__nv_bfloat16>(p, nullptr, hw), tab, SX_CALL);
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
    SX_CUDA(launch_pdl(finalize_kernel, dim3(1), dim3(32), 0, static_cast<cudaStream_t>(stream), sums, mean, std));
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
    } else if (dtype == SX_F16) {
        const __half *p = static_cast<const __half *>(images);
        __half *o = static_cast<__half *>(out);
#define SX_CALL(T, V, B) launch_apply<T, V, B>(p, o, n, hw, src_mean, src_std, ref_mean, ref_std, stream)
        SX_REINHARD_DISPATCH(__half, can_vectorize<__half>(p, o, hw), tab, SX_CALL);
#undef SX_CALL
    } else if (dtype == SX_BF16) {
        const __nv_bfloat16 *p = static_cast<const __nv_bfloat16 *>(images);
        __nv_bfloat16 *o = static_cast<__nv_bfloat16 *>(out);
#define SX_CALL(T, V, B) launch_apply<T, V, B>(p, o, n, hw, src_mean, src_std, ref_mean, ref_std, stream)
        SX_REINHARD_DISPATCH(__nv_bfloat16, can_vectorize<__nv_bfloat16>(p, o, hw), tab, SX_CALL);
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
    if (int rc = peer_status_check("sx_reinhard_finalize_peers")) return rc;
    SX_CUDA(launch_pdl(finalize_peers_kernel, dim3(1), dim3(64), 0, static_cast<cudaStream_t>(stream), static_cast<unsigned char *const *>(peer_buffers_dev), world, rank, (unsigned)epoch, mean, std, peer_timeout_ns(), peer_status_device_ptr()));
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
