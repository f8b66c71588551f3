// macenko.cu -- Macenko stain normalisation on B200, without a sort and without materialising
// OD / phi / concentration tensors.
//
// Reference semantics: src/stainx/backends/torch_backend.py:L362-560 (torch CPU oracle);
// code replaced: csrc/macenko.cu:L145-262 (one-CTA-per-image covariance + analytic eigh) and the
// ATen pipeline of src/stainx_cuda_torch/csrc/macenko.cu:L67-266 (OD copies, three full sorts).
//
// FOUR streaming passes over the RGB planes (128-bit loads, OD recomputed in registers) plus two
// ~3 % subsample passes:
//   moments         OD, tissue mask, shifted first/second moments, per-channel OD range   (M1-M3)
//   hist(ANGLE,0)   SAMPLE pass: 12-bit histogram of the angle key over a pseudo-random ~1/64
//                   subsample -> bracket [lo, hi) that contains the wanted ranks (+-8 sigma)
//   hist(ANGLE,1)   FULL pass: count keys below the bracket, 4096-cell histogram inside it with
//                   the exact float min/max of every cell                                (M5-M6)
//   hist(CONC,0/1)  the same pair for the two concentration rows over all rows           (M8-M9)
//   apply           concentrations -> rescale -> OD' -> RGB                              (M10)
// with one-CTA-per-slot kernels in between (eigen-decomposition, bracket, rank search, HE / pinv).
//
// Order statistics: phi = atan2(y, x) is never evaluated per pixel.  The selection runs on the
// "diamond angle" p(y, x) in [-2, 2], a monotone function of atan2(y, x), scaled to [0, 2^24);
// the selected pixel's (cos, sin) is recovered from p in closed form.  Concentration keys are a
// linear map of [c_lo, c_hi] (bounds from the OD range and pinv(HE)) onto the same key range.
// The full pass resolves the bracket into 4096 cells and records each cell's exact float min and
// max, so the returned value is the exact order statistic whenever the selected cell holds one
// distinct value (the normal case: a bracket is ~1 % of the data), else it is interpolated inside
// a cell of width <= bracket/4096.  The nearest-rank index is exact: ranks below the bracket are
// counted, not estimated.
#include <cstdio>
#include <mutex>
#include <type_traits>
#include <vector>

#include "common.cuh"

namespace sx {
namespace macenko {

constexpr int kThreads = 256;
constexpr int kBins = 4096;      // coarse bins of the sample pass == cells of the full pass
constexpr float kKeyMax = 16777215.0f;
constexpr int kSampleGroups = 4096;  // pixel groups sampled per image (float32: 16 K px, uint8: 64 K px)
constexpr double kBracketZ = 6.5;    // half-width of the rank bracket in sample standard deviations

constexpr float kLog2_240 = 7.906890595608519f;

// ---- workspace layout (regions are contiguous over slots so that each can be all-reduced) ----
struct SlotState {
    float e[6];           // E (3x2) row-major: [i][0] = middle eigenvector, [i][1] = largest
    int use_all;          // < 3 rows pass the mask: use every row (L409-410)
    int group_px;         // pixels per sampled group (set by the sample pass)
    long long n_sel;      // rows entering the angle selection
    long long n_all;      // rows in the slot
    long long rank[2];    // wanted 0-based rank per query (nearest-rank index)
    float proj[8];        // two affine maps of l (4 floats each) giving the stage's ranked quantities
    float lo_v[2];        // bracket [lo_v, hi_v) per query, in the value space of the ranked quantity
    float hi_v[2];
    float inv_w[2];       // (kBins - 2) / (hi_v - lo_v): value offset -> inner cell 1 .. kBins-2
    int open_lo[2];       // bracket reaches the smallest value: values < lo_v go to catch-all cell 0
    int open_hi[2];       // bracket reaches the largest value: values >= hi_v go to cell kBins-1
    float val[2];         // selected values (angle: diamond angle p; conc: concentration)
    float pinv[6];        // (HE^T HE)^-1 HE^T, 2x3 row-major
    float c_lo[2];        // concentration key mapping: key = (C - c_lo) * c_scale
    float c_scale[2];
    int miss;             // a wanted rank fell outside its bracket in the last resolve (also in STATUS)
};

struct Layout {
    int64_t moments, odrange, hist1, hist2, vmin, vmax, fit, counters, status, state, total;
    __host__ __device__ explicit Layout(int64_t slots) {
        int64_t o = 0;
        moments = o;  o += slots * 12 * 8;
        counters = o; o += slots * 8 * 8;
        odrange = o;  o += slots * 8 * 4;
        hist1 = o;    o += slots * 2 * kBins * 4;
        hist2 = o;    o += slots * 2 * kBins * 4;
        vmin = o;     o += slots * 2 * kBins * 4;
        vmax = o;     o += slots * 2 * kBins * 4;
        fit = o;      o += slots * 8 * 4;
        status = o;   o += slots * 4 * 4;
        state = (o + 15) / 16 * 16; o = state + slots * (int64_t)sizeof(SlotState);
        total = (o + 255) / 256 * 256;
    }
};

// MOMENTS region: 64-bit FIXED-POINT sums (scale 2^22), not doubles.  Integer addition is associative, so the combined
// moments of a slot do not depend on the order in which CTAs (or ranks) add their partial sums: the same call on the same
// input gives the same bits every time.  (Double atomics made the last bits order-dependent, and the float32
// eigen-decomposition can turn a last-bit difference into ~1e-7 relative in the stain vectors: one grey level on a uint8
// pixel near the truncation edge.)  A CTA's partial sum is converted once, at 2^-22 = 2.4e-7 absolute on values of 1e3-1e7
// -- below the float32 rounding of the per-thread partial sums it is made of.  Range: |sum| < 2^63 / 2^22 = 2.2e12, i.e.
// 5e10 pixels per slot (50 000 images of 1024 x 1024 in one pooled fit).  Word 10 = pixel count of the slot, unscaled.
constexpr double kMomScale = 4194304.0;
__device__ __forceinline__ long long mom_to_fx(double v) { return __double2ll_rn(v * kMomScale); }
__device__ __forceinline__ double mom_from_fx(long long v) { return (double)v * (1.0 / kMomScale); }

struct Ws {
    long long *moments;
    unsigned long long *counters;  // [slot][8]: below[2], sample count[2], pad
    float *odrange;
    unsigned *hist1, *hist2;
    float *vmin, *vmax, *fit;
    int *status;                   // [slot][4]: [0] bit q set = rank of query q fell outside its bracket;
                                   // [1..3] CTAs of the slot's image that finished moments / resolve(ANGLE) / resolve(CONC)
    SlotState *state;
    __host__ __device__ Ws(void *base, int64_t slots) {
        Layout L(slots);
        char *b = static_cast<char *>(base);
        moments = reinterpret_cast<long long *>(b + L.moments);
        counters = reinterpret_cast<unsigned long long *>(b + L.counters);
        odrange = reinterpret_cast<float *>(b + L.odrange);
        hist1 = reinterpret_cast<unsigned *>(b + L.hist1);
        hist2 = reinterpret_cast<unsigned *>(b + L.hist2);
        vmin = reinterpret_cast<float *>(b + L.vmin);
        vmax = reinterpret_cast<float *>(b + L.vmax);
        fit = reinterpret_cast<float *>(b + L.fit);
        status = reinterpret_cast<int *>(b + L.status);
        state = reinterpret_cast<SlotState *>(b + L.state);
    }
};

// ---- pixel loading -----------------------------------------------------------------------------
// A thread owns kPix consecutive pixels of one image: one 128-bit load per colour plane
// (float32: 4 px, uint8: 16 px).  Planes that are not 16-byte aligned use kPix = 1.
//
// Every pass works on l = log2(255 x + 1) (float32: one FFMA + one MUFU.LG2; uint8: a 256-entry
// table built per CTA in the reference's operation order x = v / 255, t = x * 255 + 1,
// torch_backend.py:L112, L550).  Optical density is affine in l,
//     OD = -ln((255 x + 1) / 240) = ln2 * (log2(240) - l),
// so the tissue mask, the moments, both projections and the reconstruction fold into FMAs on l:
//     min_c OD_c >= 0.15        <=>  max_c l_c <= kLThr
//     cov(OD) = ln2^2 cov(l)     (same eigenvectors)
//     OD . v  = ln2 log2(240) sum(v) - ln2 (l . v)
template <typename T, bool VEC>
struct Pix {
    static constexpr int kPix = VEC ? (int)(16 / sizeof(T)) : 1;
};

constexpr float kLThr = 7.690486339f;    // log2(240) - 0.15 / ln2
constexpr float kLShift = 6.5f;          // moments are accumulated about this value of l

__device__ __forceinline__ float f32_l(float x) { return fast_lg2(__fmaf_rn(x, 255.0f, 1.0f)); }
// uint8: byte K of `w` -> l.  The reference computes x = v / 255, t = x * 255 + 1 in float32
// (torch_backend.py:L112, L550); t == v + 1 exactly for every v in 0..255 (checked exhaustively), so
// l = log2(v + 1).  The byte goes straight into the mantissa of 2^23 (one PRMT) and 2^23 - 1 is
// subtracted: no integer-to-float conversion (SFU) and no table in shared memory (a 256-entry table
// indexed by 32 random bytes costs ~3.5 bank-conflict replays per lookup, which bound the uint8
// passes: ncu mio_throttle 8.5 warps per issue in the first version of the reconstruction).
template <int K>
__device__ __forceinline__ float u8_l(unsigned w) {
    return fast_lg2(__fsub_rn(__uint_as_float(__byte_perm(w, 0x4b000000u, 0x7650 + K)), 8388607.0f));
}
__device__ __forceinline__ float u8_l_scalar(unsigned v) { return fast_lg2(__fsub_rn(__uint_as_float(0x4b000000u | v), 8388607.0f)); }

// Raw 128-bit (or scalar) loads of one pixel group, kept in registers so that the loads of the next
// group can be in flight while the current one is processed.
template <typename T, bool VEC>
struct RawGroup {
    using Vec = typename std::conditional<VEC, typename std::conditional<sizeof(T) == 4, float4, uint4>::type, T>::type;
    Vec v[3];
    __device__ __forceinline__ void load(const T *__restrict__ base, int64_t hw) {
        if constexpr (VEC) {
#pragma unroll
            for (int c = 0; c < 3; ++c) v[c] = ld_stream(reinterpret_cast<const Vec *>(base + c * hw));
        } else {
#pragma unroll
            for (int c = 0; c < 3; ++c) v[c] = base[c * hw];
        }
    }
    // l[c][k] = log2(255 x + 1) of channel c of pixel k.
    __device__ __forceinline__ void to_l(const float *tab, float (&l)[3][Pix<T, VEC>::kPix]) const {
        constexpr int kPix = Pix<T, VEC>::kPix;
        if constexpr (sizeof(T) == 4) {
            if constexpr (VEC) {
#pragma unroll
                for (int c = 0; c < 3; ++c) { l[c][0] = f32_l(v[c].x); l[c][1] = f32_l(v[c].y); l[c][2] = f32_l(v[c].z); l[c][3] = f32_l(v[c].w); }
            } else {
#pragma unroll
                for (int c = 0; c < 3; ++c) l[c][0] = f32_l(v[c]);
            }
        } else if constexpr (sizeof(T) == 2) {  // float16 / bfloat16 storage, float32 arithmetic
            if constexpr (VEC) {
#pragma unroll
                for (int c = 0; c < 3; ++c) {
                    const unsigned w[4] = {v[c].x, v[c].y, v[c].z, v[c].w};
#pragma unroll
                    for (int k = 0; k < 4; ++k) {
                        const float2 f = Half2IO<T>::unpack(w[k]);
                        l[c][2 * k] = f32_l(f.x);
                        l[c][2 * k + 1] = f32_l(f.y);
                    }
                }
            } else {
#pragma unroll
                for (int c = 0; c < 3; ++c) l[c][0] = f32_l(Half2IO<T>::widen(v[c]));
            }
        } else {
            if constexpr (VEC) {
#pragma unroll
                for (int c = 0; c < 3; ++c) {
                    const unsigned w[4] = {v[c].x, v[c].y, v[c].z, v[c].w};
#pragma unroll
                    for (int k = 0; k < kPix; k += 4) {
                        l[c][k] = u8_l<0>(w[k >> 2]);
                        l[c][k + 1] = u8_l<1>(w[k >> 2]);
                        l[c][k + 2] = u8_l<2>(w[k >> 2]);
                        l[c][k + 3] = u8_l<3>(w[k >> 2]);
                    }
                }
            } else {
#pragma unroll
                for (int c = 0; c < 3; ++c) l[c][0] = u8_l_scalar(v[c]);
            }
        }
    }
};

// Streams the pixel groups first, first + stride, ... (< groups) of one image through `body(l, gi)`
// with one group of loads always in flight ahead of the arithmetic.  The trip count is uniform
// over the CTA (bodies may use __syncthreads through `every`), inactive threads skip the body.
template <typename T, bool VEC, typename Body, typename Every>
__device__ __forceinline__ void stream_groups(const T *__restrict__ image, int64_t hw, int64_t groups, int64_t cta_first, int64_t cta_stride, const float *tab, Body body, Every every) {
    constexpr int kPix = Pix<T, VEC>::kPix;
    RawGroup<T, VEC> cur, nxt;  // (two groups ahead was measured: the extra registers cost more than the latency they hide)
    int64_t gi = cta_first + threadIdx.x;
    if (gi < groups) cur.load(image + gi * kPix, hw);
    for (int64_t base = cta_first; base < groups; base += cta_stride) {
        const int64_t gn = gi + cta_stride;
        if (gn < groups) nxt.load(image + gn * kPix, hw);
        if (gi < groups) {
            float l[3][kPix];
            cur.to_l(tab, l);
            body(l, gi);
        }
        every();
        cur = nxt;
        gi = gn;
    }
}

// The same, but the body runs on EVERY thread of the CTA in every trip (with `valid` = the thread
// has a group; its l values are unspecified otherwise), so bodies may use full-warp collectives.
template <typename T, bool VEC, typename Body, typename Every>
__device__ __forceinline__ void stream_groups_uniform(const T *__restrict__ image, int64_t hw, int64_t groups, int64_t cta_first, int64_t cta_stride, const float *tab, Body body, Every every) {
    constexpr int kPix = Pix<T, VEC>::kPix;
    RawGroup<T, VEC> cur, nxt;
#pragma unroll
    for (int c = 0; c < 3; ++c) { cur.v[c] = typename RawGroup<T, VEC>::Vec(); nxt.v[c] = cur.v[c]; }
    int64_t gi = cta_first + threadIdx.x;
    if (gi < groups) cur.load(image + gi * kPix, hw);
    for (int64_t base = cta_first; base < groups; base += cta_stride) {
        const int64_t gn = gi + cta_stride;
        if (gn < groups) nxt.load(image + gn * kPix, hw);
        float l[3][kPix];
        cur.to_l(tab, l);
        body(l, gi < groups);
        every();
        cur = nxt;
        gi = gn;
    }
}

// Loads kPix pixels; l[c][k] = log2(255 x + 1) of channel c of pixel k (no prefetch; sample pass).
template <typename T, bool VEC>
__device__ __forceinline__ void load_l(const T *__restrict__ base, int64_t hw, const float *tab, float (&l)[3][Pix<T, VEC>::kPix]) {
    RawGroup<T, VEC> r;
    r.load(base, hw);
    r.to_l(tab, l);
}

// CTA -> (image, chunk) mapping shared by all pixel passes.
struct PassGeom {
    int64_t n_img, hw;
    int cpi;  // CTAs per image
};

// ---- ranked quantities ---------------------------------------------------------------------------
// Diamond angle: monotone in atan2(y, x) over (-pi, pi], range [-2, 2].
__device__ __forceinline__ float diamond_angle(float y, float x) {
    const float a = __fadd_rn(fabsf(x), fabsf(y));
    float r = __fmul_rn(y, fast_rcp(a));
    r = a > 0.0f ? r : 0.0f;
    const float flipped = __fsub_rn(copysignf(2.0f, y), r);  // y >= 0: 2 - r, y < 0: -2 - r
    return x >= 0.0f ? r : flipped;
}
// value = aff[3] + aff[0] l0 + aff[1] l1 + aff[2] l2   (pinned: every pass must agree bit for bit)
__device__ __forceinline__ float affine3(const float *aff, float l0, float l1, float l2) {
    return __fmaf_rn(aff[2], l2, __fmaf_rn(aff[1], l1, __fmaf_rn(aff[0], l0, aff[3])));
}
// Sample-pass keys: floats in [0, 2^24), monotone in the ranked value.
__device__ __forceinline__ float angle_key(float p) {  // p in [-2,2]
    return fminf(fmaxf(__fmul_rn(__fadd_rn(p, 2.0f), 4194304.0f), 0.0f), kKeyMax);
}
__device__ __forceinline__ float conc_key(float c, float lo, float scale) {
    return fminf(fmaxf(__fmul_rn(__fsub_rn(c, lo), scale), 0.0f), kKeyMax);
}

// The two ranked values of one pixel.  ANGLE: the diamond angle of (That1, That0) for kept rows
// (both queries share it), NaN for masked rows so that every later comparison is false.
// CONC: the two concentrations.
struct RankParams {
    float proj[8];
    int use_all;
    __device__ __forceinline__ explicit RankParams(const SlotState &s) : use_all(s.use_all) {
#pragma unroll
        for (int i = 0; i < 8; ++i) proj[i] = s.proj[i];
    }
};

template <int STAGE>
__device__ __forceinline__ void ranked_values(const RankParams &st, float l0, float l1, float l2, float &v0, float &v1) {
    if constexpr (STAGE == SX_STAGE_ANGLE) {
        const float t0 = affine3(st.proj, l0, l1, l2);      // That[:,0] (L417)
        const float t1 = affine3(st.proj + 4, l0, l1, l2);  // That[:,1]
        const float p = diamond_angle(t1, t0);              // monotone in atan2(t1, t0) (L418)
        const bool keep = st.use_all || fmaxf(l0, fmaxf(l1, l2)) <= kLThr;  // L404-405
        v0 = v1 = keep ? p : __int_as_float(0x7fc00000);
    } else {
        v0 = affine3(st.proj, l0, l1, l2);  // L444
        v1 = affine3(st.proj + 4, l0, l1, l2);
    }
}

// ---- moments (M1-M3) ---------------------------------------------------------------------------
// One pixel group into float32 partial sums s[10] (count, 3 sums, 6 products of l - kLShift over the
// kept rows) and the running per-channel range of l over all rows.
template <int kPix, bool MASKED>
__device__ __forceinline__ void moments_group(const float (&l)[3][kPix], float (&s)[10], float (&lo)[3], float (&hi)[3]) {
#pragma unroll
    for (int k = 0; k < kPix; ++k) {
        const float r = l[0][k], gg = l[1][k], b = l[2][k];
        bool keep = true;
        if (MASKED) {
            // one range over the three channels (3 instructions per pixel instead of 6): the CONC key
            // range only needs an enclosing interval; lo[0] / hi[0] carry it, the caller copies it
            const float mx = fmaxf(r, fmaxf(gg, b));
            lo[0] = fminf(lo[0], fminf(r, fminf(gg, b)));
            hi[0] = fmaxf(hi[0], mx);
            keep = mx <= kLThr;  // L404-405
        }
        const float x = keep ? r - kLShift : 0.0f, y = keep ? gg - kLShift : 0.0f, z = keep ? b - kLShift : 0.0f;
        s[0] += keep ? 1.0f : 0.0f;
        s[1] += x; s[2] += y; s[3] += z;
        s[4] = __fmaf_rn(x, x, s[4]); s[5] = __fmaf_rn(x, y, s[5]); s[6] = __fmaf_rn(x, z, s[6]);
        s[7] = __fmaf_rn(y, y, s[7]); s[8] = __fmaf_rn(y, z, s[8]); s[9] = __fmaf_rn(z, z, s[9]);
    }
}

// CTA-wide sum of acc[10] (fixed-point integers: any order gives the same bits); afterwards thread i < 10 holds total i in acc[0].
__device__ __forceinline__ long long warp_sum_ll(long long v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ void block_sum10(long long (&acc)[10], long long (*red)[10]) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
    for (int i = 0; i < 10; ++i) {
        const long long r = warp_sum_ll(acc[i]);
        if (lane == 0) red[warp][i] = r;
    }
    __syncthreads();
    if (threadIdx.x < 10) {
        long long r = 0;
        for (int k = 0; k < kThreads / 32; ++k) r += red[k][threadIdx.x];
        acc[0] = r;
    }
}

// Rows (of kThreads pixel groups) whose float32 partial sums a thread folds into one fixed-point addition.
template <typename T, bool VEC>
struct MomFlush {
    static constexpr int kRows = Pix<T, VEC>::kPix >= 16 ? 2 : 8;
};

// Streams groups of one image through moments_group.  A thread sums the float32 contributions of kRows consecutive rows
// (<= 64 pixels, fixed order) and converts that partial sum to 64-bit fixed point ONCE; everything after is integer
// addition.  Callers start at a row that is a multiple of kRows within the image, so the set of pixels behind every
// float32 partial sum -- and with it every bit of the image's moments -- is the same whatever the batch around the image,
// the grid size or the number of ranks.
template <typename T, bool VEC, bool MASKED>
__device__ __forceinline__ void moments_stream(const T *__restrict__ image, int64_t hw, int64_t groups_end, int64_t cta_first, int64_t cta_stride, const float *tab, long long (&acc)[10], float (&lo)[3], float (&hi)[3]) {
    constexpr int kPix = Pix<T, VEC>::kPix;
    constexpr int kFlush = MomFlush<T, VEC>::kRows;
    float s[10] = {0, 0, 0, 0, 0, 0, 0, 0, 0, 0};
    int pending = 0;
    stream_groups<T, VEC>(
        image, hw, groups_end, cta_first, cta_stride, tab,
        [&](const float(&l)[3][kPix], int64_t) {
            moments_group<kPix, MASKED>(l, s, lo, hi);
            if (++pending == kFlush) {
#pragma unroll
                for (int i = 0; i < 10; ++i) { acc[i] += mom_to_fx((double)s[i]); s[i] = 0.0f; }
                pending = 0;
            }
        },
        [] {});
#pragma unroll
    for (int i = 0; i < 10; ++i) acc[i] += mom_to_fx((double)s[i]);
}

// ---- symmetric 3x3 eigen-decomposition (M4) ----------------------------------------------------
// Cyclic Jacobi in float32 on the covariance scaled to unit trace (the covariance itself comes
// from float64 moments).  It runs on one thread between two image phases, so its latency is on the
// critical path of the fused pipeline: no divisions or square roots beyond MUFU.RCP / MUFU.RSQ, all
// nine + nine entries in registers.  Eigenvector accuracy ~1e-6, the same as the reference's
// float32 LAPACK eigh.  Columns sorted by ascending eigenvalue; each column's component of largest
// magnitude is made positive (LAPACK leaves the sign implementation-defined; the result of the
// normaliser does not depend on it for well-posed stain planes, SURVEY.md section 7 H-a).
__device__ __forceinline__ void jacobi_rotate(float &app, float &aqq, float &apq, float &arp, float &arq, float &v0p, float &v0q, float &v1p, float &v1q, float &v2p, float &v2q) {
    if (apq == 0.0f) return;
    const float d = aqq - app, b2 = 2.0f * apq;
    // t = sgn(theta) / (|theta| + sqrt(theta^2 + 1)), theta = d / b2, without forming theta
    const float t = copysignf(1.0f, d) * b2 * fast_rcp(fabsf(d) + sqrtf(__fmaf_rn(d, d, b2 * b2)));
    const float c = rsqrtf(__fmaf_rn(t, t, 1.0f)), sn = t * c;
    app = __fmaf_rn(-t, apq, app);
    aqq = __fmaf_rn(t, apq, aqq);
    apq = 0.0f;
    float x = arp, y = arq;
    arp = c * x - sn * y; arq = sn * x + c * y;
    x = v0p; y = v0q; v0p = c * x - sn * y; v0q = sn * x + c * y;
    x = v1p; y = v1q; v1p = c * x - sn * y; v1q = sn * x + c * y;
    x = v2p; y = v2q; v2p = c * x - sn * y; v2q = sn * x + c * y;
}

// C symmetric (only the upper triangle is read); V columns = eigenvectors, w ascending.
__device__ void eigh3(const double C[3][3], float V[3][3], float w[3]) {
    const double tr = fabs(C[0][0]) + fabs(C[1][1]) + fabs(C[2][2]);
    const double sc = tr > 0.0 ? 1.0 / tr : 1.0;
    float a00 = (float)(C[0][0] * sc), a01 = (float)(C[0][1] * sc), a02 = (float)(C[0][2] * sc);
    float a11 = (float)(C[1][1] * sc), a12 = (float)(C[1][2] * sc), a22 = (float)(C[2][2] * sc);
    float v00 = 1, v01 = 0, v02 = 0, v10 = 0, v11 = 1, v12 = 0, v20 = 0, v21 = 0, v22 = 1;
    for (int sweep = 0; sweep < 12; ++sweep) {
        const float off = fabsf(a01) + fabsf(a02) + fabsf(a12);
        if (off <= 1e-12f * (fabsf(a00) + fabsf(a11) + fabsf(a22))) break;
        jacobi_rotate(a00, a11, a01, a02, a12, v00, v01, v10, v11, v20, v21);  // (p, q) = (0, 1), r = 2
        jacobi_rotate(a00, a22, a02, a01, a12, v00, v02, v10, v12, v20, v22);  // (0, 2), r = 1
        jacobi_rotate(a11, a22, a12, a01, a02, v01, v02, v11, v12, v21, v22);  // (1, 2), r = 0
    }
    float ev[3] = {a00, a11, a22};
    float vc[3][3] = {{v00, v10, v20}, {v01, v11, v21}, {v02, v12, v22}};  // vc[j] = eigenvector j
    // sorting network on three (value, vector) pairs
#pragma unroll
    for (int pass = 0; pass < 3; ++pass) {
        const int i = pass == 1 ? 1 : 0;
        if (ev[i] > ev[i + 1]) {
            const float te = ev[i]; ev[i] = ev[i + 1]; ev[i + 1] = te;
#pragma unroll
            for (int k = 0; k < 3; ++k) { const float tv = vc[i][k]; vc[i][k] = vc[i + 1][k]; vc[i + 1][k] = tv; }
        }
    }
#pragma unroll
    for (int j = 0; j < 3; ++j) {
        const float ax = fabsf(vc[j][0]), ay = fabsf(vc[j][1]), az = fabsf(vc[j][2]);
        const float big = ax >= ay ? (ax >= az ? vc[j][0] : vc[j][2]) : (ay >= az ? vc[j][1] : vc[j][2]);
        const float sg = big < 0.0f ? -1.0f : 1.0f;
        // one normalisation removes the drift of the float32 rotations
        const float inv = rsqrtf(vc[j][0] * vc[j][0] + vc[j][1] * vc[j][1] + vc[j][2] * vc[j][2]) * sg;
        w[j] = (float)((double)ev[j] * tr);
#pragma unroll
        for (int i = 0; i < 3; ++i) V[i][j] = vc[j][i] * inv;
    }
}

// Row j of a (2 x 3) linear map of OD, rewritten as an affine map of l:
//   sum_c a[c] OD_c = ln2 log2(240) sum_c a[c] - sum_c (ln2 a[c]) l_c
__device__ __forceinline__ void od_map_to_l(const float a[3], float *aff) {
    const double ln2 = 0.693147180559945309, l240 = 7.906890595608519;
    aff[0] = (float)(-ln2 * a[0]);
    aff[1] = (float)(-ln2 * a[1]);
    aff[2] = (float)(-ln2 * a[2]);
    aff[3] = (float)(ln2 * l240 * ((double)a[0] + (double)a[1] + (double)a[2]));
}

// E = eigvecs[:, (1, 2)] of the unbiased covariance held (as shifted raw moments of l) in m[0..9];
// cov(OD) = ln2^2 cov(l) has the same eigenvectors.  Also the projection maps of the ANGLE stage.
__device__ void basis_from_moments(const double *m, SlotState &st) {
    const double n = m[0];
    st.n_sel = (long long)(n + 0.5);
    double C[3][3] = {{0, 0, 0}, {0, 0, 0}, {0, 0, 0}};
    if (n > 1.0) {  // unbiased covariance about the mean (L393-397); n <= 1 -> zeros (L395-396)
        const double sx = m[1], sy = m[2], sz = m[3], rn = 1.0 / n, rd = 1.0 / (n - 1.0);
        C[0][0] = (m[4] - sx * sx * rn) * rd; C[0][1] = (m[5] - sx * sy * rn) * rd; C[0][2] = (m[6] - sx * sz * rn) * rd;
        C[1][1] = (m[7] - sy * sy * rn) * rd; C[1][2] = (m[8] - sy * sz * rn) * rd; C[2][2] = (m[9] - sz * sz * rn) * rd;
        C[1][0] = C[0][1]; C[2][0] = C[0][2]; C[2][1] = C[1][2];
    }
    float V[3][3], w[3];
    eigh3(C, V, w);
    float e0[3], e1[3];
#pragma unroll
    for (int i = 0; i < 3; ++i) {
        st.e[i * 2] = e0[i] = V[i][1];      // L415: middle eigenvector
        st.e[i * 2 + 1] = e1[i] = V[i][2];  //       largest
    }
    od_map_to_l(e0, st.proj);
    od_map_to_l(e1, st.proj + 4);
}

// One thread per slot.
__global__ void basis_kernel(void *ws_base, int64_t slots, int64_t slot0, int64_t count, int allow_fallback) {
    Ws ws(ws_base, slots);
    const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= count) return;
    const int64_t slot = slot0 + idx;
    SlotState &st = ws.state[slot];
    double m[10];
    for (int i = 0; i < 10; ++i) m[i] = mom_from_fx(ws.moments[slot * 12 + i]);
    st.n_all = ws.moments[slot * 12 + 10];
    st.use_all = 0;
    if (allow_fallback && m[0] < 3.0) {  // L409-410: handled by fallback_kernel
        st.use_all = 1;
        return;
    }
    basis_from_moments(m, st);
}

// Fallback (transform only): slots with fewer than 3 masked rows use every row.  One CTA per image;
// CTAs of unflagged slots exit at once, a flagged slot's CTA re-accumulates the whole image without
// the mask and computes its basis (rare path, so it is not spread over several CTAs).
template <typename T, bool VEC>
__global__ void __launch_bounds__(kThreads) fallback_kernel(const T *__restrict__ img, int64_t hw, int64_t slot0, void *ws_base, int64_t slots) {
    __shared__ float tab[256];
    __shared__ long long red[kThreads / 32][10];
    __shared__ double tot[10];
    Ws ws(ws_base, slots);
    const int64_t n = blockIdx.x;
    const int64_t slot = slot0 + n;
    if (!ws.state[slot].use_all) return;
    long long acc[10] = {0, 0, 0, 0, 0, 0, 0, 0, 0, 0};
    float lo[3] = {0, 0, 0}, hi[3] = {0, 0, 0};
    moments_stream<T, VEC, false>(img + n * 3 * hw, hw, hw / Pix<T, VEC>::kPix, 0, kThreads, tab, acc, lo, hi);
    block_sum10(acc, red);
    if (threadIdx.x < 10) {
        ws.moments[slot * 12 + threadIdx.x] = acc[0];
        tot[threadIdx.x] = mom_from_fx(acc[0]);
    }
    __syncthreads();
    if (threadIdx.x == 0) basis_from_moments(tot, ws.state[slot]);
}

// ---- order-statistic passes ----------------------------------------------------------------------
__device__ __forceinline__ unsigned mix32(unsigned x) {
    x ^= x >> 16; x *= 0x7feb352du; x ^= x >> 15; x *= 0x846ca68bu; x ^= x >> 16;
    return x;
}

// LEVEL 0 -- sample pass: every image contributes ~kSampleGroups pixel groups, one per stride-sized
// window at a hashed offset (so that periodic image structure cannot alias with the sampling).
// The offsets depend only on the window index: an image's result does not depend on where in
// the batch it sits.
template <typename T, bool VEC, int STAGE>
__global__ void __launch_bounds__(kThreads) sample_kernel(const T *__restrict__ img, PassGeom g, int pooled, int64_t slot0, void *ws_base, int64_t slots, int64_t max_groups) {
    constexpr int kPix = Pix<T, VEC>::kPix;
    constexpr int kQ = (STAGE == SX_STAGE_CONC) ? 2 : 1;
    __shared__ float tab[256];
    __shared__ unsigned sh[kQ * kBins];
    __shared__ unsigned s_cnt;
    Ws ws(ws_base, slots);
    const int64_t n = blockIdx.x / g.cpi;
    const int chunk = blockIdx.x % g.cpi;
    const int64_t slot = pooled ? 0 : slot0 + n;
    const SlotState &gst = ws.state[slot];
    const RankParams st(gst);
    const float c_lo0 = gst.c_lo[0], c_lo1 = gst.c_lo[1], c_sc0 = gst.c_scale[0], c_sc1 = gst.c_scale[1];
    for (int i = threadIdx.x; i < kQ * kBins; i += kThreads) sh[i] = 0u;
    if (threadIdx.x == 0) {
        s_cnt = 0u;
        if (chunk == 0) ws.state[slot].group_px = kPix;
    }
    __syncthreads();

    const T *image = img + n * 3 * g.hw;
    const int64_t groups = g.hw / kPix;
    const int64_t stride = groups / max_groups > 1 ? groups / max_groups : 1;  // max_groups >= groups: every group (level 2, exact)
    const int64_t nsamp = groups / stride;
    unsigned cnt = 0;
    for (int64_t i = (int64_t)chunk * kThreads + threadIdx.x; i < nsamp; i += (int64_t)g.cpi * kThreads) {
        const unsigned off = stride > 1 ? __umulhi(mix32((unsigned)i * 0x9e3779b9u + 0x85ebca6bu), (unsigned)stride) : 0u;
        const int64_t gi = i * stride + off;
        float l[3][kPix];
        load_l<T, VEC>(image + gi * kPix, g.hw, tab, l);
#pragma unroll
        for (int k = 0; k < kPix; ++k) {
            float v0, v1;
            ranked_values<STAGE>(st, l[0][k], l[1][k], l[2][k], v0, v1);
            if (v0 != v0) continue;  // masked row
            if (STAGE == SX_STAGE_ANGLE) {
                atomicAdd(&sh[__float2int_rz(angle_key(v0)) >> 12], 1u);
            } else {
                atomicAdd(&sh[__float2int_rz(conc_key(v0, c_lo0, c_sc0)) >> 12], 1u);
                atomicAdd(&sh[kBins + (__float2int_rz(conc_key(v1, c_lo1, c_sc1)) >> 12)], 1u);
            }
            ++cnt;
        }
    }
    cnt = (unsigned)__reduce_add_sync(0xffffffffu, cnt);
    if ((threadIdx.x & 31) == 0 && cnt) atomicAdd(&s_cnt, cnt);
    __syncthreads();
    unsigned *h1 = ws.hist1 + slot * 2 * kBins;
    for (int i = threadIdx.x; i < kQ * kBins; i += kThreads)
        if (sh[i]) atomicAdd(&h1[i], sh[i]);
    if (threadIdx.x == 0 && s_cnt) {
        atomicAdd(&ws.counters[slot * 8 + 2], (unsigned long long)s_cnt);
        if (kQ == 2) atomicAdd(&ws.counters[slot * 8 + 3], (unsigned long long)s_cnt);
    }
}

// Rare path of the full pass: value v of query q lies in the bracket (or in an open end).
__device__ __noinline__ void record_cell(const SlotState &st, int q, float v, unsigned *h2, float *vmin, float *vmax) {
    int cell;
    if (v < st.lo_v[q]) cell = 0;
    else if (v < st.hi_v[q]) cell = 1 + min(__float2int_rz(__fmul_rn(__fsub_rn(v, st.lo_v[q]), st.inv_w[q])), kBins - 3);
    else cell = kBins - 1;
    const int sub = q * kBins + cell;
    atomicAdd(&h2[sub], 1u);
    atomic_min_f32(&vmin[sub], v);
    atomic_max_f32(&vmax[sub], v);
}

// LEVEL 1 -- full pass: count the values below each bracket, resolve the bracket into kBins cells.
// A bracket holds 1-3 % of the rows: rare per pixel, but not per warp (a warp tests 256 pixel-queries
// per iteration: ~4 hits).  Hits are appended to a shared-memory queue: a lane with hits reserves its
// slots with ONE shared atomic and writes its (value, query) pairs with predicated stores; the CTA
// drains the queue into the slot's cells with all lanes busy every kDrainEvery iterations.  (A warp-wide
// reservation -- shuffle scan of the lanes' hit counts, one atomic per warp -- costs ~50 instructions per
// warp and iteration against ~5 here: resolve passes 197 / 173 us against 178 / 153 us on 64 x 1024^2
// float32.)  The common path per query and pixel is a subtraction, a shifted add (rows below) and two
// compares.
constexpr int kQueueCap = 2048;

// Shared scratch of the resolve pass: the hit queue.
struct ResolveSmem {
    unsigned below[2];
    unsigned qn;
    unsigned pad;
    uint2 q[kQueueCap];  // (value bits, query)
};

// Streams the groups first, first + stride, ... of one image.  `st` is a shared-memory copy of the
// slot's state, `rs` shared scratch; results go to the slot's global cells / counters.
template <typename T, bool VEC, int STAGE>
__device__ __forceinline__ void resolve_pass(const T *__restrict__ image, int64_t hw, int64_t groups_end, int64_t first, int64_t stride, const float *tab, const SlotState &st, ResolveSmem &rs, unsigned *h2, float *vmin, float *vmax, unsigned long long *counters) {
    constexpr int kPix = Pix<T, VEC>::kPix;
    // ~2.5 % of the pixel-queries hit: 256 threads x kPix x kDrainEvery x 2.5 % stays far below the queue capacity
    constexpr int kDrainEvery = kPix >= 16 ? 4 : (kPix >= 4 ? 12 : 32);
    if (threadIdx.x < 2) rs.below[threadIdx.x] = 0u;
    if (threadIdx.x == 2) rs.qn = 0u;
    __syncthreads();
    // an open end is encoded by moving the bound to -/+ infinity for the below / inside tests;
    // record_cell sorts such values into the catch-all cells
    const float cl0 = st.open_lo[0] ? -INFINITY : st.lo_v[0], ch0 = st.open_hi[0] ? INFINITY : st.hi_v[0];
    const float cl1 = st.open_lo[1] ? -INFINITY : st.lo_v[1], ch1 = st.open_hi[1] ? INFINITY : st.hi_v[1];
    const RankParams rp(st);
    const int lane = threadIdx.x & 31;
    const unsigned q_addr = smem_u32(rs.q), qn_addr = smem_u32(&rs.qn);

    auto drain = [&]() {
        __syncthreads();
        const int cnt = min((int)rs.qn, kQueueCap);
        for (int j = threadIdx.x; j < cnt; j += kThreads) {
            const uint2 e = rs.q[j];
            record_cell(st, (int)e.y, __uint_as_float(e.x), h2, vmin, vmax);
        }
        __syncthreads();
        if (threadIdx.x == 0) rs.qn = 0u;
        __syncthreads();
    };

    unsigned below0 = 0u, below1 = 0u;
    int it = 0;
    stream_groups_uniform<T, VEC>(
        image, hw, groups_end, first, stride, tab,
        [&](const float(&l)[3][kPix], bool valid) {
            float v0[kPix], v1[kPix];
            unsigned m0 = 0u, m1 = 0u, b0 = 0u, b1 = 0u;
#pragma unroll
            for (int k = 0; k < kPix; ++k) {
                ranked_values<STAGE>(rp, l[0][k], l[1][k], l[2][k], v0[k], v1[k]);
                // NaN (masked row): every comparison is false, NaN - x keeps a clear sign bit
                b0 += __float_as_uint(__fsub_rn(v0[k], cl0)) >> 31;  // v < cl (cl = -inf: never)
                b1 += __float_as_uint(__fsub_rn(v1[k], cl1)) >> 31;
                m0 |= ((v0[k] >= cl0) & (v0[k] < ch0)) ? 1u << k : 0u;
                m1 |= ((v1[k] >= cl1) & (v1[k] < ch1)) ? 1u << k : 0u;
            }
            if (!valid) m0 = m1 = b0 = b1 = 0u;  // a thread without a group
            below0 += b0;
            below1 += b1;
            if ((m0 | m1) != 0u) {  // a lane with hits reserves its own slots: one shared atomic per such lane
                const int c = __popc(m0) + __popc(m1);
                unsigned base = 0u;
                asm volatile("atom.shared.add.u32 %0, [%1], %2;" : "=r"(base) : "r"(qn_addr), "r"((unsigned)c) : "memory");
                if (base + (unsigned)c <= (unsigned)kQueueCap) {
                    unsigned addr = q_addr + base * 8u;  // shared-space byte address
#pragma unroll
                    for (int k = 0; k < kPix; ++k) {  // static indices keep v0 / v1 in registers
                        if (m0 & (1u << k)) { st_shared_v2(addr, __float_as_uint(v0[k]), 0u); addr += 8u; }
                        if (m1 & (1u << k)) { st_shared_v2(addr, __float_as_uint(v1[k]), 1u); addr += 8u; }
                    }
                } else {  // queue (nearly) full (degenerate data): every reserved slot below the capacity
                          // must still be written, because the drain reads all of them; the rest is recorded directly
                    unsigned pos = base;
#pragma unroll
                    for (int k = 0; k < kPix; ++k) {
                        if (m0 & (1u << k)) {
                            if (pos < (unsigned)kQueueCap) st_shared_v2(q_addr + pos * 8u, __float_as_uint(v0[k]), 0u);
                            else record_cell(st, 0, v0[k], h2, vmin, vmax);
                            ++pos;
                        }
                        if (m1 & (1u << k)) {
                            if (pos < (unsigned)kQueueCap) st_shared_v2(q_addr + pos * 8u, __float_as_uint(v1[k]), 1u);
                            else record_cell(st, 1, v1[k], h2, vmin, vmax);
                            ++pos;
                        }
                    }
                }
            }
        },
        [&] {
            if (++it == kDrainEvery) { drain(); it = 0; }
        });
    drain();
    below0 = (unsigned)__reduce_add_sync(0xffffffffu, below0);
    below1 = (unsigned)__reduce_add_sync(0xffffffffu, below1);
    if (lane == 0) {
        if (below0) atomicAdd(&rs.below[0], below0);
        if (below1) atomicAdd(&rs.below[1], below1);
    }
    __syncthreads();
    if (threadIdx.x < 2 && rs.below[threadIdx.x]) atomicAdd(&counters[threadIdx.x], (unsigned long long)rs.below[threadIdx.x]);
}

// The sample pass of one slot, split over `parts` CTAs (per-image kernels of the transform
// pipeline): the same hashed groups and keys as sample_kernel, histograms in shared memory (zeroed
// by the caller), sampled rows counted in *s_cnt.  Loads are issued up to four groups at a time.
template <typename T, bool VEC, int STAGE>
__device__ __noinline__ void sample_slot(const T *__restrict__ image, int64_t hw, const float *tab, const SlotState &st, unsigned (*hist)[kBins], unsigned *s_cnt, int part, int parts, int64_t max_groups = kSampleGroups) {
    constexpr int kPix = Pix<T, VEC>::kPix;
    constexpr int kBatch = 4;
    const int64_t groups = hw / kPix;
    const int64_t stride = groups / max_groups > 1 ? groups / max_groups : 1;  // max_groups >= groups: every group (the exact coarse pass of a recovery)
    const int64_t nsamp = groups / stride;
    const RankParams rp(st);
    const float c_lo0 = st.c_lo[0], c_lo1 = st.c_lo[1], c_sc0 = st.c_scale[0], c_sc1 = st.c_scale[1];
    unsigned cnt = 0;
    const int64_t step = (int64_t)parts * kThreads;  // CTA `part` of `parts` takes samples part * kThreads + t, + step, ...
    for (int64_t i0 = (int64_t)part * kThreads + threadIdx.x; i0 < nsamp; i0 += (int64_t)kBatch * step) {
        RawGroup<T, VEC> raw[kBatch];
#pragma unroll
        for (int b = 0; b < kBatch; ++b) {
            const int64_t i = i0 + (int64_t)b * step;
            if (i < nsamp) {
                const unsigned off = stride > 1 ? __umulhi(mix32((unsigned)i * 0x9e3779b9u + 0x85ebca6bu), (unsigned)stride) : 0u;
                raw[b].load(image + (i * stride + off) * kPix, hw);
            }
        }
#pragma unroll
        for (int b = 0; b < kBatch; ++b) {
            if (i0 + (int64_t)b * step >= nsamp) break;
            float l[3][kPix];
            raw[b].to_l(tab, l);
#pragma unroll
            for (int k = 0; k < kPix; ++k) {
                float v0, v1;
                ranked_values<STAGE>(rp, l[0][k], l[1][k], l[2][k], v0, v1);
                if (v0 != v0) continue;  // masked row
                if (STAGE == SX_STAGE_ANGLE) {
                    atomicAdd(&hist[0][__float2int_rz(angle_key(v0)) >> 12], 1u);
                } else {
                    atomicAdd(&hist[0][__float2int_rz(conc_key(v0, c_lo0, c_sc0)) >> 12], 1u);
                    atomicAdd(&hist[1][__float2int_rz(conc_key(v1, c_lo1, c_sc1)) >> 12], 1u);
                }
                ++cnt;
            }
        }
    }
    cnt = (unsigned)__reduce_add_sync(0xffffffffu, cnt);
    if ((threadIdx.x & 31) == 0 && cnt) atomicAdd(s_cnt, cnt);
}

// ---- per-slot rank searches (one CTA per slot) --------------------------------------------------
// Nearest-rank index (torch_backend.py:L362-365): round_half_even(0.01 * q * (n - 1)), in double.
__device__ __forceinline__ long long rank_index(double q, long long n) { return (long long)rint(0.01 * q * (double)(n - 1)); }

// Inclusive prefix sums of TWO kBins histograms side by side (threads 0..127: h0 -> pre[0],
// threads 128..255: h1 -> pre[1]); 32 bins per thread, warp-shuffle scan, two barriers.  Counts are
// 32-bit: a slot holds fewer than 2^32 rows.  SMEM: the sources are shared memory (else global,
// read through L2).  Every source word is read before the first barrier and `pre` is written after
// it, so the sources may alias `pre` (in-place scan).
template <bool SMEM>
__device__ void dual_prefix(const unsigned *h0, const unsigned *h1, unsigned (*pre)[kBins]) {
    __shared__ unsigned wsum[kThreads / 32];
    const int half = threadIdx.x >> 7, t = threadIdx.x & 127, lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const uint4 *src = reinterpret_cast<const uint4 *>((half ? h1 : h0) + t * 32);
    unsigned v[32];
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        const uint4 x = SMEM ? src[i] : __ldcg(src + i);
        v[4 * i] = x.x; v[4 * i + 1] = x.y; v[4 * i + 2] = x.z; v[4 * i + 3] = x.w;
    }
    unsigned sum = 0;
#pragma unroll
    for (int i = 0; i < 32; ++i) { sum += v[i]; v[i] = sum; }
    unsigned incl = sum;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const unsigned y = __shfl_up_sync(0xffffffffu, incl, o);
        if (lane >= o) incl += y;
    }
    if (lane == 31) wsum[warp] = incl;
    __syncthreads();
    unsigned base = incl - sum;
    for (int k = half * 4; k < warp; ++k) base += wsum[k];
    uint4 *dst = reinterpret_cast<uint4 *>(&pre[half][t * 32]);
#pragma unroll
    for (int i = 0; i < 8; ++i) dst[i] = make_uint4(v[4 * i] + base, v[4 * i + 1] + base, v[4 * i + 2] + base, v[4 * i + 3] + base);
    __syncthreads();
}

// Index of the bin holding 0-based rank k: first b with pre[b] > k.
__device__ __forceinline__ int bin_of_rank(const unsigned *pre, long long k) {
    int lo = 0, hi = kBins - 1;
    while (lo < hi) {
        const int mid = (lo + hi) >> 1;
        if ((long long)pre[mid] > k) hi = mid; else lo = mid + 1;
    }
    return lo;
}

// Unit direction (cos, sin) of a diamond angle p in [-2, 2].
__device__ __forceinline__ void diamond_to_unit(float p, double &c, double &s) {
    double x, y;
    const double pd = (double)p;
    if (pd > 1.0) { x = -(pd - 1.0); y = 2.0 - pd; }
    else if (pd < -1.0) { x = 1.0 + pd; y = -2.0 - pd; }
    else { x = 1.0 - fabs(pd); y = pd; }
    const double h2 = x * x + y * y;  // in [0.5, 1]
    double r = (double)rsqrtf((float)h2);
    r = r * (1.5 - 0.5 * h2 * r * r);  // two Newton steps: 1e-7 -> 1e-14 -> exact to double rounding
    r = r * (1.5 - 0.5 * h2 * r * r);
    c = x * r; s = y * r;
}

// After the sample pass: wanted ranks and their brackets, converted from sample-key bins to the
// value space of the full pass.
// `st` is a shared-memory copy of the slot's state (read and updated by thread 0), `pre` a shared
// scratch array of kBins prefix sums; global accumulators are read through L2 (__ldcg).
// `pre[q]` = inclusive prefix sums of query q's sample histogram, m[q] = rows in that sample.
__device__ void bracket_from_prefix(int stage, SlotState &st, unsigned (*pre)[kBins], long long m0, long long m1) {
    if ((threadIdx.x & 127) == 0) {  // threads 0 and 128: one query each
        const int q = threadIdx.x >> 7;
        const long long n = stage == SX_STAGE_ANGLE ? st.n_sel : st.n_all;
        const double pct = stage == SX_STAGE_ANGLE ? (q == 0 ? 1.0 : 99.0) : 99.0;  // L421-422, L447-448
        long long k = rank_index(pct, n);
        if (k > n - 1) k = n - 1;
        if (k < 0) k = 0;
        st.rank[q] = k;
        const long long m = q ? m1 : m0;
        const int group_pixels = st.group_px > 0 ? st.group_px : 16;
        // Inner cells 1 .. kBins-2 tile [lo, hi).  When the rank bracket reaches an end of the
        // sample, the true order statistic may lie beyond the sample's extreme value: that
        // side is left open and its values are collected in a catch-all cell (0 or kBins-1).
        int b_lo = 0, b_hi = kBins - 1, open_lo = 1, open_hi = 1;
        if (m > 0 && n > 0) {
            long long r_lo, r_hi;
            if (m >= n) {  // the "sample" is the whole slot: the coarse bin of rank k is certain
                r_lo = r_hi = k;
            } else {
                // The pixels of one sampled group are neighbours and may be fully correlated:
                // the binomial deviation is taken over groups, not pixels.
                const double ks = (double)k * (double)m / (double)n;
                const double sd = (double)sqrtf((float)((double)group_pixels * (double)m * (0.01 * pct) * (1.0 - 0.01 * pct)));
                r_lo = (long long)floor(ks - kBracketZ * sd) - 2 * group_pixels;
                r_hi = (long long)ceil(ks + kBracketZ * sd) + 2 * group_pixels;
                open_lo = r_lo <= 0;
                open_hi = r_hi >= m - 1;
            }
            b_lo = bin_of_rank(pre[q], r_lo < 0 ? 0 : (r_lo < m - 1 ? r_lo : m - 1));
            b_hi = bin_of_rank(pre[q], r_hi < 0 ? 0 : (r_hi < m - 1 ? r_hi : m - 1));
            if (m >= n) {
                // exact bracket: one guard bin per side, because a value on a bin edge may round
                // differently in the sample key and in the value-space test of the full pass
                b_lo = b_lo > 0 ? b_lo - 1 : 0;
                b_hi = b_hi < kBins - 1 ? b_hi + 1 : kBins - 1;
                open_lo = b_lo == 0;
                open_hi = b_hi == kBins - 1;
            }
        }
        // sample keys -> values: ANGLE key = (p + 2) 2^22, CONC key = (C - c_lo) c_scale
        double lo_v, hi_v;
        if (stage == SX_STAGE_ANGLE) {
            lo_v = (double)b_lo * (4096.0 / 4194304.0) - 2.0;
            hi_v = (double)(b_hi + 1) * (4096.0 / 4194304.0) - 2.0;
        } else {
            const double bin_w = 4096.0 / (double)st.c_scale[q];
            lo_v = (double)st.c_lo[q] + (double)b_lo * bin_w;
            hi_v = (double)st.c_lo[q] + (double)(b_hi + 1) * bin_w;
        }
        st.lo_v[q] = (float)lo_v;
        st.hi_v[q] = (float)hi_v;
        st.inv_w[q] = __fdiv_rn((float)(kBins - 2), __fsub_rn(st.hi_v[q], st.lo_v[q]));
        st.open_lo[q] = open_lo;
        st.open_hi[q] = open_hi;
    }
    __syncthreads();
}

// Development (SX_ENABLE_TUNING, sx_macenko_set_tuning bit 1): every SAMPLE bracket is replaced by an empty one far
// away from the data, so that the rank search misses and the exact recovery path runs (tests of that path).
__device__ int g_force_miss = 0;
__device__ __forceinline__ void maybe_force_miss(SlotState &st) {
    if (g_force_miss && threadIdx.x < 2) {
        st.lo_v[threadIdx.x] = 3.0e30f; st.hi_v[threadIdx.x] = 3.1e30f; st.inv_w[threadIdx.x] = 0.0f;
        st.open_lo[threadIdx.x] = st.open_hi[threadIdx.x] = 0;
    }
    __syncthreads();
}

// The same from the global sample histograms of the phase-level API.
__device__ void bracket_slot(const Ws &ws, int64_t slot, int stage, SlotState &st, unsigned (*pre)[kBins]) {
    const unsigned *h = ws.hist1 + slot * 2 * kBins;
    // ANGLE: both queries read histogram 0; CONC: query q reads histogram q
    dual_prefix<false>(h, stage == SX_STAGE_ANGLE ? h : h + kBins, pre);
    const long long m0 = (long long)__ldcg(ws.counters + slot * 8 + 2);
    const long long m1 = stage == SX_STAGE_ANGLE ? m0 : (long long)__ldcg(ws.counters + slot * 8 + 3);
    if (threadIdx.x == 0) {  // the group size all ranks agree on (see t_moments_kernel)
        const float gp = __ldcg(ws.odrange + slot * 8 + 6);
        if (gp > (float)st.group_px) st.group_px = (int)gp;
    }
    __syncthreads();
    bracket_from_prefix(stage, st, pre, m0, m1);
    if (m0 < (stage == SX_STAGE_ANGLE ? st.n_sel : st.n_all)) maybe_force_miss(st);  // never an exact (whole-slot) bracket
    else __syncthreads();
}

// Cooperative copy of a slot's state between global memory (through L2) and shared memory.
__device__ __forceinline__ void load_state(SlotState *dst_smem, const SlotState *src) {
    const unsigned *s32 = reinterpret_cast<const unsigned *>(src);
    unsigned *d32 = reinterpret_cast<unsigned *>(dst_smem);
    for (int i = threadIdx.x; i < (int)(sizeof(SlotState) / 4); i += blockDim.x) d32[i] = __ldcg(s32 + i);
    __syncthreads();
}
__device__ __forceinline__ void store_state(SlotState *dst, const SlotState *src_smem) {
    __syncthreads();
    const unsigned *s32 = reinterpret_cast<const unsigned *>(src_smem);
    unsigned *d32 = reinterpret_cast<unsigned *>(dst);
    for (int i = threadIdx.x; i < (int)(sizeof(SlotState) / 4); i += blockDim.x) d32[i] = s32[i];
}

__global__ void __launch_bounds__(kThreads) bracket_kernel(void *ws_base, int64_t slots, int64_t slot0, int stage) {
    __shared__ __align__(16) unsigned pre[2][kBins];
    __shared__ SlotState st;
    Ws ws(ws_base, slots);
    const int64_t slot = slot0 + blockIdx.x;
    load_state(&st, ws.state + slot);
    bracket_slot(ws, slot, stage, st, pre);
    store_state(ws.state + slot, &st);
}

// After the full pass: the order statistics; (ANGLE) HE, pinv, concentration maps; (CONC) maxC.
// `rg` = per-channel (-min l, max l) of the slot (global odrange, or the state's copy); sets st.miss.
__device__ void select_slot(const Ws &ws, int64_t slot, int stage, SlotState &st, const float *rg, unsigned (*pre)[kBins]) {
    const int64_t base = slot * 2 * kBins;
    if (threadIdx.x == 0) st.miss = 0;
    dual_prefix<false>(ws.hist2 + base, ws.hist2 + base + kBins, pre);
    if ((threadIdx.x & 127) == 0) {  // threads 0 and 128: one query each
        const int q = threadIdx.x >> 7;
        const long long inside = (long long)pre[q][kBins - 1];
        long long k = st.rank[q] - (long long)__ldcg(ws.counters + slot * 8 + q);
        if (k < 0 || k >= inside) {  // the rank fell outside the bracket (not expected): nearest edge
            atomicOr(&ws.status[slot * 4], 1 << q);
            st.miss = 1;
            k = k < 0 ? 0 : (inside > 0 ? inside - 1 : 0);
        }
        const int cell = bin_of_rank(pre[q], k);
        const long long before = cell > 0 ? (long long)pre[q][cell - 1] : 0;
        const long long cnt = (long long)pre[q][cell] - before;
        const float lo = __ldcg(ws.vmin + base + q * kBins + cell), hi = __ldcg(ws.vmax + base + q * kBins + cell);
        float v = lo;
        if (cnt > 1 && hi > lo) v = lo + (hi - lo) * __fdiv_rn((float)(k - before), (float)(cnt - 1));
        st.val[q] = v;
    }
    __syncthreads();
    float *fit = ws.fit + slot * 8;
    if (threadIdx.x == 0) {
        if (stage == SX_STAGE_ANGLE) {
            // M7 (L425-439): v = E (cos phi, sin phi); HE columns ordered by first component.
            double c0, s0, c1, s1;
            diamond_to_unit(st.val[0], c0, s0);
            diamond_to_unit(st.val[1], c1, s1);
            float vmin[3], vmax[3];
            for (int i = 0; i < 3; ++i) {
                vmin[i] = (float)((double)st.e[i * 2] * c0 + (double)st.e[i * 2 + 1] * s0);
                vmax[i] = (float)((double)st.e[i * 2] * c1 + (double)st.e[i * 2 + 1] * s1);
            }
            const bool min_first = vmin[0] > vmax[0];
            float he[6];
            for (int i = 0; i < 3; ++i) { he[i * 2] = min_first ? vmin[i] : vmax[i]; he[i * 2 + 1] = min_first ? vmax[i] : vmin[i]; }
            for (int i = 0; i < 6; ++i) fit[i] = he[i];
            // M8 (L444): least squares via the normal equations, in double.
            double a00 = 0, a01 = 0, a11 = 0;
            for (int i = 0; i < 3; ++i) { a00 += (double)he[i * 2] * he[i * 2]; a01 += (double)he[i * 2] * he[i * 2 + 1]; a11 += (double)he[i * 2 + 1] * he[i * 2 + 1]; }
            const double rdet = 1.0 / (a00 * a11 - a01 * a01);
            for (int i = 0; i < 3; ++i) {
                st.pinv[i] = (float)((a11 * he[i * 2] - a01 * he[i * 2 + 1]) * rdet);
                st.pinv[3 + i] = (float)((-a01 * he[i * 2] + a00 * he[i * 2 + 1]) * rdet);
            }
            // concentration rows as affine maps of l for the CONC stage
            od_map_to_l(st.pinv, st.proj);
            od_map_to_l(st.pinv + 3, st.proj + 4);
            // Key range of each concentration row from the per-channel range of l (interval arithmetic).
            for (int j = 0; j < 2; ++j) {
                double lo = st.proj[j * 4 + 3], hi = lo;
                for (int c = 0; c < 3; ++c) {
                    const double w = st.proj[j * 4 + c], a = w * (double)(-rg[c]), b = w * (double)rg[3 + c];
                    lo += fmin(a, b); hi += fmax(a, b);
                }
                const double pad = 1e-6 * (fabs(lo) + fabs(hi)) + 1e-12;
                lo -= pad; hi += pad;
                st.c_lo[j] = (float)lo;
                st.c_scale[j] = __fdiv_rn(16777216.0f, (float)(hi - lo));
            }
        } else {
            fit[6] = st.val[0];  // maxC (L447-449)
            fit[7] = st.val[1];
        }
    }
}

__global__ void __launch_bounds__(kThreads) select_kernel(void *ws_base, int64_t slots, int64_t slot0, int stage) {
    __shared__ __align__(16) unsigned pre[2][kBins];
    __shared__ SlotState st;
    __shared__ float rg[8];
    Ws ws(ws_base, slots);
    const int64_t slot = slot0 + blockIdx.x;
    const int64_t base = slot * 2 * kBins;
    if (threadIdx.x < 8) rg[threadIdx.x] = ws.odrange[slot * 8 + threadIdx.x];
    load_state(&st, ws.state + slot);
    select_slot(ws, slot, stage, st, rg, pre);
    store_state(ws.state + slot, &st);
    if (stage == SX_STAGE_ANGLE) {  // re-arm the slot's histograms and counters for the CONC stage
        for (int i = threadIdx.x; i < 2 * kBins; i += kThreads) {
            ws.hist1[base + i] = 0u;
            ws.hist2[base + i] = 0u;
            ws.vmin[base + i] = INFINITY;
            ws.vmax[base + i] = -INFINITY;
        }
        if (threadIdx.x < 8) ws.counters[slot * 8 + threadIdx.x] = 0ull;
    }
}

// ---- deterministic recovery of a missed bracket ---------------------------------------------------
// The sample bracket holds the wanted rank with probability ~1 - 1e-10 per query for independent
// pixels; structured images can defeat any sampling argument, so a miss must not end in a clamped
// answer.  When select_slot reports a miss, the SAME CTA re-does the stage for its slot exactly:
// coarse histogram over EVERY pixel group of the slot's image(s) (sample_slot with stride 1), bracket =
// the coarse bin of the wanted rank plus one guard bin per side (bracket_from_prefix, m >= n branch),
// cells re-armed, full resolve pass by this CTA, rank search again.  Slow (one CTA streams the slot
// twice) and never expected to run; status[slot][3] counts the stages that took this path.
// `pre` (2 x kBins words) serves as histogram, prefix sums and then as the hit queue of the resolve pass.
static_assert(sizeof(ResolveSmem) <= sizeof(unsigned) * 2 * kBins, "the recovery pass keeps its hit queue in the histogram array");

template <typename T, bool VEC, int STAGE>
__device__ __noinline__ void recover_stage(const T *__restrict__ img, int64_t n_img, int64_t hw, const Ws &ws, int64_t slot, SlotState &st, const float *rg, unsigned (*pre)[kBins], unsigned *s_cnt) {
    const int64_t base = slot * 2 * kBins;
    if (STAGE == SX_STAGE_ANGLE && threadIdx.x == 0) {  // select(ANGLE) has replaced the projection maps by the concentration maps
        float e0[3], e1[3];
        for (int i = 0; i < 3; ++i) { e0[i] = st.e[i * 2]; e1[i] = st.e[i * 2 + 1]; }
        od_map_to_l(e0, st.proj);
        od_map_to_l(e1, st.proj + 4);
    }
    for (int i = threadIdx.x; i < 2 * kBins; i += kThreads) pre[0][i] = 0u;
    if (threadIdx.x == 0) *s_cnt = 0u;
    __syncthreads();
    for (int64_t k = 0; k < n_img; ++k) sample_slot<T, VEC, STAGE>(img + k * 3 * hw, hw, nullptr, st, pre, s_cnt, 0, 1, (int64_t)1 << 62);
    __syncthreads();
    const long long m = (long long)*s_cnt;
    dual_prefix<true>(pre[0], STAGE == SX_STAGE_ANGLE ? pre[0] : pre[1], pre);
    bracket_from_prefix(STAGE, st, pre, m, m);  // m == n: exact bracket
    for (int i = threadIdx.x; i < 2 * kBins; i += kThreads) {
        ws.hist2[base + i] = 0u;
        ws.vmin[base + i] = INFINITY;
        ws.vmax[base + i] = -INFINITY;
    }
    if (threadIdx.x < 2) ws.counters[slot * 8 + threadIdx.x] = 0ull;
    __threadfence();
    __syncthreads();
    ResolveSmem &rs = *reinterpret_cast<ResolveSmem *>(&pre[0][0]);
    for (int64_t k = 0; k < n_img; ++k) {
        resolve_pass<T, VEC, STAGE>(img + k * 3 * hw, hw, hw / Pix<T, VEC>::kPix, 0, kThreads, nullptr, st, rs, ws.hist2 + base, ws.vmin + base, ws.vmax + base, ws.counters + slot * 8);
        __syncthreads();
    }
    __threadfence();
    __syncthreads();
    if (threadIdx.x == 0) {
        atomicAnd(&ws.status[slot * 4], ~3);  // the exact pass decides
        atomicAdd(&ws.status[slot * 4 + 3], 1);
    }
    __syncthreads();
    select_slot(ws, slot, STAGE, st, rg, pre);
    __syncthreads();
}

// select_kernel with the recovery path: one CTA per slot; `pooled` != 0: the slot pools all n_img images (fit),
// else slot slot0 + blockIdx.x belongs to image blockIdx.x (transform).
template <typename T, bool VEC, int STAGE>
__global__ void __launch_bounds__(kThreads) select_recover_kernel(const T *__restrict__ img, int64_t n_img, int64_t hw, int pooled, int64_t slot0, void *ws_base, int64_t slots) {
    __shared__ __align__(16) unsigned pre[2][kBins];
    __shared__ SlotState st;
    __shared__ float rg[8];
    __shared__ unsigned s_cnt;
    Ws ws(ws_base, slots);
    const int64_t slot = slot0 + blockIdx.x;
    const int64_t base = slot * 2 * kBins;
    if (threadIdx.x < 8) rg[threadIdx.x] = ws.odrange[slot * 8 + threadIdx.x];
    load_state(&st, ws.state + slot);
    select_slot(ws, slot, STAGE, st, rg, pre);
    __syncthreads();
    if (st.miss) recover_stage<T, VEC, STAGE>(pooled ? img : img + (int64_t)blockIdx.x * 3 * hw, pooled ? n_img : 1, hw, ws, slot, st, rg, pre, &s_cnt);
    store_state(ws.state + slot, &st);
    if (STAGE == SX_STAGE_ANGLE) {  // re-arm the slot's histograms and counters for the CONC stage
        for (int i = threadIdx.x; i < 2 * kBins; i += kThreads) {
            ws.hist1[base + i] = 0u;
            ws.hist2[base + i] = 0u;
            ws.vmin[base + i] = INFINITY;
            ws.vmax[base + i] = -INFINITY;
        }
        if (threadIdx.x < 8) ws.counters[slot * 8 + threadIdx.x] = 0ull;
    }
}

// Truncation of 0 <= x < 2^23 on the FMA pipe: adding 2^23 with round-toward-zero leaves trunc(x) in
// the low mantissa bits (F2I / I2F / FRND run on the SFU, which the exponentials already load).
__device__ __forceinline__ unsigned trunc_byte(float x) { return __float_as_uint(__fadd_rz(x, 8388608.0f)) & 0xffu; }  // x in [0, 256)
__device__ __forceinline__ float trunc_small(float x) { return __fsub_rn(__fadd_rz(x, 8388608.0f), 8388608.0f); }
// q / 255 correctly rounded for integer-valued 0 <= q <= 255 (checked exhaustively against the true
// division of _template.py:L111-112): one Newton correction of q * (1 / 255).
__device__ __forceinline__ float div255_exact(float q) {
    const float c = 1.0f / 255.0f;
    const float r0 = __fmul_rn(q, c);
    return __fmaf_rn(__fmaf_rn(-255.0f, r0, q), c, r0);
}

// coef[c][0..2] = M3 row c, coef[c][3] = bias (shared memory, written by threads 0..2).
template <typename T, int OUT>
__device__ __forceinline__ void apply_coefficients(float *coef, const float *__restrict__ he_ref, const float *__restrict__ maxc_ref, const float *pinv, float maxc0, float maxc1) {
    if (threadIdx.x < 3) {
        const int c = threadIdx.x;
        const float n0 = __fdiv_rn(maxc_ref[0], maxc0), n1 = __fdiv_rn(maxc_ref[1], maxc1);  // L452
        float sum = 0.0f;
        for (int k = 0; k < 3; ++k) {
            const float m = he_ref[c * 2] * n0 * pinv[k] + he_ref[c * 2 + 1] * n1 * pinv[3 + k];
            coef[c * 4 + k] = m;
            sum += m;
        }
        float bias = kLog2_240 * (1.0f - sum);
        if (OUT == 2 && sizeof(T) == 4) bias -= 7.994353436858858f;  // log2(255): fold the /255
        coef[c * 4 + 3] = bias;
    }
}

// Streams the groups first, first + stride, ... of one image through the reconstruction.
template <typename T, bool VEC, int OUT>
__device__ __forceinline__ void apply_pass(const T *__restrict__ image, void *__restrict__ out_image, int64_t hw, int64_t first, int64_t stride, const float *tab, const float *unit_tab, const float *coef) {
    constexpr int kPix = Pix<T, VEC>::kPix;
    float A[3][4];
#pragma unroll
    for (int c = 0; c < 3; ++c)
#pragma unroll
        for (int k = 0; k < 4; ++k) A[c][k] = coef[c * 4 + k];
    const float top = (OUT == 2 && sizeof(T) == 4) ? 1.0f : 255.0f;  // clamp(.., 0, 255) (L459)

    stream_groups<T, VEC>(image, hw, hw / kPix, first, stride, tab, [&](const float(&l)[3][kPix], int64_t gi) {
        float o[3][kPix];
#pragma unroll
        for (int k = 0; k < kPix; ++k)
#pragma unroll
            for (int c = 0; c < 3; ++c) {
                const float e = __fmaf_rn(A[c][2], l[2][k], __fmaf_rn(A[c][1], l[1][k], __fmaf_rn(A[c][0], l[0][k], A[c][3])));
                o[c][k] = fminf(fast_ex2(e), top);
            }
        const int64_t off = gi * kPix;
        if constexpr (OUT == 0) {
            uint8_t *out = static_cast<uint8_t *>(out_image) + off;
#pragma unroll
            for (int c = 0; c < 3; ++c) {
                if constexpr (VEC) {
                    unsigned w[4] = {0, 0, 0, 0};
#pragma unroll
                    for (int k = 0; k < kPix; ++k) w[k >> 2] |= trunc_byte(o[c][k]) << (8 * (k & 3));
                    st_stream(reinterpret_cast<uint4 *>(out + c * hw), make_uint4(w[0], w[1], w[2], w[3]));
                } else {
                    out[c * hw] = (uint8_t)__float2int_rz(o[c][0]);
                }
            }
        } else if constexpr (sizeof(T) == 2) {
            // 16-bit float input: the reference casts its float32 [0,255] result back to the input dtype (L131), then
            // normalize_to_0_1 divides THAT by 255 in the same dtype (_template.py:L111-112): two roundings.
            T *out = static_cast<T *>(out_image) + off;
#pragma unroll
            for (int c = 0; c < 3; ++c) {
                if constexpr (OUT == 2) {
#pragma unroll
                    for (int k = 0; k < kPix; ++k) o[c][k] = __fdiv_rn(round_trip<T>(o[c][k]), 255.0f);
                }
                if constexpr (VEC) {
                    st_stream(reinterpret_cast<uint4 *>(out + c * hw), make_uint4(Half2IO<T>::pack(o[c][0], o[c][1]), Half2IO<T>::pack(o[c][2], o[c][3]), Half2IO<T>::pack(o[c][4], o[c][5]), Half2IO<T>::pack(o[c][6], o[c][7])));
                } else {
                    out[c * hw] = Half2IO<T>::narrow(o[c][0]);
                }
            }
        } else {
            float *out = static_cast<float *>(out_image) + off;
#pragma unroll
            for (int c = 0; c < 3; ++c) {
                if constexpr (sizeof(T) == 1) {
                    // uint8 input: the reference truncates to uint8 first (L560), then casts / divides
#pragma unroll
                    for (int k = 0; k < kPix; ++k) {
                        const float q = trunc_small(o[c][k]);
                        o[c][k] = OUT == 2 ? div255_exact(q) : q;
                    }
                }
                if constexpr (VEC) {
#pragma unroll
                    for (int k = 0; k < kPix; k += 4) st_stream(reinterpret_cast<float4 *>(out + c * hw + k), make_float4(o[c][k], o[c][k + 1], o[c][k + 2], o[c][k + 3]));
                } else {
                    out[c * hw] = o[c][0];
                }
            }
        }
    }, [] {});
}

// ---- apply (M10) ---------------------------------------------------------------------------------
// OD' = he_ref . diag(maxc_ref / maxC) . pinv . OD is a 3x3 map M3 of OD.  With l = log2(255x+1):
//   240 exp(-OD'_c) = 2^( b_c + sum_k M3[c][k] l_k ),   b_c = log2(240) (1 - sum_k M3[c][k]).
// OUT: 0 = uint8 (truncated), 1 = float32 in [0,255], 2 = float32 / 255 (normalize_to_0_1).
template <typename T, bool VEC, int OUT>
__global__ void __launch_bounds__(kThreads) apply_kernel(const T *__restrict__ img, void *__restrict__ out_, PassGeom g, int64_t slot0, const float *__restrict__ he_ref, const float *__restrict__ maxc_ref, void *ws_base, int64_t slots) {
    __shared__ float tab[256];
    __shared__ float unit_tab[256];
    __shared__ float coef[12];
    Ws ws(ws_base, slots);
    const int64_t n = blockIdx.x / g.cpi;
    const int chunk = blockIdx.x % g.cpi;
    const int64_t slot = slot0 + n;
    apply_coefficients<T, OUT>(coef, he_ref, maxc_ref, ws.state[slot].pinv, ws.fit[slot * 8 + 6], ws.fit[slot * 8 + 7]);
    __syncthreads();
    constexpr int kOutBytes = OUT == 0 ? 1 : (sizeof(T) == 2 ? 2 : 4);  // 16-bit float input: output in the input's dtype
    apply_pass<T, VEC, OUT>(img + n * 3 * g.hw, static_cast<char *>(out_) + n * 3 * g.hw * kOutBytes, g.hw, (int64_t)chunk * kThreads, (int64_t)g.cpi * kThreads, tab, unit_tab, coef);
}

// ---- per-image transform pipeline ------------------------------------------------------------
// sx_macenko_transform = init + EIGHT launches per chain: three lean streaming kernels over the chain's images
// (moments, resolve ANGLE, resolve CONC), the reconstruction, and between them small per-image kernels:
//   mid<ANGLE>             after moments:       basis (M3-M4), masked-row fallback (L409-410), ANGLE sample + bracket
//   select_recover<ANGLE>  after resolve ANGLE: rank search, HE / pinv (M7-M8); exact re-run of the stage on a missed bracket
//   mid<CONC>              then:                CONC sample + bracket
//   select_recover<CONC>   after resolve CONC:  rank search -> maxC (M9); same recovery
// The phase-level API needs 13 launches for the same work because a sharded fit must all-reduce
// between them.  (Folding the per-image steps into the tail of the streaming kernels -- "last CTA of
// an image finishes it" -- was built and measured: the extra registers and shared memory slowed the
// streaming loops by as much as the saved launches, 1004 us against 1012 us on 64 x 1024^2 float32,
// and made uint8 slower.)  The arithmetic is the phase-level API's (same functions, same sample
// groups), so both paths give the same statistics.
//
// Work decomposition of the streaming kernels: the batch is one list of ROWS (kThreads consecutive
// pixel groups of one image); CTA c of a grid sized to the resident capacity of the device owns the
// contiguous rows [c R / grid, (c + 1) R / grid) -- equal work for every CTA whatever the batch and
// image sizes, at most two images per CTA for batches larger than the grid.
struct RowGeom {
    int64_t n_img, hw, total_rows, slot0;  // image i of this launch uses statistics slot slot0 + i (pooled: every image slot0)
    int rows_per_img, pooled;
};

// The segment (image, rows [row0, row1)) at row `r` of a CTA's range ending at `r_end`.
struct RowSegment {
    int64_t n;
    int row0, row1;
    __device__ __forceinline__ RowSegment(const RowGeom &g, int64_t r, int64_t r_end) {
        n = r / g.rows_per_img;
        row0 = (int)(r - n * g.rows_per_img);
        const int64_t want = (int64_t)row0 + (r_end - r);
        row1 = (int)(want < (int64_t)g.rows_per_img ? want : (int64_t)g.rows_per_img);
    }
};

template <typename T, bool VEC>
__global__ void __launch_bounds__(kThreads) t_moments_kernel(const T *__restrict__ img, RowGeom g, void *ws_base, int64_t slots) {
    constexpr int kPix = Pix<T, VEC>::kPix;
    __shared__ long long red[kThreads / 32][10];
    __shared__ float redf[kThreads / 32][2];
    Ws ws(ws_base, slots);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int64_t groups = g.hw / kPix;
    // The batch is split over the CTAs in UNITS of MomFlush::kRows rows that never straddle an image and always start at
    // a multiple of kRows within it (see moments_stream): a CTA owns the units [c U / grid, (c + 1) U / grid).
    constexpr int kUnitRows = MomFlush<T, VEC>::kRows;
    const int64_t upi = ((int64_t)g.rows_per_img + kUnitRows - 1) / kUnitRows, total_units = g.n_img * upi;
    const int64_t u_end = (int64_t)(blockIdx.x + 1) * total_units / gridDim.x;
    for (int64_t u = (int64_t)blockIdx.x * total_units / gridDim.x; u < u_end;) {
        struct { int64_t n; int row0, row1; } seg;
        seg.n = u / upi;
        const int64_t ul0 = u - seg.n * upi, want = ul0 + (u_end - u), ul1 = want < upi ? want : upi;
        u += ul1 - ul0;
        seg.row0 = (int)(ul0 * kUnitRows);
        seg.row1 = (int)(ul1 * kUnitRows < (int64_t)g.rows_per_img ? ul1 * kUnitRows : (int64_t)g.rows_per_img);
        const int64_t slot = g.pooled ? g.slot0 : g.slot0 + seg.n;
        const int64_t seg_end = (int64_t)seg.row1 * kThreads < groups ? (int64_t)seg.row1 * kThreads : groups;
        long long acc[10] = {0, 0, 0, 0, 0, 0, 0, 0, 0, 0};
        float lo[3] = {INFINITY, INFINITY, INFINITY}, hi[3] = {-INFINITY, -INFINITY, -INFINITY};
        moments_stream<T, VEC, true>(img + seg.n * 3 * g.hw, g.hw, seg_end, (int64_t)seg.row0 * kThreads, kThreads, nullptr, acc, lo, hi);
        {
            const float a = warp_max(-lo[0]), b = warp_max(hi[0]);
            if (lane == 0) { redf[warp][0] = a; redf[warp][1] = b; }
        }
        block_sum10(acc, red);
        if (threadIdx.x < 10) {
            if (acc[0] != 0) atomicAdd(reinterpret_cast<unsigned long long *>(&ws.moments[slot * 12 + threadIdx.x]), (unsigned long long)acc[0]);
        } else if (threadIdx.x >= 32 && threadIdx.x < 38) {
            const int i = threadIdx.x - 32;  // [0..2] = -min l (one range for the three channels), [3..5] = max l
            float v = -INFINITY;
            for (int k = 0; k < kThreads / 32; ++k) v = fmaxf(v, redf[k][i / 3]);
            atomic_max_f32(&ws.odrange[slot * 8 + i], v);
        } else if (threadIdx.x == 64 && seg.row0 == 0) {
            atomicAdd(reinterpret_cast<unsigned long long *>(&ws.moments[slot * 12 + 10]), (unsigned long long)g.hw);  // rows in the slot (pooled fit: all images)
            // pixels per sampled group of THIS rank's kernel variant; ODRANGE is MAX-combined over the ranks of a sharded
            // fit, so every rank sizes its sample brackets with the same (largest) group -- ranks whose shards differ in
            // vectorisation, or hold no images at all, would otherwise define different cells for the summed histograms
            atomic_max_f32(&ws.odrange[slot * 8 + 6], (float)kPix);
        }
        __syncthreads();  // red / redf are rewritten by the next segment
    }
}

template <typename T, bool VEC, int STAGE>
__global__ void __launch_bounds__(kThreads) t_resolve_kernel(const T *__restrict__ img, RowGeom g, void *ws_base, int64_t slots) {
    constexpr int kPix = Pix<T, VEC>::kPix;
    __shared__ SlotState st;
    __shared__ ResolveSmem rs;
    Ws ws(ws_base, slots);
    const int64_t groups = g.hw / kPix;
    const int64_t r_end = (int64_t)(blockIdx.x + 1) * g.total_rows / gridDim.x;
    for (int64_t r = (int64_t)blockIdx.x * g.total_rows / gridDim.x; r < r_end;) {
        const RowSegment seg(g, r, r_end);
        r += seg.row1 - seg.row0;
        const int64_t slot = g.pooled ? g.slot0 : g.slot0 + seg.n, base = slot * 2 * kBins;
        const int64_t seg_end = (int64_t)seg.row1 * kThreads < groups ? (int64_t)seg.row1 * kThreads : groups;
        __syncthreads();  // the previous segment has finished with `st`
        load_state(&st, ws.state + slot);
        resolve_pass<T, VEC, STAGE>(img + seg.n * 3 * g.hw, g.hw, seg_end, (int64_t)seg.row0 * kThreads, kThreads, nullptr, st, rs, ws.hist2 + base, ws.vmin + base, ws.vmax + base, ws.counters + slot * 8);
    }
}

// Shared memory of the per-image kernels.
struct MidSmem {
    SlotState st;
    unsigned s_cnt;
    int s_last;
    double tot[12];
};

// Per-image step before a resolve pass, kMidParts CTAs per image: (ANGLE only: basis M3-M4 and the
// masked-row fallback L409-410, computed redundantly by every CTA of the image -- identical inputs,
// identical results), then each CTA samples its share of the image's sample groups into a
// shared-memory histogram and adds it to the slot's global sample histogram; the CTA that arrives
// last turns the histogram into the brackets of the stage and stores the slot's state.
constexpr int kMidParts = 8;

template <typename T, bool VEC, int STAGE>
__global__ void __launch_bounds__(kThreads) mid_kernel(const T *__restrict__ img, int64_t hw, int64_t slot0, void *ws_base, int64_t slots) {
    __shared__ MidSmem ms;
    __shared__ long long red[kThreads / 32][10];
    __shared__ __align__(16) unsigned hist[2][kBins];
    Ws ws(ws_base, slots);
    const int64_t slot = slot0 + blockIdx.x / kMidParts;
    const int part = blockIdx.x % kMidParts;
    const T *image = img + (int64_t)(blockIdx.x / kMidParts) * 3 * hw;
    const int64_t base = slot * 2 * kBins;
    if constexpr (STAGE == SX_STAGE_ANGLE) {
        if (threadIdx.x < 10) ms.tot[threadIdx.x] = mom_from_fx(ws.moments[slot * 12 + threadIdx.x]);
        __syncthreads();
        if (threadIdx.x == 0) {
            SlotState z = {};
            ms.st = z;
            ms.st.n_all = (long long)hw;
            ms.st.use_all = ms.tot[0] < 3.0;
            if (!ms.st.use_all) basis_from_moments(ms.tot, ms.st);
        }
        __syncthreads();
        if (ms.st.use_all) {  // fewer than 3 rows pass the mask: every row (rare; this CTA re-reads the image)
            long long acc[10] = {0, 0, 0, 0, 0, 0, 0, 0, 0, 0};
            float lo[3] = {0, 0, 0}, hi[3] = {0, 0, 0};
            moments_stream<T, VEC, false>(image, hw, hw / Pix<T, VEC>::kPix, 0, kThreads, nullptr, acc, lo, hi);
            block_sum10(acc, red);
            if (threadIdx.x < 10) ms.tot[threadIdx.x] = mom_from_fx(acc[0]);
            __syncthreads();
            if (threadIdx.x == 0) basis_from_moments(ms.tot, ms.st);
            __syncthreads();
        }
    } else {
        load_state(&ms.st, ws.state + slot);
    }
    for (int i = threadIdx.x; i < 2 * kBins; i += kThreads) hist[0][i] = 0u;
    if (threadIdx.x == 0) {
        ms.s_cnt = 0u;
        ms.st.group_px = Pix<T, VEC>::kPix;
    }
    __syncthreads();
    sample_slot<T, VEC, STAGE>(image, hw, nullptr, ms.st, hist, &ms.s_cnt, part, kMidParts);
    __syncthreads();
    constexpr int kQ = STAGE == SX_STAGE_CONC ? 2 : 1;
    for (int i = threadIdx.x; i < kQ * kBins; i += kThreads)
        if (hist[0][i]) atomicAdd(&ws.hist1[base + i], hist[0][i]);
    if (threadIdx.x == 0 && ms.s_cnt) {
        atomicAdd(&ws.counters[slot * 8 + 2], (unsigned long long)ms.s_cnt);
        if (kQ == 2) atomicAdd(&ws.counters[slot * 8 + 3], (unsigned long long)ms.s_cnt);
    }
    __threadfence();  // this thread's global atomics are visible before the arrival below
    __syncthreads();
    if (threadIdx.x == 0) {
        ms.s_last = atomicAdd(&ws.status[slot * 4 + 1 + STAGE], 1) == kMidParts - 1;
        __threadfence();
    }
    __syncthreads();
    if (!ms.s_last) return;
    if (STAGE == SX_STAGE_ANGLE && threadIdx.x < 11) {  // the moments the basis was computed from (fallback: every row)
        if (ms.st.use_all && threadIdx.x < 10) ws.moments[slot * 12 + threadIdx.x] = mom_to_fx(ms.tot[threadIdx.x]);
        if (threadIdx.x == 10) ws.moments[slot * 12 + 10] = (long long)hw;
    }
    bracket_slot(ws, slot, STAGE, ms.st, hist);
    store_state(ws.state + slot, &ms.st);
}

// uint8 in, float32 out (OUT 1: [0,255], OUT 2: [0,1]): FOUR pixels per thread.  With 16 pixels per
// thread each lane would store 64 contiguous bytes per plane, i.e. every 128-bit store of a warp
// hits 32 different half-filled sectors (measured: 3.3 TB/s); with 4 pixels a warp's store is one
// contiguous 512-byte run, exactly like the float32 path, and its load one 128-byte run.
template <int OUT>
__global__ void __launch_bounds__(kThreads) apply_u8_f32_kernel(const uint8_t *__restrict__ img, float *__restrict__ out, PassGeom g, int64_t slot0, const float *__restrict__ he_ref, const float *__restrict__ maxc_ref, void *ws_base, int64_t slots) {
    __shared__ float coef[12];
    Ws ws(ws_base, slots);
    const int64_t n = blockIdx.x / g.cpi;
    const int chunk = blockIdx.x % g.cpi;
    const int64_t slot = slot0 + n;
    apply_coefficients<float, 1>(coef, he_ref, maxc_ref, ws.state[slot].pinv, ws.fit[slot * 8 + 6], ws.fit[slot * 8 + 7]);
    __syncthreads();
    float A[3][4];
#pragma unroll
    for (int c = 0; c < 3; ++c)
#pragma unroll
        for (int k = 0; k < 4; ++k) A[c][k] = coef[c * 4 + k];
    const uint8_t *image = img + n * 3 * g.hw;
    float *oimg = out + n * 3 * g.hw;
    const int64_t groups = g.hw / 4, stride = (int64_t)g.cpi * kThreads;
    int64_t gi = (int64_t)chunk * kThreads + threadIdx.x;
    unsigned cur[3] = {0u, 0u, 0u}, nxt[3] = {0u, 0u, 0u};
    auto load = [&](int64_t i, unsigned (&w)[3]) {
#pragma unroll
        for (int c = 0; c < 3; ++c) asm volatile("ld.global.nc.L1::no_allocate.u32 %0, [%1];" : "=r"(w[c]) : "l"(image + c * g.hw + i * 4));
    };
    if (gi < groups) load(gi, cur);
    for (; gi < groups; gi += stride) {
        if (gi + stride < groups) load(gi + stride, nxt);
        float l[3][4];
#pragma unroll
        for (int c = 0; c < 3; ++c) { l[c][0] = u8_l<0>(cur[c]); l[c][1] = u8_l<1>(cur[c]); l[c][2] = u8_l<2>(cur[c]); l[c][3] = u8_l<3>(cur[c]); }
#pragma unroll
        for (int c = 0; c < 3; ++c) {
            float o[4];
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                const float e = __fmaf_rn(A[c][2], l[2][k], __fmaf_rn(A[c][1], l[1][k], __fmaf_rn(A[c][0], l[0][k], A[c][3])));
                const float q = trunc_small(fminf(fast_ex2(e), 255.0f));  // the reference truncates to uint8 first (L560)
                o[k] = OUT == 2 ? div255_exact(q) : q;
            }
            st_stream(reinterpret_cast<float4 *>(oimg + c * g.hw + gi * 4), make_float4(o[0], o[1], o[2], o[3]));
        }
#pragma unroll
        for (int c = 0; c < 3; ++c) cur[c] = nxt[c];
    }
}

// ---- sharded pooled fit: combine of slot 0's statistics over NVLink peer memory -----------------
// The pooled fit of a sharded reference batch must combine every rank's moments / histograms
// before each per-slot step.  Instead of 2-4 NCCL all-reduces per step (14 per fit, each ~25 us of
// launch + latency for <= 96 KB), ONE kernel per step combines the regions in place:
//   bufs[p] = rank p's one-slot workspace followed by uint32 flagsA[64], flagsB[64], mapped by all
//   ranks (symmetric memory).  (1) publish epoch in flagsA of every peer, wait for all; (2) every
//   thread combines its words of each region over the ranks in rank order with 128-bit peer loads
//   into a private scratch; (3) publish epoch in flagsB ("I have read everybody"), wait for all --
//   only then may the ranks overwrite their own regions; (4) scratch -> own regions.  Every rank
//   ends up with bit-identical combined statistics (same order of additions).
// `which`: 0 after moments (MOMENTS sum, ODRANGE max), 1 after a sample pass (HIST1, COUNTERS sum),
// 2 after a resolve pass (HIST2, COUNTERS sum, VMIN min, VMAX max).
constexpr int kCombineThreads = 1024;
constexpr int kPeerMaxWorld = 64;

__device__ __forceinline__ uint4 ld_peer_v4(const void *p) {
    uint4 v;
    asm volatile("ld.relaxed.sys.global.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void peer_rendezvous(unsigned char *const *bufs, int world, int rank, unsigned epoch, int64_t flags_off, unsigned long long budget_ns, unsigned *status) {
    __syncthreads();
    if ((int)threadIdx.x < world) {
        __threadfence_system();
        unsigned *theirs = reinterpret_cast<unsigned *>(bufs[threadIdx.x] + flags_off) + rank;
        asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(theirs), "r"(epoch) : "memory");
        const unsigned *mine = reinterpret_cast<const unsigned *>(bufs[rank] + flags_off) + threadIdx.x;
        wait_peer_flag(mine, epoch, budget_ns, status, (int)threadIdx.x, rank);  // bounded: a dead rank must not hang the node
    }
    __syncthreads();
}

enum CombineOp { kSumU32 = 0, kMinF32 = 1, kMaxF32 = 2, kSumU64 = 3, kSumF64 = 4 };

// Combines `bytes` (multiple of 16) at `off` of every rank's buffer into scratch + soff.
template <int OP>
__device__ __forceinline__ void combine_region(unsigned char *const *bufs, int world, int64_t off, int64_t bytes, unsigned char *scratch, int64_t soff) {
    for (int64_t i = (int64_t)threadIdx.x * 16; i < bytes; i += (int64_t)kCombineThreads * 16) {
        uint4 acc = ld_peer_v4(bufs[0] + off + i);
        for (int p = 1; p < world; ++p) {
            const uint4 v = ld_peer_v4(bufs[p] + off + i);
            if constexpr (OP == kSumU32) {
                acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w;
            } else if constexpr (OP == kMinF32) {
                acc.x = __float_as_uint(fminf(__uint_as_float(acc.x), __uint_as_float(v.x))); acc.y = __float_as_uint(fminf(__uint_as_float(acc.y), __uint_as_float(v.y)));
                acc.z = __float_as_uint(fminf(__uint_as_float(acc.z), __uint_as_float(v.z))); acc.w = __float_as_uint(fminf(__uint_as_float(acc.w), __uint_as_float(v.w)));
            } else if constexpr (OP == kMaxF32) {
                acc.x = __float_as_uint(fmaxf(__uint_as_float(acc.x), __uint_as_float(v.x))); acc.y = __float_as_uint(fmaxf(__uint_as_float(acc.y), __uint_as_float(v.y)));
                acc.z = __float_as_uint(fmaxf(__uint_as_float(acc.z), __uint_as_float(v.z))); acc.w = __float_as_uint(fmaxf(__uint_as_float(acc.w), __uint_as_float(v.w)));
            } else if constexpr (OP == kSumU64) {
                const unsigned long long a0 = ((unsigned long long)acc.y << 32 | acc.x) + ((unsigned long long)v.y << 32 | v.x), a1 = ((unsigned long long)acc.w << 32 | acc.z) + ((unsigned long long)v.w << 32 | v.z);
                acc = make_uint4((unsigned)a0, (unsigned)(a0 >> 32), (unsigned)a1, (unsigned)(a1 >> 32));
            } else {
                const double a0 = __hiloint2double((int)acc.y, (int)acc.x) + __hiloint2double((int)v.y, (int)v.x), a1 = __hiloint2double((int)acc.w, (int)acc.z) + __hiloint2double((int)v.w, (int)v.z);
                acc = make_uint4((unsigned)__double2loint(a0), (unsigned)__double2hiint(a0), (unsigned)__double2loint(a1), (unsigned)__double2hiint(a1));
            }
        }
        *reinterpret_cast<uint4 *>(scratch + soff + i) = acc;
    }
}

__global__ void __launch_bounds__(kCombineThreads) peer_combine_kernel(unsigned char *const *__restrict__ bufs, int world, int rank, unsigned epoch, int which, unsigned char *__restrict__ scratch, unsigned long long budget_ns, unsigned *status) {
    const Layout L(1);
    const int64_t cells = 2 * kBins * 4;  // bytes of one per-slot cell array
    peer_rendezvous(bufs, world, rank, epoch, L.total, budget_ns, status);  // (1) everybody's statistics of this step are complete
    if (which == 0) {
        combine_region<kSumU64>(bufs, world, L.moments, 12 * 8, scratch, 0);  // fixed-point sums: integer addition, order-independent
        combine_region<kMaxF32>(bufs, world, L.odrange, 8 * 4, scratch, 128);
    } else if (which == 1) {
        combine_region<kSumU32>(bufs, world, L.hist1, cells, scratch, 0);
        combine_region<kSumU64>(bufs, world, L.counters, 8 * 8, scratch, 3 * cells);
    } else {
        combine_region<kSumU32>(bufs, world, L.hist2, cells, scratch, 0);
        combine_region<kMinF32>(bufs, world, L.vmin, cells, scratch, cells);
        combine_region<kMaxF32>(bufs, world, L.vmax, cells, scratch, 2 * cells);
        combine_region<kSumU64>(bufs, world, L.counters, 8 * 8, scratch, 3 * cells);
    }
    __threadfence();
    peer_rendezvous(bufs, world, rank, epoch, L.total + kPeerMaxWorld * 4, budget_ns, status);  // (3) everybody has read everybody
    unsigned char *own = bufs[rank];
    auto put = [&](int64_t off, int64_t bytes, int64_t soff) {
        for (int64_t i = (int64_t)threadIdx.x * 16; i < bytes; i += (int64_t)kCombineThreads * 16) *reinterpret_cast<uint4 *>(own + off + i) = *reinterpret_cast<const uint4 *>(scratch + soff + i);
    };
    if (which == 0) {
        put(L.moments, 12 * 8, 0);
        put(L.odrange, 8 * 4, 128);
    } else if (which == 1) {
        put(L.hist1, cells, 0);
        put(L.counters, 8 * 8, 3 * cells);
    } else {
        put(L.hist2, cells, 0);
        put(L.vmin, cells, cells);
        put(L.vmax, cells, 2 * cells);
        put(L.counters, 8 * 8, 3 * cells);
    }
}

// fit_transform: the pooled fit's moments are the SUM of the per-image moments the transform needs anyway (fixed-point
// integers: the sum of the slots is bit for bit what a pooled moments pass over the same images accumulates), and its
// ranges the MAX of theirs -- one read of the batch instead of two.  One CTA of 12 warps: warp w < 11 sums moment word w
// (10 sums + the pixel count) of the slots [slot0, slot0 + count) of `src`, warp 11 takes the maximum of the 8 ODRANGE
// words; the result goes to `dst_slot` of `dst` (same or another workspace, e.g. a peer-mapped one-slot buffer).
constexpr int kPoolThreads = 384;
__global__ void __launch_bounds__(kPoolThreads) pool_moments_kernel(const void *src_base, int64_t src_slots, int64_t slot0, int64_t count, void *dst_base, int64_t dst_slots, int64_t dst_slot) {
    const Ws src(const_cast<void *>(src_base), src_slots);
    Ws dst(dst_base, dst_slots);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (warp < 11) {
        long long v = 0;
        for (int64_t i = lane; i < count; i += 32) v += src.moments[(slot0 + i) * 12 + warp];
        v = warp_sum_ll(v);
        if (lane == 0) dst.moments[dst_slot * 12 + warp] = v;
    } else {
        const int word = lane & 7;
        float v = -INFINITY;
        for (int64_t i = lane >> 3; i < count; i += 4) v = fmaxf(v, src.odrange[(slot0 + i) * 8 + word]);
        v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, 8));
        v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, 16));
        if (lane < 8) dst.odrange[dst_slot * 8 + word] = v;
    }
}

__global__ void init_kernel(void *ws_base, int64_t slots) {
    Ws ws(ws_base, slots);
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int64_t per = 2 * kBins;
    if (i < slots * per) {
        ws.hist1[i] = 0u;
        ws.hist2[i] = 0u;
        ws.vmin[i] = INFINITY;
        ws.vmax[i] = -INFINITY;
    }
    if (i < slots * 12) ws.moments[i] = 0;
    if (i < slots * 8) { ws.odrange[i] = -INFINITY; ws.fit[i] = 0.0f; ws.counters[i] = 0ull; }
    if (i < slots * 4) ws.status[i] = 0;
    if (i < slots) {
        SlotState z = {};
        ws.state[i] = z;
    }
}

static int g_ctas_per_sm = 8;  // per-image kernels (sample, reconstruction): CTAs per SM the grid is sized for.  The reconstruction of
                               // 64 x 1024^2 float32 takes 259 us with 3 or 8 and 283 us with 4 (640 CTAs = 4.3 per SM: uneven last wave)
static int g_split = 3;          // chains (streams) a large batch is split into
static int g_phase_kernels = 0;  // development: sx_macenko_transform through the phase-level API (one launch per step)
static int g_helper_streams = 1; // per-image kernels of a multi-chain transform on high-priority helper streams (see ChainHelper)

static PassGeom make_geom(int64_t n, int64_t hw, int kpix, int64_t groups_override = -1) {
    PassGeom g;
    g.n_img = n;
    g.hw = hw;
    const int64_t groups = groups_override >= 0 ? groups_override : hw / kpix;
    int64_t want = ((int64_t)sm_count() * g_ctas_per_sm + n - 1) / (n > 0 ? n : 1);
    const int64_t most = (groups + kThreads - 1) / kThreads;
    if (want > most) want = most;
    if (want < 1) want = 1;
    g.cpi = (int)want;
    return g;
}

template <typename T, bool VEC>
static RowGeom make_row_geom(int64_t n, int64_t hw, int64_t slot0, int pooled) {
    RowGeom g;
    g.n_img = n;
    g.hw = hw;
    g.slot0 = slot0;
    g.pooled = pooled;
    g.rows_per_img = (int)((hw / Pix<T, VEC>::kPix + kThreads - 1) / kThreads);
    g.total_rows = n * g.rows_per_img;
    return g;
}

// Grid of a pipeline kernel: every CTA resident at once (occupancy x SMs), never more than rows.
template <typename K>
static unsigned pipeline_grid(K kernel, int64_t total_rows) {
    int per_sm = 0;
    prefer_l1(kernel, kThreads);  // the pipeline kernels stream: smallest carve-out that holds the resident CTAs
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kernel, kThreads, 0) != cudaSuccess || per_sm < 1) per_sm = 1;
    const int64_t resident = (int64_t)per_sm * sm_count();
    return (unsigned)(total_rows < resident ? (total_rows > 0 ? total_rows : 1) : resident);
}

template <typename T>
static bool vec_ok(const void *a, const void *b, int64_t hw) {
    return hw % (16 / (int64_t)sizeof(T)) == 0 && aligned16(a) && (b == nullptr || aligned16(b));
}

}  // namespace macenko
}  // namespace sx

using namespace sx;
using namespace sx::macenko;

#define SX_DISPATCH_TV_ONE(TYPE, vec, ...)                                      \
    {                                                                          \
        using T = TYPE;                                                        \
        if (vec) { constexpr bool VEC = true; __VA_ARGS__; }                   \
        else { constexpr bool VEC = false; __VA_ARGS__; }                      \
    }
#define SX_DISPATCH_TV(dtype, vec, ...)                                        \
    do {                                                                       \
        if ((dtype) == SX_F32) SX_DISPATCH_TV_ONE(float, vec, __VA_ARGS__)     \
        else if ((dtype) == SX_F16) SX_DISPATCH_TV_ONE(__half, vec, __VA_ARGS__) \
        else if ((dtype) == SX_BF16) SX_DISPATCH_TV_ONE(__nv_bfloat16, vec, __VA_ARGS__) \
        else SX_DISPATCH_TV_ONE(uint8_t, vec, __VA_ARGS__)                     \
    } while (0)

static bool images_vec_ok(const void *images, const void *out, int dtype, int64_t hw) {
    const bool in_ok = dtype == SX_F32 ? vec_ok<float>(images, nullptr, hw) : (dtype == SX_U8 ? vec_ok<uint8_t>(images, nullptr, hw) : vec_ok<__half>(images, nullptr, hw));
    // output plane offsets are multiples of hw elements of the output type (>= 1 byte), so an
    // aligned base and hw % kPix == 0 keep every 128-bit store aligned
    return in_ok && (out == nullptr || aligned16(out));
}

extern "C" int sx_macenko_apply(const void *images, int dtype, int64_t n, int64_t h, int64_t w, int64_t slot0, const float *he_ref, const float *maxc_ref, void *out, int out_dtype, float out_scale, void *workspace, int64_t slots, sx_stream_t stream_);

// Development (sx_macenko_trace, inert unless SX_ENABLE_TUNING=1): completion time of every kernel of the transform
// chains, from timing events recorded behind each launch -- the only timeline tool on a box without nsys.
struct TraceRec { const char *what; int chain; cudaEvent_t ev; };
static std::vector<TraceRec> g_trace;
static cudaEvent_t g_trace_base = nullptr;
static int g_trace_on = 0;
static void trace_mark(const char *what, int chain, cudaStream_t s) {
    if (!g_trace_on) return;
    cudaEvent_t e = nullptr;
    if (cudaEventCreate(&e) != cudaSuccess || cudaEventRecord(e, s) != cudaSuccess) return;
    g_trace.push_back({what, chain, e});
}

// Helper stream of one chain (library-owned, HIGH priority) and the events that tie it to the chain's stream.
// The streaming kernels of a chain are single-wave grids sized to the resident capacity of the device: while one runs
// there is no free CTA slot, and when it drains the hardware hands the slots to whatever is pending -- in launch order
// among streams of equal priority.  With the per-image kernels on the chains' own streams a chain's `mid` / `select`
// kernels therefore queue behind the streaming kernels the other chains already have pending, and the chains drift into
// lockstep (the timeline of tools/trace_mk.py: all three `select` kernels of a stage complete within a few us of each
// other, with nothing else to run beside them): the transform took the same 780-795 us with 2, 3 or 4 chains, against
// 721 us for its four streaming kernels alone.  On a high-priority stream a chain's small kernels take the first slots
// that drain, run beside another chain's streaming kernel, and that chain's next streaming kernel is ready when its turn
// comes.  Same kernels on the same data, so no bit changes (tools/probe_helper.py: torch.equal over all settings);
// measured on 64 x 1024^2 float32 781 -> 764 us (repeat: 785 -> 766), 32 x 2048^2 uint8 -> float32 1096 -> 1069 us
// (1094 -> 1062); uint8 -> uint8 and fit_transform unchanged within the 1-2 % run-to-run spread.
struct ChainHelper {
    cudaStream_t stream = nullptr;
    cudaEvent_t ev[6] = {};
};

// One chain of the per-image pipeline for images [0, n) -> slots [slot0, slot0 + n) on `stream`.
// with_moments = false: the slots already hold the images' moments (sx_macenko_fit_transform).
// hp != nullptr: the per-image kernels go to hp->stream, ordered against `stream` by hp->ev in both directions.
static int run_pipeline(const void *images, int dtype, int64_t n, int64_t h, int64_t w, int64_t slot0, const float *he_ref, const float *maxc_ref, void *out, int out_dtype, float out_scale, void *workspace, int64_t slots, cudaStream_t stream, bool with_moments = true, cudaEvent_t before_apply = nullptr, const ChainHelper *hp = nullptr, int chain = 0) {
    const int64_t hw = h * w;
    const bool vec = images_vec_ok(images, nullptr, dtype, hw);
    cudaStream_t small = hp ? hp->stream : stream;
    cudaError_t herr = cudaSuccess;
    int hop_no = 0;
    // what `to` enqueues next runs after everything enqueued on `from` so far
    auto hop = [&](cudaStream_t from, cudaStream_t to) {
        if (!hp) return;
        cudaEvent_t e = hp->ev[hop_no++];
        if (herr == cudaSuccess) herr = cudaEventRecord(e, from);
        if (herr == cudaSuccess) herr = cudaStreamWaitEvent(to, e, 0);
    };
    SX_DISPATCH_TV(dtype, vec, {
        const T *p = static_cast<const T *>(images);
        const RowGeom g = make_row_geom<T, VEC>(n, hw, slot0, 0);
        if (with_moments) t_moments_kernel<T, VEC><<<pipeline_grid(t_moments_kernel<T, VEC>, g.total_rows), kThreads, 0, stream>>>(p, g, workspace, slots);
        trace_mark("moments", chain, stream);
        hop(stream, small);
        mid_kernel<T, VEC, SX_STAGE_ANGLE><<<(unsigned)(n * kMidParts), kThreads, 0, small>>>(p, hw, slot0, workspace, slots);
        trace_mark("mid_angle", chain, small);
        hop(small, stream);
        t_resolve_kernel<T, VEC, SX_STAGE_ANGLE><<<pipeline_grid(t_resolve_kernel<T, VEC, SX_STAGE_ANGLE>, g.total_rows), kThreads, 0, stream>>>(p, g, workspace, slots);
        trace_mark("resolve_angle", chain, stream);
        hop(stream, small);
        select_recover_kernel<T, VEC, SX_STAGE_ANGLE><<<(unsigned)n, kThreads, 0, small>>>(p, n, hw, 0, slot0, workspace, slots);
        trace_mark("select_angle", chain, small);
        mid_kernel<T, VEC, SX_STAGE_CONC><<<(unsigned)(n * kMidParts), kThreads, 0, small>>>(p, hw, slot0, workspace, slots);
        trace_mark("mid_conc", chain, small);
        hop(small, stream);
        t_resolve_kernel<T, VEC, SX_STAGE_CONC><<<pipeline_grid(t_resolve_kernel<T, VEC, SX_STAGE_CONC>, g.total_rows), kThreads, 0, stream>>>(p, g, workspace, slots);
        trace_mark("resolve_conc", chain, stream);
        hop(stream, small);
        select_recover_kernel<T, VEC, SX_STAGE_CONC><<<(unsigned)n, kThreads, 0, small>>>(p, n, hw, 0, slot0, workspace, slots);
        trace_mark("select_conc", chain, small);
        hop(small, stream);
    });
    note_launch(with_moments ? 6 : 5);
    SX_LAUNCHED("macenko::transform pipeline");
    if (herr != cudaSuccess) return sx::fail(SX_ERR_CUDA, "macenko::transform pipeline: helper-stream event failed: %s", cudaGetErrorString(herr));
    if (before_apply) SX_CUDA(cudaStreamWaitEvent(stream, before_apply, 0));  // he_ref / maxc_ref come from another stream
    const int rc = sx_macenko_apply(images, dtype, n, h, w, slot0, he_ref, maxc_ref, out, out_dtype, out_scale, workspace, slots, stream);
    trace_mark("apply", chain, stream);
    return rc;
}

// Side stream of the calling thread's current device (created on first use, never destroyed: the
// library owns no device memory, only this stream and two events per device).
constexpr int kMaxChains = 4;
struct SideStream {
    cudaStream_t stream[kMaxChains - 1] = {};
    cudaEvent_t fork = nullptr, join[kMaxChains - 1] = {};
    std::mutex mu;
    // fit_transform: the pooled fit runs beside the transform's per-image chains (they do not need it before `apply`)
    cudaStream_t fit_stream = nullptr;
    cudaEvent_t fit_fork = nullptr, fit_done = nullptr;
    std::mutex fit_mu;
    ChainHelper helper[kMaxChains];  // per chain: high-priority stream of its per-image kernels (see ChainHelper)
};
static SideStream *side_stream() {
    static SideStream table[64];
    static std::mutex init_mu;
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return nullptr;
    std::lock_guard<std::mutex> lock(init_mu);
    SideStream &s = table[dev];
    if (s.fork == nullptr) {
        for (int i = 0; i < kMaxChains - 1; ++i) {
            if (cudaStreamCreateWithFlags(&s.stream[i], cudaStreamNonBlocking) != cudaSuccess) return nullptr;
            if (cudaEventCreateWithFlags(&s.join[i], cudaEventDisableTiming) != cudaSuccess) return nullptr;
        }
        int lo_prio = 0, hi_prio = 0;  // the fit is a chain of short dependent kernels: let them jump the queue of streaming CTAs
        if (cudaDeviceGetStreamPriorityRange(&lo_prio, &hi_prio) != cudaSuccess) return nullptr;
        if (cudaStreamCreateWithPriority(&s.fit_stream, cudaStreamNonBlocking, hi_prio) != cudaSuccess) return nullptr;
        for (int c = 0; c < kMaxChains; ++c) {
            if (cudaStreamCreateWithPriority(&s.helper[c].stream, cudaStreamNonBlocking, hi_prio) != cudaSuccess) return nullptr;
            for (cudaEvent_t &e : s.helper[c].ev)
                if (cudaEventCreateWithFlags(&e, cudaEventDisableTiming) != cudaSuccess) return nullptr;
        }
        if (cudaEventCreateWithFlags(&s.fit_fork, cudaEventDisableTiming) != cudaSuccess) return nullptr;
        if (cudaEventCreateWithFlags(&s.fit_done, cudaEventDisableTiming) != cudaSuccess) return nullptr;
        if (cudaEventCreateWithFlags(&s.fork, cudaEventDisableTiming) != cudaSuccess) return nullptr;
    }
    return &s;
}

extern "C" {

// Development hook: process-global, not thread-safe, inert unless SX_ENABLE_TUNING=1.
int sx_macenko_set_tuning(int ctas_per_sm, int64_t phase_kernels) {
    if (!tuning_enabled()) return sx::fail(SX_ERR_UNSUPPORTED, "tuning hooks are disabled (set SX_ENABLE_TUNING=1 before loading the library)");
    if (ctas_per_sm > 0) g_ctas_per_sm = ctas_per_sm;
    if (phase_kernels >= 0) {  // bit 0: phase-level chain; bit 1: force sample-bracket misses (recovery tests); bit 2: no helper streams; bits 4..: number of chains (0: default)
        g_phase_kernels = (phase_kernels & 1) != 0;
        g_helper_streams = ((phase_kernels >> 2) & 1) == 0;
        const int force = (int)((phase_kernels >> 1) & 1);
        SX_CUDA(cudaMemcpyToSymbol(g_force_miss, &force, sizeof(int)));
        const int chains = (int)(phase_kernels >> 4);
        g_split = chains > 0 ? (chains < kMaxChains ? chains : kMaxChains) : 3;
    }
    return SX_OK;
}

// Development hook: enable != 0 arms the trace for the next multi-chain transform; enable == 0 synchronises the device and
// writes one line per kernel ("chain kernel completion-time-in-us", relative to the start of the call) into buf.
int sx_macenko_trace(int enable, char *buf, int64_t cap) {
    if (!tuning_enabled()) return sx::fail(SX_ERR_UNSUPPORTED, "tuning hooks are disabled (set SX_ENABLE_TUNING=1 before loading the library)");
    if (enable) {
        for (TraceRec &r : g_trace) cudaEventDestroy(r.ev);
        g_trace.clear();
        if (g_trace_base) cudaEventDestroy(g_trace_base);
        g_trace_base = nullptr;
        g_trace_on = 1;
        return SX_OK;
    }
    g_trace_on = 0;
    SX_CUDA(cudaDeviceSynchronize());
    int64_t used = 0;
    if (buf && cap > 0) buf[0] = 0;
    for (TraceRec &r : g_trace) {
        float ms = -1.0f;
        if (g_trace_base) cudaEventElapsedTime(&ms, g_trace_base, r.ev);
        if (buf && used < cap - 64) used += snprintf(buf + used, (size_t)(cap - used), "%d %s %.1f\n", r.chain, r.what, ms * 1e3f);
        cudaEventDestroy(r.ev);
    }
    g_trace.clear();
    if (g_trace_base) cudaEventDestroy(g_trace_base);
    g_trace_base = nullptr;
    return SX_OK;
}

int64_t sx_macenko_workspace_bytes(int64_t slots) { return slots > 0 ? Layout(slots).total : 0; }

int sx_macenko_region(int64_t slots, int region, int64_t *offset, int64_t *bytes) {
    SX_REQUIRE(slots > 0 && offset && bytes, "bad arguments");
    Layout L(slots);
    switch (region) {
        case SX_REGION_MOMENTS: *offset = L.moments; *bytes = slots * 12 * 8; break;
        case SX_REGION_ODRANGE: *offset = L.odrange; *bytes = slots * 8 * 4; break;
        case SX_REGION_HIST1: *offset = L.hist1; *bytes = slots * 2 * kBins * 4; break;
        case SX_REGION_HIST2: *offset = L.hist2; *bytes = slots * 2 * kBins * 4; break;
        case SX_REGION_VMIN: *offset = L.vmin; *bytes = slots * 2 * kBins * 4; break;
        case SX_REGION_VMAX: *offset = L.vmax; *bytes = slots * 2 * kBins * 4; break;
        case SX_REGION_FIT: *offset = L.fit; *bytes = slots * 8 * 4; break;
        case SX_REGION_COUNTERS: *offset = L.counters; *bytes = slots * 8 * 8; break;
        case SX_REGION_STATUS: *offset = L.status; *bytes = slots * 4 * 4; break;
        default: return sx::fail(SX_ERR_INVALID, "unknown region %d", region);
    }
    return SX_OK;
}

int64_t sx_macenko_peer_buffer_bytes(void) { return Layout(1).total + 2 * kPeerMaxWorld * 4; }
int64_t sx_macenko_peer_scratch_bytes(void) { return 3 * 2 * kBins * 4 + 256; }

int sx_macenko_peer_combine(const void *peer_buffers_dev, int world, int rank, uint32_t epoch, int which, void *scratch, sx_stream_t stream) {
    SX_REQUIRE(peer_buffers_dev && scratch, "NULL argument");
    SX_REQUIRE(world >= 1 && world <= kPeerMaxWorld && rank >= 0 && rank < world, "bad rank/world (%d, %d)", rank, world);
    SX_REQUIRE(epoch != 0 && which >= 0 && which <= 2, "bad epoch/which (%u, %d)", epoch, which);
    if (int rc = peer_status_check("sx_macenko_peer_combine")) return rc;
    peer_combine_kernel<<<1, kCombineThreads, 0, static_cast<cudaStream_t>(stream)>>>(static_cast<unsigned char *const *>(peer_buffers_dev), world, rank, epoch, which, static_cast<unsigned char *>(scratch), peer_timeout_ns(), peer_status_device_ptr());
    SX_LAUNCHED("macenko::peer_combine_kernel");
    return SX_OK;
}

int sx_macenko_begin(void *workspace, int64_t slots, sx_stream_t stream) {
    SX_REQUIRE(workspace && slots > 0, "bad workspace");
    const int64_t items = slots * 2 * kBins;
    init_kernel<<<(unsigned)((items + 255) / 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(workspace, slots);
    SX_LAUNCHED("macenko::init_kernel");
    return SX_OK;
}

static int check_slots(int64_t n, int pooled, int64_t slot0, int64_t slots) {
    SX_REQUIRE(slots > 0, "slots must be > 0");
    if (pooled) return SX_OK;
    SX_REQUIRE(slot0 >= 0 && slot0 + n <= slots, "slot range [%lld, %lld) outside [0, %lld)", (long long)slot0, (long long)(slot0 + n), (long long)slots);
    return SX_OK;
}

static int check_range(int64_t slot0, int64_t count, int64_t slots) {
    SX_REQUIRE(slots > 0 && slot0 >= 0 && count >= 0 && slot0 + count <= slots, "slot range [%lld, %lld) outside [0, %lld)", (long long)slot0, (long long)(slot0 + count), (long long)slots);
    return SX_OK;
}

int sx_macenko_moments(const void *images, int dtype, int64_t n, int64_t h, int64_t w, int pooled, int64_t slot0, void *workspace, int64_t slots, sx_stream_t stream_) {
    if (int rc = check_images(images, dtype, n, h, w)) return rc;
    if (int rc = check_slots(n, pooled, slot0, slots)) return rc;
    SX_REQUIRE(workspace, "workspace is NULL");
    const int64_t hw = h * w;
    if (n == 0 || hw == 0) return SX_OK;
    cudaStream_t stream = static_cast<cudaStream_t>(stream_);
    const bool vec = images_vec_ok(images, nullptr, dtype, hw);
    SX_DISPATCH_TV(dtype, vec, {
        const RowGeom g = make_row_geom<T, VEC>(n, hw, pooled ? 0 : slot0, pooled);
        t_moments_kernel<T, VEC><<<pipeline_grid(t_moments_kernel<T, VEC>, g.total_rows), kThreads, 0, stream>>>(static_cast<const T *>(images), g, workspace, slots);
    });
    SX_LAUNCHED("macenko::moments_kernel");
    return SX_OK;
}

int sx_macenko_basis(void *workspace, int64_t slots, int64_t slot0, int64_t count, int allow_fallback, sx_stream_t stream) {
    SX_REQUIRE(workspace, "workspace is NULL");
    if (int rc = check_range(slot0, count, slots)) return rc;
    if (count == 0) return SX_OK;
    basis_kernel<<<(unsigned)((count + 63) / 64), 64, 0, static_cast<cudaStream_t>(stream)>>>(workspace, slots, slot0, count, allow_fallback);
    SX_LAUNCHED("macenko::basis_kernel");
    return SX_OK;
}

int sx_macenko_moments_fallback(const void *images, int dtype, int64_t n, int64_t h, int64_t w, int64_t slot0, void *workspace, int64_t slots, sx_stream_t stream_) {
    if (int rc = check_images(images, dtype, n, h, w)) return rc;
    if (int rc = check_slots(n, 0, slot0, slots)) return rc;
    SX_REQUIRE(workspace, "workspace is NULL");
    const int64_t hw = h * w;
    if (n == 0 || hw == 0) return SX_OK;
    cudaStream_t stream = static_cast<cudaStream_t>(stream_);
    const bool vec = images_vec_ok(images, nullptr, dtype, hw);
    SX_DISPATCH_TV(dtype, vec, {
        fallback_kernel<T, VEC><<<(unsigned)n, kThreads, 0, stream>>>(static_cast<const T *>(images), hw, slot0, workspace, slots);
    });
    SX_LAUNCHED("macenko::fallback_kernel");
    return SX_OK;
}

int sx_macenko_hist(const void *images, int dtype, int64_t n, int64_t h, int64_t w, int pooled, int64_t slot0, int stage, int level, void *workspace, int64_t slots, sx_stream_t stream_) {
    if (int rc = check_images(images, dtype, n, h, w)) return rc;
    if (int rc = check_slots(n, pooled, slot0, slots)) return rc;
    SX_REQUIRE(workspace, "workspace is NULL");
    SX_REQUIRE((stage == SX_STAGE_ANGLE || stage == SX_STAGE_CONC) && (level >= 0 && level <= 2), "bad stage/level (%d, %d)", stage, level);
    const int64_t hw = h * w;
    if (n == 0 || hw == 0) return SX_OK;
    cudaStream_t stream = static_cast<cudaStream_t>(stream_);
    const bool vec = images_vec_ok(images, nullptr, dtype, hw);
    SX_DISPATCH_TV(dtype, vec, {
        const T *p = static_cast<const T *>(images);
        if (level != 1) {
            const int64_t groups = hw / Pix<T, VEC>::kPix;
            const int64_t max_groups = level == 0 ? (int64_t)kSampleGroups : (int64_t)1 << 62;  // level 2: every group (exact coarse histogram)
            const int64_t stride = groups / max_groups > 1 ? groups / max_groups : 1;
            PassGeom g = make_geom(n, hw, Pix<T, VEC>::kPix, groups / stride);
            const unsigned grid = (unsigned)(n * g.cpi);
            if (stage == SX_STAGE_ANGLE) sample_kernel<T, VEC, SX_STAGE_ANGLE><<<grid, kThreads, 0, stream>>>(p, g, pooled, slot0, workspace, slots, max_groups);
            else sample_kernel<T, VEC, SX_STAGE_CONC><<<grid, kThreads, 0, stream>>>(p, g, pooled, slot0, workspace, slots, max_groups);
        } else {
            const RowGeom g = make_row_geom<T, VEC>(n, hw, pooled ? 0 : slot0, pooled);
            if (stage == SX_STAGE_ANGLE) t_resolve_kernel<T, VEC, SX_STAGE_ANGLE><<<pipeline_grid(t_resolve_kernel<T, VEC, SX_STAGE_ANGLE>, g.total_rows), kThreads, 0, stream>>>(p, g, workspace, slots);
            else t_resolve_kernel<T, VEC, SX_STAGE_CONC><<<pipeline_grid(t_resolve_kernel<T, VEC, SX_STAGE_CONC>, g.total_rows), kThreads, 0, stream>>>(p, g, workspace, slots);
        }
    });
    SX_LAUNCHED("macenko::hist kernels");
    return SX_OK;
}

int sx_macenko_select(void *workspace, int64_t slots, int64_t slot0, int64_t count, int stage, int level, sx_stream_t stream_) {
    SX_REQUIRE(workspace, "workspace is NULL");
    if (int rc = check_range(slot0, count, slots)) return rc;
    SX_REQUIRE((stage == SX_STAGE_ANGLE || stage == SX_STAGE_CONC) && (level == 0 || level == 1), "bad stage/level (%d, %d)", stage, level);
    if (count == 0) return SX_OK;
    cudaStream_t stream = static_cast<cudaStream_t>(stream_);
    if (level == 0) bracket_kernel<<<(unsigned)count, kThreads, 0, stream>>>(workspace, slots, slot0, stage);
    else select_kernel<<<(unsigned)count, kThreads, 0, stream>>>(workspace, slots, slot0, stage);
    SX_LAUNCHED("macenko::select kernels");
    return SX_OK;
}

int sx_macenko_apply(const void *images, int dtype, int64_t n, int64_t h, int64_t w, int64_t slot0, const float *he_ref, const float *maxc_ref, void *out, int out_dtype, float out_scale, void *workspace, int64_t slots, sx_stream_t stream_) {
    if (int rc = check_images(images, dtype, n, h, w)) return rc;
    if (int rc = check_slots(n, 0, slot0, slots)) return rc;
    if (n == 0 || h * w == 0) return SX_OK;
    SX_REQUIRE(workspace && he_ref && maxc_ref && out, "NULL argument");
    const bool half_in = dtype == SX_F16 || dtype == SX_BF16;
    SX_REQUIRE(half_in ? out_dtype == dtype : (out_dtype == SX_F32 || (out_dtype == SX_U8 && dtype == SX_U8)), "output dtype %d does not go with input dtype %d (uint8 -> uint8 | float32, float32 -> float32, float16 / bfloat16 -> the same)", out_dtype, dtype);
    const bool unit = out_scale != 1.0f;
    SX_REQUIRE(!unit || fabsf(out_scale - 1.0f / 255.0f) < 1e-9f, "out_scale must be 1 or 1/255");
    const int64_t hw = h * w;
    cudaStream_t stream = static_cast<cudaStream_t>(stream_);
    const bool vec = images_vec_ok(images, out, dtype, hw);
    if (dtype == SX_U8 && out_dtype == SX_F32 && hw % 4 == 0 && (reinterpret_cast<uintptr_t>(images) & 3u) == 0 && aligned16(out)) {
        PassGeom g = make_geom(n, hw, 4);
        const unsigned grid = (unsigned)(n * g.cpi);
        const uint8_t *p = static_cast<const uint8_t *>(images);
        prefer_l1(apply_u8_f32_kernel<2>, kThreads);
        prefer_l1(apply_u8_f32_kernel<1>, kThreads);
        if (unit) apply_u8_f32_kernel<2><<<grid, kThreads, 0, stream>>>(p, static_cast<float *>(out), g, slot0, he_ref, maxc_ref, workspace, slots);
        else apply_u8_f32_kernel<1><<<grid, kThreads, 0, stream>>>(p, static_cast<float *>(out), g, slot0, he_ref, maxc_ref, workspace, slots);
        SX_LAUNCHED("macenko::apply_u8_f32_kernel");
        return SX_OK;
    }
    SX_DISPATCH_TV(dtype, vec, {
        PassGeom g = make_geom(n, hw, Pix<T, VEC>::kPix);
        const unsigned grid = (unsigned)(n * g.cpi);
        const T *p = static_cast<const T *>(images);
        prefer_l1(apply_kernel<T, VEC, 0>, kThreads);
        prefer_l1(apply_kernel<T, VEC, 1>, kThreads);
        prefer_l1(apply_kernel<T, VEC, 2>, kThreads);
        if (out_dtype == SX_U8) apply_kernel<T, VEC, 0><<<grid, kThreads, 0, stream>>>(p, out, g, slot0, he_ref, maxc_ref, workspace, slots);
        else if (!unit) apply_kernel<T, VEC, 1><<<grid, kThreads, 0, stream>>>(p, out, g, slot0, he_ref, maxc_ref, workspace, slots);
        else apply_kernel<T, VEC, 2><<<grid, kThreads, 0, stream>>>(p, out, g, slot0, he_ref, maxc_ref, workspace, slots);
    });
    SX_LAUNCHED("macenko::apply_kernel");
    return SX_OK;
}

// The per-image pipeline of images [0, n) -> slots [slot0, slot0 + n) (see "per-image transform pipeline" above).  Large
// batches run as up to kMaxChains part-batch chains, all but the first on side streams: the small per-image kernels of one
// chain (~25 us per stage of dependent global round trips on a few SMs) overlap the streaming kernels of the other.  The
// caller's stream forks into the side streams and joins them again, so the call keeps its stream-ordered,
// host-asynchronous contract (and can be captured into a CUDA graph).
static int transform_chains(const void *images, int dtype, int64_t n, int64_t h, int64_t w, int64_t slot0, const float *he_ref, const float *maxc_ref, void *out, int out_dtype, float out_scale, void *workspace, int64_t slots, cudaStream_t stream, bool with_moments, cudaEvent_t before_apply = nullptr) {
    const int64_t hw = h * w;
    const int64_t in_bytes = (int64_t)dtype_bytes(dtype) * 3 * hw, out_bytes = (int64_t)dtype_bytes(out_dtype) * 3 * hw;
    int chains = g_split;
    if (n * in_bytes < ((int64_t)64 << 20)) chains = 1;
    while (chains > 1 && n / chains < 4) --chains;
    if (chains <= 1) return run_pipeline(images, dtype, n, h, w, slot0, he_ref, maxc_ref, out, out_dtype, out_scale, workspace, slots, stream, with_moments, before_apply);
    SideStream *side = side_stream();
    SX_REQUIRE(side != nullptr, "could not create the side streams");
    std::lock_guard<std::mutex> lock(side->mu);
    SX_CUDA(cudaEventRecord(side->fork, stream));
    if (g_trace_on && g_trace_base == nullptr && cudaEventCreate(&g_trace_base) == cudaSuccess) cudaEventRecord(g_trace_base, stream);
    int rc = SX_OK;
    for (int c = 0; c < chains; ++c) {  // chain 0 on the caller's stream, chain c > 0 on side stream c - 1
        const int64_t i0 = n * c / chains, i1 = n * (c + 1) / chains;
        cudaStream_t cs = c == 0 ? stream : side->stream[c - 1];
        if (c > 0) SX_CUDA(cudaStreamWaitEvent(cs, side->fork, 0));
        const int r = run_pipeline(static_cast<const char *>(images) + i0 * in_bytes, dtype, i1 - i0, h, w, slot0 + i0, he_ref, maxc_ref, static_cast<char *>(out) + i0 * out_bytes, out_dtype, out_scale, workspace, slots, cs, with_moments, before_apply,
                                   g_helper_streams ? &side->helper[c] : nullptr, c);
        if (r && !rc) rc = r;
        if (c > 0) {
            SX_CUDA(cudaEventRecord(side->join[c - 1], cs));
            SX_CUDA(cudaStreamWaitEvent(stream, side->join[c - 1], 0));
        }
    }
    return rc;
}

// What sx_macenko_apply accepts, checked before the first launch of a whole-call entry point.
static int check_output(int dtype, const void *out, int out_dtype, float out_scale) {
    SX_REQUIRE(out, "NULL output");
    const bool half_in = dtype == SX_F16 || dtype == SX_BF16;
    SX_REQUIRE(half_in ? out_dtype == dtype : (out_dtype == SX_F32 || (out_dtype == SX_U8 && dtype == SX_U8)), "output dtype %d does not go with input dtype %d (uint8 -> uint8 | float32, float32 -> float32, float16 / bfloat16 -> the same)", out_dtype, dtype);
    SX_REQUIRE(out_scale == 1.0f || fabsf(out_scale - 1.0f / 255.0f) < 1e-9f, "out_scale must be 1 or 1/255");
    return SX_OK;
}

int sx_macenko_transform(const void *images, int dtype, int64_t n, int64_t h, int64_t w, const float *he_ref, const float *maxc_ref, void *out, int out_dtype, float out_scale, void *workspace, int64_t workspace_bytes, sx_stream_t s) {
    if (int rc = check_images(images, dtype, n, h, w)) return rc;
    if (n == 0 || h * w == 0) return SX_OK;
    SX_REQUIRE(workspace && workspace_bytes >= sx_macenko_workspace_bytes(n), "workspace too small (%lld < %lld)", (long long)workspace_bytes, (long long)sx_macenko_workspace_bytes(n));
    SX_REQUIRE(he_ref && maxc_ref && out, "NULL argument");
    cudaStream_t stream = static_cast<cudaStream_t>(s);
    int rc;
    if ((rc = sx_macenko_begin(workspace, n, s))) return rc;
    if (g_phase_kernels) {  // development: the phase-level API chained on one stream (13 launches)
        if ((rc = sx_macenko_moments(images, dtype, n, h, w, 0, 0, workspace, n, s))) return rc;
        if ((rc = sx_macenko_basis(workspace, n, 0, n, 1, s))) return rc;
        if ((rc = sx_macenko_moments_fallback(images, dtype, n, h, w, 0, workspace, n, s))) return rc;
        for (int stage = 0; stage < 2; ++stage)
            for (int level = 0; level < 2; ++level) {
                if ((rc = sx_macenko_hist(images, dtype, n, h, w, 0, 0, stage, level, workspace, n, s))) return rc;
                if ((rc = sx_macenko_select(workspace, n, 0, n, stage, level, s))) return rc;
            }
        return sx_macenko_apply(images, dtype, n, h, w, 0, he_ref, maxc_ref, out, out_dtype, out_scale, workspace, n, s);
    }
    return transform_chains(images, dtype, n, h, w, 0, he_ref, maxc_ref, out, out_dtype, out_scale, workspace, n, stream, true);
}

// The sharded pooled fit of one NVLink node as ONE library call: the phase sequence of the header's
// "Phase order" with sx_macenko_peer_combine after moments and after every hist pass (five exchanges, epochs
// first_epoch .. first_epoch + 4).  The same kernels as the phase-level calls; what it removes is the host time between
// ~25 separate calls (each ~8-10 us of interpreter + ctypes), which at 16-64 images per rank is a third of the fit.
// `own_buffer` = this rank's peer-mapped buffer (the one-slot workspace lives at its start); `exact` != 0 replaces the
// sample passes by the exact coarse pass (level 2) -- the repeat after a missed bracket.  n may be 0 (a rank without
// reference images still takes part in every exchange).  he / maxc may be NULL (read the FIT region instead).
// `per_image` (optional): a workspace of n slots that already holds the per-image moments of `images`
// (sx_macenko_fit_transform_peers) -- slot 0 of own_buffer then gets their sum instead of a second moments pass.
static int fit_peers_impl(const void *images, int dtype, int64_t n, int64_t h, int64_t w, const void *peer_buffers_dev, void *own_buffer, int world, int rank, uint32_t first_epoch, int exact,
                          void *scratch, float *he, float *maxc, sx_stream_t s, const void *per_image) {
    if (int rc = check_images(images, dtype, n, h, w)) return rc;
    SX_REQUIRE(peer_buffers_dev && own_buffer && scratch, "NULL argument");
    SX_REQUIRE(first_epoch != 0 && first_epoch + 4 >= first_epoch, "bad first epoch %u", first_epoch);
    const bool have = n > 0 && h * w > 0;
    const int coarse = exact ? 2 : 0;
    int rc;
    if ((rc = sx_macenko_begin(own_buffer, 1, s))) return rc;
    if (have && per_image) {
        pool_moments_kernel<<<1, kPoolThreads, 0, static_cast<cudaStream_t>(s)>>>(per_image, n, 0, n, own_buffer, 1, 0);
        SX_LAUNCHED("macenko::pool_moments_kernel");
    } else if (have && (rc = sx_macenko_moments(images, dtype, n, h, w, 1, 0, own_buffer, 1, s))) return rc;
    uint32_t epoch = first_epoch;
    if ((rc = sx_macenko_peer_combine(peer_buffers_dev, world, rank, epoch++, 0, scratch, s))) return rc;
    if ((rc = sx_macenko_basis(own_buffer, 1, 0, 1, 0, s))) return rc;  // no fallback in fit (L483-487)
    for (int stage = 0; stage < 2; ++stage) {
        if (have && (rc = sx_macenko_hist(images, dtype, n, h, w, 1, 0, stage, coarse, own_buffer, 1, s))) return rc;
        if ((rc = sx_macenko_peer_combine(peer_buffers_dev, world, rank, epoch++, 1, scratch, s))) return rc;
        if ((rc = sx_macenko_select(own_buffer, 1, 0, 1, stage, 0, s))) return rc;  // ranks + brackets, identical on every rank
        if (have && (rc = sx_macenko_hist(images, dtype, n, h, w, 1, 0, stage, 1, own_buffer, 1, s))) return rc;
        if ((rc = sx_macenko_peer_combine(peer_buffers_dev, world, rank, epoch++, 2, scratch, s))) return rc;
        if ((rc = sx_macenko_select(own_buffer, 1, 0, 1, stage, 1, s))) return rc;
    }
    Ws ws(own_buffer, 1);
    if (he) SX_CUDA(cudaMemcpyAsync(he, ws.fit, 6 * sizeof(float), cudaMemcpyDeviceToDevice, static_cast<cudaStream_t>(s)));
    if (maxc) SX_CUDA(cudaMemcpyAsync(maxc, ws.fit + 6, 2 * sizeof(float), cudaMemcpyDeviceToDevice, static_cast<cudaStream_t>(s)));
    return SX_OK;
}

int sx_macenko_fit_peers(const void *images, int dtype, int64_t n, int64_t h, int64_t w, const void *peer_buffers_dev, void *own_buffer, int world, int rank, uint32_t first_epoch, int exact,
                         void *scratch, float *he, float *maxc, sx_stream_t s) {
    return fit_peers_impl(images, dtype, n, h, w, peer_buffers_dev, own_buffer, world, rank, first_epoch, exact, scratch, he, maxc, s, nullptr);
}

// The pooled fit after its moments: basis, then per stage sample pass -> brackets -> full pass -> rank search, all on
// slot 0 of `workspace`; the result is in the FIT region of slot 0.
static int fit_stages(const void *images, int dtype, int64_t n, int64_t h, int64_t w, void *workspace, int64_t slots, sx_stream_t s) {
    int rc;
    if ((rc = sx_macenko_basis(workspace, slots, 0, 1, 0, s))) return rc;  // no fallback in fit (L483-487)
    const int64_t hw = h * w;
    const bool vec = images_vec_ok(images, nullptr, dtype, hw);
    for (int stage = 0; stage < 2; ++stage) {
        if ((rc = sx_macenko_hist(images, dtype, n, h, w, 1, 0, stage, 0, workspace, slots, s))) return rc;
        if ((rc = sx_macenko_select(workspace, slots, 0, 1, stage, 0, s))) return rc;
        if ((rc = sx_macenko_hist(images, dtype, n, h, w, 1, 0, stage, 1, workspace, slots, s))) return rc;
        // rank search; a missed bracket is re-done exactly by the same kernel (see "deterministic recovery")
        SX_DISPATCH_TV(dtype, vec, {
            const T *p = static_cast<const T *>(images);
            if (stage == SX_STAGE_ANGLE) select_recover_kernel<T, VEC, SX_STAGE_ANGLE><<<1, kThreads, 0, static_cast<cudaStream_t>(s)>>>(p, n, hw, 1, 0, workspace, slots);
            else select_recover_kernel<T, VEC, SX_STAGE_CONC><<<1, kThreads, 0, static_cast<cudaStream_t>(s)>>>(p, n, hw, 1, 0, workspace, slots);
        });
        SX_LAUNCHED("macenko::select_recover_kernel");
    }
    return SX_OK;
}

int sx_macenko_fit(const void *images, int dtype, int64_t n, int64_t h, int64_t w, float *he, float *maxc, void *workspace, int64_t workspace_bytes, sx_stream_t s) {
    if (int rc = check_images(images, dtype, n, h, w)) return rc;
    SX_REQUIRE(n > 0 && h * w > 0, "empty reference batch");
    SX_REQUIRE(he && maxc, "NULL output");
    SX_REQUIRE(workspace && workspace_bytes >= sx_macenko_workspace_bytes(1), "workspace too small");
    int rc;
    if ((rc = sx_macenko_begin(workspace, 1, s))) return rc;
    if ((rc = sx_macenko_moments(images, dtype, n, h, w, 1, 0, workspace, 1, s))) return rc;
    if ((rc = fit_stages(images, dtype, n, h, w, workspace, 1, s))) return rc;
    Ws ws(workspace, 1);
    SX_CUDA(cudaMemcpyAsync(he, ws.fit, 6 * sizeof(float), cudaMemcpyDeviceToDevice, static_cast<cudaStream_t>(s)));
    SX_CUDA(cudaMemcpyAsync(maxc, ws.fit + 6, 2 * sizeof(float), cudaMemcpyDeviceToDevice, static_cast<cudaStream_t>(s)));
    return SX_OK;
}

// fit_transform (src/stainx/base.py:L59-61 = fit(images) then transform(images)) in one call that reads the batch for
// the moments ONCE: the per-image moments the transform needs go to slots 1..n, their sum -- bit for bit the moments a
// pooled pass accumulates, see pool_moments_kernel -- to slot 0; the fit runs on slot 0, the transform pipeline on
// slots 1..n without its moments pass.  Results equal sx_macenko_fit followed by sx_macenko_transform exactly.
int sx_macenko_fit_transform(const void *images, int dtype, int64_t n, int64_t h, int64_t w, float *he, float *maxc, void *out, int out_dtype, float out_scale, void *workspace, int64_t workspace_bytes, sx_stream_t s) {
    if (int rc = check_images(images, dtype, n, h, w)) return rc;
    SX_REQUIRE(n > 0 && h * w > 0, "empty reference batch");
    SX_REQUIRE(he && maxc, "NULL output");
    if (int rc = check_output(dtype, out, out_dtype, out_scale)) return rc;
    const int64_t slots = n + 1;
    SX_REQUIRE(workspace && workspace_bytes >= sx_macenko_workspace_bytes(slots), "workspace too small (%lld < %lld: n + 1 slots)", (long long)workspace_bytes, (long long)sx_macenko_workspace_bytes(slots));
    cudaStream_t stream = static_cast<cudaStream_t>(s);
    int rc;
    if ((rc = sx_macenko_begin(workspace, slots, s))) return rc;
    if ((rc = sx_macenko_moments(images, dtype, n, h, w, 0, 1, workspace, slots, s))) return rc;
    // The fit (slot 0) forks onto its own stream: the transform's rank searches (slots 1..n) need nothing from it, only
    // `apply` does, and the fit's ~18 short dependent kernels leave the SMs idle that the transform's streaming fills.
    SideStream *side = side_stream();
    SX_REQUIRE(side != nullptr, "could not create the side streams");
    std::lock_guard<std::mutex> lock(side->fit_mu);
    cudaStream_t fs = side->fit_stream;
    SX_CUDA(cudaEventRecord(side->fit_fork, stream));
    SX_CUDA(cudaStreamWaitEvent(fs, side->fit_fork, 0));
    Ws ws(workspace, slots);
    pool_moments_kernel<<<1, kPoolThreads, 0, fs>>>(workspace, slots, 1, n, workspace, slots, 0);
    note_launch();
    rc = cudaGetLastError() == cudaSuccess ? SX_OK : sx::fail(SX_ERR_CUDA, "macenko::pool_moments_kernel launch failed");
    if (!rc) rc = fit_stages(images, dtype, n, h, w, workspace, slots, fs);
    if (!rc && (cudaMemcpyAsync(he, ws.fit, 6 * sizeof(float), cudaMemcpyDeviceToDevice, fs) != cudaSuccess || cudaMemcpyAsync(maxc, ws.fit + 6, 2 * sizeof(float), cudaMemcpyDeviceToDevice, fs) != cudaSuccess))
        rc = sx::fail(SX_ERR_CUDA, "copy of the fitted parameters failed");
    SX_CUDA(cudaEventRecord(side->fit_done, fs));  // recorded on every path: the caller's stream always joins the fit stream
    if (rc) {
        SX_CUDA(cudaStreamWaitEvent(stream, side->fit_done, 0));
        return rc;
    }
    return transform_chains(images, dtype, n, h, w, 1, ws.fit, ws.fit + 6, out, out_dtype, out_scale, workspace, slots, stream, false, side->fit_done);
}

// The same over the ranks of one NVLink node: sx_macenko_fit_peers on the sum of this rank's per-image moments, then the
// transform of this rank's images with the pooled fit.  `workspace`: private device memory of n slots (n = 0: may be
// NULL); he / maxc are required here (the transform reads them).  After a non-zero STATUS word 0 of own_buffer the caller
// repeats the call with exact = 1, as for sx_macenko_fit_peers.
int sx_macenko_fit_transform_peers(const void *images, int dtype, int64_t n, int64_t h, int64_t w, const void *peer_buffers_dev, void *own_buffer, int world, int rank, uint32_t first_epoch, int exact,
                                   void *scratch, float *he, float *maxc, void *out, int out_dtype, float out_scale, void *workspace, int64_t workspace_bytes, sx_stream_t s) {
    if (int rc = check_images(images, dtype, n, h, w)) return rc;
    SX_REQUIRE(he && maxc, "NULL output");
    const bool have = n > 0 && h * w > 0;
    if (have) {
        if (int rc = check_output(dtype, out, out_dtype, out_scale)) return rc;
        SX_REQUIRE(workspace && workspace_bytes >= sx_macenko_workspace_bytes(n), "workspace too small (%lld < %lld)", (long long)workspace_bytes, (long long)sx_macenko_workspace_bytes(n));
    }
    int rc;
    if (have) {
        if ((rc = sx_macenko_begin(workspace, n, s))) return rc;
        if ((rc = sx_macenko_moments(images, dtype, n, h, w, 0, 0, workspace, n, s))) return rc;
    }
    if (!have) return fit_peers_impl(images, dtype, n, h, w, peer_buffers_dev, own_buffer, world, rank, first_epoch, exact, scratch, he, maxc, s, nullptr);
    // as in sx_macenko_fit_transform: the fit (with its five exchanges) beside the transform's rank searches
    cudaStream_t stream = static_cast<cudaStream_t>(s);
    SideStream *side = side_stream();
    SX_REQUIRE(side != nullptr, "could not create the side streams");
    std::lock_guard<std::mutex> lock(side->fit_mu);
    cudaStream_t fs = side->fit_stream;
    SX_CUDA(cudaEventRecord(side->fit_fork, stream));
    SX_CUDA(cudaStreamWaitEvent(fs, side->fit_fork, 0));
    rc = fit_peers_impl(images, dtype, n, h, w, peer_buffers_dev, own_buffer, world, rank, first_epoch, exact, scratch, he, maxc, fs, workspace);
    SX_CUDA(cudaEventRecord(side->fit_done, fs));
    if (rc) {
        SX_CUDA(cudaStreamWaitEvent(stream, side->fit_done, 0));
        return rc;
    }
    return transform_chains(images, dtype, n, h, w, 0, he, maxc, out, out_dtype, out_scale, workspace, n, stream, false, side->fit_done);
}

}  // extern "C"
