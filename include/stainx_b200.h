/*
 * stainx_b200.h -- C ABI of libstainx_b200.so, the B200 (sm_100a) implementation of StainX's
 * per-pixel stain-normalization hot path.
 *
 * This is the drop-in boundary.  It replaces the reference's pybind11
 * module `stainx_cuda_torch.stainx_cuda_torch` (rendeirolab/stainx v0.1.4,
 * src/stainx_cuda_torch/csrc/bindings.cpp:L19-35), whose four entry points
 *     histogram_matching(images, ref_hist)            src/stainx_cuda_torch/csrc/histogram_matching.cu:L25-169
 *     reinhard(images, target_mean, target_std)        src/stainx_cuda_torch/csrc/reinhard.cu:L25-121
 *     macenko / macenko_fast(images, HE, maxC)         src/stainx_cuda_torch/csrc/macenko.cu:L67-274
 * take and return torch tensors, and the torch-only fit path
 *     HistogramMatchingTorch.compute_reference_histograms_torch   src/stainx/backends/torch_backend.py:L143-179
 *     ReinhardTorch.compute_reference_mean_std_torch               src/stainx/backends/torch_backend.py:L308-323
 *     MacenkoTorch.compute_reference_stain_matrix_torch            src/stainx/backends/torch_backend.py:L463-519
 *
 * Conventions
 *   - plain C: raw DEVICE pointers, int64 sizes, enums as int, the CUDA stream as void*.
 *   - every function returns an int status (SX_OK = 0); sx_last_error() gives the thread-local
 *     message of the last failure.  (Reference behaviour: TORCH_CHECK -> RuntimeError.)
 *   - the library owns no memory: the caller allocates inputs, outputs and the workspace
 *     (sx_*_workspace_bytes) and keeps them alive until the stream has drained.
 *   - asynchronous: work is enqueued on `stream`; no call synchronises the host.
 *   - images are contiguous, C = 3.  SX_NCHW everywhere; SX_NHWC additionally for histogram
 *     matching (the reference accepts it there through a permute view,
 *     src/stainx/backends/torch_cuda_backend.py:L46-49).
 *   - uint8 images are [0,255]; float32 / float16 / bfloat16 images are assumed [0,1] and never
 *     max-rescaled (torch_backend.py:L103-113).
 *   - phase-level functions are split exactly where a sharded (multi-GPU) run must all-reduce
 *     statistics; the *_transform / *_fit conveniences chain the phases on one stream.
 */
#ifndef STAINX_B200_H
#define STAINX_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define SX_ABI_VERSION 2 /* 2: SX_F16 / SX_BF16, sx_peer_status*, sx_macenko_hist level 2, STATUS word 3, development hooks declared */

enum sx_status { SX_OK = 0, SX_ERR_INVALID = 1, SX_ERR_CUDA = 2, SX_ERR_UNSUPPORTED = 3 };
/* SX_F16 / SX_BF16: float images in [0,1] stored as IEEE half / bfloat16.  The reference widens such tensors to
 * float32, computes, and casts the result back (torch_backend.py:L103-131, src/stainx_cuda_torch/csrc/
 * histogram_matching.cu:L63-65); here the kernels load and store the 16-bit values themselves (8 pixels per
 * 128-bit vector, float32 arithmetic in registers, round-to-nearest-even stores): half the bytes per pixel.
 * Outputs have the dtype of the input. */
enum sx_dtype { SX_U8 = 0, SX_F32 = 1, SX_F16 = 2, SX_BF16 = 3 };
enum sx_layout { SX_NCHW = 0, SX_NHWC = 1 };

typedef void *sx_stream_t; /* cudaStream_t */

/* ---- library ------------------------------------------------------------------------------- */
int sx_abi_version(void);
const char *sx_last_error(void);
/* SM count, compute capability and L2 size of the current device. */
int sx_device_info(int *sm_count, int *cc_major, int *cc_minor, int64_t *l2_bytes);
/* Number of kernels this library has launched in the calling process (monotonic). */
int64_t sx_kernel_launches(void);

/* Peer exchanges (the *_peers functions below wait inside a kernel for the other ranks of the node).
 * Every such wait is bounded: SX_PEER_TIMEOUT_MS in the environment at first use, default 20000.  A
 * kernel whose wait expires records it in a per-device status record (pinned host memory, read
 * without synchronising) and carries on with meaningless results; from then on every *_peers call on
 * that device fails with SX_ERR_CUDA until sx_peer_status_clear().  sx_peer_status reports the record
 * of the current device: *timed_out != 0, the rank that was waited for, and the epoch. */
int sx_peer_status(int *timed_out, int *waited_for_rank, uint32_t *epoch);
int sx_peer_status_clear(void);

/* ---- histogram matching --------------------------------------------------------------------
 * Reference: torch_backend.py:L194-301 (oracle), csrc/histogram_matching.cu:L49-226 (CUDA). */

/* H3: per-channel 256-bin counts of the whole batch, ADDED into counts[3][256] (uint64; the caller
 * zeroes it, or keeps accumulating over shards).  float32 input is quantised by truncating
 * clamp(x*255, 0, 255) (torch_backend.py:L115-120). */
int sx_hm_hist(const void *images, int dtype, int layout, int64_t n, int64_t h, int64_t w,
               uint64_t *counts, sx_stream_t stream);
/* H1 (fit): ref_hist[c][b] = float(counts) / (torch-order float32 sum + 1e-8)  (L139-141). */
int sx_hm_ref_hist(const uint64_t *counts, float *ref_hist, sx_stream_t stream);
/* H2a: reference CDF: h/(sum(h)+1e-8), cumsum with a double accumulator (L221-223). */
int sx_hm_ref_cdf(const float *ref_hist, float *ref_cdf, sx_stream_t stream);
/* H2b: source CDF from counts and npix = N*H*W, searchsorted + interpolation -> lut[3][256]
 * float32 in [0,255] (L234-281).  npix < 0: use the sum of each channel's counts (sharded runs). */
int sx_hm_build_lut(const uint64_t *counts, int64_t npix, const float *ref_cdf, float *lut,
                    sx_stream_t stream);
/* H4: out = lut[c][v]; uint8 -> uint8 (truncated), float32 -> float32 clamp(lut/255, 0, 1)
 * (L285-298).  `out` has the dtype and layout of `images`. */
int sx_hm_apply(const void *images, int dtype, int layout, int64_t n, int64_t h, int64_t w,
                const float *lut, void *out, sx_stream_t stream);
/* Sharded batches on one NVLink node: H2b with the SUM all-reduce of the counts fused into the
 * kernel (peer loads/stores over NVLink; no NCCL call).  Every rank owns a zero-initialised buffer
 * of sx_hm_peer_buffer_bytes() that all ranks have mapped (symmetric memory / CUDA IPC):
 * uint64 counts[2][3][256] then uint32 flags[64].  peer_buffers_dev is a DEVICE array of `world`
 * pointers, entry p = rank p's buffer as mapped on this device.  Step `epoch` (1, 2, 3, ... the same
 * on every rank): zero counts[epoch & 1] of the own buffer, sx_hm_hist into it, then this call; it
 * publishes the own counts, waits for every rank's, sums them in rank order (bit-identical on all
 * ranks) and builds the LUT; counts_out (optional) receives the summed counts.  Replaces
 * dist.all_reduce + sx_hm_build_lut(npix = -1). */
int64_t sx_hm_peer_buffer_bytes(void);
int sx_hm_build_lut_peers(const void *peer_buffers_dev, int world, int rank, uint32_t epoch,
                          const float *ref_cdf, float *lut, uint64_t *counts_out, sx_stream_t stream);
/* The whole sharded transform in one call (one chain of dependent launches): zeroes counts[epoch & 1] of
 * own_buffer (this rank's peer-mapped buffer, i.e. peer_buffers[rank]), sx_hm_hist into them,
 * sx_hm_build_lut_peers, sx_hm_apply.  workspace: >= 3072 bytes (the LUT). */
int sx_hm_transform_peers(const void *images, int dtype, int layout, int64_t n, int64_t h, int64_t w,
                          const void *peer_buffers_dev, void *own_buffer, int world, int rank, uint32_t epoch,
                          const float *ref_cdf, void *out, void *workspace, int64_t workspace_bytes,
                          sx_stream_t stream);
/* Workspace for the chained transform/fit below (bytes). */
int64_t sx_hm_workspace_bytes(void);
/* zero counts -> hist -> ref_cdf + build_lut -> apply on one stream (single-device transform), enqueued as one
 * chain of programmatic dependent launches: only the first kernel relies on ordinary stream order, each later
 * kernel is resident before the one in front of it has drained and waits (griddepcontrol.wait) before it reads
 * what that kernel wrote.  To the caller the call is ordered like a plain sequence of launches. */
int sx_hm_transform(const void *images, int dtype, int layout, int64_t n, int64_t h, int64_t w,
                    const float *ref_hist, void *out, void *workspace, int64_t workspace_bytes,
                    sx_stream_t stream);
/* hist -> ref_hist (single-device fit). */
int sx_hm_fit(const void *images, int dtype, int layout, int64_t n, int64_t h, int64_t w,
              float *ref_hist, void *workspace, int64_t workspace_bytes, sx_stream_t stream);

/* ---- Reinhard ------------------------------------------------------------------------------
 * Reference: torch_backend.py:L16-101, L308-355 (oracle), csrc/reinhard.cu:L45-139 (CUDA). */

/* R1+R2: fused RGB->LAB and whole-batch statistics.  ADDS into sums[7] (double):
 *   [0..2] sum(lab_c - 128), [3..5] sum((lab_c - 128)^2), [6] pixel count.
 * The caller zeroes sums (and, when sharded, all-reduces it with SUM before finalize). */
int sx_reinhard_stats(const void *images, int dtype, int64_t n, int64_t h, int64_t w,
                      double *sums, sx_stream_t stream);
/* mean[3], std[3] (unbiased, Bessel) float32 from sums (L320-321 / L345-346). */
int sx_reinhard_finalize(const double *sums, float *mean, float *std, sx_stream_t stream);
/* R1+R3+R4: RGB->LAB, ((lab-src_mean)/(src_std+1e-8))*ref_std+ref_mean, LAB->RGB, clamp.
 * float32 -> float32 [0,1]; uint8 -> uint8 = trunc(clamp(rgb*255,0,255)) (L349-355). */
int sx_reinhard_apply(const void *images, int dtype, int64_t n, int64_t h, int64_t w,
                      const float *src_mean, const float *src_std, const float *ref_mean,
                      const float *ref_std, void *out, sx_stream_t stream);
/* Sharded batches on one NVLink node: finalize with the SUM all-reduce of the sums fused into the
 * kernel (peer loads over NVLink, no NCCL call; protocol of sx_hm_build_lut_peers).  Every rank owns
 * a zero-initialised, peer-mapped buffer of sx_reinhard_peer_buffer_bytes(): double sums[2][8] then
 * uint32 flags[64].  Step `epoch` (1, 2, ...): zero sums[epoch & 1] of the own buffer,
 * sx_reinhard_stats into it, then this call. */
int64_t sx_reinhard_peer_buffer_bytes(void);
int sx_reinhard_finalize_peers(const void *peer_buffers_dev, int world, int rank, uint32_t epoch,
                               float *mean, float *std, sx_stream_t stream);
int64_t sx_reinhard_workspace_bytes(void);
/* stats -> finalize -> apply on one stream. */
int sx_reinhard_transform(const void *images, int dtype, int64_t n, int64_t h, int64_t w,
                          const float *ref_mean, const float *ref_std, void *out, void *workspace,
                          int64_t workspace_bytes, sx_stream_t stream);
/* stats -> finalize (single-device fit, L308-323). */
int sx_reinhard_fit(const void *images, int dtype, int64_t n, int64_t h, int64_t w, float *mean,
                    float *std, void *workspace, int64_t workspace_bytes, sx_stream_t stream);

/* ---- Macenko -------------------------------------------------------------------------------
 * Reference: torch_backend.py:L362-560 (oracle), src/stainx_cuda_torch/csrc/macenko.cu:L67-266
 * and csrc/macenko.cu:L145-262 (CUDA).
 *
 * Statistics live in "slots": transform uses one slot per image (every statistic is per image,
 * torch_backend.py:L556-558); fit pools all images into slot 0 (L477-485).  The workspace holds,
 * per slot, the OD moments, the order-statistic histograms and the fitted HE / maxC.
 *
 * Phase order (`hist` level 0 is a ~3 % subsample pass, level 1 a full streaming pass; each `select`
 * is a tiny one-CTA-per-slot kernel):
 *   begin -> moments -> basis -> [moments_fallback]                            (M1-M4)
 *         -> hist(ANGLE,0) -> select(ANGLE,0) -> hist(ANGLE,1) -> select(ANGLE,1)   (M5-M7)
 *         -> hist(CONC,0)  -> select(CONC,0)  -> hist(CONC,1)  -> select(CONC,1)    (M8-M9)
 *         -> apply (M10)   or   read the FIT region (M11)
 * A sharded pooled fit combines the regions named by sx_macenko_region() over the ranks after
 * `moments` (MOMENTS, ODRANGE) and after every `hist` (level 0: HIST1, COUNTERS; level 1: HIST2,
 * COUNTERS, VMIN, VMAX) -- with all-reduces, or on one NVLink node with sx_macenko_peer_combine.
 * sx_macenko_transform does not use the phase functions one by one: it chains lean streaming kernels
 * and per-image kernels of its own (eight launches per chain; batches of >= 64 MB run as up to three
 * part-batch chains, all but the first on library-owned side streams, and every chain's four per-image
 * kernels on a library-owned high-priority stream of its own; all of them are forked from and joined
 * into `stream`, so the call stays stream-ordered and graph-capturable; the library streams are per device,
 * so while one thread CAPTURES such a call into a CUDA graph no other thread may enqueue a multi-chain Macenko
 * call on the same device -- the shared streams are in capture mode until the capture ends). */
enum sx_macenko_stage { SX_STAGE_ANGLE = 0, SX_STAGE_CONC = 1 };
enum sx_macenko_region_id {
    SX_REGION_MOMENTS = 0,  /* int64   [slots][12]   reduce: SUM.  Fixed-point (scale 2^22) count and shifted first / second moments
                             *                        of log2(255 x + 1) over the kept rows, [10] = pixels in the slot (unscaled).
                             *                        Integer sums: the combined value does not depend on the order of addition. */
    SX_REGION_ODRANGE = 1,  /* float32 [slots][8]    reduce: MAX  (-min_c x3, max_c x3, [6] pixels per sampled group of the
                             *                        rank's kernel variant -- the combined maximum sizes every rank's brackets) */
    SX_REGION_HIST1 = 2,    /* int32   [slots][2][4096]  sample histogram; reduce: SUM (wraps mod 2^32) */
    SX_REGION_HIST2 = 3,    /* int32   [slots][2][4096]  cells inside the bracket; reduce: SUM */
    SX_REGION_VMIN = 4,     /* float32 [slots][2][4096]  reduce: MIN */
    SX_REGION_VMAX = 5,     /* float32 [slots][2][4096]  reduce: MAX */
    SX_REGION_FIT = 6,      /* float32 [slots][8]: HE row-major (6) + maxC (2); read-only result */
    SX_REGION_COUNTERS = 7, /* int64   [slots][8]: rows below the bracket (2), sampled rows (2), pad; reduce: SUM */
    SX_REGION_STATUS = 8,   /* int32   [slots][4]: [0] bit q set = rank q fell outside its bracket and was not recovered (never
                             *           expected); [1] = stages of this slot that sx_macenko_transform re-ran with an exact bracket */
    SX_REGION_PIPECTRL = 9  /* uint32  [16]: control words of the transform pipeline (tile / role dispensers, development counters) */
};

int64_t sx_macenko_workspace_bytes(int64_t slots);
/* Byte offset and size of a region inside the workspace. */
int sx_macenko_region(int64_t slots, int region, int64_t *offset, int64_t *bytes);
/* Sharded pooled fit on one NVLink node: the combine of slot 0's statistics over all ranks in ONE
 * kernel (peer loads over NVLink) instead of 2-4 NCCL all-reduces per step.  Every rank's ONE-SLOT
 * workspace lives at the start of a zero-initialised, peer-mapped buffer of
 * sx_macenko_peer_buffer_bytes() (workspace, then uint32 flags[2][64]); peer_buffers_dev = device
 * array of `world` pointers.  Call with the same increasing `epoch` (1, 2, ...) on every rank after
 * moments (which = 0: MOMENTS sum, ODRANGE max), after hist(., 0) (1: HIST1, COUNTERS sum) and
 * after hist(., 1) (2: HIST2, COUNTERS sum, VMIN min, VMAX max); `scratch` is private device memory
 * of sx_macenko_peer_scratch_bytes().  On return (stream order) the own regions hold the combined
 * values, bit-identical on every rank. */
int64_t sx_macenko_peer_buffer_bytes(void);
int64_t sx_macenko_peer_scratch_bytes(void);
int sx_macenko_peer_combine(const void *peer_buffers_dev, int world, int rank, uint32_t epoch,
                            int which, void *scratch, sx_stream_t stream);
/* The whole sharded pooled fit in one call: begin, moments, the five exchanges (epochs first_epoch .. first_epoch + 4; the
 * caller advances its epoch counter by 5), basis, both stages' hist / select pairs.  own_buffer = this rank's peer-mapped
 * buffer (peer_buffers[rank]); exact != 0 uses the exact coarse pass (hist level 2) instead of the sample pass; n = 0 is
 * allowed (a rank without reference images still takes part in the exchanges); he / maxc (device, optional) receive the
 * result, which is also in the FIT region of the workspace.  The caller should read STATUS word 0 afterwards and repeat
 * with exact = 1 if it is non-zero (identical on every rank). */
int sx_macenko_fit_peers(const void *images, int dtype, int64_t n, int64_t h, int64_t w, const void *peer_buffers_dev,
                         void *own_buffer, int world, int rank, uint32_t first_epoch, int exact, void *scratch, float *he,
                         float *maxc, sx_stream_t stream);
/* Initialise the workspace (must precede moments). */
int sx_macenko_begin(void *workspace, int64_t slots, sx_stream_t stream);
/* M1-M3: OD = -ln((255x+1)/240); mask min_c OD >= 0.15; accumulates count and shifted first and
 * second moments of the kept rows, and the per-channel OD range of all rows.
 * pooled != 0: every image goes to slot 0, else image i goes to slot slot0 + i. */
int sx_macenko_moments(const void *images, int dtype, int64_t n, int64_t h, int64_t w, int pooled,
                       int64_t slot0, void *workspace, int64_t slots, sx_stream_t stream);
/* M3-M4: unbiased covariance -> symmetric 3x3 eigen-decomposition -> E = eigvecs[:, (1, 2)].
 * allow_fallback != 0 (transform): slots with fewer than 3 kept rows are flagged to use all rows
 * (torch_backend.py:L409-410); run moments_fallback afterwards.
 * basis / select act on the slot range [slot0, slot0 + count). */
int sx_macenko_basis(void *workspace, int64_t slots, int64_t slot0, int64_t count,
                     int allow_fallback, sx_stream_t stream);
/* Transform only: slots flagged by `basis` re-accumulate every row and get their basis (one CTA per
 * image; CTAs of unflagged slots exit at once). */
int sx_macenko_moments_fallback(const void *images, int dtype, int64_t n, int64_t h, int64_t w,
                                int64_t slot0, void *workspace, int64_t slots, sx_stream_t stream);
/* One order-statistic pass.  stage ANGLE: nearest-rank 1st/99th percentile of
 * phi = atan2(OD.e_large, OD.e_mid) over the kept rows (M5-M6); stage CONC: 99th percentile of
 * each least-squares concentration row over all rows (M8-M9).  level 0 histograms a 12-bit
 * prefix of a monotone key over a pseudo-random subsample (about 4096 pixel groups per image);
 * level 1 is the full pass: it counts the rows below the bracket chosen by select(.,0), resolves
 * the bracket into 4096 cells and records the exact min/max value of every cell.
 * level 2 = level 0 over EVERY pixel group: select(.,0) after it yields an exact bracket (the coarse bin of
 * the wanted rank).  It is the recovery of a missed sample bracket (STATUS region): sx_macenko_transform and
 * sx_macenko_fit run it inside their rank-search kernels for the slots that need it; a sharded fit repeats
 * the stage with it (stainx_b200/backends/torch_cuda_backend.py). */
int sx_macenko_hist(const void *images, int dtype, int64_t n, int64_t h, int64_t w, int pooled,
                    int64_t slot0, int stage, int level, void *workspace, int64_t slots,
                    sx_stream_t stream);
/* Per-slot step after a hist pass.  level 0: wanted ranks and their brackets (+-8 sigma of the
 * sample rank, rounded out to 12-bit prefixes).  level 1: rank search inside the bracket;
 * (ANGLE,1) also forms HE, its pseudo-inverse and the key range of the concentrations (M7) and
 * re-arms the histograms; (CONC,1) stores maxC. */
int sx_macenko_select(void *workspace, int64_t slots, int64_t slot0, int64_t count, int stage,
                      int level, sx_stream_t stream);
/* M10: C = pinv(HE_src).OD scaled by maxc_ref/maxC_src; OD' = he_ref.C; rgb = clamp(240 exp(-OD'),
 * 0, 255) * out_scale.  out_dtype SX_U8 (only for uint8 input, truncated; out_scale ignored),
 * SX_F32 (uint8 / float32 input), or the input's own 16-bit float type (SX_F16 / SX_BF16 input: the
 * [0,255] value rounded to that type, then -- out_scale 1/255 -- divided by 255 and rounded again,
 * which is what the reference's cast-back followed by `/ 255.0` gives).  out_scale = 1 keeps the reference's [0,255] float output, 1/255 folds the
 * normalize_to_0_1 division (src/stainx/normalizers/_template.py:L111-112) into the store. */
int sx_macenko_apply(const void *images, int dtype, int64_t n, int64_t h, int64_t w, int64_t slot0,
                     const float *he_ref, const float *maxc_ref, void *out, int out_dtype,
                     float out_scale, void *workspace, int64_t slots, sx_stream_t stream);
/* Whole per-image transform on one stream (slots = n). */
int sx_macenko_transform(const void *images, int dtype, int64_t n, int64_t h, int64_t w,
                         const float *he_ref, const float *maxc_ref, void *out, int out_dtype,
                         float out_scale, void *workspace, int64_t workspace_bytes,
                         sx_stream_t stream);
/* Whole pooled fit on one stream (slots = 1): he[3][2] row-major, maxc[2] (device pointers). */
int sx_macenko_fit(const void *images, int dtype, int64_t n, int64_t h, int64_t w, float *he,
                   float *maxc, void *workspace, int64_t workspace_bytes, sx_stream_t stream);
/* fit_transform (src/stainx/base.py:L59-61: fit(images) then transform(images)) as one call that reads the batch
 * for the moments ONCE instead of twice: the per-image moments of the transform go to slots 1..n, their sum -- the
 * MOMENTS region holds fixed-point integers, so the sum of the slots is bit for bit what a pooled moments pass
 * accumulates -- to slot 0, on which the pooled fit runs; the transform pipeline then skips its moments pass.
 * he / maxc / out equal sx_macenko_fit followed by sx_macenko_transform on the same images exactly.
 * workspace_bytes >= sx_macenko_workspace_bytes(n + 1). */
int sx_macenko_fit_transform(const void *images, int dtype, int64_t n, int64_t h, int64_t w, float *he, float *maxc,
                             void *out, int out_dtype, float out_scale, void *workspace, int64_t workspace_bytes,
                             sx_stream_t stream);
/* The same for a batch sharded over the ranks of one NVLink node (the batch-mode pooled fit of BASELINE config c4):
 * sx_macenko_fit_peers on the sum of this rank's per-image moments, then the transform of this rank's images with the
 * pooled fit.  workspace = private device memory of sx_macenko_workspace_bytes(n) (n = 0: may be NULL, the rank still
 * takes part in the five exchanges); he / maxc are required.  As for sx_macenko_fit_peers the caller reads STATUS word 0
 * of own_buffer afterwards and repeats the call with exact = 1 if it is non-zero. */
int sx_macenko_fit_transform_peers(const void *images, int dtype, int64_t n, int64_t h, int64_t w,
                                   const void *peer_buffers_dev, void *own_buffer, int world, int rank,
                                   uint32_t first_epoch, int exact, void *scratch, float *he, float *maxc, void *out,
                                   int out_dtype, float out_scale, void *workspace, int64_t workspace_bytes,
                                   sx_stream_t stream);

/* ---- development hooks ----------------------------------------------------------------------
 * Launch-geometry knobs for tools/probe.py and the ncu scripts.  NOT part of the drop-in surface:
 * they set process-global state, are not thread-safe, and return SX_ERR_UNSUPPORTED (changing
 * nothing) unless SX_ENABLE_TUNING=1 was in the environment before their first call.  With the
 * hooks inert the library has no mutable global state besides per-device cached attributes, the
 * launch counter, the side / helper streams of sx_macenko_transform and the peer status record above. */
int sx_hm_set_tuning(int hist_byte_counters, int hist_ctas_per_sm, int apply_ctas_per_sm);
int sx_reinhard_set_tuning(int ctas_per_sm);
int sx_macenko_set_tuning(int ctas_per_sm, int64_t phase_kernels);
/* Timeline of the transform chains: enable != 0 arms it for the next multi-chain sx_macenko_transform; enable == 0
 * synchronises the device and writes one line per kernel ("chain kernel completion-time-us") into buf. */
int sx_macenko_trace(int enable, char *buf, int64_t cap);

#ifdef __cplusplus
}
#endif
#endif /* STAINX_B200_H */
