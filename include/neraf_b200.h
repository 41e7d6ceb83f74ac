/*
 * neraf_b200.h -- C ABI of the B200-native NeRAF acoustic-field hot path.
 *
 * The reference (AmandineBtto/NeRAF) is pure Python/PyTorch and has no FFI of its
 * own; this header is the boundary a maintainer binds from the nerfstudio plugin
 * (INTEGRATION.md shows the ctypes stub).  Each entry point names the reference
 * interface it replaces (paths relative to /root/reference).
 *
 * Conventions
 *   - every pointer argument marked "dev" is a CUDA device pointer owned by the
 *     caller (normally the storage of a torch tensor); the library never allocates
 *     persistent device memory and never frees caller memory.  Scratch space is the
 *     caller-provided `workspace` whose size comes from the matching *_sizes call.
 *   - pointer arrays (`weights`, `biases`, ...) are HOST arrays of device pointers.
 *   - `stream` is a cudaStream_t (pass torch.cuda.current_stream().cuda_stream);
 *     all work is enqueued on it, nothing synchronises the device.
 *   - every function returns NERAF_OK (0) or an error code; the text of the last
 *     error of the calling thread is available through neraf_last_error().
 *     No C++ exception crosses this boundary.
 *   - there is NO CPU implementation behind these symbols: without a sm_100
 *     device they fail with NERAF_ERR_CUDA.
 */
#ifndef NERAF_B200_H
#define NERAF_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define NERAF_ABI_VERSION 9

#if defined(__GNUC__)
#define NERAF_API __attribute__((visibility("default")))
#else
#define NERAF_API
#endif

typedef void* neraf_stream_t; /* cudaStream_t */

enum {
  NERAF_OK = 0,
  NERAF_ERR_INVALID = 1,     /* bad argument (shape, alignment, null pointer) */
  NERAF_ERR_CUDA = 2,        /* CUDA runtime/driver error, text in neraf_last_error() */
  NERAF_ERR_UNSUPPORTED = 3, /* combination not implemented */
  NERAF_ERR_WORKSPACE = 4    /* workspace/pack buffer too small */
};

enum { NERAF_PREC_FP32 = 0, NERAF_PREC_BF16 = 1 };
enum { NERAF_ACT_NONE = 0, NERAF_ACT_LEAKY = 1, NERAF_ACT_TANH10 = 2 };
enum { NERAF_CRIT_SC_SLMSE = 0, NERAF_CRIT_SC_SLL1 = 1, NERAF_CRIT_MSE = 2 };
/* column order of the per-query encodings inside h */
enum { NERAF_ORDER_TIME_MIC_SRC_ROT = 0, /* NeRAF_model.py:560 (after the grid block) */
       NERAF_ORDER_MIC_SRC_TIME_ROT = 1  /* NeRAF_model.py:562 (no-grid variant)      */ };

NERAF_API int neraf_version(void);
NERAF_API const char* neraf_last_error(void);
/* sizeof() of a public struct by its type name ("neraf_gemm_job", ...), 0 for an unknown name: lets a binding check
 * its mirror of the layouts (tests/test_abi.py does, for the ctypes table). */
NERAF_API size_t neraf_abi_sizeof(const char* name);
/* Cumulative number of CUDA kernels launched by this library in this process (bench.py: gpu_launches). */
NERAF_API long long neraf_launch_count(void);
/* 1 when the current device is compute capability 10.x, else 0 (no error raised). */
NERAF_API int neraf_device_supported(void);

/* ------------------------------------------------------------------------------------------------
 * Acoustic field  (replaces NeRAFAudioModel.get_outputs NeRAF_model.py:531-566 and
 * NeRAFAudioSoundField.forward NeRAF_field.py:47-65, and their autograd backward)
 * ---------------------------------------------------------------------------------------------- */
#define NERAF_MAX_TRUNK 8

typedef struct {
  int32_t n_grid;                 /* width of the batch-invariant block of h (1024; 0 = none)          */
  int32_t n_enc;                  /* width of the per-query block of h (163 = 21 + 63 + 63 + 16)       */
  int32_t n_trunk;                /* number of trunk layers (5)                                        */
  int32_t trunk[NERAF_MAX_TRUNK]; /* trunk output widths (5096, 2048, 1024, 1024, W)                   */
  int32_t n_channels;             /* sound_rez C: number of heads                                      */
  int32_t n_freq;                 /* N_frequencies F: width of every head                              */
} neraf_field_dims;

/* Bytes of the packed-weight buffer and of the per-call workspace for batches up to max_batch. */
NERAF_API int neraf_field_sizes(const neraf_field_dims* dims, int precision, int64_t max_batch,
                      size_t* pack_bytes, size_t* workspace_bytes);

/* Derive the tensor-core operand copies of the fp32 parameters (bf16, rows padded to a multiple of 8; one copy per
 * matrix -- forward reads it K-major, backward MN-major).
 * weights[l] : dev fp32 (out,in) row-major, l = 0..n_trunk-1 trunk then n_channels heads
 *              (state_dict order soundfield.{l}.weight, STFT_linear.{c}.weight; NeRAF_field.py:41-45)
 * biases[l]  : dev fp32 (out).  Must be re-run whenever the parameters change.  No-op for FP32. */
NERAF_API int neraf_field_pack(const neraf_field_dims* dims, int precision, const float* const* weights,
                     const float* const* biases, void* pack, size_t pack_bytes, neraf_stream_t stream);

typedef struct {
  int64_t batch;                /* B: number of queries (STFT columns)                                  */
  const int64_t* time_query;    /* dev int64 (B)      batch['time_query']  NeRAF_model.py:533          */
  const double* mic_pose;       /* dev f64  (B,3)     batch['mic_pose']    :537                          */
  const double* source_pose;    /* dev f64  (B,3)     batch['source_pose'] :538                          */
  const double* rot;            /* dev f64  (B,3)     batch['rot']         :539                          */
  const float* aabb;            /* dev f32  (2,3)     scene_box.aabb       :541                          */
  float time_denominator;       /* float(max_len - 1) :534                                              */
  int32_t order;                /* NERAF_ORDER_*                                                        */
  /* If non-null the encodings above are skipped and this (B, n_enc) fp32 matrix (row stride
   * enc_ld) is used as the per-query block of h: the dense NeRAFAudioSoundField.forward(h). */
  const float* enc;
  int64_t enc_ld;
} neraf_queries;

/* out: dev fp32 (B, C*F) == (B, C, F) contiguous, log-magnitude STFT columns.
 * grid_feature: dev fp32 (n_grid) or NULL when n_grid == 0.
 * pack / repack: the bf16 operand copies (neraf_field_sizes bytes).  repack != 0 re-derives them from the
 *   current fp32 parameters inside this call (on a helper stream, overlapped with the encodings and the earlier
 *   layers) -- what a training step needs after every optimizer update; repack == 0 trusts the buffer.
 * keep != 0 (training): the workspace afterwards holds what neraf_field_backward needs (the bf16 activations of every
 * layer and the sign bit masks of the LeakyReLU gates); keep == 0 (inference) skips the masks. */
NERAF_API int neraf_field_forward(const neraf_field_dims* dims, int precision, const neraf_queries* q,
                        const float* grid_feature, const float* const* weights,
                        const float* const* biases, void* pack, size_t pack_bytes, int repack,
                        void* workspace, size_t workspace_bytes, float* out, int keep,
                        neraf_stream_t stream);

/* Backward of neraf_field_forward(keep=1) on the same workspace.
 * dout, out : dev fp32 (B, C*F).   dweights[l]/dbiases[l]: dev fp32, parameter shapes, OVERWRITTEN (bias gradients
 *             and dgrid are zeroed then accumulated: laid out back to back -- dbiases[0..], then dgrid -- they are
 *             zeroed by a single memset).
 * dgrid     : dev fp32 (n_grid) or NULL.   denc: dev fp32 (B, n_enc) row stride denc_ld, or NULL. */
NERAF_API int neraf_field_backward(const neraf_field_dims* dims, int precision, int64_t batch, const float* dout,
                         const float* out, const float* grid_feature, const float* const* weights,
                         const void* pack, void* workspace, size_t workspace_bytes,
                         float* const* dweights, float* const* dbiases, float* dgrid, float* denc,
                         int64_t denc_ld, neraf_stream_t stream);

/* Data-parallel backward (the reference refuses world_size > 1, NeRAF_pipeline.py:154-155):
 * neraf_field_backward plus the hooks a gradient exchange needs.  All options are independent; opt == NULL or all
 * zero is neraf_field_backward.
 *   defer_grid_grads : skip the two gradients of the hoisted grid block (dW1[:, :n_grid] = db1 (x) g and dgrid =
 *                W1[:, :n_grid]^T db1).  Both are LINEAR in db1 and g is replicated, so under data parallelism they are
 *                formed once from the all-reduced db1 by neraf_field_grid_grads -- 20.9 MB less to all-reduce.
 *   dw0_compact (needs defer_grid_grads): the per-query block of dW1 is written to this (trunk[0], n_enc) fp32 buffer
 *                with row stride round_up(n_enc, 8) instead of into the strided dweights[0][:, n_grid:], so that
 *                everything that must be all-reduced can sit in one contiguous buffer WITHOUT the grid block;
 *                neraf_field_grid_grads copies it back.
 *   dweights_bf16 (needs defer_grid_grads; not with mc): n_trunk + n_channels bf16 matrices; the weight-gradient GEMMs
 *                store their results THERE, rounded to bf16 (TMA stores from the epilogue), instead of fp32 into
 *                dweights / dw0_compact -- for a bf16 gradient exchange there is then no fp32 -> bf16 pass over the
 *                gradients at all.  Entry 0 is the compact (trunk[0], round_up(n_enc, 8)) block, entry l >= 1 has the
 *                shape of dweights[l] (all row lengths are multiples of 8); head entries must be contiguous.  Bases
 *                16-byte aligned.  Bias gradients (and dgrid / denc) stay fp32.
 *   phase      : 0 = the whole backward.  1 = everything except the last dgrad (the gradient of layer 1's output) and
 *                the per-query block of dW1 -- after it, every gradient except dW1 / db1 is final and can be
 *                all-reduced; 2 = exactly the rest, to be launched while that all-reduce runs.
 *   max_ctas   : upper bound on the grid of the job-list launch (0 = one CTA per SM): a concurrent collective kernel
 *                needs SMs of its own, and the persistent kernel requires all its CTAs to be co-resident.
 *   mc         : (experimental) the weight-gradient tiles are not stored but added into every rank's copy of a
 *                symmetric gradient buffer through its multicast alias (NVLS multimem.red in the epilogue; bf16 path
 *                only).  Every dweights[l] / dw0_compact must lie inside [local_base, local_base + bytes); the caller
 *                zeroes that buffer on EVERY rank before ANY rank starts this call (e.g. memset, then the loss's
 *                all-reduce), and makes the ranks meet again afterwards (e.g. the bias all-reduce) before reading the
 *                sums.  dbiases / dgrid / denc stay local sums, to be all-reduced by the caller. */
typedef struct {
  const void* local_base;   /* this device's address of the symmetric buffer                         */
  void* multicast_base;     /* multicast address of the same buffer (cuMulticast* / torch symm_mem) */
  size_t bytes;
} neraf_multicast;

/* neraf_field_forward with the spectral loss's partial sums formed by the epilogue that stores the prediction (bf16
 * path; the fp32 path runs neraf_spectral_loss_sums behind the forward): gt is the (B, C, F) target, sums f64[>=5] is
 * zeroed by the call and holds the four sums of neraf_spectral_loss_sums afterwards.  With neraf_loss_grad.losses in
 * the backward, a training step needs no loss launch at all.  gt == NULL: the call only clears sums (inside the launch
 * that precedes the GEMMs), for a following neraf_spectral_loss_sums(..., accumulate = 1) -- the variant that measured
 * faster at the reference batch, see neraf_b200/model.py. */
NERAF_API int neraf_field_forward_loss_sums(const neraf_field_dims* dims, int precision, const neraf_queries* queries,
                                  const float* grid_feature, const float* const* weights, const float* const* biases,
                                  void* pack, size_t pack_bytes, int repack, void* workspace, size_t workspace_bytes,
                                  float* out, int keep, const float* gt, double* sums, neraf_stream_t stream);

/* The spectral loss's gradient formed inside the backward instead of being read from `dout` (which may then be NULL):
 * dout[i] = neraf_spectral_loss_backward(out, gt, ..., sums, upstream = 1)[i], evaluated on the fly by the kernel that
 * applies the heads' 10*tanh derivative -- one launch and a 2 x 4 B/element round trip less per step. */
/* Data parallel: the ranks' copies of one small exchange buffer in symmetric (peer-mapped) memory -- what
 * torch.distributed._symmetric_memory hands out: peers[r] is rank r's buffer as mapped into THIS process
 * (peers[rank] the local one), NERAF_EXCHANGE_BYTES each, zero before the first use.  Kernels exchange a few words
 * per step through it (stores to every peer + a flag, spin on the local flags) instead of a host-issued collective,
 * so a whole training step stays one CUDA graph. */
#define NERAF_MAX_RANKS 16
#define NERAF_EXCHANGE_BYTES 8192
typedef struct {
  int32_t world, rank;
  void* peers[NERAF_MAX_RANKS];
} neraf_rank_exchange;

typedef struct {
  const float* gt;          /* (B, C, F) target log-magnitudes, same layout as out          */
  int64_t n_total;          /* elements over ALL ranks (the mean's denominator)             */
  int32_t criterion;        /* NERAF_CRIT_*                                                  */
  const double* sums;       /* device f64[>=4]: the (all-reduced) partial sums of the loss   */
  float w_sc, w_mag;        /* loss weights (NeRAF_model.py:597-598)                         */
  float* losses;            /* optional dev f32[2]: the two weighted losses (what                */
                            /* neraf_spectral_loss_finalize writes), formed by the same launch  */
  float* total;             /* optional dev f32[1] (with losses): their sum, what a Trainer reads */
  /* fuse_sums != 0: ONE launch forms the partial sums of (out, gt) itself, meets at a grid barrier (where, under
   * data parallelism, the last block exchanges the four sums with the other ranks through `exchange`), and then forms
   * the gradient -- no loss launch and no host collective between forward and backward.  `sums` must hold zeros on
   * entry (neraf_field_forward_loss_sums with gt == NULL clears it) and receives the GLOBAL sums; `sync` is a device
   * u32[4] that is zero before its first use and owned by this call sequence (barrier state + step counter). */
  int32_t fuse_sums;
  uint32_t* sync;
  const neraf_rank_exchange* exchange;   /* NULL: single process */
} neraf_loss_grad;

/* ------------------------------------------------------------------------------------------------
 * Data-parallel gradient exchange (SURVEY.md 8e collective 1; the reference has no multi-GPU path at all,
 * NeRAF_pipeline.py:153-157) as ONE kernel of this library over peer-mapped ("symmetric") memory: a two-shot
 * all-reduce -- multimem.ld_reduce of this rank's slice (the NVSwitch adds the ranks' copies), multimem.st of the sum
 * into every rank's copy -- chunk by chunk in the order the backward finishes the gradients.  Passed to
 * neraf_field_backward_dp (neraf_dp_options.exchange) it runs BESIDE the backward's GEMM launch: every chunk waits for
 * the completion counter (neraf_dp_options.notify) the backward advances, so gradients travel while the remaining GEMMs
 * run, and the whole training step -- no host-issued collective anywhere -- is one CUDA graph.  Called on its own it is
 * an ordinary in-stream all-reduce of the chunks.
 *   region   : every rank's gradient buffer in symmetric memory (same layout on every rank): peers[r] is rank r's as
 *              mapped into this process (peers[rank] the local one), multicast its NVLS alias or NULL (then plain peer
 *              loads / stores are used).  The sums replace the addends in place, on every rank.
 *   chunks   : byte ranges of the region (16-byte multiples), bf16 (f32 == 0) or fp32 elements, fp32 accumulation;
 *              notify / notify_count / notify_increment: dev u32 counters that this rank's producer advances by
 *              notify_increment each per step as the chunk is stored (NULL: ready when the call starts).  EVERY step
 *              that advances the counters must run this call exactly once (the kernel counts steps in `state`).
 *   signals  : every rank's NERAF_EXCHANGE_BYTES signal buffer (symmetric memory, zero before first use; may be the
 *              buffer of neraf_rank_exchange); state: dev u32[NERAF_EXCHANGE_STATE_WORDS] of this rank, zero before first use.
 *   max_ctas : 0 = one CTA per SM (128 threads, fits beside a CTA of the job-list kernel). */
#define NERAF_MAX_EXCHANGE_CHUNKS 32
typedef struct {
  int64_t offset, bytes;
  const uint32_t* notify;      /* notify_count consecutive counters: the chunk is ready when ALL have advanced */
  uint32_t notify_count;
  uint32_t notify_increment;
  int32_t f32;
  /* pull form only: where the sums of this chunk go, as fp32, instead of replacing the addends in the region (NULL:
   * they stay in the region, in the chunk's element type).  row_elems == 0: element i of the chunk -> dst[i] (dst
   * 16-byte aligned).  row_elems > 0: the chunk is a (rows, src_ld) matrix with row_elems valid columns (src_ld a
   * multiple of 8 bf16 / 4 fp32 elements): element (r, c) -> dst[r * dst_ld + c]. */
  float* dst;
  int64_t dst_ld;
  int32_t row_elems, src_ld;
} neraf_exchange_chunk;
#define NERAF_EXCHANGE_STATE_WORDS 64
typedef struct {
  int32_t world, rank, n_chunks, max_ctas;
  neraf_exchange_chunk chunks[NERAF_MAX_EXCHANGE_CHUNKS];
  void* multicast;
  void* peers[NERAF_MAX_RANKS];
  void* signals[NERAF_MAX_RANKS];
  uint32_t* state;
  int32_t pull;        /* 0: push form (multimem.st / peer stores of the sums into every rank's region).  1: pull form:
                          a rank stores the sums of its slice into its own region only and every rank fetches the other
                          ranks' slices with peer loads, delivering them to chunk.dst as fp32 when that is set -- the
                          exchange then ends with the gradients in the optimizer's buffers, no widening pass follows */
  uint64_t* trace;     /* optional dev u64[4 + 4 * n_chunks], %globaltimer stamps of the last call: [0] start, [1] this rank's
                          workers done, [2] every rank done; per chunk c at [4 + 4 c]: announced by this rank, announced
                          by every rank (seen by worker block 1), issued by worker block 1 */
} neraf_grad_exchange;
NERAF_API int neraf_dp_exchange_grads(const neraf_grad_exchange* x, neraf_stream_t stream);

#define NERAF_NOTIFY_COUNTERS 1024
typedef struct {
  const neraf_multicast* mc;
  float* dw0_compact;
  int32_t defer_grid_grads;
  int32_t phase;
  int32_t max_ctas;
  const neraf_loss_grad* loss;   /* NULL: read the upstream gradient from dout */
  void* const* dweights_bf16;    /* NULL: fp32 weight gradients in dweights / dw0_compact */
  uint32_t* notify;              /* optional dev u32[NERAF_NOTIFY_COUNTERS], never cleared by the library: completion
                                    counters for a kernel running BESIDE the backward (neraf_dp_exchange_grads).  Matrix
                                    k (k < n_trunk: weight gradient of trunk layer k; k == n_trunk: the head weight
                                    gradients, all heads as one (C F, W) matrix; k == n_trunk + 1: "every bias gradient
                                    is final") owns notify_count[k] = ceil(rows_k / 256) consecutive counters from
                                    notify_offset[k] = sum of the counts before it (rows: trunk[k], C F, batch): counter r
                                    advances by notify_increment[k] per launch when rows [256 r, 256 r + 256) of the
                                    gradient matrix are stored (k == n_trunk + 1: all of its counters must advance). */
  uint32_t* notify_offset;       /* HOST u32[n_trunk + 2] each, written by the call (with notify) */
  uint32_t* notify_count;
  uint32_t* notify_increment;
  const neraf_grad_exchange* exchange;   /* optional (with notify): the exchange kernel is launched right behind the backward's
                                    GEMM kernel as a programmatic dependent that never waits for it -- it moves in once
                                    that persistent grid is resident and runs beside it; chunks whose notify points into
                                    `notify` get their notify_increment from this call */
  int32_t zero_tail_slack;       /* the caller owns up to 3 floats behind the last bias gradient (/ dgrid): the fused loss
                                    kernel may clear the back-to-back gradient vectors in whole 16-byte words */
} neraf_dp_options;

NERAF_API int neraf_field_backward_dp(const neraf_field_dims* dims, int precision, int64_t batch, const float* dout,
                            const float* out, const float* grid_feature, const float* const* weights,
                            const void* pack, void* workspace, size_t workspace_bytes,
                            float* const* dweights, float* const* dbiases, float* dgrid, float* denc,
                            int64_t denc_ld, const neraf_dp_options* opt, neraf_stream_t stream);

/* dweight0[n, k] = dbias0[n] * grid_feature[k] (k < n_grid; row stride n_grid + n_enc; may be NULL),
 * dweight0[n, n_grid + e] = dw0_compact[n, e] (when dw0_compact != NULL; fp32, or bf16 with compact_bf16 != 0; row stride
 * round_up(n_enc, 8)) and
 * dgrid[k] = sum_n weight0[n, k] * dbias0[n] (may be NULL; overwritten; bit-reproducible: the row blocks' partial sums
 * meet in fp64): the deferred part of neraf_field_backward_dp.
 * scratch (with dgrid): dev, >= 8 * n_grid + 8 bytes, 8-byte aligned, zero before the first use; the call leaves it zero.
 * widen_*: optional second duty of the same launch after a bf16 gradient exchange: widen_dst[i] = (float)widen_src_bf16[i]
 * for i < widen_n (both 16-byte aligned): the exchanged sums land in the fp32 .grad buffer without a launch of their own. */
NERAF_API int neraf_field_grid_grads(const neraf_field_dims* dims, const float* grid_feature, const float* weight0,
                           const float* dbias0, const void* dw0_compact, int32_t compact_bf16, float* dweight0,
                           float* dgrid, void* scratch, const void* widen_src_bf16, float* widen_dst, int64_t widen_n,
                           neraf_stream_t stream);

/* Encodings only (NeRFEncoding x3 + SHEncoding + normalisation/zeroing, NeRAF_model.py:533-551):
 * enc_out dev fp32 (B, 163) row stride ld. */
NERAF_API int neraf_encode_queries(const neraf_queries* q, float* enc_out, int64_t ld, neraf_stream_t stream);

/* ------------------------------------------------------------------------------------------------
 * Spectral loss (replaces STFTLoss.forward NeRAF_evaluator.py:88-108 + weights NeRAF_model.py:593-598)
 * ---------------------------------------------------------------------------------------------- */
/* sums: dev f64[4] = { S_num = sum (e^y - e^x)^2, S_den = sum (e^y - 1e-3)^2, S_sq = sum (y-x)^2,
 *                      S_abs = sum |y-x| } over n elements; OVERWRITTEN (accumulate=0) or added to. */
NERAF_API int neraf_spectral_loss_sums(const float* pred, const float* gt, int64_t n, double* sums, int accumulate,
                             neraf_stream_t stream);
/* losses: dev f32[2] = { w_sc * sqrt(S_num)/sqrt(S_den), w_mag * (S_sq or S_abs)/n_total };
 * plain-MSE criterion: { 0, w_mag * S_sq/n_total }.  The reference weights are w_sc = 0.1*loss_factor,
 * w_mag = loss_factor (NeRAF_model.py:597-598).  n_total is the GLOBAL element count (all DP ranks). */
NERAF_API int neraf_spectral_loss_finalize(const double* sums, int64_t n_total, int criterion, float w_sc, float w_mag,
                                 float* losses, neraf_stream_t stream);
/* Single-GPU forward in ONE launch: sums + finalize (n_total == n).  scratch: dev f64[5] (the four sums, readable
 * by neraf_spectral_loss_backward, + a completion ticket), OVERWRITTEN. */
NERAF_API int neraf_spectral_loss_forward(const float* pred, const float* gt, int64_t n, int criterion, float w_sc,
                                float w_mag, double* scratch, float* losses, neraf_stream_t stream);
/* dpred[i] = upstream_sc[0] * d losses[0]/d pred[i] + upstream_mag[0] * d losses[1]/d pred[i] for the n local
 * elements (upstream_*: dev f32 scalars -- the autograd gradients of the two loss terms -- or NULL for 1). */
NERAF_API int neraf_spectral_loss_backward(const float* pred, const float* gt, int64_t n, int64_t n_total, int criterion,
                                 const double* sums, const float* upstream_sc, const float* upstream_mag,
                                 float w_sc, float w_mag, float* dpred, neraf_stream_t stream);

/* ------------------------------------------------------------------------------------------------
 * Griffin-Lim (replaces torchaudio GriffinLim as configured at NeRAF_model.py:139, used :229,:753-754,
 * and the log->magnitude conversion :746-747)
 * ---------------------------------------------------------------------------------------------- */
typedef struct {
  int32_t n_fft;        /* (N_freq_stft - 1) * 2, power of two in [64, 2048]                         */
  int32_t win_length;   /* hann(win_length, periodic) centred in n_fft                               */
  int32_t hop;          /* hop_length                                                                 */
  int32_t n_frames;     /* T                                                                          */
  int32_t n_iter;       /* 32                                                                         */
  float momentum;       /* 0.99 (the 1/(1+m) rescale of torchaudio is applied inside)                 */
  int32_t input_is_log; /* 1: spec holds log-magnitudes, convert with clip(exp(x) - 1e-3, 0, 1e4)     */
} neraf_gl_params;

NERAF_API int neraf_griffinlim_sizes(const neraf_gl_params* p, int64_t n_signals, size_t* workspace_bytes);

/* Signals are indexed (item n, channel c), n < n_items, c < n_channels.
 * spec       : dev fp32, element (n, c, frame t, bin f) at
 *              spec[n*stride_n + c*stride_c + t*stride_t + f*stride_f]
 *              (torchaudio layout (N, C, F, T): stride_f = T, stride_t = 1; field layout (N, T, C, F):
 *              stride_n = T*C*F, stride_t = C*F, stride_c = F, stride_f = 1).
 * init_phase : dev fp32 interleaved complex with the same logical indexing (strides in complex
 *              elements), or NULL for an all-ones start (rand_init=False).
 * wave       : dev fp32 (n_items, n_channels, hop*(n_frames-1)) contiguous. */
NERAF_API int neraf_griffinlim(const neraf_gl_params* p, int64_t n_items, int32_t n_channels, const float* spec,
                     int64_t stride_n, int64_t stride_c, int64_t stride_t, int64_t stride_f,
                     const float* init_phase, int64_t init_stride_n, int64_t init_stride_c,
                     int64_t init_stride_t, int64_t init_stride_f, void* workspace, size_t workspace_bytes,
                     float* wave, neraf_stream_t stream);

/* ------------------------------------------------------------------------------------------------
 * Acoustic metrics of rendered impulse responses, batched (replaces the per-RIR numpy of NeRAF_helper.py:
 * compute_t60 :48-64 / measure_rt60_advance :66-77 (pyroomacoustics measure_rt60), measure_clarity :104-107,
 * measure_edt :124-146, called from NeRAF_evaluator.py:131-190 after a device -> host copy of every waveform)
 *
 * wave : dev fp32 (n_signals, n_samples), e.g. the output of neraf_griffinlim viewed as (N*C, L).
 * t60  : dev f64 (n_signals) or NULL.  t60_highpass_hz == 0: measure_rt60(h, fs, t60_decay_db) (SoundSpaces: 30);
 *        > 0: the response is first high-passed (torchaudio highpass_biquad, Q 0.707, clamped to [-1, 1]) into the
 *        workspace (n_signals * n_samples floats) -- RAF: 200 Hz, decay 10.  A failed fit reads -1 (compute_t60's except).
 * edt  : dev f64 (n_signals) or NULL: measure_edt(h, fs, decay_db = 10); NaN when the curve never drops 10 dB.
 * c50  : dev f64 (n_signals) or NULL: measure_clarity(h, 50 ms, fs).
 * ---------------------------------------------------------------------------------------------- */
typedef struct {
  int32_t n_samples;
  double fs;
  float t60_decay_db;
  double t60_highpass_hz;
} neraf_metric_params;

NERAF_API int neraf_acoustic_metrics(const neraf_metric_params* p, const float* wave, int64_t n_signals,
                           void* workspace, size_t workspace_bytes, double* t60, double* edt, double* c50,
                           neraf_stream_t stream);

/* ------------------------------------------------------------------------------------------------
 * Audio data feed (replaces the per-sample host work of NeRAFDataset / SoundSpacesDataset.get_data,
 * NeRAF_dataset.py:89-132 and :272-296, and the DataLoader collate in NeRAF_datamanager.py:80-119)
 *
 * cache        : dev fp32 (n_rirs * n_frames, column_floats): row rir * n_frames + t is the target column
 *                log(|STFT[rir][:, :, t]| + 1e-3) flattened (C, F); columns past the end of a recording are filled
 *                the reference's way (log(min + 1e-3)) when the cache is built, so dataset index == cache row
 *                (get_id_tmp, NeRAF_dataset.py:86-87).
 * *_table      : dev f64 (n_rirs, 3) poses per RIR (dataparser outputs).
 * sample_idx   : dev i64 (batch) dataset indices in [0, n_rirs * n_frames).
 * Outputs (dev, the batch dict of NeRAF_dataset.py:129-130): data fp32 (batch, column_floats); time_query i64 (batch);
 * audio_idx i64 (batch) or NULL; mic_pose / source_pose / rot f64 (batch, 3).
 * status       : dev i32 or NULL: set to 1 when an index was out of range (that sample then reads row 0).
 * ---------------------------------------------------------------------------------------------- */
NERAF_API int neraf_gather_batch(const float* cache, int64_t n_rirs, int32_t n_frames, int32_t column_floats,
                       const double* mic_table, const double* source_table, const double* rot_table,
                       const int64_t* sample_idx, int64_t batch, float* data, int64_t* time_query,
                       int64_t* audio_idx, double* mic_pose, double* source_pose, double* rot,
                       int32_t* status, neraf_stream_t stream);

/* ------------------------------------------------------------------------------------------------
 * Operator-level entry points (used by the dense module path and by the parity tests)
 * ---------------------------------------------------------------------------------------------- */
/* fp32 CUDA-core GEMM:  C[m,n] (+)= act( sum_k A[m*a_rs + k*a_cs] * B[n*b_rs + k*b_cs] + bias[n] ) * gate'
 * gate (optional, (M,N) row stride ldg): multiply by 1 where gate > 0 else 0.1 (LeakyReLU backward). */
NERAF_API int neraf_gemm_f32(int64_t M, int64_t N, int64_t K, const float* A, int64_t a_rs, int64_t a_cs,
                   const float* B, int64_t b_rs, int64_t b_cs, const float* bias, int act,
                   const float* gate, int64_t ldg, float* C, int64_t ldc, int accumulate,
                   neraf_stream_t stream);

/* bf16 tcgen05 GEMM: D[M,N] = A[M,K] * B[N,K]^T, A and B bf16 K-major (row strides lda/ldb, multiples
 * of 8 elements, 16-byte aligned bases), fp32 accumulation in TMEM, fused epilogue.
 * Outputs (any subset): out_bf16 (M,N) row stride ld_bf16; out_bf16_t (N,M) row stride ld_t;
 * out_f32 (M,N) row stride ld_f32 (accumulate_f32: add instead of overwrite). */
typedef struct {
  const float* bias;       /* (N) or NULL                                     */
  int32_t act;             /* NERAF_ACT_*                                     */
  const void* gate;        /* bf16 (M,N) row stride ldg or NULL               */
  int64_t ldg;
  void* out_bf16;
  int64_t ld_bf16;
  void* out_bf16_t;
  int64_t ld_t;
  float* out_f32;
  int64_t ld_f32;
  int32_t accumulate_f32;
  /* job-list kernel only (neraf_gemm_bf16_jobs); must be NULL for neraf_gemm_bf16:
   * mask_out  : with act == NERAF_ACT_LEAKY and out_bf16, also store the sign bits of the 32 stored activations
   *             (m, 32 w .. 32 w + 31) in the uint32 word mask_out[w * ld_mask + m]: bit i <- column 32 w + 2 i,
   *             bit 16 + i <- column 32 w + 2 i + 1, set when the activation is negative;
   * gate_mask : LeakyReLU' gate read from such a mask instead of `gate` (same indexing). */
  void* mask_out;
  const void* gate_mask;
  int64_t ld_mask;
  /* job-list kernel only: multicast (NVLS) alias of out_f32 -- same layout, address from cuMulticast* / torch
   * symmetric memory.  When set, the results are not stored but ADDED into every GPU's copy of the buffer by the
   * NVSwitch (multimem.red.add.f32): a gradient all-reduce fused into the weight-gradient GEMM.  The buffers must be
   * zero on every rank before any rank's launch starts. */
  void* out_f32_multicast;
  /* job-list kernel only, with out_f32: the spectral loss's four partial sums over the stored values x against the
   * targets y = loss_gt[m * ld_gt + n] are ADDED to loss_sums[0..3] (f64, as neraf_spectral_loss_sums defines them:
   * sum (e^y - e^x)^2, sum (e^y - 1e-3)^2, sum (y - x)^2, sum |y - x|) by the epilogue that stores x. */
  const float* loss_gt;
  int64_t ld_gt;
  double* loss_sums;
} neraf_gemm_epilogue;

NERAF_API int neraf_gemm_bf16(int64_t M, int64_t N, int64_t K, const void* A, int64_t lda, const void* B, int64_t ldb,
                    const neraf_gemm_epilogue* epi, neraf_stream_t stream);

/* A LIST of dependent bf16 GEMMs executed by ONE persistent launch (tile-level scheduling over CTA pairs, row-block
 * dependency counters): the form in which neraf_field_forward/backward run the MLP.  Same TN contract and epilogue as
 * neraf_gemm_bf16 except: no transposed / accumulating outputs, at most one of out_bf16 / out_f32 per job.
 *   a_mn / b_mn : operand is stored (K, M) resp. (K, N) row-major and consumed MN-major (weight gradients contract
 *                 over the batch, the outer dimension of row-major activations) instead of (M, K) / (N, K).
 *   bn          : tile width 64 / 128 / 256 (>= 128 when b_mn); the tile height is 256 rows (one CTA pair).
 *   wait_job    : index (< own index) of the job whose outputs this job reads, or -1.  wait_all == 0: tile of row
 *                 block r (256 rows) starts when row block r of wait_job is complete (both jobs have the same M);
 *                 wait_all == 1: when every row block of wait_job is complete.
 *   colsum      : optional dev fp32 (N), must be zeroed by the caller: += column sums of the fp32 results.
 *   notify      : see below.
 *   bias        : must not be produced by a job of the same launch (it is prefetched before dependencies resolve);
 *                 A, B, gate and gate_mask may be.
 *   out_bf16    : ld_bf16 is a multiple of 8, so a row has round_up(N, 8) - N pad columns: they may be overwritten
 *                 with zeros (TMA stores clip with 16-byte granularity); nothing beyond them is touched.
 *   out_f32     : only the M x N results are written.
 * counters: dev scratch, >= 4 * sum_j ceil(M_j / 256) bytes. */
#define NERAF_MAX_GEMM_JOBS 24
typedef struct {
  int64_t M, N, K;
  const void* A; int64_t lda;
  const void* B; int64_t ldb;
  int32_t a_mn, b_mn;
  int32_t b_static;   /* 1: B is not written by any job of this launch (weights, data of an earlier launch): it may be
                         fetched before this job's dependency is resolved */
  int32_t bn;
  int32_t wait_job;
  int32_t wait_all;
  int32_t merge_next;   /* 1: interleave this job's tiles with those of the NEXT job (which must not wait for this one) */
  neraf_gemm_epilogue epi;
  float* colsum;
  uint32_t* notify;     /* optional dev u32[ceil(M / 256)], never cleared: counter r is advanced (gpu-scope release) as the
                           tiles of rows [256 r, 256 r + 256) are stored, by a fixed amount per launch -- lets a
                           CONCURRENT kernel wait for (parts of) this job's output */
} neraf_gemm_job;

NERAF_API int neraf_gemm_bf16_jobs(const neraf_gemm_job* jobs, int n_jobs, void* counters, size_t counters_bytes,
                         neraf_stream_t stream);

/* Pin the tile shape of neraf_gemm_bf16 (tuning / tests): code = cta_group*1000 + tile_n with cta_group in
 * {1,2}, tile_n in {64,128,256}; 0 restores the built-in cost model. */
NERAF_API int neraf_gemm_bf16_set_tile(int code);

/* fp32 (rows, cols) row stride ld_in  ->  bf16 (rows, cols) row stride ld_out (and/or its transpose
 * (cols, rows) row stride ld_t).  Either output may be NULL. */
NERAF_API int neraf_convert_bf16(const float* in, int64_t rows, int64_t cols, int64_t ld_in, void* out, int64_t ld_out,
                       void* out_t, int64_t ld_t, neraf_stream_t stream);

/* ------------------------------------------------------------------------------------------------
 * Grid-feature producer: the operators of the reference's ResNet3D (NeRAF_resnet3d.py:116-201, called at
 * NeRAF_model.py:554-556 every training step and :680-683 per rendered RIR) that are not GEMMs.  A convolution is
 * neraf_grid_im2col + neraf_gemm_bf16_jobs (bf16) / neraf_gemm_f32 (fp32); its data gradient is a GEMM +
 * neraf_grid_col2im, its weight gradient one GEMM that contracts over the voxels.  neraf_b200/gridnet.py assembles
 * them into ResNet3D_helper with the reference's parameter names.
 *
 * Activations are matrices (V, C): row v = (d * H + h) * W + w is a voxel, its C channels are contiguous (channels
 * last), row stride ld >= C elements; dtype NERAF_DT_F32 or NERAF_DT_BF16.  Every function processes one grid
 * (the reference's batch is 1).
 * ---------------------------------------------------------------------------------------------- */
enum { NERAF_DT_F32 = 0, NERAF_DT_BF16 = 1 };

typedef struct {
  int32_t in_d, in_h, in_w; /* input extent in voxels                                     */
  int32_t channels;         /* channels of the tensor the window slides over               */
  int32_t k, stride, pad;   /* cubic window; output extent (in + 2 pad - k) / stride + 1   */
} neraf_window3d;

/* col[v_out, ((kd k + kh) k + kw) C + c] = in[v_in(v_out, kd, kh, kw), c] (zero outside the grid); columns
 * [k^3 C, ld_col) are zero-filled.  The input element (v, c) is read at in[v * voxel_stride + c * channel_stride]:
 * (ld, 1) for an activation matrix, (1, V) for the reference's channels-first grid (1, C, D, H, W). */
NERAF_API int neraf_grid_im2col(const neraf_window3d* w, const void* in, int32_t in_dtype, int64_t voxel_stride,
                      int64_t channel_stride, void* col, int32_t col_dtype, int64_t ld_col, neraf_stream_t stream);
/* dx[v_in, c] = sum of dcol[v_out, kidx C + c] over the window positions that read v_in (gather, deterministic). */
NERAF_API int neraf_grid_col2im(const neraf_window3d* w, const void* dcol, int32_t dtype, int64_t ld_col, void* dx,
                      int64_t ld_dx, neraf_stream_t stream);
/* nn.Conv3d weight (c_out, c_in, k, k, k) fp32 -> GEMM operand out[co, kidx c_in + ci], row stride ld_out (pad
 * columns zero), and the weight gradient back: dw[co, ci, kidx] = sum_s dw_mat[s * partial_stride + co * ld + kidx c_in + ci]
 * over n_partials (>= 1) matrices -- the partial results of a weight-gradient GEMM split over the voxels (split-K). */
NERAF_API int neraf_grid_pack_weight(const float* weight, int64_t c_out, int64_t c_in, int64_t k3, void* out,
                           int32_t out_dtype, int64_t ld_out, neraf_stream_t stream);
NERAF_API int neraf_grid_unpack_wgrad(const float* dw_mat, int64_t ld, int64_t c_out, int64_t c_in, int64_t k3,
                            int32_t n_partials, int64_t partial_stride, float* dweight, neraf_stream_t stream);

/* nn.BatchNorm3d.  sums: dev f64 (2, C), cleared by the call: sum x, sum x^2 over the V rows. */
NERAF_API int neraf_grid_bn_stats(const void* x, int32_t dtype, int64_t V, int64_t C, int64_t ld, double* sums,
                        neraf_stream_t stream);
/* training != 0: mean / invstd (dev fp32 (C)) from the batch sums (biased variance), running estimates updated with
 * the unbiased variance when momentum > 0 (running_* may be NULL otherwise); training == 0: from the running
 * estimates (sums ignored).  invstd may be NULL (global average pooling = stats + this call's mean). */
NERAF_API int neraf_grid_bn_finalize(const double* sums, int64_t V, int64_t C, float eps, float momentum, int32_t training,
                           float* running_mean, float* running_var, float* mean, float* invstd,
                           neraf_stream_t stream);
/* y = [relu]( gamma (x - mean) invstd + beta [+ residual] ); residual may be NULL. */
NERAF_API int neraf_grid_bn_apply(const void* x, int32_t dtype, int64_t V, int64_t C, int64_t ld_x, const float* mean,
                        const float* invstd, const float* gamma, const float* beta, const void* residual,
                        int64_t ld_res, int32_t relu, void* y, int64_t ld_y, neraf_stream_t stream);
/* Backward in two passes (all matrices (V, C) with the same row stride ld).
 * reduce: g = (dy [+ dy2]) * [y > 0] (dy2, y optional) is stored to g_out and sums (dev f64 (2, C), cleared by the
 *         call) receive sum g and sum g xhat;
 * apply : dx = gamma invstd (g - (sum g + xhat sum g xhat) / V) (training) or gamma invstd g (running statistics);
 *         dgamma = sum g xhat, dbeta = sum g (dev fp32 (C), either may be NULL).  dx may alias g. */
NERAF_API int neraf_grid_bn_backward_reduce(const void* dy, const void* dy2, const void* y, const void* x, int32_t dtype,
                                  int64_t V, int64_t C, int64_t ld, const float* mean, const float* invstd,
                                  void* g_out, double* sums, neraf_stream_t stream);
NERAF_API int neraf_grid_bn_backward_apply(const void* g, const void* x, int32_t dtype, int64_t V, int64_t C, int64_t ld,
                                 const float* mean, const float* invstd, const float* gamma, const double* sums,
                                 int32_t training, void* dx, float* dgamma, float* dbeta, neraf_stream_t stream);

/* nn.MaxPool3d(k, stride, pad) (NeRAF_resnet3d.py:122): y (V_out, C), argmax dev i32 (V_out, C) = the winning input
 * voxel (first maximum in scan order, as torch's CPU kernel); backward gathers dy (+ dy2, optional: the pooled tensor
 * has two consumers, same row stride) through argmax into dx (V_in, C). */
NERAF_API int neraf_grid_maxpool(const neraf_window3d* w, const void* x, int32_t dtype, int64_t ld_x, void* y, int64_t ld_y,
                       int32_t* argmax, neraf_stream_t stream);
NERAF_API int neraf_grid_maxpool_backward(const neraf_window3d* w, const void* dy, const void* dy2, int32_t dtype, int64_t ld_dy,
                                const int32_t* argmax, void* dx, int64_t ld_dx, neraf_stream_t stream);
/* out[r, c] = v[c] * scale for r < V: gradient of the global average pooling (NeRAF_resnet3d.py:141-157). */
NERAF_API int neraf_grid_broadcast_rows(const float* v, float scale, int64_t V, int64_t C, void* out, int32_t dtype,
                              int64_t ld, neraf_stream_t stream);

#ifdef __cplusplus
}
#endif
#endif /* NERAF_B200_H */
