// Acoustic-field forward / backward: the launch sequence behind neraf_field_forward/backward.
//
// Replaces NeRAFAudioModel.get_outputs (/root/reference/NeRAF/NeRAF_model.py:531-566) +
// NeRAFAudioSoundField.forward (NeRAF_field.py:47-65) and their autograd backward.
//
// Layer 1 is factored: the first n_grid columns of h are the same grid feature g for every query
// (NeRAF_model.py:557-558 `expand`), so  h W1^T + b1 = enc W1[:, G:]^T + (b1 + W1[:, :G] g)  -- one
// mat-vec per step instead of a (B x 1024) GEMM slice; in backward  dW1[:, :G] = db1 (x) g  and
// dg = W1[:, :G]^T db1.  The reference's dense path costs 20.42 M MAC/query, this one 15.20 M.
//
// Two precisions share the sequence:
//   FP32  CUDA-core GEMMs on the fp32 parameters directly (parity path, 1e-5 gate)
//   BF16  the job-list tcgen05 kernel (gemm_mega.cu) on bf16 operand copies of the parameters.  One copy per
//         weight matrix, (out, in) row-major: forward reads it K-major, dgrad reads the SAME copy MN-major, and the
//         weight gradients read the row-major activations / gradients MN-major, so nothing is ever transposed.
//
// Launches per bf16 train step: prep (encodings + c1 + layer-1 operand copy) | pack of the other layers on a helper
// stream | layer 1 | layers 2..heads | ... loss ... | memset + head gradient | whole backward | grid-block gradients.
#include <cmath>
#include <cstdlib>
#include <mutex>

#include "common.cuh"
#include "kernels.h"

namespace neraf {

namespace {

struct Layout {
  int L;                 // trunk layers
  int C, F, CF, W;       // heads
  int G, E;              // grid / per-query widths of h
  int n[NERAF_MAX_TRUNK];   // trunk widths
  int k[NERAF_MAX_TRUNK];   // trunk input widths (k[0] = E: the per-query block only)
  // ---- pack (bf16) offsets in bytes
  size_t w[NERAF_MAX_TRUNK];
  int64_t ldw[NERAF_MAX_TRUNK];
  size_t wh, bh;
  int64_t ldwh;
  size_t pack_bytes;
  // ---- workspace offsets in bytes
  size_t c1, enc;
  int64_t ld_enc;
  size_t x[NERAF_MAX_TRUNK], dz[NERAF_MAX_TRUNK];
  int64_t ldx[NERAF_MAX_TRUNK];
  size_t mask[NERAF_MAX_TRUNK];      // sign bit masks of the pre-activations: (ceil(n/32), ld_mask) uint32, bf16 path
  int64_t ld_mask;
  size_t dzh;
  int64_t ld_h;
  size_t counters, counters_bytes;   // dependency counters of the job-list kernel
  size_t gscratch, gscratch_bytes;   // fp64 accumulator of dg + ticket (grid_grads), directly behind the counters: both are
                                     // cleared by field_prep_kernel in every forward
  size_t dwh;                        // (C*F, W) fp32 joint head weight gradient (only used when C > 1)
  size_t ws_bytes;
};

inline size_t take(size_t& cursor, size_t bytes) {
  const size_t at = cursor;
  cursor += (bytes + 255) / 256 * 256;
  return at;
}

int make_layout(const neraf_field_dims* d, int precision, int64_t batch, Layout* lo) {
  NERAF_REQUIRE(d, "field: dims is null");
  NERAF_REQUIRE(d->n_trunk >= 1 && d->n_trunk <= NERAF_MAX_TRUNK, "field: n_trunk %d out of range", d->n_trunk);
  NERAF_REQUIRE(d->n_enc > 0 && d->n_grid >= 0 && d->n_channels >= 1 && d->n_freq >= 1, "field: bad dims");
  NERAF_REQUIRE(d->n_channels <= 8, "field: at most 8 heads");
  NERAF_REQUIRE(precision == NERAF_PREC_FP32 || precision == NERAF_PREC_BF16, "field: unknown precision %d", precision);
  NERAF_REQUIRE(batch >= 0 && batch < (1ll << 30), "field: batch %lld out of range", (long long)batch);
  Layout& l = *lo;
  l.L = d->n_trunk; l.C = d->n_channels; l.F = d->n_freq; l.CF = l.C * l.F; l.G = d->n_grid; l.E = d->n_enc;
  for (int i = 0; i < l.L; ++i) {
    NERAF_REQUIRE(d->trunk[i] > 0, "field: trunk[%d] <= 0", i);
    l.n[i] = d->trunk[i];
    l.k[i] = i == 0 ? l.E : d->trunk[i - 1];
  }
  l.W = l.n[l.L - 1];
  const bool bf = precision == NERAF_PREC_BF16;
  const size_t es = bf ? 2 : 4;

  size_t cur = 0;
  if (bf) {
    for (int i = 0; i < l.L; ++i) {
      l.ldw[i] = round_up(l.k[i], 8);
      l.w[i] = take(cur, (size_t)l.n[i] * l.ldw[i] * 2);
    }
    l.ldwh = round_up(l.W, 8);
    l.wh = take(cur, (size_t)l.CF * l.ldwh * 2);
    l.bh = take(cur, (size_t)l.CF * 4);
  }
  l.pack_bytes = cur;

  cur = 0;
  const size_t B = (size_t)batch;
  l.c1 = take(cur, (size_t)l.n[0] * 4);
  l.ld_enc = bf ? round_up(l.E, 8) : l.E;
  l.enc = take(cur, B * l.ld_enc * es);
  l.ld_mask = round_up(batch > 0 ? batch : 1, 64);
  for (int i = 0; i < l.L; ++i) {
    l.ldx[i] = bf ? round_up(l.n[i], 8) : l.n[i];
    l.x[i] = take(cur, B * l.ldx[i] * es);
    l.dz[i] = take(cur, B * l.ldx[i] * es);
    l.mask[i] = bf ? take(cur, (size_t)ceil_div(l.n[i], 32) * l.ld_mask * 4) : 0;
  }
  l.ld_h = bf ? round_up(l.CF, 8) : l.CF;
  l.dzh = take(cur, B * l.ld_h * es);
  l.counters_bytes = (size_t)NERAF_MEGA_MAX_JOBS * (size_t)(ceil_div(batch > 0 ? batch : 1, 256) + 32) * 4;
  l.counters = take(cur, l.counters_bytes);
  l.gscratch_bytes = (size_t)round_up((int64_t)l.G * 8 + 8, 256);
  l.gscratch = take(cur, l.gscratch_bytes);
  l.dwh = (bf && l.C > 1) ? take(cur, (size_t)l.CF * l.W * 4) : 0;
  l.ws_bytes = cur;
  return NERAF_OK;
}

// Helper stream for the per-step re-pack of layers 2.. (off the critical chain encodings -> layer 1): forked from and
// joined back into the caller's stream with events, so it is capturable in a CUDA graph.
struct SideStream {
  cudaStream_t stream = nullptr;
  cudaEvent_t fork = nullptr, done = nullptr;
  bool ok = false;
};

SideStream* side_stream() {
  static SideStream per_dev[64];
  static std::mutex mu;
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return nullptr;
  std::lock_guard<std::mutex> lock(mu);
  SideStream& s = per_dev[dev];
  if (!s.ok) {
    if (getenv("NERAF_NO_SIDE_STREAM")) return nullptr;
    if (cudaStreamCreateWithFlags(&s.stream, cudaStreamNonBlocking) != cudaSuccess) { cudaGetLastError(); return nullptr; }
    if (cudaEventCreateWithFlags(&s.fork, cudaEventDisableTiming) != cudaSuccess) { cudaGetLastError(); return nullptr; }
    if (cudaEventCreateWithFlags(&s.done, cudaEventDisableTiming) != cudaSuccess) { cudaGetLastError(); return nullptr; }
    s.ok = true;
  }
  return &s;
}

// Tile width of a job from a small cost model calibrated on the per-tile timeline (tools/mega_trace.py): a CTA pair
// spends max(MMA, epilogue) per tile -- k-blocks stream from L2 at ~0.36 us for bn <= 128 and ~0.41 us for bn = 256,
// the epilogue costs ~0.9 us per 32-column chunk of a warp plus ~0.9 us to publish the tile -- and the job takes
// ceil(tiles / pairs) rounds of that (the epilogue of the last round is exposed).
int choose_bn(int64_t M, int64_t N, int64_t K, int min_bn) {
  const double pairs = sm_count() / 2;
  const double kb = (double)ceil_div(K, 64);
  int best = min_bn;
  double best_t = 1e30;
  for (int bn = min_bn; bn <= 256; bn *= 2) {
    const double tiles = (double)ceil_div(M, 256) * (double)ceil_div(N, bn);
    const double mma = kb * (bn == 256 ? 0.41 : 0.36);
    const double epi = 0.9 * (bn / 64) + 0.9;
    const double t = (ceil(tiles / pairs) - 1.0) * (mma > epi ? mma : epi) + mma + epi;   // last round: nothing overlaps
    // ties go to the narrower tile (shorter epilogue tail) -- except for epilogue-bound jobs (layer 1 forward: 3
    // k-blocks), where the wider tile means fewer rounds and fewer publishes: 256 instead of 128 columns measured
    // 4.5 us per step faster at B=2048 (tools/ab_step.py)
    if (t < best_t * (mma < epi ? 1.02 : 0.97)) { best_t = t; best = bn; }
  }
  return best;
}

MegaJob make_job(int64_t M, int64_t N, int64_t K, const void* A, int64_t lda, const void* B, int64_t ldb, int wait_job,
                 int wait_all) {
  MegaJob j = {};
  j.M = M; j.N = N; j.K = K; j.A = A; j.lda = lda; j.B = B; j.ldb = ldb;
  j.bn = choose_bn(M, N, K, 64);
  j.wait_job = wait_job; j.wait_all = wait_all;
  j.b_static = 1;      // every B operand of the field is a weight matrix or an activation of the forward launch
  return j;
}

// dX (B, k_in) = dZ (B, n_out) W with W (n_out, k_in) row-major: the B operand is consumed MN-major.
MegaJob make_dgrad_job(int64_t batch, int64_t k_in, int64_t n_out, const void* dz, int64_t ld_dz, const void* w,
                       int64_t ld_w, int wait_job) {
  MegaJob j = make_job(batch, k_in, n_out, dz, ld_dz, w, ld_w, wait_job, 0);
  j.b_mn = 1;
  j.bn = choose_bn(batch, k_in, n_out, 128);
  return j;
}

// dW (n_out, k_in) = dZ^T X with dZ (B, n_out) and X (B, k_in) both row-major: both operands MN-major.
MegaJob make_wgrad_job(int64_t n_out, int64_t k_in, int64_t batch, const void* dz, int64_t ld_dz, const void* x,
                       int64_t ld_x, int wait_job) {
  MegaJob j = make_job(n_out, k_in, batch, dz, ld_dz, x, ld_x, wait_job, 1);
  j.a_mn = 1; j.b_mn = 1;
  // A weight-gradient job is never on the dgrad chain, it fills the pairs the chain leaves idle -- and every pair it
  // occupies starts its next chain tile later.  Rows of >= 2048 columns take the widest tile (half as many tiles, less
  // operand traffic): dW3 (1024 x 2048) at 256 instead of 128 columns measured 6.5 us per step faster at B=2048, while
  // 256 columns for the 1024-wide rows (dW4, dW5) measured slower or equal (tools/ab_step.py).
  j.bn = k_in >= 2048 ? 256 : choose_bn(n_out, k_in, batch, 128);
  return j;
}

inline uint8_t* at(void* base, size_t off) { return reinterpret_cast<uint8_t*>(base) + off; }
inline const uint8_t* at(const void* base, size_t off) { return reinterpret_cast<const uint8_t*>(base) + off; }

// bf16 operand copies of trunk layers [first, L) and of the heads (+ their concatenated fp32 bias vector), one launch.
int pack_layers(const Layout& l, int first, const float* const* weights, const float* const* biases, void* pack,
                cudaStream_t stream) {
  PackList pl = {};
  for (int i = first; i < l.L; ++i) {
    PackMatrix& m = pl.m[pl.n++];
    m.in = weights[i] + (i == 0 ? l.G : 0);
    m.rows = l.n[i]; m.cols = l.k[i]; m.ld_in = i == 0 ? l.G + l.E : l.k[i];
    m.out = reinterpret_cast<__nv_bfloat16*>(at(pack, l.w[i])); m.ld_out = l.ldw[i];
    if (pl.n == 7) {                                   // keep one slot for the heads
      NERAF_TRY(pack_list(pl, stream));
      pl = PackList{};
    }
  }
  // heads are concatenated along N: every head is (F, W) with the same row stride, so they form one (C*F, W) matrix
  // in the pack; in the parameters they are C separate tensors -> one list entry per head would exceed 8 for C = 8,
  // so heads go in a list of their own when needed.
  if (pl.n + l.C > 8) {
    NERAF_TRY(pack_list(pl, stream));
    pl = PackList{};
  }
  for (int c = 0; c < l.C; ++c) {
    PackMatrix& m = pl.m[pl.n++];
    m.in = weights[l.L + c]; m.rows = l.F; m.cols = l.W; m.ld_in = l.W;
    m.out = reinterpret_cast<__nv_bfloat16*>(at(pack, l.wh)) + (size_t)c * l.F * l.ldwh; m.ld_out = l.ldwh;
    pl.copy_src[c] = biases[l.L + c];
  }
  pl.copy_dst = reinterpret_cast<float*>(at(pack, l.bh));
  pl.copy_width = l.F; pl.n_copy = l.CF;
  return pack_list(pl, stream);
}

int check_ptr_list(const float* const* list, int n, const char* what) {
  NERAF_REQUIRE(list, "field: %s array is null", what);
  for (int i = 0; i < n; ++i) NERAF_REQUIRE(list[i], "field: %s[%d] is null", what, i);
  return NERAF_OK;
}

}  // namespace

}  // namespace neraf

using namespace neraf;

extern "C" int neraf_field_sizes(const neraf_field_dims* dims, int precision, int64_t max_batch, size_t* pack_bytes,
                                 size_t* workspace_bytes) {
  Layout l;
  NERAF_TRY(make_layout(dims, precision, max_batch, &l));
  if (pack_bytes) *pack_bytes = l.pack_bytes;
  if (workspace_bytes) *workspace_bytes = l.ws_bytes;
  return NERAF_OK;
}

extern "C" int neraf_field_pack(const neraf_field_dims* dims, int precision, const float* const* weights,
                                const float* const* biases, void* pack, size_t pack_bytes, neraf_stream_t stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  Layout l;
  NERAF_TRY(make_layout(dims, precision, 0, &l));
  if (precision == NERAF_PREC_FP32) return NERAF_OK;
  NERAF_TRY(check_ptr_list(weights, l.L + l.C, "weights"));
  NERAF_TRY(check_ptr_list(biases, l.L + l.C, "biases"));
  NERAF_REQUIRE(pack, "field_pack: pack buffer is null");
  if (pack_bytes < l.pack_bytes)
    return set_error(NERAF_ERR_WORKSPACE, "field_pack: pack buffer %zu < %zu bytes", pack_bytes, l.pack_bytes);
  return pack_layers(l, 0, weights, biases, pack, stream);
}

// loss_gt / loss_sums (optional, together): the spectral loss's partial sums of `out` against the targets, formed by
// the heads' epilogue (bf16) or by loss_sums_kernel behind the GEMMs (fp32); loss_sums[0..4] is cleared first.
static int field_forward_impl(const neraf_field_dims* dims, int precision, const neraf_queries* q,
                              const float* grid_feature, const float* const* weights, const float* const* biases,
                              void* pack, size_t pack_bytes, int repack, void* ws, size_t ws_bytes, float* out,
                              int keep, const float* loss_gt, double* loss_sums, cudaStream_t stream) {
  NERAF_REQUIRE(q, "field_forward: queries is null");
  Layout l;
  NERAF_TRY(make_layout(dims, precision, q->batch, &l));
  const int64_t B = q->batch;
  if (B == 0) {
    if (loss_sums) NERAF_CHECK_CUDA(cudaMemsetAsync(loss_sums, 0, 5 * sizeof(double), stream));
    return NERAF_OK;
  }
  NERAF_TRY(check_ptr_list(weights, l.L + l.C, "weights"));
  NERAF_TRY(check_ptr_list(biases, l.L + l.C, "biases"));
  NERAF_REQUIRE(out && ws, "field_forward: out/workspace is null");
  NERAF_REQUIRE(l.G == 0 || grid_feature, "field_forward: grid_feature is null but n_grid = %d", l.G);
  if (ws_bytes < l.ws_bytes)
    return set_error(NERAF_ERR_WORKSPACE, "field_forward: workspace %zu < %zu bytes for batch %lld", ws_bytes,
                     l.ws_bytes, (long long)B);
  const bool bf = precision == NERAF_PREC_BF16;
  NERAF_REQUIRE(!bf || pack, "field_forward: pack is null");
  NERAF_REQUIRE(q->enc || l.E == 163, "field_forward: query encodings produce 163 columns, dims->n_enc = %d", l.E);
  NERAF_REQUIRE(!q->enc || q->enc_ld >= l.E, "field_forward: enc_ld < n_enc");
  if (bf && pack_bytes < l.pack_bytes)
    return set_error(NERAF_ERR_WORKSPACE, "field_forward: pack buffer %zu < %zu bytes", pack_bytes, l.pack_bytes);

  // effective layer-1 bias  c1 = b1 + W1[:, :G] g   (once per step, not per query)
  float* c1w = reinterpret_cast<float*>(at(ws, l.c1));
  const float* c1 = l.G > 0 ? c1w : biases[0];
  const int64_t ldw0 = l.G + l.E;

  if (!bf) {
    // the per-query block of h is kept in the workspace: backward reads it for dW1
    float* enc = reinterpret_cast<float*>(at(ws, l.enc));
    if (q->enc)
      NERAF_CHECK_CUDA(cudaMemcpy2DAsync(enc, (size_t)l.E * 4, q->enc, (size_t)q->enc_ld * 4, (size_t)l.E * 4, (size_t)B,
                                         cudaMemcpyDeviceToDevice, stream));
    NERAF_TRY(field_prep(q->enc ? nullptr : q, enc, l.E, nullptr, 0, l.E, weights[0], ldw0, biases[0], grid_feature, l.n[0],
                         l.G, c1w, l.E, nullptr, 0, stream, nullptr, reinterpret_cast<unsigned int*>(at(ws, l.counters)),
                         (int64_t)((l.gscratch + l.gscratch_bytes - l.counters) / 4)));
    const float* x = enc;
    int64_t ldx = l.E;
    for (int i = 0; i < l.L; ++i) {
      float* y = reinterpret_cast<float*>(at(ws, l.x[i]));
      const float* w = weights[i] + (i == 0 ? l.G : 0);
      const int64_t ldw = i == 0 ? ldw0 : l.k[i];
      NERAF_TRY(gemm_f32(B, l.n[i], l.k[i], x, ldx, 1, w, ldw, 1, i == 0 ? c1 : biases[i], NERAF_ACT_LEAKY, nullptr, 0, y,
                         l.n[i], 0, stream));
      x = y; ldx = l.n[i];
    }
    for (int c = 0; c < l.C; ++c)
      NERAF_TRY(gemm_f32(B, l.F, l.W, x, ldx, 1, weights[l.L + c], l.W, 1, biases[l.L + c], NERAF_ACT_TANH10, nullptr, 0,
                         out + (size_t)c * l.F, l.CF, 0, stream));
    if (loss_sums) {
      NERAF_CHECK_CUDA(cudaMemsetAsync(loss_sums, 0, 5 * sizeof(double), stream));
      if (loss_gt) NERAF_TRY(neraf_spectral_loss_sums(out, loss_gt, B * l.CF, loss_sums, 1, stream));
    }
    return NERAF_OK;
  }

  // ---- bf16 tensor-core path
  const bool do_pack = repack != 0;
  SideStream* side = do_pack ? side_stream() : nullptr;
  if (side) {                                        // layers 2.. are re-packed beside the encodings and layer 1
    NERAF_CHECK_CUDA(cudaEventRecord(side->fork, stream));
    NERAF_CHECK_CUDA(cudaStreamWaitEvent(side->stream, side->fork, 0));
    NERAF_TRY(pack_layers(l, 1, weights, biases, pack, side->stream));
    NERAF_CHECK_CUDA(cudaEventRecord(side->done, side->stream));
  } else if (do_pack) {
    NERAF_TRY(pack_layers(l, 1, weights, biases, pack, stream));
  }
  void* enc = at(ws, l.enc);
  if (q->enc) NERAF_TRY(convert_bf16(q->enc, B, l.E, q->enc_ld, enc, l.ld_enc, nullptr, 0, stream));
  NERAF_TRY(field_prep(q->enc ? nullptr : q, nullptr, 0, enc, l.ld_enc, (int)l.ld_enc, weights[0], ldw0, biases[0],
                       grid_feature, l.n[0], l.G, c1w, l.E, do_pack ? at(pack, l.w[0]) : nullptr, l.ldw[0], stream,
                       loss_sums, reinterpret_cast<unsigned int*>(at(ws, l.counters)),
                       (int64_t)((l.gscratch + l.gscratch_bytes - l.counters) / 4)));

  // One persistent launch for the whole MLP (two when the operands are being re-packed on the helper stream:
  // layer 1 starts as soon as its own copy exists, the remaining layers once the helper stream has finished).
  MegaJob jobs[NERAF_MAX_TRUNK + 1];
  const void* xin = enc;
  int64_t ldin = l.ld_enc;
  for (int i = 0; i <= l.L; ++i) {
    const bool head = i == l.L;
    MegaJob& j = jobs[i];
    j = make_job(B, head ? l.CF : l.n[i], head ? l.W : l.k[i], xin, ldin, head ? at(pack, l.wh) : at(pack, l.w[i]),
                 head ? l.ldwh : l.ldw[i], i - 1, 0);
    if (head) {
      j.epi.bias = reinterpret_cast<const float*>(at(pack, l.bh));
      j.epi.act = NERAF_ACT_TANH10;
      j.epi.out_f32 = out; j.epi.ld_f32 = l.CF;
      j.epi.loss_gt = loss_gt; j.epi.ld_gt = l.CF; j.epi.loss_sums = loss_sums;
    } else {
      j.epi.bias = i == 0 ? c1 : biases[i];
      j.epi.act = NERAF_ACT_LEAKY;
      j.epi.out_bf16 = at(ws, l.x[i]); j.epi.ld_bf16 = l.ldx[i];
      if (keep) { j.epi.mask_out = at(ws, l.mask[i]); j.epi.ld_mask = l.ld_mask; }   // LeakyReLU' gate of the backward pass
      xin = j.epi.out_bf16; ldin = l.ldx[i];
    }
  }
  // ONE persistent launch for the whole MLP: row blocks of layer 2 start while layer 1 is still finishing others.
  // (An earlier version launched layer 1 on its own so that it could start before the helper stream had re-packed
  // the other layers; the second launch's fixed cost and the lost layer-1 / layer-2 overlap cost more than the
  // few microseconds of waiting for the pack.)
  if (side) NERAF_CHECK_CUDA(cudaStreamWaitEvent(stream, side->done, 0));
  // the counters were cleared by field_prep_kernel above, and every job-list launch clears them again before it exits:
  // neither this launch nor the backward's needs a memset node
  // (programmatic dependent launch only when the launch has ONE predecessor, the prep kernel: not beside the helper stream)
  NERAF_TRY(mega_run(jobs, l.L + 1, at(ws, l.counters), l.counters_bytes, stream, 0, true, side == nullptr));
  return NERAF_OK;
}

extern "C" int neraf_field_forward(const neraf_field_dims* dims, int precision, const neraf_queries* q,
                                   const float* grid_feature, const float* const* weights, const float* const* biases,
                                   void* pack, size_t pack_bytes, int repack, void* ws, size_t ws_bytes, float* out,
                                   int keep, neraf_stream_t stream) {
  return field_forward_impl(dims, precision, q, grid_feature, weights, biases, pack, pack_bytes, repack, ws, ws_bytes, out,
                            keep, nullptr, nullptr, (cudaStream_t)stream);
}

extern "C" int neraf_field_forward_loss_sums(const neraf_field_dims* dims, int precision, const neraf_queries* q,
                                             const float* grid_feature, const float* const* weights,
                                             const float* const* biases, void* pack, size_t pack_bytes, int repack,
                                             void* ws, size_t ws_bytes, float* out, int keep, const float* gt,
                                             double* sums, neraf_stream_t stream) {
  NERAF_REQUIRE(sums, "field_forward_loss_sums: sums is null");
  return field_forward_impl(dims, precision, q, grid_feature, weights, biases, pack, pack_bytes, repack, ws, ws_bytes, out,
                            keep, gt, sums, (cudaStream_t)stream);
}

static int field_backward_impl(const neraf_field_dims* dims, int precision, int64_t B, const float* dout,
                               const float* out, const float* grid_feature, const float* const* weights,
                               const void* pack, void* ws, size_t ws_bytes, float* const* dweights,
                               float* const* dbiases, float* dgrid, float* denc, int64_t denc_ld,
                               const neraf_dp_options* opt, cudaStream_t stream) {
  const neraf_multicast* mc = opt ? opt->mc : nullptr;
  float* dw0_compact = opt ? opt->dw0_compact : nullptr;
  const int defer_grid_grads = opt ? opt->defer_grid_grads : 0;
  int phase = opt ? opt->phase : 0;
  const int max_ctas = opt ? opt->max_ctas : 0;
  const neraf_loss_grad* loss = opt ? opt->loss : nullptr;
  const bool zero_tail_slack = opt && opt->zero_tail_slack;
  uint32_t* notify = opt ? opt->notify : nullptr;
  void* const* dw16 = opt ? opt->dweights_bf16 : nullptr;
  NERAF_REQUIRE(phase >= 0 && phase <= 2, "field_backward_dp: phase must be 0, 1 or 2");
  Layout l;
  NERAF_TRY(make_layout(dims, precision, B, &l));
  if (B == 0) return NERAF_OK;
  NERAF_TRY(check_ptr_list(weights, l.L + l.C, "weights"));
  NERAF_TRY(check_ptr_list(dweights, l.L + l.C, "dweights"));
  NERAF_TRY(check_ptr_list(dbiases, l.L + l.C, "dbiases"));
  NERAF_REQUIRE((dout || loss) && out && ws, "field_backward: dout/out/workspace is null");
  NERAF_REQUIRE(l.G == 0 || grid_feature, "field_backward: grid_feature is null but n_grid = %d", l.G);
  NERAF_REQUIRE(!denc || denc_ld >= l.E, "field_backward: denc_ld < n_enc");
  if (ws_bytes < l.ws_bytes)
    return set_error(NERAF_ERR_WORKSPACE, "field_backward: workspace %zu < %zu bytes for batch %lld", ws_bytes,
                     l.ws_bytes, (long long)B);
  const bool bf = precision == NERAF_PREC_BF16;
  NERAF_REQUIRE(!bf || pack, "field_backward: pack is null");
  const int last = l.L - 1;
  const int64_t ldw0 = l.G + l.E;
  if (l.L < 2) {                                     // nothing to split with a single trunk layer
    if (phase == 2) return NERAF_OK;
    phase = 0;
  }

  NERAF_REQUIRE(!dw0_compact || defer_grid_grads, "field_backward_dp: dw0_compact needs defer_grid_grads");
  NERAF_REQUIRE(!dw16 || (defer_grid_grads && !mc && precision == NERAF_PREC_BF16),
                "field_backward_dp: dweights_bf16 needs the bf16 path with defer_grid_grads and no multicast");
  if (dw16) {
    for (int i = 0; i < l.L + l.C; ++i)
      NERAF_REQUIRE(dw16[i] && ((uintptr_t)dw16[i] & 15) == 0, "field_backward_dp: dweights_bf16[%d] is null or misaligned", i);
    for (int c = 1; c < l.C; ++c)
      NERAF_REQUIRE((uint8_t*)dw16[l.L + c] == (uint8_t*)dw16[l.L] + (size_t)c * l.F * l.W * 2,
                    "field_backward_dp: bf16 head gradients must be contiguous");
    for (int i = 1; i < l.L; ++i)
      NERAF_REQUIRE(l.k[i] % 8 == 0, "field_backward_dp: dweights_bf16 needs layer widths that are multiples of 8");
    NERAF_REQUIRE(l.W % 8 == 0, "field_backward_dp: dweights_bf16 needs a feature width that is a multiple of 8");
  }
  if (!bf && (mc || defer_grid_grads || phase))
    return set_error(NERAF_ERR_UNSUPPORTED, "field_backward_dp: the fused all-reduce exists for the bf16 path only");
  if (!bf) {
    float* dzh = reinterpret_cast<float*>(at(ws, l.dzh));
    NERAF_TRY(head_backward(dout, out, B, l.CF, dzh, l.CF, nullptr, 0, nullptr, 0, stream, loss));
    const float* x_last = reinterpret_cast<const float*>(at(ws, l.x[last]));
    float* dz_last = reinterpret_cast<float*>(at(ws, l.dz[last]));
    for (int c = 0; c < l.C; ++c) {
      const float* dzc = dzh + (size_t)c * l.F;
      NERAF_TRY(colsum_f32(dzc, B, l.F, l.CF, dbiases[l.L + c], stream));
      // dW_head[f, w] = sum_b dzc[b, f] x_last[b, w]
      NERAF_TRY(gemm_f32(l.F, l.W, B, dzc, 1, l.CF, x_last, 1, l.W, nullptr, NERAF_ACT_NONE, nullptr, 0, dweights[l.L + c],
                         l.W, 0, stream));
      // dZ_last = (sum_c dzc W_c) * leaky'(x_last); the gate is applied by the call that adds the last head
      NERAF_TRY(gemm_f32(B, l.W, l.F, dzc, l.CF, 1, weights[l.L + c], 1, l.W, nullptr, NERAF_ACT_NONE,
                         c == l.C - 1 ? x_last : nullptr, l.W, dz_last, l.W, c > 0, stream));
    }
    for (int i = last; i >= 0; --i) {
      const float* dz = reinterpret_cast<const float*>(at(ws, l.dz[i]));
      NERAF_TRY(colsum_f32(dz, B, l.n[i], l.n[i], dbiases[i], stream));
      if (i > 0) {
        const float* xin = reinterpret_cast<const float*>(at(ws, l.x[i - 1]));
        NERAF_TRY(gemm_f32(l.n[i], l.k[i], B, dz, 1, l.n[i], xin, 1, l.k[i], nullptr, NERAF_ACT_NONE, nullptr, 0,
                           dweights[i], l.k[i], 0, stream));
        NERAF_TRY(gemm_f32(B, l.k[i], l.n[i], dz, l.n[i], 1, weights[i], 1, l.k[i], nullptr, NERAF_ACT_NONE, xin, l.k[i],
                           reinterpret_cast<float*>(at(ws, l.dz[i - 1])), l.k[i], 0, stream));
      } else {
        const float* enc = reinterpret_cast<const float*>(at(ws, l.enc));
        NERAF_TRY(gemm_f32(l.n[0], l.E, B, dz, 1, l.n[0], enc, 1, l.E, nullptr, NERAF_ACT_NONE, nullptr, 0,
                           dweights[0] + l.G, ldw0, 0, stream));
        if (l.G > 0)
          NERAF_TRY(grid_grads(dbiases[0], grid_feature, weights[0], ldw0, l.n[0], l.G, dweights[0], dgrid, at(ws, l.gscratch), stream));
        if (denc)
          NERAF_TRY(gemm_f32(B, l.E, l.n[0], dz, l.n[0], 1, weights[0] + l.G, 1, ldw0, nullptr, NERAF_ACT_NONE, nullptr, 0,
                             denc, denc_ld, 0, stream));
      }
    }
    return NERAF_OK;
  }

  // ---- bf16 tensor-core path: the dgrad chain and all weight-gradient GEMMs in ONE launch.
  // Bias gradients are column sums of the fp32 dZ accumulated by the epilogues (atomics on zeroed buffers): the
  // buffers are zeroed with as few memsets as their addresses allow -- one when the caller laid them out back to back
  // (neraf_b200/field.py does); dgrid is overwritten by grid_grads.
  float* zero_start = nullptr;                       // one 16-byte-tileable span: cleared by the fused loss kernel instead
  int64_t zero_n = 0;
  if (phase != 2) {
    const int nb = l.L + l.C;
    struct Span { float* a; float* b; };
    Span spans[NERAF_MAX_TRUNK + 8 + 1];
    int ns = 0;
    int i = 0;
    while (i < nb) {
      float* start = dbiases[i];
      float* end = start + (i < l.L ? l.n[i] : l.F);
      int j = i + 1;
      while (j < nb && dbiases[j] == end) { end += j < l.L ? l.n[j] : l.F; ++j; }
      spans[ns++] = Span{start, end};
      i = j;
    }
    const bool fused_clear = loss && loss->fuse_sums && ns == 1 && ((uintptr_t)spans[0].a & 15) == 0 &&
                             ((spans[0].b - spans[0].a) % 4 == 0 || zero_tail_slack);
    if (fused_clear) {
      zero_start = spans[0].a;
      zero_n = round_up(spans[0].b - spans[0].a, 4);
    } else {
      for (int k = 0; k < ns; ++k)
        NERAF_CHECK_CUDA(cudaMemsetAsync(spans[k].a, 0, (size_t)(spans[k].b - spans[k].a) * 4, stream));
    }
  }
  void* dzh = at(ws, l.dzh);
  if (phase != 2)
    NERAF_TRY(head_backward(dout, out, B, l.CF, nullptr, 0, dzh, l.ld_h, dbiases + l.L, l.F, stream, loss, zero_start, zero_n));
  bool heads_contiguous = true;                      // the C head gradients form one (C*F, W) matrix?
  for (int c = 1; c < l.C; ++c) heads_contiguous = heads_contiguous && dweights[l.L + c] == dweights[l.L] + (size_t)c * l.F * l.W;
  // Fused all-reduce (data parallel): weight gradients are not stored but added into every rank's copy of the
  // caller's symmetric gradient buffer through its multicast alias (NVLS multimem.red in the GEMM epilogue).
  auto mc_alias = [&](float* p, void** out_mc) -> int {
    *out_mc = nullptr;
    if (!mc) return NERAF_OK;
    const uintptr_t lo = (uintptr_t)mc->local_base, a = (uintptr_t)p;
    NERAF_REQUIRE(mc->multicast_base && a >= lo && a < lo + mc->bytes,
                  "field_backward_dp: a weight gradient lies outside the symmetric buffer");
    *out_mc = reinterpret_cast<uint8_t*>(mc->multicast_base) + (a - lo);
    return NERAF_OK;
  };
  NERAF_REQUIRE(!mc || heads_contiguous, "field_backward_dp: head gradients must be contiguous");
  MegaJob jobs[NERAF_MEGA_MAX_JOBS];
  int notify_slot[NERAF_MEGA_MAX_JOBS];               // which completion counter of neraf_dp_options.notify a job advances
  for (int i = 0; i < NERAF_MEGA_MAX_JOBS; ++i) notify_slot[i] = -1;
  int nj = 0;
  // phase 1 stops before the last dgrad (dZ of layer 1) and the per-query block of dW1; phase 2 is exactly those:
  // the caller all-reduces what phase 1 finished while phase 2 computes (neraf_dp_options).
  int producer = nj;
  if (phase != 2) {                                    // dZ_last = (dZ_head W_head) * leaky'(x_last)
    MegaJob& j = jobs[nj++];
    j = make_dgrad_job(B, l.W, l.CF, dzh, l.ld_h, at(pack, l.wh), l.ldwh, -1);
    j.epi.gate_mask = at(ws, l.mask[last]); j.epi.ld_mask = l.ld_mask;
    j.epi.out_bf16 = at(ws, l.dz[last]); j.epi.ld_bf16 = l.ldx[last];
    j.colsum = dbiases[last];
  }
  if (phase != 2) {                                    // head weight gradients (all heads in one GEMM)
    notify_slot[nj] = l.L;
    MegaJob& j = jobs[nj++];
    j = make_wgrad_job(l.CF, l.W, B, dzh, l.ld_h, at(ws, l.x[last]), l.ldx[last], -1);
    j.wait_all = 0;
    if (dw16) {
      j.epi.out_bf16 = dw16[l.L]; j.epi.ld_bf16 = l.W;
    } else {
      j.epi.out_f32 = heads_contiguous ? dweights[l.L] : reinterpret_cast<float*>(at(ws, l.dwh));
      j.epi.ld_f32 = l.W;
      NERAF_TRY(mc_alias(j.epi.out_f32, &j.epi.out_f32_multicast));
    }
  }
  // With the gradient exchange running beside this launch, the largest weight gradient (dW of trunk layer 1: two thirds
  // of the bytes) is scheduled BEFORE the dgrad job of the same step instead of behind it: it is then complete -- and
  // travelling -- while the last dgrad job and the compact dW1 block are computed, instead of finishing with the launch.
  const bool exchange_beside = opt && opt->exchange;
  int status = NERAF_OK;
  for (int i = last; i >= 0; --i) {
    if (phase == 2 && i > 1) continue;
    if (phase == 1 && i == 0) break;
    const int dz_producer = (phase == 2 && i == 1) ? -1 : producer;   // job that writes dZ_i (phase 2: an earlier launch)
    auto add_dgrad = [&]() {                           // chain: dZ_{i-1} = (dZ_i W_i) * leaky'(x_{i-1})
      if (!(i > 0 && !(phase == 1 && i == 1))) return;
      producer = nj;
      MegaJob& j = jobs[nj++];
      j = make_dgrad_job(B, l.k[i], l.n[i], at(ws, l.dz[i]), l.ldx[i], at(pack, l.w[i]), l.ldw[i], dz_producer);
      j.epi.gate_mask = at(ws, l.mask[i - 1]); j.epi.ld_mask = l.ld_mask;
      j.epi.out_bf16 = at(ws, l.dz[i - 1]); j.epi.ld_bf16 = l.ldx[i - 1];
      j.colsum = dbiases[i - 1];
      // Fused all-reduce: the tiles of dW_i push their results over NVLink, the tiles of this dgrad job do not, and
      // both only need dZ_i -- interleaved, the link drains behind the dgrad tiles instead of throttling every CTA
      // pair at once (the largest pair, dgrad 2->1 / dW_2, is 70 % of the bytes and sits at the end of the backward).
      if (mc && phase == 0) j.merge_next = 1;
    };
    auto add_wgrad = [&]() {                           // dW_i = dZ_i^T x_{i-1}: needs every row block of dZ_i
      if (phase == 2 && i == 1) return;                // dW_1 belongs to phase 1
      notify_slot[nj] = i;
      MegaJob& w = jobs[nj++];
      if (i > 0) {
        w = make_wgrad_job(l.n[i], l.k[i], B, at(ws, l.dz[i]), l.ldx[i], at(ws, l.x[i - 1]), l.ldx[i - 1], dz_producer);
        if (dw16) {
          w.epi.out_bf16 = dw16[i]; w.epi.ld_bf16 = l.k[i];
        } else {
          w.epi.out_f32 = dweights[i]; w.epi.ld_f32 = l.k[i];
          if (status == NERAF_OK) status = mc_alias(w.epi.out_f32, &w.epi.out_f32_multicast);
        }
      } else {
        w = make_wgrad_job(l.n[0], l.E, B, at(ws, l.dz[0]), l.ldx[0], at(ws, l.enc), l.ld_enc, dz_producer);
        w.epi.out_f32 = dweights[0] + l.G; w.epi.ld_f32 = ldw0;
        if (dw0_compact) { w.epi.out_f32 = dw0_compact; w.epi.ld_f32 = round_up(l.E, 8); }   // (n_1, E) with 32-byte rows
        if (dw16) { w.epi.out_f32 = nullptr; w.epi.out_bf16 = dw16[0]; w.epi.ld_bf16 = round_up(l.E, 8); }
        if (status == NERAF_OK) status = mc_alias(w.epi.out_f32, &w.epi.out_f32_multicast);
        if (denc) {
          MegaJob& e = jobs[nj++];
          e = make_dgrad_job(B, l.E, l.n[0], at(ws, l.dz[0]), l.ldx[0], at(pack, l.w[0]), l.ldw[0], dz_producer);
          e.epi.out_f32 = denc; e.epi.ld_f32 = denc_ld;
        }
      }
    };
    if (exchange_beside && phase == 0 && i == 1 && !mc) { add_wgrad(); add_dgrad(); }
    else { add_dgrad(); add_wgrad(); }
  }
  NERAF_TRY(status);
  // Completion counters for a concurrent consumer (the data-parallel gradient exchange, neraf_dp_exchange_grads): one per
  // weight gradient, one for "every bias gradient is final" = the last job of the dgrad chain (its row blocks need all
  // row blocks of every earlier link; the heads' bias gradients were complete before this launch).
  unsigned int increments[NERAF_MEGA_MAX_JOBS] = {0};
  unsigned int slot_off[NERAF_MAX_TRUNK + 2] = {0}, slot_cnt[NERAF_MAX_TRUNK + 2] = {0}, slot_inc[NERAF_MAX_TRUNK + 2] = {0};
  if (notify) {
    NERAF_REQUIRE(phase == 0, "field_backward_dp: completion counters need the one-launch backward (phase 0)");
    notify_slot[producer] = l.L + 1;
    // canonical layout (a binding can compute it): matrix k owns ceil(rows_k / 256) counters behind those of matrix k - 1
    unsigned int next = 0;
    for (int k = 0; k < l.L + 2; ++k) {
      const int64_t rows = k < l.L ? l.n[k] : (k == l.L ? l.CF : B);
      slot_off[k] = next;
      slot_cnt[k] = (unsigned int)ceil_div(rows, 256);
      next += slot_cnt[k];
    }
    for (int j = 0; j < nj; ++j) {
      if (notify_slot[j] < 0) continue;
      const int k = notify_slot[j];
      NERAF_REQUIRE((unsigned int)ceil_div(jobs[j].M, 256) == slot_cnt[k], "field_backward_dp: counter layout mismatch (job %d)", j);
      jobs[j].notify = notify + slot_off[k];
    }
    NERAF_REQUIRE(next <= NERAF_NOTIFY_COUNTERS, "field_backward_dp: %u completion counters needed, %d available", next,
                  NERAF_NOTIFY_COUNTERS);
  }
  // Data parallel: the gradient exchange kernel runs BESIDE the job-list launch and follows its completion counters.  It
  // is launched right behind it in the same stream as a programmatic dependent that never waits: it moves in when every
  // CTA pair of the persistent GEMM grid is resident (they release their dependents at start), so its CTAs -- wherever
  // the block scheduler puts them -- can never keep a pair from becoming resident.  (Launched on a second stream, two of
  // its CTAs on one SM could: the pairs spin on each other's progress, the exchange spins on theirs -- a deadlock that
  // the watchdog turned into a launch failure, intermittently, on graph replays.)
  const neraf_grad_exchange* xg = opt ? opt->exchange : nullptr;
  NERAF_REQUIRE(!xg || notify, "field_backward_dp: exchange needs notify");
  // left clean by the forward's launch; set up under the tail of the head-gradient kernel that precedes it
  NERAF_TRY(mega_run(jobs, nj, at(ws, l.counters), l.counters_bytes, stream, max_ctas, true, phase != 2, increments,
                     xg != nullptr));
  if (notify) {
    for (int j = 0; j < nj; ++j)
      if (notify_slot[j] >= 0) slot_inc[notify_slot[j]] = increments[j];
    for (int k = 0; k < l.L + 2; ++k) {
      if (opt->notify_offset) opt->notify_offset[k] = slot_off[k];
      if (opt->notify_count) opt->notify_count[k] = slot_cnt[k];
      if (opt->notify_increment) opt->notify_increment[k] = slot_inc[k];
    }
  }
  if (xg) {
    neraf_grad_exchange x = *xg;                       // chunks that name counters of this call get their increment
    for (int c = 0; c < x.n_chunks && c < NERAF_MAX_EXCHANGE_CHUNKS; ++c) {
      const uint32_t* nf = x.chunks[c].notify;
      if (!nf || nf < notify || nf >= notify + NERAF_NOTIFY_COUNTERS) continue;
      const unsigned int idx = (unsigned int)(nf - notify);
      bool found = false;
      for (int k = 0; k < l.L + 2 && !found; ++k)
        if (idx >= slot_off[k] && idx + x.chunks[c].notify_count <= slot_off[k] + slot_cnt[k]) {
          x.chunks[c].notify_increment = slot_inc[k];
          found = true;
        }
      NERAF_REQUIRE(found, "field_backward_dp: chunk %d of the exchange names counters of no single gradient matrix", c);
    }
    NERAF_TRY(dp_exchange_grads(&x, stream, true));
  }
  if (!heads_contiguous && phase != 2 && !dw16)
    for (int c = 0; c < l.C; ++c)
      NERAF_CHECK_CUDA(cudaMemcpyAsync(dweights[l.L + c], at(ws, l.dwh) + (size_t)c * l.F * l.W * 4, (size_t)l.F * l.W * 4,
                                       cudaMemcpyDeviceToDevice, stream));
  if (l.G > 0 && !defer_grid_grads && phase != 1)
    NERAF_TRY(grid_grads(dbiases[0], grid_feature, weights[0], ldw0, l.n[0], l.G, dweights[0], dgrid, at(ws, l.gscratch), stream));
  return NERAF_OK;
}

extern "C" int neraf_field_backward(const neraf_field_dims* dims, int precision, int64_t B, const float* dout,
                                    const float* out, const float* grid_feature, const float* const* weights,
                                    const void* pack, void* ws, size_t ws_bytes, float* const* dweights,
                                    float* const* dbiases, float* dgrid, float* denc, int64_t denc_ld,
                                    neraf_stream_t stream) {
  return field_backward_impl(dims, precision, B, dout, out, grid_feature, weights, pack, ws, ws_bytes, dweights, dbiases,
                             dgrid, denc, denc_ld, nullptr, (cudaStream_t)stream);
}

extern "C" int neraf_field_backward_dp(const neraf_field_dims* dims, int precision, int64_t B, const float* dout,
                                       const float* out, const float* grid_feature, const float* const* weights,
                                       const void* pack, void* ws, size_t ws_bytes, float* const* dweights,
                                       float* const* dbiases, float* dgrid, float* denc, int64_t denc_ld,
                                       const neraf_dp_options* opt, neraf_stream_t stream) {
  return field_backward_impl(dims, precision, B, dout, out, grid_feature, weights, pack, ws, ws_bytes, dweights, dbiases,
                             dgrid, denc, denc_ld, opt, (cudaStream_t)stream);
}

extern "C" int neraf_field_grid_grads(const neraf_field_dims* dims, const float* grid_feature, const float* weight0,
                                      const float* dbias0, const void* dw0_compact, int32_t compact_bf16, float* dweight0,
                                      float* dgrid, void* scratch, const void* widen_src_bf16, float* widen_dst,
                                      int64_t widen_n, neraf_stream_t stream) {
  NERAF_REQUIRE(dims && dims->n_trunk >= 1, "field_grid_grads: dims is null");
  if (dims->n_grid <= 0) return NERAF_OK;
  NERAF_REQUIRE(grid_feature && weight0 && dbias0 && (dweight0 || dgrid), "field_grid_grads: null pointer");
  NERAF_REQUIRE(!dgrid || scratch, "field_grid_grads: dgrid needs scratch");
  // the widening (28.7 MB -> 57.4 MB, no dependence on db1) runs on the helper stream beside the grid-gradient kernel
  SideStream* side = (widen_n > 0 && !getenv("NERAF_WIDEN_FUSED")) ? side_stream() : nullptr;
  if (side) {
    NERAF_CHECK_CUDA(cudaEventRecord(side->fork, (cudaStream_t)stream));
    NERAF_CHECK_CUDA(cudaStreamWaitEvent(side->stream, side->fork, 0));
    NERAF_TRY(widen_bf16(widen_src_bf16, widen_dst, widen_n, side->stream));
    NERAF_CHECK_CUDA(cudaEventRecord(side->done, side->stream));
  }
  NERAF_TRY(grid_grads(dbias0, grid_feature, weight0, (int64_t)dims->n_grid + dims->n_enc, dims->trunk[0], dims->n_grid, dweight0,
                       dgrid, scratch, (cudaStream_t)stream, dw0_compact, dims->n_enc, round_up(dims->n_enc, 8), compact_bf16 != 0,
                       side ? nullptr : widen_src_bf16, side ? nullptr : widen_dst, side ? 0 : widen_n));
  if (side) NERAF_CHECK_CUDA(cudaStreamWaitEvent((cudaStream_t)stream, side->done, 0));
  return NERAF_OK;
}
