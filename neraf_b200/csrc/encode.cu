// Query encodings of the acoustic field, fused into one kernel.
//
// Replaces the ~35 small kernels of NeRAFAudioModel.get_outputs before the MLP
// (/root/reference/NeRAF/NeRAF_model.py:533-551): time normalisation, AABB normalisation and
// whole-vector zeroing of out-of-box poses, NeRFEncoding x3 (float64 for positions, float32 for
// time, exactly like the reference's dtype flow) and the tiny-cuda-nn degree-4 spherical harmonics
// (float32 math, float16 rounding).  Writes the 163 per-query columns of h either as fp32 (parity
// path) or as bf16 in the two layouts the tcgen05 GEMMs consume (row-major K-major operand for the
// layer-1 GEMM, transposed copy for the layer-1 weight gradient).
#include <cuda_fp16.h>

#include "common.cuh"
#include "kernels.h"

namespace neraf {

// float32 values of 2 ** torch.linspace(0, 8, 10) (nerfstudio NeRFEncoding), bit-exact.
__constant__ float kFreqs[10] = {0x1.000000p+0f, 0x1.da0c40p+0f, 0x1.b6e8b0p+1f, 0x1.965fecp+2f, 0x1.784086p+3f,
                                 0x1.5c5cc0p+4f, 0x1.428a32p+5f, 0x1.2aa1a8p+6f, 0x1.147eccp+7f, 0x1.000000p+8f};

struct EncodeArgs {
  int64_t B;
  const int64_t* time_query;
  const double* mic;
  const double* src;
  const double* rot;
  const float* aabb;
  float time_den;
  int order;
  float* out_f32; int64_t ld_f32;
  __nv_bfloat16* out_bf16; int64_t ld_bf16;      // row-major (B, ld)
  __nv_bfloat16* out_bf16_t; int64_t ld_t;       // transposed (ncols_padded, ld_t)
  int ncols_padded;                               // columns [163, ncols_padded) are zero-filled
};

__device__ __forceinline__ void put(const EncodeArgs& a, int64_t b, int col, float v) {
  if (a.out_f32) a.out_f32[b * a.ld_f32 + col] = v;
  if (a.out_bf16) a.out_bf16[b * a.ld_bf16 + col] = __float2bfloat16_rn(v);
  if (a.out_bf16_t) a.out_bf16_t[(int64_t)col * a.ld_t + b] = __float2bfloat16_rn(v);
}

// One thread per (query, group, frequency): 80 threads per query.  Groups: 0 time, 1-3 mic xyz, 4-6 source xyz,
// 7 rot (spherical harmonics + padding, done by the k == 0 thread).  The k == 0 thread of a group also writes the
// raw input column (include_input=True appends it last).
__device__ __forceinline__ void encode_one(const EncodeArgs& a, int64_t gid) {
  const int64_t b = gid / 80;
  const int r = (int)(gid - b * 80);
  const int grp = r / 10, k = r - grp * 10;
  if (b >= a.B) return;
  const int off_time = a.order == NERAF_ORDER_TIME_MIC_SRC_ROT ? 0 : 126;
  const int off_mic = a.order == NERAF_ORDER_TIME_MIC_SRC_ROT ? 21 : 0;
  const int off_src = a.order == NERAF_ORDER_TIME_MIC_SRC_ROT ? 84 : 63;
  const int off_rot = 147;

  if (grp == 0) {
    // NeRAF_model.py:533-535 then NeRFEncoding(in_dim=1) entirely in float32.
    const float t = __fdiv_rn((float)a.time_query[b], a.time_den);
    const float scaled = __fmul_rn(0x1.921fb6p+2f, t);            // float32(2*pi) * t
    const float s = __fmul_rn(scaled, kFreqs[k]);
    const float c = __fadd_rn(s, 0x1.921fb6p+0f);                 // + float32(pi/2)
    put(a, b, off_time + k, (float)sin((double)s));
    put(a, b, off_time + 10 + k, (float)sin((double)c));
    if (k == 0) put(a, b, off_time + 20, t);
  } else if (grp <= 6) {
    const bool is_mic = grp <= 3;
    const int d = is_mic ? grp - 1 : grp - 4;
    const double* p = (is_mic ? a.mic : a.src) + b * 3;
    // SceneBox.get_normalized_positions: lengths formed in float32, promoted against float64 poses.
    double pn[3];
    bool inside = true;
#pragma unroll
    for (int i = 0; i < 3; ++i) {
      const float lo = a.aabb[i], hi = a.aabb[3 + i];
      const float len = __fsub_rn(hi, lo);
      pn[i] = (p[i] - (double)lo) / (double)len;
      inside = inside && (pn[i] > 0.0) && (pn[i] < 1.0);
    }
    const double x = inside ? pn[d] : pn[d] * 0.0;                // NeRAF_model.py:543-546 (whole vector)
    const double s = 6.283185307179586 * x * (double)kFreqs[k];
    const int base = is_mic ? off_mic : off_src;
    put(a, b, base + d * 10 + k, (float)sin(s));
    put(a, b, base + 30 + d * 10 + k, (float)sin(s + 1.5707963267948966));
    if (k == 0) put(a, b, base + 60 + d, (float)x);
  } else if (k == 0) {
    // tiny-cuda-nn SphericalHarmonics degree 4 on 2*rot-1; operation order == oracle/encodings.py sh4_tcnn.
    const float x = __fsub_rn(__fmul_rn((float)a.rot[b * 3 + 0], 2.f), 1.f);
    const float y = __fsub_rn(__fmul_rn((float)a.rot[b * 3 + 1], 2.f), 1.f);
    const float z = __fsub_rn(__fmul_rn((float)a.rot[b * 3 + 2], 2.f), 1.f);
    const float c0 = 0.28209479177387814f, c1 = 0.48860251190291987f, c2 = 1.0925484305920792f,
                c3 = 0.94617469575755997f, c3b = 0.31539156525251999f, c4 = 0.54627421529603959f,
                c5 = 0.59004358992664352f, c6 = 2.8906114426405538f, c7 = 0.45704579946446572f,
                c8 = 0.3731763325901154f, c9 = 1.4453057213202769f;
    const float xy = __fmul_rn(x, y), xz = __fmul_rn(x, z), yz = __fmul_rn(y, z);
    const float x2 = __fmul_rn(x, x), y2 = __fmul_rn(y, y), z2 = __fmul_rn(z, z);
    float o[16];
    o[0] = c0;
    o[1] = __fmul_rn(-c1, y);
    o[2] = __fmul_rn(c1, z);
    o[3] = __fmul_rn(-c1, x);
    o[4] = __fmul_rn(c2, xy);
    o[5] = __fmul_rn(-c2, yz);
    o[6] = __fsub_rn(__fmul_rn(c3, z2), c3b);
    o[7] = __fmul_rn(-c2, xz);
    o[8] = __fsub_rn(__fmul_rn(c4, x2), __fmul_rn(c4, y2));
    o[9] = __fmul_rn(__fmul_rn(c5, y), __fadd_rn(__fmul_rn(-3.f, x2), y2));
    o[10] = __fmul_rn(__fmul_rn(c6, xy), z);
    o[11] = __fmul_rn(__fmul_rn(c7, y), __fsub_rn(1.f, __fmul_rn(5.f, z2)));
    o[12] = __fmul_rn(__fmul_rn(c8, z), __fsub_rn(__fmul_rn(5.f, z2), 3.f));
    o[13] = __fmul_rn(__fmul_rn(c7, x), __fsub_rn(1.f, __fmul_rn(5.f, z2)));
    o[14] = __fmul_rn(__fmul_rn(c9, z), __fsub_rn(x2, y2));
    o[15] = __fmul_rn(__fmul_rn(c5, x), __fadd_rn(-x2, __fmul_rn(3.f, y2)));
#pragma unroll
    for (int i = 0; i < 16; ++i) put(a, b, off_rot + i, __half2float(__float2half_rn(o[i])));
    for (int c = 163; c < a.ncols_padded; ++c) put(a, b, c, 0.f);
  }
}

// What precedes layer 1 of a forward pass, in ONE launch (block ranges):
//   [0, enc_blocks)                      the query encodings above
//   [enc_blocks, +gb_blocks)             c1[n] = b1[n] + W1[n, :G] . g   (batch-invariant grid block of layer 1, one warp per row)
//   [.., +pk_blocks)                     bf16 operand copy of the per-query columns W1[:, G:] (only when re-packing)
struct PrepArgs {
  EncodeArgs enc;
  int enc_blocks, gb_blocks, pk_blocks;
  const float* W1; int64_t ldw; const float* b1; const float* g; int64_t n1, G; float* c1;
  int64_t E; __nv_bfloat16* w1_out; int64_t w1_ld;
  double* zero5;          // optional: five doubles cleared by block 0 (the loss sums a later launch of the step adds to)
  unsigned int* zero_u32; int64_t n_zero_u32;     // optional: words cleared by block 0 (the job-list kernel's counters)
};

__global__ void __launch_bounds__(256) field_prep_kernel(const PrepArgs p) {
  int blk = blockIdx.x;
  if (blk == 0) {
    if (threadIdx.x < 5 && p.zero5) p.zero5[threadIdx.x] = 0.0;
    for (int64_t i = threadIdx.x; i < p.n_zero_u32; i += 256) p.zero_u32[i] = 0u;
  }
  if (blk < p.enc_blocks) {
    encode_one(p.enc, (int64_t)blk * 256 + threadIdx.x);
    return;
  }
  blk -= p.enc_blocks;
  if (blk < p.gb_blocks) {
    const int64_t n = (int64_t)blk * 8 + threadIdx.x / 32;
    const int lane = threadIdx.x % 32;
    if (n >= p.n1) return;
    const float* w = p.W1 + n * p.ldw;
    // 32 independent loads per lane (1024 columns of the row) before the first use: with 4 in flight the row took 8
    // dependent round trips to a cold HBM (ncu: 24.5 MB in 15 us, DRAM 20 % busy); this way the whole 20.9 MB block of
    // W1 is in flight at once and the pass is bandwidth-bound
    float acc[4] = {0.f, 0.f, 0.f, 0.f};
    int64_t k = lane;
    for (; k + 31 * 32 < p.G; k += 1024) {
      float v[32];
#pragma unroll
      for (int u = 0; u < 32; ++u) v[u] = __ldg(w + k + 32 * u);
#pragma unroll
      for (int u = 0; u < 32; ++u) acc[u & 3] = fmaf(v[u], __ldg(p.g + k + 32 * u), acc[u & 3]);
    }
    for (; k + 96 < p.G; k += 128) {                              // 4 independent loads in flight per lane
#pragma unroll
      for (int u = 0; u < 4; ++u) acc[u] = fmaf(__ldg(w + k + 32 * u), __ldg(p.g + k + 32 * u), acc[u]);
    }
    for (; k < p.G; k += 32) acc[0] = fmaf(__ldg(w + k), __ldg(p.g + k), acc[0]);
    float t = (acc[0] + acc[1]) + (acc[2] + acc[3]);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) t += __shfl_xor_sync(0xffffffffu, t, o);
    if (lane == 0) p.c1[n] = t + (p.b1 ? p.b1[n] : 0.f);
    return;
  }
  blk -= p.gb_blocks;
  {                                                               // W1[:, G:G+E] -> bf16 (n1, w1_ld): 8 rows per block
    const int64_t n = (int64_t)blk * 8 + threadIdx.x / 32;
    const int lane = threadIdx.x % 32;
    if (n >= p.n1) return;
    const float* w = p.W1 + n * p.ldw + p.G;
    __nv_bfloat16* o = p.w1_out + n * p.w1_ld;
    for (int64_t k = lane; k < p.w1_ld; k += 32) o[k] = __float2bfloat16_rn(k < p.E ? __ldg(w + k) : 0.f);
  }
}

__global__ void __launch_bounds__(256) encode_kernel(EncodeArgs a) {
  encode_one(a, (int64_t)blockIdx.x * blockDim.x + threadIdx.x);
}

static int check_queries(const neraf_queries* q) {
  NERAF_REQUIRE(q && q->batch >= 0, "encode: bad query struct");
  if (q->batch == 0) return NERAF_OK;
  NERAF_REQUIRE(q->time_query && q->mic_pose && q->source_pose && q->rot && q->aabb,
                "encode: null query pointer");
  NERAF_REQUIRE(q->time_denominator != 0.f, "encode: time_denominator must be max_len - 1 != 0");
  return NERAF_OK;
}

int encode_queries(const neraf_queries* q, float* out_f32, int64_t ld_f32, void* out_bf16, int64_t ld_bf16,
                   void* out_bf16_t, int64_t ld_t, int ncols_padded, cudaStream_t stream) {
  NERAF_TRY(check_queries(q));
  if (q->batch == 0) return NERAF_OK;
  EncodeArgs a{q->batch, q->time_query, q->mic_pose, q->source_pose, q->rot, q->aabb, q->time_denominator,
               q->order, out_f32, ld_f32, (__nv_bfloat16*)out_bf16, ld_bf16, (__nv_bfloat16*)out_bf16_t, ld_t,
               ncols_padded};
  encode_kernel<<<(unsigned)ceil_div(q->batch * 80, 256), 256, 0, stream>>>(a);
  NERAF_CHECK_LAUNCH("encode_kernel");
  return NERAF_OK;
}

int field_prep(const neraf_queries* q, float* enc_f32, int64_t ld_f32, void* enc_bf16, int64_t ld_bf16, int ncols_padded,
               const float* W1, int64_t ldw, const float* b1, const float* g, int64_t n1, int64_t G, float* c1,
               int64_t E, void* w1_out, int64_t w1_ld, cudaStream_t stream, double* zero5, unsigned int* zero_u32,
               int64_t n_zero_u32) {
  PrepArgs p = {};
  p.zero5 = zero5;
  p.zero_u32 = zero_u32; p.n_zero_u32 = zero_u32 ? n_zero_u32 : 0;
  if (q) {
    NERAF_TRY(check_queries(q));
    p.enc = EncodeArgs{q->batch, q->time_query, q->mic_pose, q->source_pose, q->rot, q->aabb, q->time_denominator,
                       q->order, enc_f32, ld_f32, (__nv_bfloat16*)enc_bf16, ld_bf16, nullptr, 0, ncols_padded};
    p.enc_blocks = (int)ceil_div(q->batch * 80, 256);
  }
  p.W1 = W1; p.ldw = ldw; p.b1 = b1; p.g = g; p.n1 = n1; p.G = G; p.c1 = c1;
  p.gb_blocks = (G > 0 && c1) ? (int)ceil_div(n1, 8) : 0;
  p.E = E; p.w1_out = (__nv_bfloat16*)w1_out; p.w1_ld = w1_ld;
  p.pk_blocks = w1_out ? (int)ceil_div(n1, 8) : 0;
  const int blocks = p.enc_blocks + p.gb_blocks + p.pk_blocks;
  if (blocks == 0) {
    if (zero5) NERAF_CHECK_CUDA(cudaMemsetAsync(zero5, 0, 5 * sizeof(double), stream));
    if (p.n_zero_u32 > 0) NERAF_CHECK_CUDA(cudaMemsetAsync(zero_u32, 0, (size_t)p.n_zero_u32 * 4, stream));
    return NERAF_OK;
  }
  field_prep_kernel<<<(unsigned)blocks, 256, 0, stream>>>(p);
  NERAF_CHECK_LAUNCH("field_prep_kernel");
  return NERAF_OK;
}

}  // namespace neraf

extern "C" int neraf_encode_queries(const neraf_queries* q, float* enc_out, int64_t ld, neraf_stream_t stream) {
  NERAF_REQUIRE(q, "neraf_encode_queries: queries is null");
  if (q->batch == 0) return NERAF_OK;
  NERAF_REQUIRE(enc_out && ld >= 163, "neraf_encode_queries: enc_out null or ld < 163");
  return neraf::encode_queries(q, enc_out, ld, nullptr, 0, nullptr, 0, 163, (cudaStream_t)stream);
}
