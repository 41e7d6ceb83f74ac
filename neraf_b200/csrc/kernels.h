// Internal (non-ABI) launch helpers shared between translation units.
#pragma once
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/neraf_b200.h"

namespace neraf {

// encode.cu
int encode_queries(const neraf_queries* q, float* out_f32, int64_t ld_f32, void* out_bf16, int64_t ld_bf16,
                   void* out_bf16_t, int64_t ld_t, int ncols_padded, cudaStream_t stream);

// What precedes layer 1 of a forward pass in one launch: query encodings (q may be null: skipped), the effective
// layer-1 bias c1 = b1 + W1[:, :G] g (skipped when G == 0 or c1 is null) and the bf16 copy of W1[:, G:G+E]
// (skipped when w1_out is null).
int field_prep(const neraf_queries* q, float* enc_f32, int64_t ld_f32, void* enc_bf16, int64_t ld_bf16, int ncols_padded,
               const float* W1, int64_t ldw, const float* b1, const float* g, int64_t n1, int64_t G, float* c1,
               int64_t E, void* w1_out, int64_t w1_ld, cudaStream_t stream, double* zero5 = nullptr,
               unsigned int* zero_u32 = nullptr, int64_t n_zero_u32 = 0);

// gemm_simt.cu
int gemm_f32(int64_t M, int64_t N, int64_t K, const float* A, int64_t a_rs, int64_t a_cs, const float* B, int64_t b_rs,
             int64_t b_cs, const float* bias, int act, const float* gate, int64_t ldg, float* C, int64_t ldc,
             int accumulate, cudaStream_t stream);

// gemm_umma.cu
int gemm_bf16(int64_t M, int64_t N, int64_t K, const void* A, int64_t lda, const void* B, int64_t ldb,
              const neraf_gemm_epilogue* epi, cudaStream_t stream);

// gemm_mega.cu: a list of dependent bf16 GEMMs (same TN form and epilogue as gemm_bf16) executed by ONE
// persistent launch with tile-level scheduling and row-block dependency tracking.
#define NERAF_MEGA_MAX_JOBS NERAF_MAX_GEMM_JOBS
typedef neraf_gemm_job MegaJob;      // public contract: include/neraf_b200.h
// max_ctas: 0 = one CTA per SM; otherwise an upper bound on the grid (a concurrent kernel gets the other SMs)
// counters_clean: the caller guarantees the counter buffer is zero (cleared explicitly, or last used by this kernel,
// which clears what it used before it exits): no memset node is issued
// pdl: launch with programmatic stream serialization (the previous work of the stream must be a kernel of this library
// that executes griddepcontrol.launch_dependents, or any kernel: the prologue then simply starts when that one ends)
// notify_increment (optional, host, n_jobs entries): by how much one launch advances each job's `notify` counter
// release_dependents_early: griddepcontrol.launch_dependents as soon as every CTA of the grid is resident -- for a kernel
// launched behind this one with programmatic stream serialization that never waits for it (griddepcontrol.wait) but
// runs BESIDE it: it moves in only once this persistent grid holds all its SM slots, so it can never keep a CTA pair
// of this grid from becoming resident (the pairs spin on each other's progress)
int mega_run(const MegaJob* jobs, int n_jobs, void* counters, size_t counters_bytes, cudaStream_t stream, int max_ctas = 0,
             bool counters_clean = false, bool pdl = false, unsigned int* notify_increment = nullptr,
             bool release_dependents_early = false);
// comm.cu: neraf_dp_exchange_grads; beside_previous: launched as a programmatic dependent of the previous kernel of the
// stream (which must release its dependents early, see above) without ever waiting for it
int dp_exchange_grads(const neraf_grad_exchange* x, cudaStream_t stream, bool beside_previous);

// elementwise.cu
int convert_bf16(const float* in, int64_t rows, int64_t cols, int64_t ld_in, void* out, int64_t ld_out, void* out_t,
                 int64_t ld_t, cudaStream_t stream);
// bf16 operand copies of up to 8 fp32 matrices + one gathered fp32 vector, one launch (elementwise.cu)
struct PackMatrix {
  const float* in; int64_t rows, cols, ld_in;
  __nv_bfloat16* out; int64_t ld_out;
  int64_t unit_start; int vec;                 // filled by pack_list
};
struct PackList {
  int n;
  PackMatrix m[8];
  int64_t total_units;
  const float* copy_src[8]; float* copy_dst; int64_t copy_width, n_copy;   // copy_dst[j] = copy_src[j / width][j % width]
};
int pack_list(PackList& L, cudaStream_t stream);
// dW[n*ldw + k] = s[n] * g[k] (dW may be null) and dg[k] = sum_n W[n*ldw + k] * s[n] (dg may be null; OVERWRITTEN,
// bit-reproducible) in one launch.  scratch (needed with dg): K doubles + 8 bytes, zero before the first use, left zero.
// compact (optional): (N, E) block (fp32, or bf16 with compact_bf16) with row stride ld_c copied into dW[:, K:K+E]
// widen_* (optional, data parallel): widen_dst[i] = float(widen_src[i]) for i < widen_n, bf16 -> fp32, by extra blocks of
// the same launch
int grid_grads(const float* s, const float* g, const float* W, int64_t ldw, int64_t N, int64_t K, float* dW, float* dg,
               void* scratch, cudaStream_t stream, const void* compact = nullptr, int64_t E = 0, int64_t ld_c = 0,
               bool compact_bf16 = false, const void* widen_src = nullptr, float* widen_dst = nullptr, int64_t widen_n = 0);
// dst[i] = float(src[i]), bf16 -> fp32, i < n (both 16-byte aligned)
int widen_bf16(const void* src, float* dst, int64_t n, cudaStream_t stream);
// out[n] = sum_m X[m*ld + n]   fp32 row-major (M,N)
int colsum_f32(const float* X, int64_t M, int64_t N, int64_t ld, float* out, cudaStream_t stream);
struct HeadColsum {
  float* ptr[8];          // one bias-gradient vector per head (C <= 8)
  long long width;        // columns per head; 0 = no column sums
};
// NERAF_PDL=0 disables programmatic dependent launch everywhere (debugging / A-B timing)
bool pdl_enabled();
// loss.cu: the loss's partial sums, (data parallel) their exchange, the loss gradient through 10 tanh and the heads'
// bias gradients in ONE launch (neraf_loss_grad.fuse_sums); zero[0 .. n_zero) is cleared before the gradient half
int loss_head_fused(const float* y, int64_t M, int64_t N, float* dz_f32, int64_t ld_f32, void* dz_bf16, int64_t ld_bf16,
                    const HeadColsum& cs, const neraf_loss_grad* loss, float* zero, int64_t n_zero, cudaStream_t stream);
// dz = dout * (10 - y^2/10): gradient through 10*tanh; outputs fp32 (M,N) and/or bf16 (M,N) row-major.
// colsum (optional): per-head bias-gradient buffers, column n is added (atomics) to colsum[n / head_width][n % head_width]
// loss (optional): dout is not read but evaluated from (y, loss->gt, loss->sums) as neraf_spectral_loss_backward does.
int head_backward(const float* dout, const float* y, int64_t M, int64_t N, float* dz_f32, int64_t ld_f32, void* dz_bf16,
                  int64_t ld_bf16, float* const* colsum, int64_t head_width, cudaStream_t stream,
                  const neraf_loss_grad* loss = nullptr, float* zero = nullptr, int64_t n_zero = 0);

}  // namespace neraf
