// Internal (non-ABI) launch helpers shared between translation units.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/neraf_b200.h"

namespace neraf {

// encode.cu
int encode_queries(const neraf_queries* q, float* out_f32, int64_t ld_f32, void* out_bf16, int64_t ld_bf16,
                   void* out_bf16_t, int64_t ld_t, int ncols_padded, cudaStream_t stream);

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
int mega_run(const MegaJob* jobs, int n_jobs, void* counters, size_t counters_bytes, cudaStream_t stream);

// elementwise.cu
int convert_bf16(const float* in, int64_t rows, int64_t cols, int64_t ld_in, void* out, int64_t ld_out, void* out_t,
                 int64_t ld_t, cudaStream_t stream);
// y[n] = bias[n] + sum_k W[n*ldw + k] * g[k]   (the batch-invariant part of layer 1)
int grid_bias(const float* W, int64_t ldw, const float* bias, const float* g, int64_t N, int64_t K, float* y,
              cudaStream_t stream);
// out[k] = sum_n W[n*ldw + k] * s[n]           (d loss / d grid feature)
int grid_backward(const float* W, int64_t ldw, const float* s, int64_t N, int64_t K, float* out, cudaStream_t stream);
// dW[n*ldw + k] = s[n] * g[k], k < K           (rank-1 weight gradient of the grid block)
int outer_product(const float* s, const float* g, int64_t N, int64_t K, float* dW, int64_t ldw, cudaStream_t stream);
// out[n] = sum_m X[m*ld + n]   fp32 row-major (M,N)
int colsum_f32(const float* X, int64_t M, int64_t N, int64_t ld, float* out, cudaStream_t stream);
// out[n] = sum_m Xt[n*ld + m]  bf16 transposed (N,M)
int rowsum_bf16(const void* Xt, int64_t N, int64_t M, int64_t ld, float* out, cudaStream_t stream);
// dz = dout * (10 - y^2/10): gradient through 10*tanh; outputs fp32 (M,N) and/or bf16 (M,N) + bf16^T (N,M)
// colsum (optional): per-head bias-gradient buffers, column n is added (atomics) to colsum[n / head_width][n % head_width]
int head_backward(const float* dout, const float* y, int64_t M, int64_t N, float* dz_f32, int64_t ld_f32, void* dz_bf16,
                  int64_t ld_bf16, void* dz_bf16_t, int64_t ld_t, float* const* colsum, int64_t head_width,
                  cudaStream_t stream);

}  // namespace neraf
