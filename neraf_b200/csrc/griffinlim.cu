// Batched Griffin-Lim (K3): magnitude STFT -> room impulse response.
//
// Replaces torchaudio.transforms.GriffinLim as configured by the reference
// (/root/reference/NeRAF/NeRAF_model.py:139, used :229 and :753-754; algorithm
// torchaudio/functional/functional.py:255-353 on torch.stft/istft = 33 ISTFT + 32 STFT cuFFT passes and
// ~200 small kernels per call, spectrogram-sized state round-tripping HBM every iteration).
//
// Design.  The Griffin-Lim state is held in the TIME domain: by linearity of the STFT,
//   angles_{k+1} = normalize(STFT(w_k) - m STFT(w_{k-1})) = normalize(STFT(w_k - m w_{k-1})),
// so one iteration is, per frame:  window -> rFFT -> unit-phase * magnitude -> irFFT -> window ->
// overlap-add, entirely inside one warp, and the only per-signal state is two waveforms.  One CTA owns
// one signal: D = w_k - m w_{k-1} (read by the frames) and the overlap-add accumulator live in shared
// memory for all 33 passes; w_{k-1} is a 60 KB L2-resident line in the workspace touched once per
// iteration; the magnitude is re-read from L2.  HBM traffic is therefore the compulsory
// magnitude read + waveform write (0.18 MB/RIR for RAF) instead of the 39.75 MB/RIR streaming model
// of SURVEY.md section 8d.  Frames that overlap are processed in different "colour" passes so the
// overlap-add is race-free and deterministic without atomics.
#include "common.cuh"
#include "gl_core.h"

namespace neraf {
namespace gl {

struct GlArgs {
  long long n_signals;
  int T, F, L, hop, win_length, n_iter, n_colors;
  float m;                       // momentum / (1 + momentum)
  const float* mag_t;            // (S, T, F) staged magnitudes
  const float* init;             // interleaved complex or nullptr
  int n_channels;                // signal s = item * n_channels + channel
  long long init_sn, init_sc, init_st, init_sf;
  const float* inv_env;          // (L)
  float* prev;                   // (S, L) previous normalised waveform
  float* wave;                   // (S, L)
};

template <int R, int H, bool INV>
__device__ __forceinline__ void run_pass(int lane, int Ns, float* re, float* im, const C2* tw) {
  C2 v[PassShape<R, H>::PER_LANE][R];
  pass_load<R, H, INV>(lane, Ns, re, im, tw, v);
  __syncwarp();
  pass_store<R, H>(lane, Ns, re, im, v);
  __syncwarp();
}

// twp: the concatenated per-pass twiddle tables (gl_core.h PassTables)
template <int H, bool INV>
__device__ __forceinline__ void fft_warp(int lane, float* re, float* im, const C2* twp) {
  int Ns = 1;
#pragma unroll
  for (int i = 0; i < Schedule<H>::N8; ++i) {
    run_pass<8, H, INV>(lane, Ns, re, im, twp + PassTables<H>::offset8(i));
    Ns *= 8;
  }
  if (Schedule<H>::TAIL == 4) run_pass<4, H, INV>(lane, Ns, re, im, twp + PassTables<H>::TAIL_OFFSET);
  if (Schedule<H>::TAIL == 2) run_pass<2, H, INV>(lane, Ns, re, im, twp + PassTables<H>::TAIL_OFFSET);
}

template <int H>
__global__ void __launch_bounds__(512, 1) griffinlim_kernel(GlArgs a) {
  constexpr int N = 2 * H;
  constexpr int HP = padded_size(H);
  extern __shared__ __align__(16) float smem_f[];
  const int Lp = (a.L + 3) & ~3;
  float* D = smem_f;
  float* ACC = D + Lp;
  float* win = ACC + Lp;
  C2* tw = reinterpret_cast<C2*>(win + N);
  C2* twp = tw + N;
  float* fftbuf = reinterpret_cast<float*>(twp + PassTables<H>::TOTAL);

  const int tid = threadIdx.x, nthreads = blockDim.x;
  const int warp = tid / 32, lane = tid % 32, nwarps = nthreads / 32;
  float* re = fftbuf + warp * 2 * HP;
  float* im = re + HP;

  const int left = (N - a.win_length) / 2;
  for (int n = tid; n < N; n += nthreads) {
    const int i = n - left;
    win[n] = (i >= 0 && i < a.win_length) ? (float)(0.5 - 0.5 * cospi(2.0 * (double)i / (double)a.win_length)) : 0.f;
    double s, c;
    sincospi(2.0 * (double)n / (double)N, &s, &c);
    tw[n] = C2{(float)c, (float)(-s)};
  }
  for (int e = tid; e < PassTables<H>::TOTAL; e += nthreads) {
    int num, den;
    pass_table_angle<H>(e, &num, &den);
    double s, c;
    sincospi(2.0 * (double)num / (double)den, &s, &c);
    twp[e] = C2{(float)c, (float)(-s)};
  }
  const float scale = 1.f / (float)N;
  const int j_lo = left / 2, j_hi = (left + a.win_length + 1) / 2;      // complex samples inside the window support

  for (long long sig = blockIdx.x; sig < a.n_signals; sig += gridDim.x) {
    for (int n = tid; n < a.L; n += nthreads) ACC[n] = 0.f;
    __syncthreads();
    const float* mag_sig = a.mag_t + sig * (long long)a.T * a.F;
    float* prev = a.prev + sig * (long long)a.L;
    for (int it = 0; it <= a.n_iter; ++it) {
      // One frame per warp at a time; frames that overlap in time must not overlap-add concurrently: frame t has
      // colour t % n_colors, same-coloured frames are >= n_colors hops apart, one barrier per colour.  (Walking a block
      // of consecutive frames per warp instead saves one of SoundSpaces' 8 rounds and measured no faster; two frames
      // per warp for the 512-point transform halves the rounds but spills at 128 registers and ran 1.9x slower: the
      // kernel is throughput-bound on the shared-memory pipe, not on rounds.)
      auto frame = [&](int t) {
        const float* mag_row = mag_sig + (long long)t * a.F;
        if (it == 0) {
          const float* init_row = a.init ? a.init + 2 * ((sig / a.n_channels) * a.init_sn + (sig % a.n_channels) * a.init_sc +
                                                         (long long)t * a.init_st)
                                         : nullptr;
          init_step<H>(lane, tw, mag_row, init_row, a.init_sf, re, im);
          __syncwarp();
        } else {
          MagRegs<H> mag;                       // consumed after the forward FFT: its latency hides behind it
          load_mag<H>(lane, mag_row, mag);
          load_frame<H>(lane, t, a.hop, a.L, j_lo, j_hi, D, win, re, im);
          __syncwarp();
          fft_warp<H, false>(lane, re, im, twp);
          spectrum_step<H>(lane, tw, mag, re, im);
          __syncwarp();
        }
        fft_warp<H, true>(lane, re, im, twp);
        ola_frame<H>(lane, t, a.hop, a.L, j_lo, j_hi, win, re, im, scale, ACC);
        __syncwarp();
      };
      for (int color = 0; color < a.n_colors; ++color) {
        for (int t = color + a.n_colors * warp; t < a.T; t += a.n_colors * nwarps) frame(t);
        __syncthreads();
      }
      // whole-waveform pass: window-envelope normalisation (torch.istft) + momentum combination.  Every thread issues
      // ALL its global loads (envelope, previous waveform) before it touches them: one exposed L2 latency per
      // iteration instead of one per unrolled group (ncu: 16 % of all stall samples sat on these loads).
      const bool last = it == a.n_iter;
      float* wave = a.wave + sig * (long long)a.L;
      if ((a.L & 3) == 0) {
        constexpr int UNR = 4;
        const int L4 = a.L >> 2;
        const float4* env4 = reinterpret_cast<const float4*>(a.inv_env);
        float4* prev4 = reinterpret_cast<float4*>(prev);
        float4* acc4 = reinterpret_cast<float4*>(ACC);
        float4* d4 = reinterpret_cast<float4*>(D);
        for (int base = tid; base < L4; base += nthreads * UNR) {
          float4 e[UNR], p[UNR];
#pragma unroll
          for (int u = 0; u < UNR; ++u) {
            const int i = base + u * nthreads;
            if (i < L4) {
              e[u] = __ldg(env4 + i);
              if (!last && it > 0) p[u] = prev4[i];
            }
          }
#pragma unroll
          for (int u = 0; u < UNR; ++u) {
            const int i = base + u * nthreads;
            if (i < L4) {
              const float4 acc = acc4[i];
              const float4 b = make_float4(acc.x * e[u].x, acc.y * e[u].y, acc.z * e[u].z, acc.w * e[u].w);
              acc4[i] = make_float4(0.f, 0.f, 0.f, 0.f);
              if (last) {
                reinterpret_cast<float4*>(wave)[i] = b;
              } else {
                const float4 q = it > 0 ? p[u] : make_float4(0.f, 0.f, 0.f, 0.f);
                d4[i] = make_float4(b.x - a.m * q.x, b.y - a.m * q.y, b.z - a.m * q.z, b.w - a.m * q.w);
                prev4[i] = b;
              }
            }
          }
        }
      } else {
        for (int n = tid; n < a.L; n += nthreads) {
          const float b = ACC[n] * __ldg(a.inv_env + n);
          ACC[n] = 0.f;
          if (last) {
            wave[n] = b;
          } else {
            const float p = it > 0 ? prev[n] : 0.f;
            D[n] = b - a.m * p;
            prev[n] = b;
          }
        }
      }
      __syncthreads();
    }
  }
}

__global__ void __launch_bounds__(256) gl_stage_kernel(const float* __restrict__ spec, int n_channels, long long sn,
                                                       long long sc, long long st, long long sf, long long total,
                                                       int T, int F, int is_log, float* __restrict__ mag_t) {
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= total) return;
  const int f = (int)(idx % F);
  const long long r = idx / F;
  const int t = (int)(r % T);
  const long long s = r / T;
  float v = __ldg(spec + (s / n_channels) * sn + (s % n_channels) * sc + t * st + f * sf);
  if (is_log) v = fminf(fmaxf(expf(v) - 1e-3f, 0.f), 10000.f);      // NeRAF_model.py:746-747
  mag_t[idx] = v;
}

__global__ void __launch_bounds__(256) gl_env_kernel(float* __restrict__ inv_env, int L, int T, int hop, int N,
                                                     int win_length) {
  const int n = blockIdx.x * blockDim.x + threadIdx.x;
  if (n >= L) return;
  const int left = (N - win_length) / 2;
  double e = 0.0;
  // frames t with 0 <= n + N/2 - t*hop < N
  const int p = n + N / 2;
  int t_hi = p / hop;
  if (t_hi > T - 1) t_hi = T - 1;
  for (int t = t_hi; t >= 0; --t) {
    const int i = p - t * hop;
    if (i >= N) break;
    const int j = i - left;
    if (j >= 0 && j < win_length) {
      const double w = 0.5 - 0.5 * cospi(2.0 * (double)j / (double)win_length);
      e += w * w;
    }
  }
  inv_env[n] = e > 1e-11 ? (float)(1.0 / e) : 0.f;
}

static size_t pass_tables_total(int H) {
  switch (H) {
    case 32: return PassTables<32>::TOTAL;
    case 64: return PassTables<64>::TOTAL;
    case 128: return PassTables<128>::TOTAL;
    case 256: return PassTables<256>::TOTAL;
    case 512: return PassTables<512>::TOTAL;
    case 1024: return PassTables<1024>::TOTAL;
  }
  return 0;
}

struct Plan {
  int H, L, n_colors, nwarps;
  size_t smem_bytes;
  size_t off_env, off_mag, off_prev, ws_bytes;
};

static int make_plan(const neraf_gl_params* p, long long S, Plan* pl) {
  NERAF_REQUIRE(p, "griffinlim: params is null");
  const int N = p->n_fft;
  NERAF_REQUIRE(N >= 64 && N <= 2048 && (N & (N - 1)) == 0, "griffinlim: n_fft %d must be a power of two in [64, 2048]", N);
  NERAF_REQUIRE(p->win_length >= 1 && p->win_length <= N, "griffinlim: win_length %d out of range", p->win_length);
  NERAF_REQUIRE(p->hop >= 1 && p->hop <= p->win_length, "griffinlim: hop %d out of range", p->hop);
  NERAF_REQUIRE(p->n_frames >= 2, "griffinlim: need at least 2 frames");
  NERAF_REQUIRE(p->n_iter >= 0, "griffinlim: n_iter < 0");
  NERAF_REQUIRE(p->momentum >= 0.f && p->momentum < 1.f, "griffinlim: momentum must be in [0, 1)");   // functional.py:299
  NERAF_REQUIRE(S >= 0, "griffinlim: n_signals < 0");
  pl->H = N / 2;
  pl->L = p->hop * (p->n_frames - 1);
  NERAF_REQUIRE(pl->L > N / 2, "griffinlim: signal too short for reflect padding (L=%d, n_fft=%d)", pl->L, N);
  {
    // samples one frame overlap-adds: the window support rounded outwards to whole complex pairs (kernel: j_lo, j_hi)
    const int left = (N - p->win_length) / 2;
    const int span = 2 * ((left + p->win_length + 1) / 2 - left / 2);
    pl->n_colors = (int)ceil_div(span, p->hop);
  }
  const int F = N / 2 + 1;
  const size_t Lp = (size_t)((pl->L + 3) & ~3);
  pl->nwarps = 0;
  for (int nw = 16; nw >= 2; nw /= 2) {
    const size_t bytes = (2 * Lp + N) * 4 + ((size_t)N + pass_tables_total(pl->H)) * 8 + (size_t)nw * 2 * padded_size(pl->H) * 4;
    if (bytes <= 227 * 1024) { pl->nwarps = nw; pl->smem_bytes = bytes; break; }
  }
  if (pl->nwarps == 0)
    return set_error(NERAF_ERR_UNSUPPORTED, "griffinlim: waveform of %d samples does not fit in shared memory", pl->L);
  size_t cur = 0;
  auto take = [&](size_t b) { size_t at = cur; cur += (b + 255) / 256 * 256; return at; };
  pl->off_env = take((size_t)pl->L * 4);
  pl->off_mag = take((size_t)S * p->n_frames * F * 4);
  pl->off_prev = take((size_t)S * pl->L * 4);
  pl->ws_bytes = cur;
  return NERAF_OK;
}

template <int H>
static int launch_gl(const GlArgs& a, const Plan& pl, cudaStream_t stream) {
  NERAF_CHECK_CUDA(cudaFuncSetAttribute(griffinlim_kernel<H>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)pl.smem_bytes));
  const long long grid = a.n_signals < sm_count() ? a.n_signals : sm_count();
  griffinlim_kernel<H><<<(unsigned)grid, pl.nwarps * 32, pl.smem_bytes, stream>>>(a);
  NERAF_CHECK_LAUNCH("griffinlim_kernel");
  return NERAF_OK;
}

}  // namespace gl
}  // namespace neraf

using namespace neraf;

extern "C" int neraf_griffinlim_sizes(const neraf_gl_params* p, int64_t n_signals, size_t* workspace_bytes) {
  gl::Plan pl;
  NERAF_TRY(gl::make_plan(p, n_signals, &pl));
  if (workspace_bytes) *workspace_bytes = pl.ws_bytes;
  return NERAF_OK;
}

extern "C" int neraf_griffinlim(const neraf_gl_params* p, int64_t n_items, int32_t n_channels, const float* spec,
                                int64_t sn, int64_t sc, int64_t st, int64_t sf, const float* init, int64_t isn,
                                int64_t isc, int64_t ist, int64_t isf, void* ws, size_t ws_bytes, float* wave,
                                neraf_stream_t stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  NERAF_REQUIRE(n_items >= 0 && n_channels >= 1, "griffinlim: bad n_items / n_channels");
  const int64_t S = n_items * n_channels;
  gl::Plan pl;
  NERAF_TRY(gl::make_plan(p, S, &pl));
  if (S == 0) return NERAF_OK;
  NERAF_REQUIRE(spec && wave && ws, "griffinlim: null pointer");
  if (ws_bytes < pl.ws_bytes)
    return set_error(NERAF_ERR_WORKSPACE, "griffinlim: workspace %zu < %zu bytes", ws_bytes, pl.ws_bytes);
  const int F = p->n_fft / 2 + 1;
  uint8_t* base = reinterpret_cast<uint8_t*>(ws);
  float* inv_env = reinterpret_cast<float*>(base + pl.off_env);
  float* mag_t = reinterpret_cast<float*>(base + pl.off_mag);
  float* prev = reinterpret_cast<float*>(base + pl.off_prev);

  gl::gl_env_kernel<<<(unsigned)ceil_div(pl.L, 256), 256, 0, stream>>>(inv_env, pl.L, p->n_frames, p->hop, p->n_fft,
                                                                       p->win_length);
  NERAF_CHECK_LAUNCH("gl_env_kernel");
  const long long total = (long long)S * p->n_frames * F;
  gl::gl_stage_kernel<<<(unsigned)ceil_div(total, 256), 256, 0, stream>>>(spec, n_channels, sn, sc, st, sf, total,
                                                                          p->n_frames, F, p->input_is_log, mag_t);
  NERAF_CHECK_LAUNCH("gl_stage_kernel");

  gl::GlArgs a;
  a.n_signals = S; a.T = p->n_frames; a.F = F; a.L = pl.L; a.hop = p->hop; a.win_length = p->win_length;
  a.n_iter = p->n_iter; a.n_colors = pl.n_colors; a.m = p->momentum / (1.f + p->momentum);
  a.mag_t = mag_t; a.init = init; a.n_channels = n_channels;
  a.init_sn = isn; a.init_sc = isc; a.init_st = ist; a.init_sf = isf;
  a.inv_env = inv_env; a.prev = prev; a.wave = wave;
  switch (pl.H) {
    case 32: return gl::launch_gl<32>(a, pl, stream);
    case 64: return gl::launch_gl<64>(a, pl, stream);
    case 128: return gl::launch_gl<128>(a, pl, stream);
    case 256: return gl::launch_gl<256>(a, pl, stream);
    case 512: return gl::launch_gl<512>(a, pl, stream);
    case 1024: return gl::launch_gl<1024>(a, pl, stream);
  }
  return set_error(NERAF_ERR_UNSUPPORTED, "griffinlim: unsupported n_fft %d", p->n_fft);
}
