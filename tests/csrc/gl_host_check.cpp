// CPU harness for neraf_b200/csrc/gl_core.h: drives the per-lane building blocks of the fused
// Griffin-Lim kernel lane by lane (32 emulated lanes, the two halves of every FFT pass separated exactly
// where the GPU kernel places its warp barriers) so the index arithmetic can be checked against the
// oracle without a GPU.  Test infrastructure only.
//
//   gl_host_check fft                      -> self-check of the H-point FFTs against a naive DFT
//   gl_host_check gl n_fft win hop T n_iter momentum has_init < in.bin > out.bin
//        in : mag [T][F] f32, then (if has_init) init [T][F] complex64 ;  out : wave [hop*(T-1)] f32
#include <cmath>
#include <complex>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

#include "../../neraf_b200/csrc/gl_core.h"

using namespace neraf::gl;

template <int R, int H, bool INV>
static void run_pass(int Ns, float* re, float* im, const C2* tw) {
  constexpr int PL = PassShape<R, H>::PER_LANE;
  static C2 regs[32][PL][R];
  for (int lane = 0; lane < 32; ++lane) pass_load<R, H, INV>(lane, Ns, re, im, tw, regs[lane]);
  for (int lane = 0; lane < 32; ++lane) pass_store<R, H>(lane, Ns, re, im, regs[lane]);
}

// per-pass twiddle tables exactly as the kernel builds them (gl_core.h PassTables / pass_table_angle)
template <int H>
static const C2* pass_tables() {
  static std::vector<C2> t;
  if (t.empty()) {
    t.resize(PassTables<H>::TOTAL + 1);
    for (int e = 0; e < PassTables<H>::TOTAL; ++e) {
      int num, den;
      pass_table_angle<H>(e, &num, &den);
      const double a = -2.0 * M_PI * num / den;
      t[e] = C2{(float)std::cos(a), (float)std::sin(a)};
    }
  }
  return t.data();
}

template <int H, bool INV>
static void fft_host(float* re, float* im, const C2*) {
  const C2* twp = pass_tables<H>();
  int Ns = 1;
  for (int i = 0; i < Schedule<H>::N8; ++i) { run_pass<8, H, INV>(Ns, re, im, twp + PassTables<H>::offset8(i)); Ns *= 8; }
  if (Schedule<H>::TAIL == 4) run_pass<4, H, INV>(Ns, re, im, twp + PassTables<H>::TAIL_OFFSET);
  if (Schedule<H>::TAIL == 2) run_pass<2, H, INV>(Ns, re, im, twp + PassTables<H>::TAIL_OFFSET);
}

static std::vector<C2> make_tw(int n_fft) {
  std::vector<C2> tw(n_fft);
  for (int m = 0; m < n_fft; ++m) {
    const double a = -2.0 * M_PI * m / n_fft;
    tw[m] = C2{(float)std::cos(a), (float)std::sin(a)};
  }
  return tw;
}

template <int H>
static double check_fft() {
  auto tw = make_tw(2 * H);
  std::vector<float> re(padded_size(H)), im(padded_size(H));
  std::vector<std::complex<double>> x(H);
  srand(H);
  for (int i = 0; i < H; ++i) {
    x[i] = {rand() / (double)RAND_MAX - 0.5, rand() / (double)RAND_MAX - 0.5};
    re[padi(i)] = (float)x[i].real();
    im[padi(i)] = (float)x[i].imag();
  }
  fft_host<H, false>(re.data(), im.data(), tw.data());
  double worst = 0;
  for (int k = 0; k < H; ++k) {
    std::complex<double> s = 0;
    for (int n = 0; n < H; ++n) s += x[n] * std::polar(1.0, -2.0 * M_PI * n * k / H);
    worst = std::fmax(worst, std::abs(s - std::complex<double>(re[padi(k)], im[padi(k)])));
  }
  fft_host<H, true>(re.data(), im.data(), tw.data());
  for (int n = 0; n < H; ++n)
    worst = std::fmax(worst, std::abs(x[n] - std::complex<double>(re[padi(n)], im[padi(n)]) / (double)H));
  return worst;
}

template <int H>
static int run_gl(int win_length, int hop, int T, int n_iter, float momentum, bool has_init) {
  const int N = 2 * H, F = H + 1, L = hop * (T - 1);
  std::vector<float> mag((size_t)T * F), init;
  if (fread(mag.data(), 4, mag.size(), stdin) != mag.size()) return 2;
  if (has_init) {
    init.resize((size_t)T * F * 2);
    if (fread(init.data(), 4, init.size(), stdin) != init.size()) return 2;
  }
  auto tw = make_tw(N);
  std::vector<float> win(N, 0.f);
  const int left = (N - win_length) / 2;
  for (int i = 0; i < win_length; ++i) win[left + i] = (float)(0.5 - 0.5 * std::cos(2.0 * M_PI * i / win_length));
  std::vector<float> inv_env(L);
  for (int n = 0; n < L; ++n) {
    double e = 0;
    for (int t = 0; t < T; ++t) {
      const int i = n + H - t * hop;
      if (i >= 0 && i < N) e += (double)win[i] * win[i];
    }
    inv_env[n] = (float)(1.0 / e);
  }
  const float m = momentum / (1.f + momentum);
  const float scale = 1.f / (float)N;
  std::vector<float> D(L, 0.f), ACC(L, 0.f), G(L, 0.f), re(padded_size(H)), im(padded_size(H));
  for (int it = 0; it <= n_iter; ++it) {
    for (int t = 0; t < T; ++t) {
      if (it == 0) {
        for (int lane = 0; lane < 32; ++lane)
          init_step<H>(lane, tw.data(), &mag[(size_t)t * F], has_init ? &init[(size_t)t * F * 2] : nullptr, 1, re.data(), im.data());
      } else {
        for (int lane = 0; lane < 32; ++lane) load_frame<H>(lane, t, hop, L, left / 2, (left + win_length + 1) / 2, D.data(), win.data(), re.data(), im.data());
        fft_host<H, false>(re.data(), im.data(), tw.data());
        for (int lane = 0; lane < 32; ++lane) {
          MagRegs<H> m;
          load_mag<H>(lane, &mag[(size_t)t * F], m);
          spectrum_step<H>(lane, tw.data(), m, re.data(), im.data());
        }
      }
      fft_host<H, true>(re.data(), im.data(), tw.data());
      for (int lane = 0; lane < 32; ++lane) ola_frame<H>(lane, t, hop, L, left / 2, (left + win_length + 1) / 2, win.data(), re.data(), im.data(), scale, ACC.data());
    }
    for (int n = 0; n < L; ++n) {
      const float b = ACC[n] * inv_env[n];
      D[n] = b - m * G[n];
      G[n] = b;
      ACC[n] = 0.f;
    }
  }
  fwrite(G.data(), 4, L, stdout);
  return 0;
}

int main(int argc, char** argv) {
  if (argc >= 2 && !strcmp(argv[1], "fft")) {
    const double e[6] = {check_fft<32>(), check_fft<64>(), check_fft<128>(), check_fft<256>(), check_fft<512>(), check_fft<1024>()};
    int bad = 0;
    for (int i = 0; i < 6; ++i) {
      printf("H=%d max_abs_err=%.3e\n", 32 << i, e[i]);
      if (!(e[i] < 2e-4)) bad = 1;
    }
    return bad;
  }
  if (argc == 9 && !strcmp(argv[1], "gl")) {
    const int n_fft = atoi(argv[2]), win = atoi(argv[3]), hop = atoi(argv[4]), T = atoi(argv[5]), n_iter = atoi(argv[6]);
    const float mom = (float)atof(argv[7]);
    const bool has_init = atoi(argv[8]) != 0;
    switch (n_fft) {
      case 64: return run_gl<32>(win, hop, T, n_iter, mom, has_init);
      case 128: return run_gl<64>(win, hop, T, n_iter, mom, has_init);
      case 256: return run_gl<128>(win, hop, T, n_iter, mom, has_init);
      case 512: return run_gl<256>(win, hop, T, n_iter, mom, has_init);
      case 1024: return run_gl<512>(win, hop, T, n_iter, mom, has_init);
      case 2048: return run_gl<1024>(win, hop, T, n_iter, mom, has_init);
    }
  }
  fprintf(stderr, "usage: gl_host_check fft | gl n_fft win hop T n_iter momentum has_init\n");
  return 64;
}
