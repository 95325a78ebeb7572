// stft_tables.hpp -- host-side table construction for the STFT plan (pure C++, no CUDA), shared by
// ssr_stft_plan_create and the CPU emulation harness tests/host_emul.cu.
#pragma once
#include <math.h>

#include <utility>
#include <vector>

#include "fft_core.cuh"

namespace ssr {

constexpr long double kPiL = 3.14159265358979323846264338327950288L;

// iterative radix-2 forward FFT in long double (used once per plan for the Bluestein filter)
inline void host_fft_ld(std::vector<long double>& re, std::vector<long double>& im) {
  const size_t n = re.size();
  for (size_t i = 1, j = 0; i < n; ++i) {
    size_t bit = n >> 1;
    for (; j & bit; bit >>= 1) j ^= bit;
    j ^= bit;
    if (i < j) {
      std::swap(re[i], re[j]);
      std::swap(im[i], im[j]);
    }
  }
  for (size_t len = 2; len <= n; len <<= 1) {
    for (size_t i = 0; i < n; i += len) {
      for (size_t k = 0; k < len / 2; ++k) {
        long double ang = -2 * kPiL * (long double)k / (long double)len;
        long double wr = cosl(ang), wi = sinl(ang);
        size_t a = i + k, b = i + k + len / 2;
        long double xr = re[b] * wr - im[b] * wi, xi = re[b] * wi + im[b] * wr;
        re[b] = re[a] - xr;
        im[b] = im[a] - xi;
        re[a] += xr;
        im[a] += xi;
      }
    }
  }
}

struct StftTables {
  int n_fft = 0, M = 0, logM = 0;
  bool bluestein = false;
  std::vector<cd> tw;            // exp(-2 pi i n / M), n < M
  std::vector<double> win_half;  // 0.5 * window[n]
  std::vector<uint16_t> ppos;    // padded smem slot of frequency k after the forward DIF passes
  std::vector<cd> cw;            // bluestein: 0.5 * window[n] * chirp[n]
  std::vector<cd> bfilt;         // bluestein: FFT_M(conj chirp, wrapped) / M in DIF order
  std::vector<cd> cpost;         // bluestein: chirp[k]
};

// returns false when n_fft is unsupported
inline bool build_stft_tables(int n_fft, const double* window, StftTables* t) {
  if (n_fft < 65 || n_fft > 8192) return false;
  const bool pow2 = (n_fft & (n_fft - 1)) == 0;
  int M = 1, logM = 0;
  const int need = pow2 ? n_fft : 2 * n_fft - 1;
  while (M < need) {
    M <<= 1;
    ++logM;
  }
  if (logM < 8) {
    logM = 8;
    M = 256;
  }
  if (logM > 13) return false;
  t->n_fft = n_fft;
  t->M = M;
  t->logM = logM;
  t->bluestein = !(pow2 && M == n_fft);
  std::vector<double> win(n_fft);
  for (int n = 0; n < n_fft; ++n)
    win[n] = window ? window[n]
                    : (double)(0.5L - 0.5L * cosl(2 * kPiL * (long double)n / (long double)n_fft));
  t->tw.resize(M);
  for (int n = 0; n < M; ++n) {
    long double a = -2 * kPiL * (long double)n / (long double)M;
    t->tw[n] = cd{(double)cosl(a), (double)sinl(a)};
  }
  t->win_half.resize(n_fft);
  for (int n = 0; n < n_fft; ++n) t->win_half[n] = 0.5 * win[n];
  t->ppos.resize(M);
  for (int k = 0; k < M; ++k) t->ppos[k] = (uint16_t)pad_idx(dif_position(k, logM));
  t->cw.assign(n_fft, cd{0, 0});
  t->bfilt.assign(M, cd{0, 0});
  t->cpost.assign(n_fft, cd{0, 0});
  if (t->bluestein) {
    std::vector<long double> cr(n_fft), ci(n_fft);
    for (long long n = 0; n < n_fft; ++n) {
      long long m = (n * n) % (2LL * n_fft);  // exact phase reduction of pi*n^2/N
      long double a = -kPiL * (long double)m / (long double)n_fft;
      cr[n] = cosl(a);
      ci[n] = sinl(a);
      t->cw[n] = cd{(double)(0.5L * (long double)win[n] * cr[n]),
                    (double)(0.5L * (long double)win[n] * ci[n])};
      t->cpost[n] = cd{(double)cr[n], (double)ci[n]};
    }
    std::vector<long double> br(M, 0.0L), bi(M, 0.0L);
    for (int m = 0; m < n_fft; ++m) {
      br[m] = cr[m];
      bi[m] = -ci[m];
      if (m) {
        br[M - m] = cr[m];
        bi[M - m] = -ci[m];
      }
    }
    host_fft_ld(br, bi);
    for (int k = 0; k < M; ++k)
      t->bfilt[dif_position(k, logM)] =
          cd{(double)(br[k] / (long double)M), (double)(bi[k] / (long double)M)};
  }
  return true;
}

// ---------------------------------------------------------------------------------------------
// PFA / Cooley-Tukey split for non-power-of-two n_fft = R * P with P <= 1024 (2229 = 3*743 at 48 kHz,
// 1114 = 2*557 at 24 kHz, 743 at 16 kHz, 1486 = 2*743 at 32 kHz):
//   Z[k + P m] = sum_r W_R^{rm} * ( W_N^{rk} * DFT_P(z[R n + r])[k] ),
// each P-point DFT by Bluestein on the 2048-point radix 16x16x8 transform:
//   DFT_P(a)[k] = c[k] * IFFT_2048( FFT_2048(a * c) * B )[k],  c[n] = exp(-i pi n^2 / P),
//   B = FFT_2048(conj c, wrapped) / 2048 stored in the DIF (16,16,8) digit-reversed order.
// ---------------------------------------------------------------------------------------------
inline int pos2048(int k) { return (k % 16) * 128 + ((k / 16) % 16) * 8 + k / 256; }

struct PfaTables {
  int n_fft = 0, R = 0, P = 0;
  std::vector<cd> tw;     // exp(-2 pi i n / 2048)
  std::vector<cd> cwin;   // [r*P + n] = 0.5 * window[R n + r] * c[n]
  std::vector<cd> post;   // [r*P + k] = c[k] * W_N^{rk}
  std::vector<cd> bfilt;  // [pos2048(k)] = FFT_2048(b)[k] / 2048
  std::vector<cd> wr;     // [r*R + m] = W_R^{rm}
};

// smallest R in 1..4 with n_fft % R == 0 and 65 <= n_fft / R <= 1024; 0 if none
inline int pfa_choose_r(int n_fft) {
  for (int R = 1; R <= 4; ++R)
    if (n_fft % R == 0 && n_fft / R <= 1024 && n_fft / R >= 65) return R;
  return 0;
}

inline bool build_pfa_tables(int n_fft, const double* window, PfaTables* t) {
  const int R = pfa_choose_r(n_fft);
  if (!R) return false;
  const int P = n_fft / R, M = 2048;
  t->n_fft = n_fft;
  t->R = R;
  t->P = P;
  t->tw.resize(M);
  for (int n = 0; n < M; ++n) {
    long double a = -2 * kPiL * (long double)n / (long double)M;
    t->tw[n] = cd{(double)cosl(a), (double)sinl(a)};
  }
  std::vector<long double> cr(P), ci(P);
  for (long long n = 0; n < P; ++n) {
    long long m = (n * n) % (2LL * P);
    long double a = -kPiL * (long double)m / (long double)P;
    cr[n] = cosl(a);
    ci[n] = sinl(a);
  }
  t->cwin.resize(n_fft);
  t->post.resize(n_fft);
  for (int r = 0; r < R; ++r)
    for (long long n = 0; n < P; ++n) {
      long double w = window ? (long double)window[R * n + r]
                             : (0.5L - 0.5L * cosl(2 * kPiL * (long double)(R * n + r) / (long double)n_fft));
      t->cwin[r * P + n] = cd{(double)(0.5L * w * cr[n]), (double)(0.5L * w * ci[n])};
      long long e = ((long long)r * n) % n_fft;  // W_N^{r k}
      long double a = -2 * kPiL * (long double)e / (long double)n_fft;
      long double tr = cosl(a), ti = sinl(a);
      t->post[r * P + n] = cd{(double)(cr[n] * tr - ci[n] * ti), (double)(cr[n] * ti + ci[n] * tr)};
    }
  std::vector<long double> br(M, 0.0L), bi(M, 0.0L);
  for (int m = 0; m < P; ++m) {
    br[m] = cr[m];
    bi[m] = -ci[m];
    if (m) {
      br[M - m] = cr[m];
      bi[M - m] = -ci[m];
    }
  }
  host_fft_ld(br, bi);
  t->bfilt.assign(M, cd{0, 0});
  for (int k = 0; k < M; ++k)
    t->bfilt[pos2048(k)] = cd{(double)(br[k] / (long double)M), (double)(bi[k] / (long double)M)};
  t->wr.resize(R * R);
  for (int r = 0; r < R; ++r)
    for (int m = 0; m < R; ++m) {
      long double a = -2 * kPiL * (long double)((r * m) % R) / (long double)R;
      t->wr[r * R + m] = cd{(double)cosl(a), (double)sinl(a)};
    }
  return true;
}

}  // namespace ssr
