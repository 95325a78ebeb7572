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

}  // namespace ssr
