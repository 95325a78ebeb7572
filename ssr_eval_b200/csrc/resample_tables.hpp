// resample_tables.hpp -- host-side construction of the tap tables of the staged K3 kernels (resample.cu) and the index
// arithmetic of k_resample_pair, in a form tests/host_emul.cu can run on the CPU: the emulation replays the kernel's
// window / filter indexing thread by thread and must reproduce the plain polyphase sum bit for bit.
#pragma once
#include <vector>

namespace ssr {

constexpr int kBankStride = 512;  // row stride of the transposed banks (floats / float2s): up <= 512
#ifndef SSR_K3_MIN_TP
#define SSR_K3_MIN_TP 128  // smallest CTA of k_resample_pair (measured: 160 threads x 5 CTAs per SM beat 320 x 3)
#endif
#ifndef SSR_K3_RP
#define SSR_K3_RP 16  // output pairs per thread of k_resample_pair
#endif

// phase of the outputs j = n (mod up): (half_len + n * down) % up
inline std::vector<int> k3_output_order_phases(int up, int down, int half_len) {
  std::vector<int> phase(up);
  for (int n = 0; n < up; ++n) phase[n] = (int)(((long long)half_len + (long long)n * down) % up);
  return phase;
}

// bank_t[k * kBankStride + n] = bank[phase(n)][k]  (bank: [up][K])
inline std::vector<float> k3_build_bank_t(int up, int down, int K, int half_len, const float* bank) {
  const std::vector<int> phase = k3_output_order_phases(up, down, half_len);
  std::vector<float> t((size_t)K * kBankStride, 0.f);
  for (int k = 0; k < K; ++k)
    for (int n = 0; n < up; ++n) t[(size_t)k * kBankStride + n] = bank[(size_t)phase[n] * K + k];
  return t;
}

struct K3PairTables {
  int nl = 0, tp = 0, m = 0;  // 64-bit loads per window (0: no pair kernel for this plan), threads per CTA, 2 tp / up
  std::vector<float> g;       // float2 pair_g[(par * 2 nl + q) * kBankStride + c], flattened (x, y)
  std::vector<int> thr;       // int2 pair_thr[t] = {newest(jb + 2 t) - newest(jb), c(t)}, flattened
};

// k_resample_pair is instantiated for windows of 24 / 26 words (K = 21 / 22 of the evaluation's sample-rate pairs).
// Block: 2 * TP = m * up outputs with m * down even (the window alignment repeats), smallest TP >= SSR_K3_MIN_TP.
// Returns false when the plan gets no pair kernel.
inline bool k3_build_pair_tables(int up, int down, int K, int half_len, const float* bank, K3PairTables* out) {
  *out = K3PairTables();
  if (up > kBankStride) return false;
  const int d_max = (down + up - 1) / up;
  const int NL = (K + d_max + 1 + 1) / 2;
  if (NL != 12 && NL != 13) return false;
  int m = 0;
  for (int c = 1; (long long)c * up <= 1024; ++c)
    if (((long long)c * up) % 2 == 0 && ((long long)c * down) % 2 == 0 && c * up / 2 >= SSR_K3_MIN_TP) {
      m = c;
      break;
    }
  if (m == 0) return false;
  const int TP = m * up / 2;
  const std::vector<int> phase = k3_output_order_phases(up, down, half_len);
  auto tap = [&](int ph, int k) { return (k >= 0 && k < K) ? bank[(size_t)ph * K + k] : 0.f; };
  out->g.assign((size_t)2 * 2 * NL * kBankStride * 2, 0.f);
  const int P = (up % 2 == 0) ? up / 2 : up;  // distinct output pairs (mod up) a CTA's threads see
  for (int par = 0; par < 2; ++par)
    for (int c = 0; c < P; ++c) {
      const int n = (2 * c) % up;
      const int pa = phase[n], pb = phase[(n + 1) % up];
      const int ka0 = K - 1 + par, kb0 = ka0 + (pa + down) / up;
      for (int q = 0; q < NL; ++q) {
        float* ga = &out->g[2 * (((size_t)par * 2 * NL + q) * kBankStride + c)];
        float* gb = &out->g[2 * (((size_t)par * 2 * NL + NL + q) * kBankStride + c)];
        ga[0] = tap(pa, ka0 - 2 * q);
        ga[1] = tap(pa, ka0 - 2 * q - 1);
        gb[0] = tap(pb, kb0 - 2 * q);
        gb[1] = tap(pb, kb0 - 2 * q - 1);
      }
    }
  out->thr.resize((size_t)2 * TP);
  for (int th = 0; th < TP; ++th) {
    const long long c = (long long)half_len + 2LL * th * down;
    out->thr[2 * th] = (int)(c / up - half_len / up);
    out->thr[2 * th + 1] = th % P;
  }
  out->nl = NL;
  out->tp = TP;
  out->m = m;
  return true;
}

// input samples one CTA of k_resample_pair stages: newest(last) - newest(first) + K, + d_max, + the zero-tap overhang
inline long long k3_pair_span(int up, int down, int K, int nl, int tp, int rp) {
  const int d_max = (down + up - 1) / up;
  const long long outs = 2LL * tp * rp;
  return (outs - 1) * down / up + 2 + K + d_max + 2 * nl - K + 2;
}

}  // namespace ssr
