// k1_map.cuh -- index maps of the specialised 2048-point STFT kernel (radix 16 x 16 x 8, in-place
// DIF, 128 threads).  Host+device so tests/host_emul.cu can check them exhaustively.
//
// After the three DIF passes frequency k sits at position
//   pos(k) = (k % 16) * 128 + ((k / 16) % 16) * 8 + k / 256,
// i.e. the last-pass butterfly i (elements 8i .. 8i+7) holds k = k_low(i) + 256 q, q = 0..7, with
//   k_low(i) = (i / 16) + 16 * (i % 16).
// The Hermitian partner N - k of a bin lives in butterfly partner(i):
//   partner(i) = 271 - i  for i >= 16,   16 - i  for 1 <= i <= 15 (i != 8),   i for i in {0, 8},
// at q' = 7 - q (q' = (8 - q) % 8 inside butterfly 0).  One thread owns a butterfly AND its partner,
// so the est/target spectra are separated entirely in registers.
#pragma once
#include "fft_core.cuh"

namespace ssr {

constexpr int kV2Threads = 128;

SSR_HD int v2_klow(int i) { return (i >> 4) + ((i & 15) << 4); }

// thread t -> (butterfly a, butterfly b); consecutive threads take consecutive a (and descending b)
// so the stride-8 loads stay bank-conflict free.  Thread 127 owns the two self-paired butterflies.
SSR_HD void v2_thread_butterflies(int t, int* ia, int* ib) {
  if (t < 120) {
    *ia = 16 + t;
    *ib = 255 - t;
  } else if (t < 127) {
    *ia = t - 119;       // 1..7
    *ib = 16 - *ia;      // 15..9
  } else {
    *ia = 0;
    *ib = 8;
  }
}

}  // namespace ssr
