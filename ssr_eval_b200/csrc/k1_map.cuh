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

// Warp-local variant of the same maps (SSR_WARPLOCAL): warp w owns the length-128 sub-transforms ("blocks",
// block q = pass-1 output index = k % 16) {2w, 16-2w, 2w+1, 15-2w} (w = 0: {1, 15, 8, 0}) in pass 2 AND in pass 3.
// A block and the block of its Hermitian partners (16 - q) sit in the same warp, so the pass-2 -> pass-3 exchange
// never leaves the warp: __syncwarp() instead of a CTA barrier.
SSR_HD int v2w_pass2_block(int t) {  // block of thread t's pass-2 butterfly (its index inside the block is t & 7)
  // order inside the warp: an even and an odd block per half-warp (the float32 kernels' 8-byte slots, padded
  // i + i/16, are conflict-free per half-warp only then)
  const int w = t >> 5, s = (t >> 3) & 3, g = 2 * w;
  if (w == 0) return s == 0 ? 1 : (s == 1 ? 8 : (s == 2 ? 15 : 0));
  return s == 0 ? g : (s == 1 ? g + 1 : (s == 2 ? 16 - g : 15 - g));
}

// thread t -> (butterfly a, butterfly b = partner(a)); lane 31 of warp 0 owns the two self-paired butterflies
SSR_HD void v2w_thread_butterflies(int t, int* ia, int* ib, bool* special) {
  const int w = t >> 5, l = t & 31;
  *special = false;
  if (w == 0) {
    if (l < 16) {
      *ia = 16 + l;
      *ib = 255 - l;
    } else if (l < 24) {
      *ia = 128 + (l - 16);
      *ib = 143 - (l - 16);
    } else if (l < 31) {
      *ia = l - 23;   // 1..7
      *ib = 16 - *ia;  // 15..9
    } else {
      *ia = 0;
      *ib = 8;
      *special = true;
    }
  } else {
    *ia = (l < 16) ? 32 * w + l : 32 * w + 16 + (l - 16);  // blocks 2w, 2w+1
    *ib = 271 - *ia;
  }
}

}  // namespace ssr
