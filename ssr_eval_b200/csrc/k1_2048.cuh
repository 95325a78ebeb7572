// k1_2048.cuh -- K1, specialised n_fft = 2048 kernel (radix 16x16x8, register-paired epilogue).
#pragma once
#include "k1_common.cuh"
#include "k1_map.cuh"
#include "tmem.cuh"
#include <type_traits>

namespace ssr {

// loads the compiler may not sink towards their use (issued where they are written)
__device__ __forceinline__ double ldg_f64_pinned(const double* p) {
  double r;
  asm volatile("ld.global.nc.f64 %0, [%1];" : "=d"(r) : "l"(p));
  return r;
}

__device__ __forceinline__ float ldg_f32_pinned(const float* p) {
  float r;
  asm volatile("ld.global.nc.f32 %0, [%1];" : "=f"(r) : "l"(p));
  return r;
}
__device__ __forceinline__ float ldg_pinned(const float* p) { return ldg_f32_pinned(p); }
__device__ __forceinline__ double ldg_pinned(const double* p) { return ldg_f64_pinned(p); }

// ---------------------------------------------------------------------------------------------
// K1, specialised: n_fft = 2048 (BASELINE config 2 and every evaluation at 44.1 kHz).
// 128 threads, 16 points per thread: radix 16 x 16 x 8 in-place DIF.
//   pass 1: samples come from global memory (coalesced, window folded in); the 15 pass-1 twiddles
//           of a thread never change and live in TENSOR MEMORY (60 columns of the thread's TMEM lane,
//           written once with tcgen05.st, fetched per frame with tcgen05.ld in four double-buffered
//           chunks): 60 registers less than keeping them in the register file, which is what lets
//           4 CTAs (16 warps) share an SM, and no LSU / shared-memory bandwidth;
//           RING (hop == 512 == 4 x 128): a thread's 16 (target, est) samples of a frame are the same
//           as the previous frame's shifted by 4, so they are kept -- already converted to float64 --
//           in a 64-column TMEM ring; an interior frame loads only its 4 new samples per signal from
//           global memory (one frame ahead, into 8 registers) and converts 8 instead of 32 values;
//   pass 2: twiddles W_128^{jq} (120 values) from a conflict-free shared table (RING) or, when there is no sample
//           ring, from the other half of the CTA's tensor-memory columns;
//   pass 3: no twiddles; every thread transforms a butterfly AND its Hermitian partner
//           (k1_map.cuh), so Z[k] and Z[N-k] meet in registers and the epilogue needs no
//           further shared-memory traffic.
// Shared-memory traffic per frame: 2 exchanges (4 x 32 KB) + 30 KB of twiddles.
// ---------------------------------------------------------------------------------------------
// FIXED >= 0: the metric flags are the compile-time constant FIXED (bit 3 = the magnitude
// spectrograms are written for K2); hot configurations: 1 = LSD only, 7 = LSD + log-sispec + sispec,
// 15 = those + spectrograms.  FIXED < 0: run-time flags.


// ET = double: the ESTIMATE is a float64 waveform (the reference's IIR low-pass keys, see k1_generic.cuh): E stays
// float64 from the waveform to the sums, only T is rounded to complex64 / float32 -- the arithmetic of the generic
// kernel's float64-estimate path on this kernel's machinery.
template <int FIXED, bool RING, typename ET = float>
__global__ void __launch_bounds__(kV2Threads, 4)
k_stft_metrics_2048(StftDev P, const ET* __restrict__ est, const float* __restrict__ tgt,
                    const long long* __restrict__ offsets, const int* __restrict__ item_start,
                    const int* __restrict__ item_pair, int n_items, int chunk, unsigned flags,
                    double* __restrict__ partials, float* __restrict__ spec_e,
                    float* __restrict__ spec_t, const long long* __restrict__ spec_off, int* __restrict__ next_item) {
  constexpr int N = 2048, F = 1025, NW = kV2Threads / 32;
  // tensor-memory columns of the CTA: [0, 60) the pass-1 twiddles; [64, 128): RING -- the float64 sample ring;
  // otherwise (SSR_K1_TW2_TMEM) the 15 pass-2 twiddles W_128^{j2 q} of the thread, stored group by group
  // (q0, q0+4, q0+8, q0+12): no shared-memory loads for them (240 of a frame's ~1480 LSU wavefronts)
  constexpr bool kTw2Tmem = !RING;  // (hop 441: +3.4 %; for hop 512 the sample ring is the better use, +1 %)
  constexpr int kTmemCols = (RING || kTw2Tmem) ? 128 : 64;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  cd* const buf = reinterpret_cast<cd*>(smem_raw);                                // N + N/8 slots
  float2* const edge_raw = reinterpret_cast<float2*>(smem_raw);  // edge frames stage N raw pairs inside buf
  float* const row_t = reinterpret_cast<float*>(smem_raw + sizeof(cd) * (N + N / 8));  // magnitude rows (store mode)
  float* const row_e = row_t + 1104;
  __shared__ __align__(16) cd tw2[15 * 8];
  constexpr bool E64 = sizeof(ET) == 8;
  using LT = typename std::conditional<E64, double, float>::type;  // type of the per-frame LSD sums
  __shared__ LT lsd_part[kMaxChunk][NW];
  __shared__ double red[NW][kPartials];
  __shared__ unsigned tmem_slot;

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int hop = P.hop;
  if (FIXED >= 0) flags = (unsigned)FIXED;
  const bool want_lsd = flags & SSR_METRIC_LSD, want_log = flags & SSR_METRIC_LOG_SISPEC,
             want_lin = flags & SSR_METRIC_SISPEC;
  if (FIXED >= 0 && !(FIXED & 8)) {
    spec_e = nullptr;
    spec_t = nullptr;
  }

  // per-thread constants: pass-1 twiddles W^{tid q}, q = 1..15, into this thread's TMEM lane
  const unsigned tmem_base = tmem_alloc<kTmemCols>(&tmem_slot, warp);
#pragma unroll
  for (int c = 0; c < 4; ++c) {
    cd w4[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int q = 4 * c + i + 1;
      w4[i] = q < 16 ? P.tw[tid * q] : cd{0.0, 0.0};
    }
    unsigned r[16];
    tmem_pack4(w4, r);
    tmem_st16(tmem_base + 16 * c, r);
  }
  if (kTw2Tmem) {
#pragma unroll
    for (int g = 0; g < 4; ++g) {
      cd w4[4];
#pragma unroll
      for (int q1 = 0; q1 < 4; ++q1) {
        const int q = g + 4 * q1;
        w4[q1] = P.tw[(16 * (tid & 7) * q) & 2047];  // W_128^{j2 q} (q = 0: 1, unused)
      }
      unsigned r[16];
      tmem_pack4(w4, r);
      tmem_st16(tmem_base + 64 + 16 * g, r);
    }
  }
  tmem_wait_st();
  if (tid < 120) tw2[tid] = P.tw[16 * (tid & 7) * ((tid >> 3) + 1)];  // tw2[(q-1)*8 + j] = W_128^{jq}
  int ia, ib;
#ifdef SSR_WARPLOCAL
  // passes 2 and 3 of a sub-transform block AND of its Hermitian-partner block run in one warp (k1_map.cuh):
  // the second exchange is warp-local
  bool special;
  v2w_thread_butterflies(tid, &ia, &ib, &special);
  const int blk2 = v2w_pass2_block(tid);
#else
  v2_thread_butterflies(tid, &ia, &ib);
  const bool special = (tid == kV2Threads - 1);
  const int blk2 = tid >> 3;
#endif
  const int ka = v2_klow(ia), kb = v2_klow(ib);
  const int j2 = tid & 7;
  // padded slots: pass 1 element q -> p1 + 144 q; pass 2 element r -> p2 + 9 r; pass 3 -> 9 i + r
  cd* const b1 = buf + pad_idx(tid);
  cd* const b2 = buf + pad_idx(blk2 * 128 + j2);
  const cd* const b3a = buf + 9 * ia;
  const cd* const b3b = buf + 9 * ib;
  const cd* const t2 = tw2 + j2;
  __syncthreads();

  // Work items are drawn from the device counter ONE ITEM AHEAD: thread 0 requests the next item when the current one
  // starts, so the ~1 us atomic round trip hides behind the item's frames, and the CTA knows early enough which samples
  // to pull towards its L1 for the next item's first frame (the only frame that loads all 2 x 2048 samples).
#ifndef SSR_K1_ITEM_AHEAD
  // (drawing the items one ahead -- SSR_K1_ITEM_AHEAD -- hides the atomic round trip and allows prefetching the next
  // item's first frame, but measured 2 % slower: the extra live state costs more than the latency it hides)
  __shared__ int item_slot1;
  for (int item = next_work_item(next_item, &item_slot1); item < n_items; item = next_work_item(next_item, &item_slot1)) {
#else
  __shared__ int item_slot[2];
  if (tid == 0) item_slot[0] = atomicAdd(next_item, 1);
  __syncthreads();
  int item_par = 0;
  for (int item = item_slot[0]; item < n_items; item = item_slot[item_par ^= 1]) {
    if (tid == 0) item_slot[item_par ^ 1] = atomicAdd(next_item, 1);  // read after this item's barriers
#endif
    const int p = item_pair[item];
    const int c = item - item_start[p];
    const long long off = offsets[p];
    const long long L = offsets[p + 1] - off;
    const long long T = stft_frames(L, N, hop);
    const long long f0 = (long long)c * chunk;
    const int nf = (int)min((long long)chunk, T - f0);
    const ET* xe = est + off;
    const float* xt = tgt + off;
    double s_et = 0, s_tt = 0, s_ee = 0, l_et = 0, l_tt = 0, l_ee = 0;
    float* pend_t = nullptr;
    float* pend_e = nullptr;
    const SpecLayout sl = spec_layout(spec_e, spec_t, F);
    // staged magnitude rows -> global, consecutive threads to consecutive bins.  (ncu shows the stores of this loop
    // waiting for their shared-memory loads, 8 % of the warp time of <15>; issuing the loads of four bins ahead of
    // their stores was measured SLOWER -- 5.68 vs 5.62 ms for all four metrics -- and removed again.)
    auto copy_out_rows = [&]() {
      if (sl.step == 2) {  // interleaved (estimate, target) pairs: one 8-byte store per bin
        for (int k = tid; k < F; k += kV2Threads)
          reinterpret_cast<float2*>(pend_e)[k] = make_float2(row_e[k + (k >> 4)], row_t[k + (k >> 4)]);
      } else {
        for (int k = tid; k < F; k += kV2Threads) {
          pend_t[k] = row_t[k + (k >> 4)];
          if (pend_e) pend_e[k] = row_e[k + (k >> 4)];
        }
      }
    };

    long long ring_next = -1;  // RING: frame whose 12 older sample blocks sit in the tensor-memory ring
    float pre_t[4];  // RING: the frame's 4 new samples per signal
    ET pre_e[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      pre_t[i] = 0.f;
      pre_e[i] = (ET)0;
    }
    for (int fi = 0; fi < nf; ++fi) {
      const long long f = f0 + fi;
      const long long start = f * hop - N / 2;
      cd v[16];
      // ---- pass 1: load + window, radix-16, twiddle, store
      if (start >= 0 && start + N <= L) {
        const float* pt = xt + start + tid;
        const ET* pe = xe + start + tid;
        if (RING) {
          // sample block r of frame f is global block 4 f - 8 + r (blocks of 128 samples); it lives in ring
          // chunk (f + 2 + r / 4) mod 4 (4 blocks x (target, est) x float64 = 16 columns per chunk)
          const unsigned ring = tmem_base + 64;
          const int c0 = (int)((f + 2) & 3);
          // the 16 half-window values are requested first (L1 hits, but ~40+ cycles): issued behind the tensor-memory
          // traffic they left the FP64 multiplies below waiting on the long scoreboard (ncu: ~3 % of all warp time)
#ifndef SSR_K1_WIN_AHEAD
#define SSR_K1_WIN_AHEAD 16
#endif
          constexpr int kWinAhead = SSR_K1_WIN_AHEAD;  // window values requested ahead of the tensor-memory wait
          if (ring_next == f) {
#ifndef SSR_K1_PRE_REGS
            // the 4 new samples per signal: their lines were pulled into L1 during the previous frame (prefetch below),
            // so they are fetched here, next to the tensor-memory and window loads, instead of riding in 8 registers
            // through passes 2 and 3 of the previous frame (which spilled once the twiddle prefetch needed registers)
#pragma unroll
            for (int i = 0; i < 4; ++i) {
              pre_t[i] = ldg_f32_pinned(pt + 128 * (12 + i));
              pre_e[i] = ldg_pinned(pe + 128 * (12 + i));
            }
#endif
            tmem_wait_st();  // the previous frame's ring stores
            unsigned r0[16], r1[16], r2[16], r3[16];
            tmem_ld16(ring + 16 * ((c0 + 0) & 3), r0);
            tmem_ld16(ring + 16 * ((c0 + 1) & 3), r1);
            tmem_ld16(ring + 16 * ((c0 + 2) & 3), r2);
            // the first half-window values are requested now (L1 hits, but ~40+ cycles): all 16 issued behind the
            // tensor-memory traffic left the FP64 multiplies below waiting on the long scoreboard (ncu: ~3 % of all
            // warp time); the register file does not hold all 16 next to the three TMEM chunks
            double wv[kWinAhead > 0 ? kWinAhead : 1];
#pragma unroll
            for (int r = 0; r < kWinAhead; ++r) wv[r] = ldg_f64_pinned(P.win_half + tid + 128 * r);
#pragma unroll
            for (int i = 0; i < 4; ++i)
              v[12 + i] = cd{(double)pre_t[i], (double)pre_e[i]};
            tmem_pack4(v + 12, r3);
            tmem_st16(ring + 16 * ((c0 + 3) & 3), r3);
            tmem_wait_ld(r0);
            tmem_pin(r1);
            tmem_pin(r2);
            tmem_unpack4(r0, v);
            tmem_unpack4(r1, v + 4);
            tmem_unpack4(r2, v + 8);
#pragma unroll
            for (int r = 0; r < 16; ++r) {
              const double w = r < kWinAhead ? wv[r] : __ldg(P.win_half + tid + 128 * r);
              v[r] = cd{w * v[r].x, w * v[r].y};
            }
          } else {
#pragma unroll
            for (int r = 0; r < 16; ++r) v[r] = cd{(double)__ldg(pt + 128 * r), (double)__ldg(pe + 128 * r)};
#pragma unroll
            for (int k = 0; k < 4; ++k) {
              unsigned rr[16];
              tmem_pack4(v + 4 * k, rr);
              tmem_st16(ring + 16 * ((c0 + k) & 3), rr);
            }
#pragma unroll
            for (int r = 0; r < 16; ++r) {
              const double w = __ldg(P.win_half + tid + 128 * r);
              v[r] = cd{w * v[r].x, w * v[r].y};
            }
          }
          ring_next = f + 1;
        } else {
#pragma unroll
          for (int r = 0; r < 16; ++r) {
            const double w = __ldg(P.win_half + tid + 128 * r);
            v[r] = cd{w * (double)__ldg(pt + 128 * r), w * (double)__ldg(pe + 128 * r)};
          }
        }
        if (tid < 32) {
          // next frame's new samples: [start + N, start + N + hop) of both signals, one 128 B line per lane
          const long long nxt = start + N + (long long)(tid & 15) * 32;
          if (nxt < L && (tid & 15) * 32 < hop) {
            if (tid < 16) prefetch_l1(xt + nxt);
            else prefetch_l1(xe + nxt);  // (a float64 estimate: every other line; the rest come in with the loads)
          }
        }
      } else {
        // edge frame (reflect padding; < 1 % of the frames): gather through a small staging array so
        // the 64-bit reflect arithmetic stays out of the unrolled hot path
        __syncthreads();  // the staging area aliases buf: the previous frame's pass-3 loads must be done
        if (E64) {
          // (float64 estimate: no float2 staging; one rolled gather loop, the element selected by a uniform compare)
#pragma unroll 1
          for (int r = 0; r < 16; ++r) {
            const long long idx = reflect_index(start + tid + 128 * r, L);
            const double w = __ldg(P.win_half + tid + 128 * r);
            const cd val{w * (double)__ldg(xt + idx), w * (double)__ldg(xe + idx)};
#pragma unroll
            for (int rr = 0; rr < 16; ++rr)
              if (rr == r) v[rr] = val;
          }
          __syncthreads();
        } else {
#pragma unroll 1
          for (int n = tid; n < N; n += kV2Threads) {
            const long long idx = reflect_index(start + n, L);
            edge_raw[n] = make_float2(__ldg(xt + idx), (float)__ldg(xe + idx));
          }
          __syncthreads();
#pragma unroll
          for (int r = 0; r < 16; ++r) {
            const double w = __ldg(P.win_half + tid + 128 * r);
            const float2 x = edge_raw[tid + 128 * r];
            v[r] = cd{w * (double)x.x, w * (double)x.y};
          }
        }
      }
      bfly16<false>(v);
      {  // twiddles from tensor memory, four chunks of four, the next chunk in flight while one is applied
        unsigned r[2][16];
        tmem_ld16(tmem_base, r[0]);
#pragma unroll
        for (int c = 0; c < 4; ++c) {
          tmem_wait_ld(r[c & 1]);
          if (c < 3) tmem_ld16(tmem_base + 16 * (c + 1), r[(c + 1) & 1]);
          cd w4[4];
          tmem_unpack4(r[c & 1], w4);
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            const int q = 4 * c + i + 1;
            if (q < 16) v[q] = cmul(v[q], w4[i]);
          }
        }
      }
      // the previous frame's pass-3 loads must be done before buf is overwritten; placed here (after
      // this frame's loads and butterfly) the barrier finds every warp long past that point
      __syncthreads();
      if (pend_t) {  // coalesced copy-out of the previous frame's magnitude rows (all epilogues are done)
        copy_out_rows();
        pend_t = nullptr;
      }
#pragma unroll
      for (int q = 0; q < 16; ++q) b1[144 * q] = v[q];
#ifdef SSR_K1_PRE_REGS
      if (RING) {
        const long long ns = start + hop;  // the next frame of this item, if it is an interior one
        if (fi + 1 < nf && start >= 0 && ns + N <= L) {
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            pre_t[i] = __ldg(xt + ns + tid + 128 * (12 + i));
            pre_e[i] = __ldg(xe + ns + tid + 128 * (12 + i));
          }
        }
      }
#endif
      __syncthreads();
#if defined(SSR_K1_ITEM_AHEAD) && !defined(SSR_K1_NO_NEXT_PREFETCH)
      if (fi == nf - 1) {
        // last frame of the item: pull the first frame of the NEXT item (known since this item began) towards L1 / L2,
        // one 128-byte line per thread (64 lines per signal)
        const int nxt_item = item_slot[item_par ^ 1];
        if (nxt_item < n_items) {
          const int np = item_pair[nxt_item];
          const long long noff = offsets[np];
          const long long nL = offsets[np + 1] - noff;
          const long long nstart = (long long)(nxt_item - item_start[np]) * chunk * hop - N / 2 + (long long)(tid & 63) * 32;
          if (nstart >= 0 && nstart < nL) {
            if (tid < 64) prefetch_l1(tgt + noff + nstart);
            else prefetch_l1(est + noff + nstart);
          }
        }
      }
#endif
      // ---- pass 2: sub-transforms of length 128 (stride 8)
#pragma unroll
      for (int r = 0; r < 16; ++r) v[r] = b2[9 * r];
#ifdef SSR_K1_PASS2_SERIAL
      bfly16<false>(v);
      b2[0] = v[0];
#pragma unroll
      for (int q = 1; q < 16; ++q) b2[9 * q] = cmul(v[q], t2[(q - 1) * 8]);
#else
      // last radix-4 stage group by group with the twiddles of the next group fetched ahead (fft_core.cuh)
      bfly16_first<false>(v);
      if (kTw2Tmem) {
        unsigned r[2][16];
        tmem_ld16(tmem_base + 64, r[0]);
#pragma unroll
        for (int g = 0; g < 4; ++g) {
          tmem_wait_ld(r[g & 1]);
          if (g < 3) tmem_ld16(tmem_base + 64 + 16 * (g + 1), r[(g + 1) & 1]);
          cd w4[4];
          tmem_unpack4(r[g & 1], w4);
          bfly4<false>(v[4 * g], v[4 * g + 1], v[4 * g + 2], v[4 * g + 3]);
#pragma unroll
          for (int q1 = 0; q1 < 4; ++q1) {
            const int q = g + 4 * q1;
            b2[9 * q] = q > 0 ? cmul(v[4 * g + q1], w4[q1]) : v[0];
          }
        }
      } else {
        bfly16_second_twiddled<8>(v, t2, [&](int q, cd val) { b2[9 * q] = val; });
      }
#endif
#ifdef SSR_WARPLOCAL
      __syncwarp();
#else
      __syncthreads();
#endif
      // ---- pass 3: two radix-8 butterflies (a and its Hermitian partner b), no twiddles
      cd* a = v;
      cd* b = v + 8;
#pragma unroll
      for (int r = 0; r < 8; ++r) {
        a[r] = b3a[r];
        b[r] = b3b[r];
      }
      bfly8<false>(a);
      bfly8<false>(b);
      // ---- epilogue, from registers
      LT lsd_acc = 0;
      float* st = spec_t ? spec_t + spec_off[p] + f * sl.pitch : nullptr;
      float* se = spec_e ? spec_e + spec_off[p] + f * sl.pitch : nullptr;
      auto emit = [&](int k, cd zk, cd zn) {
        // T = (Z[k] + conj Z[N-k]) / 2,  E = (Z[k] - conj Z[N-k]) / (2i); the 1/2 is in the window.
        // complex64 rounding as librosa stores it, then float32 arithmetic as torch runs it; the
        // special functions are the hardware approximations (MUFU sqrt / rcp / lg2, <= 2 ulp), well
        // inside the differences that already exist between numpy's hypotf / torch's log10 and any
        // other libm (SSR_EXACT_F32_EPILOGUE switches to the IEEE-rounded forms for A/B tests).
        const float tre = (float)(zk.x + zn.x), tim = (float)(zk.y - zn.y);
        const float tx = tre * tre + tim * tim;  // |T|^2
        if (E64) {  // float64 estimate: the arithmetic of k1_generic.cuh's E64 branch
          const float mt = sqrtf(tx);
          const double me = hypot(zk.y + zn.y, zn.x - zk.x);  // np.abs(complex128)
          if (st) {
            row_t[k + (k >> 4)] = mt;
            row_e[k + (k >> 4)] = (float)me;
          }
          if (want_lsd) {
            const double den = me + 1e-12;
            const double l = log10((double)(mt * mt) / (den * den) + 1e-12);  // target ** 2 is float32
            lsd_acc += l * l;
          }
          if (want_lin) {
            const double dt = (double)mt;
            s_et = fma(me, dt, s_et);
            s_tt = fma(dt, dt, s_tt);
            s_ee = fma(me, me, s_ee);
          }
          if (want_log) {
            const double le = log10(me + 1e-12), lt = (double)log10f(mt + 1e-12f);
            l_et = fma(le, lt, l_et);
            l_tt = fma(lt, lt, l_tt);
            l_ee = fma(le, le, l_ee);
          }
          return;
        }
        const float ere = (float)(zk.y + zn.y), eim = (float)(zn.x - zk.x);
        const float ey = ere * ere + eim * eim;  // |E|^2
#ifdef SSR_EXACT_F32_EPILOGUE
        const float mt = sqrtf(tx), me = sqrtf(ey);
#else
        const float me = __fsqrt_approx(ey);
        const float mt = (st || want_lin || want_log) ? __fsqrt_approx(tx) : 0.f;
#endif
        if (st) {  // staged through shared memory (slot k + k/16: conflict-free for the scattered k of a warp)
          row_t[k + (k >> 4)] = mt;
          row_e[k + (k >> 4)] = me;
        }
        if (want_lsd) {
          const float den = me + 1e-12f;
#ifdef SSR_EXACT_F32_EPILOGUE
          const float l = log10f((mt * mt) / (den * den) + 1e-12f);
#else
          const float l = __log10f(__fdividef(tx, den * den) + 1e-12f);
#endif
          lsd_acc += l * l;
        }
        if (want_lin) {
          const double de = (double)me, dt = (double)mt;
          s_et = fma(de, dt, s_et);
          s_tt = fma(dt, dt, s_tt);
          s_ee = fma(de, de, s_ee);
        }
        if (want_log) {
#ifdef SSR_EXACT_F32_EPILOGUE
          const double le = (double)log10f(me + 1e-12f), lt = (double)log10f(mt + 1e-12f);
#else
          const double le = (double)__log10f(me + 1e-12f), lt = (double)__log10f(mt + 1e-12f);
#endif
          l_et = fma(le, lt, l_et);
          l_tt = fma(lt, lt, l_tt);
          l_ee = fma(le, le, l_ee);
        }
      };
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        const cd za = special ? a[(8 - q) & 7] : b[7 - q];
        const cd zb = special ? b[7 - q] : a[7 - q];
        emit(ka + 256 * q, a[q], za);
        emit(kb + 256 * q, b[q], zb);
      }
      if (special) emit(1024, a[4], a[4]);
      if (want_lsd) {
        const LT w = warp_sum(lsd_acc);
        if (lane == 0) lsd_part[fi][warp] = w;
      }
      pend_t = st;  // copied out after the next barrier (next frame's pass 1, or the item epilogue)
      pend_e = se;
    }
    __syncthreads();
    if (pend_t) {
      copy_out_rows();
      pend_t = nullptr;
    }
    // ---- per-item reduction -> partials[item][0..7]
    double lsd_sum = 0.0;
    if (want_lsd && tid < nf) {
      LT sacc = 0;
#pragma unroll
      for (int w = 0; w < NW; ++w) sacc += lsd_part[tid][w];
      // torch.mean(dim=3) ** 0.5 in float32 (float64 when the estimate is float64)
      lsd_sum = E64 ? sqrt((double)sacc / (double)F) : (double)sqrtf((float)sacc / (float)F);
    }
    double vals[7] = {lsd_sum, s_et, s_tt, s_ee, l_et, l_tt, l_ee};
    // only the sums the compile-time metric set fills are reduced (LSD-only: 1 of 7; the others stay 0)
    constexpr int kFirst = 0, kLast = (FIXED >= 0 && !(FIXED & 6)) ? 1 : ((FIXED >= 0 && !(FIXED & 2)) ? 4 : 7);
#pragma unroll
    for (int i = kFirst; i < 7; ++i) {
      const double r = i < kLast ? warp_sum(vals[i]) : 0.0;
      if (lane == 0) red[warp][i] = r;
    }
    __syncthreads();
    if (tid < 7) {
      double r = 0.0;
#pragma unroll
      for (int w = 0; w < NW; ++w) r += red[w][tid];
      partials[(size_t)item * kPartials + tid] = r;
    }
    __syncthreads();
  }
  __syncthreads();
  tmem_free<kTmemCols>(&tmem_slot, warp);
}


}  // namespace ssr
