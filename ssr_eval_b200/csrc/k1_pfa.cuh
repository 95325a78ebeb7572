// k1_pfa.cuh -- K1, PFA kernel for n_fft = R * P (R Bluestein sub-transforms on the 2048-point machinery).
#pragma once
#include <type_traits>
#include "k1_common.cuh"
#include "k1_map.cuh"
#include "tmem.cuh"

namespace ssr {

// ---------------------------------------------------------------------------------------------
// K1 for non-power-of-two n_fft = R * P (2229 = 3 x 743 at 48 kHz -- the reference's own default --,
// 1114, 743, 1486 ...): R Bluestein sub-transforms of length P on the 2048-point radix 16x16x8
// machinery of k_stft_metrics_2048 (forward DIF, filter multiply and inverse butterfly of the last /
// first pass in registers, inverse DIT), recombined with a radix-R butterfly on the fly in the
// epilogue.  ~6 FFT-2048 per frame instead of 2 FFT-8192 (1.6x fewer flops, 2x less shared traffic
// than the generic Bluestein kernel).  NQ = ceil(P / 128): pass-1 inputs / pass-3' outputs beyond
// NQ are structurally zero / unused and are pruned at compile time.
// ---------------------------------------------------------------------------------------------
// 3 CTAs/SM (168 registers): the pass-1 twiddles and the Bluestein filter values of the thread's two
// butterflies live in tensor memory (tmem.cuh), and the samples + window*chirp factors of the NEXT
// sub-transform are fetched into registers one sub-transform ahead -- the tables (103 KB) do not fit the L1
// left beside the CTAs' shared memory, so every table load is an L2 round trip that has to be hidden in
// software.  Measured (LSD only, 256 pairs x 5 s): 21.6k pairs/s (3 CTAs, loads in place) -> 25.7k (2 CTAs,
// 255 registers, constants in registers, software pipeline) -> 31.2k (3 CTAs, constants in TMEM, pipeline).
// ET = double: the ESTIMATE is a float64 waveform (the reference's IIR low-pass keys, see k1_generic.cuh): E stays
// float64 from the waveform to the sums -- np.abs(complex128), float64 log10 / products -- only T is rounded to
// complex64 / float32; same arithmetic as the generic kernel's float64-estimate path, 4x its speed.
template <int NQ, int FIXED, typename ET = float>
#ifndef SSR_PFA_CTAS
#define SSR_PFA_CTAS 3  // CTAs per SM of the PFA kernels (168 registers; 2 CTAs with the ~220 registers ptxas then
                        // takes and no spill: 32.2 k pairs/s against 37.3 k -- the warps matter more)
#endif
__global__ void __launch_bounds__(kV2Threads, SSR_PFA_CTAS)
k_stft_metrics_pfa(PfaDev D, const ET* __restrict__ est, const float* __restrict__ tgt,
                   const long long* __restrict__ offsets, const int* __restrict__ item_start,
                   const int* __restrict__ item_pair, int n_items, int chunk, unsigned flags,
                   double* __restrict__ partials, float* __restrict__ spec_e,
                   float* __restrict__ spec_t, const long long* __restrict__ spec_off, int* __restrict__ next_item) {
  constexpr int M = 2048, NW = kV2Threads / 32;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  cd* const buf = reinterpret_cast<cd*>(smem_raw);                       // M + M/8 slots
  // Y_r, r < R-1, live behind buf; the LAST sub-transform's Y is written over buf itself (dead by then),
  // which keeps the CTA at ~65 KB of shared memory = 3 CTAs per SM
  cd* const Yx = reinterpret_cast<cd*>(smem_raw + sizeof(cd) * (M + M / 8));
  __shared__ __align__(16) cd tw2[15 * 8];
  __shared__ __align__(16) cd wr_s[16];
  constexpr bool E64 = sizeof(ET) == 8;
  using LT = typename std::conditional<E64, double, float>::type;  // type of the per-frame LSD sums
  __shared__ LT lsd_part[kMaxChunk][NW];
  __shared__ double red[NW][kPartials];

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int N = D.n_fft, F = D.F, hop = D.hop, R = D.R, P = D.P;
  if (FIXED >= 0) flags = (unsigned)FIXED;
  const bool want_lsd = flags & SSR_METRIC_LSD, want_log = flags & SSR_METRIC_LOG_SISPEC,
             want_lin = flags & SSR_METRIC_SISPEC;
  if (FIXED >= 0) {
    spec_e = nullptr;
    spec_t = nullptr;
  }
  const SpecLayout sl = spec_layout(spec_e, spec_t, F);

  __shared__ unsigned tmem_slot;
  if (tid < 120) tw2[tid] = D.tw[16 * (tid & 7) * ((tid >> 3) + 1)];
  if (tid < R * R) wr_s[tid] = D.wr[tid];
  int ia, ib;
#ifdef SSR_WARPLOCAL
  // forward passes 2 / 3 and inverse passes 1 / 2 of a sub-transform block and of its Hermitian-partner block run in
  // one warp (k1_map.cuh): those two exchanges are warp-local, __syncwarp() instead of a CTA barrier
  bool special_unused;
  v2w_thread_butterflies(tid, &ia, &ib, &special_unused);
  const int blk2 = v2w_pass2_block(tid);
#define SSR_PFA_SYNC_LOCAL() __syncwarp()
#else
  v2_thread_butterflies(tid, &ia, &ib);
  const int blk2 = tid >> 3;
#define SSR_PFA_SYNC_LOCAL() __syncthreads()
#endif
  const int j2 = tid & 7;
  cd* const b1 = buf + pad_idx(tid);
  cd* const b2 = buf + pad_idx(blk2 * 128 + j2);
  cd* const b3a = buf + 9 * ia;
  cd* const b3b = buf + 9 * ib;
  const cd* const t2 = tw2 + j2;
  __syncthreads();

  // Per-thread constants live in TENSOR MEMORY, this thread's lane: columns [0, 60) the pass-1 twiddles
  // W^{tid q}, [64, 96) the Bluestein filter at butterfly a's 8 slots, [96, 128) at butterfly b's -- 124
  // registers' worth that the register file no longer has to hold (written once with tcgen05.st).
  const unsigned tmem_base = tmem_alloc<128>(&tmem_slot, warp);
#pragma unroll
  for (int c = 0; c < 8; ++c) {
    cd w4[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      if (c < 4) {
        const int q = 4 * c + i + 1;
        w4[i] = q < 16 ? D.tw[tid * q] : cd{0.0, 0.0};
      } else {
        const int q = 4 * (c & 1) + i;
        w4[i] = D.bfilt[8 * (c < 6 ? ia : ib) + q];
      }
    }
    unsigned r[16];
    tmem_pack4(w4, r);
    tmem_st16(tmem_base + 16 * c, r);
  }
  tmem_wait_st();
  // v[q] *= W^{tid q} (or its conjugate), q = 1..15: four chunks of four twiddles from tensor memory, the next
  // chunk in flight while one is applied
  auto apply_tw1 = [&](cd* v, auto conj_tag) {
    constexpr bool CONJ = decltype(conj_tag)::value;
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
        if (q < 16) v[q] = CONJ ? cmul_conj(v[q], w4[i]) : cmul(v[q], w4[i]);
      }
    }
  };

  auto combine = [&](int kap) {  // Z[kap] = sum_r W_R^{r m} Y_r[k], kap = k + P m
    int m = 0;
    while (kap >= P) {
      kap -= P;
      ++m;
    }
    cd z = (R > 1) ? Yx[kap] : buf[kap];
    if (R > 1) {
      z = cmul(z, wr_s[m]);
      for (int r = 1; r < R; ++r) {
        const cd yv = (r == R - 1) ? buf[kap] : Yx[r * P + kap];
        z = cadd(z, cmul(yv, wr_s[r * R + m]));
      }
    }
    return z;
  };

  __shared__ int item_slot;
  for (int item = next_work_item(next_item, &item_slot); item < n_items; item = next_work_item(next_item, &item_slot)) {
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

    // inputs of sub-transform (f, r): samples of both signals and window*chirp, NQ per thread
    // The samples (stride-R gathers) are held in registers from one sub-transform to the next; their window * chirp
    // factors are only pulled towards L1 and loaded where they are used.  Measured (gpurun sessions s36 / s37): holding
    // the factors in registers too (24 more, ptxas spills) 36.2 k pairs/s, this form 37.3 k, no register prefetch at
    // all 32.6 k.
    float ptx[NQ];
    ET pex[NQ];
    auto fetch_inputs = [&](long long f, int r) {
      const long long start = f * hop - N / 2;
      const bool interior = (start >= 0 && start + N <= L);
#pragma unroll
      for (int q = 0; q < NQ; ++q) {
        const int n = tid + 128 * q;
        ptx[q] = 0.f;
        pex[q] = (ET)0;
        if (n < P) {
          const long long si = start + (long long)R * n + r;
          const long long idx = interior ? si : reflect_index(si, L);
          ptx[q] = __ldg(xt + idx);
          pex[q] = __ldg(xe + idx);
          prefetch_l1(D.cwin + r * P + n);
        }
      }
    };
    fetch_inputs(f0, 0);

    for (int fi = 0; fi < nf; ++fi) {
      const long long f = f0 + fi;
      for (int r = 0; r < R; ++r) {
        cd v[16];
        // ---- forward pass 1: a[n] = z[R n + r] * (0.5 window * chirp), zero padded to 2048
#pragma unroll
        for (int q = 0; q < 16; ++q) {
          v[q] = cd{0.0, 0.0};
          if (q < NQ) {
            const double tt = (double)ptx[q], ee = (double)pex[q];
            const int n = tid + 128 * q;
            const cd cw = (n < P) ? D.cwin[r * P + n] : cd{0.0, 0.0};
            v[q] = cd{tt * cw.x - ee * cw.y, tt * cw.y + ee * cw.x};
          }
        }
#ifdef SSR_PFA_FETCH_EARLY
        if (r + 1 < R) fetch_inputs(f, r + 1);
        else if (fi + 1 < nf) fetch_inputs(f + 1, 0);
#endif
        bfly16<false>(v);
        apply_tw1(v, std::false_type{});
        __syncthreads();  // previous sub-transform's last loads are done
#pragma unroll
        for (int q = 0; q < 16; ++q) b1[144 * q] = v[q];
#ifndef SSR_PFA_FETCH_EARLY
        // The loads of the next sub-transform fly during the remaining five passes of this one.  Issued here, where
        // v[] has just been stored and nothing else is live, not before pass 1: there the prefetch registers met
        // the butterfly and the tensor-memory twiddle buffers, ptxas spilled four of the just-loaded values and each
        // spill store waited for its own load (ncu: 5.8 % of the warp time on STL / LDL)
        if (r + 1 < R) fetch_inputs(f, r + 1);
        else if (fi + 1 < nf) fetch_inputs(f + 1, 0);
#endif
        __syncthreads();
        // ---- forward pass 2
#pragma unroll
        for (int q = 0; q < 16; ++q) v[q] = b2[9 * q];
#ifdef SSR_PFA_PASS2_SERIAL
        bfly16<false>(v);
        b2[0] = v[0];
#pragma unroll
        for (int q = 1; q < 16; ++q) b2[9 * q] = cmul(v[q], t2[(q - 1) * 8]);
#else
        // last radix-4 stage group by group with the twiddles of the next group fetched ahead (fft_core.cuh)
        bfly16_first<false>(v);
        bfly16_second_twiddled<8>(v, t2, [&](int q, cd val) { b2[9 * q] = val; });
#endif
        SSR_PFA_SYNC_LOCAL();
        // ---- forward pass 3, Bluestein filter, inverse pass 1: all in registers
        cd* a = v;
        cd* b = v + 8;
#pragma unroll
        for (int q = 0; q < 8; ++q) {
          a[q] = b3a[q];
          b[q] = b3b[q];
        }
        {
          unsigned r[2][16];
          tmem_ld16(tmem_base + 64, r[0]);  // the filter values arrive while the butterflies run
          bfly8<false>(a);
          bfly8<false>(b);
#pragma unroll
          for (int c = 0; c < 4; ++c) {
            tmem_wait_ld(r[c & 1]);
            if (c < 3) tmem_ld16(tmem_base + 64 + 16 * (c + 1), r[(c + 1) & 1]);
            cd w4[4];
            tmem_unpack4(r[c & 1], w4);
#pragma unroll
            for (int i = 0; i < 4; ++i) v[4 * c + i] = cmul(v[4 * c + i], w4[i]);  // v = a[0..7], b[0..7]
          }
        }
        bfly8<true>(a);
        bfly8<true>(b);
#pragma unroll
        for (int q = 0; q < 8; ++q) {
          b3a[q] = a[q];
          b3b[q] = b[q];
        }
        SSR_PFA_SYNC_LOCAL();
        // ---- inverse pass 2 (the chirp * W_N^{rk} factors of the output stage are pulled towards L1 meanwhile:
        // they are L2 round trips otherwise, and holding them in registers across two passes would spill)
#pragma unroll
        for (int q = 0; q < NQ; ++q) {
          const int k = tid + 128 * q;
          if (k < P) prefetch_l1(D.post + r * P + k);
        }
        v[0] = b2[0];
#pragma unroll
        for (int q = 1; q < 16; ++q) v[q] = cmul_conj(b2[9 * q], t2[(q - 1) * 8]);
        bfly16<true>(v);
#pragma unroll
        for (int q = 0; q < 16; ++q) b2[9 * q] = v[q];
        __syncthreads();
        // ---- inverse pass 3 -> conv[k], k = tid + 128 q; Y_r[k] = conv[k] * chirp[k] * W_N^{rk}
#pragma unroll
        for (int q = 0; q < 16; ++q) v[q] = b1[144 * q];
        apply_tw1(v, std::true_type{});
        bfly16<true>(v);
        cd* Yr = Yx + r * P;
        if (r == R - 1) {
          __syncthreads();  // every thread has finished reading buf
          Yr = buf;
        }
#pragma unroll
        for (int q = 0; q < NQ; ++q) {
          const int k = tid + 128 * q;
          if (k < P) Yr[k] = cmul(v[q], D.post[r * P + k]);
        }
      }
      __syncthreads();
      // ---- epilogue over the F bins (recombination on the fly)
      LT lsd_acc = 0;
      float* st = spec_t ? spec_t + spec_off[p] + f * sl.pitch : nullptr;
      float* se = spec_e ? spec_e + spec_off[p] + f * sl.pitch : nullptr;
      for (int k = tid; k < F; k += kV2Threads) {
        const cd zk = combine(k);
        const cd zn = combine(k ? N - k : 0);
        const float tre = (float)(zk.x + zn.x), tim = (float)(zk.y - zn.y);
        const float tx = tre * tre + tim * tim;
        if (E64) {  // float64 estimate: the arithmetic of k1_generic.cuh's E64 branch
          const float mt = sqrtf(tx);
          const double me = hypot(zk.y + zn.y, zn.x - zk.x);  // np.abs(complex128)
          spec_store(sl, st, se, k, mt, (float)me);
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
          continue;
        }
        const float ere = (float)(zk.y + zn.y), eim = (float)(zn.x - zk.x);
        const float ey = ere * ere + eim * eim;
        const float me = __fsqrt_approx(ey);
        const float mt = (st || want_lin || want_log) ? __fsqrt_approx(tx) : 0.f;
        spec_store(sl, st, se, k, mt, me);
        if (want_lsd) {
          const float den = me + 1e-12f;
          const float l = __log10f(__fdividef(tx, den * den) + 1e-12f);
          lsd_acc += l * l;
        }
        if (want_lin) {
          const double de = (double)me, dt = (double)mt;
          s_et = fma(de, dt, s_et);
          s_tt = fma(dt, dt, s_tt);
          s_ee = fma(de, de, s_ee);
        }
        if (want_log) {
          const double le = (double)__log10f(me + 1e-12f), lt = (double)__log10f(mt + 1e-12f);
          l_et = fma(le, lt, l_et);
          l_tt = fma(lt, lt, l_tt);
          l_ee = fma(le, le, l_ee);
        }
      }
      if (want_lsd) {
        const LT w = warp_sum(lsd_acc);
        if (lane == 0) lsd_part[fi][warp] = w;
      }
    }
    __syncthreads();
    double lsd_sum = 0.0;
    if (want_lsd && tid < nf) {
      LT sacc = 0;
#pragma unroll
      for (int w = 0; w < NW; ++w) sacc += lsd_part[tid][w];
      // torch.mean(dim=3) ** 0.5 in float32 (float64 when the estimate is float64)
      lsd_sum = E64 ? sqrt((double)sacc / (double)F) : (double)sqrtf((float)sacc / (float)F);
    }
    double vals[7] = {lsd_sum, s_et, s_tt, s_ee, l_et, l_tt, l_ee};
#pragma unroll
    for (int i = 0; i < 7; ++i) {
      const double rr = warp_sum(vals[i]);
      if (lane == 0) red[warp][i] = rr;
    }
    __syncthreads();
    if (tid < 7) {
      double rr = 0.0;
#pragma unroll
      for (int w = 0; w < NW; ++w) rr += red[w][tid];
      partials[(size_t)item * kPartials + tid] = rr;
    }
    __syncthreads();
  }
  __syncthreads();
  tmem_free<128>(&tmem_slot, warp);
}


}  // namespace ssr
