// k1_generic.cuh -- K1, generic kernel: any power-of-two n_fft (direct) or Bluestein with M <= 8192 (radix-8/4 in-place FFT).
#pragma once
#include <type_traits>
#include "k1_common.cuh"

namespace ssr {

// ---------------------------------------------------------------------------------------------
// K1: one CTA walks the frames of its work items; per frame:
//   load (window folded in) -> forward FFT in shared memory [-> Bluestein filter -> inverse FFT]
//   -> separate the two spectra -> complex64 rounding -> float32 magnitudes -> metric terms.
//
// ET = double: the ESTIMATE is a float64 waveform (what the reference's IIR low-pass filters hand to an
// identity-like testee: scipy's sosfiltfilt returns float64).  librosa then keeps the estimate's spectrum
// in complex128 / its magnitude in float64 (dtype_r2c), and every torch formula that mixes it with the
// float32 target promotes to float64 (metrics.py:109-121).  Reproduced here: E stays float64 from the
// waveform to the sums; only T is rounded to complex64 / float32.  (SSIM gets the float32-rounded |E|:
// skimage would run in float64, a ~1e-7 effect against the 1e-3 tolerance.)
// ---------------------------------------------------------------------------------------------
// TT = double (only together with ET = double): the TARGET is a float64 waveform too (e.g. soundfile.read's default
// dtype handed straight to AudioMetrics.evaluation): librosa keeps both spectra in complex128 and every torch formula
// runs in float64 -- here T stays float64 as well.
template <int LOGM, bool BLUE, typename ET, typename TT = float>
__global__ void __launch_bounds__(kThreads)
k_stft_metrics(StftDev P, const ET* __restrict__ est, const TT* __restrict__ tgt,
               const long long* __restrict__ offsets, const int* __restrict__ item_start,
               const int* __restrict__ item_pair, int n_items, int chunk, unsigned flags,
               double* __restrict__ partials, float* __restrict__ spec_e,
               float* __restrict__ spec_t, const long long* __restrict__ spec_off, int* __restrict__ next_item) {
  constexpr int M = 1 << LOGM;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  cd* buf = reinterpret_cast<cd*>(smem_raw);
  constexpr bool E64 = sizeof(ET) == 8;
  constexpr bool T64 = sizeof(TT) == 8;
  static_assert(!T64 || E64, "a float64 target is scored together with a float64 estimate");
  using LT = typename std::conditional<E64, double, float>::type;  // type of the per-frame LSD sums
  __shared__ LT lsd_part[kMaxChunk][kWarps];
  __shared__ double red[kWarps][kPartials];

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int N = P.n_fft, F = P.F, hop = P.hop;
  const bool want_lsd = flags & SSR_METRIC_LSD, want_log = flags & SSR_METRIC_LOG_SISPEC,
             want_lin = flags & SSR_METRIC_SISPEC;
  const SpecLayout sl = spec_layout(spec_e, spec_t, F);

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
    const TT* xt = tgt + off;
    double s_et = 0, s_tt = 0, s_ee = 0, l_et = 0, l_tt = 0, l_ee = 0;

    for (int fi = 0; fi < nf; ++fi) {
      const long long f = f0 + fi;
      const long long start = f * hop - N / 2;
      // ---- load: z[n] = 0.5*w[n]*(target + i*est)  (Bluestein: times the chirp, zero padded)
      if (!BLUE) {
        for (int n = tid; n < M; n += kThreads) {
          long long idx = reflect_index(start + n, L);
          double w = P.win_half[n];
          buf[pad_idx(n)] = cd{w * (double)__ldg(xt + idx), w * (double)__ldg(xe + idx)};
        }
      } else {
        for (int n = tid; n < M; n += kThreads) {
          cd v{0.0, 0.0};
          if (n < N) {
            long long idx = reflect_index(start + n, L);
            double t = (double)__ldg(xt + idx), e = (double)__ldg(xe + idx);
            cd w = P.cw[n];
            v = cd{t * w.x - e * w.y, t * w.y + e * w.x};
          }
          buf[pad_idx(n)] = v;
        }
      }
      __syncthreads();
      fft_forward_dif<LOGM>(buf, P.tw, tid, kThreads, SyncThreads());
      __syncthreads();
      if (BLUE) {
        for (int i = tid; i < M; i += kThreads) buf[pad_idx(i)] = cmul(buf[pad_idx(i)], P.bfilt[i]);
        __syncthreads();
        fft_inverse_dit<LOGM>(buf, P.tw, tid, kThreads, SyncThreads());
        __syncthreads();
      }
      // ---- epilogue over the F = n_fft/2+1 bins
      LT lsd_acc = 0;
      float* se = spec_e ? spec_e + spec_off[p] + f * sl.pitch : nullptr;
      float* st = spec_t ? spec_t + spec_off[p] + f * sl.pitch : nullptr;
      for (int k = tid; k < F; k += kThreads) {
        cd a, b;
        if (!BLUE) {
          a = buf[P.ppos[k]];
          b = buf[P.ppos[(N - k) & (N - 1)]];
        } else {
          int k2 = k ? N - k : 0;
          a = cmul(buf[pad_idx(k)], P.cpost[k]);
          b = cmul(buf[pad_idx(k2)], P.cpost[k2]);
        }
        // T = (Z[k] + conj Z[N-k]) / 2,  E = (Z[k] - conj Z[N-k]) / (2i); the 1/2 is in the window
        float tre = (float)(a.x + b.x), tim = (float)(a.y - b.y);
        float mt = sqrtf(tre * tre + tim * tim);
        if (E64) {
          const double me = hypot(a.y + b.y, b.x - a.x);  // np.abs(complex128)
          const double mt64 = T64 ? hypot(a.x + b.x, a.y - b.y) : (double)mt;
          spec_store(sl, st, se, k, T64 ? (float)mt64 : mt, (float)me);
          if (want_lsd) {
            const double den = me + 1e-12;
            // target ** 2 is float32 for a float32 target, float64 for a float64 one
            const double l = log10((T64 ? mt64 * mt64 : (double)(mt * mt)) / (den * den) + 1e-12);
            lsd_acc += l * l;
          }
          if (want_lin) {
            const double dt = mt64;
            s_et = fma(me, dt, s_et);
            s_tt = fma(dt, dt, s_tt);
            s_ee = fma(me, me, s_ee);
          }
          if (want_log) {
            const double le = log10(me + 1e-12), lt = T64 ? log10(mt64 + 1e-12) : (double)log10f(mt + 1e-12f);
            l_et = fma(le, lt, l_et);
            l_tt = fma(lt, lt, l_tt);
            l_ee = fma(le, le, l_ee);
          }
          continue;
        }
        float ere = (float)(a.y + b.y), eim = (float)(b.x - a.x);
        float me = sqrtf(ere * ere + eim * eim);
        spec_store(sl, st, se, k, mt, me);
        if (want_lsd) {
          float den = me + 1e-12f;
          float q = (mt * mt) / (den * den) + 1e-12f;
          float l = log10f(q);
          lsd_acc += l * l;
        }
        // sispec is evaluated in closed form from three sums (finalize); at 40+ dB the difference
        // S_ee - S_et^2/S_tt cancels 4+ digits, so the products (exact in float64) are summed in float64.
        if (want_lin) {
          const double de = (double)me, dt = (double)mt;
          s_et = fma(de, dt, s_et);
          s_tt = fma(dt, dt, s_tt);
          s_ee = fma(de, de, s_ee);
        }
        if (want_log) {
          const double le = (double)log10f(me + 1e-12f), lt = (double)log10f(mt + 1e-12f);
          l_et = fma(le, lt, l_et);
          l_tt = fma(lt, lt, l_tt);
          l_ee = fma(le, le, l_ee);
        }
      }
      if (want_lsd) {
        const LT w = warp_sum(lsd_acc);
        if (lane == 0) lsd_part[fi][warp] = w;
      }
      __syncthreads();  // buf is rewritten by the next frame's load
    }

    // ---- per-item reduction -> partials[item][0..7]
    double lsd_sum = 0.0;
    if (want_lsd && tid < nf) {
      LT s = 0;
#pragma unroll
      for (int w = 0; w < kWarps; ++w) s += lsd_part[tid][w];
      // torch.mean(dim=3) ** 0.5 in float32 (float64 when the estimate is float64)
      lsd_sum = E64 ? sqrt((double)s / (double)F) : (double)sqrtf((float)s / (float)F);
    }
    double v[7] = {lsd_sum, s_et, s_tt, s_ee, l_et, l_tt, l_ee};
#pragma unroll
    for (int i = 0; i < 7; ++i) {
      double r = warp_sum(v[i]);
      if (lane == 0) red[warp][i] = r;
    }
    __syncthreads();
    if (tid < 7) {
      double r = 0.0;
#pragma unroll
      for (int w = 0; w < kWarps; ++w) r += red[w][tid];
      partials[(size_t)item * kPartials + tid] = r;
    }
    __syncthreads();
  }
}


}  // namespace ssr
