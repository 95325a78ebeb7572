// stft_metrics.cu -- K1 (batched STFT -> magnitude -> LSD / sispec / log-sispec accumulators)
// and K2 (7x7 box-window SSIM) of the ssr_eval hot path, plus their C ABI.
//
// Reference semantics (paths relative to the reference repo):
//   ssr_eval/metrics.py:16-19   n_fft / hop from the sample rate
//   ssr_eval/metrics.py:26-30   |librosa.stft| : reflect pad, periodic Hann (f64), f64 FFT, c64
//   ssr_eval/metrics.py:109-112 lsd
//   ssr_eval/metrics.py:114-121 + ssr_eval/utils.py:68-92   sispec (energy_unify, pow_norm, ...)
//   ssr_eval/utils.py:43-44     to_log
//   ssr_eval/metrics.py:123-132 ssim -> skimage structural_similarity(win_size=7)
//
// Numerics: est and target of a pair are packed as ONE complex float64 transform
// (z = target + i*est), because the 1e-4 tolerance on LSD / log-sispec needs ~float64 FFT accuracy
// on hard-low-passed estimates (hard part 1).  After the transform the two
// spectra are separated, rounded to complex64 like librosa's store, and every metric formula runs
// in float32 exactly as torch does on the CPU; cross-frame sums are kept in float64.
#include <math.h>
#include <stdlib.h>
#include <string.h>

#include <vector>

#include "common.cuh"
#include "fft_core.cuh"
#include "stft_tables.hpp"
#include "k1_map.cuh"

namespace ssr {

std::string& last_error_ref() {
  static thread_local std::string s;
  return s;
}
std::atomic<uint64_t>& launch_counter() {
  static std::atomic<uint64_t> c{0};
  return c;
}

struct TimingState {
  bool on = false;
  std::vector<std::pair<cudaEvent_t, cudaEvent_t>> pending, pool;
};
static TimingState& timing() {
  static TimingState t;
  return t;
}

constexpr int kThreads = 256;
constexpr int kWarps = kThreads / 32;
constexpr int kMaxChunk = 64;   // frames per work item (upper bound)
constexpr int kPartials = 8;    // doubles per work item
constexpr int kSsimTR = 64;     // SSIM tile: output rows
constexpr int kSsimTC = 256;    // SSIM tile: output cols (2 per thread)

struct StftDev {
  int n_fft, hop, F, M;
  const cd* tw;            // exp(-2 pi i n / M)
  const double* win_half;  // direct: 0.5 * window[n]
  const uint16_t* ppos;    // direct: padded smem slot of frequency k after the DIF passes
  const cd* cw;            // bluestein: 0.5 * window[n] * chirp[n]
  const cd* bfilt;         // bluestein: FFT_M(conj chirp) / M in DIF (digit-reversed) order
  const cd* cpost;         // bluestein: chirp[k]
};

// tables of the PFA path (n_fft = R * P, P <= 1024; see stft_tables.hpp)
struct PfaDev {
  int n_fft, hop, F, R, P;
  const cd* tw;     // exp(-2 pi i n / 2048)
  const cd* cwin;   // [r*P + n] = 0.5 * window[R n + r] * chirp_P[n]
  const cd* post;   // [r*P + k] = chirp_P[k] * W_N^{rk}
  const cd* bfilt;  // Bluestein filter spectrum / 2048, DIF (16,16,8) order
  const cd* wr;     // [r*R + m] = W_R^{rm}
};

}  // namespace ssr

struct ssr_stft_plan {
  int n_fft, hop, F, M, logM, bluestein, device;
  int pfa;          // 1: PFA path (pdev valid)
  void* blob;       // one device allocation holding all tables
  void* blob_pfa;
  ssr::StftDev dev;
  ssr::PfaDev pdev;
};

namespace ssr {

__host__ __device__ inline long long stft_frames(long long L, int n_fft, int hop) {
  return 1 + (L + 2 * (n_fft / 2) - n_fft) / hop;
}

__device__ __forceinline__ long long reflect_index(long long i, long long L) {
  if (i >= 0 && i < L) return i;
  if (L == 1) return 0;
  long long period = 2 * (L - 1);
  i %= period;
  if (i < 0) i += period;
  return i < L ? i : period - i;
}

// ---------------------------------------------------------------------------------------------
// setup: work-item table.  item_start[p] = first work item of pair p (item = chunk of <= `chunk`
// consecutive frames), item_pair[item] = p, spec_off[p] = first spectrogram element of pair p.
// ---------------------------------------------------------------------------------------------
__global__ void k_setup(const long long* __restrict__ offsets, int n, int n_fft, int hop, int chunk,
                        int F, int* __restrict__ item_start, int* __restrict__ item_pair,
                        long long* __restrict__ spec_off) {
  __shared__ long long s_items[1024], s_frames[1024];
  const int t = threadIdx.x;
  const int per = (n + 1023) / 1024;
  const int lo = min(n, t * per), hi = min(n, lo + per);
  long long it = 0, fr = 0;
  for (int p = lo; p < hi; ++p) {
    long long T = stft_frames(offsets[p + 1] - offsets[p], n_fft, hop);
    it += (T + chunk - 1) / chunk;
    fr += T;
  }
  s_items[t] = it;
  s_frames[t] = fr;
  __syncthreads();
  if (t == 0) {
    long long a = 0, b = 0;
    for (int i = 0; i < 1024; ++i) {
      long long x = s_items[i], y = s_frames[i];
      s_items[i] = a;
      s_frames[i] = b;
      a += x;
      b += y;
    }
  }
  __syncthreads();
  it = s_items[t];
  fr = s_frames[t];
  for (int p = lo; p < hi; ++p) {
    long long T = stft_frames(offsets[p + 1] - offsets[p], n_fft, hop);
    int nc = (int)((T + chunk - 1) / chunk);
    item_start[p] = (int)it;
    spec_off[p] = fr * F;
    for (int c = 0; c < nc; ++c) item_pair[it + c] = p;
    it += nc;
    fr += T;
  }
  if (hi == n) {
    item_start[n] = (int)it;
    spec_off[n] = fr * F;
  }
}

__device__ __forceinline__ float __fsqrt_approx(float x) {
  float r;
  asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
  return r;
}

__device__ __forceinline__ void prefetch_l1(const void* p) {
  asm volatile("prefetch.global.L1 [%0];" ::"l"(p));
}

__device__ __forceinline__ void cp_async4(void* smem_dst, const void* gmem_src, int src_bytes) {
  const unsigned d = (unsigned)__cvta_generic_to_shared(smem_dst);
  asm volatile("cp.async.ca.shared.global [%0], [%1], 4, %2;" ::"r"(d), "l"(gmem_src), "r"(src_bytes) : "memory");
}

struct SyncThreads {
  __device__ __forceinline__ void operator()() const { __syncthreads(); }
};

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_down_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_down_sync(0xffffffffu, v, o);
  return v;
}

// ---------------------------------------------------------------------------------------------
// K1: one CTA walks the frames of its work items; per frame:
//   load (window folded in) -> forward FFT in shared memory [-> Bluestein filter -> inverse FFT]
//   -> separate the two spectra -> complex64 rounding -> float32 magnitudes -> metric terms.
// ---------------------------------------------------------------------------------------------
template <int LOGM, bool BLUE>
__global__ void __launch_bounds__(kThreads)
k_stft_metrics(StftDev P, const float* __restrict__ est, const float* __restrict__ tgt,
               const long long* __restrict__ offsets, const int* __restrict__ item_start,
               const int* __restrict__ item_pair, int n_items, int chunk, unsigned flags,
               double* __restrict__ partials, float* __restrict__ spec_e,
               float* __restrict__ spec_t, const long long* __restrict__ spec_off) {
  constexpr int M = 1 << LOGM;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  cd* buf = reinterpret_cast<cd*>(smem_raw);
  __shared__ float lsd_part[kMaxChunk][kWarps];
  __shared__ double red[kWarps][kPartials];

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int N = P.n_fft, F = P.F, hop = P.hop;
  const bool want_lsd = flags & SSR_METRIC_LSD, want_log = flags & SSR_METRIC_LOG_SISPEC,
             want_lin = flags & SSR_METRIC_SISPEC;

  for (int item = blockIdx.x; item < n_items; item += gridDim.x) {
    const int p = item_pair[item];
    const int c = item - item_start[p];
    const long long off = offsets[p];
    const long long L = offsets[p + 1] - off;
    const long long T = stft_frames(L, N, hop);
    const long long f0 = (long long)c * chunk;
    const int nf = (int)min((long long)chunk, T - f0);
    const float* xe = est + off;
    const float* xt = tgt + off;
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
      float lsd_acc = 0.f;
      float* se = spec_e ? spec_e + spec_off[p] + f * F : nullptr;
      float* st = spec_t ? spec_t + spec_off[p] + f * F : nullptr;
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
        float ere = (float)(a.y + b.y), eim = (float)(b.x - a.x);
        float mt = sqrtf(tre * tre + tim * tim);
        float me = sqrtf(ere * ere + eim * eim);
        if (st) st[k] = mt;
        if (se) se[k] = me;
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
        float w = warp_sum(lsd_acc);
        if (lane == 0) lsd_part[fi][warp] = w;
      }
      __syncthreads();  // buf is rewritten by the next frame's load
    }

    // ---- per-item reduction -> partials[item][0..7]
    double lsd_sum = 0.0;
    if (want_lsd && tid < nf) {
      float s = 0.f;
#pragma unroll
      for (int w = 0; w < kWarps; ++w) s += lsd_part[tid][w];
      lsd_sum = (double)sqrtf(s / (float)F);  // torch.mean(dim=3) ** 0.5 in float32
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

// ---------------------------------------------------------------------------------------------
// K1, specialised: n_fft = 2048 (BASELINE config 2 and every evaluation at 44.1 kHz).
// 128 threads, 16 points per thread: radix 16 x 16 x 8 in-place DIF.
//   pass 1: samples come straight from global memory (coalesced, window folded in), the 15
//           pass-1 twiddles of a thread never change and live in registers for the CTA's lifetime;
//   pass 2: twiddles W_128^{jq} (120 values) from a conflict-free shared table;
//   pass 3: no twiddles; every thread transforms a butterfly AND its Hermitian partner
//           (k1_map.cuh), so Z[k] and Z[N-k] meet in registers and the epilogue needs no
//           further shared-memory traffic.
// Shared-memory traffic per frame: 2 exchanges (4 x 32 KB) + 30 KB of twiddles.
// ---------------------------------------------------------------------------------------------
// FIXED >= 0: the metric flags are the compile-time constant FIXED (bit 3 = the magnitude
// spectrograms are written for K2); hot configurations: 1 = LSD only, 7 = LSD + log-sispec + sispec,
// 15 = those + spectrograms.  FIXED < 0: run-time flags.
template <int FIXED>
__global__ void __launch_bounds__(kV2Threads, 3)
k_stft_metrics_2048(StftDev P, const float* __restrict__ est, const float* __restrict__ tgt,
                    const long long* __restrict__ offsets, const int* __restrict__ item_start,
                    const int* __restrict__ item_pair, int n_items, int chunk, unsigned flags,
                    double* __restrict__ partials, float* __restrict__ spec_e,
                    float* __restrict__ spec_t, const long long* __restrict__ spec_off) {
  constexpr int N = 2048, F = 1025, NW = kV2Threads / 32;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  cd* const buf = reinterpret_cast<cd*>(smem_raw);                                // N + N/8 slots
  float2* const edge_raw = reinterpret_cast<float2*>(smem_raw);  // edge frames stage N raw pairs inside buf
  float* const row_t = reinterpret_cast<float*>(smem_raw + sizeof(cd) * (N + N / 8));  // magnitude rows (store mode)
  float* const row_e = row_t + 1104;
  __shared__ __align__(16) cd tw2[15 * 8];
  __shared__ float lsd_part[kMaxChunk][NW];
  __shared__ double red[NW][kPartials];

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int hop = P.hop;
  if (FIXED >= 0) flags = (unsigned)FIXED;
  const bool want_lsd = flags & SSR_METRIC_LSD, want_log = flags & SSR_METRIC_LOG_SISPEC,
             want_lin = flags & SSR_METRIC_SISPEC;
  if (FIXED >= 0 && !(FIXED & 8)) {
    spec_e = nullptr;
    spec_t = nullptr;
  }

  // per-thread constants
  cd tw1[15];
#pragma unroll
  for (int q = 1; q < 16; ++q) tw1[q - 1] = P.tw[tid * q];
  if (tid < 120) tw2[tid] = P.tw[16 * (tid & 7) * ((tid >> 3) + 1)];  // tw2[(q-1)*8 + j] = W_128^{jq}
  int ia, ib;
  v2_thread_butterflies(tid, &ia, &ib);
  const bool special = (tid == kV2Threads - 1);
  const int ka = v2_klow(ia), kb = v2_klow(ib);
  const int j2 = tid & 7;
  // padded slots: pass 1 element q -> p1 + 144 q; pass 2 element r -> p2 + 9 r; pass 3 -> 9 i + r
  cd* const b1 = buf + pad_idx(tid);
  cd* const b2 = buf + pad_idx((tid >> 3) * 128 + j2);
  const cd* const b3a = buf + 9 * ia;
  const cd* const b3b = buf + 9 * ib;
  const cd* const t2 = tw2 + j2;
  __syncthreads();

  for (int item = blockIdx.x; item < n_items; item += gridDim.x) {
    const int p = item_pair[item];
    const int c = item - item_start[p];
    const long long off = offsets[p];
    const long long L = offsets[p + 1] - off;
    const long long T = stft_frames(L, N, hop);
    const long long f0 = (long long)c * chunk;
    const int nf = (int)min((long long)chunk, T - f0);
    const float* xe = est + off;
    const float* xt = tgt + off;
    double s_et = 0, s_tt = 0, s_ee = 0, l_et = 0, l_tt = 0, l_ee = 0;
    float* pend_t = nullptr;
    float* pend_e = nullptr;

    for (int fi = 0; fi < nf; ++fi) {
      const long long f = f0 + fi;
      const long long start = f * hop - N / 2;
      cd v[16];
      // ---- pass 1: load + window, radix-16, twiddle, store
      if (start >= 0 && start + N <= L) {
        const float* pt = xt + start + tid;
        const float* pe = xe + start + tid;
#pragma unroll
        for (int r = 0; r < 16; ++r) {
          const double w = __ldg(P.win_half + tid + 128 * r);
          v[r] = cd{w * (double)__ldg(pt + 128 * r), w * (double)__ldg(pe + 128 * r)};
        }
        if (tid < 32) {
          // next frame's new samples: [start + N, start + N + hop) of both signals, one 128 B line per lane
          const long long nxt = start + N + (long long)(tid & 15) * 32;
          if (nxt < L && (tid & 15) * 32 < hop) prefetch_l1((tid < 16 ? xt : xe) + nxt);
        }
      } else {
        // edge frame (reflect padding; < 1 % of the frames): gather through a small staging array so
        // the 64-bit reflect arithmetic stays out of the unrolled hot path
        __syncthreads();  // the staging area aliases buf: the previous frame's pass-3 loads must be done
#pragma unroll 1
        for (int n = tid; n < N; n += kV2Threads) {
          const long long idx = reflect_index(start + n, L);
          edge_raw[n] = make_float2(__ldg(xt + idx), __ldg(xe + idx));
        }
        __syncthreads();
#pragma unroll
        for (int r = 0; r < 16; ++r) {
          const double w = __ldg(P.win_half + tid + 128 * r);
          const float2 x = edge_raw[tid + 128 * r];
          v[r] = cd{w * (double)x.x, w * (double)x.y};
        }
      }
      bfly16<false>(v);
#pragma unroll
      for (int q = 1; q < 16; ++q) v[q] = cmul(v[q], tw1[q - 1]);
      // the previous frame's pass-3 loads must be done before buf is overwritten; placed here (after
      // this frame's loads and butterfly) the barrier finds every warp long past that point
      __syncthreads();
      if (pend_t) {  // coalesced copy-out of the previous frame's magnitude rows (all epilogues are done)
        for (int k = tid; k < F; k += kV2Threads) {
          pend_t[k] = row_t[k + (k >> 4)];
          if (pend_e) pend_e[k] = row_e[k + (k >> 4)];
        }
        pend_t = nullptr;
      }
#pragma unroll
      for (int q = 0; q < 16; ++q) b1[144 * q] = v[q];
      __syncthreads();
      // ---- pass 2: sub-transforms of length 128 (stride 8)
#pragma unroll
      for (int r = 0; r < 16; ++r) v[r] = b2[9 * r];
      bfly16<false>(v);
      b2[0] = v[0];
#pragma unroll
      for (int q = 1; q < 16; ++q) b2[9 * q] = cmul(v[q], t2[(q - 1) * 8]);
      __syncthreads();
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
      float lsd_acc = 0.f;
      float* st = spec_t ? spec_t + spec_off[p] + f * F : nullptr;
      float* se = spec_e ? spec_e + spec_off[p] + f * F : nullptr;
      auto emit = [&](int k, cd zk, cd zn) {
        // T = (Z[k] + conj Z[N-k]) / 2,  E = (Z[k] - conj Z[N-k]) / (2i); the 1/2 is in the window.
        // complex64 rounding as librosa stores it, then float32 arithmetic as torch runs it; the
        // special functions are the hardware approximations (MUFU sqrt / rcp / lg2, <= 2 ulp), well
        // inside the differences that already exist between numpy's hypotf / torch's log10 and any
        // other libm (SSR_EXACT_F32_EPILOGUE switches to the IEEE-rounded forms for A/B tests).
        const float tre = (float)(zk.x + zn.x), tim = (float)(zk.y - zn.y);
        const float ere = (float)(zk.y + zn.y), eim = (float)(zn.x - zk.x);
        const float tx = tre * tre + tim * tim;  // |T|^2
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
        const float w = warp_sum(lsd_acc);
        if (lane == 0) lsd_part[fi][warp] = w;
      }
      pend_t = st;  // copied out after the next barrier (next frame's pass 1, or the item epilogue)
      pend_e = se;
    }
    __syncthreads();
    if (pend_t) {
      for (int k = tid; k < F; k += kV2Threads) {
        pend_t[k] = row_t[k + (k >> 4)];
        if (pend_e) pend_e[k] = row_e[k + (k >> 4)];
      }
      pend_t = nullptr;
    }
    // ---- per-item reduction -> partials[item][0..7]
    double lsd_sum = 0.0;
    if (want_lsd && tid < nf) {
      float sacc = 0.f;
#pragma unroll
      for (int w = 0; w < NW; ++w) sacc += lsd_part[tid][w];
      lsd_sum = (double)sqrtf(sacc / (float)F);  // torch.mean(dim=3) ** 0.5 in float32
    }
    double vals[7] = {lsd_sum, s_et, s_tt, s_ee, l_et, l_tt, l_ee};
#pragma unroll
    for (int i = 0; i < 7; ++i) {
      const double r = warp_sum(vals[i]);
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
}

// ---------------------------------------------------------------------------------------------
// K1 for non-power-of-two n_fft = R * P (2229 = 3 x 743 at 48 kHz -- the reference's own default --,
// 1114, 743, 1486 ...): R Bluestein sub-transforms of length P on the 2048-point radix 16x16x8
// machinery of k_stft_metrics_2048 (forward DIF, filter multiply and inverse butterfly of the last /
// first pass in registers, inverse DIT), recombined with a radix-R butterfly on the fly in the
// epilogue.  ~6 FFT-2048 per frame instead of 2 FFT-8192 (1.6x fewer flops, 2x less shared traffic
// than the generic Bluestein kernel).  NQ = ceil(P / 128): pass-1 inputs / pass-3' outputs beyond
// NQ are structurally zero / unused and are pruned at compile time.
// ---------------------------------------------------------------------------------------------
// PIPE = 1: 2 CTAs/SM (255 registers): the Bluestein filter values of the thread's two butterflies
// live in registers and the samples + window*chirp factors of the NEXT sub-transform are fetched into
// registers one sub-transform ahead (the tables do not fit the L1 left beside 3 CTAs' shared memory,
// so every table load is an L2 round trip that has to be hidden in software).
template <int NQ, int FIXED, int PIPE>
__global__ void __launch_bounds__(kV2Threads, PIPE ? 2 : 3)
k_stft_metrics_pfa(PfaDev D, const float* __restrict__ est, const float* __restrict__ tgt,
                   const long long* __restrict__ offsets, const int* __restrict__ item_start,
                   const int* __restrict__ item_pair, int n_items, int chunk, unsigned flags,
                   double* __restrict__ partials, float* __restrict__ spec_e,
                   float* __restrict__ spec_t, const long long* __restrict__ spec_off) {
  constexpr int M = 2048, NW = kV2Threads / 32;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  cd* const buf = reinterpret_cast<cd*>(smem_raw);                       // M + M/8 slots
  // Y_r, r < R-1, live behind buf; the LAST sub-transform's Y is written over buf itself (dead by then),
  // which keeps the CTA at ~65 KB of shared memory = 3 CTAs per SM
  cd* const Yx = reinterpret_cast<cd*>(smem_raw + sizeof(cd) * (M + M / 8));
  __shared__ __align__(16) cd tw2[15 * 8];
  __shared__ __align__(16) cd wr_s[16];
  __shared__ float lsd_part[kMaxChunk][NW];
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

  cd tw1[15];
#pragma unroll
  for (int q = 1; q < 16; ++q) tw1[q - 1] = D.tw[tid * q];
  if (tid < 120) tw2[tid] = D.tw[16 * (tid & 7) * ((tid >> 3) + 1)];
  if (tid < R * R) wr_s[tid] = D.wr[tid];
  int ia, ib;
  v2_thread_butterflies(tid, &ia, &ib);
  const int j2 = tid & 7;
  cd* const b1 = buf + pad_idx(tid);
  cd* const b2 = buf + pad_idx((tid >> 3) * 128 + j2);
  cd* const b3a = buf + 9 * ia;
  cd* const b3b = buf + 9 * ib;
  const cd* const t2 = tw2 + j2;
  __syncthreads();

  cd fa[8], fb[8];  // PIPE: Bluestein filter at this thread's 16 slots
  if (PIPE) {
#pragma unroll
    for (int q = 0; q < 8; ++q) {
      fa[q] = D.bfilt[8 * ia + q];
      fb[q] = D.bfilt[8 * ib + q];
    }
  }

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

  for (int item = blockIdx.x; item < n_items; item += gridDim.x) {
    const int p = item_pair[item];
    const int c = item - item_start[p];
    const long long off = offsets[p];
    const long long L = offsets[p + 1] - off;
    const long long T = stft_frames(L, N, hop);
    const long long f0 = (long long)c * chunk;
    const int nf = (int)min((long long)chunk, T - f0);
    const float* xe = est + off;
    const float* xt = tgt + off;
    double s_et = 0, s_tt = 0, s_ee = 0, l_et = 0, l_tt = 0, l_ee = 0;

    // (PIPE) inputs of sub-transform (f, r): samples of both signals and window*chirp, NQ per thread
    float ptx[NQ], pex[NQ];
    cd pcw[NQ];
    auto fetch_inputs = [&](long long f, int r) {
      const long long start = f * hop - N / 2;
      const bool interior = (start >= 0 && start + N <= L);
#pragma unroll
      for (int q = 0; q < NQ; ++q) {
        const int n = tid + 128 * q;
        ptx[q] = 0.f;
        pex[q] = 0.f;
        pcw[q] = cd{0.0, 0.0};
        if (n < P) {
          const long long si = start + (long long)R * n + r;
          const long long idx = interior ? si : reflect_index(si, L);
          ptx[q] = __ldg(xt + idx);
          pex[q] = __ldg(xe + idx);
          pcw[q] = D.cwin[r * P + n];
        }
      }
    };
    if (PIPE) fetch_inputs(f0, 0);

    for (int fi = 0; fi < nf; ++fi) {
      const long long f = f0 + fi;
      const long long start = f * hop - N / 2;
      const bool interior = (start >= 0 && start + N <= L);
      for (int r = 0; r < R; ++r) {
        cd v[16];
        // ---- forward pass 1: a[n] = z[R n + r] * (0.5 window * chirp), zero padded to 2048
        if (PIPE) {
#pragma unroll
          for (int q = 0; q < 16; ++q) {
            v[q] = cd{0.0, 0.0};
            if (q < NQ) {
              const double tt = (double)ptx[q], ee = (double)pex[q];
              v[q] = cd{tt * pcw[q].x - ee * pcw[q].y, tt * pcw[q].y + ee * pcw[q].x};
            }
          }
          // the loads of the next sub-transform fly during all passes of this one
          if (r + 1 < R) fetch_inputs(f, r + 1);
          else if (fi + 1 < nf) fetch_inputs(f + 1, 0);
        } else {
#pragma unroll
          for (int q = 0; q < 16; ++q) {
            v[q] = cd{0.0, 0.0};
            if (q < NQ) {
              const int n = tid + 128 * q;
              if (n < P) {
                const long long si = start + (long long)R * n + r;
                const long long idx = interior ? si : reflect_index(si, L);
                const double tt = (double)__ldg(xt + idx), ee = (double)__ldg(xe + idx);
                const cd w = D.cwin[r * P + n];
                v[q] = cd{tt * w.x - ee * w.y, tt * w.y + ee * w.x};
              }
            }
          }
        }
        bfly16<false>(v);
#pragma unroll
        for (int q = 1; q < 16; ++q) v[q] = cmul(v[q], tw1[q - 1]);
        __syncthreads();  // previous sub-transform's last loads are done
#pragma unroll
        for (int q = 0; q < 16; ++q) b1[144 * q] = v[q];
        __syncthreads();
        // ---- forward pass 2
#pragma unroll
        for (int q = 0; q < 16; ++q) v[q] = b2[9 * q];
        bfly16<false>(v);
        b2[0] = v[0];
#pragma unroll
        for (int q = 1; q < 16; ++q) b2[9 * q] = cmul(v[q], t2[(q - 1) * 8]);
        __syncthreads();
        // ---- forward pass 3, Bluestein filter, inverse pass 1: all in registers
        cd* a = v;
        cd* b = v + 8;
#pragma unroll
        for (int q = 0; q < 8; ++q) {
          a[q] = b3a[q];
          b[q] = b3b[q];
        }
        bfly8<false>(a);
        bfly8<false>(b);
#pragma unroll
        for (int q = 0; q < 8; ++q) {
          a[q] = cmul(a[q], PIPE ? fa[q] : D.bfilt[8 * ia + q]);
          b[q] = cmul(b[q], PIPE ? fb[q] : D.bfilt[8 * ib + q]);
        }
        bfly8<true>(a);
        bfly8<true>(b);
#pragma unroll
        for (int q = 0; q < 8; ++q) {
          b3a[q] = a[q];
          b3b[q] = b[q];
        }
        __syncthreads();
        // ---- inverse pass 2
        v[0] = b2[0];
#pragma unroll
        for (int q = 1; q < 16; ++q) v[q] = cmul_conj(b2[9 * q], t2[(q - 1) * 8]);
        bfly16<true>(v);
#pragma unroll
        for (int q = 0; q < 16; ++q) b2[9 * q] = v[q];
        __syncthreads();
        // ---- inverse pass 3 -> conv[k], k = tid + 128 q; Y_r[k] = conv[k] * chirp[k] * W_N^{rk}
        v[0] = b1[0];
#pragma unroll
        for (int q = 1; q < 16; ++q) v[q] = cmul_conj(b1[144 * q], tw1[q - 1]);
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
      float lsd_acc = 0.f;
      float* st = spec_t ? spec_t + spec_off[p] + f * F : nullptr;
      float* se = spec_e ? spec_e + spec_off[p] + f * F : nullptr;
      for (int k = tid; k < F; k += kV2Threads) {
        const cd zk = combine(k);
        const cd zn = combine(k ? N - k : 0);
        const float tre = (float)(zk.x + zn.x), tim = (float)(zk.y - zn.y);
        const float ere = (float)(zk.y + zn.y), eim = (float)(zn.x - zk.x);
        const float tx = tre * tre + tim * tim;
        const float ey = ere * ere + eim * eim;
        const float me = __fsqrt_approx(ey);
        const float mt = (st || want_lin || want_log) ? __fsqrt_approx(tx) : 0.f;
        if (st) st[k] = mt;
        if (se) se[k] = me;
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
        const float w = warp_sum(lsd_acc);
        if (lane == 0) lsd_part[fi][warp] = w;
      }
    }
    __syncthreads();
    double lsd_sum = 0.0;
    if (want_lsd && tid < nf) {
      float sacc = 0.f;
#pragma unroll
      for (int w = 0; w < NW; ++w) sacc += lsd_part[tid][w];
      lsd_sum = (double)sqrtf(sacc / (float)F);
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
}

// ---------------------------------------------------------------------------------------------
// K2: SSIM of two (T, F) float32 magnitude images, valid 7x7 windows only (skimage crops the
// 3-pixel border, so the reflect boundary mode of uniform_filter never reaches the mean).
// One CTA = one tile of kSsimTR x kSsimTC window positions, 128 threads, TWO adjacent columns per
// thread.  Rows stream through a double-buffered shared row buffer; per row a thread forms the
// horizontal 7-sums of (x, y, xx, yy, xy) for its two columns (sliding: the second column reuses the
// first column's inner sum) and updates RUNNING vertical 7-sums: V += h_new - h_oldest, with the last
// seven h kept in a register ring (unrolled-by-7 loop).  The running sums restart in every tile, so
// the result does not depend on how the batch was partitioned.
// ---------------------------------------------------------------------------------------------
constexpr int kSsimThreads = 128;

__global__ void __launch_bounds__(kSsimThreads)
k_ssim(const float* __restrict__ spec_e, const float* __restrict__ spec_t,
       const long long* __restrict__ spec_off, const long long* __restrict__ offsets, int pair0,
       int n_fft, int hop, int F, int tiles_x, int tiles_per_pair, double* __restrict__ ssim_part) {
  const int p = pair0 + blockIdx.y;
  const int tile = blockIdx.x;
  const int ty = tile / tiles_x, tx = tile % tiles_x;
  const long long T = stft_frames(offsets[p + 1] - offsets[p], n_fft, hop);
  const int rows_out = (int)T - 6, cols_out = F - 6;
  const int r0 = ty * kSsimTR;
  double* out = ssim_part + (size_t)p * tiles_per_pair + tile;
  if (r0 >= rows_out || cols_out <= 0) {
    if (threadIdx.x == 0) *out = 0.0;
    return;
  }
  const int r_end = min(r0 + kSsimTR, rows_out) + 6;  // input rows [r0, r_end)
  const int c0 = tx * kSsimTC;
  const int t = threadIdx.x;
  const int c = 2 * t;  // first of this thread's two columns inside the tile
  const bool ok0 = (c0 + c) < cols_out, ok1 = (c0 + c + 1) < cols_out;
  const float* E = spec_e + spec_off[p];
  const float* G = spec_t + spec_off[p];
  constexpr int RB = kSsimTC + 8, STAGES = 4;
  __shared__ __align__(16) float rowbuf[STAGES][2][RB];
  __shared__ double red[kSsimThreads / 32];

  float ring[7][10];
#pragma unroll
  for (int s = 0; s < 7; ++s)
#pragma unroll
    for (int q = 0; q < 10; ++q) ring[s][q] = 0.f;
  float V[10];
#pragma unroll
  for (int q = 0; q < 10; ++q) V[q] = 0.f;
  float acc = 0.f;
  const float inv49 = 1.0f / 49.0f, cov_norm = 49.0f / 48.0f;
  const float C1 = 0.0004f, C2 = 0.0036f;  // (0.01*2)^2, (0.03*2)^2

  // rows stream global -> shared with cp.async (LDGSTS), STAGES-1 rows in flight; columns beyond the
  // image are zero-filled by the copy itself (src-size 0)
  auto issue_row = [&](int r) {
    if (r < r_end) {
      const int stg = (r - r0) % STAGES;
      const float* er = E + (long long)r * F + c0;
      const float* gr = G + (long long)r * F + c0;
#pragma unroll
      for (int u = 0; u < 3; ++u) {
        const int col = t + u * kSsimThreads;
        if (u < 2 || t < 8) {
          const bool in = (c0 + col) < F;
          cp_async4(&rowbuf[stg][0][col], in ? er + col : er, in ? 4 : 0);
          cp_async4(&rowbuf[stg][1][col], in ? gr + col : gr, in ? 4 : 0);
        }
      }
    }
    asm volatile("cp.async.commit_group;" ::: "memory");
  };
#pragma unroll
  for (int k = 0; k < STAGES - 1; ++k) issue_row(r0 + k);

  for (int rb = r0; rb < r_end; rb += 7) {
#pragma unroll
    for (int s = 0; s < 7; ++s) {
      const int r = rb + s;
      if (r < r_end) {  // uniform across the CTA
        const int par = (r - r0) % STAGES;
        asm volatile("cp.async.wait_group %0;" ::"n"(STAGES - 2) : "memory");
        __syncthreads();               // row r has landed for everyone; row r-1 is fully consumed
        issue_row(r + STAGES - 1);     // refills the stage row r-1 occupied
        float x[8], y[8];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const float2 xv = *reinterpret_cast<const float2*>(&rowbuf[par][0][c + 2 * j]);
          const float2 yv = *reinterpret_cast<const float2*>(&rowbuf[par][1][c + 2 * j]);
          x[2 * j] = xv.x;
          x[2 * j + 1] = xv.y;
          y[2 * j] = yv.x;
          y[2 * j + 1] = yv.y;
        }
        // inner sums over columns c+1 .. c+6, then the two outputs add their own end column
        float ix = 0.f, iy = 0.f, ixx = 0.f, iyy = 0.f, ixy = 0.f;
#pragma unroll
        for (int j = 1; j < 7; ++j) {
          ix += x[j];
          iy += y[j];
          ixx += x[j] * x[j];
          iyy += y[j] * y[j];
          ixy += x[j] * y[j];
        }
        float h[10];
        h[0] = ix + x[0];
        h[1] = iy + y[0];
        h[2] = ixx + x[0] * x[0];
        h[3] = iyy + y[0] * y[0];
        h[4] = ixy + x[0] * y[0];
        h[5] = ix + x[7];
        h[6] = iy + y[7];
        h[7] = ixx + x[7] * x[7];
        h[8] = iyy + y[7] * y[7];
        h[9] = ixy + x[7] * y[7];
#pragma unroll
        for (int q = 0; q < 10; ++q) {
          V[q] += h[q] - ring[s][q];
          ring[s][q] = h[q];
        }
        if (r - r0 >= 6) {
#pragma unroll
          for (int o = 0; o < 2; ++o) {
            const float ux = V[5 * o] * inv49, uy = V[5 * o + 1] * inv49;
            const float uxx = V[5 * o + 2] * inv49, uyy = V[5 * o + 3] * inv49, uxy = V[5 * o + 4] * inv49;
            const float vx = cov_norm * (uxx - ux * ux);
            const float vy = cov_norm * (uyy - uy * uy);
            const float vxy = cov_norm * (uxy - ux * uy);
            const float A1 = 2.f * ux * uy + C1, A2 = 2.f * vxy + C2;
            const float B1 = ux * ux + uy * uy + C1, B2 = vx + vy + C2;
            const float S = __fdividef(A1 * A2, B1 * B2);
            if (o == 0 ? ok0 : ok1) acc += S;
          }
        }
      }
    }
  }
  asm volatile("cp.async.wait_group 0;" ::: "memory");
  double r = warp_sum((double)acc);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = r;
  __syncthreads();
  if (threadIdx.x == 0) {
    double sum = 0.0;
    for (int w = 0; w < kSsimThreads / 32; ++w) sum += red[w];
    *out = sum;
  }
}

// ---------------------------------------------------------------------------------------------
// finalize: fixed-order sum of the per-item partials -> the four metrics of each pair (float64).
// ---------------------------------------------------------------------------------------------
__device__ inline double sispec_from_sums(double s_et, double s_tt, double s_ee) {
  const double EPS = 1e-12;
  double alpha = s_et / (s_tt + EPS);            // energy_unify: target' = alpha * target
  double tt = alpha * alpha * s_tt;              // ||target'||^2
  double nn = s_ee - 2.0 * alpha * s_et + tt;    // ||est - target'||^2
  if (nn < 0.0) nn = 0.0;
  return 10.0 * log10(tt / (nn + EPS) + EPS);
}

__global__ void __launch_bounds__(128)
k_finalize(const long long* __restrict__ offsets, int n, int n_fft, int hop, int F,
           const int* __restrict__ item_start, const double* __restrict__ partials,
           const double* __restrict__ ssim_part, int tiles_per_pair, unsigned flags,
           double* __restrict__ out) {
  // one warp per pair; lanes stride over the items / tiles, then a fixed-shape shuffle tree:
  // the summation order depends only on the pair's own item / tile count
  const int p = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (p >= n) return;
  const double nan = __longlong_as_double(0x7ff8000000000000LL);
  const long long T = stft_frames(offsets[p + 1] - offsets[p], n_fft, hop);
  double v[7] = {0, 0, 0, 0, 0, 0, 0};
  for (int it = item_start[p] + lane; it < item_start[p + 1]; it += 32)
#pragma unroll
    for (int i = 0; i < 7; ++i) v[i] += partials[(size_t)it * kPartials + i];
  double s = 0.0;
  if (flags & SSR_METRIC_SSIM)
    for (int t = lane; t < tiles_per_pair; t += 32) s += ssim_part[(size_t)p * tiles_per_pair + t];
#pragma unroll
  for (int i = 0; i < 7; ++i) v[i] = warp_sum(v[i]);
  s = warp_sum(s);
  if (lane != 0) return;
  out[p * 4 + 0] = (flags & SSR_METRIC_LSD) ? v[0] / (double)T : nan;
  out[p * 4 + 1] = (flags & SSR_METRIC_LOG_SISPEC) ? sispec_from_sums(v[4], v[5], v[6]) : nan;
  out[p * 4 + 2] = (flags & SSR_METRIC_SISPEC) ? sispec_from_sums(v[1], v[2], v[3]) : nan;
  const double cnt = (double)(T - 6) * (double)(F - 6);
  out[p * 4 + 3] = ((flags & SSR_METRIC_SSIM) && T > 6 && F > 6) ? s / cnt : nan;
}

// ---------------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------------
// SSR_FORCE_GENERIC_K1=1 routes n_fft 2048 through the generic radix-8 kernel (A/B tests only)
static bool force_generic_k1() {
  static int v = -1;
  if (v < 0) {
    const char* e = getenv("SSR_FORCE_GENERIC_K1");
    v = (e && e[0] == '1') ? 1 : 0;
  }
  return v == 1;
}

// SSR_PFA_PIPE=0/1: PFA kernel without / with software-pipelined inputs (A/B tests)
static int pfa_pipe() {
  static int v = -1;
  if (v < 0) {
    const char* e = getenv("SSR_PFA_PIPE");
    v = e ? atoi(e) : 1;
  }
  return v ? 1 : 0;
}

struct WsLayout {
  size_t item_start, item_pair, spec_off, partials, ssim_part, spec_e, spec_t, total;
  int chunk, n_items, tiles_x, tiles_per_pair;
  long long total_frames;
};

static int plan_layout(const ssr_stft_plan* plan, const int64_t* offs, int n, unsigned flags,
                       bool need_spec_t_only, WsLayout* w) {
  long long total_frames = 0, max_T = 0;
  for (int p = 0; p < n; ++p) {
    long long L = offs[p + 1] - offs[p];
    if (L < 1) return fail(SSR_ERR_INVALID, "empty utterance in batch");
    long long T = stft_frames(L, plan->n_fft, plan->hop);
    total_frames += T;
    if (T > max_T) max_T = T;
  }
  // aim for ~8 work items per resident CTA slot (148 SMs x 2), bounded to [4, kMaxChunk] frames
  long long want = 148LL * 3 * 8;
  long long chunk = (total_frames + want - 1) / want;
  if (chunk < 4) chunk = 4;
  if (chunk > kMaxChunk) chunk = kMaxChunk;
  long long n_items = 0;
  for (int p = 0; p < n; ++p) {
    long long T = stft_frames(offs[p + 1] - offs[p], plan->n_fft, plan->hop);
    n_items += (T + chunk - 1) / chunk;
  }
  if (n_items > 0x7fffffffLL) return fail(SSR_ERR_INVALID, "batch too large");
  w->chunk = (int)chunk;
  w->n_items = (int)n_items;
  w->total_frames = total_frames;
  w->tiles_x = (plan->F - 6 + kSsimTC - 1) / kSsimTC;
  if (w->tiles_x < 1) w->tiles_x = 1;
  long long rows = max_T - 6;
  int tiles_y = rows > 0 ? (int)((rows + kSsimTR - 1) / kSsimTR) : 1;
  w->tiles_per_pair = w->tiles_x * tiles_y;
  size_t o = 0;
  w->item_start = o;
  o = align_up(o + sizeof(int) * (size_t)(n + 1), 256);
  w->item_pair = o;
  o = align_up(o + sizeof(int) * (size_t)n_items, 256);
  w->spec_off = o;
  o = align_up(o + sizeof(long long) * (size_t)(n + 1), 256);
  w->partials = o;
  o = align_up(o + sizeof(double) * kPartials * (size_t)n_items, 256);
  w->ssim_part = o;
  if (flags & SSR_METRIC_SSIM) o = align_up(o + sizeof(double) * (size_t)n * w->tiles_per_pair, 256);
  w->spec_e = o;
  if (flags & SSR_METRIC_SSIM) o = align_up(o + sizeof(float) * (size_t)total_frames * plan->F, 256);
  w->spec_t = o;
  if ((flags & SSR_METRIC_SSIM) && !need_spec_t_only)
    o = align_up(o + sizeof(float) * (size_t)total_frames * plan->F, 256);
  w->total = o;
  return SSR_OK;
}

template <int LOGM, bool BLUE>
static int launch_k1(const ssr_stft_plan* plan, int grid, size_t smem, cudaStream_t st,
                     const float* est, const float* tgt, const long long* offs_dev,
                     const int* item_start, const int* item_pair, int n_items, int chunk,
                     unsigned flags, double* partials, float* spec_e, float* spec_t,
                     const long long* spec_off) {
  auto kern = k_stft_metrics<LOGM, BLUE>;
  SSR_CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  TimingState& tm = timing();
  std::pair<cudaEvent_t, cudaEvent_t> ev{nullptr, nullptr};
  if (tm.on) {
    if (!tm.pool.empty()) {
      ev = tm.pool.back();
      tm.pool.pop_back();
    } else {
      SSR_CUDA_TRY(cudaEventCreate(&ev.first));
      SSR_CUDA_TRY(cudaEventCreate(&ev.second));
    }
    SSR_CUDA_TRY(cudaEventRecord(ev.first, st));
  }
  kern<<<grid, kThreads, smem, st>>>(plan->dev, est, tgt, offs_dev, item_start, item_pair, n_items,
                                     chunk, flags, partials, spec_e, spec_t, spec_off);
  SSR_LAUNCH_CHECK("k_stft_metrics");
  if (tm.on) {
    SSR_CUDA_TRY(cudaEventRecord(ev.second, st));
    tm.pending.push_back(ev);
  }
  return SSR_OK;
}

template <bool BLUE>
static int dispatch_k1(const ssr_stft_plan* plan, int grid, size_t smem, cudaStream_t st,
                       const float* est, const float* tgt, const long long* offs_dev,
                       const int* item_start, const int* item_pair, int n_items, int chunk,
                       unsigned flags, double* partials, float* spec_e, float* spec_t,
                       const long long* spec_off) {
#define SSR_CASE(LM)                                                                            \
  case LM:                                                                                      \
    return launch_k1<LM, BLUE>(plan, grid, smem, st, est, tgt, offs_dev, item_start, item_pair, \
                               n_items, chunk, flags, partials, spec_e, spec_t, spec_off);
  switch (plan->logM) {
    SSR_CASE(8)
    SSR_CASE(9)
    SSR_CASE(10)
    SSR_CASE(11)
    SSR_CASE(12)
    SSR_CASE(13)
    default:
      return fail(SSR_ERR_INVALID, "unsupported FFT size");
  }
#undef SSR_CASE
}

static int run_k1(const ssr_stft_plan* plan, const WsLayout& w, cudaStream_t st, const float* est,
                  const float* tgt, const long long* offs_dev, int n, unsigned flags,
                  unsigned char* ws, float* spec_e, float* spec_t) {
  int* item_start = reinterpret_cast<int*>(ws + w.item_start);
  int* item_pair = reinterpret_cast<int*>(ws + w.item_pair);
  long long* spec_off = reinterpret_cast<long long*>(ws + w.spec_off);
  double* partials = reinterpret_cast<double*>(ws + w.partials);
  k_setup<<<1, 1024, 0, st>>>(offs_dev, n, plan->n_fft, plan->hop, w.chunk, plan->F, item_start,
                              item_pair, spec_off);
  SSR_LAUNCH_CHECK("k_setup");
  size_t smem = sizeof(cd) * (size_t)padded_size(plan->M);
  int dev = 0, sms = 148;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  int per_sm = (int)((200 * 1024) / (smem + 4096));
  if (per_sm < 1) per_sm = 1;
  if (per_sm > 2) per_sm = 2;
  int grid = sms * per_sm;
  if (grid > w.n_items) grid = w.n_items;
  if (plan->pfa && !force_generic_k1()) {
    const size_t smem_p = sizeof(cd) * (2048 + 256) + sizeof(cd) * (size_t)(plan->n_fft - plan->pdev.P);
    const int pipe = pfa_pipe();
    int gp = sms * (pipe ? 2 : 3);
    if (gp > w.n_items) gp = w.n_items;
    const bool store = spec_e || spec_t;
    const bool lsd_only = !store && (flags & 7u) == 1u;
    const int nq = (plan->pdev.P + 127) / 128;
    TimingState& tm = timing();
    std::pair<cudaEvent_t, cudaEvent_t> ev{nullptr, nullptr};
    if (tm.on) {
      if (!tm.pool.empty()) {
        ev = tm.pool.back();
        tm.pool.pop_back();
      } else {
        SSR_CUDA_TRY(cudaEventCreate(&ev.first));
        SSR_CUDA_TRY(cudaEventCreate(&ev.second));
      }
      SSR_CUDA_TRY(cudaEventRecord(ev.first, st));
    }
#define SSR_PFA_LAUNCH(NQ_, FX)                                                                      \
  do {                                                                                               \
    if (pipe) SSR_PFA_LAUNCH_(NQ_, FX, 1);                                                           \
    else SSR_PFA_LAUNCH_(NQ_, FX, 0);                                                                \
  } while (0)
#define SSR_PFA_LAUNCH_(NQ_, FX, PP)                                                                 \
  do {                                                                                               \
    auto kern = k_stft_metrics_pfa<NQ_, FX, PP>;                                                     \
    SSR_CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_p)); \
    kern<<<gp, kV2Threads, smem_p, st>>>(plan->pdev, est, tgt, offs_dev, item_start, item_pair, w.n_items, \
                                         w.chunk, flags, partials, spec_e, spec_t, spec_off);        \
  } while (0)
    if (nq <= 6) {
      if (lsd_only) SSR_PFA_LAUNCH(6, 1);
      else SSR_PFA_LAUNCH(6, -1);
    } else {
      if (lsd_only) SSR_PFA_LAUNCH(8, 1);
      else SSR_PFA_LAUNCH(8, -1);
    }
#undef SSR_PFA_LAUNCH
#undef SSR_PFA_LAUNCH_
    SSR_LAUNCH_CHECK("k_stft_metrics_pfa");
    if (tm.on) {
      SSR_CUDA_TRY(cudaEventRecord(ev.second, st));
      tm.pending.push_back(ev);
    }
    return SSR_OK;
  }
  if (!plan->bluestein && plan->logM == 11 && !force_generic_k1()) {
    int g2 = sms * 3;
    if (g2 > w.n_items) g2 = w.n_items;
    TimingState& tm = timing();
    std::pair<cudaEvent_t, cudaEvent_t> ev{nullptr, nullptr};
    if (tm.on) {
      if (!tm.pool.empty()) {
        ev = tm.pool.back();
        tm.pool.pop_back();
      } else {
        SSR_CUDA_TRY(cudaEventCreate(&ev.first));
        SSR_CUDA_TRY(cudaEventCreate(&ev.second));
      }
      SSR_CUDA_TRY(cudaEventRecord(ev.first, st));
    }
    const unsigned m3 = flags & 7u;
    const size_t smem2 = sizeof(cd) * (2048 + 256) + sizeof(float) * 2 * 1104;
    const bool store = spec_e || spec_t;
    const int fixed = (!store && m3 == 1u) ? 1 : ((!store && m3 == 7u) ? 7 : ((spec_e && spec_t && m3 == 7u) ? 15 : -1));
    if (g2 > sms * 3) g2 = sms * 3;
#define SSR_V2_LAUNCH(FX)                                                                           \
  do {                                                                                              \
    auto kern = k_stft_metrics_2048<FX>;                                                            \
    SSR_CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem2)); \
    kern<<<g2, kV2Threads, smem2, st>>>(plan->dev, est, tgt, offs_dev, item_start, item_pair, w.n_items, \
                                        w.chunk, flags, partials, spec_e, spec_t, spec_off);        \
  } while (0)
    if (fixed == 1) SSR_V2_LAUNCH(1);
    else if (fixed == 7) SSR_V2_LAUNCH(7);
    else if (fixed == 15) SSR_V2_LAUNCH(15);
    else SSR_V2_LAUNCH(-1);
#undef SSR_V2_LAUNCH
    SSR_LAUNCH_CHECK("k_stft_metrics_2048");
    if (tm.on) {
      SSR_CUDA_TRY(cudaEventRecord(ev.second, st));
      tm.pending.push_back(ev);
    }
    return SSR_OK;
  }
  if (plan->bluestein)
    return dispatch_k1<true>(plan, grid, smem, st, est, tgt, offs_dev, item_start, item_pair,
                             w.n_items, w.chunk, flags, partials, spec_e, spec_t, spec_off);
  return dispatch_k1<false>(plan, grid, smem, st, est, tgt, offs_dev, item_start, item_pair,
                            w.n_items, w.chunk, flags, partials, spec_e, spec_t, spec_off);
}

}  // namespace ssr

using namespace ssr;

extern "C" {

int ssr_version(void) { return 100; }
const char* ssr_last_error(void) { return last_error_ref().c_str(); }
uint64_t ssr_launch_count(void) { return launch_counter().load(); }

int ssr_timing_enable(int on) {
  timing().on = on != 0;
  return SSR_OK;
}

int ssr_timing_collect(double* total_ms, int* n_launches) {
  TimingState& tm = timing();
  double tot = 0.0;
  int n = 0;
  for (auto& ev : tm.pending) {
    SSR_CUDA_TRY(cudaEventSynchronize(ev.second));
    float ms = 0.f;
    SSR_CUDA_TRY(cudaEventElapsedTime(&ms, ev.first, ev.second));
    tot += ms;
    ++n;
    tm.pool.push_back(ev);
  }
  tm.pending.clear();
  if (total_ms) *total_ms = tot;
  if (n_launches) *n_launches = n;
  return SSR_OK;
}

int ssr_stft_plan_create(ssr_stft_plan** out, int n_fft, int hop, const double* window_host) {
  if (!out) return fail(SSR_ERR_INVALID, "plan pointer is NULL");
  *out = nullptr;
  if (n_fft < 65 || n_fft > 8192) return fail(SSR_ERR_INVALID, "n_fft must be in [65, 8192]");
  if (hop < 1) return fail(SSR_ERR_INVALID, "hop must be >= 1");
  StftTables tb;
  if (!build_stft_tables(n_fft, window_host, &tb))
    return fail(SSR_ERR_INVALID, "n_fft too large for the Bluestein path (max 4096)");
  const int M = tb.M, logM = tb.logM;
  const bool blue = tb.bluestein;
  size_t o = 0;
  size_t o_tw = o;
  o = align_up(o + sizeof(cd) * (size_t)M, 256);
  size_t o_win = o;
  o = align_up(o + sizeof(double) * (size_t)n_fft, 256);
  size_t o_pos = o;
  o = align_up(o + sizeof(uint16_t) * (size_t)M, 256);
  size_t o_cw = o;
  o = align_up(o + sizeof(cd) * (size_t)n_fft, 256);
  size_t o_bf = o;
  o = align_up(o + sizeof(cd) * (size_t)M, 256);
  size_t o_cp = o;
  o = align_up(o + sizeof(cd) * (size_t)n_fft, 256);
  std::vector<unsigned char> host(o, 0);
  memcpy(host.data() + o_tw, tb.tw.data(), sizeof(cd) * (size_t)M);
  memcpy(host.data() + o_win, tb.win_half.data(), sizeof(double) * (size_t)n_fft);
  memcpy(host.data() + o_pos, tb.ppos.data(), sizeof(uint16_t) * (size_t)M);
  memcpy(host.data() + o_cw, tb.cw.data(), sizeof(cd) * (size_t)n_fft);
  memcpy(host.data() + o_bf, tb.bfilt.data(), sizeof(cd) * (size_t)M);
  memcpy(host.data() + o_cp, tb.cpost.data(), sizeof(cd) * (size_t)n_fft);
  ssr_stft_plan* p = new ssr_stft_plan();
  p->n_fft = n_fft;
  p->hop = hop;
  p->F = n_fft / 2 + 1;
  p->M = M;
  p->logM = logM;
  p->bluestein = blue ? 1 : 0;
  p->blob = nullptr;
  cudaError_t e = cudaGetDevice(&p->device);
  if (e == cudaSuccess) e = cudaMalloc(&p->blob, o);
  if (e == cudaSuccess) e = cudaMemcpy(p->blob, host.data(), o, cudaMemcpyHostToDevice);
  if (e != cudaSuccess) {
    if (p->blob) cudaFree(p->blob);
    delete p;
    return fail(SSR_ERR_CUDA, std::string("plan upload: ") + cudaGetErrorString(e));
  }
  p->pfa = 0;
  p->blob_pfa = nullptr;
  PfaTables pt;
  if (blue && build_pfa_tables(n_fft, window_host, &pt)) {
    const int R = pt.R;
    size_t q = 0;
    size_t q_tw = q;
    q = align_up(q + sizeof(cd) * 2048, 256);
    size_t q_cw = q;
    q = align_up(q + sizeof(cd) * (size_t)n_fft, 256);
    size_t q_po = q;
    q = align_up(q + sizeof(cd) * (size_t)n_fft, 256);
    size_t q_bf = q;
    q = align_up(q + sizeof(cd) * 2048, 256);
    size_t q_wr = q;
    q = align_up(q + sizeof(cd) * (size_t)(R * R), 256);
    std::vector<unsigned char> hp(q, 0);
    memcpy(hp.data() + q_tw, pt.tw.data(), sizeof(cd) * 2048);
    memcpy(hp.data() + q_cw, pt.cwin.data(), sizeof(cd) * (size_t)n_fft);
    memcpy(hp.data() + q_po, pt.post.data(), sizeof(cd) * (size_t)n_fft);
    memcpy(hp.data() + q_bf, pt.bfilt.data(), sizeof(cd) * 2048);
    memcpy(hp.data() + q_wr, pt.wr.data(), sizeof(cd) * (size_t)(R * R));
    cudaError_t e2 = cudaMalloc(&p->blob_pfa, q);
    if (e2 == cudaSuccess) e2 = cudaMemcpy(p->blob_pfa, hp.data(), q, cudaMemcpyHostToDevice);
    if (e2 != cudaSuccess) {
      if (p->blob_pfa) cudaFree(p->blob_pfa);
      cudaFree(p->blob);
      delete p;
      return fail(SSR_ERR_CUDA, std::string("plan upload (pfa): ") + cudaGetErrorString(e2));
    }
    unsigned char* dp = static_cast<unsigned char*>(p->blob_pfa);
    p->pfa = 1;
    p->pdev.n_fft = n_fft;
    p->pdev.hop = hop;
    p->pdev.F = n_fft / 2 + 1;
    p->pdev.R = R;
    p->pdev.P = pt.P;
    p->pdev.tw = reinterpret_cast<const cd*>(dp + q_tw);
    p->pdev.cwin = reinterpret_cast<const cd*>(dp + q_cw);
    p->pdev.post = reinterpret_cast<const cd*>(dp + q_po);
    p->pdev.bfilt = reinterpret_cast<const cd*>(dp + q_bf);
    p->pdev.wr = reinterpret_cast<const cd*>(dp + q_wr);
  }
  unsigned char* d = static_cast<unsigned char*>(p->blob);
  p->dev.n_fft = n_fft;
  p->dev.hop = hop;
  p->dev.F = p->F;
  p->dev.M = M;
  p->dev.tw = reinterpret_cast<const cd*>(d + o_tw);
  p->dev.win_half = reinterpret_cast<const double*>(d + o_win);
  p->dev.ppos = reinterpret_cast<const uint16_t*>(d + o_pos);
  p->dev.cw = reinterpret_cast<const cd*>(d + o_cw);
  p->dev.bfilt = reinterpret_cast<const cd*>(d + o_bf);
  p->dev.cpost = reinterpret_cast<const cd*>(d + o_cp);
  *out = p;
  return SSR_OK;
}

int ssr_stft_plan_destroy(ssr_stft_plan* plan) {
  if (!plan) return SSR_OK;
  if (plan->blob) cudaFree(plan->blob);
  if (plan->blob_pfa) cudaFree(plan->blob_pfa);
  delete plan;
  return SSR_OK;
}

int64_t ssr_stft_num_frames(const ssr_stft_plan* plan, int64_t length) {
  return plan ? (int64_t)stft_frames(length, plan->n_fft, plan->hop) : -1;
}

size_t ssr_stft_metrics_workspace_bytes(const ssr_stft_plan* plan, const int64_t* offsets_host,
                                        int n_pairs, unsigned flags) {
  if (!plan || !offsets_host || n_pairs < 1) return 0;
  WsLayout w;
  if (plan_layout(plan, offsets_host, n_pairs, flags, false, &w) != SSR_OK) return 0;
  return w.total;
}

int ssr_stft_metrics_batched(const ssr_stft_plan* plan, const float* est_dev, const float* tgt_dev,
                             const int64_t* offsets_host, const int64_t* offsets_dev, int n_pairs,
                             unsigned flags, double* out_dev, void* workspace_dev,
                             size_t workspace_bytes, void* stream) {
  if (!plan || !est_dev || !tgt_dev || !offsets_host || !offsets_dev || !out_dev || n_pairs < 1)
    return fail(SSR_ERR_INVALID, "ssr_stft_metrics_batched: bad argument");
  if (flags & ~SSR_METRIC_ALL) return fail(SSR_ERR_INVALID, "unknown metric flag");
  WsLayout w;
  int rc = plan_layout(plan, offsets_host, n_pairs, flags, false, &w);
  if (rc != SSR_OK) return rc;
  if (!workspace_dev || workspace_bytes < w.total)
    return fail(SSR_ERR_WORKSPACE, "workspace too small");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  unsigned char* ws = static_cast<unsigned char*>(workspace_dev);
  const long long* offs = reinterpret_cast<const long long*>(offsets_dev);
  const bool do_ssim = flags & SSR_METRIC_SSIM;
  float* spec_e = do_ssim ? reinterpret_cast<float*>(ws + w.spec_e) : nullptr;
  float* spec_t = do_ssim ? reinterpret_cast<float*>(ws + w.spec_t) : nullptr;
  rc = run_k1(plan, w, st, est_dev, tgt_dev, offs, n_pairs, flags, ws, spec_e, spec_t);
  if (rc != SSR_OK) return rc;
  double* ssim_part = reinterpret_cast<double*>(ws + w.ssim_part);
  if (do_ssim) {
    for (int p0 = 0; p0 < n_pairs; p0 += 32768) {
      int np = n_pairs - p0 < 32768 ? n_pairs - p0 : 32768;
      dim3 grid(w.tiles_per_pair, np);
      k_ssim<<<grid, kSsimThreads, 0, st>>>(spec_e, spec_t, reinterpret_cast<long long*>(ws + w.spec_off),
                                       offs, p0, plan->n_fft, plan->hop, plan->F, w.tiles_x,
                                       w.tiles_per_pair, ssim_part);
      SSR_LAUNCH_CHECK("k_ssim");
    }
  }
  k_finalize<<<(n_pairs + 3) / 4, 128, 0, st>>>(
      offs, n_pairs, plan->n_fft, plan->hop, plan->F, reinterpret_cast<int*>(ws + w.item_start),
      reinterpret_cast<double*>(ws + w.partials), ssim_part, w.tiles_per_pair, flags, out_dev);
  SSR_LAUNCH_CHECK("k_finalize");
  return SSR_OK;
}

int ssr_stft_magnitude_batched(const ssr_stft_plan* plan, const float* x_dev,
                               const int64_t* offsets_host, const int64_t* offsets_dev, int n,
                               float* spec_dev, void* workspace_dev, size_t workspace_bytes,
                               void* stream) {
  if (!plan || !x_dev || !offsets_host || !offsets_dev || !spec_dev || n < 1)
    return fail(SSR_ERR_INVALID, "ssr_stft_magnitude_batched: bad argument");
  WsLayout w;
  int rc = plan_layout(plan, offsets_host, n, 0, false, &w);
  if (rc != SSR_OK) return rc;
  if (!workspace_dev || workspace_bytes < w.total)
    return fail(SSR_ERR_WORKSPACE, "workspace too small");
  return run_k1(plan, w, static_cast<cudaStream_t>(stream), x_dev, x_dev,
                reinterpret_cast<const long long*>(offsets_dev), n, 0,
                static_cast<unsigned char*>(workspace_dev), nullptr, spec_dev);
}

}  // extern "C"
