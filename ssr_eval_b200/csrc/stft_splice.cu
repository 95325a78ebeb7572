// stft_splice.cu -- K6: STFT splice + ISTFT of BasicTestee.postprocessing.
//
// Replaces (paths relative to the reference repo) ssr_eval/eval.py:33-41:
//   stft_gt = librosa.stft(x); stft_out = librosa.stft(out)            (n_fft 2048, hop 512, float64 -> c64)
//   stft_out[:cutoff] = stft_gt[:cutoff]; out = librosa.istft(stft_out, length=len(out))
// One CTA owns `chunk_hops` hops of output.  Per frame pair (f, f+1):
//   forward  z = w*x_f + i*w*out_f  (ONE complex float64 FFT gives both spectra), split, complex64
//            rounding, bin-wise choice S[k] = k < cutoff ? X[k] : O[k]          -- twice (f and f+1);
//   inverse  Y = S_f + i*S_{f+1} (Hermitian extended) -> ONE complex float64 inverse FFT whose real /
//            imaginary parts are irfft(S_f), irfft(S_{f+1}); x float64 window / n_fft;
//   overlap-add into a float32 accumulator exactly as librosa does (y32 = float32(y32 + float64 term)),
//   divide by the float32 window-sum-square where it exceeds tiny, trim n_fft/2, length samples.
// Same radix 16x16x8 / register-paired machinery as k_stft_metrics_2048 (k1_map.cuh).
#include <math.h>

#include <vector>

#include "common.cuh"
#include "fft_core.cuh"
#include "k1_map.cuh"
#include "stft_tables.hpp"

struct ssr_splice_plan {
  int n_fft, hop, device;
  void* blob;
  const ssr::cd* tw;
  const double* win_half;    // 0.5 * hann[n]
  const double* win_over_n;  // hann[n] / n_fft
  const double* win_sq;      // hann[n]^2
  const float* ws_tab;       // [hop]: window_sumsquare of an interior sample m, index m % hop
};

namespace ssr {

struct SpDev {
  int hop;
  const cd* tw;
  const double* win_half;
  const double* win_over_n;
  const double* win_sq;
  const float* ws_tab;
};

__device__ __forceinline__ long long sp_reflect(long long i, long long L) {
  if (i >= 0 && i < L) return i;
  if (L == 1) return 0;
  long long period = 2 * (L - 1);
  i %= period;
  if (i < 0) i += period;
  return i < L ? i : period - i;
}

__global__ void __launch_bounds__(kV2Threads, 2)
k_stft_splice_istft_2048(SpDev P, const float* __restrict__ xin, const float* __restrict__ xout,
                         const long long* __restrict__ offsets, const int* __restrict__ cut_bins,
                         float* __restrict__ y, int u0, int chunk_hops) {
  constexpr int N = 2048;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  cd* const buf = reinterpret_cast<cd*>(smem_raw);
  float* const acc = reinterpret_cast<float*>(smem_raw + sizeof(cd) * (N + N / 8));
  __shared__ __align__(16) cd tw2[15 * 8];

  const int tid = threadIdx.x;
  const int u = u0 + blockIdx.y;
  const long long off = offsets[u];
  const long long L = offsets[u + 1] - off;
  const int hop = P.hop;
  const long long n0 = (long long)blockIdx.x * chunk_hops * hop;
  if (n0 >= L) return;
  const long long n1 = min(L, n0 + (long long)chunk_hops * hop);
  const long long m0 = n0 + N / 2, m1 = n1 + N / 2;
  const int span = (int)(m1 - m0);
  const long long T = 1 + L / hop;  // librosa.stft frames (n_fft even); istft uses all of them
  const long long f_lo = (m0 - N >= 0) ? (m0 - N) / hop + 1 : 0;
  const long long f_hi = min(T - 1, (m1 - 1) / hop);
  const int cut = cut_bins[u];
  const float* xa = xin + off;
  const float* xb = xout + off;

  cd tw1[15];
#pragma unroll
  for (int q = 1; q < 16; ++q) tw1[q - 1] = P.tw[tid * q];
  if (tid < 120) tw2[tid] = P.tw[16 * (tid & 7) * ((tid >> 3) + 1)];
  int ia, ib;
  v2_thread_butterflies(tid, &ia, &ib);
  const bool special = (tid == kV2Threads - 1);
  const int ka = v2_klow(ia), kb = v2_klow(ib);
  const int j2 = tid & 7;
  cd* const b1 = buf + pad_idx(tid);
  cd* const b2 = buf + pad_idx((tid >> 3) * 128 + j2);
  cd* const b3a = buf + 9 * ia;
  cd* const b3b = buf + 9 * ib;
  const cd* const t2 = tw2 + j2;

  for (int i = tid; i < span; i += kV2Threads) acc[i] = 0.f;
  __syncthreads();

  for (long long f = f_lo; f <= f_hi; f += 2) {
    const bool two = (f + 1) <= f_hi;
    float2 S[2][9];
    cd v[16];
#pragma unroll
    for (int g = 0; g < 2; ++g) {
#pragma unroll
      for (int j = 0; j < 9; ++j) S[g][j] = make_float2(0.f, 0.f);
      if (g == 0 || two) {
        const long long start = (f + g) * hop - N / 2;
        if (start >= 0 && start + N <= L) {
#pragma unroll
          for (int r = 0; r < 16; ++r) {
            const double w = __ldg(P.win_half + tid + 128 * r);
            v[r] = cd{w * (double)__ldg(xa + start + tid + 128 * r), w * (double)__ldg(xb + start + tid + 128 * r)};
          }
        } else {
#pragma unroll 1
          for (int r = 0; r < 16; ++r) {
            const long long idx = sp_reflect(start + tid + 128 * r, L);
            const double w = __ldg(P.win_half + tid + 128 * r);
            const cd val{w * (double)__ldg(xa + idx), w * (double)__ldg(xb + idx)};
#pragma unroll
            for (int rr = 0; rr < 16; ++rr)
              if (rr == r) v[rr] = val;
          }
        }
        bfly16<false>(v);
#pragma unroll
        for (int q = 1; q < 16; ++q) v[q] = cmul(v[q], tw1[q - 1]);
        __syncthreads();
#pragma unroll
        for (int q = 0; q < 16; ++q) b1[144 * q] = v[q];
        __syncthreads();
#pragma unroll
        for (int r = 0; r < 16; ++r) v[r] = b2[9 * r];
        // last radix-4 stage group by group with the twiddles of the next group fetched ahead (fft_core.cuh)
        bfly16_first<false>(v);
        bfly16_second_twiddled<8>(v, t2, [&](int q, cd val) { b2[9 * q] = val; });
        __syncthreads();
        cd* a = v;
        cd* b = v + 8;
#pragma unroll
        for (int r = 0; r < 8; ++r) {
          a[r] = b3a[r];
          b[r] = b3b[r];
        }
        bfly8<false>(a);
        bfly8<false>(b);
        auto pick = [&](int k, cd zk, cd zn) {
          // X = (Z[k] + conj Z[N-k]) / 2 (input), O = (Z[k] - conj Z[N-k]) / (2i) (model output); c64 rounding
          const float2 X = make_float2((float)(zk.x + zn.x), (float)(zk.y - zn.y));
          const float2 O = make_float2((float)(zk.y + zn.y), (float)(zn.x - zk.x));
          return k < cut ? X : O;
        };
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          const cd za = special ? a[(8 - q) & 7] : b[7 - q];
          const cd zb = special ? b[7 - q] : a[7 - q];
          S[g][q] = pick(ka + 256 * q, a[q], za);
          S[g][4 + q] = pick(kb + 256 * q, b[q], zb);
        }
        if (special) S[g][8] = pick(1024, a[4], a[4]);
      }
    }
    // ---- Y = S_f + i S_{f+1}; irfft ignores the imaginary parts of DC and Nyquist
    {
      cd* a = v;
      cd* b = v + 8;
      auto yk = [&](float2 s0, float2 s1) { return cd{(double)s0.x - (double)s1.y, (double)s0.y + (double)s1.x}; };
      auto ynk = [&](float2 s0, float2 s1) { return cd{(double)s0.x + (double)s1.y, (double)s1.x - (double)s0.y}; };
      if (!special) {
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          a[q] = yk(S[0][q], S[1][q]);
          b[7 - q] = ynk(S[0][q], S[1][q]);
          b[q] = yk(S[0][4 + q], S[1][4 + q]);
          a[7 - q] = ynk(S[0][4 + q], S[1][4 + q]);
        }
      } else {
        a[0] = cd{(double)S[0][0].x, (double)S[1][0].x};
        a[4] = cd{(double)S[0][8].x, (double)S[1][8].x};
#pragma unroll
        for (int q = 1; q < 4; ++q) {
          a[q] = yk(S[0][q], S[1][q]);
          a[8 - q] = ynk(S[0][q], S[1][q]);
        }
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          b[q] = yk(S[0][4 + q], S[1][4 + q]);
          b[7 - q] = ynk(S[0][4 + q], S[1][4 + q]);
        }
      }
      bfly8<true>(a);
      bfly8<true>(b);
#pragma unroll
      for (int r = 0; r < 8; ++r) {
        b3a[r] = a[r];
        b3b[r] = b[r];
      }
    }
    __syncthreads();
    v[0] = b2[0];
#pragma unroll
    for (int q = 1; q < 16; ++q) v[q] = cmul_conj(b2[9 * q], t2[(q - 1) * 8]);
    bfly16<true>(v);
#pragma unroll
    for (int r = 0; r < 16; ++r) b2[9 * r] = v[r];
    __syncthreads();
    v[0] = b1[0];
#pragma unroll
    for (int q = 1; q < 16; ++q) v[q] = cmul_conj(b1[144 * q], tw1[q - 1]);
    bfly16<true>(v);
    // ---- overlap-add into float32, frame f then f+1 (librosa: y32 += float64 block)
#pragma unroll
    for (int r = 0; r < 16; ++r) {
      const long long m = f * hop + tid + 128 * r;
      if (m >= m0 && m < m1) {
        const double term = __ldg(P.win_over_n + tid + 128 * r) * v[r].x;
        acc[m - m0] = (float)((double)acc[m - m0] + term);
      }
    }
    __syncthreads();
    if (two) {
#pragma unroll
      for (int r = 0; r < 16; ++r) {
        const long long m = (f + 1) * hop + tid + 128 * r;
        if (m >= m0 && m < m1) {
          const double term = __ldg(P.win_over_n + tid + 128 * r) * v[r].y;
          acc[m - m0] = (float)((double)acc[m - m0] + term);
        }
      }
    }
    __syncthreads();
  }

  // librosa.filters.window_sumsquare: float32 buffer += float64 hann^2, frame by frame.  Interior samples (every
  // frame that can cover them exists) take the sum from a plan table built with the same additions in the same
  // order, indexed by m % hop through 32-bit counters (m0 = blockIdx.x * chunk_hops * hop + N/2); only the ends
  // of an utterance run the 64-bit divisions and the loop.
  {
    // interior samples [lo, hi) of the item: multiplied by the tabulated reciprocal of the window sum (<= 1 ulp from the
    // quotient; the division behind a dependent table load was ~6 % of this kernel); the ends of an utterance keep the
    // exact sum and division (see K4, stft_lowpass.cu)
    const long long q0 = (long long)blockIdx.x * chunk_hops;
    const long long lo_ll = (long long)(N - hop) - m0, hi_ll = (T - q0) * hop - N / 2;
    const int lo = (int)max(0LL, min((long long)span, lo_ll));
    const int hi = (int)max((long long)lo, min((long long)span, hi_ll));
    auto slow = [&](int i) {
      const long long m = m0 + i;
      long long fa = (m - N >= 0) ? (m - N) / hop + 1 : 0;
      long long fb = min(T - 1, m / hop);
      float ws = 0.f;
      for (long long f = fa; f <= fb; ++f) ws = (float)((double)ws + __ldg(P.win_sq + (m - f * hop)));
      float val = acc[i];
      if (ws > 1.17549435e-38f) val = val / ws;
      y[off + (m - N / 2)] = val;
    };
    for (int i = tid; i < lo; i += kV2Threads) slow(i);
    for (int i = hi + tid; i < span; i += kV2Threads) slow(i);
    const float* inv = P.ws_tab + hop;
    float* yo = y + off + (m0 - N / 2);
    int i = lo + tid;
    int r = (N / 2 + i) % hop;
    const int step = kV2Threads % hop;
    for (; i + 3 * kV2Threads < hi; i += 4 * kV2Threads) {
      int r1 = r + step, r2, r3;
      if (r1 >= hop) r1 -= hop;
      r2 = r1 + step;
      if (r2 >= hop) r2 -= hop;
      r3 = r2 + step;
      if (r3 >= hop) r3 -= hop;
      const float w0 = __ldg(inv + r), w1 = __ldg(inv + r1), w2 = __ldg(inv + r2), w3 = __ldg(inv + r3);
      yo[i] = acc[i] * w0;
      yo[i + kV2Threads] = acc[i + kV2Threads] * w1;
      yo[i + 2 * kV2Threads] = acc[i + 2 * kV2Threads] * w2;
      yo[i + 3 * kV2Threads] = acc[i + 3 * kV2Threads] * w3;
      r = r3 + step;
      if (r >= hop) r -= hop;
    }
    for (; i < hi; i += kV2Threads) {
      yo[i] = acc[i] * __ldg(inv + r);
      r += step;
      if (r >= hop) r -= hop;
    }
  }
}

}  // namespace ssr

using namespace ssr;

extern "C" {

int ssr_splice_plan_create(ssr_splice_plan** out, int n_fft, int hop) {
  if (!out) return fail(SSR_ERR_INVALID, "plan pointer is NULL");
  *out = nullptr;
  if (n_fft != 2048) return fail(SSR_ERR_INVALID, "splice/istft supports n_fft 2048 (librosa default) only");
  if (hop < 1 || hop > n_fft) return fail(SSR_ERR_INVALID, "hop must be in [1, n_fft]");
  const int N = n_fft;
  size_t o = 0;
  size_t o_tw = o;
  o = align_up(o + sizeof(cd) * (size_t)N, 256);
  size_t o_wh = o;
  o = align_up(o + sizeof(double) * (size_t)N, 256);
  size_t o_wn = o;
  o = align_up(o + sizeof(double) * (size_t)N, 256);
  size_t o_w2 = o;
  o = align_up(o + sizeof(double) * (size_t)N, 256);
  size_t o_ws = o;
  o = align_up(o + sizeof(float) * 2 * (size_t)hop, 256);  // [0, hop): the sums, [hop, 2 hop): their reciprocals
  std::vector<unsigned char> host(o, 0);
  cd* tw = reinterpret_cast<cd*>(host.data() + o_tw);
  double* wh = reinterpret_cast<double*>(host.data() + o_wh);
  double* wn = reinterpret_cast<double*>(host.data() + o_wn);
  double* w2 = reinterpret_cast<double*>(host.data() + o_w2);
  for (int n = 0; n < N; ++n) {
    long double a = -2 * kPiL * (long double)n / (long double)N;
    tw[n] = cd{(double)cosl(a), (double)sinl(a)};
    double h = (double)(0.5L - 0.5L * cosl(2 * kPiL * (long double)n / (long double)N));
    wh[n] = 0.5 * h;
    wn[n] = h / (double)N;
    w2[n] = h * h;
  }
  float* ws_tab = reinterpret_cast<float*>(host.data() + o_ws);
  for (int r = 0; r < hop; ++r) {  // frames ascending = window offsets r + j*hop descending
    volatile float ws = 0.f;
    for (int j = (N - 1 - r) / hop; j >= 0; --j) ws = (float)((double)ws + w2[r + j * hop]);
    ws_tab[r] = ws;
    ws_tab[hop + r] = ws > 1.17549435e-38f ? 1.0f / (float)ws : 1.0f;  // librosa divides only where the sum is > tiny
  }
  ssr_splice_plan* p = new ssr_splice_plan();
  p->n_fft = N;
  p->hop = hop;
  p->blob = nullptr;
  cudaError_t e = cudaGetDevice(&p->device);
  if (e == cudaSuccess) e = cudaMalloc(&p->blob, o);
  if (e == cudaSuccess) e = cudaMemcpy(p->blob, host.data(), o, cudaMemcpyHostToDevice);
  if (e != cudaSuccess) {
    if (p->blob) cudaFree(p->blob);
    delete p;
    return fail(SSR_ERR_CUDA, std::string("splice plan upload: ") + cudaGetErrorString(e));
  }
  unsigned char* d = static_cast<unsigned char*>(p->blob);
  p->tw = reinterpret_cast<const cd*>(d + o_tw);
  p->win_half = reinterpret_cast<const double*>(d + o_wh);
  p->win_over_n = reinterpret_cast<const double*>(d + o_wn);
  p->win_sq = reinterpret_cast<const double*>(d + o_w2);
  p->ws_tab = reinterpret_cast<const float*>(d + o_ws);
  *out = p;
  return SSR_OK;
}

int ssr_splice_plan_destroy(ssr_splice_plan* plan) {
  if (!plan) return SSR_OK;
  if (plan->blob) cudaFree(plan->blob);
  delete plan;
  return SSR_OK;
}

int ssr_stft_splice_istft_batched(const ssr_splice_plan* plan, const float* x_dev, const float* out_dev,
                                  const int64_t* offsets_host, const int64_t* offsets_dev, int n,
                                  const int32_t* cut_bins_dev, float* y_dev, void* stream) {
  if (!plan || !x_dev || !out_dev || !offsets_host || !offsets_dev || !cut_bins_dev || !y_dev || n < 1)
    return fail(SSR_ERR_INVALID, "ssr_stft_splice_istft_batched: bad argument");
  if (int rc0 = check_offsets(offsets_host, n, "ssr_stft_splice_istft_batched")) return rc0;
  long long max_len = 0;
  for (int u = 0; u < n; ++u) {
    long long L = offsets_host[u + 1] - offsets_host[u];
    if (L < 1) return fail(SSR_ERR_INVALID, "empty utterance in batch");
    if (L > max_len) max_len = L;
  }
  int ch = 16384 / plan->hop;  // accumulator <= 64 KB
  if (ch > 32) ch = 32;
  if (ch < 1) ch = 1;
  SpDev P{plan->hop, plan->tw, plan->win_half, plan->win_over_n, plan->win_sq, plan->ws_tab};
  size_t smem = sizeof(cd) * (size_t)(2048 + 256) + sizeof(float) * (size_t)ch * plan->hop;
  SSR_CUDA_TRY(cudaFuncSetAttribute(k_stft_splice_istft_2048, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                    (int)smem));
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  long long per = (long long)ch * plan->hop;
  unsigned gx = (unsigned)((max_len + per - 1) / per);
  for (int u0 = 0; u0 < n; u0 += 32768) {
    int nu = n - u0 < 32768 ? n - u0 : 32768;
    k_stft_splice_istft_2048<<<dim3(gx, nu), kV2Threads, smem, st>>>(
        P, x_dev, out_dev, reinterpret_cast<const long long*>(offsets_dev), cut_bins_dev, y_dev, u0, ch);
    SSR_LAUNCH_CHECK("k_stft_splice_istft_2048");
  }
  return SSR_OK;
}

}  // extern "C"
