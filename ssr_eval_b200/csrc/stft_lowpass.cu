// stft_lowpass.cu -- K4: STFT hard low-pass (STFT -> zero bins >= cut -> ISTFT), float32.
//
// Replaces (paths relative to the reference repo):
//   ssr_eval/lowpass.py:17-28   stft_hard_lowpass_v0
//   ssr_eval/dsp.py:76-81       spectrogram_phase: mag = clamp(re^2+im^2, eps)^0.5, cos, sin
//   ssr_eval/dsp.py:83-105      wav_to_spectrogram_phase (eps = 1e-8)
//   ssr_eval/dsp.py:107-119     spectrogram_phase_to_wav -> torchlibrosa ISTFT
// torchlibrosa computes the STFT/ISTFT as dense conv1d DFTs in float32; here each CTA runs a
// shared-memory float32 FFT instead, two real frames packed per complex transform:
//   z = w*x_f + i*w*x_{f+1} -> FFT -> split -> zero k >= cut (generic kernel: + the mag/cos/sin round trip)
//   -> Y = X'_f + i*X'_{f+1} (Hermitian extended) -> inverse FFT -> Re = frame f, Im = frame f+1
//   -> x window / n_fft -> overlap-add in shared memory -> / clamp(sum window^2, 1e-11) -> trim.
// A work item owns `chunk_hops` hops of output samples and recomputes the <= n_fft/hop halo frames.
#include <math.h>
#include <stdlib.h>

#include <vector>

#include "common.cuh"
#include "fft_core.cuh"
#include "k1_map.cuh"

struct ssr_lowpass_plan {
  int n_fft, hop, logM, device;
  void* blob;
  const ssr::cf* tw;         // exp(-2 pi i n / N), float
  const float* win;          // float32(hann_periodic[n])
  const float* win_over_n;   // float32(hann[n] / N)
  const float* win_sq;       // float32(hann[n]^2)
  const float* ws_tab;       // [hop]: overlap-added window^2 of an interior sample m, index m % hop; then [hop]: 1 / max(., 1e-11)
  const uint16_t* ppos;      // padded slot of frequency k after the DIF passes
};

namespace ssr {

constexpr int kLpThreads = 256;

struct LpDev {
  int N, hop;
  const cf* tw;
  const float* win;
  const float* win_over_n;
  const float* win_sq;
  const uint16_t* ppos;
  const float* ws_tab;
};

__device__ __forceinline__ long long lp_reflect(long long i, long long L) {
  if (i >= 0 && i < L) return i;
  if (L == 1) return 0;
  long long period = 2 * (L - 1);
  i %= period;
  if (i < 0) i += period;
  return i < L ? i : period - i;
}

struct LpSync {
  __device__ __forceinline__ void operator()() const { __syncthreads(); }
};

// mag/cos/sin round trip of one bin (dsp.py:76-81 with eps 1e-8, lowpass.py:24-25 zeroing)
__device__ __forceinline__ cf phase_roundtrip(float re, float im, bool keep) {
  float mag = sqrtf(fmaxf(re * re + im * im, 1e-8f));
  float c = re / mag, s = im / mag;
  if (!keep) mag = 0.f;
  return cf{mag * c, mag * s};
}

template <int LOGM>
__global__ void __launch_bounds__(kLpThreads)
k_stft_hard_lowpass(LpDev P, const float* __restrict__ x, const long long* __restrict__ offsets,
                    const int* __restrict__ cut_bins, float* __restrict__ y, int u0, int chunk_hops) {
  constexpr int N = 1 << LOGM;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  cf* buf = reinterpret_cast<cf*>(smem_raw);
  float* acc = reinterpret_cast<float*>(smem_raw + sizeof(cf) * padded_size(N));

  const int tid = threadIdx.x;
  const int u = u0 + blockIdx.y;
  const long long off = offsets[u];
  const long long L = offsets[u + 1] - off;
  const int hop = P.hop;
  const long long n0 = (long long)blockIdx.x * chunk_hops * hop;  // first output sample of this item
  if (n0 >= L) return;
  const long long n1 = min(L, n0 + (long long)chunk_hops * hop);
  const long long m0 = n0 + N / 2, m1 = n1 + N / 2;  // overlap-add coordinates
  const int span = (int)(m1 - m0);
  const long long T = L / hop + 1;  // frames of the centred STFT (L + 2*(N/2) - N) / hop + 1
  const long long f_lo = (m0 - N >= 0) ? (m0 - N) / hop + 1 : 0;
  const long long f_hi = min(T - 1, (m1 - 1) / hop);
  const int cut = cut_bins[u];
  const float* xu = x + off;

  for (int i = tid; i < span; i += kLpThreads) acc[i] = 0.f;
  __syncthreads();

  for (long long f = f_lo; f <= f_hi; f += 2) {
    const bool two = (f + 1) <= f_hi;
    const long long s0 = f * hop - N / 2, s1 = s0 + hop;
    for (int n = tid; n < N; n += kLpThreads) {
      float w = P.win[n];
      float a = w * __ldg(xu + lp_reflect(s0 + n, L));
      float b = two ? w * __ldg(xu + lp_reflect(s1 + n, L)) : 0.f;
      buf[pad_idx(n)] = cf{a, b};
    }
    __syncthreads();
    fft_forward_dif<LOGM>(buf, P.tw, tid, kLpThreads, LpSync());
    __syncthreads();
    for (int k = tid; k <= N / 2; k += kLpThreads) {
      const int pa = P.ppos[k], pb = P.ppos[(N - k) & (N - 1)];
      cf a = buf[pa], b = buf[pb];
      // X_f = (Z[k] + conj Z[N-k]) / 2 ; X_{f+1} = (Z[k] - conj Z[N-k]) / (2i)
      float x1r = 0.5f * (a.x + b.x), x1i = 0.5f * (a.y - b.y);
      float x2r = 0.5f * (a.y + b.y), x2i = 0.5f * (b.x - a.x);
      const bool keep = k < cut;
      cf A = phase_roundtrip(x1r, x1i, keep);
      cf B = phase_roundtrip(x2r, x2i, keep);
      if (k == 0 || k == N / 2) {  // the imaginary parts of DC / Nyquist never reach the real IDFT
        A.y = 0.f;
        B.y = 0.f;
      }
      buf[pa] = cf{A.x - B.y, A.y + B.x};                      // Y[k]   = A + iB
      if (pb != pa) buf[pb] = cf{A.x + B.y, B.x - A.y};        // Y[N-k] = conj(A) + i conj(B)
    }
    __syncthreads();
    fft_inverse_dit<LOGM>(buf, P.tw, tid, kLpThreads, LpSync());
    __syncthreads();
    // overlap-add: frame f from the real part, then frame f+1 from the imaginary part
    for (int n = tid; n < N; n += kLpThreads) {
      long long m = f * hop + n;
      if (m >= m0 && m < m1) acc[m - m0] += P.win_over_n[n] * buf[pad_idx(n)].x;
    }
    __syncthreads();
    if (two) {
      for (int n = tid; n < N; n += kLpThreads) {
        long long m = (f + 1) * hop + n;
        if (m >= m0 && m < m1) acc[m - m0] += P.win_over_n[n] * buf[pad_idx(n)].y;
      }
    }
    __syncthreads();
  }

  for (int i = tid; i < span; i += kLpThreads) {
    const long long m = m0 + i;
    long long fa = (m - N >= 0) ? (m - N) / hop + 1 : 0;
    long long fb = min(T - 1, m / hop);
    float ws = 0.f;
    for (long long f = fa; f <= fb; ++f) ws += P.win_sq[m - f * hop];
    ws = fmaxf(ws, 1e-11f);
    y[off + (m - N / 2)] = acc[i] / ws;
  }
}

// ---------------------------------------------------------------------------------------------
// Specialised n_fft = 2048 kernel (the only size the reference uses, dsp.py:9): 128 threads, radix
// 16 x 16 x 8 forward DIF and 8 x 16 x 16 inverse DIT in float32.  As in K1 every thread owns a last-
// pass butterfly AND its Hermitian partner (k1_map.cuh), so the split of the two packed frames, the
// zeroing and the re-packing all happen in registers between the forward
// pass 3 and the inverse pass 1 -- 4 shared-memory exchanges per frame PAIR instead of 8.
// float2 slots are padded as i + i/16 (conflict-free for the 8-byte accesses of all three passes).
// ---------------------------------------------------------------------------------------------
SSR_HD int pad16(int i) { return i + (i >> 4); }

__device__ __forceinline__ float rcp_approx(float x) {
  float r;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
  return r;
}
__device__ __forceinline__ float sqrt_approx(float x) {
  float r;
  asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
  return r;
}

// mag = clamp(re^2+im^2, 1e-8)^0.5, cos = re/mag, sin = im/mag, out = mag*cos, mag*sin (or 0)
__device__ __forceinline__ cf phase_roundtrip_fast(float re, float im, bool keep) {
  const float mag = sqrt_approx(fmaxf(re * re + im * im, 1e-8f));
  const float inv = rcp_approx(mag);
  const float m = keep ? mag : 0.f;
  return cf{m * (re * inv), m * (im * inv)};
}

__global__ void __launch_bounds__(kV2Threads, 3)
k_stft_hard_lowpass_2048(LpDev P, const float* __restrict__ x, const long long* __restrict__ offsets,
                         const int* __restrict__ cut_bins, float* __restrict__ y, int u0,
                         int chunk_hops) {
  constexpr int N = 2048;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  cf* const buf = reinterpret_cast<cf*>(smem_raw);                       // N + N/16 slots
  float* const acc = reinterpret_cast<float*>(smem_raw + sizeof(cf) * (N + N / 16));
  __shared__ __align__(8) cf tw2[15 * 8];

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
  const long long T = L / hop + 1;
  const long long f_lo = (m0 - N >= 0) ? (m0 - N) / hop + 1 : 0;
  const long long f_hi = min(T - 1, (m1 - 1) / hop);
  const int cut = cut_bins[u];
  const float* xu = x + off;

  cf tw1[15];
#pragma unroll
  for (int q = 1; q < 16; ++q) tw1[q - 1] = P.tw[tid * q];
  if (tid < 120) tw2[tid] = P.tw[16 * (tid & 7) * ((tid >> 3) + 1)];
  int ia, ib;
#ifdef SSR_WARPLOCAL
  // passes 2 / 3 (and inverse passes 1 / 2) of a sub-transform block and of its Hermitian-partner block run in one
  // warp (k1_map.cuh): those two exchanges are warp-local, __syncwarp() instead of a CTA barrier
  bool special;
  v2w_thread_butterflies(tid, &ia, &ib, &special);
  const int blk2 = v2w_pass2_block(tid);
#define SSR_SYNC_LOCAL() __syncwarp()
#else
  v2_thread_butterflies(tid, &ia, &ib);
  const bool special = (tid == kV2Threads - 1);
  const int blk2 = tid >> 3;
#define SSR_SYNC_LOCAL() __syncthreads()
#endif
  const int ka = v2_klow(ia), kb = v2_klow(ib);
  const int j2 = tid & 7;
  cf* const b1 = buf + pad16(tid);                      // element tid + 128 q -> b1[136 q]
  cf* const b2 = buf + pad16(blk2 * 128 + j2);          // element base + 8 r   -> b2[8 r + r/2]
  cf* const b3a = buf + 8 * ia + (ia >> 1);             // element 8 i + r      -> b3[r]
  cf* const b3b = buf + 8 * ib + (ib >> 1);
  const cf* const t2 = tw2 + j2;
  float wn[16];  // window / n_fft of this thread's 16 output samples
#pragma unroll
  for (int r = 0; r < 16; ++r) wn[r] = P.win_over_n[tid + 128 * r];

  for (int i = tid; i < span; i += kV2Threads) acc[i] = 0.f;
  __syncthreads();

  // samples of the NEXT interior frame pair, loaded while the current pair is transformed (32 registers);
  // the analysis window is wn * N exactly (N is a power of two), so it costs a multiply instead of 16 loads
  float nx0[16], nx1[16];
  long long nx_f = -1;  // frame pair the prefetched samples belong to
  auto prefetch_pair = [&](long long f) {
    const long long s0 = f * hop - N / 2, s1 = s0 + hop;
    if (f <= f_hi && s0 >= 0 && s1 + N <= L) {
      const bool two = (f + 1) <= f_hi;
      const float* p0 = xu + s0 + tid;
      const float* p1 = xu + s1 + tid;
#pragma unroll
      for (int r = 0; r < 16; ++r) {
        nx0[r] = __ldg(p0 + 128 * r);
        nx1[r] = two ? __ldg(p1 + 128 * r) : 0.f;
      }
      nx_f = f;
    }
  };
  prefetch_pair(f_lo);

  for (long long f = f_lo; f <= f_hi; f += 2) {
    const bool two = (f + 1) <= f_hi;
    const long long s0 = f * hop - N / 2, s1 = s0 + hop;
    cf v[16];
    // ---- forward pass 1: z = w*x_f + i*w*x_{f+1}
    if (nx_f == f) {
#pragma unroll
      for (int r = 0; r < 16; ++r) {
        const float w = wn[r] * (float)N;
        v[r] = cf{w * nx0[r], w * nx1[r]};
      }
      prefetch_pair(f + 2);
    } else {
      // edge frame pair (reflect padding): gather through the FFT buffer (free here: the previous
      // pair's inverse-pass-3 loads were followed by two barriers) to keep 64-bit reflect math out of
      // the unrolled path
#pragma unroll 1
      for (int n = tid; n < N; n += kV2Threads) {
        const float a = __ldg(xu + lp_reflect(s0 + n, L));
        const float b = two ? __ldg(xu + lp_reflect(s1 + n, L)) : 0.f;
        buf[n] = cf{a, b};
      }
      __syncthreads();
#pragma unroll
      for (int r = 0; r < 16; ++r) {
        const float w = wn[r] * (float)N;
        const cf raw = buf[tid + 128 * r];
        v[r] = cf{w * raw.x, w * raw.y};
      }
      __syncthreads();  // all raw values are in registers before pass 1 overwrites the buffer
      prefetch_pair(f + 2);
    }
    bfly16<false>(v);
    b1[0] = v[0];
#pragma unroll
    for (int q = 1; q < 16; ++q) b1[136 * q] = cmul(v[q], tw1[q - 1]);
    __syncthreads();
    // ---- forward pass 2
#pragma unroll
    for (int r = 0; r < 16; ++r) v[r] = b2[8 * r + (r >> 1)];
    // last radix-4 stage group by group with the twiddles of the next group fetched ahead (fft_core.cuh)
    bfly16_first<false>(v);
    bfly16_second_twiddled<8>(v, t2, [&](int q, cf val) { b2[8 * q + (q >> 1)] = val; });
    SSR_SYNC_LOCAL();
    // ---- forward pass 3 (registers), bin processing, inverse pass 1 (registers)
    cf* a = v;
    cf* b = v + 8;
#pragma unroll
    for (int r = 0; r < 8; ++r) {
      a[r] = b3a[r];
      b[r] = b3b[r];
    }
    bfly8<false>(a);
    bfly8<false>(b);
    auto process = [&](int k, cf& zk, cf& zn, bool self) {
      // X_f = (Z[k] + conj Z[N-k]) / 2 ; X_{f+1} = (Z[k] - conj Z[N-k]) / (2i)
      const float x1r = 0.5f * (zk.x + zn.x), x1i = 0.5f * (zk.y - zn.y);
      const float x2r = 0.5f * (zk.y + zn.y), x2i = 0.5f * (zn.x - zk.x);
      const bool keep = k < cut;
#ifdef SSR_K4_PHASE_ROUNDTRIP
      cf A = phase_roundtrip_fast(x1r, x1i, keep);
      cf B = phase_roundtrip_fast(x2r, x2i, keep);
#else
      // The reference rebuilds a kept bin as mag * (re / mag), mag * (im / mag) (dsp.py:76-88): the bin itself up to
      // two roundings (also below the 1e-8 clamp: mag = 1e-4 both times).  This FFT mode -- whose float32 transform
      // already differs from the reference's dense products by 2e-5 per sample -- keeps the bin as it is: 9 instructions
      // and two MUFU per bin and signal less, 14 % of the kernel.  The dense mode (K4d) has the reference's exact form.
      cf A = keep ? cf{x1r, x1i} : cf{0.f, 0.f};
      cf B = keep ? cf{x2r, x2i} : cf{0.f, 0.f};
#endif
      if (self) {  // DC / Nyquist: imaginary parts never reach the real IDFT
        A.y = 0.f;
        B.y = 0.f;
      }
      zk = cf{A.x - B.y, A.y + B.x};             // Y[k]   = A + iB
      if (!self) zn = cf{A.x + B.y, B.x - A.y};  // Y[N-k] = conj(A) + i conj(B)
    };
    if (!special) {
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        process(ka + 256 * q, a[q], b[7 - q], false);
        process(kb + 256 * q, b[q], a[7 - q], false);
      }
    } else {
      // butterfly 0 holds k = 256 q (partner (8-q)%8 in the same butterfly), butterfly 8 holds 128 + 256 q
      process(0, a[0], a[0], true);
      process(256, a[1], a[7], false);
      process(512, a[2], a[6], false);
      process(768, a[3], a[5], false);
      process(1024, a[4], a[4], true);
#pragma unroll
      for (int q = 0; q < 4; ++q) process(128 + 256 * q, b[q], b[7 - q], false);
    }
    bfly8<true>(a);
    bfly8<true>(b);
#pragma unroll
    for (int r = 0; r < 8; ++r) {
      b3a[r] = a[r];
      b3b[r] = b[r];
    }
    SSR_SYNC_LOCAL();
    // ---- inverse pass 2
    v[0] = b2[0];
#pragma unroll
    for (int q = 1; q < 16; ++q) v[q] = cmul_conj(b2[8 * q + (q >> 1)], t2[(q - 1) * 8]);
    bfly16<true>(v);
#pragma unroll
    for (int r = 0; r < 16; ++r) b2[8 * r + (r >> 1)] = v[r];
    __syncthreads();
    // ---- inverse pass 3 -> natural order: v[r] = sample tid + 128 r (Re: frame f, Im: frame f+1)
    v[0] = b1[0];
#pragma unroll
    for (int q = 1; q < 16; ++q) v[q] = cmul_conj(b1[136 * q], tw1[q - 1]);
    bfly16<true>(v);
    // ---- overlap-add, frame f then frame f+1 (32-bit indices relative to the item's first sample)
    const int o0 = (int)(f * hop - m0) + tid;
#pragma unroll
    for (int r = 0; r < 16; ++r) {
      const int i = o0 + 128 * r;
      if ((unsigned)i < (unsigned)span) acc[i] += wn[r] * v[r].x;
    }
    __syncthreads();
    if (two) {
#pragma unroll
      for (int r = 0; r < 16; ++r) {
        const int i = o0 + hop + 128 * r;
        if ((unsigned)i < (unsigned)span) acc[i] += wn[r] * v[r].y;
      }
    }
    __syncthreads();
  }

  // Division by the overlap-added window^2.  For an interior sample (every frame that can cover it exists) the
  // sum depends only on m % hop and comes from a table the plan built with the same float32 additions in the
  // same order; m / hop and m % hop are tracked with 32-bit counters (m0 = blockIdx.x * chunk_hops * hop + N/2),
  // so the 64-bit divisions and the 4-5 step loop only run for the few samples at the ends of an utterance.
  {
    // Samples [lo, hi) of the item are interior (every frame that can cover them exists): their window sum depends only
    // on (N/2 + i) % hop and they are multiplied by the tabulated reciprocal -- one FMUL where the IEEE division with
    // its FCHK slow path, behind a dependent table load, made this loop 36 % of the kernel's warp time (ncu source
    // page).  The product differs from the quotient by at most 1 ulp, 3 decades below this kernel's distance to the
    // reference's dense arithmetic.  The few samples at the ends of an utterance keep the exact sum + division.
    const long long q0 = (long long)blockIdx.x * chunk_hops;
    long long lo_ll = (long long)(N - hop) - m0, hi_ll = (T - q0) * hop - N / 2;
    const int lo = (int)max(0LL, min((long long)span, lo_ll));
    const int hi = (int)max((long long)lo, min((long long)span, hi_ll));
    auto slow = [&](int i) {
      const long long m = m0 + i;
      long long fa = (m - N >= 0) ? (m - N) / hop + 1 : 0;
      long long fb = min(T - 1, m / hop);
      float ws = 0.f;
      for (long long f = fa; f <= fb; ++f) ws += P.win_sq[m - f * hop];
      y[off + (m - N / 2)] = acc[i] / fmaxf(ws, 1e-11f);
    };
    for (int i = tid; i < lo; i += kV2Threads) slow(i);
    for (int i = hi + tid; i < span; i += kV2Threads) slow(i);
    const float* inv = P.ws_tab + hop;
    float* yo = y + off + (m0 - N / 2);
    int i = lo + tid;
    int r = (N / 2 + i) % hop;
    const int step = kV2Threads % hop;  // (hop may be smaller than the block)
    for (; i + 3 * kV2Threads < hi; i += 4 * kV2Threads) {  // four independent table loads in flight
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

// SSR_FORCE_GENERIC_K4=1 routes n_fft 2048 through the generic radix-8 kernel (A/B tests only)
static bool force_generic_k4() {
  static int v = -1;
  if (v < 0) {
    const char* e = getenv("SSR_FORCE_GENERIC_K4");
    v = (e && e[0] == '1') ? 1 : 0;
  }
  return v == 1;
}

static int launch_lp_2048(const ssr_lowpass_plan* plan, const float* x, const long long* offs,
                          const int* cut, float* y, int n, long long max_len, cudaStream_t st);

static int lp_chunk_hops(int hop) {
  int c = 14336 / hop;  // accumulator <= 56 KB
  if (c > 32) c = 32;
  if (c < 1) c = 1;
  return c;
}

template <int LOGM>
static int launch_lp(const ssr_lowpass_plan* plan, const float* x, const long long* offs,
                     const int* cut, float* y, int n, long long max_len, cudaStream_t st) {
  LpDev P{plan->n_fft, plan->hop, plan->tw, plan->win, plan->win_over_n, plan->win_sq, plan->ppos, plan->ws_tab};
  const int ch = lp_chunk_hops(plan->hop);
  size_t smem = sizeof(cf) * (size_t)padded_size(plan->n_fft) + sizeof(float) * (size_t)ch * plan->hop;
  auto kern = k_stft_hard_lowpass<LOGM>;
  SSR_CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  long long per = (long long)ch * plan->hop;
  unsigned gx = (unsigned)((max_len + per - 1) / per);
  for (int u0 = 0; u0 < n; u0 += 32768) {
    int nu = n - u0 < 32768 ? n - u0 : 32768;
    kern<<<dim3(gx, nu), kLpThreads, smem, st>>>(P, x, offs, cut, y, u0, ch);
    SSR_LAUNCH_CHECK("k_stft_hard_lowpass");
  }
  return SSR_OK;
}

static int launch_lp_2048(const ssr_lowpass_plan* plan, const float* x, const long long* offs,
                          const int* cut, float* y, int n, long long max_len, cudaStream_t st) {
  LpDev P{plan->n_fft, plan->hop, plan->tw, plan->win, plan->win_over_n, plan->win_sq, plan->ppos, plan->ws_tab};
  const int ch = lp_chunk_hops(plan->hop);
  size_t smem = sizeof(cf) * (size_t)(2048 + 128) + sizeof(float) * (size_t)ch * plan->hop;
  SSR_CUDA_TRY(cudaFuncSetAttribute(k_stft_hard_lowpass_2048, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                    (int)smem));
  long long per = (long long)ch * plan->hop;
  unsigned gx = (unsigned)((max_len + per - 1) / per);
  for (int u0 = 0; u0 < n; u0 += 32768) {
    int nu = n - u0 < 32768 ? n - u0 : 32768;
    k_stft_hard_lowpass_2048<<<dim3(gx, nu), kV2Threads, smem, st>>>(P, x, offs, cut, y, u0, ch);
    SSR_LAUNCH_CHECK("k_stft_hard_lowpass_2048");
  }
  return SSR_OK;
}

}  // namespace ssr

using namespace ssr;

extern "C" {

int ssr_lowpass_plan_create(ssr_lowpass_plan** out, int n_fft, int hop) {
  if (!out) return fail(SSR_ERR_INVALID, "plan pointer is NULL");
  *out = nullptr;
  int logM = 0;
  while ((1 << logM) < n_fft) ++logM;
  if ((1 << logM) != n_fft || logM < 8 || logM > 12)
    return fail(SSR_ERR_INVALID, "stft_hard low-pass needs a power-of-two n_fft in [256, 4096]");
  if (hop < 1 || hop > n_fft) return fail(SSR_ERR_INVALID, "hop must be in [1, n_fft]");
  const long double PI = 3.14159265358979323846264338327950288L;
  const int N = n_fft;
  size_t o = 0;
  size_t o_tw = o;
  o = align_up(o + sizeof(cf) * (size_t)N, 256);
  size_t o_w = o;
  o = align_up(o + sizeof(float) * (size_t)N, 256);
  size_t o_wn = o;
  o = align_up(o + sizeof(float) * (size_t)N, 256);
  size_t o_w2 = o;
  o = align_up(o + sizeof(float) * (size_t)N, 256);
  size_t o_pos = o;
  o = align_up(o + sizeof(uint16_t) * (size_t)N, 256);
  size_t o_ws = o;
  o = align_up(o + sizeof(float) * 2 * (size_t)hop, 256);  // [0, hop): the sums, [hop, 2 hop): their clamped reciprocals
  std::vector<unsigned char> host(o, 0);
  cf* tw = reinterpret_cast<cf*>(host.data() + o_tw);
  float* w = reinterpret_cast<float*>(host.data() + o_w);
  float* wn = reinterpret_cast<float*>(host.data() + o_wn);
  float* w2 = reinterpret_cast<float*>(host.data() + o_w2);
  uint16_t* ppos = reinterpret_cast<uint16_t*>(host.data() + o_pos);
  for (int n = 0; n < N; ++n) {
    long double a = -2 * PI * (long double)n / (long double)N;
    tw[n] = cf{(float)cosl(a), (float)sinl(a)};
    double h = (double)(0.5L - 0.5L * cosl(2 * PI * (long double)n / (long double)N));
    w[n] = (float)h;
    wn[n] = (float)(h / (double)N);
    w2[n] = (float)(h * h);
    ppos[n] = (uint16_t)pad_idx(dif_position(n, logM));
  }
  // interior overlap-added window^2: frames in ascending order = window offsets r + j*hop in DESCENDING j,
  // float32 additions exactly as the kernels' edge loop performs them
  float* ws_tab = reinterpret_cast<float*>(host.data() + o_ws);
  for (int r = 0; r < hop; ++r) {
    volatile float ws = 0.f;
    for (int j = (N - 1 - r) / hop; j >= 0; --j) ws = ws + w2[r + j * hop];
    ws_tab[r] = ws;
    const float clamped = ws > 1e-11f ? (float)ws : 1e-11f;
    ws_tab[hop + r] = 1.0f / clamped;
  }
  ssr_lowpass_plan* p = new ssr_lowpass_plan();
  p->n_fft = N;
  p->hop = hop;
  p->logM = logM;
  p->blob = nullptr;
  cudaError_t e = cudaGetDevice(&p->device);
  if (e == cudaSuccess) e = cudaMalloc(&p->blob, o);
  if (e == cudaSuccess) e = cudaMemcpy(p->blob, host.data(), o, cudaMemcpyHostToDevice);
  if (e != cudaSuccess) {
    if (p->blob) cudaFree(p->blob);
    delete p;
    return fail(SSR_ERR_CUDA, std::string("lowpass plan upload: ") + cudaGetErrorString(e));
  }
  unsigned char* d = static_cast<unsigned char*>(p->blob);
  p->tw = reinterpret_cast<const cf*>(d + o_tw);
  p->win = reinterpret_cast<const float*>(d + o_w);
  p->win_over_n = reinterpret_cast<const float*>(d + o_wn);
  p->win_sq = reinterpret_cast<const float*>(d + o_w2);
  p->ppos = reinterpret_cast<const uint16_t*>(d + o_pos);
  p->ws_tab = reinterpret_cast<const float*>(d + o_ws);
  *out = p;
  return SSR_OK;
}

int ssr_lowpass_plan_destroy(ssr_lowpass_plan* plan) {
  if (!plan) return SSR_OK;
  if (plan->blob) cudaFree(plan->blob);
  delete plan;
  return SSR_OK;
}

int ssr_stft_hard_lowpass_batched(const ssr_lowpass_plan* plan, const float* x_dev,
                                  const int64_t* offsets_host, const int64_t* offsets_dev, int n,
                                  const int32_t* cut_bins_dev, float* y_dev, void* stream) {
  if (!plan || !x_dev || !offsets_host || !offsets_dev || !cut_bins_dev || !y_dev || n < 1)
    return fail(SSR_ERR_INVALID, "ssr_stft_hard_lowpass_batched: bad argument");
  if (int rc0 = check_offsets(offsets_host, n, "ssr_stft_hard_lowpass_batched")) return rc0;
  long long max_len = 0;
  for (int u = 0; u < n; ++u) {
    long long L = offsets_host[u + 1] - offsets_host[u];
    if (L < 1) return fail(SSR_ERR_INVALID, "empty utterance in batch");
    if (L > max_len) max_len = L;
  }
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const long long* offs = reinterpret_cast<const long long*>(offsets_dev);
  switch (plan->logM) {
    case 8: return launch_lp<8>(plan, x_dev, offs, cut_bins_dev, y_dev, n, max_len, st);
    case 9: return launch_lp<9>(plan, x_dev, offs, cut_bins_dev, y_dev, n, max_len, st);
    case 10: return launch_lp<10>(plan, x_dev, offs, cut_bins_dev, y_dev, n, max_len, st);
    case 11:
      if (!force_generic_k4()) return launch_lp_2048(plan, x_dev, offs, cut_bins_dev, y_dev, n, max_len, st);
      return launch_lp<11>(plan, x_dev, offs, cut_bins_dev, y_dev, n, max_len, st);
    case 12: return launch_lp<12>(plan, x_dev, offs, cut_bins_dev, y_dev, n, max_len, st);
    default: return fail(SSR_ERR_INVALID, "unsupported n_fft");
  }
}

}  // extern "C"
