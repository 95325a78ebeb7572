// fft_core.cuh -- in-place shared-memory FFT building blocks (float / double), host+device.
//
// Forward transform  : decimation-in-frequency passes, natural order in -> digit-reversed out.
// Inverse transform  : decimation-in-time passes,  digit-reversed in  -> natural order out
//                      (unnormalised; the exact conjugate-transpose of the forward passes).
// Radices 8 and 4 only: M = 2^m = 8^a * 4^b.  A pass touches, per butterfly, the same R
// locations for read and write, so one buffer suffices (no Stockham ping-pong) and the only
// synchronisation is one barrier between passes.
//
// Bank conflicts: element i lives at pad(i) = i + (i >> 3).  With 16-byte (double2) or 8-byte
// (float2) elements this makes the stride-8 / stride-4 accesses of the last passes and the
// unit-stride accesses of the first passes conflict-free per quarter/half warp.
//
// Everything is SSR_HD so that tests/host_emul.cu can run the very same code on the CPU
// (threads emulated by a loop) -- the container that builds this has no GPU.
#pragma once
#include <stdint.h>

#if defined(__CUDACC__)
#define SSR_HD __host__ __device__ __forceinline__
#else
#define SSR_HD inline
#endif

namespace ssr {

template <typename T>
struct alignas(2 * sizeof(T)) C2 {
  T x, y;
};
using cd = C2<double>;
using cf = C2<float>;

template <typename T>
SSR_HD C2<T> cmul(C2<T> a, C2<T> b) {
  return C2<T>{a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x};
}
template <typename T>
SSR_HD C2<T> cmul_conj(C2<T> a, C2<T> b) {  // a * conj(b)
  return C2<T>{a.x * b.x + a.y * b.y, a.y * b.x - a.x * b.y};
}
template <typename T>
SSR_HD C2<T> cadd(C2<T> a, C2<T> b) {
  return C2<T>{a.x + b.x, a.y + b.y};
}
template <typename T>
SSR_HD C2<T> csub(C2<T> a, C2<T> b) {
  return C2<T>{a.x - b.x, a.y - b.y};
}
// float32 complex add / sub on sm_100: the (re, im) pair is one 64-bit register pair, so FADD2 (PTX add / sub
// .rn.f32x2) does both halves in one instruction -- same IEEE roundings, half the issue slots (K4's float32 FFT)
#if defined(__CUDA_ARCH__) && __CUDA_ARCH__ >= 1000
template <>
__device__ __forceinline__ C2<float> cadd<float>(C2<float> a, C2<float> b) {
  unsigned long long ra, rb, rd;
  asm("mov.b64 %0, {%1, %2};" : "=l"(ra) : "f"(a.x), "f"(a.y));
  asm("mov.b64 %0, {%1, %2};" : "=l"(rb) : "f"(b.x), "f"(b.y));
  asm("add.rn.f32x2 %0, %1, %2;" : "=l"(rd) : "l"(ra), "l"(rb));
  C2<float> d;
  asm("mov.b64 {%0, %1}, %2;" : "=f"(d.x), "=f"(d.y) : "l"(rd));
  return d;
}
template <>
__device__ __forceinline__ C2<float> csub<float>(C2<float> a, C2<float> b) {
  unsigned long long ra, rb, rd;
  asm("mov.b64 %0, {%1, %2};" : "=l"(ra) : "f"(a.x), "f"(a.y));
  asm("mov.b64 %0, {%1, %2};" : "=l"(rb) : "f"(b.x), "f"(b.y));
  asm("sub.rn.f32x2 %0, %1, %2;" : "=l"(rd) : "l"(ra), "l"(rb));
  C2<float> d;
  asm("mov.b64 {%0, %1}, %2;" : "=f"(d.x), "=f"(d.y) : "l"(rd));
  return d;
}
#endif

// multiply by -i (forward) or +i (inverse)
template <bool INV, typename T>
SSR_HD C2<T> mul_mi(C2<T> v) {
  return INV ? C2<T>{-v.y, v.x} : C2<T>{v.y, -v.x};
}

SSR_HD int pad_idx(int i) { return i + (i >> 3); }
SSR_HD int padded_size(int M) { return M + (M >> 3); }

// y_q = sum_r x_r W4^{rq}, W4 = exp(-+ 2 pi i / 4)
template <bool INV, typename T>
SSR_HD void bfly4(C2<T>& x0, C2<T>& x1, C2<T>& x2, C2<T>& x3) {
  C2<T> a0 = cadd(x0, x2), a1 = csub(x0, x2), a2 = cadd(x1, x3), a3 = csub(x1, x3);
  C2<T> a3r = mul_mi<INV>(a3);
  x0 = cadd(a0, a2);
  x2 = csub(a0, a2);
  x1 = cadd(a1, a3r);
  x3 = csub(a1, a3r);
}

template <bool INV, typename T>
SSR_HD void bfly8(C2<T>* x) {
  // even / odd radix-4 sub-transforms
  C2<T> e0 = x[0], e1 = x[2], e2 = x[4], e3 = x[6];
  C2<T> o0 = x[1], o1 = x[3], o2 = x[5], o3 = x[7];
  bfly4<INV>(e0, e1, e2, e3);
  bfly4<INV>(o0, o1, o2, o3);
  const T h = (T)0.70710678118654752440;
  // W8^1 = (1 -+ i)/sqrt2, W8^2 = -+i, W8^3 = (-1 -+ i)/sqrt2
  C2<T> t1 = INV ? C2<T>{(o1.x - o1.y) * h, (o1.x + o1.y) * h}
                 : C2<T>{(o1.x + o1.y) * h, (o1.y - o1.x) * h};
  C2<T> t2 = mul_mi<INV>(o2);
  C2<T> t3 = INV ? C2<T>{(-o3.x - o3.y) * h, (o3.x - o3.y) * h}
                 : C2<T>{(o3.y - o3.x) * h, (-o3.x - o3.y) * h};
  x[0] = cadd(e0, o0);
  x[4] = csub(e0, o0);
  x[1] = cadd(e1, t1);
  x[5] = csub(e1, t1);
  x[2] = cadd(e2, t2);
  x[6] = csub(e2, t2);
  x[3] = cadd(e3, t3);
  x[7] = csub(e3, t3);
}

// multiply by W16^e (forward) / W16^-e (inverse), e in {1,2,3,6,9}; W16^4 = -+i is mul_mi
template <bool INV, int E, typename T>
SSR_HD C2<T> mul_w16(C2<T> v) {
  const T c1 = (T)0.92387953251128675613, s1 = (T)0.38268343236508977173;  // cos, sin(pi/8)
  const T h = (T)0.70710678118654752440;
  // W16^e = cos(e*pi/8) - i sin(e*pi/8) (forward); conjugate for the inverse
  T c, s;
  if (E == 1) { c = c1; s = s1; }
  else if (E == 2) { c = h; s = h; }
  else if (E == 3) { c = s1; s = c1; }
  else if (E == 6) { c = -h; s = h; }
  else { c = -c1; s = -s1; }  // E == 9
  if (INV) s = -s;
  return C2<T>{v.x * c + v.y * s, v.y * c - v.x * s};
}

// 16-point DFT, natural order in and out: y_q = sum_r x_r W16^{rq}.  4x4 decomposition:
// radix-4 over r1 (r = 4 r1 + r0), internal twiddles W16^{r0 q0}, radix-4 over r0 (q = q0 + 4 q1).
template <bool INV, typename T>
SSR_HD void bfly16(C2<T>* x) {
#pragma unroll
  for (int r0 = 0; r0 < 4; ++r0) bfly4<INV>(x[r0], x[4 + r0], x[8 + r0], x[12 + r0]);
  // now x[4*q0 + r0] holds u[r0][q0]
  x[5] = mul_w16<INV, 1>(x[5]);
  x[6] = mul_w16<INV, 2>(x[6]);
  x[7] = mul_w16<INV, 3>(x[7]);
  x[9] = mul_w16<INV, 2>(x[9]);
  x[10] = mul_mi<INV>(x[10]);
  x[11] = mul_w16<INV, 6>(x[11]);
  x[13] = mul_w16<INV, 3>(x[13]);
  x[14] = mul_w16<INV, 6>(x[14]);
  x[15] = mul_w16<INV, 9>(x[15]);
#pragma unroll
  for (int q0 = 0; q0 < 4; ++q0) bfly4<INV>(x[4 * q0], x[4 * q0 + 1], x[4 * q0 + 2], x[4 * q0 + 3]);
  // x[4*q0 + q1] holds y[q0 + 4 q1]: transpose the 4x4 index to natural order
#pragma unroll
  for (int a = 0; a < 4; ++a)
#pragma unroll
    for (int b = a + 1; b < 4; ++b) {
      C2<T> t = x[4 * a + b];
      x[4 * a + b] = x[4 * b + a];
      x[4 * b + a] = t;
    }
}

// The same 16-point DFT in two steps, for callers that interleave the second radix-4 stage with other work (the
// twiddle prefetch of K1's pass 2): after bfly16_first, bfly16_group<q0> leaves y[q0 + 4 q1] in x[4 q0 + q1].
template <bool INV, typename T>
SSR_HD void bfly16_first(C2<T>* x) {
#pragma unroll
  for (int r0 = 0; r0 < 4; ++r0) bfly4<INV>(x[r0], x[4 + r0], x[8 + r0], x[12 + r0]);
  x[5] = mul_w16<INV, 1>(x[5]);
  x[6] = mul_w16<INV, 2>(x[6]);
  x[7] = mul_w16<INV, 3>(x[7]);
  x[9] = mul_w16<INV, 2>(x[9]);
  x[10] = mul_mi<INV>(x[10]);
  x[11] = mul_w16<INV, 6>(x[11]);
  x[13] = mul_w16<INV, 3>(x[13]);
  x[14] = mul_w16<INV, 6>(x[14]);
  x[15] = mul_w16<INV, 9>(x[15]);
}
template <bool INV, int Q0, typename T>
SSR_HD void bfly16_group(C2<T>* x) {
  bfly4<INV>(x[4 * Q0], x[4 * Q0 + 1], x[4 * Q0 + 2], x[4 * Q0 + 3]);
}

#if defined(__CUDACC__)
// shared-memory loads the compiler may not sink towards their use (they are issued where they are written): used to
// fetch the twiddles of a pass AHEAD of the butterflies that need them
__device__ __forceinline__ C2<double> lds_pinned(const C2<double>* p) {
  C2<double> r;
  asm volatile("ld.shared.v2.f64 {%0, %1}, [%2];" : "=d"(r.x), "=d"(r.y) : "r"((unsigned)__cvta_generic_to_shared(p)));
  return r;
}
__device__ __forceinline__ C2<float> lds_pinned(const C2<float>* p) {
  C2<float> r;
  asm volatile("ld.shared.v2.f32 {%0, %1}, [%2];" : "=f"(r.x), "=f"(r.y) : "r"((unsigned)__cvta_generic_to_shared(p)));
  return r;
}

// Second half of a radix-16 DIF pass whose outputs are multiplied by twiddles from a shared table (stride TWS between
// consecutive q) and stored with stride SS: the last radix-4 stage runs group by group (outputs q0, q0+4, q0+8, q0+12)
// and the 4 twiddles of the NEXT group are requested before the current group is computed, multiplied and stored.
// The plain form (load twiddle -> multiply -> store, 15 times) serialises 12 shared-memory latencies at the end of the
// pass with nothing else for the warp to issue (ncu on K1: ~19 % of all warp time on the short scoreboard).
// STORE(q, value) is the caller's store of output q.
template <int TWS, typename T, typename STORE>
__device__ __forceinline__ void bfly16_second_twiddled(C2<T>* v, const C2<T>* tw, STORE store) {
  C2<T> w[2][4];
#pragma unroll
  for (int q1 = 1; q1 < 4; ++q1) w[0][q1] = lds_pinned(tw + (4 * q1 - 1) * TWS);
#pragma unroll
  for (int q0 = 0; q0 < 4; ++q0) {
    if (q0 < 3) {
#pragma unroll
      for (int q1 = 0; q1 < 4; ++q1) w[(q0 + 1) & 1][q1] = lds_pinned(tw + (q0 + 1 + 4 * q1 - 1) * TWS);
    }
    bfly4<false>(v[4 * q0], v[4 * q0 + 1], v[4 * q0 + 2], v[4 * q0 + 3]);
#pragma unroll
    for (int q1 = 0; q1 < 4; ++q1) {
      const int q = q0 + 4 * q1;
      store(q, q > 0 ? cmul(v[4 * q0 + q1], w[q0 & 1][q1]) : v[0]);
    }
  }
}
#endif

template <int R, bool INV, typename T>
SSR_HD void bfly(C2<T>* x) {
  if (R == 16) {
    bfly16<INV>(x);
  } else if (R == 8) {
    bfly8<INV>(x);
  } else {
    bfly4<INV>(x[0], x[1], x[2], x[3]);
  }
}

// number of radix-8 / radix-4 passes for M = 2^logM
SSR_HD constexpr int n_r4(int logM) { return (logM % 3 == 0) ? 0 : ((logM % 3 == 2) ? 1 : 2); }
SSR_HD constexpr int n_r8(int logM) { return (logM - 2 * n_r4(logM)) / 3; }

// digit-reversed position of frequency k after the forward DIF passes
inline int dif_position(int k, int logM) {
  int M = 1 << logM, pos = 0, N = M;
  int a = n_r8(logM), b = n_r4(logM);
  for (int s = 0; s < a + b; ++s) {
    int R = (s < a) ? 8 : 4;
    pos += (k % R) * (N / R);
    k /= R;
    N /= R;
  }
  return pos;
}

// One forward DIF radix-R pass over the whole M-point buffer (sub-transform length Ncur).
// tw = exp(-2 pi i n / M), n in [0, M).
template <int R, typename T>
SSR_HD void dif_pass(C2<T>* buf, int M, int Ncur, const C2<T>* __restrict__ tw, int tid,
                     int nthreads) {
  const int sub = Ncur / R;
  const int tws = M / Ncur;
  for (int i = tid; i < M / R; i += nthreads) {
    const int j = i & (sub - 1);
    const int base = (i - j) * R + j;  // (i / sub) * Ncur + j
    C2<T> x[R];
#pragma unroll
    for (int r = 0; r < R; ++r) x[r] = buf[pad_idx(base + r * sub)];
    bfly<R, false>(x);
    if (sub > 1) {
#pragma unroll
      for (int q = 1; q < R; ++q) x[q] = cmul(x[q], tw[j * q * tws]);
    }
#pragma unroll
    for (int q = 0; q < R; ++q) buf[pad_idx(base + q * sub)] = x[q];
  }
}

// One inverse DIT radix-R pass (conjugate twiddles first, then the conjugate butterfly).
template <int R, typename T>
SSR_HD void dit_pass(C2<T>* buf, int M, int Ncur, const C2<T>* __restrict__ tw, int tid,
                     int nthreads) {
  const int sub = Ncur / R;
  const int tws = M / Ncur;
  for (int i = tid; i < M / R; i += nthreads) {
    const int j = i & (sub - 1);
    const int base = (i - j) * R + j;
    C2<T> x[R];
#pragma unroll
    for (int q = 0; q < R; ++q) x[q] = buf[pad_idx(base + q * sub)];
    if (sub > 1) {
#pragma unroll
      for (int q = 1; q < R; ++q) x[q] = cmul_conj(x[q], tw[j * q * tws]);
    }
    bfly<R, true>(x);
#pragma unroll
    for (int r = 0; r < R; ++r) buf[pad_idx(base + r * sub)] = x[r];
  }
}

// Whole transforms; SYNC is a functor called between passes (and NOT after the last one).
template <int LOGM, typename T, typename SYNC>
SSR_HD void fft_forward_dif(C2<T>* buf, const C2<T>* __restrict__ tw, int tid, int nthreads,
                            SYNC sync) {
  constexpr int M = 1 << LOGM;
  constexpr int A = n_r8(LOGM), B = n_r4(LOGM);
  int N = M;
#pragma unroll
  for (int s = 0; s < A + B; ++s) {
    if (s) sync();
    if (s < A) {
      dif_pass<8>(buf, M, N, tw, tid, nthreads);
      N /= 8;
    } else {
      dif_pass<4>(buf, M, N, tw, tid, nthreads);
      N /= 4;
    }
  }
}

template <int LOGM, typename T, typename SYNC>
SSR_HD void fft_inverse_dit(C2<T>* buf, const C2<T>* __restrict__ tw, int tid, int nthreads,
                            SYNC sync) {
  constexpr int M = 1 << LOGM;
  constexpr int A = n_r8(LOGM), B = n_r4(LOGM);
  // reverse pass order: the radix-4 passes (smallest sub-transforms) first
  int N = 1;
#pragma unroll
  for (int s = A + B - 1; s >= 0; --s) {
    if (s != A + B - 1) sync();
    if (s < A) {
      N *= 8;
      dit_pass<8>(buf, M, N, tw, tid, nthreads);
    } else {
      N *= 4;
      dit_pass<4>(buf, M, N, tw, tid, nthreads);
    }
  }
}

}  // namespace ssr
