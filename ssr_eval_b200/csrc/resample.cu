// resample.cu -- K3: batched polyphase FIR resampler with scipy.signal.resample_poly semantics.
//
// Replaces (paths relative to the reference repo):
//   ssr_eval/lowpass.py:137,140  resample_poly(data, fs_down, fs_ori) / (y, fs_ori, fs_down)
//   ssr_eval/eval.py:144-150     librosa.resample(..., res_type="polyphase") -> resample_poly
// scipy (scipy/signal/_signaltools.py resample_poly + _upfirdn_apply.pyx) does, for float32 x:
//   n_out = ceil(n_in*up/down); half_len = (len(h)-1)/2; n_pre_pad = down - half_len % down;
//   n_pre_remove = (half_len + n_pre_pad) / down; h padded with n_pre_pad zeros in front;
//   y = upfirdn(h_padded, x, up, down)[n_pre_remove : n_pre_remove + n_out]   (zero extension)
// i.e.  y[j] = sum_i x[i] * h[(j + n_pre_remove)*down - n_pre_pad - i*up], accumulated in float32
// in ascending i with a separate multiply and add -- reproduced here term for term.
//
// Kernels: k_resample_pair (the evaluation's sample-rate pairs: TMA-staged span, two outputs
// per thread over one LDS.64 window, table-driven prologue), k_resample_bulk (any K <= 48: TMA-staged span, one
// output per thread-step), k_resample_tiled (index ranges beyond 32 bits), k_resample<T> (everything else, float64).
#include <stdlib.h>

#include <algorithm>
#include <vector>

#include "common.cuh"
#include "f32x2.cuh"
#include "resample_tables.hpp"

struct ssr_resample_plan {
  int up, down, n_taps, half_len, n_pre_pad, n_pre_remove, K, device;
  int floor_len;  // 1: explicit-bank plan (resampy semantics): floor(n_in * up / down) outputs instead of ceil
  int is_f64;   // 1: float64 plan (bank holds doubles), for float64 waveforms
  void* bank;   // [up][K]: bank[phase*K + k] = h[phase + k*up] (0 beyond n_taps); float or double
  // Copies of a float bank for the staged kernels, TRANSPOSED and in OUTPUT order: column n holds the phase of the
  // outputs j = n (mod up), phase(n) = (half_len + n * down) % up, so that the threads of a warp (consecutive outputs)
  // fetch their taps from consecutive addresses (ncu: with bank[phase][k] every tap load touched 32 lines), and with a
  // FIXED row stride kBankStride so that the K loads of a thread are one base address + compile-time offsets.
  //   bank_t[k * kBankStride + n] = bank[phase(n)][k]                                    (k_resample_bulk)
  // k_resample_pair (pair_nl > 0) gets the two filters of the output pair (j, j + 1), j = n (mod up), already shifted
  // onto its even-aligned window of 2 * pair_nl words, for both alignments par = 0 / 1 of the window's oldest sample:
  //   pair_g[(par * 2 * pair_nl + q) * kBankStride + c] = (gA[2q], gA[2q+1]) for q < pair_nl, (gB[..]) for q >= pair_nl,
  //   c = t % (up / gcd(up, 2)) for thread t of a CTA, n = (2 c) % up (consecutive threads -> consecutive columns),
  //   gA[w] = bank[phase(n)][K - 1 + par - w], gB[w] = bank[phase(n + 1)][K - 1 + par + d(n) - w], 0 outside [0, K),
  //   d(n) = (phase(n) + down) / up = how many samples the window of j + 1 starts after that of j
  // and pair_thr[t] = {newest(jb + 2 t) - newest(jb), c(t)} for the threads t of a CTA (jb = the CTA's first
  // output, a multiple of up, so neither depends on the CTA): no division in the kernel.
  float* bank_t;
  float2* pair_g;
  int2* pair_thr;
  int pair_nl, pair_tp, pair_m;  // 64-bit loads per window (0: no pair kernel), threads per CTA, 2 * pair_tp / up
};

namespace ssr {


__device__ __forceinline__ float mul_rn(float a, float b) { return __fmul_rn(a, b); }
__device__ __forceinline__ float add_rn(float a, float b) { return __fadd_rn(a, b); }
__device__ __forceinline__ double mul_rn(double a, double b) { return __dmul_rn(a, b); }
__device__ __forceinline__ double add_rn(double a, double b) { return __dadd_rn(a, b); }

// one output per thread; T = float (fallback for exotic ratios) or double (float64 waveforms: what
// librosa.resample hands to scipy when a testee returns float64, eval.py:144-150)
template <typename T>
__global__ void __launch_bounds__(256)
k_resample(const T* __restrict__ x, const long long* __restrict__ in_off, T* __restrict__ y,
           const long long* __restrict__ out_off, int u0, int up, int down, int n_pre_pad,
           int n_pre_remove, int K, const T* __restrict__ bank) {
  const int u = u0 + blockIdx.y;
  const long long n_out = out_off[u + 1] - out_off[u];
  const long long j = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= n_out) return;
  const long long n_in = in_off[u + 1] - in_off[u];
  const T* xu = x + in_off[u];
  const long long c = (j + n_pre_remove) * (long long)down - n_pre_pad;
  long long i_hi = c / up;
  long long phase = c - i_hi * up;
  if (phase < 0) {  // floor division for negative c
    phase += up;
    i_hi -= 1;
  }
  const T* hb = bank + phase * K;
  T acc = 0;
  for (int k = K - 1; k >= 0; --k) {
    long long i = i_hi - k;
    if (i >= 0 && i < n_in) acc = add_rn(acc, mul_rn(__ldg(xu + i), __ldg(hb + k)));
  }
  y[out_off[u] + j] = acc;
}

// Register-tiled variant.  Outputs j and j + TP (TP a multiple of `up`) share their polyphase phase, so a
// thread keeps its phase's K <= KMAX taps in registers and produces R outputs j, j+TP, ..., j+(R-1)TP;
// only x is loaded in the inner loop (through L1: consecutive lanes read consecutive-ish samples).
// The float32 multiply / add order per output is unchanged (ascending input index), so results stay
// bit-identical to scipy.
// EXACT: K == KMAX (the common ratios get their own instantiation, so the unrolled tap loops carry no run-time
// `k < K` predicates and no dead iterations).
// The input span a CTA needs ((TP*R - 1) * down / up + K samples, 9-19 KB) is staged in shared memory first,
// zero-filled outside the utterance: scipy's zero extension then needs no bounds logic (adding x*h with x = 0
// leaves the float32 sum unchanged), and the ~21 loads per output are LDS instead of LDG -- the global-load
// instruction queue was what bound the first version (ncu: lg_throttle 4.0, long_scoreboard 5.4 warps / issue).
template <int KMAX, int R, bool EXACT>
__global__ void __launch_bounds__(512)
k_resample_tiled(const float* __restrict__ x, const long long* __restrict__ in_off, float* __restrict__ y,
                 const long long* __restrict__ out_off, int u0, int up, int down, int n_pre_pad,
                 int n_pre_remove, int K, const float* __restrict__ bank, int span) {
  extern __shared__ float xs[];
  const int u = u0 + blockIdx.y;
  const long long n_out = out_off[u + 1] - out_off[u];
  const int TP = blockDim.x;
  const long long jb = (long long)blockIdx.x * TP * R;  // first output of this CTA
  if (jb >= n_out) return;
  const long long n_in = in_off[u + 1] - in_off[u];
  const float* xu = x + in_off[u];
  float* yu = y + out_off[u];
  // input index of the newest sample of output j: floor(((j + n_pre_remove) * down - n_pre_pad) / up)
  auto newest = [&](long long j, long long* phase) {
    const long long c = (j + n_pre_remove) * (long long)down - n_pre_pad;
    long long i = c / up;
    long long ph = c - i * up;
    if (ph < 0) {  // floor division for negative c
      ph += up;
      i -= 1;
    }
    *phase = ph;
    return i;
  };
  long long ph0;
  const long long i_base = newest(jb, &ph0) - (K - 1);  // oldest sample the CTA touches
  for (int i = threadIdx.x; i < span; i += TP) {
    const long long gi = i_base + i;
    xs[i] = (gi >= 0 && gi < n_in) ? __ldg(xu + gi) : 0.f;
  }
  __syncthreads();

  const long long j0 = jb + threadIdx.x;
  long long phase;
  const int p0 = (int)(newest(j0, &phase) - i_base);  // index of the newest sample of output j0 inside xs
  float h[KMAX];
#pragma unroll
  for (int k = 0; k < KMAX; ++k) h[k] = (EXACT || k < K) ? __ldg(bank + phase * K + k) : 0.f;
  const int step = (TP / up) * down;  // input advance per TP outputs (TP % up == 0)
  // G outputs are accumulated together: G independent float32 add chains and G*K loads in flight
  constexpr int G = 4;
  static_assert(R % G == 0, "R must be a multiple of G");
#pragma unroll 1
  for (int i0 = 0; i0 < R; i0 += G) {
    float acc[G];
    const float* px[G];
#pragma unroll
    for (int g = 0; g < G; ++g) {
      acc[g] = 0.f;
      px[g] = xs + p0 + (i0 + g) * step;
    }
#pragma unroll
    for (int k = KMAX - 1; k >= 0; --k)
      if (EXACT || k < K) {
#pragma unroll
        for (int g = 0; g < G; ++g) acc[g] = __fadd_rn(acc[g], __fmul_rn(px[g][-k], h[k]));
      }
#pragma unroll
    for (int g = 0; g < G; ++g) {
      const long long j = j0 + (long long)(i0 + g) * TP;
      if (j < n_out) yu[j] = acc[g];
    }
  }
}

// ---------------------------------------------------------------------------------------------
// k_resample_bulk: the production kernel for utterances whose index arithmetic fits 32 bits.  Same per-output
// float32 operation order as k_resample_tiled (bit-identical to scipy), but
//   * the input span of the CTA is staged by the TMA engine: ONE cp.async.bulk (global -> shared, completion on an
//     mbarrier) issued by thread 0 instead of a per-element LDG + bounds test + STS loop (~10 instructions per
//     staged sample); the threads compute their phases and fetch their taps while the copy is in flight.  Tiles
//     that touch an end of the utterance (zero extension) or the end of the batch buffer keep the element-wise staging;
//   * all index arithmetic is 32-bit (ncu on the previous kernel: 129 instructions per output against 63 in the
//     tap loop -- two 64-bit divisions per thread and the staging loop made up the rest);
//   * a thread produces R = 16 outputs, so the prologue is amortised over twice as many.
// c = j * down + half_len is the position of output j on the up-sampled grid (>= 0: n_pre_remove * down - n_pre_pad
// = half_len by construction), newest input sample i = c / up, phase = c % up.
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ void mbar_init(unsigned long long* bar, unsigned count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"((unsigned)__cvta_generic_to_shared(bar)), "r"(count) : "memory");
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void bulk_load_1d(void* smem_dst, const void* gmem_src, unsigned bytes, unsigned long long* bar) {
  const unsigned b = (unsigned)__cvta_generic_to_shared(bar);
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(b), "r"(bytes) : "memory");
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                   (unsigned)__cvta_generic_to_shared(smem_dst)),
               "l"(gmem_src), "r"(bytes), "r"(b)
               : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned long long* bar, unsigned parity) {
  const unsigned b = (unsigned)__cvta_generic_to_shared(bar);
  unsigned done = 0;
  while (!done) {
    asm volatile(
        "{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(done)
        : "r"(b), "r"(parity)
        : "memory");
  }
}

template <int KMAX, int R, bool EXACT>
__global__ void __launch_bounds__(512)
k_resample_bulk(const float* __restrict__ x, const long long* __restrict__ in_off, float* __restrict__ y,
                const long long* __restrict__ out_off, int u0, unsigned up, unsigned down, unsigned half_len, int K,
                const float* __restrict__ bank_t, int span, long long x_total) {
  extern __shared__ __align__(16) float xs_raw[];  // span + 8 floats (alignment slack of the bulk copy)
  __shared__ __align__(8) unsigned long long bar;
  const int u = u0 + blockIdx.y;
  const unsigned n_out = (unsigned)(out_off[u + 1] - out_off[u]);
  const unsigned TP = blockDim.x;
  const unsigned jb = blockIdx.x * TP * R;  // first output of this CTA
  if (jb >= n_out) return;
  const long long xoff = in_off[u];
  const int n_in = (int)(in_off[u + 1] - xoff);
  float* yu = y + out_off[u];
  const unsigned ib = (jb * down + half_len) / up;      // newest sample of the CTA's first output
  const int i_base = (int)ib - (K - 1);                 // oldest sample the CTA touches (negative at the start)
  // the bulk copy needs 16-byte aligned addresses and sizes: copy from the aligned address below the span
  const long long g0 = xoff + i_base;
  const int shift = (int)(g0 & 3);
  const int n_copy = (span + shift + 3) & ~3;
  const bool bulk = i_base >= 0 && i_base + span <= n_in && g0 - shift + n_copy <= x_total;
  const float* xs = xs_raw + (bulk ? shift : 0);
  if (bulk) {
    if (threadIdx.x == 0) mbar_init(&bar, 1);
    __syncthreads();
    if (threadIdx.x == 0) bulk_load_1d(xs_raw, x + (g0 - shift), (unsigned)n_copy * 4u, &bar);
  } else {
    const float* xu = x + xoff;
    for (int i = threadIdx.x; i < span; i += TP) {
      const int gi = i_base + i;
      xs_raw[i] = (gi >= 0 && gi < n_in) ? __ldg(xu + gi) : 0.f;
    }
  }
  // per-thread phase and taps (overlaps the copy)
  const unsigned j0 = jb + threadIdx.x;
  const unsigned c0 = j0 * down + half_len;
  const unsigned i0 = c0 / up;
  const unsigned n0 = threadIdx.x % up;     // j0 % up (jb is a multiple of up): the row of bank_t with j0's phase
  const int p0 = (int)(i0 - ib) + (K - 1);  // index of the newest sample of output j0 inside xs
  float h[KMAX];
#pragma unroll
  for (int k = 0; k < KMAX; ++k) h[k] = (EXACT || k < K) ? __ldg(bank_t + k * kBankStride + n0) : 0.f;
  const int step = (int)((TP / up) * down);  // input advance per TP outputs (TP % up == 0)
  if (bulk) mbar_wait(&bar, 0);
  else __syncthreads();
  constexpr int G = 4;
  static_assert(R % G == 0, "R must be a multiple of G");
#pragma unroll 1
  for (int r0 = 0; r0 < R; r0 += G) {
    float acc[G];
    const float* px[G];
#pragma unroll
    for (int g = 0; g < G; ++g) {
      acc[g] = 0.f;
      px[g] = xs + p0 + (r0 + g) * step;
    }
#pragma unroll
    for (int k = KMAX - 1; k >= 0; --k)
      if (EXACT || k < K) {
#pragma unroll
        for (int g = 0; g < G; ++g) acc[g] = __fadd_rn(acc[g], __fmul_rn(px[g][-k], h[k]));
      }
#pragma unroll
    for (int g = 0; g < G; ++g) {
      const unsigned j = j0 + (unsigned)(r0 + g) * TP;
      if (j < n_out) yu[j] = acc[g];
    }
  }
}


// ---------------------------------------------------------------------------------------------
// k_resample_pair: k_resample_bulk's successor for the production ratios.  ncu on k_resample_bulk: the LSU binds
// (79 % of its wavefront peak) -- every tap of every output is one 4-byte LDS.  Here a thread produces the two
// CONSECUTIVE outputs j, j+1, whose input windows overlap in all but d = 0 .. ceil(down / up) samples, from ONE set of
// window loads, and those loads are 8-byte LDS.64: the window start is rounded down to an even shared-memory word
// and the two filters are shifted to match ONCE per thread (taps outside [0, K) are zero), so the inner loop has no
// data-dependent indexing.  Per output: 6 LDS.64 + 12 FMUL2 + 24 FADD against 21 LDS + 21 FMUL + 21 FADD.
// The products of two neighbouring taps come from one packed FMUL2 (f32x2.cuh: each half rounded like the scalar
// FMUL); the additions stay one scalar chain per output, oldest sample first -- scipy's order.  A zero tap adds
// x * 0 = +-0 to an accumulator that started at +0 and can therefore never be -0: bit-identical for finite input.
// Outputs j + m * 2 * TP (m = 0 .. RP-1) share j's phase (2 * TP % up == 0) and alignment (their windows are
// `step` = (2 * TP / up) * down samples apart, even by choice of TP), so the shifted filters stay in registers.
// Everything that depends on the thread alone comes from two host-built tables (ssr_resample_plan): the kernel has
// no division and no range test; ncu on the first version showed half of its instructions in that prologue.
// ---------------------------------------------------------------------------------------------
template <int NL, int RP, int MAXT, int MINB>
__global__ void __launch_bounds__(MAXT, MINB)
k_resample_pair(const float* __restrict__ x, const long long* __restrict__ in_off, float* __restrict__ y,
                const long long* __restrict__ out_off, int u0, int K, const float2* __restrict__ pair_g,
                const int2* __restrict__ pair_thr, unsigned ib0, unsigned ib_step, int step, int span,
                long long x_total) {
  extern __shared__ __align__(16) float xs_raw[];  // span + 8 floats (alignment slack of the bulk copy)
  __shared__ __align__(8) unsigned long long bar;
  const int u = u0 + blockIdx.y;
  const long long yoff = out_off[u];
  const unsigned n_out = (unsigned)(out_off[u + 1] - yoff);
  const unsigned TP = blockDim.x;
  const unsigned jb = blockIdx.x * TP * (2 * RP);  // first output of this CTA (a multiple of up)
  if (jb >= n_out) return;
  const long long xoff = in_off[u];
  const int n_in = (int)(in_off[u + 1] - xoff);
  float* yu = y + yoff;
  const unsigned ib = ib0 + blockIdx.x * ib_step;  // newest sample of the CTA's first output: (jb * down + half_len) / up
  const int i_base = (int)ib - (K - 1);            // oldest sample the CTA touches (negative at the start)
  const long long g0 = xoff + i_base;
  const int shift = (int)(g0 & 3);
  const int n_copy = (span + shift + 3) & ~3;
  const bool bulk = i_base >= 0 && i_base + span <= n_in && g0 - shift + n_copy <= x_total;
  const int sh = bulk ? shift : 0;  // sample i sits at xs_raw[i - i_base + sh]
  if (bulk) {
    if (threadIdx.x == 0) mbar_init(&bar, 1);
    __syncthreads();
    if (threadIdx.x == 0) bulk_load_1d(xs_raw, x + (g0 - shift), (unsigned)n_copy * 4u, &bar);
  } else {
    const float* xu = x + xoff;
    for (int i = threadIdx.x; i < span; i += TP) {
      const int gi = i_base + i;
      xs_raw[i] = (gi >= 0 && gi < n_in) ? __ldg(xu + gi) : 0.f;
    }
  }
  // the thread's output pair: window position and the two shifted filters (overlaps the copy)
  const unsigned ja = jb + 2 * threadIdx.x;
  const int2 tt = __ldg(pair_thr + threadIdx.x);  // {newest(ja) - ib, column of pair_g}
  const int a_lo = tt.x + sh;                     // xs_raw index of the oldest sample of output ja
  const int A0 = a_lo & ~1;                       // the even word at or below it: the window is xs_raw[A0 .. A0 + 2 NL)
  float2 gA[NL], gB[NL];
  {
    const float2* g = pair_g + ((a_lo & 1) * (2 * NL) * kBankStride + tt.y);
#pragma unroll
    for (int n = 0; n < NL; ++n) {
      gA[n] = __ldg(g + n * kBankStride);
      gB[n] = __ldg(g + (NL + n) * kBankStride);
    }
  }
  const bool st2 = (reinterpret_cast<uintptr_t>(yu) & 7) == 0;  // yu + ja is then 8-byte aligned (ja is even)
  if (bulk) mbar_wait(&bar, 0);
  else __syncthreads();
  const unsigned xs_addr = (unsigned)__cvta_generic_to_shared(xs_raw + A0);
#pragma unroll 1
  for (int r = 0; r < RP; ++r) {
    // the whole window first (issued where written: the compiler otherwise serialises load -> 6 operations -> load
    // through one register pair, every step waiting out the shared-memory latency), then the arithmetic
    float2 v[NL];
    const unsigned pa0 = xs_addr + (unsigned)(r * step) * 4u;
#pragma unroll
    for (int n = 0; n < NL; ++n)
      asm volatile("ld.shared.v2.f32 {%0, %1}, [%2];" : "=f"(v[n].x), "=f"(v[n].y) : "r"(pa0 + 8u * n));
    float accA = 0.f, accB = 0.f;
#pragma unroll
    for (int n = 0; n < NL; ++n) {
      const float2 pa = mul2(v[n], gA[n]), pb = mul2(v[n], gB[n]);
      accA = __fadd_rn(__fadd_rn(accA, pa.x), pa.y);
      accB = __fadd_rn(__fadd_rn(accB, pb.x), pb.y);
    }
    const unsigned j = ja + (unsigned)r * 2u * TP;
    if (j + 1 < n_out) {
      if (st2) *reinterpret_cast<float2*>(yu + j) = make_float2(accA, accB);
      else {
        yu[j] = accA;
        yu[j + 1] = accB;
      }
    } else if (j < n_out) {
      yu[j] = accA;
    }
  }
}

}  // namespace ssr

using namespace ssr;

// SSR_FORCE_OLD_K3=1 routes everything through k_resample_tiled (A/B tests only)
static bool force_old_k3() {
  static int v = -1;
  if (v < 0) {
    const char* e = getenv("SSR_FORCE_OLD_K3");
    v = (e && e[0] == '1') ? 1 : 0;
  }
  return v == 1;
}

// SSR_FORCE_BULK_K3=1 skips k_resample_pair (A/B tests only)
static bool force_bulk_k3() {
  static int v = -1;
  if (v < 0) {
    const char* e = getenv("SSR_FORCE_BULK_K3");
    v = (e && e[0] == '1') ? 1 : 0;
  }
  return v == 1;
}

// the transposed output-order copies of a float bank and k_resample_pair's tables (see ssr_resample_plan)
static void free_staged_banks(ssr_resample_plan* p) {
  if (p->bank_t) cudaFree(p->bank_t);
  if (p->pair_g) cudaFree(p->pair_g);
  if (p->pair_thr) cudaFree(p->pair_thr);
  p->bank_t = nullptr;
  p->pair_g = nullptr;
  p->pair_thr = nullptr;
}

template <typename V>
static cudaError_t upload(V** dst, const std::vector<V>& v) {
  cudaError_t e = cudaMalloc(dst, v.size() * sizeof(V));
  if (e == cudaSuccess) e = cudaMemcpy(*dst, v.data(), v.size() * sizeof(V), cudaMemcpyHostToDevice);
  return e;
}

static cudaError_t make_staged_banks(ssr_resample_plan* p, const float* bank_host) {
  p->bank_t = nullptr;
  p->pair_g = nullptr;
  p->pair_thr = nullptr;
  p->pair_nl = p->pair_tp = p->pair_m = 0;
  if (p->up > kBankStride) return cudaSuccess;  // the staged kernels are not used for such plans
  // table construction: resample_tables.hpp (shared with the CPU emulation of k_resample_pair, tests/host_emul.cu)
  cudaError_t e = upload(&p->bank_t, k3_build_bank_t(p->up, p->down, p->K, p->half_len, bank_host));
  if (e != cudaSuccess) return e;
  K3PairTables pt;
  if (!k3_build_pair_tables(p->up, p->down, p->K, p->half_len, bank_host, &pt)) return cudaSuccess;
  static_assert(sizeof(float2) == 2 * sizeof(float) && sizeof(int2) == 2 * sizeof(int), "flattened pair layouts");
  e = cudaMalloc(&p->pair_g, pt.g.size() * sizeof(float));
  if (e == cudaSuccess) e = cudaMemcpy(p->pair_g, pt.g.data(), pt.g.size() * sizeof(float), cudaMemcpyHostToDevice);
  if (e == cudaSuccess) e = cudaMalloc(&p->pair_thr, pt.thr.size() * sizeof(int));
  if (e == cudaSuccess) e = cudaMemcpy(p->pair_thr, pt.thr.data(), pt.thr.size() * sizeof(int), cudaMemcpyHostToDevice);
  if (e != cudaSuccess) return e;
  p->pair_nl = pt.nl;
  p->pair_tp = pt.tp;
  p->pair_m = pt.m;
  return cudaSuccess;
}

template <typename T>
static int resample_plan_create(ssr_resample_plan** out, int up, int down, const T* taps_host, int n_taps) {
  if (!out) return fail(SSR_ERR_INVALID, "plan pointer is NULL");
  *out = nullptr;
  if (up < 1 || down < 1 || !taps_host || n_taps < 1 || (n_taps % 2) == 0)
    return fail(SSR_ERR_INVALID, "resample plan: up, down >= 1 and an odd number of taps required");
  ssr_resample_plan* p = new ssr_resample_plan();
  p->up = up;
  p->down = down;
  p->n_taps = n_taps;
  p->half_len = (n_taps - 1) / 2;
  p->n_pre_pad = down - p->half_len % down;
  p->n_pre_remove = (p->half_len + p->n_pre_pad) / down;
  p->K = (n_taps + up - 1) / up;
  p->is_f64 = sizeof(T) == 8;
  p->floor_len = 0;
  std::vector<T> bank((size_t)up * p->K, (T)0);
  for (int ph = 0; ph < up; ++ph)
    for (int k = 0; k < p->K; ++k) {
      long long q = ph + (long long)k * up;
      if (q < n_taps) bank[(size_t)ph * p->K + k] = taps_host[q];
    }
  p->bank = nullptr;
  cudaError_t e = cudaGetDevice(&p->device);
  if (e == cudaSuccess) e = cudaMalloc(&p->bank, bank.size() * sizeof(T));
  if (e == cudaSuccess)
    e = cudaMemcpy(p->bank, bank.data(), bank.size() * sizeof(T), cudaMemcpyHostToDevice);
  p->bank_t = nullptr;
  p->pair_g = nullptr;
  p->pair_thr = nullptr;
  p->pair_nl = p->pair_tp = p->pair_m = 0;
  if constexpr (sizeof(T) == 4) {
    if (e == cudaSuccess) e = make_staged_banks(p, reinterpret_cast<const float*>(bank.data()));
  }
  if (e != cudaSuccess) {
    free_staged_banks(p);
    if (p->bank) cudaFree(p->bank);
    delete p;
    return fail(SSR_ERR_CUDA, std::string("resample plan upload: ") + cudaGetErrorString(e));
  }
  *out = p;
  return SSR_OK;
}

extern "C" {

int ssr_resample_plan_create(ssr_resample_plan** out, int up, int down, const float* taps_host,
                             int n_taps) {
  return resample_plan_create<float>(out, up, down, taps_host, n_taps);
}

int ssr_resample_plan_create_f64(ssr_resample_plan** out, int up, int down, const double* taps_host,
                                 int n_taps) {
  return resample_plan_create<double>(out, up, down, taps_host, n_taps);
}

/* Explicit polyphase bank (see include/ssr_b200.h): output j sits at time j * down / up (in input samples), n = floor,
 * phase = (j * down) % up; tap k of that phase multiplies x[n + lead - k], k = 0 .. K-1. */
int ssr_resample_plan_create_bank(ssr_resample_plan** out, int up, int down, const float* bank_host, int K, int lead) {
  if (!out) return fail(SSR_ERR_INVALID, "plan pointer is NULL");
  *out = nullptr;
  if (up < 1 || down < 1 || !bank_host || K < 1 || lead < 0 || lead >= K)
    return fail(SSR_ERR_INVALID, "resample bank plan: up, down >= 1, K >= 1 and 0 <= lead < K required");
  ssr_resample_plan* p = new ssr_resample_plan();
  p->up = up;
  p->down = down;
  p->n_taps = K * up;
  p->K = K;
  p->half_len = lead * up;      // c = j * down + half_len: newest sample floor(j * down / up) + lead, phase (j * down) % up
  p->n_pre_remove = 0;
  p->n_pre_pad = -lead * up;    // the same c for the kernels that use (j + n_pre_remove) * down - n_pre_pad
  p->is_f64 = 0;
  p->floor_len = 1;
  p->bank = nullptr;
  cudaError_t e = cudaGetDevice(&p->device);
  if (e == cudaSuccess) e = cudaMalloc(&p->bank, sizeof(float) * (size_t)up * K);
  if (e == cudaSuccess) e = cudaMemcpy(p->bank, bank_host, sizeof(float) * (size_t)up * K, cudaMemcpyHostToDevice);
  p->bank_t = nullptr;
  p->pair_g = nullptr;
  p->pair_thr = nullptr;
  p->pair_nl = p->pair_tp = p->pair_m = 0;
  if (e == cudaSuccess) e = make_staged_banks(p, bank_host);
  if (e != cudaSuccess) {
    free_staged_banks(p);
    if (p->bank) cudaFree(p->bank);
    delete p;
    return fail(SSR_ERR_CUDA, std::string("resample bank plan upload: ") + cudaGetErrorString(e));
  }
  *out = p;
  return SSR_OK;
}

int ssr_resample_plan_destroy(ssr_resample_plan* plan) {
  if (!plan) return SSR_OK;
  if (plan->bank) cudaFree(plan->bank);
  free_staged_banks(plan);
  delete plan;
  return SSR_OK;
}

int64_t ssr_resample_out_len(const ssr_resample_plan* plan, int64_t n_in) {
  if (!plan) return -1;
  long long t = (long long)n_in * plan->up;
  if (plan->floor_len) return (int64_t)(t / plan->down);
  return (int64_t)(t / plan->down + (t % plan->down ? 1 : 0));
}

int ssr_resample_poly_batched_f64(const ssr_resample_plan* plan, const double* x_dev,
                                  const int64_t* in_offsets_host, const int64_t* in_offsets_dev,
                                  double* y_dev, const int64_t* out_offsets_host,
                                  const int64_t* out_offsets_dev, int n, void* stream) {
  if (!plan || !x_dev || !y_dev || !in_offsets_host || !in_offsets_dev || !out_offsets_host ||
      !out_offsets_dev || n < 1)
    return fail(SSR_ERR_INVALID, "ssr_resample_poly_batched_f64: bad argument");
  if (int rc = check_offsets(in_offsets_host, n, "ssr_resample_poly_batched_f64 (input)")) return rc;
  if (int rc = check_offsets(out_offsets_host, n, "ssr_resample_poly_batched_f64 (output)")) return rc;
  if (!plan->is_f64) return fail(SSR_ERR_INVALID, "float32 plan used with float64 data");
  long long max_out = 0;
  for (int u = 0; u < n; ++u) {
    long long n_in = in_offsets_host[u + 1] - in_offsets_host[u];
    long long n_out = out_offsets_host[u + 1] - out_offsets_host[u];
    if (n_out != ssr_resample_out_len(plan, n_in))
      return fail(SSR_ERR_INVALID, "output offsets do not match ssr_resample_out_len(n_in)");
    if (n_out > max_out) max_out = n_out;
  }
  if (max_out == 0) return SSR_OK;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  for (int u0 = 0; u0 < n; u0 += 32768) {
    int nu = n - u0 < 32768 ? n - u0 : 32768;
    dim3 grid((unsigned)((max_out + 255) / 256), nu);
    k_resample<double><<<grid, 256, 0, st>>>(x_dev, reinterpret_cast<const long long*>(in_offsets_dev), y_dev,
                                             reinterpret_cast<const long long*>(out_offsets_dev), u0, plan->up,
                                             plan->down, plan->n_pre_pad, plan->n_pre_remove, plan->K,
                                             static_cast<const double*>(plan->bank));
    SSR_LAUNCH_CHECK("k_resample<double>");
  }
  return SSR_OK;
}

int ssr_resample_poly_batched(const ssr_resample_plan* plan, const float* x_dev,
                              const int64_t* in_offsets_host, const int64_t* in_offsets_dev,
                              float* y_dev, const int64_t* out_offsets_host,
                              const int64_t* out_offsets_dev, int n, void* stream) {
  if (plan && plan->is_f64) return fail(SSR_ERR_INVALID, "float64 plan used with float32 data");
  if (!plan || !x_dev || !y_dev || !in_offsets_host || !in_offsets_dev || !out_offsets_host ||
      !out_offsets_dev || n < 1)
    return fail(SSR_ERR_INVALID, "ssr_resample_poly_batched: bad argument");
  if (int rc = check_offsets(in_offsets_host, n, "ssr_resample_poly_batched (input)")) return rc;
  if (int rc = check_offsets(out_offsets_host, n, "ssr_resample_poly_batched (output)")) return rc;
  long long max_out = 0;
  for (int u = 0; u < n; ++u) {
    long long n_in = in_offsets_host[u + 1] - in_offsets_host[u];
    long long n_out = out_offsets_host[u + 1] - out_offsets_host[u];
    if (n_out != ssr_resample_out_len(plan, n_in))
      return fail(SSR_ERR_INVALID, "output offsets do not match ssr_resample_out_len(n_in)");
    if (n_out > max_out) max_out = n_out;
  }
  if (max_out == 0) return SSR_OK;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  long long max_in = 0;
  for (int u = 0; u < n; ++u) max_in = std::max<long long>(max_in, in_offsets_host[u + 1] - in_offsets_host[u]);
  {
    // production path for the sample-rate pairs of the evaluation (K = 21 / 22): two consecutive outputs per thread
    // (block size and tables: make_staged_banks)
    constexpr int RP = SSR_K3_RP;
    const int NL = plan->pair_nl;  // 64-bit loads covering K + d_max samples at either alignment
    const int TPP = plan->pair_tp;
    const long long outs = 2LL * TPP * RP;
    const long long spanp = k3_pair_span(plan->up, plan->down, plan->K, NL, TPP, RP);
    const long long c_max = (max_out + outs) * plan->down + plan->half_len;
    if (NL > 0 && plan->pair_g && plan->pair_thr && (spanp + 8) * (long long)sizeof(float) <= 64 * 1024 &&
        c_max < 0x7fffffffLL && max_in < 0x7fffffffLL && !force_old_k3() && !force_bulk_k3()) {
      const int span = (int)spanp;
      const size_t smem = sizeof(float) * (size_t)(span + 8);
      const long long x_total = in_offsets_host[n];
      const unsigned ib0 = (unsigned)(plan->half_len / plan->up);
      const unsigned ib_step = (unsigned)(plan->pair_m * RP * plan->down);  // (2 TP RP / up) * down
      const int step = plan->pair_m * plan->down;
      for (int u0 = 0; u0 < n; u0 += 32768) {
        int nu = n - u0 < 32768 ? n - u0 : 32768;
        dim3 grid((unsigned)((max_out + outs - 1) / outs), nu);
        const long long* io = reinterpret_cast<const long long*>(in_offsets_dev);
        const long long* oo = reinterpret_cast<const long long*>(out_offsets_dev);
#define SSR_K3P_LAUNCH(NLV, MAXT, MINB)                                                                          \
  do {                                                                                                           \
    auto kern = k_resample_pair<NLV, RP, MAXT, MINB>;                                                                      \
    SSR_CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));            \
    kern<<<grid, TPP, smem, st>>>(x_dev, io, y_dev, oo, u0, plan->K, plan->pair_g, plan->pair_thr, ib0, ib_step, \
                                  step, span, x_total);                                                          \
  } while (0)
        // register caps: the largest that keep MINB CTAs of MAXT threads on an SM (64 / 72 / 80 registers)
        if (NL == 12 && TPP <= 160) SSR_K3P_LAUNCH(12, 160, 5);
        else if (NL == 12 && TPP <= 320) SSR_K3P_LAUNCH(12, 320, 3);  // 44.1k -> 48k (160/147)
        else if (NL == 12) SSR_K3P_LAUNCH(12, 448, 2);                // 16k -> 44.1k (441/160), 8k / 12k / 24k -> 44.1k
        else if (TPP <= 160) SSR_K3P_LAUNCH(13, 160, 5);              // 48k -> 44.1k (147/160)
        else if (TPP <= 320) SSR_K3P_LAUNCH(13, 320, 2);              // 48k -> 44.1k (147/160)
        else SSR_K3P_LAUNCH(13, 448, 2);
#undef SSR_K3P_LAUNCH
        SSR_LAUNCH_CHECK("k_resample_pair");
      }
      return SSR_OK;
    }
  }
  // tiled kernels: block = smallest multiple of `up` that is >= 256 threads (<= 512)
  int TP = plan->up * ((256 + plan->up - 1) / plan->up);
  {
    // production path: TMA-staged spans, 32-bit index arithmetic, 16 outputs per thread
    constexpr int R2 = 16;
    const long long span2 = ((long long)TP * R2 - 1) * plan->down / plan->up + 2 + plan->K;
    const long long c_max = (max_out + (long long)TP * R2) * plan->down + plan->half_len;
    if (plan->bank_t && TP <= 512 && plan->K <= 48 && (span2 + 8) * (long long)sizeof(float) <= 64 * 1024 && c_max < 0x7fffffffLL &&
        max_in < 0x7fffffffLL && !force_old_k3()) {
      const int span = (int)span2;
      const size_t smem = sizeof(float) * (size_t)(span + 8);
      const long long x_total = in_offsets_host[n];
      for (int u0 = 0; u0 < n; u0 += 32768) {
        int nu = n - u0 < 32768 ? n - u0 : 32768;
        dim3 grid((unsigned)((max_out + (long long)TP * R2 - 1) / ((long long)TP * R2)), nu);
        const long long* io = reinterpret_cast<const long long*>(in_offsets_dev);
        const long long* oo = reinterpret_cast<const long long*>(out_offsets_dev);
#define SSR_K3B_LAUNCH(KM, EX)                                                                                   \
  do {                                                                                                           \
    auto kern = k_resample_bulk<KM, R2, EX>;                                                                     \
    SSR_CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));            \
    kern<<<grid, TP, smem, st>>>(x_dev, io, y_dev, oo, u0, (unsigned)plan->up, (unsigned)plan->down,            \
                                 (unsigned)plan->half_len, plan->K, plan->bank_t, span, x_total);                \
  } while (0)
        if (plan->K == 21) SSR_K3B_LAUNCH(21, true);       // 44.1k <-> 48k up (160/147), 16k -> 44.1k (441/160)
        else if (plan->K == 22) SSR_K3B_LAUNCH(22, true);  // 48k -> 44.1k (147/160)
        else if (plan->K <= 24) SSR_K3B_LAUNCH(24, false);
        else if (plan->K <= 32) SSR_K3B_LAUNCH(32, false);
        else if (plan->K <= 40) SSR_K3B_LAUNCH(40, false);
        else SSR_K3B_LAUNCH(48, false);
#undef SSR_K3B_LAUNCH
        SSR_LAUNCH_CHECK("k_resample_bulk");
      }
      return SSR_OK;
    }
  }
  constexpr int R = 8;
  // input samples one CTA touches: newest(j_last) - newest(j_first) + K <= (TP*R - 1) * down / up + 1 + K
  const long long span_ll = ((long long)TP * R - 1) * plan->down / plan->up + 2 + plan->K;
  if (TP <= 512 && plan->K <= 48 && span_ll * (long long)sizeof(float) <= 48 * 1024) {
    const int span = (int)span_ll;
    const size_t smem = sizeof(float) * (size_t)span;
    for (int u0 = 0; u0 < n; u0 += 32768) {
      int nu = n - u0 < 32768 ? n - u0 : 32768;
      dim3 grid((unsigned)((max_out + (long long)TP * R - 1) / ((long long)TP * R)), nu);
      const long long* io = reinterpret_cast<const long long*>(in_offsets_dev);
      const long long* oo = reinterpret_cast<const long long*>(out_offsets_dev);
#define SSR_K3_LAUNCH(KM, EX)                                                                              \
  k_resample_tiled<KM, R, EX><<<grid, TP, smem, st>>>(x_dev, io, y_dev, oo, u0, plan->up, plan->down,     \
                                                      plan->n_pre_pad, plan->n_pre_remove, plan->K,       \
                                                      static_cast<const float*>(plan->bank), span)
      if (plan->K == 21) SSR_K3_LAUNCH(21, true);       // 44.1k <-> 48k up (160/147), 16k -> 44.1k (441/160)
      else if (plan->K == 22) SSR_K3_LAUNCH(22, true);  // 48k -> 44.1k (147/160)
      else if (plan->K <= 24) SSR_K3_LAUNCH(24, false);
      else if (plan->K <= 32) SSR_K3_LAUNCH(32, false);
      else if (plan->K <= 40) SSR_K3_LAUNCH(40, false);
      else SSR_K3_LAUNCH(48, false);
#undef SSR_K3_LAUNCH
      SSR_LAUNCH_CHECK("k_resample_tiled");
    }
    return SSR_OK;
  }
  for (int u0 = 0; u0 < n; u0 += 32768) {
    int nu = n - u0 < 32768 ? n - u0 : 32768;
    dim3 grid((unsigned)((max_out + 255) / 256), nu);
    k_resample<float><<<grid, 256, 0, st>>>(x_dev, reinterpret_cast<const long long*>(in_offsets_dev), y_dev,
                                            reinterpret_cast<const long long*>(out_offsets_dev), u0, plan->up,
                                            plan->down, plan->n_pre_pad, plan->n_pre_remove, plan->K,
                                            static_cast<const float*>(plan->bank));
    SSR_LAUNCH_CHECK("k_resample");
  }
  return SSR_OK;
}

}  // extern "C"
