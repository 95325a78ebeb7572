// sosfiltfilt.cu -- K7: zero-phase IIR filtering = scipy.signal.sosfiltfilt(sos, x) (SURVEY.md section 8f
// rank 3), the arithmetic behind lowpass_filter / bandpass_filter (ssr_eval/lowpass.py:54-131; callers
// ssr_eval/eval.py:334-399).  Filter DESIGN (butter / cheby1 / ellip / bessel -> sos, sosfilt_zi,
// padlen) stays on the host with scipy, exactly like the reference; this kernel replaces the recursion:
//   ext = odd_ext(x, edge)                         (float32 arithmetic, as numpy does it)
//   y1  = sosfilt(sos, ext, zi = zi * ext[0])      (float64)
//   y2  = sosfilt(sos, reverse(y1), zi = zi * y1[-1])
//   out = reverse(y2)[edge : -edge]                (float64)
// scipy's inner loop (signal/_sosfilt.pyx), per sample and section:
//   x_new = b0 * x_cur + z0;  z0 = (b1 * x_cur - a1 * x_new) + z1;  z1 = b2 * x_cur - a2 * x_new
// is reproduced with separately rounded multiplies and adds in the same order.
//
// The recursion is sequential in time but the sections form a systolic pipeline: a GROUP of SP lanes
// (SP = smallest power of two >= the number of sections: 1, 2, 4 or 8 for filter orders 2 .. 10) owns one
// utterance, lane s of the group owns section s and at micro-step t filters sample t - s, taking its input
// from lane s-1 by a width-SP shuffle.  A warp therefore filters 32 / SP utterances at once (the first
// version gave every utterance a whole warp and left 32 - S lanes idle: 31 k utt/s at order 8).  Inputs are
// fetched and outputs stored SP samples at a time per group, the inputs eight blocks ahead (register ring).
// Thousands of utterances run concurrently, which is where the throughput comes from.
#include <math.h>

#include <vector>

#include "common.cuh"

namespace ssr {

constexpr int kSosMaxSections = 32;

struct SosDev {
  int n_sections, edge;
  double b0[kSosMaxSections], b1[kSosMaxSections], b2[kSosMaxSections];
  double a1[kSosMaxSections], a2[kSosMaxSections];
  double zi0[kSosMaxSections], zi1[kSosMaxSections];
};

// odd extension in float32 arithmetic: 2*x[0] - x[edge - i] | x | 2*x[L-1] - x[L - 2 - j]
__device__ __forceinline__ float ext_sample_f32(const float* __restrict__ x, long long L, int edge, long long i) {
  if (i < edge) return __fsub_rn(__fmul_rn(2.0f, x[0]), x[edge - i]);
  const long long m = i - edge;
  if (m < L) return x[m];
  return __fsub_rn(__fmul_rn(2.0f, x[L - 1]), x[L - 2 - (m - L)]);
}
__device__ __forceinline__ double ext_sample(const float* __restrict__ x, long long L, int edge, long long i) {
  return (double)ext_sample_f32(x, L, edge, i);
}

template <int SP>
__global__ void __launch_bounds__(128)
k_sosfiltfilt(SosDev P, const float* __restrict__ x, const long long* __restrict__ offsets, int n,
              double* __restrict__ y, double* __restrict__ ws) {
  constexpr int G = 32 / SP;  // utterances per warp
  const int lane = threadIdx.x & 31;
  const int s = lane & (SP - 1), g = lane / SP;
  const int warp_global = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int n_warps = (gridDim.x * blockDim.x) >> 5;
  const int S = P.n_sections, edge = P.edge;
  const bool owner = s < S;
  const int sl = owner ? s : 0;
  const double b0 = P.b0[sl], b1 = P.b1[sl], b2 = P.b2[sl], a1 = P.a1[sl], a2 = P.a2[sl];
  const double zi0 = P.zi0[sl], zi1 = P.zi1[sl];
  const unsigned full = 0xffffffffu;

  for (long long u0 = (long long)warp_global * G; u0 < n; u0 += (long long)n_warps * G) {
    const long long u = u0 + g;
    const bool live = u < n;  // this group has an utterance
    const long long off = live ? offsets[u] : 0;
    const long long L = live ? offsets[u + 1] - off : 0;
    const long long n_tot = live ? L + 2LL * edge : 0;
    const float* xu = x + off;
    double* w = ws + off + 2LL * edge * u;  // forward-pass output, n_tot doubles
    double* yu = y + off;
    const long long steps = live ? n_tot + S - 1 : 0;
    long long steps_max = steps;  // the warp runs until its longest utterance is done
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) steps_max = max(steps_max, __shfl_xor_sync(full, steps_max, o));

    const int nt = (int)n_tot, Li = (int)L;  // per-utterance sizes fit 32 bits (checked on the host)
    for (int pass = 0; pass < 2; ++pass) {
      // initial conditions: zi * (first input sample of this pass)
      double first = 0.0;
      if (live) first = pass == 0 ? ext_sample(xu, L, edge, 0) : w[n_tot - 1];
      double z0 = __dmul_rn(zi0, first), z1 = __dmul_rn(zi1, first);
      double carry = 0.0;  // x_new of the previous micro-step (input of lane s+1)
      // a fetched value stays as it was loaded (float32 in the forward pass, float64 in the backward pass) until it is
      // used: converting at fetch time would make the fetch itself wait for the load
      auto fetch = [&](int base, float* vf, double* vd) {
        const int ii = base + s;
        *vf = 0.f;
        *vd = 0.0;
        if (ii < nt) {
          if (pass == 0) *vf = ext_sample_f32(xu, L, edge, ii);
          else *vd = w[nt - 1 - ii];
        }
      };
      // Inputs are requested PD blocks (PD * SP samples) ahead through a register ring: a group's loads hit a cache
      // line of its own utterance (nothing coalesces across groups), and one block ahead left the float -> double
      // conversion of the loaded sample waiting on the long scoreboard for 45 % of all warp time (ncu source page).
      constexpr int PD = 8;
      float xqf[PD];
      double xqd[PD];
#pragma unroll
      for (int b = 0; b < PD; ++b) fetch(b * SP, &xqf[b], &xqd[b]);
      const int smax = (int)steps_max;
      for (int base0 = 0; base0 < smax; base0 += SP * PD) {
#pragma unroll
        for (int b = 0; b < PD; ++b) {
          const int base = base0 + b * SP;
          if (base < smax) {  // uniform across the warp
            const double xin = pass == 0 ? (double)xqf[b] : xqd[b];
            fetch(base + PD * SP, &xqf[b], &xqd[b]);
#pragma unroll
            for (int j = 0; j < SP; ++j) {
              const int t = base + j;
              const double from_prev = __shfl_up_sync(full, carry, 1, SP);
              const double from_mem = __shfl_sync(full, xin, j, SP);
              const double x_cur = s == 0 ? from_mem : from_prev;
              const int idx = t - s;  // sample this lane filters now
              // branch-free step: every lane computes, inactive lanes keep their state and pass on 0
              const bool act = owner && (unsigned)idx < (unsigned)nt;
              const double xn = __dadd_rn(__dmul_rn(b0, x_cur), z0);
              const double z0n = __dadd_rn(__dsub_rn(__dmul_rn(b1, x_cur), __dmul_rn(a1, xn)), z1);
              const double z1n = __dsub_rn(__dmul_rn(b2, x_cur), __dmul_rn(a2, xn));
              z0 = act ? z0n : z0;
              z1 = act ? z1n : z1;
              carry = act ? xn : 0.0;
              if (act && s == S - 1) {  // the last section's lane stores output `idx` itself
                if (pass == 0) {
                  w[idx] = xn;
                } else {
                  const int m = nt - 1 - idx - edge;  // reverse + strip the padding
                  if ((unsigned)m < (unsigned)Li) yu[m] = xn;
                }
              }
            }
          }
        }
      }
      __syncwarp();
    }
  }
}

}  // namespace ssr

using namespace ssr;

extern "C" {

int ssr_sosfiltfilt_batched(const double* sos_host, int n_sections, const double* zi_host, int edge,
                            const float* x_dev, const int64_t* offsets_host, const int64_t* offsets_dev,
                            int n, double* y_dev, void* workspace_dev, size_t workspace_bytes, void* stream) {
  if (!sos_host || !zi_host || !x_dev || !offsets_host || !offsets_dev || !y_dev || n < 1)
    return fail(SSR_ERR_INVALID, "ssr_sosfiltfilt_batched: bad argument");
  if (n_sections < 1 || n_sections > kSosMaxSections) return fail(SSR_ERR_INVALID, "1..32 sections supported");
  if (edge < 0) return fail(SSR_ERR_INVALID, "edge must be >= 0");
  if (int rc = check_offsets(offsets_host, n, "ssr_sosfiltfilt_batched")) return rc;
  long long total = 0;
  for (int u = 0; u < n; ++u) {
    long long L = offsets_host[u + 1] - offsets_host[u];
    if (L <= edge) return fail(SSR_ERR_INVALID, "The length of the input vector x must be greater than padlen");
    if (L + 2LL * edge + 64 > 0x7fffffffLL) return fail(SSR_ERR_INVALID, "utterance too long");
    total += L + 2LL * edge;
  }
  if (!workspace_dev || workspace_bytes < sizeof(double) * (size_t)total)
    return fail(SSR_ERR_WORKSPACE, "workspace too small");
  SosDev P;
  P.n_sections = n_sections;
  P.edge = edge;
  for (int s = 0; s < n_sections; ++s) {
    const double* r = sos_host + 6 * s;  // b0 b1 b2 a0 a1 a2 with a0 == 1 (scipy normalises)
    P.b0[s] = r[0];
    P.b1[s] = r[1];
    P.b2[s] = r[2];
    P.a1[s] = r[4];
    P.a2[s] = r[5];
    P.zi0[s] = zi_host[2 * s];
    P.zi1[s] = zi_host[2 * s + 1];
  }
  const int sms = sm_count();
  int sp = 1;
  while (sp < n_sections) sp *= 2;  // lanes per utterance
  const int warps_needed = (int)(((long long)n * sp + 31) / 32);
  int blocks = (warps_needed + 3) / 4;
  if (blocks > sms * 16) blocks = sms * 16;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const long long* od = reinterpret_cast<const long long*>(offsets_dev);
  double* wsd = static_cast<double*>(workspace_dev);
  switch (sp) {
    case 1: k_sosfiltfilt<1><<<blocks, 128, 0, st>>>(P, x_dev, od, n, y_dev, wsd); break;
    case 2: k_sosfiltfilt<2><<<blocks, 128, 0, st>>>(P, x_dev, od, n, y_dev, wsd); break;
    case 4: k_sosfiltfilt<4><<<blocks, 128, 0, st>>>(P, x_dev, od, n, y_dev, wsd); break;
    case 8: k_sosfiltfilt<8><<<blocks, 128, 0, st>>>(P, x_dev, od, n, y_dev, wsd); break;
    case 16: k_sosfiltfilt<16><<<blocks, 128, 0, st>>>(P, x_dev, od, n, y_dev, wsd); break;
    default: k_sosfiltfilt<32><<<blocks, 128, 0, st>>>(P, x_dev, od, n, y_dev, wsd); break;
  }
  SSR_LAUNCH_CHECK("k_sosfiltfilt");
  return SSR_OK;
}

size_t ssr_sosfiltfilt_workspace_bytes(const int64_t* offsets_host, int n, int edge) {
  if (!offsets_host || n < 1 || edge < 0) return 0;
  return sizeof(double) * (size_t)((offsets_host[n] - offsets_host[0]) + 2LL * edge * n);
}

}  // extern "C"
