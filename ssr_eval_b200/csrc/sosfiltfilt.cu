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
// The recursion is sequential in time but the sections form a systolic pipeline: one WARP per
// utterance, lane s owns section s and at micro-step t filters sample t - s, taking its input from
// lane s-1 by shuffle.  Inputs are fetched and outputs stored 32 samples at a time (coalesced).
// Thousands of utterances (one per warp) run concurrently, which is where the throughput comes from.
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
__device__ __forceinline__ double ext_sample(const float* __restrict__ x, long long L, int edge, long long i) {
  if (i < edge) return (double)__fsub_rn(__fmul_rn(2.0f, x[0]), x[edge - i]);
  const long long m = i - edge;
  if (m < L) return (double)x[m];
  return (double)__fsub_rn(__fmul_rn(2.0f, x[L - 1]), x[L - 2 - (m - L)]);
}

__global__ void __launch_bounds__(128)
k_sosfiltfilt(SosDev P, const float* __restrict__ x, const long long* __restrict__ offsets, int n,
              double* __restrict__ y, double* __restrict__ ws) {
  const int lane = threadIdx.x & 31;
  const int warp_global = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int n_warps = (gridDim.x * blockDim.x) >> 5;
  const int S = P.n_sections, edge = P.edge;
  const bool owner = lane < S;
  const int sl = owner ? lane : 0;
  const double b0 = P.b0[sl], b1 = P.b1[sl], b2 = P.b2[sl], a1 = P.a1[sl], a2 = P.a2[sl];
  const double zi0 = P.zi0[sl], zi1 = P.zi1[sl];
  const unsigned full = 0xffffffffu;

  for (int u = warp_global; u < n; u += n_warps) {
    const long long off = offsets[u];
    const long long L = offsets[u + 1] - off;
    const long long n_tot = L + 2LL * edge;
    const float* xu = x + off;
    double* w = ws + off + 2LL * edge * u;  // forward-pass output, n_tot doubles
    double* yu = y + off;

    for (int pass = 0; pass < 2; ++pass) {
      // initial conditions: zi * (first input sample of this pass)
      const double first = pass == 0 ? ext_sample(xu, L, edge, 0) : w[n_tot - 1];
      double z0 = __dmul_rn(zi0, first), z1 = __dmul_rn(zi1, first);
      double carry = 0.0;   // x_new of the previous micro-step (input of lane s+1)
      double yout = 0.0;    // output staging: lane (index & 31) keeps output `index`
      const long long steps = n_tot + S - 1;
      for (long long base = 0; base < steps; base += 32) {
        // coalesced fetch of inputs base .. base+31
        const long long ii = base + lane;
        double xin = 0.0;
        if (ii < n_tot) xin = pass == 0 ? ext_sample(xu, L, edge, ii) : w[n_tot - 1 - ii];
#pragma unroll 4
        for (int j = 0; j < 32; ++j) {
          const long long t = base + j;
          if (t >= steps) break;
          const double from_prev = __shfl_up_sync(full, carry, 1);
          const double from_mem = __shfl_sync(full, xin, j);
          const double x_cur = lane == 0 ? from_mem : from_prev;
          const long long idx = t - lane;  // sample this lane filters now
          double x_new = 0.0;
          if (owner && idx >= 0 && idx < n_tot) {
            x_new = __dadd_rn(__dmul_rn(b0, x_cur), z0);
            z0 = __dadd_rn(__dsub_rn(__dmul_rn(b1, x_cur), __dmul_rn(a1, x_new)), z1);
            z1 = __dsub_rn(__dmul_rn(b2, x_cur), __dmul_rn(a2, x_new));
          }
          carry = x_new;
          // the last section's output is sample o = t - (S - 1)
          const double done = __shfl_sync(full, x_new, S - 1);
          const long long o = t - (S - 1);
          if (o >= 0) {
            if (lane == (int)(o & 31)) yout = done;
            if ((o & 31) == 31 || o == n_tot - 1) {  // flush the staged block (coalesced)
              const long long blk = o & ~31LL;
              const long long oi = blk + lane;
              if (oi <= o) {
                if (pass == 0) {
                  w[oi] = yout;
                } else {
                  const long long m = n_tot - 1 - oi - edge;  // reverse + strip the padding
                  if (m >= 0 && m < L) yu[m] = yout;
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
  long long total = 0;
  for (int u = 0; u < n; ++u) {
    long long L = offsets_host[u + 1] - offsets_host[u];
    if (L <= edge) return fail(SSR_ERR_INVALID, "The length of the input vector x must be greater than padlen");
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
  int dev = 0, sms = 148;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  int warps_needed = n;
  int blocks = (warps_needed + 3) / 4;
  if (blocks > sms * 16) blocks = sms * 16;
  k_sosfiltfilt<<<blocks, 128, 0, static_cast<cudaStream_t>(stream)>>>(
      P, x_dev, reinterpret_cast<const long long*>(offsets_dev), n, y_dev, static_cast<double*>(workspace_dev));
  SSR_LAUNCH_CHECK("k_sosfiltfilt");
  return SSR_OK;
}

size_t ssr_sosfiltfilt_workspace_bytes(const int64_t* offsets_host, int n, int edge) {
  if (!offsets_host || n < 1 || edge < 0) return 0;
  return sizeof(double) * (size_t)((offsets_host[n] - offsets_host[0]) + 2LL * edge * n);
}

}  // extern "C"
