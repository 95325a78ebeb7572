// k1_common.cuh -- device helpers, constants and plan-side structs shared by the K1 / K2 kernels.
#pragma once
#include <stdint.h>

#include "../../include/ssr_b200.h"
#include "fft_core.cuh"

namespace ssr {

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
// Spectrogram layouts the K1 kernels write.  Plain: one (T, F) float image per signal, row pitch F (the magnitude
// entry point).  Interleaved (K1 -> K2, spec_t == spec_e + 1): ONE image of (estimate, target) float2 pairs with an
// even row pitch, so that every row starts on a 16-byte boundary and K2 streams it with 16-byte copies straight into
// the (x, y)-pair layout its packed arithmetic works on.
// ---------------------------------------------------------------------------------------------
__host__ __device__ inline int spec_pitch_pairs(int F) { return (F + 1) & ~1; }  // float2 per interleaved row
struct SpecLayout {
  int step;         // floats between consecutive bins of one signal (1 plain, 2 interleaved)
  long long pitch;  // floats per row
};
__device__ __forceinline__ SpecLayout spec_layout(const float* spec_e, const float* spec_t, int F) {
  const bool inter = spec_e != nullptr && spec_t == spec_e + 1;
  SpecLayout l;
  l.step = inter ? 2 : 1;
  l.pitch = inter ? 2 * spec_pitch_pairs(F) : F;
  return l;
}

// magnitudes of bin k of one frame -> the spectrogram row(s); st / se = row starts (null: not wanted)
__device__ __forceinline__ void spec_store(const SpecLayout& l, float* st, float* se, int k, float mt, float me) {
  if (l.step == 2) {
    reinterpret_cast<float2*>(se)[k] = make_float2(me, mt);  // rows are 8-byte aligned (even pitch, even offsets)
  } else {
    if (st) st[k] = mt;
    if (se) se[k] = me;
  }
}

// ---------------------------------------------------------------------------------------------
// setup: work-item table.  item_start[p] = first work item of pair p (item = chunk of <= `chunk`
// consecutive frames), item_pair[item] = p, spec_off[p] = first spectrogram element (float) of pair p: frames before
// it x spec_pitch, the floats per spectrogram row (SpecLayout below).
// ---------------------------------------------------------------------------------------------
__global__ void k_setup(const long long* __restrict__ offsets, int n, int n_fft, int hop, int chunk,
                        int spec_pitch, int* __restrict__ item_start, int* __restrict__ item_pair,
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
    spec_off[p] = fr * spec_pitch;
    for (int c = 0; c < nc; ++c) item_pair[it + c] = p;
    it += nc;
    fr += T;
  }
  if (hi == n) {
    item_start[n] = (int)it;
    item_start[n + 1] = 0;  // work-item counter of the persistent kernels (dynamic scheduling)
    spec_off[n] = fr * spec_pitch;
  }
}

// Persistent CTAs draw their work items from a counter instead of striding by gridDim.x: items differ in
// length (the last chunk of a pair is short) and a static stride resonates with the items-per-pair period
// (592 CTAs, 8 items per 5 s pair: every 8th CTA would only ever see the short items).  Which CTA computes an
// item does not change its partial sums, so results stay bit-identical.
__device__ __forceinline__ int next_work_item(int* counter, int* slot_smem) {
  if (threadIdx.x == 0) *slot_smem = atomicAdd(counter, 1);
  __syncthreads();
  const int item = *slot_smem;
  __syncthreads();  // the slot may be rewritten by the next call
  return item;
}


}  // namespace ssr
