// k1_2048s.cuh -- K1, n_fft = 2048, phase-staggered form of the radix 16x16x8 kernel (k1_2048.cuh).
//
// Same arithmetic, same index maps, bit-identical results.  What changes is the schedule: one CTA of
// kSG x 128 threads per SM; each 128-thread group owns its own FFT buffer and walks its own frames
// through the three passes, but the groups are offset by one pass and advance in lock step (ONE
// CTA-wide barrier per pass).  At any time one group is in pass 1 (global loads, F2F conversions,
// FP64 butterflies), one in pass 2 (shared-memory exchanges + FP64 butterflies) and one in pass 3 +
// epilogue (F2F / MUFU / float32 work), so the FP64, LSU and XU pipes of every SM sub-partition are
// fed concurrently instead of the co-resident CTAs of k1_2048.cuh drifting into the same pass.
#pragma once
#include "k1_common.cuh"
#include "k1_map.cuh"

namespace ssr {

constexpr int kSG = 3;  // groups (frames in flight) per CTA

__device__ __forceinline__ void group_sync(int g) {
  asm volatile("bar.sync %0, 128;" ::"r"(1 + g) : "memory");
}

template <int FIXED>
__global__ void __launch_bounds__(kV2Threads* kSG, 1)
k_stft_metrics_2048s(StftDev P, const float* __restrict__ est, const float* __restrict__ tgt,
                     const long long* __restrict__ offsets, const int* __restrict__ item_start,
                     const int* __restrict__ item_pair, int n_items, int chunk, unsigned flags,
                     double* __restrict__ partials, float* __restrict__ spec_e,
                     float* __restrict__ spec_t, const long long* __restrict__ spec_off) {
  constexpr int N = 2048, F = 1025, NW = kV2Threads / 32;
  constexpr int kBufBytes = sizeof(cd) * (N + N / 8);
  extern __shared__ __align__(16) unsigned char smem_raw[];
  __shared__ __align__(16) cd tw2[15 * 8];
  __shared__ float lsd_part[kSG][kMaxChunk][NW];
  __shared__ double red[kSG][NW][kPartials];

  const int g = threadIdx.x >> 7;  // group
  const int tid = threadIdx.x & 127, lane = tid & 31, warp = tid >> 5;
  cd* const buf = reinterpret_cast<cd*>(smem_raw + (size_t)g * kBufBytes);
  float2* const edge_raw = reinterpret_cast<float2*>(buf);  // edge frames stage N raw pairs inside buf
  // magnitude rows (store mode only): after the kSG FFT buffers
  float* const row_t = reinterpret_cast<float*>(smem_raw + (size_t)kSG * kBufBytes) + (size_t)g * 2 * 1104;
  float* const row_e = row_t + 1104;

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
  if (threadIdx.x < 120) tw2[threadIdx.x] = P.tw[16 * (threadIdx.x & 7) * ((threadIdx.x >> 3) + 1)];
  int ia, ib;
  v2_thread_butterflies(tid, &ia, &ib);
  const bool special = (tid == kV2Threads - 1);
  const int ka = v2_klow(ia), kb = v2_klow(ib);
  const int j2 = tid & 7;
  cd* const b1 = buf + pad_idx(tid);
  cd* const b2 = buf + pad_idx((tid >> 3) * 128 + j2);
  const cd* const b3a = buf + 9 * ia;
  const cd* const b3b = buf + 9 * ib;
  const cd* const t2 = tw2 + j2;

  // ---- per-group work state (uniform inside a group)
  const int vstride = gridDim.x * kSG;
  int item = blockIdx.x * kSG + g;
  bool done = item >= n_items;
  int phase = 0, delay = g;
  int pend_item = -1, pend_nf = 0;  // item whose partial sums still have to be reduced and stored
  int fi = 0, nf = 0, p = 0;
  long long L = 0, f0 = 0;
  const float* xe = est;
  const float* xt = tgt;
  double s_et = 0, s_tt = 0, s_ee = 0, l_et = 0, l_tt = 0, l_ee = 0;
  float* pend_t = nullptr;
  float* pend_e = nullptr;

  auto open_item = [&]() {
    p = item_pair[item];
    const int c = item - item_start[p];
    const long long off = offsets[p];
    L = offsets[p + 1] - off;
    const long long T = stft_frames(L, N, hop);
    f0 = (long long)c * chunk;
    nf = (int)min((long long)chunk, T - f0);
    xe = est + off;
    xt = tgt + off;
    fi = 0;
  };
  if (!done) open_item();
  __syncthreads();

  while (true) {
    // ---- work left over from this group's previous step (the CTA barrier has made it visible)
    if (pend_t) {  // coalesced copy-out of the previous frame's magnitude rows
      for (int k = tid; k < F; k += kV2Threads) {
        pend_t[k] = row_t[k + (k >> 4)];
        if (pend_e) pend_e[k] = row_e[k + (k >> 4)];
      }
      pend_t = nullptr;
    }
    if (pend_item >= 0) {  // per-item reduction -> partials[item][0..6]
      double lsd_sum = 0.0;
      if (want_lsd && tid < pend_nf) {
        float sacc = 0.f;
#pragma unroll
        for (int w = 0; w < NW; ++w) sacc += lsd_part[g][tid][w];
        lsd_sum = (double)sqrtf(sacc / (float)F);  // torch.mean(dim=3) ** 0.5 in float32
      }
      double vals[7] = {lsd_sum, s_et, s_tt, s_ee, l_et, l_tt, l_ee};
#pragma unroll
      for (int i = 0; i < 7; ++i) {
        const double r = warp_sum(vals[i]);
        if (lane == 0) red[g][warp][i] = r;
      }
      group_sync(g);
      if (tid < 7) {
        double r = 0.0;
#pragma unroll
        for (int w = 0; w < NW; ++w) r += red[g][w][tid];
        partials[(size_t)pend_item * kPartials + tid] = r;
      }
      s_et = s_tt = s_ee = l_et = l_tt = l_ee = 0.0;
      pend_item = -1;
      if (item >= n_items) done = true;
    }

    if (delay > 0) {
      --delay;
    } else if (!done) {
      if (phase == 0) {
        // ================= pass 1: load + window, radix-16, twiddle, store
        const long long f = f0 + fi;
        const long long start = f * hop - N / 2;
        cd v[16];
        const bool interior = (start >= 0 && start + N <= L);
        if (interior) {
          const float* pt = xt + start + tid;
          const float* pe = xe + start + tid;
#pragma unroll
          for (int r = 0; r < 16; ++r) {
            const double w = __ldg(P.win_half + tid + 128 * r);
            v[r] = cd{w * (double)__ldg(pt + 128 * r), w * (double)__ldg(pe + 128 * r)};
          }
          if (tid < 32) {
            const long long nxt = start + N + (long long)(tid & 15) * 32;
            if (nxt < L && (tid & 15) * 32 < hop) prefetch_l1((tid < 16 ? xt : xe) + nxt);
          }
        } else {
          // edge frame (reflect padding): gather through a staging area inside this group's buffer
#pragma unroll 1
          for (int n = tid; n < N; n += kV2Threads) {
            const long long idx = reflect_index(start + n, L);
            edge_raw[n] = make_float2(__ldg(xt + idx), __ldg(xe + idx));
          }
          group_sync(g);
#pragma unroll
          for (int r = 0; r < 16; ++r) {
            const double w = __ldg(P.win_half + tid + 128 * r);
            const float2 x = edge_raw[tid + 128 * r];
            v[r] = cd{w * (double)x.x, w * (double)x.y};
          }
          group_sync(g);  // staging reads done before the buffer is overwritten below
        }
        bfly16<false>(v);
#pragma unroll
        for (int q = 1; q < 16; ++q) v[q] = cmul(v[q], tw1[q - 1]);
#pragma unroll
        for (int q = 0; q < 16; ++q) b1[144 * q] = v[q];
        phase = 1;
      } else if (phase == 1) {
        // ================= pass 2: sub-transforms of length 128 (stride 8), in place per thread
        cd v[16];
#pragma unroll
        for (int r = 0; r < 16; ++r) v[r] = b2[9 * r];
        bfly16<false>(v);
        b2[0] = v[0];
#pragma unroll
        for (int q = 1; q < 16; ++q) b2[9 * q] = cmul(v[q], t2[(q - 1) * 8]);
        phase = 2;
      } else {
        // ================= pass 3: two radix-8 butterflies (a and its Hermitian partner b) + epilogue
        cd v[16];
        cd* a = v;
        cd* b = v + 8;
#pragma unroll
        for (int r = 0; r < 8; ++r) {
          a[r] = b3a[r];
          b[r] = b3b[r];
        }
        bfly8<false>(a);
        bfly8<false>(b);
        const long long f = f0 + fi;
        float lsd_acc = 0.f;
        float* st = spec_t ? spec_t + spec_off[p] + f * F : nullptr;
        float* se = spec_e ? spec_e + spec_off[p] + f * F : nullptr;
        auto emit = [&](int k, cd zk, cd zn) {
          // see k1_2048.cuh: un-packing, complex64 rounding, float32 metric terms
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
          if (st) {
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
          if (lane == 0) lsd_part[g][fi][warp] = w;
        }
        pend_t = st;  // copied out at the start of this group's next step (after the CTA barrier)
        pend_e = se;
        phase = 0;
        if (++fi == nf) {  // item finished: its reduction runs at the start of the next step
          pend_item = item;
          pend_nf = nf;
          item += vstride;
          if (item < n_items) open_item();
        }
      }
    }
    // one CTA-wide barrier per pass; the vote ends the loop once every group has drained its items
    if (__syncthreads_and(done && pend_item < 0 && pend_t == nullptr)) break;
  }
}

}  // namespace ssr
