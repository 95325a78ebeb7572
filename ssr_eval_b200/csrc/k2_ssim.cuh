// k2_ssim.cuh -- K2 (7x7 box-window SSIM, cp.async row pipeline) and the fixed-order finalize kernel.
#pragma once
#include "f32x2.cuh"
#include "k1_common.cuh"

namespace ssr {

__device__ __forceinline__ float rcp_approx(float x) {
  float r;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
  return r;
}

// ---------------------------------------------------------------------------------------------
// K2: SSIM of two (T, F) float32 magnitude images, valid 7x7 windows only (skimage crops the
// 3-pixel border, so the reflect boundary mode of uniform_filter never reaches the mean).
// One CTA = one tile of kSsimTR x kSsimTC window positions, 128 threads, TWO adjacent columns per
// thread.  Rows stream through a 7-stage shared row buffer that holds, per column, the PAIR (estimate, target);
// per row a thread forms the horizontal 7-sums of (x, y), (xx, yy) and xy for its two columns (sliding: the
// second column reuses the first column's inner sum) and updates RUNNING vertical 7-sums: V += h_new -
// h_oldest, with the last seven h kept in a register ring (unrolled-by-7 loop).  Everything that exists for x
// and for y is one packed FADD2 / FMUL2 / FFMA2 on the pair (f32x2.cuh): same roundings, half the issue slots.
// The running sums restart in every tile, so the result does not depend on how the batch was partitioned.
// ---------------------------------------------------------------------------------------------
constexpr int kSsimThreads = 128;

__global__ void __launch_bounds__(kSsimThreads)
k_ssim(const float2* __restrict__ spec2, const long long* __restrict__ spec_off,
       const long long* __restrict__ offsets, int pair0, int n_fft, int hop, int F, int tiles_x, int tiles_per_pair,
       double* __restrict__ ssim_part) {
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
  const int Fp = spec_pitch_pairs(F);                // row pitch of the interleaved image (float2; even)
  const float2* img = spec2 + (spec_off[p] >> 1);    // spec_off counts floats
  // 7 stages = the unroll factor of the row loop: the stage a row lives in and the stage the next copy refills are
  // compile-time constants inside the unrolled body (no wrap-around arithmetic, selects or index registers)
  constexpr int RB = kSsimTC + 8, STAGES = 7;
  // row buffer: per column the pair (estimate, target) -- the pair is the unit every packed operation below works on
  __shared__ __align__(16) float2 rowbuf[STAGES][RB];
  __shared__ double red[kSsimThreads / 32];

  // per output column o (2 per thread): vertical running sums V1 = (sum x, sum y), V2 = (sum xx, sum yy), Vxy and
  // the ring of the last seven horizontal sums; all float2 arithmetic is FADD2 / FMUL2 / FFMA2 (f32x2.cuh)
  float2 R1[7][2], R2[7][2];
  float Rxy[7][2];
#pragma unroll
  for (int s = 0; s < 7; ++s)
#pragma unroll
    for (int o = 0; o < 2; ++o) {
      R1[s][o] = make_float2(0.f, 0.f);
      R2[s][o] = make_float2(0.f, 0.f);
      Rxy[s][o] = 0.f;
    }
  float2 V1[2], V2[2];
  float Vxy[2];
#pragma unroll
  for (int o = 0; o < 2; ++o) {
    V1[o] = make_float2(0.f, 0.f);
    V2[o] = make_float2(0.f, 0.f);
    Vxy[o] = 0.f;
  }
  float acc = 0.f;
  const float inv49 = 1.0f / 49.0f, cov_norm = 49.0f / 48.0f;
  const float C1 = 0.0004f, C2 = 0.0036f;  // (0.01*2)^2, (0.03*2)^2
  const float2 inv49_2 = make_float2(inv49, inv49), cov2 = make_float2(cov_norm, cov_norm);

  // rows stream global -> shared with 16-byte cp.async (LDGSTS.128), STAGES-1 rows in flight.  K1 writes the image
  // as (estimate, target) pairs with an even row pitch, so a row of the tile is RB / 2 = 132 aligned 16-byte chunks that
  // land directly in the pair layout: ONE copy per thread and row (+ one for threads 0..3), one running pointer.
  // (Round 1 copied 4 bytes at a time from two separate images: 6 LDGSTS + their 64-bit addresses per thread and row,
  // a fifth of the kernel's instructions.)  Chunks beyond the row are zero-filled (src-size 0); the pad column of an
  // odd F holds whatever K1 left there: it only reaches outputs that ok0 / ok1 mask.
  const bool in0 = (c0 + c) < Fp, in1 = t < 4 && (c0 + kSsimTC + c) < Fp;
  const float2* g_next = img + (long long)r0 * Fp + c0 + (in0 ? c : 0);  // this thread's chunk of the row issued next
  const int off1 = in1 ? kSsimTC : 0;
  const unsigned sdst0 = (unsigned)__cvta_generic_to_shared(&rowbuf[0][c]);
  const unsigned sdst1 = (unsigned)__cvta_generic_to_shared(&rowbuf[0][t < 4 ? kSsimTC + c : c]);
  int rows_left = r_end - r0;
  constexpr unsigned kStageBytes = RB * sizeof(float2);
  auto issue_row = [&](int stage) {
    if (rows_left > 0) {
      const unsigned sb = stage * kStageBytes;
      asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(sdst0 + sb), "l"(g_next), "r"(in0 ? 16 : 0)
                   : "memory");
      if (t < 4)
        asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(sdst1 + sb), "l"(g_next + off1),
                     "r"(in1 ? 16 : 0)
                     : "memory");
      g_next += Fp;
      --rows_left;
    }
    asm volatile("cp.async.commit_group;" ::: "memory");
  };
#pragma unroll
  for (int k = 0; k < STAGES - 1; ++k) issue_row(k);

  for (int rb = r0; rb < r_end; rb += 7) {
#pragma unroll
    for (int s = 0; s < 7; ++s) {
      const int r = rb + s;
      if (r < r_end) {  // uniform across the CTA
        asm volatile("cp.async.wait_group %0;" ::"n"(STAGES - 2) : "memory");
        __syncthreads();  // row r has landed for everyone; row r-1 is fully consumed
        issue_row((s + STAGES - 1) % STAGES);  // refills the stage row r-1 occupied
        const int par = s;  // row r = rb + s lives in stage s (rb - r0 is a multiple of 7 = STAGES)
        float2 P[8];      // (x, y) of columns c .. c+7
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const float4 v = *reinterpret_cast<const float4*>(&rowbuf[par][c + 2 * j]);
          P[2 * j] = make_float2(v.x, v.y);
          P[2 * j + 1] = make_float2(v.z, v.w);
        }
        // inner sums over columns c+1 .. c+6, then the two outputs add their own end column
        float2 S1 = make_float2(0.f, 0.f), S2 = make_float2(0.f, 0.f);
        float sxy = 0.f;
#pragma unroll
        for (int j = 1; j < 7; ++j) {
          S1 = add2(S1, P[j]);
          S2 = fma2(P[j], P[j], S2);
          sxy = fmaf(P[j].x, P[j].y, sxy);
        }
#pragma unroll
        for (int o = 0; o < 2; ++o) {
          const float2 pe = P[o == 0 ? 0 : 7];
          const float2 H1 = add2(S1, pe);
          const float2 H2 = fma2(pe, pe, S2);
          const float hxy = fmaf(pe.x, pe.y, sxy);
          // (V - oldest) + newest: the oldest row's sums are dead before the newest are formed, so the ring slot is
          // overwritten in place (V + (newest - oldest) cost 10 register moves per row)
          V1[o] = add2(sub2(V1[o], R1[s][o]), H1);
          V2[o] = add2(sub2(V2[o], R2[s][o]), H2);
          Vxy[o] = (Vxy[o] - Rxy[s][o]) + hxy;
          R1[s][o] = H1;
          R2[s][o] = H2;
          Rxy[s][o] = hxy;
        }
        if (r - r0 >= 6) {
#pragma unroll
          for (int o = 0; o < 2; ++o) {
            const float2 u1 = mul2(V1[o], inv49_2);                    // (ux, uy)
            const float2 u2 = mul2(V2[o], inv49_2);                    // (uxx, uyy)
            const float uxy = Vxy[o] * inv49;
            const float2 var = mul2(cov2, fma2(make_float2(-u1.x, -u1.y), u1, u2));  // (vx, vy)
            const float vxy = cov_norm * (uxy - u1.x * u1.y);
            const float A1 = 2.f * u1.x * u1.y + C1, A2 = 2.f * vxy + C2;
            const float B1 = u1.x * u1.x + u1.y * u1.y + C1, B2 = var.x + var.y + C2;
            // B1 * B2 >= C1 * C2 = 1.44e-6: no denormal-range rescue needed, the bare MUFU.RCP (<= 1 ulp) + one multiply
            const float S = (A1 * A2) * rcp_approx(B1 * B2);
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


}  // namespace ssr
