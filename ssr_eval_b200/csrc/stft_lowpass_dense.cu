// stft_lowpass_dense.cu -- K4d: the reference-faithful ("dense") mode of the STFT hard low-pass.
//
// ssr_eval/lowpass.py:17-28 runs torchlibrosa's STFT / ISTFT (ssr_eval/dsp.py:21-39): NOT an FFT but dense float32
// dot products of length n_fft -- conv1d with (DFT matrix x Hann) kernels, and a 1x1 conv with the
// (IDFT matrix / n_fft x Hann) kernels over the Hermitian-extended spectrum.  The bins above the cutoff of the
// low-passed waveform ARE the rounding noise of that arithmetic, and LSD / log-sispec of such an estimate measure it
// (DESIGN.md section 3): an FFT (K4, ~11 roundings per output) leaves a lower floor than 2048-term float32 dot
// products and moves LSD by ~+0.25.  This mode reproduces the dense arithmetic: the same float32 matrices (built on
// the host with torchlibrosa's formula and handed in through the C ABI) and the accumulation order of the reference's
// convolutions as torch 2.x / oneDNN run them on an AVX-512 host (identified by bit-comparing emulations against
// torch's CPU conv1d, profiles/r02_dense_dft_study.md): STFT (stride-hop conv1d, one input channel) = ONE float32 FMA
// accumulator per output over all n_fft taps in ascending order; ISTFT (1x1 conv over n_fft channels) = one FMA
// accumulator per K-block of 256 channels, the block sums added in ascending order; IEEE sqrt / division for the
// magnitude / cos / sin round trip; real and imaginary products subtracted at the end (torchlibrosa ISTFT.forward);
// overlap-add (F.fold) from the LAST covering frame to the first; division by the clamped overlap-added window^2.
// With the same matrices the GPU result is bit-identical to the CPU reference for almost every sample (tests).
//
// This IS a dense contraction (the reference's own choice); it runs on the FP32 FMA pipe, not on tensor cores -- TF32 /
// BF16 would change the very rounding noise the mode exists to reproduce.  ~12.6 GFLOP per 5 s utterance: an opt-in
// parity mode (a few k utterances/s), not the fast path.
#include <vector>

#include "common.cuh"

struct ssr_lowpass_dense_plan {
  int n_fft, hop, F, device;
  float* blob;     // one allocation: w_fwd [2F][n_fft] (real rows, then imag rows), v_real / v_imag [n_fft][n_fft], win_sq [n_fft]
  const float* w_fwd;
  const float* v_real;
  const float* v_imag;
  const float* win_sq;
};

namespace ssr {

__device__ __forceinline__ long long dense_reflect(long long i, long long L) {
  if (i >= 0 && i < L) return i;
  if (L == 1) return 0;
  const long long period = 2 * (L - 1);
  i %= period;
  if (i < 0) i += period;
  return i < L ? i : period - i;
}

// frame_off[u] = first frame (row of the frame matrix) of utterance u0 + u; frame_off[nu] = total
__global__ void k_dense_setup(const long long* __restrict__ offsets, int u0, int nu, int hop, long long* __restrict__ frame_off) {
  if (threadIdx.x == 0 && blockIdx.x == 0) {
    long long t = 0;
    for (int u = 0; u < nu; ++u) {
      frame_off[u] = t;
      t += (offsets[u0 + u + 1] - offsets[u0 + u]) / hop + 1;
    }
    frame_off[nu] = t;
  }
}

// A[row][n] = x_padded[t * hop + n] (reflect padding n_fft / 2), one CTA per frame
__global__ void __launch_bounds__(256) k_dense_frames(const float* __restrict__ x, const long long* __restrict__ offsets,
                                                      int u0, const long long* __restrict__ frame_off, int n_fft, int hop,
                                                      float* __restrict__ A) {
  const int u = blockIdx.y;
  const long long T = frame_off[u + 1] - frame_off[u];
  const long long off = offsets[u0 + u], L = offsets[u0 + u + 1] - off;
  for (long long t = blockIdx.x; t < T; t += gridDim.x) {
    float* row = A + (frame_off[u] + t) * n_fft;
    const long long s = t * hop - n_fft / 2;
    for (int n = threadIdx.x; n < n_fft; n += blockDim.x) row[n] = __ldg(x + off + dense_reflect(s + n, L));
  }
}

// C[M][N] (op)= A[M][K] * B[N][K]^T, float32.  One FMA accumulator per output over K-blocks of KBLOCK taps (ascending k;
// KBLOCK = 0: a single block), block sums added in ascending order: ctot = ctot + acc.  SUB: C = C_old - ctot (the
// "conv_real(..) - conv_imag(..)" of torchlibrosa's ISTFT).  128 x 128 tile, BK = 16, 256 threads x (8 x 8) outputs,
// double-buffered shared tiles.
constexpr int kGM = 128, kGN = 128, kGK = 16, kGPad = 4;

template <bool SUB, int KBLOCK>
__global__ void __launch_bounds__(256) k_dense_sgemm(const float* __restrict__ A, const float* __restrict__ B,
                                                     float* __restrict__ C, long long M, int N, int K, int ldc) {
  __shared__ __align__(16) float As[2][kGK][kGM + kGPad];
  __shared__ __align__(16) float Bs[2][kGK][kGN + kGPad];
  const int tid = threadIdx.x;
  const long long m0 = (long long)blockIdx.y * kGM;
  const int n0 = blockIdx.x * kGN;
  const int tx = tid & 15, ty = tid >> 4;  // 16 x 16 threads; thread owns rows ty*4+{0..3}, 64+ty*4+{0..3}; cols likewise
  // global -> shared: 128 rows x 16 k per operand = 512 float4; thread loads 2 per operand
  const int lr = tid >> 2;        // 0..63 (+64 for the second)
  const int lk = (tid & 3) * 4;   // 0,4,8,12
  float4 ra[2], rb[2];
  auto gload = [&](int k0) {
#pragma unroll
    for (int i = 0; i < 2; ++i) {
      const long long r = m0 + lr + 64 * i;
      ra[i] = r < M ? __ldg(reinterpret_cast<const float4*>(A + r * K + k0 + lk)) : make_float4(0.f, 0.f, 0.f, 0.f);
      const int c = n0 + lr + 64 * i;
      rb[i] = c < N ? __ldg(reinterpret_cast<const float4*>(B + (long long)c * K + k0 + lk)) : make_float4(0.f, 0.f, 0.f, 0.f);
    }
  };
  auto sstore = [&](int s) {
#pragma unroll
    for (int i = 0; i < 2; ++i) {
      const int r = lr + 64 * i;
      As[s][lk + 0][r] = ra[i].x;
      As[s][lk + 1][r] = ra[i].y;
      As[s][lk + 2][r] = ra[i].z;
      As[s][lk + 3][r] = ra[i].w;
      Bs[s][lk + 0][r] = rb[i].x;
      Bs[s][lk + 1][r] = rb[i].y;
      Bs[s][lk + 2][r] = rb[i].z;
      Bs[s][lk + 3][r] = rb[i].w;
    }
  };
  float acc[8][8], tot[8][8];
#pragma unroll
  for (int i = 0; i < 8; ++i)
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[i][j] = tot[i][j] = 0.f;
  gload(0);
  sstore(0);
  __syncthreads();
  const int nk = K / kGK;
  for (int kt = 0; kt < nk; ++kt) {
    const int s = kt & 1;
    if (kt + 1 < nk) gload((kt + 1) * kGK);
#pragma unroll
    for (int k = 0; k < kGK; ++k) {
      const float4 a0 = *reinterpret_cast<const float4*>(&As[s][k][ty * 4]);
      const float4 a1 = *reinterpret_cast<const float4*>(&As[s][k][64 + ty * 4]);
      const float4 b0 = *reinterpret_cast<const float4*>(&Bs[s][k][tx * 4]);
      const float4 b1 = *reinterpret_cast<const float4*>(&Bs[s][k][64 + tx * 4]);
      const float av[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
      const float bv[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
#pragma unroll
      for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[i][j] = __fmaf_rn(av[i], bv[j], acc[i][j]);
    }
    if ((KBLOCK > 0 && ((kt + 1) * kGK) % KBLOCK == 0) || kt + 1 == nk) {  // end of a K-block: fold it into the total
#pragma unroll
      for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          tot[i][j] = __fadd_rn(tot[i][j], acc[i][j]);
          acc[i][j] = 0.f;
        }
    }
    if (kt + 1 < nk) {
      sstore(s ^ 1);
      __syncthreads();
    }
  }
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const long long r = m0 + (i < 4 ? ty * 4 + i : 64 + ty * 4 + (i - 4));
    if (r >= M) continue;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const int c = n0 + (j < 4 ? tx * 4 + j : 64 + tx * 4 + (j - 4));
      if (c >= N) continue;
      float* p = C + r * ldc + c;
      *p = SUB ? __fsub_rn(*p, tot[i][j]) : tot[i][j];
    }
  }
}

// spectrogram_phase + zeroing + spectrogram_phase_to_wav's products (dsp.py:76-81, 107-119; lowpass.py:22-25), then
// torchlibrosa's Hermitian extension: D_re[t][k] = D_re[t][N-k] = mag'*cos, D_im[t][k] = -D_im[t][N-k] = mag'*sin
__global__ void __launch_bounds__(256) k_dense_phase(const float* __restrict__ Cfwd, const long long* __restrict__ frame_off,
                                                     const int* __restrict__ cut_bins, int u0, int n_fft, int F,
                                                     float* __restrict__ Dre, float* __restrict__ Dim) {
  const int u = blockIdx.y;
  const long long T = frame_off[u + 1] - frame_off[u];
  const int cut = cut_bins[u0 + u];
  for (long long t = blockIdx.x; t < T; t += gridDim.x) {
    const long long row = frame_off[u] + t;
    const float* c = Cfwd + row * (2LL * F);
    float* dre = Dre + row * n_fft;
    float* dim = Dim + row * n_fft;
    for (int k = threadIdx.x; k < F; k += blockDim.x) {
      const float re = c[k], im = c[F + k];
      const float p = __fadd_rn(__fmul_rn(re, re), __fmul_rn(im, im));
      const float mag = __fsqrt_rn(fmaxf(p, 1e-8f));     // torch.clamp(.., 1e-8, inf) ** 0.5
      const float cs = __fdiv_rn(re, mag), sn = __fdiv_rn(im, mag);
      const float m = k < cut ? mag : 0.f;                // sps[..., cut:] = 0
      const float r2 = __fmul_rn(m, cs), i2 = __fmul_rn(m, sn);
      dre[k] = r2;
      dim[k] = i2;
      if (k >= 1 && k <= F - 2) {
        dre[n_fft - k] = r2;
        dim[n_fft - k] = -i2;
      }
    }
  }
}

// overlap-add from the last covering frame to the first (the order of torch's F.fold), / clamp(overlap-added
// window^2, 1e-11), drop n_fft / 2, keep `length` samples
__global__ void __launch_bounds__(256) k_dense_ola(const float* __restrict__ S, const long long* __restrict__ frame_off,
                                                   const long long* __restrict__ offsets, int u0, int n_fft, int hop,
                                                   const float* __restrict__ win_sq, float* __restrict__ y) {
  const int u = blockIdx.y;
  const long long T = frame_off[u + 1] - frame_off[u];
  const long long off = offsets[u0 + u], L = offsets[u0 + u + 1] - off;
  const float* Su = S + frame_off[u] * n_fft;
  for (long long m = (long long)blockIdx.x * blockDim.x + threadIdx.x; m < L; m += (long long)gridDim.x * blockDim.x) {
    const long long p = m + n_fft / 2;
    const long long ta = p - n_fft >= 0 ? (p - n_fft) / hop + 1 : 0;
    const long long tb = min(T - 1, p / hop);
    float acc = 0.f, ws = 0.f;
    for (long long t = tb; t >= ta; --t) {
      const int i = (int)(p - t * hop);
      acc = __fadd_rn(acc, Su[t * n_fft + i]);
      ws = __fadd_rn(ws, win_sq[i]);
    }
    y[off + m] = __fdiv_rn(acc, fmaxf(ws, 1e-11f));
  }
}

static size_t dense_bytes_per_frame(int n_fft, int F) {
  // A (n_fft) + Cfwd (2F) + Dre (n_fft) + Dim (n_fft) + S (n_fft) floats
  return sizeof(float) * ((size_t)4 * n_fft + 2 * (size_t)F);
}
// the five arrays of a chunk start on 256-byte boundaries (the sgemm loads float4): slack reserved per chunk
constexpr size_t kDenseSlack = 5 * 256;

}  // namespace ssr

using namespace ssr;

extern "C" {

int ssr_lowpass_dense_plan_create(ssr_lowpass_dense_plan** out, int n_fft, int hop, const float* stft_w_real_host,
                                  const float* stft_w_imag_host, const float* istft_w_real_host,
                                  const float* istft_w_imag_host, const float* ola_window_host) {
  if (!out) return fail(SSR_ERR_INVALID, "plan pointer is NULL");
  *out = nullptr;
  if (n_fft < 64 || n_fft > 4096 || (n_fft % 16) != 0) return fail(SSR_ERR_INVALID, "dense low-pass: n_fft must be a multiple of 16 in [64, 4096]");
  if (hop < 1 || hop > n_fft) return fail(SSR_ERR_INVALID, "dense low-pass: hop must be in [1, n_fft]");
  if (!stft_w_real_host || !stft_w_imag_host || !istft_w_real_host || !istft_w_imag_host || !ola_window_host)
    return fail(SSR_ERR_INVALID, "dense low-pass: matrix pointer is NULL");
  const int F = n_fft / 2 + 1;
  const size_t n_w = (size_t)2 * F * n_fft, n_v = (size_t)n_fft * n_fft;
  ssr_lowpass_dense_plan* p = new ssr_lowpass_dense_plan();
  p->n_fft = n_fft;
  p->hop = hop;
  p->F = F;
  p->blob = nullptr;
  cudaError_t e = cudaGetDevice(&p->device);
  if (e == cudaSuccess) e = cudaMalloc(&p->blob, sizeof(float) * (n_w + 2 * n_v + n_fft));
  float* d = p->blob;
  if (e == cudaSuccess) e = cudaMemcpy(d, stft_w_real_host, sizeof(float) * (size_t)F * n_fft, cudaMemcpyHostToDevice);
  if (e == cudaSuccess) e = cudaMemcpy(d + (size_t)F * n_fft, stft_w_imag_host, sizeof(float) * (size_t)F * n_fft, cudaMemcpyHostToDevice);
  if (e == cudaSuccess) e = cudaMemcpy(d + n_w, istft_w_real_host, sizeof(float) * n_v, cudaMemcpyHostToDevice);
  if (e == cudaSuccess) e = cudaMemcpy(d + n_w + n_v, istft_w_imag_host, sizeof(float) * n_v, cudaMemcpyHostToDevice);
  if (e == cudaSuccess) e = cudaMemcpy(d + n_w + 2 * n_v, ola_window_host, sizeof(float) * n_fft, cudaMemcpyHostToDevice);
  if (e != cudaSuccess) {
    if (p->blob) cudaFree(p->blob);
    delete p;
    return fail(SSR_ERR_CUDA, std::string("dense low-pass plan upload: ") + cudaGetErrorString(e));
  }
  p->w_fwd = d;
  p->v_real = d + n_w;
  p->v_imag = d + n_w + n_v;
  p->win_sq = d + n_w + 2 * n_v;
  *out = p;
  return SSR_OK;
}

int ssr_lowpass_dense_plan_destroy(ssr_lowpass_dense_plan* plan) {
  if (!plan) return SSR_OK;
  if (plan->blob) cudaFree(plan->blob);
  delete plan;
  return SSR_OK;
}

/* bytes that let the whole batch run as one chunk; any workspace that holds the longest utterance is accepted
 * (the batch is then processed in several chunks) */
size_t ssr_stft_hard_lowpass_dense_workspace_bytes(const ssr_lowpass_dense_plan* plan, const int64_t* offsets_host, int n) {
  if (!plan || !offsets_host || n < 1) return 0;
  long long frames = 0;
  for (int u = 0; u < n; ++u) frames += (offsets_host[u + 1] - offsets_host[u]) / plan->hop + 1;
  return align_up(sizeof(long long) * (size_t)(n + 1), 256) + (size_t)frames * dense_bytes_per_frame(plan->n_fft, plan->F) + kDenseSlack;
}

int ssr_stft_hard_lowpass_dense_batched(const ssr_lowpass_dense_plan* plan, const float* x_dev, const int64_t* offsets_host,
                                        const int64_t* offsets_dev, int n, const int32_t* cut_bins_dev, float* y_dev,
                                        void* workspace_dev, size_t workspace_bytes, void* stream) {
  if (!plan || !x_dev || !offsets_host || !offsets_dev || !cut_bins_dev || !y_dev || !workspace_dev || n < 1)
    return fail(SSR_ERR_INVALID, "ssr_stft_hard_lowpass_dense_batched: bad argument");
  if (offsets_host[0] != 0) return fail(SSR_ERR_INVALID, "offsets must start at 0");
  const int N = plan->n_fft, F = plan->F, hop = plan->hop;
  for (int u = 0; u < n; ++u) {
    const long long L = offsets_host[u + 1] - offsets_host[u];
    if (L < 1) return fail(SSR_ERR_INVALID, "empty utterance in batch");
    if (L <= N / 2) return fail(SSR_ERR_INVALID, "utterance too short for reflect padding (needs > n_fft/2 samples)");
  }
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const long long* offs = reinterpret_cast<const long long*>(offsets_dev);
  const size_t head = align_up(sizeof(long long) * (size_t)(n + 1), 256);
  if (workspace_bytes <= head + kDenseSlack) return fail(SSR_ERR_WORKSPACE, "workspace too small");
  const size_t per_frame = dense_bytes_per_frame(N, F);
  const long long cap = (long long)((workspace_bytes - head - kDenseSlack) / per_frame);
  unsigned char* ws = static_cast<unsigned char*>(workspace_dev);
  long long* frame_off = reinterpret_cast<long long*>(ws);
  int u0 = 0;
  while (u0 < n) {
    long long frames = 0;
    int nu = 0;
    while (u0 + nu < n && nu < 32768) {
      const long long T = (offsets_host[u0 + nu + 1] - offsets_host[u0 + nu]) / hop + 1;
      if (frames + T > cap) break;
      frames += T;
      ++nu;
    }
    if (nu == 0) return fail(SSR_ERR_WORKSPACE, "workspace too small for the longest utterance");
    unsigned char* p = ws + head;
    auto carve = [&](size_t floats) {
      float* r = reinterpret_cast<float*>(p);
      p += align_up(sizeof(float) * floats, 256);
      return r;
    };
    float* A = carve((size_t)frames * N);
    float* Cf = carve((size_t)frames * 2 * F);
    float* Dre = carve((size_t)frames * N);
    float* Dim = carve((size_t)frames * N);
    float* S = carve((size_t)frames * N);
    k_dense_setup<<<1, 32, 0, st>>>(offs, u0, nu, hop, frame_off);
    SSR_LAUNCH_CHECK("k_dense_setup");
    k_dense_frames<<<dim3(64, nu), 256, 0, st>>>(x_dev, offs, u0, frame_off, N, hop, A);
    SSR_LAUNCH_CHECK("k_dense_frames");
    const unsigned gy = (unsigned)((frames + kGM - 1) / kGM);
    k_dense_sgemm<false, 0><<<dim3((2 * F + kGN - 1) / kGN, gy), 256, 0, st>>>(A, plan->w_fwd, Cf, frames, 2 * F, N, 2 * F);
    SSR_LAUNCH_CHECK("k_dense_sgemm");
    k_dense_phase<<<dim3(64, nu), 256, 0, st>>>(Cf, frame_off, cut_bins_dev, u0, N, F, Dre, Dim);
    SSR_LAUNCH_CHECK("k_dense_phase");
    k_dense_sgemm<false, 256><<<dim3((N + kGN - 1) / kGN, gy), 256, 0, st>>>(Dre, plan->v_real, S, frames, N, N, N);
    SSR_LAUNCH_CHECK("k_dense_sgemm");
    k_dense_sgemm<true, 256><<<dim3((N + kGN - 1) / kGN, gy), 256, 0, st>>>(Dim, plan->v_imag, S, frames, N, N, N);
    SSR_LAUNCH_CHECK("k_dense_sgemm");
    k_dense_ola<<<dim3(32, nu), 256, 0, st>>>(S, frame_off, offs, u0, N, hop, plan->win_sq, y_dev);
    SSR_LAUNCH_CHECK("k_dense_ola");
    u0 += nu;
  }
  return SSR_OK;
}

}  // extern "C"
