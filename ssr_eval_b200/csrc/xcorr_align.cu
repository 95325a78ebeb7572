// xcorr_align.cu -- K8 ("next" row): the alignment step of the mp3 degradation,
//   shft01 = np.argmax(scipy.signal.correlate(decoded, x)) - x.shape[0]          (ssr_eval/eval.py:319-320)
// for a batch of (decoded, x) pairs of equal length L (eval.py:318 unifies the lengths first).  The codec itself
// (sox, eval.py:308-316) stays out of scope; this is the part that is arithmetic.
//
// scipy's 'full' cross-correlation z[k] = sum_l a[l] x[l - k + L - 1], k = 0 .. 2L-2, is the circular correlation
// c[m] = sum_l a[l + m] x[l] at lag m = k - (L - 1) once both signals are zero-padded to N >= 2L - 1.  Per pair:
//   1. z = a + i x, zero-padded to N = 2^m (one complex transform carries both real signals);
//   2. forward FFT of length N = N1 x N2 in two steps through shared memory ("four-step" FFT): N2 column transforms
//      of length N1 (8 columns per CTA, one warp each, 64-byte coalesced segments) times the twiddles W_N^{n2 k1},
//      then N1 contiguous row transforms of length N2; bin k = k1 + N1 k2 ends up at [k1][k2];
//   3. cross spectrum S[k] = A[k] conj(X[k]) from Z[k] and Z[N - k] (A = (Z[k] + conj Z[N-k]) / 2, X = (Z[k] - conj
//      Z[N-k]) / 2i), written in place for k and N - k by one thread (S is Hermitian);
//   4. inverse FFT (the same two kernels with conjugated input / output, rows first), unnormalised -- a positive
//      scale does not move the argmax;
//   5. argmax over k = 0 .. 2L-2 of Re c[(k - L + 1) mod N], first maximum wins (np.argmax).
// float32 complex arithmetic, like scipy's own FFT path for float32 inputs (scipy.signal.correlate -> fftconvolve ->
// pocketfft in single precision): the result is an INDEX, compared bit-exactly with scipy in the tests on delayed /
// noisy copies.  Sub-transforms run warp-locally on fft_core.cuh (radix 8 / 4, __syncwarp between passes).
#include <math.h>

#include <algorithm>
#include <vector>

#include "common.cuh"
#include "fft_core.cuh"

namespace ssr {

constexpr int kXcLanes = 8;  // columns (or rows) per CTA, one warp each

struct SyncWarp {
  __device__ __forceinline__ void operator()() const { __syncwarp(); }
};

template <int LOGM>
__device__ __forceinline__ int xc_dif_position(int k) {
  constexpr int A = n_r8(LOGM), B = n_r4(LOGM);
  int pos = 0, n = 1 << LOGM;
#pragma unroll
  for (int s = 0; s < A + B; ++s) {
    const int r = (s < A) ? 8 : 4;
    pos += (k % r) * (n / r);
    k /= r;
    n /= r;
  }
  return pos;
}

__device__ __forceinline__ cf xc_twiddle(unsigned prod, int logN) {  // exp(-2 pi i prod / N)
  const unsigned r = prod & ((1u << logN) - 1u);
  float s, c;
  sincospif(-2.0f * (float)r / (float)(1u << logN), &s, &c);
  return cf{c, s};
}

// z[n] = a[n] + i x[n] (n < L), 0 up to N
__global__ void __launch_bounds__(256) k_xc_pack(const float* __restrict__ a, const float* __restrict__ x,
                                                 const long long* __restrict__ offsets, const int* __restrict__ ids,
                                                 int logN, cf* __restrict__ z) {
  const int u = ids[blockIdx.y];
  const long long off = offsets[u], L = offsets[u + 1] - off;
  const long long N = 1LL << logN;
  cf* zu = z + (long long)blockIdx.y * N;
  for (long long n = (long long)blockIdx.x * blockDim.x + threadIdx.x; n < N; n += (long long)gridDim.x * blockDim.x)
    zu[n] = n < L ? cf{__ldg(a + off + n), __ldg(x + off + n)} : cf{0.f, 0.f};
}

// One step of the four-step FFT.  The data of an utterance is a [rows][cols] matrix (rows = 1 << logN1 = N1,
// cols = N2).  COLS: transform along the row index for 8 adjacent columns (length N1 = 1 << LOGM); else along the
// column index for 8 adjacent rows (length N2 = 1 << LOGM).  TW: multiply output [r][c] by W_N^{r c}.
// CONJ: conjugate on load and on store (inverse transform through the forward machinery).
template <int LOGM, bool COLS, bool TW, bool CONJ>
__global__ void __launch_bounds__(32 * kXcLanes) k_xc_fft_step(cf* __restrict__ z, int logN, int logN1) {
  constexpr int M = 1 << LOGM;
  extern __shared__ __align__(16) unsigned char xc_smem[];
  cf* const tw = reinterpret_cast<cf*>(xc_smem);                       // exp(-2 pi i n / M)
  unsigned short* const pos = reinterpret_cast<unsigned short*>(tw + M);  // slot of output k after the DIF passes
  cf* const bufs = reinterpret_cast<cf*>(xc_smem + sizeof(cf) * M + sizeof(unsigned short) * M);
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int logN2 = logN - logN1;
  const long long N = 1LL << logN;
  const int n_cols = 1 << logN2;
  cf* const zu = z + (long long)blockIdx.y * N;
  for (int n = tid; n < M; n += blockDim.x) {
    float s, c;
    sincospif(-2.0f * (float)n / (float)M, &s, &c);
    tw[n] = cf{c, s};
    pos[n] = (unsigned short)pad_idx(xc_dif_position<LOGM>(n));
  }
  cf* const buf = bufs + (size_t)warp * padded_size(M);
  const int first = blockIdx.x * kXcLanes;  // first column (COLS) or row of this CTA
  if (COLS) {
    // element (n1, first + c): 8 consecutive threads read 8 consecutive complex values (64 bytes)
    for (int i = tid; i < M * kXcLanes; i += blockDim.x) {
      const int c = i & (kXcLanes - 1), n1 = i / kXcLanes;
      cf v = zu[(long long)n1 * n_cols + first + c];
      if (CONJ) v.y = -v.y;
      bufs[(size_t)c * padded_size(M) + pad_idx(n1)] = v;
    }
    __syncthreads();
  } else {
    const cf* row = zu + (long long)(first + warp) * n_cols;
    for (int n2 = lane; n2 < M; n2 += 32) {
      cf v = row[n2];
      if (CONJ) v.y = -v.y;
      buf[pad_idx(n2)] = v;
    }
    __syncthreads();  // (also orders the twiddle / position tables)
  }
  fft_forward_dif<LOGM>(buf, tw, lane, 32, SyncWarp{});
  if (COLS) {
    __syncthreads();
    for (int i = tid; i < M * kXcLanes; i += blockDim.x) {
      const int c = i & (kXcLanes - 1), k1 = i / kXcLanes;
      cf v = bufs[(size_t)c * padded_size(M) + pos[k1]];
      if (TW) v = cmul(v, xc_twiddle((unsigned)k1 * (unsigned)(first + c), logN));
      if (CONJ) v.y = -v.y;
      zu[(long long)k1 * n_cols + first + c] = v;
    }
  } else {
    __syncwarp();
    cf* row = zu + (long long)(first + warp) * n_cols;
    for (int k2 = lane; k2 < M; k2 += 32) {
      cf v = buf[pos[k2]];
      if (TW) v = cmul(v, xc_twiddle((unsigned)(first + warp) * (unsigned)k2, logN));
      if (CONJ) v.y = -v.y;
      row[k2] = v;
    }
  }
}

// S[k] = A[k] conj(X[k]) with A, X the spectra of the real and imaginary part of z; bin k = k1 + N1 k2 sits at
// [k1][k2].  One thread owns the pair (k, N - k), k <= N / 2.
__global__ void __launch_bounds__(256) k_xc_cross(cf* __restrict__ z, int logN, int logN1) {
  const long long N = 1LL << logN;
  const int logN2 = logN - logN1;
  const unsigned N1m = (1u << logN1) - 1u;
  cf* zu = z + (long long)blockIdx.y * N;
  for (long long k = (long long)blockIdx.x * blockDim.x + threadIdx.x; k <= N / 2; k += (long long)gridDim.x * blockDim.x) {
    const long long kn = (N - k) & (N - 1);
    const long long ia = ((long long)(k & N1m) << logN2) + (k >> logN1);
    const long long ib = ((long long)(kn & N1m) << logN2) + (kn >> logN1);
    const cf zk = zu[ia], zn = zu[ib];
    const cf A = cf{0.5f * (zk.x + zn.x), 0.5f * (zk.y - zn.y)};
    const cf X = cf{0.5f * (zk.y + zn.y), 0.5f * (zn.x - zk.x)};
    const cf S = cmul_conj(A, X);
    zu[ia] = S;
    if (ib != ia) zu[ib] = cf{S.x, -S.y};
  }
}

// argmax over scipy's index k = 0 .. 2L-2 (lag k - (L-1)), first maximum wins
__global__ void __launch_bounds__(1024) k_xc_argmax(const cf* __restrict__ z, const long long* __restrict__ offsets,
                                                    const int* __restrict__ ids, int logN, long long* __restrict__ out) {
  const int u = ids[blockIdx.x];
  const long long L = offsets[u + 1] - offsets[u];
  const long long N = 1LL << logN;
  const cf* zu = z + (long long)blockIdx.x * N;
  constexpr long long kNone = 0x7fffffffffffffffLL;
  float best = -INFINITY;
  long long arg = kNone;
  for (long long k = threadIdx.x; k < 2 * L - 1; k += blockDim.x) {  // ascending k: ">" keeps the first maximum
    const float v = zu[(k - (L - 1)) & (N - 1)].x;
    if (v > best) {
      best = v;
      arg = k;
    }
  }
  __shared__ float sb[1024];
  __shared__ long long sa[1024];
  sb[threadIdx.x] = best;
  sa[threadIdx.x] = arg;
  __syncthreads();
  for (int s = 512; s > 0; s >>= 1) {
    if ((int)threadIdx.x < s) {
      const float v = sb[threadIdx.x + s];
      const long long a2 = sa[threadIdx.x + s];
      if (a2 != kNone && (sa[threadIdx.x] == kNone || v > sb[threadIdx.x] || (v == sb[threadIdx.x] && a2 < sa[threadIdx.x]))) {
        sb[threadIdx.x] = v;
        sa[threadIdx.x] = a2;
      }
    }
    __syncthreads();
  }
  if (threadIdx.x == 0) out[u] = sa[0] == kNone ? 0 : sa[0];  // all -inf / NaN: np.argmax gives 0
}

static int xc_log2_size(long long L) {
  int m = 12;
  while ((1LL << m) < 2 * L - 1) ++m;
  return m;
}

template <int LOGM, bool COLS, bool TW, bool CONJ>
static int xc_launch_step(cf* z, int logN, int logN1, int nu, cudaStream_t st) {
  constexpr int M = 1 << LOGM;
  const size_t smem = sizeof(cf) * M + sizeof(unsigned short) * M + sizeof(cf) * (size_t)kXcLanes * padded_size(M);
  auto kern = k_xc_fft_step<LOGM, COLS, TW, CONJ>;
  SSR_CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  const int other = COLS ? (1 << (logN - logN1)) : (1 << logN1);  // number of columns (COLS) / rows
  kern<<<dim3(other / kXcLanes, nu), 32 * kXcLanes, smem, st>>>(z, logN, logN1);
  SSR_LAUNCH_CHECK("k_xc_fft_step");
  return SSR_OK;
}

template <bool COLS, bool TW, bool CONJ>
static int xc_step(int logM, cf* z, int logN, int logN1, int nu, cudaStream_t st) {
  switch (logM) {
    case 6: return xc_launch_step<6, COLS, TW, CONJ>(z, logN, logN1, nu, st);
    case 7: return xc_launch_step<7, COLS, TW, CONJ>(z, logN, logN1, nu, st);
    case 8: return xc_launch_step<8, COLS, TW, CONJ>(z, logN, logN1, nu, st);
    case 9: return xc_launch_step<9, COLS, TW, CONJ>(z, logN, logN1, nu, st);
    case 10: return xc_launch_step<10, COLS, TW, CONJ>(z, logN, logN1, nu, st);
    default: return fail(SSR_ERR_INVALID, "xcorr: unsupported sub-transform size");
  }
}

}  // namespace ssr

using namespace ssr;

extern "C" {

/* bytes that let the whole batch run in one pass per FFT size; anything that holds the longest pair is accepted */
size_t ssr_xcorr_workspace_bytes(const int64_t* offsets_host, int n) {
  if (!offsets_host || n < 1) return 0;
  size_t total = 0;
  for (int u = 0; u < n; ++u) {
    const long long L = offsets_host[u + 1] - offsets_host[u];
    if (L < 1 || 2 * L - 1 > (1LL << 20)) return 0;
    total += sizeof(cf) << xc_log2_size(L);
  }
  return align_up(sizeof(int) * (size_t)n, 256) + total;
}

int ssr_xcorr_argmax_batched(const float* a_dev, const float* x_dev, const int64_t* offsets_host,
                             const int64_t* offsets_dev, int n, int64_t* argmax_dev, void* workspace_dev,
                             size_t workspace_bytes, void* stream) {
  if (!a_dev || !x_dev || !offsets_host || !offsets_dev || !argmax_dev || !workspace_dev || n < 1)
    return fail(SSR_ERR_INVALID, "ssr_xcorr_argmax_batched: bad argument");
  if (offsets_host[0] != 0) return fail(SSR_ERR_INVALID, "offsets must start at 0");
  std::vector<int> logs(n);
  for (int u = 0; u < n; ++u) {
    const long long L = offsets_host[u + 1] - offsets_host[u];
    if (L < 1) return fail(SSR_ERR_INVALID, "empty utterance in batch");
    if (2 * L - 1 > (1LL << 20)) return fail(SSR_ERR_INVALID, "xcorr: utterance longer than 524288 samples");
    logs[u] = xc_log2_size(L);
  }
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const long long* offs = reinterpret_cast<const long long*>(offsets_dev);
  const size_t head = align_up(sizeof(int) * (size_t)n, 256);
  if (workspace_bytes <= head) return fail(SSR_ERR_WORKSPACE, "workspace too small");
  unsigned char* ws = static_cast<unsigned char*>(workspace_dev);
  int* ids_dev = reinterpret_cast<int*>(ws);
  cf* z = reinterpret_cast<cf*>(ws + head);
  for (int m = 12; m <= 20; ++m) {
    std::vector<int> ids;
    for (int u = 0; u < n; ++u)
      if (logs[u] == m) ids.push_back(u);
    if (ids.empty()) continue;
    const size_t per = sizeof(cf) << m;
    const size_t cap = (workspace_bytes - head) / per;
    if (cap == 0) return fail(SSR_ERR_WORKSPACE, "workspace too small for the longest pair");
    const int logN1 = m / 2, logN2 = m - logN1;
    for (size_t s = 0; s < ids.size(); s += std::min<size_t>(cap, 32768)) {
      const int nu = (int)std::min<size_t>(std::min<size_t>(cap, 32768), ids.size() - s);
      // the id list of this pass; a pageable source is staged before the call returns, stream order keeps it safe
      SSR_CUDA_TRY(cudaMemcpyAsync(ids_dev, ids.data() + s, sizeof(int) * (size_t)nu, cudaMemcpyHostToDevice, st));
      k_xc_pack<<<dim3((unsigned)std::min<long long>((1LL << m) / 256, 1024), nu), 256, 0, st>>>(a_dev, x_dev, offs, ids_dev, m, z);
      SSR_LAUNCH_CHECK("k_xc_pack");
      int rc;
      if ((rc = xc_step<true, true, false>(logN1, z, m, logN1, nu, st)) != SSR_OK) return rc;    // columns + twiddles
      if ((rc = xc_step<false, false, false>(logN2, z, m, logN1, nu, st)) != SSR_OK) return rc;  // rows
      k_xc_cross<<<dim3((unsigned)std::min<long long>(((1LL << m) / 2 + 256) / 256, 1024), nu), 256, 0, st>>>(z, m, logN1);
      SSR_LAUNCH_CHECK("k_xc_cross");
      if ((rc = xc_step<false, true, true>(logN2, z, m, logN1, nu, st)) != SSR_OK) return rc;    // inverse rows + twiddles
      if ((rc = xc_step<true, false, true>(logN1, z, m, logN1, nu, st)) != SSR_OK) return rc;    // inverse columns
      k_xc_argmax<<<nu, 1024, 0, st>>>(z, offs, ids_dev, m, reinterpret_cast<long long*>(argmax_dev));
      SSR_LAUNCH_CHECK("k_xc_argmax");
    }
  }
  return SSR_OK;
}

}  // extern "C"
