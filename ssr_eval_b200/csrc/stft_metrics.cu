// stft_metrics.cu -- K1 (batched STFT -> magnitude -> LSD / sispec / log-sispec accumulators)
// and K2 (7x7 box-window SSIM) of the ssr_eval hot path, plus their C ABI.
//
// Reference semantics (paths relative to the reference repo):
//   ssr_eval/metrics.py:16-19   n_fft / hop from the sample rate
//   ssr_eval/metrics.py:26-30   |librosa.stft| : reflect pad, periodic Hann (f64), f64 FFT, c64
//   ssr_eval/metrics.py:109-112 lsd
//   ssr_eval/metrics.py:114-121 + ssr_eval/utils.py:68-92   sispec (energy_unify, pow_norm, ...)
//   ssr_eval/utils.py:43-44     to_log
//   ssr_eval/metrics.py:123-132 ssim -> skimage structural_similarity(win_size=7)
//
// Numerics: est and target of a pair are packed as ONE complex float64 transform
// (z = target + i*est), because the 1e-4 tolerance on LSD / log-sispec needs ~float64 FFT accuracy
// on hard-low-passed estimates (hard part 1).  After the transform the two
// spectra are separated, rounded to complex64 like librosa's store, and every metric formula runs
// in float32 exactly as torch does on the CPU; cross-frame sums are kept in float64.
#include <math.h>
#include <stdlib.h>
#include <string.h>

#include <vector>

#include "common.cuh"
#include "fft_core.cuh"
#include "stft_tables.hpp"
#include "k1_map.cuh"
#include "k1_common.cuh"
#include "k1_generic.cuh"
#include "k1_2048.cuh"
#include "k1_pfa.cuh"
#include "k2_ssim.cuh"

namespace ssr {

std::string& last_error_ref() {
  static thread_local std::string s;
  return s;
}
std::atomic<uint64_t>& launch_counter() {
  static std::atomic<uint64_t> c{0};
  return c;
}

struct TimingState {
  bool on = false;
  std::vector<std::pair<cudaEvent_t, cudaEvent_t>> pending, pool;
};
static TimingState& timing() {
  static TimingState t;
  return t;
}

}  // namespace ssr

struct ssr_stft_plan {
  int n_fft, hop, F, M, logM, bluestein, device;
  int pfa;          // 1: PFA path (pdev valid)
  void* blob;       // one device allocation holding all tables
  void* blob_pfa;
  ssr::StftDev dev;
  ssr::PfaDev pdev;
};

namespace ssr {

// ---------------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------------
// SSR_FORCE_GENERIC_K1=1 routes n_fft 2048 through the generic radix-8 kernel (A/B tests only)
static bool force_generic_k1() {
  static int v = -1;
  if (v < 0) {
    const char* e = getenv("SSR_FORCE_GENERIC_K1");
    v = (e && e[0] == '1') ? 1 : 0;
  }
  return v == 1;
}

struct WsLayout {
  size_t item_start, item_pair, spec_off, partials, ssim_part, spec_e, spec_t, total;
  int chunk, n_items, tiles_x, tiles_per_pair;
  long long total_frames;
};

static int plan_layout(const ssr_stft_plan* plan, const int64_t* offs, int n, unsigned flags, WsLayout* w) {
  long long total_frames = 0, max_T = 0;
  for (int p = 0; p < n; ++p) {
    long long L = offs[p + 1] - offs[p];
    if (L < 1) return fail(SSR_ERR_INVALID, "empty utterance in batch");
    long long T = stft_frames(L, plan->n_fft, plan->hop);
    total_frames += T;
    if (T > max_T) max_T = T;
  }
  // ~32 work items per resident CTA slot (SMs x 4): with dynamic scheduling the tail is at most one item
  long long want = (long long)sm_count() * 4 * 32;
  long long chunk = (total_frames + want - 1) / want;
  if (chunk < 4) chunk = 4;
  if (chunk > kMaxChunk) chunk = kMaxChunk;
  long long n_items = 0;
  for (int p = 0; p < n; ++p) {
    long long T = stft_frames(offs[p + 1] - offs[p], plan->n_fft, plan->hop);
    n_items += (T + chunk - 1) / chunk;
  }
  if (n_items > 0x7fffffffLL) return fail(SSR_ERR_INVALID, "batch too large");
  w->chunk = (int)chunk;
  w->n_items = (int)n_items;
  w->total_frames = total_frames;
  w->tiles_x = (plan->F - 6 + kSsimTC - 1) / kSsimTC;
  if (w->tiles_x < 1) w->tiles_x = 1;
  long long rows = max_T - 6;
  int tiles_y = rows > 0 ? (int)((rows + kSsimTR - 1) / kSsimTR) : 1;
  w->tiles_per_pair = w->tiles_x * tiles_y;
  size_t o = 0;
  w->item_start = o;
  o = align_up(o + sizeof(int) * (size_t)(n + 2), 256);  // + the work-item counter
  w->item_pair = o;
  o = align_up(o + sizeof(int) * (size_t)n_items, 256);
  w->spec_off = o;
  o = align_up(o + sizeof(long long) * (size_t)(n + 1), 256);
  w->partials = o;
  o = align_up(o + sizeof(double) * kPartials * (size_t)n_items, 256);
  w->ssim_part = o;
  if (flags & SSR_METRIC_SSIM) o = align_up(o + sizeof(double) * (size_t)n * w->tiles_per_pair, 256);
  // K1 -> K2 spectrograms: one interleaved image of (estimate, target) pairs, even row pitch (k1_common.cuh: SpecLayout)
  w->spec_e = o;
  w->spec_t = o + sizeof(float);
  if (flags & SSR_METRIC_SSIM)
    o = align_up(o + sizeof(float2) * (size_t)total_frames * spec_pitch_pairs(plan->F), 256);
  w->total = o;
  return SSR_OK;
}

template <int LOGM, bool BLUE, typename ET, typename TT>
static int launch_k1(const ssr_stft_plan* plan, int grid, size_t smem, cudaStream_t st,
                     const ET* est, const TT* tgt, const long long* offs_dev,
                     const int* item_start, const int* item_pair, int n_items, int chunk,
                     unsigned flags, double* partials, float* spec_e, float* spec_t,
                     const long long* spec_off, int* next_item) {
  auto kern = k_stft_metrics<LOGM, BLUE, ET, TT>;
  SSR_CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  TimingState& tm = timing();
  std::pair<cudaEvent_t, cudaEvent_t> ev{nullptr, nullptr};
  if (tm.on) {
    if (!tm.pool.empty()) {
      ev = tm.pool.back();
      tm.pool.pop_back();
    } else {
      SSR_CUDA_TRY(cudaEventCreate(&ev.first));
      SSR_CUDA_TRY(cudaEventCreate(&ev.second));
    }
    SSR_CUDA_TRY(cudaEventRecord(ev.first, st));
  }
  kern<<<grid, kThreads, smem, st>>>(plan->dev, est, tgt, offs_dev, item_start, item_pair, n_items,
                                     chunk, flags, partials, spec_e, spec_t, spec_off, next_item);
  SSR_LAUNCH_CHECK("k_stft_metrics");
  if (tm.on) {
    SSR_CUDA_TRY(cudaEventRecord(ev.second, st));
    tm.pending.push_back(ev);
  }
  return SSR_OK;
}

template <bool BLUE, typename ET, typename TT>
static int dispatch_k1(const ssr_stft_plan* plan, int grid, size_t smem, cudaStream_t st,
                       const ET* est, const TT* tgt, const long long* offs_dev,
                       const int* item_start, const int* item_pair, int n_items, int chunk,
                       unsigned flags, double* partials, float* spec_e, float* spec_t,
                       const long long* spec_off, int* next_item) {
#define SSR_CASE(LM)                                                                            \
  case LM:                                                                                      \
    return launch_k1<LM, BLUE, ET, TT>(plan, grid, smem, st, est, tgt, offs_dev, item_start, item_pair, \
                               n_items, chunk, flags, partials, spec_e, spec_t, spec_off, next_item);
  switch (plan->logM) {
    SSR_CASE(8)
    SSR_CASE(9)
    SSR_CASE(10)
    SSR_CASE(11)
    SSR_CASE(12)
    SSR_CASE(13)
    default:
      return fail(SSR_ERR_INVALID, "unsupported FFT size");
  }
#undef SSR_CASE
}

// est64 != nullptr: the estimate is a float64 batch (generic kernel, see k1_generic.cuh); est is ignored then
static int run_k1(const ssr_stft_plan* plan, const WsLayout& w, cudaStream_t st, const float* est,
                  const float* tgt, const long long* offs_dev, int n, unsigned flags,
                  unsigned char* ws, float* spec_e, float* spec_t, const double* est64 = nullptr,
                  const double* tgt64 = nullptr) {
  int* item_start = reinterpret_cast<int*>(ws + w.item_start);
  int* item_pair = reinterpret_cast<int*>(ws + w.item_pair);
  long long* spec_off = reinterpret_cast<long long*>(ws + w.spec_off);
  double* partials = reinterpret_cast<double*>(ws + w.partials);
  const bool interleaved = spec_e && spec_t == spec_e + 1;
  k_setup<<<1, 1024, 0, st>>>(offs_dev, n, plan->n_fft, plan->hop, w.chunk,
                              interleaved ? 2 * spec_pitch_pairs(plan->F) : plan->F, item_start, item_pair, spec_off);
  SSR_LAUNCH_CHECK("k_setup");
  size_t smem = sizeof(cd) * (size_t)padded_size(plan->M);
  const int sms = sm_count();
  int per_sm = (int)((200 * 1024) / (smem + 4096));
  if (per_sm < 1) per_sm = 1;
  if (per_sm > 2) per_sm = 2;
  int grid = sms * per_sm;
  if (grid > w.n_items) grid = w.n_items;
  if (est64 && tgt64) {  // float64 estimate AND target: everything stays float64 (generic kernel)
    if (plan->bluestein)
      return dispatch_k1<true, double, double>(plan, grid, smem, st, est64, tgt64, offs_dev, item_start, item_pair,
                                               w.n_items, w.chunk, flags, partials, spec_e, spec_t, spec_off, item_start + n + 1);
    return dispatch_k1<false, double, double>(plan, grid, smem, st, est64, tgt64, offs_dev, item_start, item_pair,
                                              w.n_items, w.chunk, flags, partials, spec_e, spec_t, spec_off, item_start + n + 1);
  }
  if (est64 && plan->pfa && !force_generic_k1()) {
    // float64 estimate on the PFA kernel (n_fft = R * P, e.g. 2229 at 48 kHz): same arithmetic as the generic kernel's
    // float64-estimate path at ~4x its speed -- the IIR low-pass keys of setting_lowpass_filtering all come this way
    const size_t smem_p = sizeof(cd) * (2048 + 256) + sizeof(cd) * (size_t)(plan->n_fft - plan->pdev.P);
    int gp = sms * SSR_PFA_CTAS;
    if (gp > w.n_items) gp = w.n_items;
    const int nq = (plan->pdev.P + 127) / 128;
#define SSR_PFA64_LAUNCH(NQ_)                                                                        \
  do {                                                                                               \
    auto kern = k_stft_metrics_pfa<NQ_, -1, double>;                                                 \
    SSR_CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_p)); \
    kern<<<gp, kV2Threads, smem_p, st>>>(plan->pdev, est64, tgt, offs_dev, item_start, item_pair, w.n_items, \
                                         w.chunk, flags, partials, spec_e, spec_t, spec_off, item_start + n + 1); \
  } while (0)
    if (nq <= 6) SSR_PFA64_LAUNCH(6);
    else SSR_PFA64_LAUNCH(8);
#undef SSR_PFA64_LAUNCH
    SSR_LAUNCH_CHECK("k_stft_metrics_pfa<double>");
    return SSR_OK;
  }
  if (est64 && !plan->bluestein && plan->logM == 11 && !force_generic_k1()) {
    // float64 estimate on the 2048-point kernel (evaluation at 44.1 kHz and BASELINE cfg 2 sizes)
    int g2 = sms * 4;
    if (g2 > w.n_items) g2 = w.n_items;
    const size_t smem2 = sizeof(cd) * (2048 + 256) + sizeof(float) * 2 * 1104;
    if (plan->hop == 512) {
      auto kern = k_stft_metrics_2048<-1, true, double>;
      SSR_CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem2));
      kern<<<g2, kV2Threads, smem2, st>>>(plan->dev, est64, tgt, offs_dev, item_start, item_pair, w.n_items, w.chunk,
                                          flags, partials, spec_e, spec_t, spec_off, item_start + n + 1);
    } else {
      auto kern = k_stft_metrics_2048<-1, false, double>;
      SSR_CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem2));
      kern<<<g2, kV2Threads, smem2, st>>>(plan->dev, est64, tgt, offs_dev, item_start, item_pair, w.n_items, w.chunk,
                                          flags, partials, spec_e, spec_t, spec_off, item_start + n + 1);
    }
    SSR_LAUNCH_CHECK("k_stft_metrics_2048<double>");
    return SSR_OK;
  }
  if (est64) {
    if (plan->bluestein)
      return dispatch_k1<true, double, float>(plan, grid, smem, st, est64, tgt, offs_dev, item_start, item_pair,
                                       w.n_items, w.chunk, flags, partials, spec_e, spec_t, spec_off, item_start + n + 1);
    return dispatch_k1<false, double, float>(plan, grid, smem, st, est64, tgt, offs_dev, item_start, item_pair,
                                      w.n_items, w.chunk, flags, partials, spec_e, spec_t, spec_off, item_start + n + 1);
  }
  if (plan->pfa && !force_generic_k1()) {
    const size_t smem_p = sizeof(cd) * (2048 + 256) + sizeof(cd) * (size_t)(plan->n_fft - plan->pdev.P);
    int gp = sms * SSR_PFA_CTAS;
    if (gp > w.n_items) gp = w.n_items;
    const bool store = spec_e || spec_t;
    const bool lsd_only = !store && (flags & 7u) == 1u;
    const int nq = (plan->pdev.P + 127) / 128;
    TimingState& tm = timing();
    std::pair<cudaEvent_t, cudaEvent_t> ev{nullptr, nullptr};
    if (tm.on) {
      if (!tm.pool.empty()) {
        ev = tm.pool.back();
        tm.pool.pop_back();
      } else {
        SSR_CUDA_TRY(cudaEventCreate(&ev.first));
        SSR_CUDA_TRY(cudaEventCreate(&ev.second));
      }
      SSR_CUDA_TRY(cudaEventRecord(ev.first, st));
    }
#define SSR_PFA_LAUNCH(NQ_, FX)                                                                      \
  do {                                                                                               \
    auto kern = k_stft_metrics_pfa<NQ_, FX>;                                                         \
    SSR_CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_p)); \
    kern<<<gp, kV2Threads, smem_p, st>>>(plan->pdev, est, tgt, offs_dev, item_start, item_pair, w.n_items, \
                                         w.chunk, flags, partials, spec_e, spec_t, spec_off, item_start + n + 1); \
  } while (0)
    if (nq <= 6) {
      if (lsd_only) SSR_PFA_LAUNCH(6, 1);
      else SSR_PFA_LAUNCH(6, -1);
    } else {
      if (lsd_only) SSR_PFA_LAUNCH(8, 1);
      else SSR_PFA_LAUNCH(8, -1);
    }
#undef SSR_PFA_LAUNCH
    SSR_LAUNCH_CHECK("k_stft_metrics_pfa");
    if (tm.on) {
      SSR_CUDA_TRY(cudaEventRecord(ev.second, st));
      tm.pending.push_back(ev);
    }
    return SSR_OK;
  }
  if (!plan->bluestein && plan->logM == 11 && !force_generic_k1()) {
    int g2 = sms * 4;  // 4 CTAs (16 warps) per SM: 128 registers, 50 KB shared memory, 64-128 TMEM columns each
    if (g2 > w.n_items) g2 = w.n_items;
    TimingState& tm = timing();
    std::pair<cudaEvent_t, cudaEvent_t> ev{nullptr, nullptr};
    if (tm.on) {
      if (!tm.pool.empty()) {
        ev = tm.pool.back();
        tm.pool.pop_back();
      } else {
        SSR_CUDA_TRY(cudaEventCreate(&ev.first));
        SSR_CUDA_TRY(cudaEventCreate(&ev.second));
      }
      SSR_CUDA_TRY(cudaEventRecord(ev.first, st));
    }
    const unsigned m3 = flags & 7u;
    const size_t smem2 = sizeof(cd) * (2048 + 256) + sizeof(float) * 2 * 1104;
    const bool store = spec_e || spec_t;
    const bool ring = plan->hop == 512;  // TMEM sample ring: the hop has to be 4 blocks of 128 samples
    const int fixed = (!store && m3 == 1u) ? 1 : ((!store && m3 == 7u) ? 7 : ((spec_e && spec_t && m3 == 7u) ? 15 : -1));
#define SSR_V2_LAUNCH(FX)                                                                           \
  do {                                                                                              \
    auto kern = ring ? k_stft_metrics_2048<FX, true> : k_stft_metrics_2048<FX, false>;              \
    SSR_CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem2)); \
    kern<<<g2, kV2Threads, smem2, st>>>(plan->dev, est, tgt, offs_dev, item_start, item_pair, w.n_items, \
                                        w.chunk, flags, partials, spec_e, spec_t, spec_off, item_start + n + 1); \
  } while (0)
    if (fixed == 1) SSR_V2_LAUNCH(1);
    else if (fixed == 7) SSR_V2_LAUNCH(7);
    else if (fixed == 15) SSR_V2_LAUNCH(15);
    else SSR_V2_LAUNCH(-1);
#undef SSR_V2_LAUNCH
    SSR_LAUNCH_CHECK("k_stft_metrics_2048");
    if (tm.on) {
      SSR_CUDA_TRY(cudaEventRecord(ev.second, st));
      tm.pending.push_back(ev);
    }
    return SSR_OK;
  }
  if (plan->bluestein)
    return dispatch_k1<true, float, float>(plan, grid, smem, st, est, tgt, offs_dev, item_start, item_pair,
                             w.n_items, w.chunk, flags, partials, spec_e, spec_t, spec_off, item_start + n + 1);
  return dispatch_k1<false, float, float>(plan, grid, smem, st, est, tgt, offs_dev, item_start, item_pair,
                            w.n_items, w.chunk, flags, partials, spec_e, spec_t, spec_off, item_start + n + 1);
}

}  // namespace ssr

using namespace ssr;

extern "C" {

int ssr_version(void) { return 200; }  // 2xx: round-2 ABI (K0, K4d, K8, explicit resampler banks, float64 targets)
const char* ssr_last_error(void) { return last_error_ref().c_str(); }
uint64_t ssr_launch_count(void) { return launch_counter().load(); }

int ssr_timing_enable(int on) {
  timing().on = on != 0;
  return SSR_OK;
}

int ssr_timing_collect(double* total_ms, int* n_launches) {
  TimingState& tm = timing();
  double tot = 0.0;
  int n = 0;
  for (auto& ev : tm.pending) {
    SSR_CUDA_TRY(cudaEventSynchronize(ev.second));
    float ms = 0.f;
    SSR_CUDA_TRY(cudaEventElapsedTime(&ms, ev.first, ev.second));
    tot += ms;
    ++n;
    tm.pool.push_back(ev);
  }
  tm.pending.clear();
  if (total_ms) *total_ms = tot;
  if (n_launches) *n_launches = n;
  return SSR_OK;
}

int ssr_stft_plan_create(ssr_stft_plan** out, int n_fft, int hop, const double* window_host) {
  if (!out) return fail(SSR_ERR_INVALID, "plan pointer is NULL");
  *out = nullptr;
  if (n_fft < 65 || n_fft > 8192) return fail(SSR_ERR_INVALID, "n_fft must be in [65, 8192]");
  if (hop < 1) return fail(SSR_ERR_INVALID, "hop must be >= 1");
  StftTables tb;
  if (!build_stft_tables(n_fft, window_host, &tb))
    return fail(SSR_ERR_INVALID, "n_fft too large for the Bluestein path (max 4096)");
  const int M = tb.M, logM = tb.logM;
  const bool blue = tb.bluestein;
  size_t o = 0;
  size_t o_tw = o;
  o = align_up(o + sizeof(cd) * (size_t)M, 256);
  size_t o_win = o;
  o = align_up(o + sizeof(double) * (size_t)n_fft, 256);
  size_t o_pos = o;
  o = align_up(o + sizeof(uint16_t) * (size_t)M, 256);
  size_t o_cw = o;
  o = align_up(o + sizeof(cd) * (size_t)n_fft, 256);
  size_t o_bf = o;
  o = align_up(o + sizeof(cd) * (size_t)M, 256);
  size_t o_cp = o;
  o = align_up(o + sizeof(cd) * (size_t)n_fft, 256);
  std::vector<unsigned char> host(o, 0);
  memcpy(host.data() + o_tw, tb.tw.data(), sizeof(cd) * (size_t)M);
  memcpy(host.data() + o_win, tb.win_half.data(), sizeof(double) * (size_t)n_fft);
  memcpy(host.data() + o_pos, tb.ppos.data(), sizeof(uint16_t) * (size_t)M);
  memcpy(host.data() + o_cw, tb.cw.data(), sizeof(cd) * (size_t)n_fft);
  memcpy(host.data() + o_bf, tb.bfilt.data(), sizeof(cd) * (size_t)M);
  memcpy(host.data() + o_cp, tb.cpost.data(), sizeof(cd) * (size_t)n_fft);
  ssr_stft_plan* p = new ssr_stft_plan();
  p->n_fft = n_fft;
  p->hop = hop;
  p->F = n_fft / 2 + 1;
  p->M = M;
  p->logM = logM;
  p->bluestein = blue ? 1 : 0;
  p->blob = nullptr;
  cudaError_t e = cudaGetDevice(&p->device);
  if (e == cudaSuccess) e = cudaMalloc(&p->blob, o);
  if (e == cudaSuccess) e = cudaMemcpy(p->blob, host.data(), o, cudaMemcpyHostToDevice);
  if (e != cudaSuccess) {
    if (p->blob) cudaFree(p->blob);
    delete p;
    return fail(SSR_ERR_CUDA, std::string("plan upload: ") + cudaGetErrorString(e));
  }
  p->pfa = 0;
  p->blob_pfa = nullptr;
  PfaTables pt;
  if (blue && build_pfa_tables(n_fft, window_host, &pt)) {
    const int R = pt.R;
    size_t q = 0;
    size_t q_tw = q;
    q = align_up(q + sizeof(cd) * 2048, 256);
    size_t q_cw = q;
    q = align_up(q + sizeof(cd) * (size_t)n_fft, 256);
    size_t q_po = q;
    q = align_up(q + sizeof(cd) * (size_t)n_fft, 256);
    size_t q_bf = q;
    q = align_up(q + sizeof(cd) * 2048, 256);
    size_t q_wr = q;
    q = align_up(q + sizeof(cd) * (size_t)(R * R), 256);
    std::vector<unsigned char> hp(q, 0);
    memcpy(hp.data() + q_tw, pt.tw.data(), sizeof(cd) * 2048);
    memcpy(hp.data() + q_cw, pt.cwin.data(), sizeof(cd) * (size_t)n_fft);
    memcpy(hp.data() + q_po, pt.post.data(), sizeof(cd) * (size_t)n_fft);
    memcpy(hp.data() + q_bf, pt.bfilt.data(), sizeof(cd) * 2048);
    memcpy(hp.data() + q_wr, pt.wr.data(), sizeof(cd) * (size_t)(R * R));
    cudaError_t e2 = cudaMalloc(&p->blob_pfa, q);
    if (e2 == cudaSuccess) e2 = cudaMemcpy(p->blob_pfa, hp.data(), q, cudaMemcpyHostToDevice);
    if (e2 != cudaSuccess) {
      if (p->blob_pfa) cudaFree(p->blob_pfa);
      cudaFree(p->blob);
      delete p;
      return fail(SSR_ERR_CUDA, std::string("plan upload (pfa): ") + cudaGetErrorString(e2));
    }
    unsigned char* dp = static_cast<unsigned char*>(p->blob_pfa);
    p->pfa = 1;
    p->pdev.n_fft = n_fft;
    p->pdev.hop = hop;
    p->pdev.F = n_fft / 2 + 1;
    p->pdev.R = R;
    p->pdev.P = pt.P;
    p->pdev.tw = reinterpret_cast<const cd*>(dp + q_tw);
    p->pdev.cwin = reinterpret_cast<const cd*>(dp + q_cw);
    p->pdev.post = reinterpret_cast<const cd*>(dp + q_po);
    p->pdev.bfilt = reinterpret_cast<const cd*>(dp + q_bf);
    p->pdev.wr = reinterpret_cast<const cd*>(dp + q_wr);
  }
  unsigned char* d = static_cast<unsigned char*>(p->blob);
  p->dev.n_fft = n_fft;
  p->dev.hop = hop;
  p->dev.F = p->F;
  p->dev.M = M;
  p->dev.tw = reinterpret_cast<const cd*>(d + o_tw);
  p->dev.win_half = reinterpret_cast<const double*>(d + o_win);
  p->dev.ppos = reinterpret_cast<const uint16_t*>(d + o_pos);
  p->dev.cw = reinterpret_cast<const cd*>(d + o_cw);
  p->dev.bfilt = reinterpret_cast<const cd*>(d + o_bf);
  p->dev.cpost = reinterpret_cast<const cd*>(d + o_cp);
  *out = p;
  return SSR_OK;
}

int ssr_stft_plan_destroy(ssr_stft_plan* plan) {
  if (!plan) return SSR_OK;
  if (plan->blob) cudaFree(plan->blob);
  if (plan->blob_pfa) cudaFree(plan->blob_pfa);
  delete plan;
  return SSR_OK;
}

int64_t ssr_stft_num_frames(const ssr_stft_plan* plan, int64_t length) {
  return plan ? (int64_t)stft_frames(length, plan->n_fft, plan->hop) : -1;
}

size_t ssr_stft_metrics_workspace_bytes(const ssr_stft_plan* plan, const int64_t* offsets_host,
                                        int n_pairs, unsigned flags) {
  if (!plan || !offsets_host || n_pairs < 1) return 0;
  WsLayout w;
  if (plan_layout(plan, offsets_host, n_pairs, flags, &w) != SSR_OK) return 0;
  return w.total;
}

static int metrics_batched_impl(const ssr_stft_plan* plan, const float* est_dev, const double* est64_dev,
                                const float* tgt_dev, const double* tgt64_dev, const int64_t* offsets_host,
                                const int64_t* offsets_dev, int n_pairs, unsigned flags, double* out_dev,
                                void* workspace_dev, size_t workspace_bytes, void* stream) {
  if (!plan || (!est_dev && !est64_dev) || (!tgt_dev && !tgt64_dev) || !offsets_host || !offsets_dev || !out_dev || n_pairs < 1)
    return fail(SSR_ERR_INVALID, "ssr_stft_metrics_batched: bad argument");
  if (int rc0 = check_offsets(offsets_host, n_pairs, "ssr_stft_metrics_batched")) return rc0;
  if (flags & ~SSR_METRIC_ALL) return fail(SSR_ERR_INVALID, "unknown metric flag");
  WsLayout w;
  int rc = plan_layout(plan, offsets_host, n_pairs, flags, &w);
  if (rc != SSR_OK) return rc;
  if (!workspace_dev || workspace_bytes < w.total)
    return fail(SSR_ERR_WORKSPACE, "workspace too small");
  if (reinterpret_cast<uintptr_t>(workspace_dev) & 15)  // K2 streams the spectrogram rows with 16-byte copies
    return fail(SSR_ERR_INVALID, "ssr_stft_metrics_batched: the workspace must be 16-byte aligned");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  unsigned char* ws = static_cast<unsigned char*>(workspace_dev);
  const long long* offs = reinterpret_cast<const long long*>(offsets_dev);
  const bool do_ssim = flags & SSR_METRIC_SSIM;
  float* spec_e = do_ssim ? reinterpret_cast<float*>(ws + w.spec_e) : nullptr;
  float* spec_t = do_ssim ? reinterpret_cast<float*>(ws + w.spec_t) : nullptr;
  rc = run_k1(plan, w, st, est_dev, tgt_dev, offs, n_pairs, flags, ws, spec_e, spec_t, est64_dev, tgt64_dev);
  if (rc != SSR_OK) return rc;
  double* ssim_part = reinterpret_cast<double*>(ws + w.ssim_part);
  if (do_ssim) {
    for (int p0 = 0; p0 < n_pairs; p0 += 32768) {
      int np = n_pairs - p0 < 32768 ? n_pairs - p0 : 32768;
      dim3 grid(w.tiles_per_pair, np);
      k_ssim<<<grid, kSsimThreads, 0, st>>>(reinterpret_cast<const float2*>(spec_e),
                                            reinterpret_cast<long long*>(ws + w.spec_off), offs, p0, plan->n_fft,
                                            plan->hop, plan->F, w.tiles_x, w.tiles_per_pair, ssim_part);
      SSR_LAUNCH_CHECK("k_ssim");
    }
  }
  k_finalize<<<(n_pairs + 3) / 4, 128, 0, st>>>(
      offs, n_pairs, plan->n_fft, plan->hop, plan->F, reinterpret_cast<int*>(ws + w.item_start),
      reinterpret_cast<double*>(ws + w.partials), ssim_part, w.tiles_per_pair, flags, out_dev);
  SSR_LAUNCH_CHECK("k_finalize");
  return SSR_OK;
}

int ssr_stft_metrics_batched(const ssr_stft_plan* plan, const float* est_dev, const float* tgt_dev,
                             const int64_t* offsets_host, const int64_t* offsets_dev, int n_pairs,
                             unsigned flags, double* out_dev, void* workspace_dev,
                             size_t workspace_bytes, void* stream) {
  if (!est_dev) return fail(SSR_ERR_INVALID, "ssr_stft_metrics_batched: bad argument");
  return metrics_batched_impl(plan, est_dev, nullptr, tgt_dev, nullptr, offsets_host, offsets_dev, n_pairs, flags,
                              out_dev, workspace_dev, workspace_bytes, stream);
}

int ssr_stft_metrics_batched_f64est(const ssr_stft_plan* plan, const double* est_dev, const float* tgt_dev,
                                    const int64_t* offsets_host, const int64_t* offsets_dev, int n_pairs,
                                    unsigned flags, double* out_dev, void* workspace_dev,
                                    size_t workspace_bytes, void* stream) {
  if (!est_dev) return fail(SSR_ERR_INVALID, "ssr_stft_metrics_batched_f64est: bad argument");
  return metrics_batched_impl(plan, nullptr, est_dev, tgt_dev, nullptr, offsets_host, offsets_dev, n_pairs, flags,
                              out_dev, workspace_dev, workspace_bytes, stream);
}

int ssr_stft_metrics_batched_f64(const ssr_stft_plan* plan, const double* est_dev, const double* tgt_dev,
                                 const int64_t* offsets_host, const int64_t* offsets_dev, int n_pairs,
                                 unsigned flags, double* out_dev, void* workspace_dev, size_t workspace_bytes,
                                 void* stream) {
  if (!est_dev || !tgt_dev) return fail(SSR_ERR_INVALID, "ssr_stft_metrics_batched_f64: bad argument");
  return metrics_batched_impl(plan, nullptr, est_dev, nullptr, tgt_dev, offsets_host, offsets_dev, n_pairs, flags,
                              out_dev, workspace_dev, workspace_bytes, stream);
}

int ssr_stft_magnitude_batched(const ssr_stft_plan* plan, const float* x_dev,
                               const int64_t* offsets_host, const int64_t* offsets_dev, int n,
                               float* spec_dev, void* workspace_dev, size_t workspace_bytes,
                               void* stream) {
  if (!plan || !x_dev || !offsets_host || !offsets_dev || !spec_dev || n < 1)
    return fail(SSR_ERR_INVALID, "ssr_stft_magnitude_batched: bad argument");
  if (int rc0 = check_offsets(offsets_host, n, "ssr_stft_magnitude_batched")) return rc0;
  WsLayout w;
  int rc = plan_layout(plan, offsets_host, n, 0, &w);
  if (rc != SSR_OK) return rc;
  if (!workspace_dev || workspace_bytes < w.total)
    return fail(SSR_ERR_WORKSPACE, "workspace too small");
  return run_k1(plan, w, static_cast<cudaStream_t>(stream), x_dev, x_dev,
                reinterpret_cast<const long long*>(offsets_dev), n, 0,
                static_cast<unsigned char*>(workspace_dev), nullptr, spec_dev);
}

}  // extern "C"

