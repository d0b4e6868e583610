// capi.cu -- C ABI of libmdctgan_b200.so (see include/mdctgan_b200.h for the contract).
#include <atomic>
#include <cstdarg>
#include <cstdio>
#include <climits>
#include <cstring>
#include <string>
#include <type_traits>
#include <vector>

#include "../../include/mdctgan_b200.h"
#include "mdct_kernels.cuh"
#include "mdct_plan_tables.h"

using namespace mdctk;

namespace {

thread_local std::string g_err;
std::atomic<int64_t> g_launches{0};

int fail(int code, const char* fmt, ...) {
  char buf[512];
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(buf, sizeof buf, fmt, ap);
  va_end(ap);
  g_err = buf;
  return code;
}
}  // namespace

// shared with the other translation units of the library (nn_capi.cu)
int mdctgan_set_error(int code, const char* fmt, ...) {
  char buf[512];
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(buf, sizeof buf, fmt, ap);
  va_end(ap);
  g_err = buf;
  return code;
}
void mdctgan_count_launch() { g_launches.fetch_add(1, std::memory_order_relaxed); }

namespace {
int cuda_fail(cudaError_t e, const char* what) {
  return fail((int)e, "%s: %s", what, cudaGetErrorString(e));
}
#define CK(call)                                         \
  do {                                                   \
    cudaError_t e_ = (call);                             \
    if (e_ != cudaSuccess) return cuda_fail(e_, #call);  \
  } while (0)

constexpr int kNumStreams = 3;

}  // namespace

struct mdctgan_plan {
  int device = 0;
  int num_sms = 0;
  float* tabT32 = nullptr;
  double* tabT64 = nullptr;
  float* tabW = nullptr;
  float* window = nullptr;       // the reference's synthesis window (= analysis window), fp64 flavour
  float* window_syn = nullptr;   // TDAC-exact synthesis window of the fp32 flavour (mdct_plan_tables.h)
  // occupancy-derived persistent grid sizes, indexed by kernel variant and frames-per-tile (ft = 4, 8, 12, 16)
  int grid_fwd[6][4] = {};   // [f32 raw, f32 fused, f64 raw, f64 fused, mixed raw, mixed fused][ft/4 - 1]
  int grid_inv[6][4] = {};   // [f32 raw, f64 raw, f32 fused, f64 fused, mixed raw, mixed fused][ft/4 - 1]
  // host-API scratch
  cudaStream_t streams[kNumStreams] = {nullptr, nullptr, nullptr};
  void* scratch_in[kNumStreams] = {nullptr, nullptr, nullptr};
  void* scratch_out[kNumStreams] = {nullptr, nullptr, nullptr};
  size_t scratch_in_bytes = 0, scratch_out_bytes = 0;
};

namespace {

// Persistent grid = resident CTAs per SM x SM count, for each tile height ft in {4, 8, 12, 16}.
template <typename K, typename SmemFn> int setup_kernel(K kernel, SmemFn smem_of, int num_sms, int* grid_out, int max_ft) {
  CK(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_of(max_ft)));
  for (int i = 0; i < max_ft / 4; ++i) {
    const int ft = 4 * (i + 1);
    int per_sm = 0;
    CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kernel, 32 * (i + 1), smem_of(ft)));
    if (per_sm < 1) return fail(-3, "kernel does not fit on an SM (ft %d, smem %zu)", ft, smem_of(ft));
    grid_out[i] = per_sm * num_sms;
  }
  return 0;
}

// Frames per tile: minimise (tiles per clip) x (rows staged per tile); `halo` = frames recomputed per tile.
int pick_ft(int64_t units, int halo, int max_ft) {
  int best = max_ft;
  int64_t best_cost = INT64_MAX;
  for (int ft = max_ft; ft >= 4; ft -= 4) {
    const int64_t per = ft - halo;
    const int64_t cost = ((units + per - 1) / per) * (ft + 1);
    if (cost < best_cost) { best_cost = cost; best = ft; }
  }
  return best;
}

NormParams make_norm(const mdctgan_norm* n, float* inv_a, float* inv_b) {
  NormParams p;
  p.mode = n->mode;
  p.gain = n->gain;
  // (s - smin)/(smax - smin)*(hi - lo) + lo
  const double a = ((double)n->norm_hi - (double)n->norm_lo) / ((double)n->src_hi - (double)n->src_lo);
  p.aff_a = (float)a;
  p.aff_b = (float)((double)n->norm_lo - (double)n->src_lo * a);
  p.lo = n->norm_lo;
  if (inv_a) {
    const double ia = 1.0 / a;
    *inv_a = (float)ia;
    *inv_b = (float)((double)n->src_lo - (double)n->norm_lo * ia);
  }
  return p;
}

int check_norm(const mdctgan_norm* n) {
  if (!n) return fail(-1, "norm is NULL");
  if (n->mode != MDCTGAN_MODE_RAW && n->mode != MDCTGAN_MODE_ARCSINH) return fail(-2, "unsupported norm mode %d", n->mode);
  if (n->mode == MDCTGAN_MODE_ARCSINH && !(n->gain > 0.f)) return fail(-1, "arcsinh gain must be > 0");
  if (!(n->src_hi > n->src_lo) || !(n->norm_hi > n->norm_lo)) return fail(-1, "empty src_range / norm_range");
  return 0;
}

template <typename R, int EPI, typename OutT, bool EXACT>
int launch_fwd(const mdctgan_plan* pl, FwdParams& p, const int* grid_cap, cudaStream_t st) {
  p.ft = pick_ft(p.F, 0, KCfg<R>::kMaxFt);
  p.tiles_per_clip = (p.F + p.ft - 1) / p.ft;
  p.ntiles = p.B * p.tiles_per_clip;
  if (p.ntiles == 0) return 0;
  p.tabT = std::is_same<R, float>::value ? (const void*)pl->tabT32 : (const void*)pl->tabT64;
  p.tabW = pl->tabW;
  const int grid = (int)std::min<int64_t>(p.ntiles, grid_cap[p.ft / 4 - 1]);
  mdct4_fwd_kernel<R, EPI, OutT, EXACT><<<grid, 8 * p.ft, fwd_smem_bytes<R>(p.ft), st>>>(p);
  g_launches.fetch_add(1, std::memory_order_relaxed);
  CK(cudaGetLastError());
  return 0;
}

template <typename R, typename S, typename OutT, int PRO, bool EXACT>
int launch_inv(const mdctgan_plan* pl, InvParams& p, const int* grid_cap, cudaStream_t st) {
  if (p.B == 0 || p.F < 2 || p.out_len == 0) return 0;
  const int64_t nout = (p.out_len + kHop - 1) / kHop;   // output blocks actually needed (out_length crop)
  p.ft = pick_ft(nout, 1, KCfg<R>::kMaxFtInv);
  p.tiles_per_clip = (nout + p.ft - 2) / (p.ft - 1);
  p.ntiles = p.B * p.tiles_per_clip;
  p.tabT = std::is_same<R, float>::value ? (const void*)pl->tabT32 : (const void*)pl->tabT64;
  p.window = EXACT ? pl->window : pl->window_syn;   // only the bit-faithful flavour keeps the reference's synthesis window
  const int grid = (int)std::min<int64_t>(p.ntiles, grid_cap[p.ft / 4 - 1]);
  imdct4_inv_kernel<R, S, OutT, PRO, EXACT><<<grid, 8 * p.ft, inv_smem_bytes<R, S, typename UType<R, OutT, EXACT>::type>(p.ft), st>>>(p);
  g_launches.fetch_add(1, std::memory_order_relaxed);
  CK(cudaGetLastError());
  return 0;
}

int check_common(const mdctgan_plan* plan, int64_t B, int64_t F, int precision) {
  if (!plan) return fail(-1, "plan is NULL");
  if (B < 0 || F < 0) return fail(-1, "negative size (B=%lld, F=%lld)", (long long)B, (long long)F);
  if (precision != MDCTGAN_F32 && precision != MDCTGAN_F64 && precision != MDCTGAN_MIXED)
    return fail(-1, "precision must be MDCTGAN_F32, MDCTGAN_F64 or MDCTGAN_MIXED");
  return 0;
}

}  // namespace

extern "C" {

int mdctgan_abi_version(void) { return MDCTGAN_ABI_VERSION; }
const char* mdctgan_last_error(void) { return g_err.c_str(); }
int64_t mdctgan_launch_count(void) { return g_launches.load(); }

int64_t mdctgan_frame_count(int64_t T, int64_t dim0, int hop, int win, int center) {
  const int64_t start = center ? hop : 0;
  const int64_t extra = dim0 % hop;
  const int64_t end = start + (extra ? hop - extra : 0);
  const int64_t total = start + T + end;
  return total >= win ? (total - win) / hop + 1 : 0;
}

int mdctgan_plan_create(mdctgan_plan** out, int n_fft, int hop, int win, const float* window_host) {
  if (!out || !window_host) return fail(-1, "NULL argument");
  *out = nullptr;
  if (n_fft != kWin || hop != kHop || win != kWin)
    return fail(-2, "unsupported transform size n_fft=%d hop=%d win=%d (this build: 512/256/512)", n_fft, hop, win);
  if (!window_is_symmetric(window_host, win)) return fail(-2, "window must be symmetric (w[m] == w[win-1-m])");
  mdctgan_plan* pl = new mdctgan_plan();
  CK(cudaGetDevice(&pl->device));
  CK(cudaDeviceGetAttribute(&pl->num_sms, cudaDevAttrMultiProcessorCount, pl->device));
  PlanTablesHost t;
  build_plan_tables(window_host, t);
  CK(cudaMalloc(&pl->tabT32, t.T32.size() * sizeof(float)));
  CK(cudaMalloc(&pl->tabT64, t.T64.size() * sizeof(double)));
  CK(cudaMalloc(&pl->tabW, t.W.size() * sizeof(float)));
  CK(cudaMalloc(&pl->window, win * sizeof(float)));
  CK(cudaMalloc(&pl->window_syn, win * sizeof(float)));
  std::vector<float> wsyn;
  build_synthesis_window(window_host, win, wsyn);
  CK(cudaMemcpy(pl->window_syn, wsyn.data(), win * sizeof(float), cudaMemcpyHostToDevice));
  CK(cudaMemcpy(pl->tabT32, t.T32.data(), t.T32.size() * sizeof(float), cudaMemcpyHostToDevice));
  CK(cudaMemcpy(pl->tabT64, t.T64.data(), t.T64.size() * sizeof(double), cudaMemcpyHostToDevice));
  CK(cudaMemcpy(pl->tabW, t.W.data(), t.W.size() * sizeof(float), cudaMemcpyHostToDevice));
  CK(cudaMemcpy(pl->window, window_host, win * sizeof(float), cudaMemcpyHostToDevice));
  int rc = 0;
  auto fs32 = [](int ft) { return fwd_smem_bytes<float>(ft); };
  auto fs64 = [](int ft) { return fwd_smem_bytes<double>(ft); };
  auto is32 = [](int ft) { return inv_smem_bytes<float, float>(ft); };
  auto is64 = [](int ft) { return inv_smem_bytes<double, double>(ft); };
  auto is64f = [](int ft) { return inv_smem_bytes<double, float>(ft); };
  auto is64m = [](int ft) { return inv_smem_bytes<double, float, UType<double, float, false>::type>(ft); };      // mixed flavour
  if ((rc = setup_kernel(mdct4_fwd_kernel<float, 0, float, false>, fs32, pl->num_sms, pl->grid_fwd[0], KCfg<float>::kMaxFt))) return rc;
  if ((rc = setup_kernel(mdct4_fwd_kernel<float, 1, float, false>, fs32, pl->num_sms, pl->grid_fwd[1], KCfg<float>::kMaxFt))) return rc;
  if ((rc = setup_kernel(mdct4_fwd_kernel<double, 0, double, true>, fs64, pl->num_sms, pl->grid_fwd[2], KCfg<double>::kMaxFt))) return rc;
  if ((rc = setup_kernel(mdct4_fwd_kernel<double, 1, float, true>, fs64, pl->num_sms, pl->grid_fwd[3], KCfg<double>::kMaxFt))) return rc;
  if ((rc = setup_kernel(mdct4_fwd_kernel<double, 0, float, false>, fs64, pl->num_sms, pl->grid_fwd[4], KCfg<double>::kMaxFt))) return rc;
  if ((rc = setup_kernel(mdct4_fwd_kernel<double, 1, float, false>, fs64, pl->num_sms, pl->grid_fwd[5], KCfg<double>::kMaxFt))) return rc;
  if ((rc = setup_kernel(imdct4_inv_kernel<float, float, float, 0, false>, is32, pl->num_sms, pl->grid_inv[0], KCfg<float>::kMaxFtInv))) return rc;
  if ((rc = setup_kernel(imdct4_inv_kernel<double, double, double, 0, true>, is64, pl->num_sms, pl->grid_inv[1], KCfg<double>::kMaxFtInv))) return rc;
  if ((rc = setup_kernel(imdct4_inv_kernel<float, float, float, 1, false>, is32, pl->num_sms, pl->grid_inv[2], KCfg<float>::kMaxFtInv))) return rc;
  if ((rc = setup_kernel(imdct4_inv_kernel<double, float, double, 1, true>, is64f, pl->num_sms, pl->grid_inv[3], KCfg<double>::kMaxFtInv))) return rc;
  if ((rc = setup_kernel(imdct4_inv_kernel<double, float, float, 0, false>, is64m, pl->num_sms, pl->grid_inv[4], KCfg<double>::kMaxFtInv))) return rc;
  if ((rc = setup_kernel(imdct4_inv_kernel<double, float, float, 1, false>, is64m, pl->num_sms, pl->grid_inv[5], KCfg<double>::kMaxFtInv))) return rc;
  *out = pl;
  return 0;
}

int mdctgan_plan_destroy(mdctgan_plan* pl) {
  if (!pl) return 0;
  cudaFree(pl->tabT32); cudaFree(pl->tabT64); cudaFree(pl->tabW); cudaFree(pl->window); cudaFree(pl->window_syn);
  for (int i = 0; i < kNumStreams; ++i) {
    if (pl->scratch_in[i]) cudaFree(pl->scratch_in[i]);
    if (pl->scratch_out[i]) cudaFree(pl->scratch_out[i]);
    if (pl->streams[i]) cudaStreamDestroy(pl->streams[i]);
  }
  delete pl;
  return 0;
}

int mdctgan_mdct4_forward(const mdctgan_plan* plan, const float* audio, int64_t B, int64_t T, int64_t audio_stride,
                          int64_t F, void* spec, int64_t spec_clip_stride, int precision, void* stream) {
  if (int rc = check_common(plan, B, F, precision)) return rc;
  if (B == 0 || F == 0) return 0;
  if (!audio || !spec) return fail(-1, "NULL buffer");
  if (T < 0 || audio_stride < T) return fail(-1, "audio_stride %lld < T %lld", (long long)audio_stride, (long long)T);
  if (spec_clip_stride < F * kBins || (spec_clip_stride & 1)) return fail(-1, "bad spec_clip_stride %lld", (long long)spec_clip_stride);
  if (reinterpret_cast<uintptr_t>(spec) & 15) return fail(-1, "spec pointer must be 16-byte aligned");
  FwdParams p{};
  p.audio = audio; p.audio_stride = audio_stride; p.T = T; p.B = B; p.F = F;
  p.out = spec; p.out_clip_stride = spec_clip_stride; p.out_chan_stride = 0; p.channels = 1;
  p.np = NormParams{0, 1.f, 1.f, 0.f, 0.f};
  if (precision == MDCTGAN_F32) return launch_fwd<float, 0, float, false>(plan, p, plan->grid_fwd[0], (cudaStream_t)stream);
  if (precision == MDCTGAN_MIXED) return launch_fwd<double, 0, float, false>(plan, p, plan->grid_fwd[4], (cudaStream_t)stream);
  return launch_fwd<double, 0, double, true>(plan, p, plan->grid_fwd[2], (cudaStream_t)stream);
}

int mdctgan_audio2mdct_forward(const mdctgan_plan* plan, const float* audio, int64_t B, int64_t T, int64_t audio_stride,
                               int64_t F, const mdctgan_norm* norm, float* out, int channels,
                               int64_t out_clip_stride, int64_t out_chan_stride, int precision, void* stream) {
  if (int rc = check_common(plan, B, F, precision)) return rc;
  if (int rc = check_norm(norm)) return rc;
  if (B == 0 || F == 0) return 0;
  if (!audio || !out) return fail(-1, "NULL buffer");
  if (channels != 1 && channels != 2) return fail(-1, "channels must be 1 or 2");
  if (T < 0 || audio_stride < T) return fail(-1, "audio_stride %lld < T %lld", (long long)audio_stride, (long long)T);
  if ((out_clip_stride & 1) || (out_chan_stride & 1) || out_chan_stride < (channels - 1) * F * kBins)
    return fail(-1, "bad output strides");
  if (reinterpret_cast<uintptr_t>(out) & 7) return fail(-1, "out pointer must be 8-byte aligned");
  FwdParams p{};
  p.audio = audio; p.audio_stride = audio_stride; p.T = T; p.B = B; p.F = F;
  p.out = out; p.out_clip_stride = out_clip_stride; p.out_chan_stride = out_chan_stride; p.channels = channels;
  p.np = make_norm(norm, nullptr, nullptr);
  if (precision == MDCTGAN_F32) return launch_fwd<float, 1, float, false>(plan, p, plan->grid_fwd[1], (cudaStream_t)stream);
  if (precision == MDCTGAN_MIXED) return launch_fwd<double, 1, float, false>(plan, p, plan->grid_fwd[5], (cudaStream_t)stream);
  return launch_fwd<double, 1, float, true>(plan, p, plan->grid_fwd[3], (cudaStream_t)stream);
}

int mdctgan_imdct4_inverse(const mdctgan_plan* plan, const void* spec, int64_t B, int64_t F, int64_t spec_clip_stride,
                           void* audio, int64_t audio_stride, int64_t out_len, int precision, void* stream) {
  if (int rc = check_common(plan, B, F, precision)) return rc;
  if (B == 0 || F < 2 || out_len == 0) return 0;
  if (!spec || !audio) return fail(-1, "NULL buffer");
  if (out_len < 0 || out_len > (F - 1) * kHop) return fail(-1, "out_len %lld outside [0, (F-1)*hop]", (long long)out_len);
  if (audio_stride < out_len || spec_clip_stride < F * kBins) return fail(-1, "bad strides");
  InvParams p{};
  p.spec = spec; p.spec_clip_stride = spec_clip_stride; p.B = B; p.F = F;
  p.out = audio; p.out_clip_stride = audio_stride; p.out_len = out_len;
  p.np = NormParams{0, 1.f, 1.f, 0.f, 0.f}; p.inv_a = 1.f; p.inv_b = 0.f;
  if (precision == MDCTGAN_F32) return launch_inv<float, float, float, 0, false>(plan, p, plan->grid_inv[0], (cudaStream_t)stream);
  if (precision == MDCTGAN_MIXED) return launch_inv<double, float, float, 0, false>(plan, p, plan->grid_inv[4], (cudaStream_t)stream);
  return launch_inv<double, double, double, 0, true>(plan, p, plan->grid_inv[1], (cudaStream_t)stream);
}

int mdctgan_mdct2audio_inverse(const mdctgan_plan* plan, const float* spectro, int64_t B, int64_t F, int64_t spec_clip_stride,
                               const mdctgan_norm* norm, void* audio, int64_t audio_stride, int64_t out_len,
                               int precision, void* stream) {
  if (int rc = check_common(plan, B, F, precision)) return rc;
  if (int rc = check_norm(norm)) return rc;
  if (B == 0 || F < 2 || out_len == 0) return 0;
  if (!spectro || !audio) return fail(-1, "NULL buffer");
  if (out_len < 0 || out_len > (F - 1) * kHop) return fail(-1, "out_len %lld outside [0, (F-1)*hop]", (long long)out_len);
  if (audio_stride < out_len || spec_clip_stride < F * kBins) return fail(-1, "bad strides");
  InvParams p{};
  p.spec = spectro; p.spec_clip_stride = spec_clip_stride; p.B = B; p.F = F;
  p.out = audio; p.out_clip_stride = audio_stride; p.out_len = out_len;
  p.np = make_norm(norm, &p.inv_a, &p.inv_b);
  if (precision == MDCTGAN_F32) return launch_inv<float, float, float, 1, false>(plan, p, plan->grid_inv[2], (cudaStream_t)stream);
  if (precision == MDCTGAN_MIXED) return launch_inv<double, float, float, 1, false>(plan, p, plan->grid_inv[5], (cudaStream_t)stream);
  return launch_inv<double, float, double, 1, true>(plan, p, plan->grid_inv[3], (cudaStream_t)stream);
}

// ------------------------------------------------------------------------------------------------
// Host-buffer forms: chunked, three streams, copies overlap kernels.
// ------------------------------------------------------------------------------------------------
}  // extern "C"

namespace {

int ensure_scratch(mdctgan_plan* pl, size_t in_bytes, size_t out_bytes) {
  for (int i = 0; i < kNumStreams; ++i)
    if (!pl->streams[i]) CK(cudaStreamCreateWithFlags(&pl->streams[i], cudaStreamNonBlocking));
  if (in_bytes > pl->scratch_in_bytes) {
    for (int i = 0; i < kNumStreams; ++i) {
      if (pl->scratch_in[i]) CK(cudaFree(pl->scratch_in[i]));
      pl->scratch_in[i] = nullptr;
      CK(cudaMalloc(&pl->scratch_in[i], in_bytes));
    }
    pl->scratch_in_bytes = in_bytes;
  }
  if (out_bytes > pl->scratch_out_bytes) {
    for (int i = 0; i < kNumStreams; ++i) {
      if (pl->scratch_out[i]) CK(cudaFree(pl->scratch_out[i]));
      pl->scratch_out[i] = nullptr;
      CK(cudaMalloc(&pl->scratch_out[i], out_bytes));
    }
    pl->scratch_out_bytes = out_bytes;
  }
  return 0;
}

// Streams `B` clips through `run(chunk_in_dev, chunk_out_dev, nclips, stream)`.
template <typename Run>
int stream_clips(mdctgan_plan* pl, const void* in_host, size_t in_clip_bytes, void* out_host, size_t out_clip_bytes,
                 int64_t B, Run run) {
  if (B == 0) return 0;
  const size_t target = (size_t)32 << 20;   // ~32 MiB of the larger side per chunk
  const size_t big = std::max(in_clip_bytes, out_clip_bytes);
  int64_t chunk = std::max<int64_t>(1, (int64_t)(target / std::max<size_t>(big, 1)));
  chunk = std::min<int64_t>(chunk, (B + kNumStreams - 1) / kNumStreams);
  chunk = std::max<int64_t>(chunk, 1);
  if (int rc = ensure_scratch(pl, chunk * in_clip_bytes, chunk * out_clip_bytes)) return rc;
  int s = 0;
  for (int64_t b0 = 0; b0 < B; b0 += chunk, s = (s + 1) % kNumStreams) {
    const int64_t n = std::min<int64_t>(chunk, B - b0);
    cudaStream_t st = pl->streams[s];
    CK(cudaMemcpyAsync(pl->scratch_in[s], (const char*)in_host + b0 * in_clip_bytes, n * in_clip_bytes, cudaMemcpyHostToDevice, st));
    if (int rc = run(pl->scratch_in[s], pl->scratch_out[s], n, st)) return rc;
    CK(cudaMemcpyAsync((char*)out_host + b0 * out_clip_bytes, pl->scratch_out[s], n * out_clip_bytes, cudaMemcpyDeviceToHost, st));
  }
  for (int i = 0; i < kNumStreams; ++i) CK(cudaStreamSynchronize(pl->streams[i]));
  return 0;
}

}  // namespace

extern "C" {

int mdctgan_mdct4_forward_host(mdctgan_plan* plan, const float* audio, int64_t B, int64_t T, int64_t F, void* spec, int precision) {
  if (int rc = check_common(plan, B, F, precision)) return rc;
  const size_t esz = precision == MDCTGAN_F64 ? 8 : 4;
  return stream_clips(plan, audio, (size_t)T * 4, spec, (size_t)F * kBins * esz, B,
                      [&](void* din, void* dout, int64_t n, cudaStream_t st) {
                        return mdctgan_mdct4_forward(plan, (const float*)din, n, T, T, F, dout, F * kBins, precision, st);
                      });
}

int mdctgan_audio2mdct_forward_host(mdctgan_plan* plan, const float* audio, int64_t B, int64_t T, int64_t F,
                                    const mdctgan_norm* norm, float* out, int channels, int precision) {
  if (int rc = check_common(plan, B, F, precision)) return rc;
  if (channels != 1 && channels != 2) return fail(-1, "channels must be 1 or 2");
  return stream_clips(plan, audio, (size_t)T * 4, out, (size_t)channels * F * kBins * 4, B,
                      [&](void* din, void* dout, int64_t n, cudaStream_t st) {
                        return mdctgan_audio2mdct_forward(plan, (const float*)din, n, T, T, F, norm, (float*)dout, channels,
                                                          (int64_t)channels * F * kBins, F * kBins, precision, st);
                      });
}

int mdctgan_imdct4_inverse_host(mdctgan_plan* plan, const void* spec, int64_t B, int64_t F, void* audio, int64_t out_len, int precision) {
  if (int rc = check_common(plan, B, F, precision)) return rc;
  const size_t esz = precision == MDCTGAN_F64 ? 8 : 4;
  return stream_clips(plan, spec, (size_t)F * kBins * esz, audio, (size_t)out_len * esz, B,
                      [&](void* din, void* dout, int64_t n, cudaStream_t st) {
                        return mdctgan_imdct4_inverse(plan, din, n, F, F * kBins, dout, out_len, out_len, precision, st);
                      });
}

int mdctgan_mdct2audio_inverse_host(mdctgan_plan* plan, const float* spectro, int64_t B, int64_t F, const mdctgan_norm* norm,
                                    void* audio, int64_t out_len, int precision) {
  if (int rc = check_common(plan, B, F, precision)) return rc;
  const size_t esz = precision == MDCTGAN_F64 ? 8 : 4;
  return stream_clips(plan, spectro, (size_t)F * kBins * 4, audio, (size_t)out_len * esz, B,
                      [&](void* din, void* dout, int64_t n, cudaStream_t st) {
                        return mdctgan_mdct2audio_inverse(plan, (const float*)din, n, F, F * kBins, norm, dout, out_len, out_len,
                                                          precision, st);
                      });
}

}  // extern "C"
