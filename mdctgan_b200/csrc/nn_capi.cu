// nn_capi.cu -- C ABI of the network kernels (see include/mdctgan_b200.h, "network layers").
#include <cstdio>
#include <cstdlib>
#include <string>

#include "../../include/mdctgan_b200.h"
#include "nn_kernels.cuh"
#include "conv_umma.cuh"

using namespace nnk;

int mdctgan_set_error(int code, const char* fmt, ...);   // capi.cu
void mdctgan_count_launch();                              // capi.cu

#define CKN(call)                                                                  \
  do {                                                                             \
    cudaError_t e_ = (call);                                                       \
    if (e_ != cudaSuccess) return mdctgan_set_error((int)e_, "%s: %s", #call, cudaGetErrorString(e_)); \
  } while (0)

namespace {
int grid_for(size_t total, int block) {
  size_t g = (total + block - 1) / block;
  const size_t cap = 148 * 16;
  return (int)(g < cap ? (g ? g : 1) : cap);
}
}  // namespace

extern "C" {

int mdctgan_conv2d_nhwc(const float* x, int B, int H, int W, int Cin, const float* w, const float* bias, float* y, int Ho, int Wo,
                        int Cout, int kh, int kw, int stride, int pad, int pad_mode, int transposed, const float* in_scale,
                        const float* in_shift, int in_per_sample, int in_act, const double* in_stats, double in_count, float in_eps,
                        int act, double* stats, void* stream) {
  if (!x || !w || !y) return mdctgan_set_error(-1, "conv2d: NULL buffer");
  if (in_stats && (in_scale || !in_per_sample || in_count <= 0))
    return mdctgan_set_error(-1, "conv2d: in_stats is the InstanceNorm2d form (per sample, count > 0, no in_scale)");
  if (B < 0 || H <= 0 || W <= 0 || Cin <= 0 || Cout <= 0 || Ho <= 0 || Wo <= 0 || kh <= 0 || kw <= 0 || stride <= 0 || pad < 0)
    return mdctgan_set_error(-1, "conv2d: bad shape");
  if (Cin > 1024 && (in_scale || in_stats))      // the per-channel scale / shift tables live in shared memory
    return mdctgan_set_error(-2, "conv2d: Cin %d > 1024 with a deferred normalisation is unsupported", Cin);
  if (pad_mode == kPadReflect && (pad >= H || pad >= W)) return mdctgan_set_error(-1, "conv2d: reflection pad %d >= input size %dx%d", pad, H, W);
  if ((in_scale == nullptr) != (in_shift == nullptr)) return mdctgan_set_error(-1, "conv2d: in_scale / in_shift must come together");
  if (B == 0) return 0;
  ConvParams p{};
  p.x = x; p.B = B; p.H = H; p.W = W; p.Cin = Cin; p.w = w; p.bias = bias; p.y = y; p.Ho = Ho; p.Wo = Wo; p.Cout = Cout;
  p.kh = kh; p.kw = kw; p.stride = stride; p.pad = pad; p.pad_mode = pad_mode; p.transposed = transposed;
  p.in.scale = in_scale; p.in.shift = in_shift; p.in.per_sample = in_per_sample; p.in.act = in_act;
  p.in.stats = in_stats; p.in.count = (float)in_count; p.in.eps = in_eps;
  p.act = act; p.stats = stats;
  cudaStream_t st = (cudaStream_t)stream;
  const int HWo = Ho * Wo;
  if (Cout == 1 && !stats && Cin % 4 == 0 && Cin <= 512 && (in_act == kActNone || in_act == kActRelu || in_act == kActLeaky)) {
    // one output channel: LANES lanes per pixel with the thread's normalisation in registers (conv2d_cout1_kernel)
    if (kw == 7 && stride == 1 && !transposed && Cin <= 64 && Wo % 4 == 0) {        // the generator head: four pixels per 8-lane group
      const int blocks = B * ((Ho * (Wo / 4) + 31) / 32);
      if (Cin <= 32) conv2d_cout1_row4_kernel<7, 1><<<blocks, 256, 0, st>>>(p);
      else conv2d_cout1_row4_kernel<7, 2><<<blocks, 256, 0, st>>>(p);
    }
    else if (Cin <= 32) conv2d_cout1_kernel<8, 1><<<B * ((HWo + 31) / 32), 256, 0, st>>>(p);
    else if (Cin <= 64) conv2d_cout1_kernel<8, 2><<<B * ((HWo + 31) / 32), 256, 0, st>>>(p);
    else if (Cin <= 128) conv2d_cout1_kernel<32, 1><<<B * ((HWo + 7) / 8), 256, 0, st>>>(p);
    else if (Cin <= 256) conv2d_cout1_kernel<32, 2><<<B * ((HWo + 7) / 8), 256, 0, st>>>(p);
    else conv2d_cout1_kernel<32, 4><<<B * ((HWo + 7) / 8), 256, 0, st>>>(p);
  } else if (Cout <= 4 && !stats) {
    const int bps = (HWo + 31) / 32;
    const size_t smem = ((size_t)kh * kw * Cin * Cout + 2 * (size_t)Cin) * sizeof(float);
    if (smem > 200 * 1024) return mdctgan_set_error(-2, "conv2d: Cout<=4 kernel needs %zu bytes of shared memory", smem);
#define LAUNCH_SMALL(CO)                                                                                                       \
    do {                                                                                                                       \
      static bool attr_set = false;   /* once, outside any stream capture in practice (first eager call) */                    \
      if (!attr_set) { CKN(cudaFuncSetAttribute(conv2d_cout_small_kernel<CO>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024)); attr_set = true; } \
      conv2d_cout_small_kernel<CO><<<B * bps, 256, smem, st>>>(p);                                                             \
    } while (0)
    if (Cout == 1) LAUNCH_SMALL(1);
    else if (Cout == 2) LAUNCH_SMALL(2);
    else if (Cout == 3) LAUNCH_SMALL(3);
    else LAUNCH_SMALL(4);
#undef LAUNCH_SMALL
  } else {
    const int bn = Cout >= 64 ? 64 : 32;
    int bm = 64;
    if ((long long)B * ((HWo + 63) / 64) * ((Cout + bn - 1) / bn) < 148) bm = 32;
    p.tiles_per_sample = (HWo + bm - 1) / bm;
    dim3 grid(B * p.tiles_per_sample, (Cout + bn - 1) / bn);
    if (bm == 64 && bn == 64) conv2d_nhwc_kernel<64, 64><<<grid, 256, 0, st>>>(p);
    else if (bm == 64) conv2d_nhwc_kernel<64, 32><<<grid, 256, 0, st>>>(p);
    else if (bn == 64) conv2d_nhwc_kernel<32, 64><<<grid, 256, 0, st>>>(p);
    else conv2d_nhwc_kernel<32, 32><<<grid, 256, 0, st>>>(p);
  }
  mdctgan_count_launch();
  CKN(cudaGetLastError());
  return 0;
}

int mdctgan_norm_finalize(const double* stats, int B, int C, double count, float eps, int mode, const float* gamma, const float* beta,
                          float* running_mean, float* running_var, float momentum, float* scale, float* shift, void* stream) {
  if (!scale || !shift) return mdctgan_set_error(-1, "norm_finalize: NULL output");
  if (mode != 2 && !stats) return mdctgan_set_error(-1, "norm_finalize: NULL statistics");
  if (mode == 2 && (!running_mean || !running_var)) return mdctgan_set_error(-1, "norm_finalize: eval mode needs running statistics");
  if (mode < 0 || mode > 2) return mdctgan_set_error(-1, "norm_finalize: bad mode %d", mode);
  NormFinalizeParams p{stats, B, C, count, eps, mode, gamma, beta, running_mean, running_var, momentum, scale, shift};
  const int n = mode == 0 ? B * C : C;
  if (n == 0) return 0;
  norm_finalize_kernel<<<(n + 127) / 128, 128, 0, (cudaStream_t)stream>>>(p);
  mdctgan_count_launch();
  CKN(cudaGetLastError());
  return 0;
}

int mdctgan_norm_apply(const float* a, const float* a_scale, const float* a_shift, int a_per_sample, int a_act, const double* a_stats,
                       const float* b, const float* b_scale, const float* b_shift, int b_per_sample, int b_act, const double* b_stats,
                       double count, float eps, float* y, int B, int HW, int C, int act_out, void* stream) {
  if (!a || !y) return mdctgan_set_error(-1, "norm_apply: NULL buffer");
  if (C % 4 || C > 1024) return mdctgan_set_error(-2, "norm_apply: C %d must be a multiple of 4, <= 1024", C);
  if ((a_stats && (a_scale || !a_per_sample)) || (b_stats && (b_scale || !b_per_sample)) || ((a_stats || b_stats) && count <= 0))
    return mdctgan_set_error(-1, "norm_apply: *_stats is the InstanceNorm2d form (per sample, count > 0, no scale)");
  ApplyParams p{};
  p.a = a; p.na = InputNorm{a_scale, a_shift, a_per_sample, a_act, a_stats, (float)count, eps};
  p.b = b; p.nb = InputNorm{b_scale, b_shift, b_per_sample, b_act, b_stats, (float)count, eps};
  p.y = y; p.B = B; p.HW = HW; p.C = C; p.act_out = act_out;
  const size_t per_sample4 = (size_t)HW * C / 4;
  if (per_sample4 == 0 || B == 0) return 0;
  if (B > 65535) return mdctgan_set_error(-2, "norm_apply: batch %d > 65535", B);
  int chunks = (int)((per_sample4 + 255) / 256);
  const int cap = (148 * 8 + B - 1) / B;
  if (chunks > cap) chunks = cap;
  norm_apply_kernel<<<dim3(chunks, B), 256, 0, (cudaStream_t)stream>>>(p);
  mdctgan_count_launch();
  CKN(cudaGetLastError());
  return 0;
}

int mdctgan_avgpool3s2_nhwc(const float* x, float* y, int B, int H, int W, int C, void* stream) {
  if (!x || !y) return mdctgan_set_error(-1, "avgpool: NULL buffer");
  const int Ho = (H + 2 - 3) / 2 + 1, Wo = (W + 2 - 3) / 2 + 1;
  const size_t total = (size_t)B * Ho * Wo * C;
  if (total == 0) return 0;
  avgpool3s2_nhwc_kernel<<<grid_for(total, 256), 256, 0, (cudaStream_t)stream>>>(x, y, B, H, W, C, Ho, Wo);
  mdctgan_count_launch();
  CKN(cudaGetLastError());
  return 0;
}

int mdctgan_attention_abs_pos(const float* qkv, const float* emb_h, const float* emb_w, float* out, int B, int Hh, int Ww, int heads,
                               int d, float scale, double* stats, void* stream) {
  if (!qkv || !emb_h || !emb_w || !out) return mdctgan_set_error(-1, "attention: NULL buffer");
  const int L = Hh * Ww;
  if (d % 32 || d > 128) return mdctgan_set_error(-2, "attention: dim_head %d must be a multiple of 32, <= 128", d);
  if (L > 256) return mdctgan_set_error(-2, "attention: %d tokens > 256 unsupported (feature map %dx%d)", L, Hh, Ww);
  const size_t smem = ((size_t)L * (d + 1) + (size_t)L * d + 8 * (size_t)d) * sizeof(float);
  if (smem > 227 * 1024) return mdctgan_set_error(-2, "attention: %zu bytes of shared memory needed (L=%d, d=%d)", smem, L, d);
  if (B == 0) return 0;
  AttnParams p{qkv, emb_h, emb_w, out, B, Hh, Ww, heads, d, scale, stats};
  cudaStream_t st = (cudaStream_t)stream;
  const int kpl = (L + 31) / 32;
#define LAUNCH_ATTN(K)                                                                                                     \
  do {                                                                                                                     \
    static bool attr_set = false;                                                                                          \
    if (!attr_set) {                                                                                                       \
      CKN(cudaFuncSetAttribute(attention_abs_pos_kernel<K>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));     \
      attr_set = true;                                                                                                     \
    }                                                                                                                      \
    attention_abs_pos_kernel<K><<<B * heads, 256, smem, st>>>(p);                                                          \
  } while (0)
  if (kpl <= 1) LAUNCH_ATTN(1);
  else if (kpl <= 2) LAUNCH_ATTN(2);
  else if (kpl <= 4) LAUNCH_ATTN(4);
  else LAUNCH_ATTN(8);
#undef LAUNCH_ATTN
  mdctgan_count_launch();
  CKN(cudaGetLastError());
  return 0;
}

int mdctgan_residual_scale_add(const float* sr, const float* lr, int64_t lr_row_stride, float* y, int64_t rows, int nbins, int lr_bins,
                                float low_scale, void* stream) {
  if (!sr || !lr || !y) return mdctgan_set_error(-1, "residual: NULL buffer");
  if (rows < 0 || nbins <= 0 || lr_bins < 0 || lr_bins > nbins || lr_row_stride < nbins) return mdctgan_set_error(-1, "residual: bad shape");
  const size_t total = (size_t)rows * nbins;
  if (total == 0) return 0;
  residual_scale_add_kernel<<<grid_for(total, 256), 256, 0, (cudaStream_t)stream>>>(sr, lr, lr_row_stride, y, rows, nbins, lr_bins, low_scale);
  mdctgan_count_launch();
  CKN(cudaGetLastError());
  return 0;
}

int mdctgan_nchw_to_nhwc(const float* x, float* y, int B, int C, int HW, void* stream) {
  if (!x || !y) return mdctgan_set_error(-1, "layout: NULL buffer");
  const size_t total = (size_t)B * C * HW;
  if (total == 0) return 0;
  nchw_to_nhwc_kernel<<<grid_for(total, 256), 256, 0, (cudaStream_t)stream>>>(x, y, B, C, HW);
  mdctgan_count_launch();
  CKN(cudaGetLastError());
  return 0;
}

int mdctgan_nhwc_to_nchw(const float* x, float* y, int B, int C, int HW, void* stream) {
  if (!x || !y) return mdctgan_set_error(-1, "layout: NULL buffer");
  const size_t total = (size_t)B * C * HW;
  if (total == 0) return 0;
  nhwc_to_nchw_kernel<<<grid_for(total, 256), 256, 0, (cudaStream_t)stream>>>(x, y, B, C, HW);
  mdctgan_count_launch();
  CKN(cudaGetLastError());
  return 0;
}

}  // extern "C"

// ---- tcgen05 implicit-GEMM convolution (conv_umma.cuh) --------------------------------------------------
namespace {
// Largest split-K cluster: 16 CTAs (non-portable size) by default; MDCTGAN_UMMA_MAX_SPLIT = 8 keeps the portable limit (A/B switch)
int umma_split_cap() {
  static int cap = 0;
  if (!cap) {
    const char* e = getenv("MDCTGAN_UMMA_MAX_SPLIT");
    cap = e ? atoi(e) : umma::kMaxSplits;
    if (cap < 1) cap = 1;
    if (cap > umma::kMaxSplits) cap = umma::kMaxSplits;
  }
  return cap;
}

template <int BN, bool SPLIT3>
int umma_max_clusters(int splits) {   // how many clusters of `splits` CTAs fit on the device at once (cached)
  static int cache[umma::kMaxSplits + 1] = {};
  if (cache[splits]) return cache[splits];
  if (splits > 8) {      // beyond the portable cluster size
    static bool np_set = false;
    if (!np_set) { cudaFuncSetAttribute(umma::conv2d_umma_kernel<BN, SPLIT3>, cudaFuncAttributeNonPortableClusterSizeAllowed, 1); cudaGetLastError(); np_set = true; }
  }
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3(1, splits, 1);
  cfg.blockDim = dim3(umma::kThreads);
  cfg.dynamicSmemBytes = umma::Cfg<BN, SPLIT3>::kSmemBytes;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeClusterDimension;
  at[0].val.clusterDim.x = 1; at[0].val.clusterDim.y = splits; at[0].val.clusterDim.z = 1;
  cfg.attrs = at; cfg.numAttrs = 1;
  int n = 0;
  if (cudaOccupancyMaxActiveClusters(&n, umma::conv2d_umma_kernel<BN, SPLIT3>, &cfg) != cudaSuccess || n <= 0) {
    cudaGetLastError();
    n = splits > 8 ? -1 : 148 / splits;   // conservative guess; non-portable sizes the device refuses are never chosen
  }
  cache[splits] = n;
  return n;
}

template <int BN, bool SPLIT3>
int launch_umma(umma::ConvUmmaParams& p, int n_tiles, int min_kchunks, cudaStream_t st) {
  using C = umma::Cfg<BN, SPLIT3>;
  static bool attr_set = false;
  if (!attr_set) {
    CKN(cudaFuncSetAttribute(umma::conv2d_umma_kernel<BN, SPLIT3>, cudaFuncAttributeMaxDynamicSharedMemorySize, C::kSmemBytes));
    attr_set = true;
  }
  const int tiles = p.m_tiles * (p.span ? 1 : p.B) * p.classes * n_tiles;
  // split K over a cluster while the whole grid still fits the device in one wave
  int splits = 1;
  for (int s = umma_split_cap(); s >= 2; --s) {
    if (s > min_kchunks) continue;
    if (tiles <= umma_max_clusters<BN, SPLIT3>(s)) { splits = s; break; }
  }
  p.splits = splits;
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3(tiles, splits, 1);
  cfg.blockDim = dim3(umma::kThreads);
  cfg.dynamicSmemBytes = C::kSmemBytes;
  cfg.stream = st;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeClusterDimension;
  at[0].val.clusterDim.x = 1; at[0].val.clusterDim.y = splits; at[0].val.clusterDim.z = 1;
  cfg.attrs = at; cfg.numAttrs = 1;
  CKN(cudaLaunchKernelEx(&cfg, umma::conv2d_umma_kernel<BN, SPLIT3>, p));
  return 0;
}
}  // namespace

static long long* g_umma_trace = nullptr;

extern "C" {

/* debug: per-CTA phase timestamps of the following mdctgan_conv2d_umma launches ([grid.y][grid.x][16] int64), NULL = off */
int mdctgan_conv2d_umma_set_trace(long long* dev_buf) { g_umma_trace = dev_buf; return 0; }

int mdctgan_conv2d_umma_supported(int Cin, int Cout) { return (Cin % 4 == 0 && Cin <= umma::kMaxCin && Cout % 4 == 0 && Cout >= 16) ? 1 : 0; }

static int cout_padded(int Cout) { return (Cout + 31) / 32 * 32; }

int64_t mdctgan_conv2d_umma_packed_floats(int K, int Cout) {
  return (int64_t)((K + umma::kKC - 1) / umma::kKC) * 2 * cout_padded(Cout) * umma::kKC;
}

int mdctgan_conv2d_umma_pack_weight(const float* w_kn, int K, int Cout, float* out, void* stream) {
  if (!w_kn || !out) return mdctgan_set_error(-1, "umma pack: NULL buffer");
  if (K <= 0 || Cout <= 0 || Cout % 4) return mdctgan_set_error(-1, "umma pack: bad shape K=%d Cout=%d", K, Cout);
  const int kchunks = (K + umma::kKC - 1) / umma::kKC;
  const size_t total = (size_t)kchunks * umma::kKC * Cout;
  if (cout_padded(Cout) != Cout)       // the padding rows of the image are zero
    CKN(cudaMemsetAsync(out, 0, (size_t)mdctgan_conv2d_umma_packed_floats(K, Cout) * sizeof(float), (cudaStream_t)stream));
  umma::pack_weight_umma_kernel<<<grid_for(total, 256), 256, 0, (cudaStream_t)stream>>>(w_kn, out, K, Cout, cout_padded(Cout), kchunks);
  mdctgan_count_launch();
  CKN(cudaGetLastError());
  return 0;
}

int mdctgan_conv2d_umma(const float* x, int B, int H, int W, int Cin, const float* w_packed, const float* bias, float* y, int Ho, int Wo,
                        int Cout, int kh, int kw, int stride, int pad, int pad_mode, int transposed, const float* in_scale,
                        const float* in_shift, int in_per_sample, int in_act, const double* in_stats, double in_count, float in_eps,
                        int act, double* stats, int precision, void* stream) {
  if (!x || !w_packed || !y) return mdctgan_set_error(-1, "conv2d_umma: NULL buffer");
  if (B < 0 || H <= 0 || W <= 0 || Cin <= 0 || Cout <= 0 || Ho <= 0 || Wo <= 0 || kh <= 0 || kw <= 0 || stride <= 0 || pad < 0)
    return mdctgan_set_error(-1, "conv2d_umma: bad shape");
  if (!mdctgan_conv2d_umma_supported(Cin, Cout))
    return mdctgan_set_error(-2, "conv2d_umma: Cin %d (need %%4, <= %d) / Cout %d (need %%4, >= 16) unsupported", Cin, umma::kMaxCin, Cout);
  if (pad_mode == kPadReflect && (pad >= H || pad >= W)) return mdctgan_set_error(-1, "conv2d_umma: reflection pad %d >= input size %dx%d", pad, H, W);
  if ((in_scale == nullptr) != (in_shift == nullptr)) return mdctgan_set_error(-1, "conv2d_umma: in_scale / in_shift must come together");
  if (precision != 0 && precision != 1) return mdctgan_set_error(-1, "conv2d_umma: precision %d (0 = 3xTF32 fp32-class, 1 = TF32)", precision);
  if ((long long)B * Ho * Wo > 0x7fffffffLL - 128 || (long long)B * H * W * Cin > 0x7fffffffLL)
    return mdctgan_set_error(-2, "conv2d_umma: tensor too large for 32-bit offsets");
  if (B == 0) return 0;
  umma::ConvUmmaParams p{};
  p.x = x; p.B = B; p.H = H; p.W = W; p.Cin = Cin; p.wp = w_packed; p.bias = bias; p.y = y; p.Ho = Ho; p.Wo = Wo; p.Cout = Cout;
  p.kh = kh; p.kw = kw; p.stride = stride; p.pad = pad; p.pad_mode = pad_mode; p.transposed = transposed;
  p.in.scale = in_scale; p.in.shift = in_shift; p.in.per_sample = in_per_sample; p.in.act = in_act;
  p.act = act; p.stats = stats; p.trace = g_umma_trace;
  p.CoutP = cout_padded(Cout);
  const int CoutP = p.CoutP;
  p.K = kh * kw * Cin; p.kchunks = (p.K + umma::kKC - 1) / umma::kKC;
  int min_kchunks = p.kchunks;
  // ConvTranspose2d: tiles per output parity class, K loop over the live taps only (conv_umma.cuh TileGeom)
  p.classes = 1;
  int hw_class = Ho * Wo, hw_min = Ho * Wo;
  if (transposed && stride > 1 && Cin % umma::kKC == 0 && kh >= stride && kw >= stride && Ho >= stride && Wo >= stride) {
    p.classes = stride * stride;
    min_kchunks = (kh / stride) * (kw / stride) * (Cin / umma::kKC);
    hw_class = ((Ho + stride - 1) / stride) * ((Wo + stride - 1) / stride);   // the largest class (classes are ragged for odd sizes)
    hw_min = (Ho / stride) * (Wo / stride);                                   // the smallest one
  }
  else if (transposed && stride > 1)
    return mdctgan_set_error(-2, "conv2d_umma: ConvTranspose2d stride %d needs Cin %% 32 == 0, k >= stride, Ho, Wo >= stride", stride);
  if (transposed && pad_mode == kPadReflect) return mdctgan_set_error(-1, "conv2d_umma: reflection padding on a transposed convolution");
  if (in_stats) {
    if (in_scale) return mdctgan_set_error(-1, "conv2d_umma: pass either in_scale/in_shift or in_stats");
    if (!in_per_sample || in_count <= 0) return mdctgan_set_error(-1, "conv2d_umma: in_stats is the InstanceNorm2d form (per sample, count > 0)");
    p.in.stats = in_stats; p.in.count = (float)in_count; p.in.eps = in_eps;
  }
  // Planes smaller than a 128-row tile share tiles across samples (conv_umma.cuh `span`) when the per-sample normalisation
  // tables of every sample a tile can touch fit in shared memory.
  p.span = 0;
  p.m_tiles = (hw_class + umma::kBM - 1) / umma::kBM;
  if (hw_class < umma::kBM && B > 1) {
    const int per_tile = (umma::kBM + hw_min - 2) / hw_min + 1;                // samples a 128-row window can straddle
    const bool per_sample_norm = (in_scale || in_stats) && in_per_sample;
    if (!per_sample_norm || (long long)(per_tile < B ? per_tile : B) * Cin <= umma::kNormTab) {
      p.span = 1;
      p.m_tiles = (B * hw_class + umma::kBM - 1) / umma::kBM;
    }
  }
  cudaStream_t st = (cudaStream_t)stream;
  int rc;
  // Tile width: BN = 128 halves the shared-memory traffic per output (the A tile is re-read by every MMA), but at the reference's
  // shapes the grid is the constraint: take the widest tile that still gives about a wave of CTAs once K is split over a cluster.
  const long long m_ctas = (long long)p.m_tiles * (p.span ? 1 : B) * p.classes;
  const int max_split = min_kchunks < umma_split_cap() ? min_kchunks : umma_split_cap();
  int bn = 32;
  if (CoutP % 128 == 0 && (m_ctas * (CoutP / 128) >= 74 || m_ctas * (CoutP / 128) * max_split >= 96)) bn = 128;
  else if (CoutP % 64 == 0 && (m_ctas * (CoutP / 64) >= 74 || m_ctas * (CoutP / 64) * max_split >= 96)) bn = 64;
  else if (CoutP % 64 == 0 && m_ctas * (CoutP / 32) * max_split < 96) bn = 64;     // nothing fills the device: fewer, wider tiles
  if (bn == 128) rc = precision == 0 ? launch_umma<128, true>(p, CoutP / 128, min_kchunks, st) : launch_umma<128, false>(p, CoutP / 128, min_kchunks, st);
  else if (bn == 64) rc = precision == 0 ? launch_umma<64, true>(p, CoutP / 64, min_kchunks, st) : launch_umma<64, false>(p, CoutP / 64, min_kchunks, st);
  else rc = precision == 0 ? launch_umma<32, true>(p, CoutP / 32, min_kchunks, st) : launch_umma<32, false>(p, CoutP / 32, min_kchunks, st);
  if (rc) return rc;
  mdctgan_count_launch();
  CKN(cudaGetLastError());
  return 0;
}

}  // extern "C"
