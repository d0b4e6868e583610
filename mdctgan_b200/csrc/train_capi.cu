// train_capi.cu -- C ABI of the train-step kernels (include/mdctgan_b200.h, "train step").
#include <cstdio>
#include <cstdlib>

#include "../../include/mdctgan_b200.h"
#include "train_kernels.cuh"
#include "spectro_codec.cuh"
#include "wgrad_umma.cuh"

using namespace trk;

int mdctgan_set_error(int code, const char* fmt, ...);   // capi.cu
void mdctgan_count_launch();                              // capi.cu

#define CKT(call)                                                                  \
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

}  // extern "C"

namespace {
template <int NBLK, bool SPLIT3>
int launch_wgrad_umma(umma::WgradUmmaParams& p, cudaStream_t st) {
  using C = umma::WgCfg<NBLK, SPLIT3>;
  static bool attr_set = false;
  if (!attr_set) {
    CKT(cudaFuncSetAttribute(umma::conv_wgrad_umma_kernel<NBLK, SPLIT3>, cudaFuncAttributeMaxDynamicSharedMemorySize, C::kSmemBytes));
    attr_set = true;
  }
  p.n_tiles = (p.Cout + C::BN - 1) / C::BN;
  const int m_tiles = (p.K + umma::kBM - 1) / umma::kBM;
  const int tiles = m_tiles * p.n_tiles;
  // split the reduction over pixels while the CTAs still fit ONE wave (1 CTA per SM: the ring takes most of the shared memory);
  // every split costs one more set of output reductions
  static int cta_cap = 0;       // MDCTGAN_WGRAD_CTA_CAP: CTAs a weight-gradient launch may spread over (default: the whole device)
  if (!cta_cap) { const char* e = getenv("MDCTGAN_WGRAD_CTA_CAP"); cta_cap = e ? atoi(e) : 148; if (cta_cap < 1) cta_cap = 1; }
  int psplit = tiles >= cta_cap ? 1 : cta_cap / tiles;
  if (psplit > p.chunks) psplit = p.chunks;
  if (psplit > 65535) psplit = 65535;
  umma::conv_wgrad_umma_kernel<NBLK, SPLIT3><<<dim3(tiles, psplit), umma::kWgThreads, C::kSmemBytes, st>>>(p);
  return 0;
}

// N tile = nblk blocks of 32 output channels.  Wide tiles amortise the im2col gather (it is repeated per N tile); narrow ones give the
// grid its CTAs when the pixel reduction is too short to be split: the widest tile that still yields ~a wave of CTAs.
int pick_wgrad_nblk(int K, int Cout, int chunks) {
  const int m_tiles = (K + umma::kBM - 1) / umma::kBM;
  const int blocks = (Cout + 31) / 32;
  const int max_split = chunks / 4 > 0 ? chunks / 4 : 1;
  static int min_ctas = 0;
  if (!min_ctas) { const char* e = getenv("MDCTGAN_WGRAD_MIN_CTAS"); min_ctas = e ? atoi(e) : 120; if (min_ctas < 1) min_ctas = 1; }
  for (int nblk = blocks < 8 ? blocks : 8; nblk >= 1; --nblk) {
    const long long ctas = (long long)m_tiles * ((blocks + nblk - 1) / nblk) * max_split;
    if (ctas >= min_ctas) return nblk;
  }
  return 1;
}

template <bool SPLIT3>
int launch_wgrad_umma_n(int nblk, umma::WgradUmmaParams& p, cudaStream_t st) {
  switch (nblk) {
    case 1: return launch_wgrad_umma<1, SPLIT3>(p, st);
    case 2: return launch_wgrad_umma<2, SPLIT3>(p, st);
    case 3: return launch_wgrad_umma<3, SPLIT3>(p, st);
    case 4: return launch_wgrad_umma<4, SPLIT3>(p, st);
    case 5: return launch_wgrad_umma<5, SPLIT3>(p, st);
    case 6: return launch_wgrad_umma<6, SPLIT3>(p, st);
    case 7: return launch_wgrad_umma<7, SPLIT3>(p, st);
    default: return launch_wgrad_umma<8, SPLIT3>(p, st);
  }
}
}  // namespace

extern "C" {

int mdctgan_conv2d_wgrad_umma_supported(int Cin, int Cout) { return (Cin % 4 == 0 && Cout % 4 == 0 && Cin >= 16 && Cout >= 16) ? 1 : 0; }

/* engine: 0 = fp32 FFMA kernel, 1 = tcgen05 3xTF32 (fp32-class), 2 = tcgen05 single-pass TF32 */
int mdctgan_conv2d_wgrad(const float* x, int B, int H, int W, int Cin, const float* dy, int Ho, int Wo, int Cout, int kh, int kw, int stride,
                         int pad, int pad_mode, int transposed, const float* in_scale, const float* in_shift, int in_per_sample, int in_act,
                         const double* in_stats, double in_count, float in_eps, float* dw, int64_t s_co, int64_t s_ci, int64_t s_tap,
                         float* dbias, int engine, void* stream) {
  if (engine != 0) {
    if (!x || !dy || !dw) return mdctgan_set_error(-1, "conv2d_wgrad: NULL buffer");
    if (engine != 1 && engine != 2) return mdctgan_set_error(-1, "conv2d_wgrad: engine %d (0 fp32, 1 tcgen05 3xTF32, 2 tcgen05 TF32)", engine);
    if (B < 0 || H <= 0 || W <= 0 || Cin <= 0 || Cout <= 0 || Ho <= 0 || Wo <= 0 || kh <= 0 || kw <= 0 || stride <= 0 || pad < 0)
      return mdctgan_set_error(-1, "conv2d_wgrad: bad shape");
    if (!mdctgan_conv2d_wgrad_umma_supported(Cin, Cout))
      return mdctgan_set_error(-2, "conv2d_wgrad: the tcgen05 kernel needs Cin %% 4 == 0, Cout %% 4 == 0, both >= 16 (got %d, %d)", Cin, Cout);
    if (in_stats) return mdctgan_set_error(-1, "conv2d_wgrad: the tcgen05 kernel takes explicit in_scale / in_shift (resolve the statistics first)");
    if ((in_scale == nullptr) != (in_shift == nullptr)) return mdctgan_set_error(-1, "conv2d_wgrad: in_scale / in_shift must come together");
    if ((long long)B * Ho * Wo > 0x7fffffffLL - 64 || (long long)B * H * W * Cin > 0x7fffffffLL)
      return mdctgan_set_error(-2, "conv2d_wgrad: tensor too large for 32-bit offsets");
    if (B == 0) return 0;
    umma::WgradUmmaParams p{};
    p.x = x; p.B = B; p.H = H; p.W = W; p.Cin = Cin; p.dy = dy; p.Ho = Ho; p.Wo = Wo; p.Cout = Cout;
    p.kh = kh; p.kw = kw; p.stride = stride; p.pad = pad; p.pad_mode = pad_mode; p.transposed = transposed;
    p.in.scale = in_scale; p.in.shift = in_shift; p.in.per_sample = in_per_sample; p.in.act = in_act;
    p.dw = dw; p.s_co = s_co; p.s_ci = s_ci; p.s_tap = s_tap; p.dbias = dbias;
    p.K = kh * kw * Cin; p.P = B * Ho * Wo; p.chunks = (p.P + umma::kWgPix - 1) / umma::kWgPix;
    cudaStream_t st = (cudaStream_t)stream;
    const int nblk = pick_wgrad_nblk(p.K, Cout, p.chunks);
    const int rc = engine == 1 ? launch_wgrad_umma_n<true>(nblk, p, st) : launch_wgrad_umma_n<false>(nblk, p, st);
    if (rc) return rc;
    mdctgan_count_launch();
    CKT(cudaGetLastError());
    return 0;
  }
  if (!x || !dy || !dw) return mdctgan_set_error(-1, "conv2d_wgrad: NULL buffer");
  if (B < 0 || H <= 0 || W <= 0 || Cin <= 0 || Cout <= 0 || Ho <= 0 || Wo <= 0 || kh <= 0 || kw <= 0 || stride <= 0 || pad < 0)
    return mdctgan_set_error(-1, "conv2d_wgrad: bad shape");
  if (Cin > 1024) return mdctgan_set_error(-2, "conv2d_wgrad: Cin %d > 1024 unsupported", Cin);
  if ((in_scale == nullptr) != (in_shift == nullptr)) return mdctgan_set_error(-1, "conv2d_wgrad: in_scale / in_shift must come together");
  if (in_stats && (in_scale || !in_per_sample || in_count <= 0))
    return mdctgan_set_error(-1, "conv2d_wgrad: in_stats is the InstanceNorm2d form (per sample, count > 0, no in_scale)");
  if (B == 0) return 0;
  WgradParams p{};
  p.x = x; p.B = B; p.H = H; p.W = W; p.Cin = Cin; p.dy = dy; p.Ho = Ho; p.Wo = Wo; p.Cout = Cout;
  p.kh = kh; p.kw = kw; p.stride = stride; p.pad = pad; p.pad_mode = pad_mode; p.transposed = transposed;
  p.in.scale = in_scale; p.in.shift = in_shift; p.in.per_sample = in_per_sample; p.in.act = in_act;
  p.in.stats = in_stats; p.in.count = (float)in_count; p.in.eps = in_eps;
  p.dw = dw; p.s_co = s_co; p.s_ci = s_ci; p.s_tap = s_tap; p.dbias = dbias;
  const int K = kh * kw * Cin, HWo = Ho * Wo;
  cudaStream_t st = (cudaStream_t)stream;
  if (Cout == 1 && stride == 1 && !transposed && Cin % 4 == 0 && (kw == 7 || kw == 4) && (size_t)K * sizeof(float) <= 64 * 1024 &&
      (in_act == kActNone || in_act == kActRelu || in_act == kActLeaky) && !getenv("MDCTGAN_WGRAD_COUT1_OLD")) {
    // the heads: register reuse along the row (conv_wgrad_cout1_kernel)
    const long long units = (long long)B * Ho * ((Wo + kWgSegW - 1) / kWgSegW);
    const long long threads = units * (Cin / 4) * kh;
    const long long blocks = (threads + 255) / 256;
    if (blocks > 0x7fffffffLL) return mdctgan_set_error(-2, "conv2d_wgrad: grid too large");
    const size_t smem = (size_t)K * sizeof(float);
    if (kw == 7) {
      static bool a7 = false;
      if (!a7) { CKT(cudaFuncSetAttribute(conv_wgrad_cout1_kernel<7>, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024)); a7 = true; }
      conv_wgrad_cout1_kernel<7><<<(unsigned)blocks, 256, smem, st>>>(p);
    } else {
      static bool a4 = false;
      if (!a4) { CKT(cudaFuncSetAttribute(conv_wgrad_cout1_kernel<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024)); a4 = true; }
      conv_wgrad_cout1_kernel<4><<<(unsigned)blocks, 256, smem, st>>>(p);
    }
  } else if (Cout <= 4) {
    const int kblocks = (K + 255) / 256;
    int chunks = (1184 + kblocks * B - 1) / (kblocks * B);
    const int max_chunks = (HWo + 255) / 256;
    if (chunks > max_chunks) chunks = max_chunks;
    if (chunks < 1) chunks = 1;
    p.pix_per_chunk = ((HWo + chunks - 1) / chunks + 15) / 16 * 16;
    p.chunks_per_sample = (HWo + p.pix_per_chunk - 1) / p.pix_per_chunk;
    if ((long long)B * p.chunks_per_sample > 65535) return mdctgan_set_error(-2, "conv2d_wgrad: grid too large");
    conv_wgrad_small_cout_kernel<<<dim3(kblocks, B * p.chunks_per_sample), 256, 0, st>>>(p);
  } else {
    const int k_tiles = (K + 127) / 128;
    p.n_tiles = (Cout + 63) / 64;
    const int tiles = k_tiles * p.n_tiles;
    int chunks = (296 + tiles * B - 1) / (tiles * B);
    const int max_chunks = (HWo + 127) / 128;
    if (chunks > max_chunks) chunks = max_chunks;
    if (chunks < 1) chunks = 1;
    p.pix_per_chunk = ((HWo + chunks - 1) / chunks + 15) / 16 * 16;
    p.chunks_per_sample = (HWo + p.pix_per_chunk - 1) / p.pix_per_chunk;
    // work items = (sample, chunk); a CTA takes a contiguous run of them so that ~2 waves of CTAs cover the device
    const int items = B * p.chunks_per_sample;
    // every extra CTA along y costs a full set of output atomics (the epilogue is what bounds the small-plane layers): split the
    // reduction only while the tile grid alone leaves SMs idle
    int gy = tiles >= 148 ? 1 : (296 + tiles - 1) / tiles;
    if (gy > items) gy = items;
    if (gy < 1) gy = 1;
    if (gy > 65535) return mdctgan_set_error(-2, "conv2d_wgrad: grid too large");
    conv_wgrad_kernel<float><<<dim3(tiles, gy), 256, 0, st>>>(p);
  }
  mdctgan_count_launch();
  CKT(cudaGetLastError());
  return 0;
}

int mdctgan_norm_act_bwd(const float* x, const float* dv, float* dx, const double* stats, double count, float eps, int mode, const float* gamma,
                         const float* beta, int act, double* red, float* dgamma, float* dbeta, int B, int HW, int C, void* stream) {
  return mdctgan_norm_act_bwd_folded(x, dv, dx, stats, count, eps, mode, gamma, beta, act, red, dgamma, dbeta, B, HW, 1, C, 0, stream);
}

/* fold_pad > 0: dv is [B][H + 2 fold_pad][W + 2 fold_pad][C], the gradient of the reflection-padded view (folded while loading) */
int mdctgan_norm_act_bwd_folded(const float* x, const float* dv, float* dx, const double* stats, double count, float eps, int mode,
                                const float* gamma, const float* beta, int act, double* red, float* dgamma, float* dbeta, int B, int H, int W,
                                int C, int fold_pad, void* stream) {
  const int HW = H * W;
  if (fold_pad < 0 || (fold_pad > 0 && (fold_pad >= H || fold_pad >= W))) return mdctgan_set_error(-1, "norm_act_bwd: fold_pad %d vs %dx%d", fold_pad, H, W);
  if (!x || !dv || !dx || !stats) return mdctgan_set_error(-1, "norm_act_bwd: NULL buffer");
  if (mode != 0 && mode != 1) return mdctgan_set_error(-1, "norm_act_bwd: mode %d (0 InstanceNorm2d, 1 train-mode BatchNorm2d)", mode);
  if (C % 4 || C > 1024) return mdctgan_set_error(-2, "norm_act_bwd: C %d must be a multiple of 4, <= 1024", C);
  if (act == kActTanh) return mdctgan_set_error(-2, "norm_act_bwd: tanh after a normalisation is not a reference configuration");
  if (B == 0 || HW == 0) return 0;
  if (B > 65535) return mdctgan_set_error(-2, "norm_act_bwd: batch %d > 65535", B);
  NormBwdParams p{x, dv, dx, stats, count, eps, mode, gamma, beta, act, red, dgamma, dbeta, B, HW, C, B, fold_pad, H, W};
  cudaStream_t st = (cudaStream_t)stream;
  if (mode == 0 && C % 8 == 0 && C / 8 <= 65535 && (long long)B * (C / 8) * 256 >= (long long)HW * 2) {
    // InstanceNorm2d, one launch, no scratch: when the (sample, 8-channel) groups alone give enough CTAs for the plane size
    instnorm_bwd_fused_kernel<<<dim3(C / 8, B), 256, 0, st>>>(p);
    mdctgan_count_launch();
    CKT(cudaGetLastError());
    return 0;
  }
  if (!red) return mdctgan_set_error(-1, "norm_act_bwd: NULL reduction scratch");
  const int groups = C / 4, pstep = 256 / groups;
  int chunks = (HW + pstep - 1) / pstep;
  const int cap = (148 * 4 + B - 1) / B;
  if (chunks > cap) chunks = cap;
  norm_bwd_reduce_kernel<<<dim3(chunks, B), 256, 0, st>>>(p);
  mdctgan_count_launch();
  const size_t per_sample4 = (size_t)HW * C / 4;
  int chunks2 = (int)((per_sample4 + 255) / 256);
  const int cap2 = (148 * 8 + B - 1) / B;
  if (chunks2 > cap2) chunks2 = cap2;
  norm_bwd_apply_kernel<<<dim3(chunks2, B), 256, 0, st>>>(p);
  mdctgan_count_launch();
  CKT(cudaGetLastError());
  return 0;
}

int mdctgan_act_bwd(const float* dy, const float* y, float* g, int64_t n, int act, void* stream) {
  if (!dy || !y || !g) return mdctgan_set_error(-1, "act_bwd: NULL buffer");
  if (n <= 0) return 0;
  act_bwd_kernel<<<grid_for((size_t)n, 256), 256, 0, (cudaStream_t)stream>>>(dy, y, g, (size_t)n, act);
  mdctgan_count_launch();
  CKT(cudaGetLastError());
  return 0;
}

int mdctgan_add(const float* a, const float* b, float* y, int64_t n, void* stream) {
  if (!a || !b || !y) return mdctgan_set_error(-1, "add: NULL buffer");
  if (n <= 0) return 0;
  add_kernel<<<grid_for((size_t)n, 256), 256, 0, (cudaStream_t)stream>>>(a, b, y, (size_t)n);
  mdctgan_count_launch();
  CKT(cudaGetLastError());
  return 0;
}

int mdctgan_reflect_pad_bwd(const float* dpad, float* dx, int B, int H, int W, int C, int pad, void* stream) {
  if (!dpad || !dx) return mdctgan_set_error(-1, "reflect_pad_bwd: NULL buffer");
  if (pad < 0 || pad >= H || pad >= W) return mdctgan_set_error(-1, "reflect_pad_bwd: pad %d vs %dx%d", pad, H, W);
  const size_t total = (size_t)B * H * W * C;
  if (!total) return 0;
  reflect_fold_kernel<<<grid_for(total, 256), 256, 0, (cudaStream_t)stream>>>(dpad, dx, B, H, W, C, pad);
  mdctgan_count_launch();
  CKT(cudaGetLastError());
  return 0;
}

int mdctgan_reflect_pad_bwd_add(const float* dpad, const float* other, float* dx, int B, int H, int W, int C, int pad, void* stream) {
  if (!dpad || !other || !dx) return mdctgan_set_error(-1, "reflect_pad_bwd_add: NULL buffer");
  if (C % 4) return mdctgan_set_error(-2, "reflect_pad_bwd_add: C %d must be a multiple of 4", C);
  if (pad < 0 || pad >= H || pad >= W) return mdctgan_set_error(-1, "reflect_pad_bwd_add: pad %d vs %dx%d", pad, H, W);
  const size_t total = (size_t)B * H * W * (C / 4);
  if (!total) return 0;
  reflect_fold_add_kernel<<<grid_for(total, 256), 256, 0, (cudaStream_t)stream>>>(dpad, other, dx, B, H, W, C / 4, pad);
  mdctgan_count_launch();
  CKT(cudaGetLastError());
  return 0;
}

int mdctgan_avgpool3s2_bwd(const float* dy, float* dx, int B, int H, int W, int C, void* stream) {
  if (!dy || !dx) return mdctgan_set_error(-1, "avgpool_bwd: NULL buffer");
  const int Ho = (H + 2 - 3) / 2 + 1, Wo = (W + 2 - 3) / 2 + 1;
  const size_t total = (size_t)B * H * W * C;
  if (!total) return 0;
  avgpool3s2_bwd_kernel<<<grid_for(total, 256), 256, 0, (cudaStream_t)stream>>>(dy, dx, B, H, W, C, Ho, Wo);
  mdctgan_count_launch();
  CKT(cudaGetLastError());
  return 0;
}

int mdctgan_attention_abs_pos_bwd(const float* qkv, const float* emb_h, const float* emb_w, const float* dout, float* dqkv, float* demb_h,
                                   float* demb_w, int B, int Hh, int Ww, int heads, int d, float scale, void* stream) {
  if (!qkv || !emb_h || !emb_w || !dout || !dqkv) return mdctgan_set_error(-1, "attention_bwd: NULL buffer");
  const int L = Hh * Ww;
  if (d % 32 || d > 128) return mdctgan_set_error(-2, "attention_bwd: dim_head %d must be a multiple of 32, <= 128", d);
  if (L > 256) return mdctgan_set_error(-2, "attention_bwd: %d tokens > 256 unsupported", L);
  size_t smem = (2 * (size_t)L * (d + 1) + 2 * (size_t)L * d + 16 * (size_t)d) * sizeof(float);
  int npass = 1;
  if (smem > 227 * 1024 && (d / 32) % 2 == 0) { npass = 2; smem = (2 * (size_t)L * (d + 1) + (size_t)L * d + 16 * (size_t)d) * sizeof(float); }
  if (smem > 227 * 1024 && (d / 32) % 4 == 0) { npass = 4; smem = (2 * (size_t)L * (d + 1) + (size_t)L * d / 2 + 16 * (size_t)d) * sizeof(float); }
  if (smem > 227 * 1024) return mdctgan_set_error(-2, "attention_bwd: %zu bytes of shared memory needed (L=%d, d=%d)", smem, L, d);
  if (B == 0) return 0;
  AttnBwdParams p{qkv, emb_h, emb_w, dout, dqkv, demb_h, demb_w, B, Hh, Ww, heads, d, scale};
  cudaStream_t st = (cudaStream_t)stream;
  const int kpl = (L + 31) / 32;
#define LAUNCH_ATTN_BWD2(K, G)                                                                                         \
  do {                                                                                                                 \
    static bool attr_set = false;                                                                                      \
    if (!attr_set) {                                                                                                   \
      CKT(cudaFuncSetAttribute(attention_bwd_kernel<K, G>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));  \
      attr_set = true;                                                                                                 \
    }                                                                                                                  \
    attention_bwd_kernel<K, G><<<B * heads, 256, smem, st>>>(p);                                                       \
  } while (0)
#define LAUNCH_ATTN_BWD(K) do { if (npass == 1) LAUNCH_ATTN_BWD2(K, 1); else if (npass == 2) LAUNCH_ATTN_BWD2(K, 2); else LAUNCH_ATTN_BWD2(K, 4); } while (0)
  if (kpl <= 1) LAUNCH_ATTN_BWD(1);
  else if (kpl <= 2) LAUNCH_ATTN_BWD(2);
  else if (kpl <= 4) LAUNCH_ATTN_BWD(4);
  else LAUNCH_ATTN_BWD(8);
#undef LAUNCH_ATTN_BWD2
#undef LAUNCH_ATTN_BWD
  mdctgan_count_launch();
  CKT(cudaGetLastError());
  return 0;
}

int mdctgan_mse_const_fwd(const float* x, int64_t n, float target, double coef, double* slot, void* stream) {
  if (!x || !slot) return mdctgan_set_error(-1, "mse_const_fwd: NULL buffer");
  if (n <= 0) return 0;
  int g = grid_for((size_t)n, 256 * 8);
  mse_const_fwd_kernel<<<g, 256, 0, (cudaStream_t)stream>>>(x, (size_t)n, target, coef, slot);
  mdctgan_count_launch();
  CKT(cudaGetLastError());
  return 0;
}

int mdctgan_mse_const_bwd(const float* x, int64_t n, float target, float coef, const float* gscale, float* g, int accumulate, void* stream) {
  if (!x || !g) return mdctgan_set_error(-1, "mse_const_bwd: NULL buffer");
  if (n <= 0) return 0;
  mse_const_bwd_kernel<<<grid_for((size_t)n, 256), 256, 0, (cudaStream_t)stream>>>(x, (size_t)n, target, coef, gscale, g, accumulate);
  mdctgan_count_launch();
  CKT(cudaGetLastError());
  return 0;
}

int mdctgan_bce_const_fwd(const float* x, int64_t n, float target, double coef, double* slot, void* stream) {
  if (!x || !slot) return mdctgan_set_error(-1, "bce_const_fwd: NULL buffer");
  if (n <= 0) return 0;
  bce_const_fwd_kernel<<<grid_for((size_t)n, 256 * 8), 256, 0, (cudaStream_t)stream>>>(x, (size_t)n, target, coef, slot);
  mdctgan_count_launch();
  CKT(cudaGetLastError());
  return 0;
}

int mdctgan_bce_const_bwd(const float* x, int64_t n, float target, float coef, const float* gscale, float* g, int accumulate, void* stream) {
  if (!x || !g) return mdctgan_set_error(-1, "bce_const_bwd: NULL buffer");
  if (n <= 0) return 0;
  bce_const_bwd_kernel<<<grid_for((size_t)n, 256), 256, 0, (cudaStream_t)stream>>>(x, (size_t)n, target, coef, gscale, g, accumulate);
  mdctgan_count_launch();
  CKT(cudaGetLastError());
  return 0;
}

int mdctgan_l1_pair_fwd(const float* a, const float* b, int64_t n, double coef, double* slot, void* stream) {
  if (!a || !b || !slot) return mdctgan_set_error(-1, "l1_pair_fwd: NULL buffer");
  if (n <= 0) return 0;
  l1_pair_fwd_kernel<<<grid_for((size_t)n, 256 * 8), 256, 0, (cudaStream_t)stream>>>(a, b, (size_t)n, coef, slot);
  mdctgan_count_launch();
  CKT(cudaGetLastError());
  return 0;
}

int mdctgan_l1_pair_bwd(const float* a, const float* b, int64_t n, float coef, const float* gscale, float* g, int accumulate, void* stream) {
  if (!a || !b || !g) return mdctgan_set_error(-1, "l1_pair_bwd: NULL buffer");
  if (n <= 0) return 0;
  l1_pair_bwd_kernel<<<grid_for((size_t)n, 256), 256, 0, (cudaStream_t)stream>>>(a, b, (size_t)n, coef, gscale, g, accumulate);
  mdctgan_count_launch();
  CKT(cudaGetLastError());
  return 0;
}

/* items: n_items records of mdctgan_loss_item (host memory; copied into the kernel parameters) */
static int loss_table_of(const mdctgan_loss_item* items, int n_items, LossTable* tab, long long* max_n, const char* who) {
  if (!items || n_items <= 0) return mdctgan_set_error(-1, "%s: empty item table", who);
  if (n_items > kLossItems) return mdctgan_set_error(-2, "%s: %d items > %d", who, n_items, kLossItems);
  tab->n = n_items;
  *max_n = 0;
  for (int i = 0; i < n_items; ++i) {
    const mdctgan_loss_item& s = items[i];
    if (!s.a || (s.kind == 1 && !s.b) || s.kind < 0 || s.kind > 2 || s.n < 0) return mdctgan_set_error(-1, "%s: bad item %d", who, i);
    tab->it[i] = LossItem{s.a, s.b, s.g, (long long)s.n, s.coef, s.gscale, s.target, s.kind, s.slot};
    if (s.n > *max_n) *max_n = s.n;
  }
  return 0;
}

int mdctgan_multi_loss_fwd(const mdctgan_loss_item* items, int n_items, double* acc, int n_slots, void* stream) {
  if (!acc) return mdctgan_set_error(-1, "multi_loss_fwd: NULL accumulator");
  LossTable tab; long long max_n;
  if (int rc = loss_table_of(items, n_items, &tab, &max_n, "multi_loss_fwd")) return rc;
  for (int i = 0; i < n_items; ++i) if (items[i].slot < 0 || items[i].slot >= n_slots) return mdctgan_set_error(-1, "multi_loss_fwd: slot %d", items[i].slot);
  if (max_n == 0) return 0;
  int gx = (int)((max_n + 256 * 8 - 1) / (256 * 8)); if (gx > 64) gx = 64;
  multi_loss_fwd_kernel<<<dim3(gx, n_items), 256, 0, (cudaStream_t)stream>>>(tab, acc);
  mdctgan_count_launch();
  CKT(cudaGetLastError());
  return 0;
}

int mdctgan_multi_loss_bwd(const mdctgan_loss_item* items, int n_items, void* stream) {
  LossTable tab; long long max_n;
  if (int rc = loss_table_of(items, n_items, &tab, &max_n, "multi_loss_bwd")) return rc;
  for (int i = 0; i < n_items; ++i) if (!items[i].g) return mdctgan_set_error(-1, "multi_loss_bwd: item %d has no gradient buffer", i);
  if (max_n == 0) return 0;
  int gx = (int)((max_n + 256 * 4 - 1) / (256 * 4)); if (gx > 128) gx = 128;
  multi_loss_bwd_kernel<<<dim3(gx, n_items), 256, 0, (cudaStream_t)stream>>>(tab);
  mdctgan_count_launch();
  CKT(cudaGetLastError());
  return 0;
}

int mdctgan_f64_to_f32(const double* a, float* y, int n, void* stream) {
  if (!a || !y) return mdctgan_set_error(-1, "f64_to_f32: NULL buffer");
  if (n <= 0) return 0;
  f64_to_f32_kernel<<<(n + 127) / 128, 128, 0, (cudaStream_t)stream>>>(a, y, n);
  mdctgan_count_launch();
  CKT(cudaGetLastError());
  return 0;
}

int mdctgan_disc_input_fwd(const float* lr, int64_t lr_clip_stride, const float* s, float* out, int64_t clips, int64_t per_clip, float lo,
                           void* stream) {
  if (!lr || !s || !out) return mdctgan_set_error(-1, "disc_input_fwd: NULL buffer");
  const size_t total = (size_t)clips * per_clip;
  if (!total) return 0;
  disc_input_fwd_kernel<<<grid_for(total, 256), 256, 0, (cudaStream_t)stream>>>(lr, lr_clip_stride, s, out, clips, per_clip, lo);
  mdctgan_count_launch();
  CKT(cudaGetLastError());
  return 0;
}

int mdctgan_disc_input_bwd(const float* g, const float* s, float* ds, int64_t n, void* stream) {
  if (!g || !s || !ds) return mdctgan_set_error(-1, "disc_input_bwd: NULL buffer");
  if (n <= 0) return 0;
  disc_input_bwd_kernel<<<grid_for((size_t)n, 256), 256, 0, (cudaStream_t)stream>>>(g, s, ds, (size_t)n);
  mdctgan_count_launch();
  CKT(cudaGetLastError());
  return 0;
}

int mdctgan_adam_flat(float* p, const float* g, float* m, float* v, int64_t n, float lr, float beta1, float beta2, float eps,
                      float grad_scale, int64_t step, const int64_t* step_dev, void* stream) {
  if (!p || !g || !m || !v) return mdctgan_set_error(-1, "adam_flat: NULL buffer");
  if (!step_dev && step < 1) return mdctgan_set_error(-1, "adam_flat: step %lld (1-based)", (long long)step);
  if (((uintptr_t)p | (uintptr_t)g | (uintptr_t)m | (uintptr_t)v) & 15) return mdctgan_set_error(-1, "adam_flat: buffers must be 16-byte aligned");
  if (n <= 0) return 0;
  AdamParams a{p, g, m, v, (size_t)n, lr, beta1, beta2, eps, grad_scale, 1.f, 1.f, (const long long*)step_dev};
  if (!step_dev) {
    a.bias_c1 = (float)(1.0 - pow((double)beta1, (double)step));
    a.bias_c2_sqrt = (float)sqrt(1.0 - pow((double)beta2, (double)step));
  }
  const int grid = grid_for((size_t)n / 4 + 1, 256);
  adam_flat_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(a);
  mdctgan_count_launch();
  CKT(cudaGetLastError());
  return 0;
}

int mdctgan_pack_weights_multi(const void* descs_dev, int n_desc, int64_t total_work, void* stream) {
  if (!descs_dev) return mdctgan_set_error(-1, "pack_weights_multi: NULL descriptor table");
  if (n_desc <= 0 || total_work <= 0) return 0;
  static_assert(sizeof(PackDesc) == sizeof(mdctgan_pack_desc), "descriptor layout");
  pack_weights_multi_kernel<<<grid_for((size_t)total_work, 256 * 8), 256, 0, (cudaStream_t)stream>>>((const PackDesc*)descs_dev, n_desc,
                                                                                                      (long long)total_work);
  mdctgan_count_launch();
  CKT(cudaGetLastError());
  return 0;
}

int64_t mdctgan_segment_count(int64_t L, int seg, int ov) {
  if (L < seg) return 1;
  const int64_t nfull = (L + seg - 1) / seg;
  const int64_t padded = seg * nfull + 2 * (int64_t)ov;     // ov + L + (seg*nfull - L + ov)
  return (padded - seg) / (seg - ov) + 1;
}

int mdctgan_segment_gather(const float* audio_dev, int64_t L, float* out_dev, int64_t n_seg, int seg, int ov, void* stream) {
  if (!audio_dev || !out_dev) return mdctgan_set_error(-1, "segment_gather: NULL buffer");
  if (seg <= 0 || ov < 0 || ov >= seg || L <= 0 || n_seg <= 0) return mdctgan_set_error(-1, "segment_gather: bad shape");
  // a clip shorter than one segment is padded at the end only (audio_dataset.py:163-166)
  const int ov_eff = L < seg ? 0 : ov;
  segment_gather_kernel<<<grid_for((size_t)(n_seg * seg), 256), 256, 0, (cudaStream_t)stream>>>(audio_dev, L, out_dev, n_seg, seg, seg - ov_eff, ov_eff);
  mdctgan_count_launch();
  CKT(cudaGetLastError());
  return 0;
}

int mdctgan_segment_ola_part(const void* seg_dev, void* out_dev, int64_t n_seg, int seg, int ov, int crop_begin, int crop_end, int precision,
                             void* stream) {
  if (!seg_dev || !out_dev) return mdctgan_set_error(-1, "segment_ola: NULL buffer");
  if (seg <= 0 || ov < 0 || 2 * ov > seg || n_seg <= 0) return mdctgan_set_error(-1, "segment_ola: bad shape (0 <= 2*overlap <= segment)");
  if (crop_begin < 0 || crop_end < 0 || crop_begin > ov || crop_end > ov) return mdctgan_set_error(-1, "segment_ola: crop must lie in [0, overlap]");
  const int step = seg - ov;
  const int64_t out_len = (n_seg - 1) * step + seg - (int64_t)crop_begin - (int64_t)crop_end;
  if (out_len <= 0) return 0;
  cudaStream_t st = (cudaStream_t)stream;
  if (precision == MDCTGAN_F64) segment_ola_kernel<double><<<grid_for((size_t)out_len, 256), 256, 0, st>>>((const double*)seg_dev, (double*)out_dev, out_len, n_seg, seg, step, ov, crop_begin);
  else if (precision == MDCTGAN_F32) segment_ola_kernel<float><<<grid_for((size_t)out_len, 256), 256, 0, st>>>((const float*)seg_dev, (float*)out_dev, out_len, n_seg, seg, step, ov, crop_begin);
  else return mdctgan_set_error(-1, "segment_ola: precision %d", precision);
  mdctgan_count_launch();
  CKT(cudaGetLastError());
  return 0;
}

int mdctgan_segment_ola(const void* seg_dev, void* out_dev, int64_t n_seg, int seg, int ov, int precision, void* stream) {
  return mdctgan_segment_ola_part(seg_dev, out_dev, n_seg, seg, ov, ov, ov, precision, stream);
}

int mdctgan_add_noise(const float* lr_dev, const float* noise_dev, float* out_dev, int64_t n, double segment_length, double snr_db,
                      double* scratch3_dev, void* stream) {
  if (!lr_dev || !noise_dev || !out_dev || !scratch3_dev) return mdctgan_set_error(-1, "add_noise: NULL buffer");
  if (n < 2 || segment_length <= 0) return mdctgan_set_error(-1, "add_noise: need n >= 2 samples and segment_length > 0");
  cudaStream_t st = (cudaStream_t)stream;
  CKT(cudaMemsetAsync(scratch3_dev, 0, 3 * sizeof(double), st));
  add_noise_sums_kernel<<<grid_for((size_t)n, 256), 256, 0, st>>>(lr_dev, noise_dev, (long long)n, scratch3_dev);
  add_noise_apply_kernel<<<grid_for((size_t)n, 256), 256, 0, st>>>(lr_dev, noise_dev, out_dev, (long long)n, scratch3_dev, segment_length, snr_db);
  mdctgan_count_launch();
  mdctgan_count_launch();
  CKT(cudaGetLastError());
  return 0;
}

int mdctgan_metrics_rows(const float* hr, const float* lr, const float* sr, int64_t rows, int64_t T, double* rows_out, void* stream) {
  if (!hr || !lr || !sr || !rows_out) return mdctgan_set_error(-1, "metrics_rows: NULL buffer");
  if (rows <= 0 || T <= 0) return 0;
  if (rows > 65535) return mdctgan_set_error(-2, "metrics_rows: %lld rows > 65535", (long long)rows);
  int gx = (int)((T + 256 * 16 - 1) / (256 * 16));
  if (gx > 148 * 4) gx = 148 * 4;
  metrics_rows_kernel<<<dim3(gx, (unsigned)rows), 256, 0, (cudaStream_t)stream>>>(hr, lr, sr, T, rows_out);
  mdctgan_count_launch();
  CKT(cudaGetLastError());
  return 0;
}

int64_t mdctgan_lsd_frame_count(int64_t T, int n_fft, int hop, int center) {
  if (center) return T / hop + 1;
  return T < n_fft ? 0 : (T - n_fft) / hop + 1;
}

int mdctgan_lsd_frames(const float* hr, const float* sr, int64_t rows, int64_t T, int n_fft, int hop, const float* window_dev, int center,
                       double* acc, void* stream) {
  if (!hr || !sr || !window_dev || !acc) return mdctgan_set_error(-1, "lsd_frames: NULL buffer");
  if (n_fft != 1024 && n_fft != 512 && n_fft != 2048) return mdctgan_set_error(-2, "lsd_frames: n_fft %d (512, 1024 or 2048)", n_fft);
  if (center && T <= n_fft / 2) return mdctgan_set_error(-1, "lsd_frames: reflect padding needs T > n_fft/2");
  const int64_t frames = mdctgan_lsd_frame_count(T, n_fft, hop, center);
  if (rows <= 0 || frames <= 0) return 0;
  if (rows > 65535) return mdctgan_set_error(-2, "lsd_frames: %lld rows > 65535", (long long)rows);
  dim3 grid((unsigned)frames, (unsigned)rows);
  cudaStream_t st = (cudaStream_t)stream;
  if (n_fft == 1024) lsd_frames_kernel<1024><<<grid, 256, 0, st>>>(hr, sr, T, hop, (int)frames, window_dev, center, acc);
  else if (n_fft == 512) lsd_frames_kernel<512><<<grid, 256, 0, st>>>(hr, sr, T, hop, (int)frames, window_dev, center, acc);
  else lsd_frames_kernel<2048><<<grid, 256, 0, st>>>(hr, sr, T, hop, (int)frames, window_dev, center, acc);
  mdctgan_count_launch();
  CKT(cudaGetLastError());
  return 0;
}

int mdctgan_pack_weights_tiled(const void* descs_dev, const int64_t* tile_begin_dev, int n_desc, int64_t total_tiles, int max_taps, void* stream) {
  if (!descs_dev || !tile_begin_dev) return mdctgan_set_error(-1, "pack_weights_tiled: NULL table");
  if (n_desc <= 0 || total_tiles <= 0) return 0;
  if (max_taps <= 0 || max_taps > kPackMaxTaps) return mdctgan_set_error(-2, "pack_weights_tiled: %d taps (max %d)", max_taps, kPackMaxTaps);
  if (total_tiles > 0x7fffffffLL) return mdctgan_set_error(-2, "pack_weights_tiled: too many tiles");
  const size_t smem = (size_t)max_taps * kPackTapStride * sizeof(float);
  static size_t attr_smem = 0;
  if (smem > 48 * 1024 && smem > attr_smem) {
    CKT(cudaFuncSetAttribute(pack_weights_tiled_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(kPackMaxTaps * kPackTapStride * sizeof(float))));
    attr_smem = kPackMaxTaps * kPackTapStride * sizeof(float);
  }
  pack_weights_tiled_kernel<<<(unsigned)total_tiles, 256, smem, (cudaStream_t)stream>>>((const PackDesc*)descs_dev, (const long long*)tile_begin_dev, n_desc);
  mdctgan_count_launch();
  CKT(cudaGetLastError());
  return 0;
}

int mdctgan_plane_stats(const float* x, int B, int HW, int C, double* stats, void* stream) {
  if (!x || !stats) return mdctgan_set_error(-1, "plane_stats: NULL buffer");
  if (C % 4 || C > 1024) return mdctgan_set_error(-2, "plane_stats: C %d must be a multiple of 4, <= 1024", C);
  if (B <= 0 || HW <= 0) return 0;
  if (B > 65535) return mdctgan_set_error(-2, "plane_stats: batch %d > 65535", B);
  const int pstep = 256 / (C / 4);
  int chunks = (HW + pstep * 8 - 1) / (pstep * 8);
  const int cap = (148 * 4 + B - 1) / B;
  if (chunks > cap) chunks = cap;
  if (chunks < 1) chunks = 1;
  plane_stats_kernel<<<dim3(chunks, B), 256, 0, (cudaStream_t)stream>>>(x, HW, C, stats);
  mdctgan_count_launch();
  CKT(cudaGetLastError());
  return 0;
}

int mdctgan_upsample_nearest2x(const float* x, float* y, int B, int H, int W, int C, int backward, void* stream) {
  if (!x || !y) return mdctgan_set_error(-1, "upsample_nearest2x: NULL buffer");
  if (C % 4) return mdctgan_set_error(-2, "upsample_nearest2x: C %d must be a multiple of 4", C);
  const size_t total = (size_t)B * H * W * (C / 4) * (backward ? 1 : 4);
  if (!total) return 0;
  if (backward) upsample_nearest2x_bwd_kernel<<<grid_for(total, 256), 256, 0, (cudaStream_t)stream>>>(x, y, B, H, W, C / 4);
  else upsample_nearest2x_kernel<<<grid_for(total, 256), 256, 0, (cudaStream_t)stream>>>(x, y, B, H, W, C / 4);
  mdctgan_count_launch();
  CKT(cudaGetLastError());
  return 0;
}

static int spec_norm_of(const mdctgan_norm* n, SpecNorm* o) {
  if (!n) return mdctgan_set_error(-1, "norm is NULL");
  if (n->mode != MDCTGAN_MODE_RAW && n->mode != MDCTGAN_MODE_ARCSINH) return mdctgan_set_error(-2, "unsupported norm mode %d", n->mode);
  if (!(n->src_hi > n->src_lo) || !(n->norm_hi > n->norm_lo)) return mdctgan_set_error(-1, "empty src_range / norm_range");
  *o = SpecNorm{n->mode, (double)n->gain, (double)n->src_lo, (double)n->src_hi, (double)n->norm_lo, (double)n->norm_hi};
  return 0;
}

int mdctgan_spectro_normalize(const void* x, void* y, int64_t n, const mdctgan_norm* norm, int precision, void* stream) {
  if (!x || !y) return mdctgan_set_error(-1, "spectro_normalize: NULL buffer");
  SpecNorm p; if (int rc = spec_norm_of(norm, &p)) return rc;
  if (n <= 0) return 0;
  cudaStream_t st = (cudaStream_t)stream;
  if (precision == MDCTGAN_F64) spectro_normalize_kernel<double, double><<<grid_for((size_t)n, 256), 256, 0, st>>>((const double*)x, (double*)y, (size_t)n, p);
  else if (precision == MDCTGAN_F32) spectro_normalize_kernel<float, float><<<grid_for((size_t)n, 256), 256, 0, st>>>((const float*)x, (float*)y, (size_t)n, p);
  else return mdctgan_set_error(-1, "spectro_normalize: precision %d", precision);
  mdctgan_count_launch();
  CKT(cudaGetLastError());
  return 0;
}

int mdctgan_spectro_denormalize(const void* s, double* y, int64_t n, const mdctgan_norm* norm, int precision, void* stream) {
  if (!s || !y) return mdctgan_set_error(-1, "spectro_denormalize: NULL buffer");
  SpecNorm p; if (int rc = spec_norm_of(norm, &p)) return rc;
  if (n <= 0) return 0;
  cudaStream_t st = (cudaStream_t)stream;
  if (precision == MDCTGAN_F64) spectro_denormalize_kernel<double><<<grid_for((size_t)n, 256), 256, 0, st>>>((const double*)s, y, (size_t)n, p);
  else if (precision == MDCTGAN_F32) spectro_denormalize_kernel<float><<<grid_for((size_t)n, 256), 256, 0, st>>>((const float*)s, y, (size_t)n, p);
  else return mdctgan_set_error(-1, "spectro_denormalize: precision %d", precision);
  mdctgan_count_launch();
  CKT(cudaGetLastError());
  return 0;
}

/* ---- secondary encodings of Audio2MDCT (pix2pixHD_model.py:83-163): dB / explicit_encoding / per-sample min-max ---- */
static int codec_mode_ok(int mode) { return mode >= codec::kModeRaw && mode <= codec::kModeExplicit; }

int mdctgan_spectro_encode(const void* spec, int precision, int64_t B, int64_t plane, int mode, double gain, double alpha, double min_value,
                           double* enc, float* sign, float* minmax, void* stream) {
  if (!spec || !enc) return mdctgan_set_error(-1, "spectro_encode: NULL buffer");
  if (precision != MDCTGAN_F32 && precision != MDCTGAN_F64) return mdctgan_set_error(-1, "spectro_encode: precision %d", precision);
  if (!codec_mode_ok(mode)) return mdctgan_set_error(-2, "spectro_encode: mode %d", mode);
  if (B <= 0 || plane <= 0) return 0;
  if (B > 65535) return mdctgan_set_error(-1, "spectro_encode: batch %lld > 65535", (long long)B);
  cudaStream_t st = (cudaStream_t)stream;
  const int C = mode == codec::kModeExplicit ? 2 : 1;
  const int nmm = (int)(B * C * 2);
  if (minmax) codec::minmax_keys_init_kernel<<<(nmm + 255) / 256, 256, 0, st>>>(reinterpret_cast<unsigned*>(minmax), nmm);
  codec::EncodeParams p{spec, precision == MDCTGAN_F64, (long long)B, (long long)plane, mode, gain, alpha, min_value, enc, sign,
                        reinterpret_cast<unsigned*>(minmax)};
  long long chunks = (plane + 256 * 8 - 1) / (256 * 8);
  if (chunks > 1024) chunks = 1024;
  codec::spectro_encode_kernel<<<dim3((unsigned)chunks, (unsigned)B), 256, 0, st>>>(p);
  if (minmax) codec::minmax_keys_to_float_kernel<<<(nmm + 255) / 256, 256, 0, st>>>(reinterpret_cast<unsigned*>(minmax), nmm);
  mdctgan_count_launch();
  CKT(cudaGetLastError());
  return 0;
}

int mdctgan_spectro_affine(const double* enc, int64_t planes, int64_t plane, const float* minmax, double src_lo, double src_hi,
                           double norm_lo, double norm_hi, float* out, void* stream) {
  if (!enc || !out) return mdctgan_set_error(-1, "spectro_affine: NULL buffer");
  if (planes <= 0 || plane <= 0) return 0;
  if (planes > 65535) return mdctgan_set_error(-1, "spectro_affine: %lld planes > 65535", (long long)planes);
  if (!minmax && !(src_hi > src_lo)) return mdctgan_set_error(-1, "spectro_affine: empty src_range");
  if (!(norm_hi > norm_lo)) return mdctgan_set_error(-1, "spectro_affine: empty norm_range");
  long long chunks = (plane + 256 * 8 - 1) / (256 * 8);
  if (chunks > 1024) chunks = 1024;
  codec::AffineParams p{(long long)planes, (long long)plane, minmax, src_lo, src_hi, norm_lo, norm_hi};
  codec::spectro_affine_kernel<<<dim3((unsigned)chunks, (unsigned)planes), 256, 0, (cudaStream_t)stream>>>(enc, out, p);
  mdctgan_count_launch();
  CKT(cudaGetLastError());
  return 0;
}

int mdctgan_spectro_decode(const float* s, int64_t B, int64_t plane, int mode, double gain, double alpha, double min_value, const float* minmax,
                           double src_lo, double src_hi, double norm_lo, double norm_hi, const float* pha, double* out, void* stream) {
  if (!s || !out) return mdctgan_set_error(-1, "spectro_decode: NULL buffer");
  if (!codec_mode_ok(mode)) return mdctgan_set_error(-2, "spectro_decode: mode %d", mode);
  if (B <= 0 || plane <= 0) return 0;
  if (B > 65535) return mdctgan_set_error(-1, "spectro_decode: batch %lld > 65535", (long long)B);
  if (!(norm_hi > norm_lo)) return mdctgan_set_error(-1, "spectro_decode: empty norm_range");
  long long chunks = (plane + 256 * 8 - 1) / (256 * 8);
  if (chunks > 1024) chunks = 1024;
  codec::DecodeParams p{s, (long long)B, (long long)plane, mode, gain, alpha, min_value, minmax, src_lo, src_hi, norm_lo, norm_hi, pha, out};
  codec::spectro_decode_kernel<<<dim3((unsigned)chunks, (unsigned)B), 256, 0, (cudaStream_t)stream>>>(p);
  mdctgan_count_launch();
  CKT(cudaGetLastError());
  return 0;
}

int mdctgan_resample_fir(const float* x, int rows, int64_t L, const float* table, int K, int orig, int nw, int width, float* y,
                         int64_t target, void* stream) {
  if (!x || !table || !y) return mdctgan_set_error(-1, "resample_fir: NULL buffer");
  if (rows <= 0 || L <= 0 || target <= 0) return 0;
  if (K != 2 * width + orig || orig <= 0 || nw <= 0) return mdctgan_set_error(-1, "resample_fir: K %d != 2*width + orig (%d, %d)", K, width, orig);
  const size_t smem = (size_t)nw * K * sizeof(float);
  const int in_smem = smem <= 48 * 1024;
  resample_fir_kernel<<<grid_for((size_t)rows * (size_t)target, 256), 256, in_smem ? smem : 0, (cudaStream_t)stream>>>(
      x, L, table, K, orig, nw, width, y, target, rows, in_smem);
  mdctgan_count_launch();
  CKT(cudaGetLastError());
  return 0;
}

int mdctgan_counter_inc(int64_t* counter_dev, void* stream) {
  if (!counter_dev) return mdctgan_set_error(-1, "counter_inc: NULL buffer");
  counter_inc_kernel<<<1, 1, 0, (cudaStream_t)stream>>>((long long*)counter_dev);
  mdctgan_count_launch();
  CKT(cudaGetLastError());
  return 0;
}

}  // extern "C"
