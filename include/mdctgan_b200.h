/* mdctgan_b200.h -- C ABI of libmdctgan_b200.so (hand-written sm_100a CUDA behind the mdctGAN hot path).
 *
 * The reference (neoncloud/mdctGAN) has no FFI: its boundary for this path is the Python surface
 *   models/mdct.py:359-489            MDCT4 / IMDCT4 (.forward)
 *   models/pix2pixHD_model.py:14-200  Audio2MDCT (to_spectro / normalize / denormalize / to_audio)
 * The entry points below are what a ctypes binding of that surface needs; each one cites the
 * reference code it replaces.  INTEGRATION.md shows the reference-side stub.
 *
 * Conventions: plain C types only.  "dev" pointers are CUDA device pointers (tensor.data_ptr()),
 * "host" pointers are host memory (pinned for full copy speed).  `stream` is a cudaStream_t passed
 * as void* (0 = legacy default stream).  Device entry points never synchronise, never allocate.
 * Return value: 0 = ok; < 0 = invalid argument / unsupported configuration; > 0 = cudaError_t.
 * mdctgan_last_error() returns a thread-local message for the last non-zero return.
 * The caller owns every buffer; a plan is immutable after creation and may be shared by streams.
 */
#ifndef MDCTGAN_B200_H_
#define MDCTGAN_B200_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define MDCTGAN_ABI_VERSION 1

/* arithmetic flavour of the transform core */
#define MDCTGAN_F32 0 /* fp32 butterflies, fp32 I/O  (fast path; ~1.2e-7 rel-L2 vs the reference)          */
#define MDCTGAN_F64 1 /* fp64 butterflies, fp64 coefficient / audio I/O (reference dtypes, mdct.py:387-390) */
#define MDCTGAN_MIXED 2 /* fp64 butterflies on fp32 I/O: the fp32 tensors of F32 with the accuracy of the fp64 core
                         * (MDCT4 -> IMDCT4 round trip 0.3 eps*peak, inside the 2-ulp bar; F32 is at 2-2.8) */

/* spectrogram encodings of Audio2MDCT.normalize (pix2pixHD_model.py:83-125) handled in-kernel */
#define MDCTGAN_MODE_RAW 0     /* --raw_mdct         :102-103 */
#define MDCTGAN_MODE_ARCSINH 1 /* --arcsinh_transform :96-100  */

typedef struct mdctgan_plan mdctgan_plan;

typedef struct mdctgan_norm {
  int32_t mode;   /* MDCTGAN_MODE_*                                                    */
  float gain;     /* opt.arcsinh_gain                       (pix2pixHD_model.py:98)    */
  float src_lo;   /* abs-norm source range, opt.src_range   (:116-119)                 */
  float src_hi;
  float norm_lo;  /* opt.norm_range                         (:120-123)                 */
  float norm_hi;
} mdctgan_norm;

int mdctgan_abi_version(void);
const char* mdctgan_last_error(void);

/* MDCT4.__init__ / IMDCT4.__init__ (models/mdct.py:365-390, :429-455).  `window_host`: win_length fp32
 * values (kbdwin, util/util.py:179-186).  Supported: n_fft 512, hop 256, win 512, symmetric window,
 * center=True; anything else returns -2 (the Python layer raises).  Tables go to the current device. */
int mdctgan_plan_create(mdctgan_plan** plan, int n_fft, int hop_length, int win_length, const float* window_host);
int mdctgan_plan_destroy(mdctgan_plan* plan);

/* Frame count of MDCT4.forward for T samples (models/mdct.py:394-407), including the reference quirk
 * that the extra right pad is derived from len(signal) = dim0 (the batch size for 2-D input). */
int64_t mdctgan_frame_count(int64_t T, int64_t dim0, int hop_length, int win_length, int center);

/* MDCT4.forward (models/mdct.py:392-425): audio fp32 [B, T] (row stride audio_stride) ->
 * coefficients [B, F, 256] (clip stride spec_clip_stride, elements), fp32 or fp64 per `precision`. */
int mdctgan_mdct4_forward(const mdctgan_plan* plan, const float* audio_dev, int64_t B, int64_t T, int64_t audio_stride,
                          int64_t F, void* spec_dev, int64_t spec_clip_stride, int precision, void* stream);

/* Audio2MDCT.to_spectro with abs_norm (pix2pixHD_model.py:32-47,81 + normalize :96-123), fused with
 * the second network channel |s|*2+norm_lo (:400-402): audio fp32 [B, T] -> fp32 [B, channels, F, 256]
 * (channels 1 or 2; strides in elements).  precision F64 = fp64 core + library asinh, rounded once. */
int mdctgan_audio2mdct_forward(const mdctgan_plan* plan, const float* audio_dev, int64_t B, int64_t T, int64_t audio_stride,
                               int64_t F, const mdctgan_norm* norm, float* out_dev, int channels,
                               int64_t out_clip_stride, int64_t out_chan_stride, int precision, void* stream);

/* IMDCT4.forward (models/mdct.py:457-489): coefficients [B, F, 256] -> audio [B, out_len],
 * out_len <= (F-1)*256 (the out_length crop, :486-488).  fp32 or fp64 in AND out per `precision`. */
int mdctgan_imdct4_inverse(const mdctgan_plan* plan, const void* spec_dev, int64_t B, int64_t F, int64_t spec_clip_stride,
                           void* audio_dev, int64_t audio_stride, int64_t out_len, int precision, void* stream);

/* Audio2MDCT.to_audio, abs_norm arcsinh/raw branches (pix2pixHD_model.py:127-135,139-163):
 * normalised fp32 spectrogram [B, F, 256] -> audio; F32: fp32 audio, F64: fp64 audio (reference dtype). */
int mdctgan_mdct2audio_inverse(const mdctgan_plan* plan, const float* spectro_dev, int64_t B, int64_t F, int64_t spec_clip_stride,
                               const mdctgan_norm* norm, void* audio_dev, int64_t audio_stride, int64_t out_len,
                               int precision, void* stream);

/* Audio2MDCT.normalize / denormalize on an existing spectrogram (pix2pixHD_model.py:83-137; arcsinh / raw with abs_norm), computed
 * in fp64 like the reference.  normalize: x, y fp32 or fp64 per `precision`; denormalize: s fp32 or fp64 per `precision`, y fp64. */
int mdctgan_spectro_normalize(const void* x, void* y, int64_t n, const mdctgan_norm* norm, int precision, void* stream);
int mdctgan_spectro_denormalize(const void* s, double* y, int64_t n, const mdctgan_norm* norm, int precision, void* stream);

/* Secondary encodings of Audio2MDCT (pix2pixHD_model.py:83-163), element-wise fp64 around the raw MDCT4 / IMDCT4 launches.
 * mode: MDCTGAN_MODE_RAW / _ARCSINH / _DB (default options, :104-106) / _EXPLICIT (--explicit_encoding, :84-95: 2 channels).
 * encode: raw coefficients [B][plane] (fp32 or fp64 per `precision`) -> enc fp64 [B][C][plane]; sign (nullable) = torch.sign(spectro)
 *         as fp32 (:36); minmax (nullable) = fp32 [B][C][2] (min, max) of every plane (the no --abs_norm branch, :111-114).
 * affine: (enc - min) / (max - min) * (norm_hi - norm_lo) + norm_lo -> fp32 (:116-123); minmax NULL = the abs-norm src_range.
 * decode: denormalize (:127-137) + channel recombination / phase product of to_audio (:142-157) -> raw coefficients fp64 [B][plane];
 *         pha (nullable, dB mode) = fp32 [B][plane] multiplier. */
#define MDCTGAN_MODE_DB 2
#define MDCTGAN_MODE_EXPLICIT 3
int mdctgan_spectro_encode(const void* spec, int precision, int64_t B, int64_t plane, int mode, double gain, double alpha, double min_value,
                           double* enc, float* sign, float* minmax, void* stream);
int mdctgan_spectro_affine(const double* enc, int64_t planes, int64_t plane, const float* minmax, double src_lo, double src_hi,
                           double norm_lo, double norm_hi, float* out, void* stream);
int mdctgan_spectro_decode(const float* s, int64_t B, int64_t plane, int mode, double gain, double alpha, double min_value, const float* minmax,
                           double src_lo, double src_hi, double norm_lo, double norm_hi, const float* pha, double* out, void* stream);

/* Host-buffer forms of the four calls above (the end-to-end path a non-torch caller uses): inputs and
 * outputs are HOST arrays, densely packed; clips are streamed host->device->host in chunks over three
 * CUDA streams so copies overlap the kernels.  They synchronise before returning.  Scratch device
 * memory is owned by the plan and grown on first use (not thread-safe per plan). */
int mdctgan_mdct4_forward_host(mdctgan_plan* plan, const float* audio_host, int64_t B, int64_t T, int64_t F,
                               void* spec_host, int precision);
int mdctgan_audio2mdct_forward_host(mdctgan_plan* plan, const float* audio_host, int64_t B, int64_t T, int64_t F,
                                    const mdctgan_norm* norm, float* out_host, int channels, int precision);
int mdctgan_imdct4_inverse_host(mdctgan_plan* plan, const void* spec_host, int64_t B, int64_t F, void* audio_host,
                                int64_t out_len, int precision);
int mdctgan_mdct2audio_inverse_host(mdctgan_plan* plan, const float* spectro_host, int64_t B, int64_t F,
                                    const mdctgan_norm* norm, void* audio_host, int64_t out_len, int precision);

/* ------------------------------------------------------------------------------------------------
 * Network layers (reference: models/networks.py).  Activations are NHWC fp32 device buffers; weights are
 * packed [kh*kw*Cin][Cout] (Conv2d weight [Cout,Cin,kh,kw] permuted (2,3,1,0); ConvTranspose2d weight
 * [Cin,Cout,kh,kw] permuted (2,3,0,1)).  Normalisation layers are split between the producing
 * convolution (statistics, `stats` = [B][Cout][2] doubles (sum, sumsq), accumulated atomically: zero it
 * first) and the consumer (v = act(x*scale + shift), `in_scale/in_shift` = [B][Cin] when
 * in_per_sample, else [Cin]).  act codes: 0 none, 1 ReLU, 2 LeakyReLU(0.2), 3 tanh.
 * pad_mode: 0 zeros, 1 reflection (nn.ReflectionPad2d in front of an unpadded conv, networks.py:308,430).
 */
/* A consumer may also take the producer's InstanceNorm2d(affine=False) statistics raw: `in_stats` = the producer's
 * [B][Cin][2] (sum, sumsq), `in_count` = elements per plane, `in_eps`; it then derives scale = rstd and
 * shift = -mean*rstd itself and no norm_finalize launch is needed (in_scale / in_shift must be NULL). */
/* nn.Conv2d / nn.ConvTranspose2d forward (networks.py:308-352, :387-417, :649-670): direct fp32 FFMA kernels --
 * every shape, incl. the Cin = 2 stem and the Cout = 1 heads. */
int mdctgan_conv2d_nhwc(const float* x, int B, int H, int W, int Cin, const float* w, const float* bias, float* y, int Ho, int Wo,
                        int Cout, int kh, int kw, int stride, int pad, int pad_mode, int transposed, const float* in_scale,
                        const float* in_shift, int in_per_sample, int in_act, const double* in_stats, double in_count, float in_eps,
                        int act, double* stats, void* stream);
/* The same layer on the 5th-generation tensor cores (tcgen05.mma kind::tf32, TMEM accumulator, TMA bulk-copied
 * weights, cluster split-K; csrc/conv_umma.cuh) for Cin % 4 == 0, Cin <= 1024, Cout % 32 == 0
 * (mdctgan_conv2d_umma_supported).  `w_packed` is the image written by mdctgan_conv2d_umma_pack_weight from the
 * [kh*kw*Cin][Cout] weights (mdctgan_conv2d_umma_packed_floats floats).  precision 0: 3xTF32 split, fp32-class
 * results; 1: single TF32 pass.  With in_stats and B > 1 the (per parity class) plane must be a multiple of 128
 * pixels. */
int mdctgan_conv2d_umma_supported(int Cin, int Cout);
/* debug aid: subsequent mdctgan_conv2d_umma launches write per-CTA phase timestamps into dev_buf
 * ([grid.y * grid.x][16] int64; slot 0 globaltimer ns, slots 1.. clock64 at the phase boundaries); NULL = off */
int mdctgan_conv2d_umma_set_trace(long long* dev_buf);
int64_t mdctgan_conv2d_umma_packed_floats(int K, int Cout);
int mdctgan_conv2d_umma_pack_weight(const float* w_kn, int K, int Cout, float* out, void* stream);
int mdctgan_conv2d_umma(const float* x, int B, int H, int W, int Cin, const float* w_packed, const float* bias, float* y, int Ho, int Wo,
                        int Cout, int kh, int kw, int stride, int pad, int pad_mode, int transposed, const float* in_scale,
                        const float* in_shift, int in_per_sample, int in_act, const double* in_stats, double in_count, float in_eps,
                        int act, double* stats, int precision, void* stream);
/* InstanceNorm2d(affine=False) (mode 0, networks.py:26) / BatchNorm2d train (1) / eval (2) statistics ->
 * per-(sample,)channel scale & shift; BatchNorm also updates its running buffers in train mode. */
int mdctgan_norm_finalize(const double* stats, int B, int C, double count, float eps, int mode, const float* gamma, const float* beta,
                          float* running_mean, float* running_var, float momentum, float* scale, float* shift, void* stream);
/* y = act_out( act_a(a*sa+ta) [+ act_b(b*sb+tb)] ): ResnetBlock `x + conv_block(x)` (networks.py:461-463),
 * LocalEnhancer branch sum (:266-267), BottleStack shortcut. */
int mdctgan_norm_apply(const float* a, const float* a_scale, const float* a_shift, int a_per_sample, int a_act, const double* a_stats,
                       const float* b, const float* b_scale, const float* b_shift, int b_per_sample, int b_act, const double* b_stats,
                       double count, float eps, float* y, int B, int HW, int C, int act_out, void* stream);
/* nn.AvgPool2d(3, stride=2, padding=1, count_include_pad=False) (networks.py:249-250, :525-526). */
int mdctgan_avgpool3s2_nhwc(const float* x, float* y, int B, int H, int W, int C, void* stream);
/* BoTNet attention with absolute position embedding (bottleneck_transformer_pytorch==0.1.4 `Attention`,
 * called from networks.py:342-344): qkv NHWC [B, Hh*Ww, 3*heads*d] -> out [B, Hh*Ww, heads*d];
 * `stats` (nullable) receives the per-channel (sum, sumsq) for the BatchNorm2d that follows. */
int mdctgan_attention_abs_pos(const float* qkv, const float* emb_h, const float* emb_w, float* out, int B, int Hh, int Ww, int heads,
                               int d, float scale, double* stats, void* stream);
/* --fit_residual at inference (pix2pixHD_model.py:631-635): y = sr with the first lr_bins columns scaled by
 * low_scale, plus lr; sr, y: [rows, nbins] dense, lr: rows of nbins with row stride lr_row_stride. */
int mdctgan_residual_scale_add(const float* sr, const float* lr, int64_t lr_row_stride, float* y, int64_t rows, int nbins, int lr_bins,
                                float low_scale, void* stream);
/* layout changes at the network boundary (the reference's modules take and return NCHW) */
int mdctgan_nchw_to_nhwc(const float* x, float* y, int B, int C, int HW, void* stream);
int mdctgan_nhwc_to_nchw(const float* x, float* y, int B, int C, int HW, void* stream);

/* ------------------------------------------------------------------------------------------------
 * Train step (reference: what torch autograd derives for models/pix2pixHD_model.py:416-451 + train.py:175-202).
 * dgrad has no entry of its own: the input gradient of nn.Conv2d is mdctgan_conv2d_nhwc / _umma with transposed = 1
 * on the output gradient with the weights packed [kh*kw*Cout][Cin]; of nn.ConvTranspose2d the plain convolution.
 */
/* dW (+= , float reductions) and dbias (+=) of nn.Conv2d / nn.ConvTranspose2d: x = the layer's raw input with its deferred
 * normalisation / activation (as in mdctgan_conv2d_nhwc), dy = gradient of the raw convolution output [B,Ho,Wo,Cout];
 * dW[co][ci][tap] lives at dw[co*s_co + ci*s_ci + tap*s_tap] (the parameter's own layout, e.g. a slice of the flat bucket).
 * engine: 0 = fp32 FFMA kernel (any shape); 1 = tcgen05 MN-major implicit GEMM, 3xTF32 (fp32-class); 2 = tcgen05, single TF32
 * pass.  Engines 1 / 2 need mdctgan_conv2d_wgrad_umma_supported(Cin, Cout) and explicit in_scale / in_shift (no in_stats). */
int mdctgan_conv2d_wgrad_umma_supported(int Cin, int Cout);
int mdctgan_conv2d_wgrad(const float* x, int B, int H, int W, int Cin, const float* dy, int Ho, int Wo, int Cout, int kh, int kw, int stride,
                         int pad, int pad_mode, int transposed, const float* in_scale, const float* in_shift, int in_per_sample, int in_act,
                         const double* in_stats, double in_count, float in_eps, float* dw, int64_t s_co, int64_t s_ci, int64_t s_tap,
                         float* dbias, int engine, void* stream);
/* Backward of v = act(norm(x)): mode 0 InstanceNorm2d(affine=False) (networks.py:26), 1 train-mode BatchNorm2d (BottleStack).
 * stats = the forward (sum, sumsq) [B][C][2]; red = zeroed [B][C][2] scratch; dgamma / dbeta accumulated (mode 1, nullable). */
int mdctgan_norm_act_bwd(const float* x, const float* dv, float* dx, const double* stats, double count, float eps, int mode, const float* gamma,
                         const float* beta, int act, double* red, float* dgamma, float* dbeta, int B, int HW, int C, void* stream);
/* The same with the gradient arriving in the reflection-padded geometry of the consumer ([B][H + 2 fold_pad][W + 2 fold_pad][C], the raw
 * output of the input-gradient convolution): nn.ReflectionPad2d's backward is applied while loading.  fold_pad = 0: plain. */
int mdctgan_norm_act_bwd_folded(const float* x, const float* dv, float* dx, const double* stats, double count, float eps, int mode,
                                const float* gamma, const float* beta, int act, double* red, float* dgamma, float* dbeta, int B, int H, int W,
                                int C, int fold_pad, void* stream);
/* nn.ReflectionPad2d backward fused with the accumulation into an existing gradient: dx = fold(dpad) + other (C % 4 == 0) */
int mdctgan_reflect_pad_bwd_add(const float* dpad, const float* other, float* dx, int B, int H, int W, int C, int pad, void* stream);
/* g = dy * act'(y) from the activated value y (epilogue LeakyReLU / tanh, plain ReLU views) */
int mdctgan_act_bwd(const float* dy, const float* y, float* g, int64_t n, int act, void* stream);
int mdctgan_add(const float* a, const float* b, float* y, int64_t n, void* stream);
/* (sum, sumsq) per (sample, channel) of a materialised NHWC tensor -> stats [B][C][2] (+=): the branch sums of ConvResBlock /
 * InterpolateUpsample (networks.py:387-417) are followed by an InstanceNorm2d */
int mdctgan_plane_stats(const float* x, int B, int HW, int C, double* stats, void* stream);
/* F.interpolate(scale_factor=2.0, mode="nearest") (networks.py:396): x [B,H,W,C] -> y [B,2H,2W,C]; backward = 1: x = dy [B,2H,2W,C] ->
 * y = dx [B,H,W,C] (H, W always the SMALL size) */
int mdctgan_upsample_nearest2x(const float* x, float* y, int B, int H, int W, int C, int backward, void* stream);
/* nn.ReflectionPad2d backward: dpad [B,H+2p,W+2p,C] -> dx [B,H,W,C] */
int mdctgan_reflect_pad_bwd(const float* dpad, float* dx, int B, int H, int W, int C, int pad, void* stream);
int mdctgan_avgpool3s2_bwd(const float* dy, float* dx, int B, int H, int W, int C, void* stream);
/* backward of mdctgan_attention_abs_pos: dqkv [B,L,3*heads*d] written, demb_h / demb_w accumulated (nullable) */
int mdctgan_attention_abs_pos_bwd(const float* qkv, const float* emb_h, const float* emb_w, const float* dout, float* dqkv, float* demb_h,
                                   float* demb_w, int B, int Hh, int Ww, int heads, int d, float scale, void* stream);
/* GANLoss, LSGAN branch (networks.py:127-137): *slot += coef * sum((x - target)^2); g (+)= 2*coef*(*gscale)*(x - target) */
int mdctgan_mse_const_fwd(const float* x, int64_t n, float target, double coef, double* slot, void* stream);
int mdctgan_mse_const_bwd(const float* x, int64_t n, float target, float coef, const float* gscale, float* g, int accumulate, void* stream);
/* The same pair for nn.BCELoss on sigmoid outputs (GANLoss with --no_lsgan, networks.py:107-108; torch's -100 log clamp and 1e-12
 * denominator clamp). */
/* Every loss term of a step in ONE launch (and every gradient seed in another): up to 24 items, copied by value into the kernel
 * parameters.  kind 0: MSE of `a` against the constant `target`; 1: L1 between `a` and `b`; 2: BCE of probabilities `a` against `target`.
 * fwd: acc[slot] += coef * sum (acc: fp64 device array of n_slots, zeroed by the caller).  bwd: g = coef * (*gscale) * d(sum)/da. */
typedef struct mdctgan_loss_item {
  const float* a; const float* b; float* g;
  int64_t n; double coef; const float* gscale;
  float target; int32_t kind; int32_t slot; int32_t pad_;
} mdctgan_loss_item;
int mdctgan_multi_loss_fwd(const mdctgan_loss_item* items, int n_items, double* acc, int n_slots, void* stream);
int mdctgan_multi_loss_bwd(const mdctgan_loss_item* items, int n_items, void* stream);
int mdctgan_bce_const_fwd(const float* x, int64_t n, float target, double coef, double* slot, void* stream);
int mdctgan_bce_const_bwd(const float* x, int64_t n, float target, float coef, const float* gscale, float* g, int accumulate, void* stream);
/* feature matching (pix2pixHD_model.py:447-451): *slot += coef * sum|a - b|; g (+)= coef*(*gscale)*sign(a - b) */
int mdctgan_l1_pair_fwd(const float* a, const float* b, int64_t n, double coef, double* slot, void* stream);
int mdctgan_l1_pair_bwd(const float* a, const float* b, int64_t n, float coef, const float* gscale, float* g, int accumulate, void* stream);
int mdctgan_f64_to_f32(const double* a, float* y, int n, void* stream);
/* cat(lr_spectro, s, |s|*2+lo) (pix2pixHD_model.py:369,420-427) as NHWC [clips*per_clip][3]; backward ds = g1 + 2 sign(s) g2 */
int mdctgan_disc_input_fwd(const float* lr, int64_t lr_clip_stride, const float* s, float* out, int64_t clips, int64_t per_clip, float lo,
                           void* stream);
int mdctgan_disc_input_bwd(const float* g, const float* s, float* ds, int64_t n, void* stream);
/* torch.optim.Adam (pix2pixHD_model.py:350-364; no weight decay) on one flat fp32 buffer; g is multiplied by grad_scale
 * (1/world after the all-reduce).  step: 1-based, or read from step_dev when non-NULL (CUDA-graph replays). */
int mdctgan_adam_flat(float* p, const float* g, float* m, float* v, int64_t n, float lr, float beta1, float beta2, float eps,
                      float grad_scale, int64_t step, const int64_t* step_dev, void* stream);
int mdctgan_counter_inc(int64_t* counter_dev, void* stream);
/* Kernel-side weight images of every convolution of a model in ONE launch (run after each optimiser step; replaces the
 * per-layer host-side re-layouts).  Element (k, n) of descriptor d: tap = k / Kch, kc = k % Kch, value =
 * src[kc*s_kch + n*s_n + (flip ? taps-1-tap : tap)*s_tap]; written to dst_kn[k*N + n] (nullable) and / or as TF32 hi|lo into the
 * tcgen05 image dst_umma (nullable; layout of mdctgan_conv2d_umma_pack_weight, kchunks*2*N*32 floats, rows k >= K zero).
 * work_begin = prefix sum of kchunks*32*N.  `descs_dev`: device array of n_desc descriptors. */
typedef struct mdctgan_pack_desc {
  const float* src; float* dst_kn; float* dst_umma;
  int32_t K, N, Kch, taps, flip, kchunks;
  int64_t s_kch, s_n, s_tap;
  int64_t work_begin;
} mdctgan_pack_desc;
int mdctgan_pack_weights_multi(const void* descs_dev, int n_desc, int64_t total_work, void* stream);
/* The same for descriptors with dst_kn == NULL, Kch % 32 == 0, N % 32 == 0, taps <= 49, as a shared-memory tiled transpose (every
 * global access coalesced): tile_begin_dev[i] = prefix sum of (N/32)*(Kch/32); max_taps = largest taps in the table. */
int mdctgan_pack_weights_tiled(const void* descs_dev, const int64_t* tile_begin_dev, int n_desc, int64_t total_tiles, int max_taps, void* stream);

/* ------------------------------------------------------------------------------------------------
 * Long-form generation (generate_audio.py:29-53).
 */
/* number of segments AudioTestDataset.seg_pad_audio (data/audio_dataset.py:153-167) produces for an L-sample clip */
int64_t mdctgan_segment_count(int64_t L, int seg, int ov);
/* seg_pad_audio: clip [L] fp32 -> [n_seg, seg] (zero-pad (ov, seg*ceil(L/seg) - L + ov), unfold(size seg, step seg - ov)) */
int mdctgan_segment_gather(const float* audio_dev, int64_t L, float* out_dev, int64_t n_seg, int seg, int ov, void* stream);
/* generate_audio.py:40-53: halve the first / last ov samples of each segment, fold(stride seg - ov), crop ov at both ends;
 * [n_seg, seg] -> [(n_seg-1)*(seg-ov) + seg - 2*ov]; ov = 0 is the plain concatenation.  precision: MDCTGAN_F32 / _F64 (in and out). */
int mdctgan_segment_ola(const void* seg_dev, void* out_dev, int64_t n_seg, int seg, int ov, int precision, void* stream);
/* The same fold over a contiguous RUN of segments (multi-GPU long-form generation: every rank folds its own run):
 * crops crop_begin / crop_end (each in [0, ov]) samples instead of ov -- 0 at an interior shard edge, where the ov half-weighted
 * samples that overlap the neighbouring run are kept and summed by whoever assembles the shards.
 * [n_seg, seg] -> [(n_seg-1)*(seg-ov) + seg - crop_begin - crop_end] */
int mdctgan_segment_ola_part(const void* seg_dev, void* out_dev, int64_t n_seg, int seg, int ov, int crop_begin, int crop_end,
                             int precision, void* stream);

/* AudioDataset.__getitem__ noise injection (data/audio_dataset.py:72-78): out = lr + sqrt(sum(lr^2)/segment_length/10^(snr/10)) /
 * std(noise) * (noise - mean(noise)), `noise` = the caller's N(0,1) draw (torch.randn in the reference), std unbiased.
 * scratch3_dev: 3 doubles of device scratch.  Two launches. */
int mdctgan_add_noise(const float* lr_dev, const float* noise_dev, float* out_dev, int64_t n, double segment_length, double snr_db,
                      double* scratch3_dev, void* stream);

/* ------------------------------------------------------------------------------------------------
 * Evaluation metrics (util/util.py:132-177 compute_matrics; callers train.py:116-117, generate_audio.py:59-60).
 */
/* per row r of [rows, T]: rows_out[3r..] += (sum (sr-hr)^2, sum hr^2, sum (lr-hr)^2) in double (zero it first): MSE and the SNRs */
int mdctgan_metrics_rows(const float* hr, const float* lr, const float* sr, int64_t rows, int64_t T, double* rows_out, void* stream);
/* frames of aF.spectrogram(n_fft, hop, center) */
int64_t mdctgan_lsd_frame_count(int64_t T, int n_fft, int hop, int center);
/* *acc += sum over rows and frames of sqrt(mean_k (log10(|STFT hr|^2 + 1e-6) - log10(|STFT sr|^2 + 1e-6))^2): the LSD of
 * util/util.py:170-175 is *acc / (rows * frames).  window_dev: n_fft fp32 values (kbdwin(2*win)); n_fft 512 / 1024 / 2048. */
int mdctgan_lsd_frames(const float* hr, const float* sr, int64_t rows, int64_t T, int n_fft, int hop, const float* window_dev, int center,
                       double* acc, void* stream);

/* ------------------------------------------------------------------------------------------------
 * Data preparation in front of the path: torchaudio.functional.resample (sinc_interp_hann) as the reference calls it
 * (data/audio_dataset.py:66-71, :169-177).  `table`: [nw][K] polyphase FIR (K = 2*width + orig; orig / nw = the two rates
 * divided by their gcd); x [rows, L] -> y [rows, target], target = ceil(nw * L / orig).
 */
int mdctgan_resample_fir(const float* x, int rows, int64_t L, const float* table, int K, int orig, int nw, int width, float* y,
                         int64_t target, void* stream);

/* Introspection for tests / bench: number of kernels this library has launched in this process. */
int64_t mdctgan_launch_count(void);

#ifdef __cplusplus
}
#endif
#endif /* MDCTGAN_B200_H_ */
