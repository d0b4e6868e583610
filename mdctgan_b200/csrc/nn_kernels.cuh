// nn_kernels.cuh -- fp32 CUDA-core kernels for the generator / discriminator stacks (sm_100a).
//
// These are the shape-generic kernels of the network half of the hot path (reference: models/networks.py
// GlobalGenerator :301-372, LocalEnhancer :173-298, ResnetBlock :421-463, NLayerDiscriminator :641-692):
// every convolution flavour the reference builds (k 7/5/4/3/1, stride 1/2, zero / reflection padding,
// ConvTranspose2d k3 s2 p1 op1), InstanceNorm2d(affine=False) / BatchNorm2d folded into the producer's
// epilogue (statistics) and the consumer's prologue (normalise + activation), the residual / branch adds,
// AvgPool2d(3, 2, 1, count_include_pad=False) and the Cout = 1 heads with tanh.
// Activations are NHWC fp32 ([B, H, W, C], channels contiguous); weights are packed [kh*kw*Cin][Cout].
//
// The tcgen05 implicit-GEMM kernel (conv_umma.cuh) takes over the 3x3 stride-1 residual-block convolutions,
// which are > 80 % of the generator's FLOPs; the kernels here remain the reference
// implementation on the device for those and the production path for the degenerate shapes
// (Cin = 2 stem, Cout = 1 head, strided / transposed layers).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace nnk {

enum Act : int { kActNone = 0, kActRelu = 1, kActLeaky = 2, kActTanh = 3, kActSigmoid = 4 };   // sigmoid: the --no_lsgan PatchGAN head (networks.py:672)
enum PadMode : int { kPadZero = 0, kPadReflect = 1 };

// How a consumer sees a producer's raw output: v = act(x * scale[b][c] + shift[b][c]).
// Built by norm_finalize_kernel from the producer's (sum, sumsq) statistics (InstanceNorm: per sample and
// channel; BatchNorm: per channel) -- or absent (identity) when `scale` is null.
struct InputNorm {
  const float* scale;   // [B][C] or [C] (per_sample = 0), null = identity
  const float* shift;
  int per_sample;
  int act;              // activation applied after the affine (kActNone / kActRelu / kActLeaky)
  // alternative to scale / shift for InstanceNorm2d(affine=False): the producer's raw [B][C][2] (sum, sumsq); the
  // consumer derives scale = rstd, shift = -mean*rstd itself (no finalize launch).  tcgen05 convolution only.
  const double* stats; float count; float eps;
};

struct ConvParams {
  const float* x; int B, H, W, Cin;
  const float* w;       // [kh*kw*Cin][Cout]
  const float* bias;    // [Cout] or null
  float* y; int Ho, Wo, Cout;
  int kh, kw, stride, pad, pad_mode, transposed;
  InputNorm in;
  int act;              // epilogue activation
  double* stats;        // [B][Cout][2] (sum, sumsq) accumulated with atomics, or null
  int tiles_per_sample;
};

// scale / shift of sample `b` into shared memory: ready-made, or derived from the raw InstanceNorm statistics.  rstd = rsqrt(var + eps) in
// fp64 (1 ulp, a dozen instructions) everywhere it is derived -- `1.0 / sqrt()` is ~100 fp64-pipe instructions per value, which was
// 1 - 2 us of every consumer's prologue (four values per thread of a 4 x 512-channel table)
__device__ __forceinline__ void norm_to_smem(const InputNorm& in, int b, int C, float* s_scale, float* s_shift, int tid, int nthreads) {
  const size_t off = in.per_sample ? (size_t)b * C : 0;
  if (in.stats) {
    const double inv_n = 1.0 / (double)in.count;
    for (int c = tid; c < C; c += nthreads) {
      const double mean = in.stats[2 * (off + c)] * inv_n;
      double var = in.stats[2 * (off + c) + 1] * inv_n - mean * mean;
      var = var < 0.0 ? 0.0 : var;
      const double rstd = rsqrt(var + (double)in.eps);
      s_scale[c] = (float)rstd;
      s_shift[c] = (float)(-mean * rstd);
    }
  } else if (in.scale) {
    for (int c = tid; c < C; c += nthreads) { s_scale[c] = __ldg(in.scale + off + c); s_shift[c] = __ldg(in.shift + off + c); }
  }
}

__device__ __forceinline__ float apply_act(float v, int act) {
  if (act == kActRelu) return fmaxf(v, 0.f);
  if (act == kActLeaky) return v > 0.f ? v : 0.2f * v;
  if (act == kActTanh) return tanhf(v);
  if (act == kActSigmoid) return 1.f / (1.f + expf(-v));
  return v;
}

// input coordinate of output pixel `o` for tap `k` along one axis; returns -1 when the tap reads padding
__device__ __forceinline__ int in_coord(int o, int k, int n_in, int stride, int pad, int pad_mode, int transposed) {
  if (!transposed) {
    int i = o * stride - pad + k;
    if (pad_mode == kPadReflect) {
      if (i < 0) i = -i;
      if (i >= n_in) i = 2 * n_in - 2 - i;
      return i;
    }
    return (i >= 0 && i < n_in) ? i : -1;
  }
  const int t = o + pad - k;   // ConvTranspose2d: o = i*stride - pad + k
  if (t < 0 || (t % stride) != 0) return -1;
  const int i = t / stride;
  return i < n_in ? i : -1;
}

// ------------------------------------------------------------------------------------------------
// Implicit-GEMM convolution, fp32 FFMA.  CTA tile: BM output pixels (of ONE sample) x BN output channels,
// K = kh*kw*Cin walked in chunks of 16; 256 threads, (BM/16) x (BN/16) outputs per thread.
// ------------------------------------------------------------------------------------------------
template <int BM, int BN>
__global__ void __launch_bounds__(256) conv2d_nhwc_kernel(const ConvParams p) {
  constexpr int BK = 16;
  constexpr int TM = BM / 16, TN = BN / 16;
  __shared__ __align__(16) float As[BK][BM + 4];
  __shared__ __align__(16) float Bs[BK][BN];
  __shared__ float s_scale[1024], s_shift[1024];   // Cin <= 1024 (train.sh config: 896)

  const int tid = threadIdx.x;
  const int b = blockIdx.x / p.tiles_per_sample;
  const int m0 = (blockIdx.x - b * p.tiles_per_sample) * BM;   // first output pixel of the tile within the sample
  const int n0 = blockIdx.y * BN;
  const int HWo = p.Ho * p.Wo;
  const int K = p.kh * p.kw * p.Cin;
  const bool has_norm = p.in.scale != nullptr || p.in.stats != nullptr;
  if (has_norm) norm_to_smem(p.in, b, p.Cin, s_scale, s_shift, tid, 256);
  __syncthreads();

  // A-load assignment: thread loads 4 consecutive k for pixel (tid % BM) in rounds; with 256 threads and
  // BM x 16 elements per chunk = BM*4 float4 -> BM*4/256 rounds
  constexpr int A_ROUNDS = (BM * 4 + 255) / 256;
  constexpr int B_ROUNDS = (BN * 4 + 255) / 256;
  const float* xb = p.x + (size_t)b * p.H * p.W * p.Cin;
  const bool vec_ok = (p.Cin % 4) == 0;

  float acc[TM][TN];
#pragma unroll
  for (int i = 0; i < TM; ++i)
#pragma unroll
    for (int jn = 0; jn < TN; ++jn) acc[i][jn] = 0.f;

  const int ty = tid / 16, tx = tid % 16;

  for (int k0 = 0; k0 < K; k0 += BK) {
    // ---- stage A (im2col gather with the producer's normalisation + activation applied on the fly)
#pragma unroll
    for (int rd = 0; rd < A_ROUNDS; ++rd) {
      const int idx = tid + rd * 256;
      if (idx < BM * 4) {
        const int pm = idx / 4, kq = (idx % 4) * 4;
        const int m = m0 + pm;
        float v[4] = {0.f, 0.f, 0.f, 0.f};
        if (m < HWo) {
          const int oy = m / p.Wo, ox = m - oy * p.Wo;
          if (vec_ok) {
            const int k = k0 + kq;
            if (k < K) {
              const int tap = k / p.Cin, c = k - tap * p.Cin;
              const int ky = tap / p.kw, kx = tap - ky * p.kw;
              const int iy = in_coord(oy, ky, p.H, p.stride, p.pad, p.pad_mode, p.transposed);
              const int ix = in_coord(ox, kx, p.W, p.stride, p.pad, p.pad_mode, p.transposed);
              if (iy >= 0 && ix >= 0) {
                const float4 q = __ldg(reinterpret_cast<const float4*>(xb + ((size_t)iy * p.W + ix) * p.Cin + c));
                v[0] = q.x; v[1] = q.y; v[2] = q.z; v[3] = q.w;
                if (has_norm) {
#pragma unroll
                  for (int u = 0; u < 4; ++u) v[u] = apply_act(fmaf(v[u], s_scale[c + u], s_shift[c + u]), p.in.act);
                }
              }
            }
          } else {
#pragma unroll
            for (int u = 0; u < 4; ++u) {
              const int k = k0 + kq + u;
              if (k < K) {
                const int tap = k / p.Cin, c = k - tap * p.Cin;
                const int ky = tap / p.kw, kx = tap - ky * p.kw;
                const int iy = in_coord(oy, ky, p.H, p.stride, p.pad, p.pad_mode, p.transposed);
                const int ix = in_coord(ox, kx, p.W, p.stride, p.pad, p.pad_mode, p.transposed);
                if (iy >= 0 && ix >= 0) {
                  float t = __ldg(xb + ((size_t)iy * p.W + ix) * p.Cin + c);
                  if (has_norm) t = apply_act(fmaf(t, s_scale[c], s_shift[c]), p.in.act);
                  v[u] = t;
                }
              }
            }
          }
        }
#pragma unroll
        for (int u = 0; u < 4; ++u) As[kq + u][pm] = v[u];
      }
    }
    // ---- stage B (weights, [K][Cout])
#pragma unroll
    for (int rd = 0; rd < B_ROUNDS; ++rd) {
      const int idx = tid + rd * 256;
      if (idx < BN * 4) {
        const int kk = idx / (BN / 4), nq = (idx % (BN / 4)) * 4;
        const int k = k0 + kk, n = n0 + nq;
        float4 q = make_float4(0.f, 0.f, 0.f, 0.f);
        if (k < K) {
          if (n + 3 < p.Cout && (p.Cout % 4) == 0) {
            q = __ldg(reinterpret_cast<const float4*>(p.w + (size_t)k * p.Cout + n));
          } else {
            if (n + 0 < p.Cout) q.x = __ldg(p.w + (size_t)k * p.Cout + n + 0);
            if (n + 1 < p.Cout) q.y = __ldg(p.w + (size_t)k * p.Cout + n + 1);
            if (n + 2 < p.Cout) q.z = __ldg(p.w + (size_t)k * p.Cout + n + 2);
            if (n + 3 < p.Cout) q.w = __ldg(p.w + (size_t)k * p.Cout + n + 3);
          }
        }
        *reinterpret_cast<float4*>(&Bs[kk][nq]) = q;
      }
    }
    __syncthreads();
#pragma unroll
    for (int kk = 0; kk < BK; ++kk) {
      float a[TM], bv[TN];
#pragma unroll
      for (int i = 0; i < TM; ++i) a[i] = As[kk][ty * TM + i];
#pragma unroll
      for (int jn = 0; jn < TN; ++jn) bv[jn] = Bs[kk][tx * TN + jn];
#pragma unroll
      for (int i = 0; i < TM; ++i)
#pragma unroll
        for (int jn = 0; jn < TN; ++jn) acc[i][jn] = fmaf(a[i], bv[jn], acc[i][jn]);
    }
    __syncthreads();
  }

  // ---- epilogue: bias, activation, store, InstanceNorm / BatchNorm statistics
  // fp64 partial sums: var = E[x^2] - mean^2 must survive planes with |mean| >> std (e.g. the PatchGAN layers behind the
  // nearly constant |s|*2+lo input channel); torch's two-pass / Welford statistics do
  double csum[TN], csq[TN];
#pragma unroll
  for (int jn = 0; jn < TN; ++jn) { csum[jn] = 0.0; csq[jn] = 0.0; }
  float* yb = p.y + (size_t)b * HWo * p.Cout;
#pragma unroll
  for (int i = 0; i < TM; ++i) {
    const int m = m0 + ty * TM + i;
#pragma unroll
    for (int jn = 0; jn < TN; ++jn) {
      const int n = n0 + tx * TN + jn;
      if (m < HWo && n < p.Cout) {
        float v = acc[i][jn] + (p.bias ? __ldg(p.bias + n) : 0.f);
        if (p.stats) { csum[jn] += (double)v; csq[jn] = fma((double)v, (double)v, csq[jn]); }   // BEFORE the epilogue activation (norm follows conv)
        yb[(size_t)m * p.Cout + n] = apply_act(v, p.act);
      }
    }
  }
  if (p.stats) {
    // reduce over the 16 ty-rows of the CTA through shared memory, then one atomic per channel per CTA
    __shared__ double red[2][16][BN];
#pragma unroll
    for (int jn = 0; jn < TN; ++jn) { red[0][ty][tx * TN + jn] = csum[jn]; red[1][ty][tx * TN + jn] = csq[jn]; }
    __syncthreads();
    if (tid < BN) {
      double s = 0.0, q = 0.0;
#pragma unroll
      for (int r = 0; r < 16; ++r) { s += red[0][r][tid]; q += red[1][r][tid]; }
      const int n = n0 + tid;
      if (n < p.Cout) {
        double* st = p.stats + ((size_t)b * p.Cout + n) * 2;
        atomicAdd(st, s);
        atomicAdd(st + 1, q);
      }
    }
  }
}

// ------------------------------------------------------------------------------------------------
// Cout <= 4 convolution (generator head 7x7 -> 1 + tanh, PatchGAN heads 4x4 -> 1, input gradient of the first PatchGAN
// layer 64 -> 3): 8 lanes per output pixel, lanes split the channels in float4, shuffle-reduce.  Dead taps of a
// transposed / strided geometry are skipped per pixel.  Weights [kh*kw*Cin][COUT] in shared memory.
// ------------------------------------------------------------------------------------------------
template <int COUT>
__global__ void __launch_bounds__(256) conv2d_cout_small_kernel(const ConvParams p) {
  extern __shared__ float s_w[];   // [kh*kw*Cin][COUT] weights, then scale/shift [Cin] x2 when normalising
  const int K = p.kh * p.kw * p.Cin;
  float* s_scale = s_w + K * COUT;
  float* s_shift = s_scale + p.Cin;
  const int HWo = p.Ho * p.Wo;
  const int groups_per_block = 256 / 8;
  const int blocks_per_sample = (HWo + groups_per_block - 1) / groups_per_block;
  const int b = blockIdx.x / blocks_per_sample;
  const int m = (blockIdx.x - b * blocks_per_sample) * groups_per_block + threadIdx.x / 8;
  const int l8 = threadIdx.x % 8;
  for (int i = threadIdx.x; i < K * COUT; i += 256) s_w[i] = __ldg(p.w + i);
  const bool has_norm = p.in.scale != nullptr || p.in.stats != nullptr;
  if (has_norm) norm_to_smem(p.in, b, p.Cin, s_scale, s_shift, threadIdx.x, 256);
  __syncthreads();
  float acc[COUT];
#pragma unroll
  for (int j = 0; j < COUT; ++j) acc[j] = 0.f;
  if (m < HWo) {
    const int oy = m / p.Wo, ox = m - oy * p.Wo;
    const float* xb = p.x + (size_t)b * p.H * p.W * p.Cin;
    for (int ky = 0; ky < p.kh; ++ky) {
      const int iy = in_coord(oy, ky, p.H, p.stride, p.pad, p.pad_mode, p.transposed);
      if (iy < 0) continue;
      for (int kx = 0; kx < p.kw; ++kx) {
        const int ix = in_coord(ox, kx, p.W, p.stride, p.pad, p.pad_mode, p.transposed);
        if (ix < 0) continue;
        const float* px = xb + ((size_t)iy * p.W + ix) * p.Cin;
        const float* wt = s_w + (size_t)(ky * p.kw + kx) * p.Cin * COUT;
        for (int c = l8 * 4; c < p.Cin; c += 32) {
          if (c + 3 < p.Cin && (p.Cin % 4) == 0) {
            float4 q = __ldg(reinterpret_cast<const float4*>(px + c));
            float v[4] = {q.x, q.y, q.z, q.w};
#pragma unroll
            for (int u = 0; u < 4; ++u) {
              if (has_norm) v[u] = apply_act(fmaf(v[u], s_scale[c + u], s_shift[c + u]), p.in.act);
#pragma unroll
              for (int j = 0; j < COUT; ++j) acc[j] = fmaf(v[u], wt[(c + u) * COUT + j], acc[j]);
            }
          } else {
            for (int u = 0; u < 4 && c + u < p.Cin; ++u) {
              float v = __ldg(px + c + u);
              if (has_norm) v = apply_act(fmaf(v, s_scale[c + u], s_shift[c + u]), p.in.act);
#pragma unroll
              for (int j = 0; j < COUT; ++j) acc[j] = fmaf(v, wt[(c + u) * COUT + j], acc[j]);
            }
          }
        }
      }
    }
  }
#pragma unroll
  for (int j = 0; j < COUT; ++j) {
    acc[j] += __shfl_xor_sync(0xffffffffu, acc[j], 4);
    acc[j] += __shfl_xor_sync(0xffffffffu, acc[j], 2);
    acc[j] += __shfl_xor_sync(0xffffffffu, acc[j], 1);
  }
  if (m < HWo && l8 == 0) {
#pragma unroll
    for (int j = 0; j < COUT; ++j) {
      const float v = acc[j] + (p.bias ? __ldg(p.bias + j) : 0.f);
      p.y[((size_t)b * HWo + m) * COUT + j] = apply_act(v, p.act);
    }
    // (no statistics: the reference never normalises these maps)
  }
}

// ------------------------------------------------------------------------------------------------
// Cout == 1, Cin % 4 == 0 (generator head 7x7 -> 1 + tanh: LANES = 8; PatchGAN heads 4x4 512 -> 1: LANES = 32): LANES lanes per
// output pixel, lane l owns the channel quads c = 4 l + 4 LANES i (i < NI <= 4).  The r01 / r02 profiles had the Cout <= 4 kernel above
// at 70 - 85 us on these layers: 12 shared-memory loads per tap (weights + scale / shift, scalar) next to one global load on the
// 7x7 head (LSU-issue bound), and a 256-iteration serial load chain per thread on the 512-channel heads (latency bound).  Here the
// deferred normalisation of the thread's channels lives in registers, weights come as float4 through L1 (__ldg: every pixel
// group reads the same K floats), and the tap loop is unrolled so that several independent global loads are in flight.
// ------------------------------------------------------------------------------------------------
template <int LANES, int NI>
__global__ void __launch_bounds__(256) conv2d_cout1_kernel(const ConvParams p) {
  __shared__ float s_scale[1024], s_shift[1024];
  const int HWo = p.Ho * p.Wo;
  constexpr int kPix = 256 / LANES;
  const int blocks_per_sample = (HWo + kPix - 1) / kPix;
  const int b = blockIdx.x / blocks_per_sample;
  const int m = (blockIdx.x - b * blocks_per_sample) * kPix + threadIdx.x / LANES;
  const int l = threadIdx.x % LANES;
  const bool has_norm = p.in.scale != nullptr || p.in.stats != nullptr;
  if (has_norm) norm_to_smem(p.in, b, p.Cin, s_scale, s_shift, threadIdx.x, 256);
  __syncthreads();
  float4 sc[NI], sh[NI];
#pragma unroll
  for (int i = 0; i < NI; ++i) {
    const int c = 4 * l + 4 * LANES * i;
    sc[i] = make_float4(1.f, 1.f, 1.f, 1.f); sh[i] = make_float4(0.f, 0.f, 0.f, 0.f);
    if (has_norm && c < p.Cin) { sc[i] = *reinterpret_cast<const float4*>(s_scale + c); sh[i] = *reinterpret_cast<const float4*>(s_shift + c); }
  }
  const float slope = !has_norm ? 1.f : (p.in.act == kActRelu ? 0.f : (p.in.act == kActLeaky ? 0.2f : 1.f));   // v = max(e, slope e)
  float acc = 0.f;
  if (m < HWo) {
    const int oy = m / p.Wo, ox = m - oy * p.Wo;
    const float* xb = p.x + (size_t)b * p.H * p.W * p.Cin;
    for (int ky = 0; ky < p.kh; ++ky) {
      const int iy = in_coord(oy, ky, p.H, p.stride, p.pad, p.pad_mode, p.transposed);
      if (iy < 0) continue;
#pragma unroll 4
      for (int kx = 0; kx < p.kw; ++kx) {
        const int ix = in_coord(ox, kx, p.W, p.stride, p.pad, p.pad_mode, p.transposed);
        if (ix < 0) continue;
        const float* px = xb + ((size_t)iy * p.W + ix) * p.Cin;
        const float* wt = p.w + (size_t)(ky * p.kw + kx) * p.Cin;
#pragma unroll
        for (int i = 0; i < NI; ++i) {
          const int c = 4 * l + 4 * LANES * i;
          if (c < p.Cin) {
            const float4 q = __ldg(reinterpret_cast<const float4*>(px + c));
            const float4 w = __ldg(reinterpret_cast<const float4*>(wt + c));
            float e0 = fmaf(q.x, sc[i].x, sh[i].x), e1 = fmaf(q.y, sc[i].y, sh[i].y), e2 = fmaf(q.z, sc[i].z, sh[i].z), e3 = fmaf(q.w, sc[i].w, sh[i].w);
            e0 = fmaxf(e0, slope * e0); e1 = fmaxf(e1, slope * e1); e2 = fmaxf(e2, slope * e2); e3 = fmaxf(e3, slope * e3);
            acc = fmaf(e0, w.x, acc); acc = fmaf(e1, w.y, acc); acc = fmaf(e2, w.z, acc); acc = fmaf(e3, w.w, acc);
          }
        }
      }
    }
  }
#pragma unroll
  for (int off = LANES / 2; off; off >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, off);
  if (m < HWo && l == 0) p.y[(size_t)b * HWo + m] = apply_act(acc + (p.bias ? __ldg(p.bias) : 0.f), p.act);
}

// The 7x7 generator head (stride 1, Cin <= 64) once more, with register reuse along the row: an 8-lane group computes FOUR adjacent
// output pixels, so a row of taps needs KW + 3 input pixels instead of 4 KW, and every weight quad is loaded once per four pixels.
// conv2d_cout1_kernel<8, 1> on this layer measured 66 us: 98 float4 loads per thread through L1 (4 + 1 wavefronts each) = L1-bound.
template <int KW, int NI>
__global__ void __launch_bounds__(256) conv2d_cout1_row4_kernel(const ConvParams p) {
  __shared__ float s_scale[64], s_shift[64];
  const int groups_per_row = p.Wo / 4;                        // (host: Wo % 4 == 0)
  const int groups_per_sample = p.Ho * groups_per_row;
  const int blocks_per_sample = (groups_per_sample + 31) / 32;
  const int b = blockIdx.x / blocks_per_sample;
  const int g = (blockIdx.x - b * blocks_per_sample) * 32 + threadIdx.x / 8;
  const int l = threadIdx.x % 8;
  const bool has_norm = p.in.scale != nullptr || p.in.stats != nullptr;
  if (has_norm) norm_to_smem(p.in, b, p.Cin, s_scale, s_shift, threadIdx.x, 256);
  __syncthreads();
  const float slope = !has_norm ? 1.f : (p.in.act == kActRelu ? 0.f : (p.in.act == kActLeaky ? 0.2f : 1.f));
  float acc[4] = {0.f, 0.f, 0.f, 0.f};
  const bool live = g < groups_per_sample;
  const int oy = live ? g / groups_per_row : 0, ox0 = live ? (g - oy * groups_per_row) * 4 : 0;
  const float* xb = p.x + (size_t)b * p.H * p.W * p.Cin;
  if (live) {
#pragma unroll
    for (int i = 0; i < NI; ++i) {
      const int c = 4 * l + 32 * i;
      if (c >= p.Cin) break;
      float4 sc = make_float4(1.f, 1.f, 1.f, 1.f), sh = make_float4(0.f, 0.f, 0.f, 0.f);
      if (has_norm) { sc = *reinterpret_cast<const float4*>(s_scale + c); sh = *reinterpret_cast<const float4*>(s_shift + c); }
      int ixs[KW + 3];
#pragma unroll
      for (int t = 0; t < KW + 3; ++t) ixs[t] = in_coord(ox0, t, p.W, 1, p.pad, p.pad_mode, 0);      // input column of (ox0 + j, kx), t = j + kx
      for (int ky = 0; ky < p.kh; ++ky) {
        const int iy = in_coord(oy, ky, p.H, 1, p.pad, p.pad_mode, 0);
        if (iy < 0) continue;
        const float* row = xb + (size_t)iy * p.W * p.Cin + c;
        float4 xs[KW + 3];
#pragma unroll
        for (int t = 0; t < KW + 3; ++t) {
          float4 e = make_float4(0.f, 0.f, 0.f, 0.f);
          if (ixs[t] >= 0) {
            const float4 q = __ldg(reinterpret_cast<const float4*>(row + (size_t)ixs[t] * p.Cin));
            e.x = fmaf(q.x, sc.x, sh.x); e.y = fmaf(q.y, sc.y, sh.y); e.z = fmaf(q.z, sc.z, sh.z); e.w = fmaf(q.w, sc.w, sh.w);
            e.x = fmaxf(e.x, slope * e.x); e.y = fmaxf(e.y, slope * e.y); e.z = fmaxf(e.z, slope * e.z); e.w = fmaxf(e.w, slope * e.w);
          }
          xs[t] = e;
        }
        const float* wt = p.w + (size_t)ky * KW * p.Cin + c;
#pragma unroll
        for (int kx = 0; kx < KW; ++kx) {
          const float4 w = __ldg(reinterpret_cast<const float4*>(wt + (size_t)kx * p.Cin));
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            const float4 e = xs[kx + j];
            acc[j] = fmaf(e.x, w.x, acc[j]); acc[j] = fmaf(e.y, w.y, acc[j]); acc[j] = fmaf(e.z, w.z, acc[j]); acc[j] = fmaf(e.w, w.w, acc[j]);
          }
        }
      }
    }
  }
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    acc[j] += __shfl_xor_sync(0xffffffffu, acc[j], 4);
    acc[j] += __shfl_xor_sync(0xffffffffu, acc[j], 2);
    acc[j] += __shfl_xor_sync(0xffffffffu, acc[j], 1);
  }
  if (live && l == 0) {
    const float bias = p.bias ? __ldg(p.bias) : 0.f;
    float4 o = make_float4(apply_act(acc[0] + bias, p.act), apply_act(acc[1] + bias, p.act), apply_act(acc[2] + bias, p.act), apply_act(acc[3] + bias, p.act));
    *reinterpret_cast<float4*>(p.y + (size_t)b * p.Ho * p.Wo + (size_t)oy * p.Wo + ox0) = o;
  }
}

// ------------------------------------------------------------------------------------------------
// (sum, sumsq) -> (scale, shift).  InstanceNorm2d(affine=False, eps): per (b, c); BatchNorm2d: per c, with
// optional affine and running-statistics update (momentum) in training mode, or running statistics in eval.
// ------------------------------------------------------------------------------------------------
struct NormFinalizeParams {
  const double* stats;   // [B][C][2]
  int B, C; double count;   // elements per (b, c) plane
  float eps;
  int batch_norm;        // 0: instance (per b, c)   1: batch statistics over b   2: eval (running statistics)
  const float* gamma; const float* beta;        // BatchNorm affine (nullable)
  float* running_mean; float* running_var; float momentum;   // BatchNorm buffers (nullable)
  float* scale; float* shift;                    // out: [B][C] (instance) or [C] (batch)
};

static __global__ void norm_finalize_kernel(const NormFinalizeParams p) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (p.batch_norm == 0) {
    if (i >= p.B * p.C) return;
    const double mean = p.stats[2 * (size_t)i] / p.count;
    double var = p.stats[2 * (size_t)i + 1] / p.count - mean * mean;
    if (var < 0) var = 0;
    const double rstd = rsqrt(var + (double)p.eps);
    p.scale[i] = (float)rstd;
    p.shift[i] = (float)(-mean * rstd);
    return;
  }
  if (i >= p.C) return;
  double mean, var;
  if (p.batch_norm == 1) {
    double s = 0, q = 0;
    for (int b = 0; b < p.B; ++b) { s += p.stats[2 * ((size_t)b * p.C + i)]; q += p.stats[2 * ((size_t)b * p.C + i) + 1]; }
    const double n = p.count * p.B;
    mean = s / n;
    var = q / n - mean * mean;
    if (var < 0) var = 0;
    if (p.running_mean) {   // torch: running_var uses the unbiased estimate
      p.running_mean[i] = (1.f - p.momentum) * p.running_mean[i] + p.momentum * (float)mean;
      p.running_var[i] = (1.f - p.momentum) * p.running_var[i] + p.momentum * (float)(var * n / (n - 1.0));
    }
  } else {
    mean = p.running_mean[i];
    var = p.running_var[i];
  }
  const double rstd = rsqrt(var + (double)p.eps);
  const double g = p.gamma ? (double)p.gamma[i] : 1.0, be = p.beta ? (double)p.beta[i] : 0.0;
  p.scale[i] = (float)(rstd * g);
  p.shift[i] = (float)(be - mean * rstd * g);
}

// ------------------------------------------------------------------------------------------------
// y = act_a(a * sa + ta) [+ act_b(b * sb + tb)]   (materialises a normalised tensor: residual-block
// output x + IN(conv), LocalEnhancer branch sum, BottleStack shortcut).  float4 over channels.
// ------------------------------------------------------------------------------------------------
struct ApplyParams {
  const float* a; InputNorm na;
  const float* b; InputNorm nb;   // b nullable
  float* y; int B, HW, C; int act_out;
};

// grid = (chunks per sample, B): a block stays inside one sample so the (possibly statistics-derived) scale / shift
// of that sample sit in shared memory
static __global__ void __launch_bounds__(256) norm_apply_kernel(const ApplyParams p) {
  __shared__ float sa[2][1024], sb[2][1024];
  const int bi = blockIdx.y;
  const bool na = p.na.scale != nullptr || p.na.stats != nullptr;
  const bool nb = p.b != nullptr && (p.nb.scale != nullptr || p.nb.stats != nullptr);
  if (na) norm_to_smem(p.na, bi, p.C, sa[0], sa[1], threadIdx.x, 256);
  if (nb) norm_to_smem(p.nb, bi, p.C, sb[0], sb[1], threadIdx.x, 256);
  __syncthreads();
  const size_t per_sample4 = (size_t)p.HW * p.C / 4;
  const size_t base = (size_t)bi * p.HW * p.C;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < per_sample4; i += (size_t)gridDim.x * blockDim.x) {
    const size_t e = i * 4;
    const int c = (int)(e % p.C);
    float4 va = __ldg(reinterpret_cast<const float4*>(p.a + base + e));
    float v[4] = {va.x, va.y, va.z, va.w};
    if (na) {
#pragma unroll
      for (int u = 0; u < 4; ++u) v[u] = fmaf(v[u], sa[0][c + u], sa[1][c + u]);
    }
#pragma unroll
    for (int u = 0; u < 4; ++u) v[u] = apply_act(v[u], p.na.act);
    if (p.b) {
      float4 vb = __ldg(reinterpret_cast<const float4*>(p.b + base + e));
      float w[4] = {vb.x, vb.y, vb.z, vb.w};
      if (nb) {
#pragma unroll
        for (int u = 0; u < 4; ++u) w[u] = fmaf(w[u], sb[0][c + u], sb[1][c + u]);
      }
#pragma unroll
      for (int u = 0; u < 4; ++u) v[u] += apply_act(w[u], p.nb.act);
    }
#pragma unroll
    for (int u = 0; u < 4; ++u) v[u] = apply_act(v[u], p.act_out);
    *reinterpret_cast<float4*>(p.y + base + e) = make_float4(v[0], v[1], v[2], v[3]);
  }
}

// ------------------------------------------------------------------------------------------------
// AvgPool2d(3, stride 2, padding 1, count_include_pad=False) on NHWC (networks.py:249-250, :525-526)
// ------------------------------------------------------------------------------------------------
static __global__ void avgpool3s2_nhwc_kernel(const float* __restrict__ x, float* __restrict__ y, int B, int H, int W, int C, int Ho, int Wo) {
  const size_t total = (size_t)B * Ho * Wo * C;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
    const int c = (int)(i % C);
    size_t r = i / C;
    const int ox = (int)(r % Wo); r /= Wo;
    const int oy = (int)(r % Ho);
    const int b = (int)(r / Ho);
    float s = 0.f; int n = 0;
    for (int dy = -1; dy <= 1; ++dy) {
      const int iy = oy * 2 + dy;
      if (iy < 0 || iy >= H) continue;
      for (int dx = -1; dx <= 1; ++dx) {
        const int ix = ox * 2 + dx;
        if (ix < 0 || ix >= W) continue;
        s += __ldg(x + (((size_t)b * H + iy) * W + ix) * C + c);
        ++n;
      }
    }
    y[i] = s / (float)n;
  }
}

// ------------------------------------------------------------------------------------------------
// BoTNet multi-head self-attention with absolute position embedding (bottleneck_transformer_pytorch 0.1.4
// `Attention`, called from networks.py:342-344):  out = softmax(scale*q.(k + emb)^T) v,
// emb[(y,x)] = height[y] + width[x].  One CTA per (sample, head); K+emb and V live in shared memory, one
// warp per query row: lanes own keys for the scores (shuffle max / sum softmax) and channels for P.V.
// qkv: NHWC [B, L, 3*heads*d] (q | k | v, channel = head*d + i); out: [B, L, heads*d].
// Optionally accumulates per-channel (sum, sumsq) of the output for the BatchNorm2d that follows.
// ------------------------------------------------------------------------------------------------
struct AttnParams {
  const float* qkv; const float* emb_h; const float* emb_w; float* out;
  int B, Hh, Ww, heads, d; float scale; double* stats;
};

template <int KPL>   // keys per lane = ceil(L / 32)
__global__ void __launch_bounds__(256) attention_abs_pos_kernel(const AttnParams p) {
  extern __shared__ float sm[];
  const int L = p.Hh * p.Ww, d = p.d, C = p.heads * d;
  const int b = blockIdx.x / p.heads, h = blockIdx.x % p.heads;
  float* Kp = sm;                       // [L][d+1]
  float* V = Kp + (size_t)L * (d + 1);  // [L][d]
  float* qrow = V + (size_t)L * d;      // [8 warps][d]
  const float* base = p.qkv + (size_t)b * L * 3 * C;
  for (int i = threadIdx.x; i < L * d; i += 256) {
    const int j = i / d, dd = i - j * d;
    const int y = j / p.Ww, x = j - y * p.Ww;
    const float* tok = base + (size_t)j * 3 * C + h * d + dd;
    Kp[j * (d + 1) + dd] = __ldg(tok + C) + __ldg(p.emb_h + y * d + dd) + __ldg(p.emb_w + x * d + dd);
    V[j * d + dd] = __ldg(tok + 2 * C);
  }
  __syncthreads();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  float* q = qrow + warp * d;
  const int dpl = d / 32;               // channels per lane (d % 32 == 0)
  double ssum[4] = {0.0, 0.0, 0.0, 0.0}, ssq[4] = {0.0, 0.0, 0.0, 0.0};
  for (int i = warp; i < L; i += 8) {
    for (int dd = lane; dd < d; dd += 32) q[dd] = __ldg(base + (size_t)i * 3 * C + h * d + dd) * p.scale;
    __syncwarp();
    float sc[KPL];
    float mx = -INFINITY;
#pragma unroll
    for (int u = 0; u < KPL; ++u) {
      const int j = lane + 32 * u;
      float a = -INFINITY;
      if (j < L) {
        a = 0.f;
        const float* kr = Kp + j * (d + 1);
        for (int dd = 0; dd < d; ++dd) a = fmaf(q[dd], kr[dd], a);
      }
      sc[u] = a;
      mx = fmaxf(mx, a);
    }
#pragma unroll
    for (int o = 16; o; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
    float sum = 0.f;
#pragma unroll
    for (int u = 0; u < KPL; ++u) {
      sc[u] = (lane + 32 * u < L) ? __expf(sc[u] - mx) : 0.f;
      sum += sc[u];
    }
#pragma unroll
    for (int o = 16; o; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
    const float inv = 1.f / sum;
    float acc[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
    for (int u = 0; u < KPL; ++u) {
      for (int l = 0; l < 32; ++l) {
        const int j = l + 32 * u;
        if (j >= L) break;
        const float pj = __shfl_sync(0xffffffffu, sc[u], l);
        for (int t = 0; t < dpl; ++t) acc[t] = fmaf(pj, V[j * d + lane + 32 * t], acc[t]);
      }
    }
    for (int t = 0; t < dpl; ++t) {
      const float o = acc[t] * inv;
      p.out[((size_t)b * L + i) * C + h * d + lane + 32 * t] = o;
      ssum[t] += (double)o; ssq[t] = fma((double)o, (double)o, ssq[t]);
    }
    __syncwarp();
  }
  if (p.stats) {
    for (int t = 0; t < dpl; ++t) {
      double* st = p.stats + ((size_t)b * C + h * d + lane + 32 * t) * 2;
      atomicAdd(st, ssum[t]);
      atomicAdd(st + 1, ssq[t]);
    }
  }
}

// ------------------------------------------------------------------------------------------------
// Inference residual of --fit_residual (pix2pixHD_model.py:631-635):
//   sr[..., :lr_bins] *= low_scale (1e-3);  sr += lr          on [rows, nbins] fp32 (rows = B*F)
// ------------------------------------------------------------------------------------------------
static __global__ void residual_scale_add_kernel(const float* __restrict__ sr, const float* __restrict__ lr, int64_t lr_row_stride,
                                          float* __restrict__ y, int64_t rows, int nbins, int lr_bins, float low_scale) {
  const size_t total = (size_t)rows * nbins;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
    const int k = (int)(i % nbins);
    const size_t r = i / nbins;
    const float v = sr[i];
    y[i] = (k < lr_bins ? v * low_scale : v) + lr[r * lr_row_stride + k];
  }
}

// NCHW <-> NHWC (network boundary: the reference's modules take / return NCHW tensors)
static __global__ void nchw_to_nhwc_kernel(const float* __restrict__ x, float* __restrict__ y, int B, int C, int HW) {
  const size_t total = (size_t)B * C * HW;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
    const int c = (int)(i % C);
    const size_t r = i / C;
    const int pix = (int)(r % HW);
    const int b = (int)(r / HW);
    y[i] = __ldg(x + ((size_t)b * C + c) * HW + pix);
  }
}
static __global__ void nhwc_to_nchw_kernel(const float* __restrict__ x, float* __restrict__ y, int B, int C, int HW) {
  const size_t total = (size_t)B * C * HW;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
    const int pix = (int)(i % HW);
    const size_t r = i / HW;
    const int c = (int)(r % C);
    const int b = (int)(r / C);
    y[i] = __ldg(x + ((size_t)b * HW + pix) * C + c);
  }
}

}  // namespace nnk
