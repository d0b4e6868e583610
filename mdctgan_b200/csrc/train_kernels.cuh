// train_kernels.cuh -- backward / loss / optimiser kernels of the GAN train step (sm_100a, fp32).
//
// Reference graph (all through torch autograd there): models/pix2pixHD_model.py:416-451 (_forward: three
// discriminator passes, LSGAN + feature-matching losses), train.py:175-202 (loss_G / loss_D backward, two Adam
// steps), models/networks.py:97-137 (GANLoss).  What autograd derives there is written out here:
//
//   conv_wgrad_kernel      dW, dbias of nn.Conv2d / nn.ConvTranspose2d: implicit GEMM  dW[k][co] = sum_m A[m][k]*dY[m][co]
//                          with A gathered exactly like the forward kernels do (reflection / zero padding, stride,
//                          transposed taps, the producer's deferred InstanceNorm / BatchNorm + activation), written with
//                          float atomics straight into the parameter-layout gradient buffer (the flat NCCL bucket).
//                          (dgrad needs no kernel of its own: it is the forward convolution on the transposed geometry
//                          with the same weights, nn_kernels.cuh / conv_umma.cuh.)
//   norm_bwd_*             backward of  v = act(norm(x))  for InstanceNorm2d(affine=False) and train-mode BatchNorm2d
//   act_bwd_kernel         backward of an epilogue activation (LeakyReLU of the first PatchGAN layer, tanh head)
//   reflect_fold_kernel    backward of nn.ReflectionPad2d
//   avgpool3s2_bwd_kernel  backward of AvgPool2d(3, 2, 1, count_include_pad=False)
//   attention_bwd_kernel   backward of the BoTNet attention (abs. position embedding)
//   mse_const_* / l1_pair_*   LSGAN (networks.py:127-137) and feature matching (pix2pixHD_model.py:447-451)
//   disc_input_*           cat(lr, s, |s|*2+lo) (pix2pixHD_model.py:420-427, :369) written directly as NHWC, and its backward
//   adam_flat_kernel       torch.optim.Adam (pix2pixHD_model.py:350-364) over one flat buffer
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "nn_kernels.cuh"

namespace trk {
using nnk::InputNorm;
using nnk::apply_act;
using nnk::in_coord;
using nnk::kActLeaky;
using nnk::kActNone;
using nnk::kActRelu;
using nnk::kActTanh;

__device__ __forceinline__ float act_grad_from_pre(float pre, int act) {   // d act(pre) / d pre   (ReLU / LeakyReLU / none)
  if (act == kActRelu) return pre > 0.f ? 1.f : 0.f;
  if (act == kActLeaky) return pre > 0.f ? 1.f : 0.2f;
  return 1.f;
}
__device__ __forceinline__ float act_grad_from_out(float y, int act) {     // same, from the activated value
  if (act == kActTanh) return 1.f - y * y;
  if (act == nnk::kActSigmoid) return y * (1.f - y);
  return act_grad_from_pre(y, act);   // ReLU / LeakyReLU keep the sign
}

// ------------------------------------------------------------------------------------------------
// Weight gradient.  grid = (k_tiles * n_tiles, B * chunks_per_sample); 256 threads; tile 64 (k) x 64 (co),
// 4x4 per thread; the reduction dimension (pixels of one sample chunk) is walked 16 at a time.
// ------------------------------------------------------------------------------------------------
struct WgradParams {
  const float* x; int B, H, W, Cin;
  InputNorm in;
  const float* dy; int Ho, Wo, Cout;
  int kh, kw, stride, pad, pad_mode, transposed;
  float* dw; long long s_co, s_ci, s_tap;   // element offset of dW[co][ci][tap] in the parameter's own layout
  float* dbias;
  int n_tiles, chunks_per_sample, pix_per_chunk;
};

// CTA tile 128 (k) x 64 (co), 8 x 4 outputs per thread; the pixel (reduction) dimension is walked 16 at a time through a
// double-buffered shared-memory stage: the global gather of step i+1 is in flight while step i is multiplied.
struct WgradALoad {            // one float4 slot of the A stage: 4 consecutive k of one pixel row
  int ky[4], kx[4], c[4];      // tap / channel of each k (c < 0: beyond K)
};

__device__ __forceinline__ void wgrad_decode(const WgradParams& p, int K, int kbase, WgradALoad& a) {
#pragma unroll
  for (int u = 0; u < 4; ++u) {
    const int k = kbase + u;
    if (k < K) {
      const int tap = k / p.Cin;
      a.c[u] = k - tap * p.Cin;
      a.ky[u] = tap / p.kw;
      a.kx[u] = tap - a.ky[u] * p.kw;
    } else { a.c[u] = -1; a.ky[u] = 0; a.kx[u] = 0; }
  }
}

__device__ __forceinline__ float4 wgrad_gather(const WgradParams& p, const float* xb, const WgradALoad& a, int oy, int ox, bool vec_a, bool has_norm,
                                                const float* s_scale, const float* s_shift) {
  float v[4] = {0.f, 0.f, 0.f, 0.f};
  if (vec_a) {
    if (a.c[0] >= 0) {
      const int iy = in_coord(oy, a.ky[0], p.H, p.stride, p.pad, p.pad_mode, p.transposed);
      const int ix = in_coord(ox, a.kx[0], p.W, p.stride, p.pad, p.pad_mode, p.transposed);
      if (iy >= 0 && ix >= 0) {
        const int c = a.c[0];
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
      if (a.c[u] >= 0) {
        const int iy = in_coord(oy, a.ky[u], p.H, p.stride, p.pad, p.pad_mode, p.transposed);
        const int ix = in_coord(ox, a.kx[u], p.W, p.stride, p.pad, p.pad_mode, p.transposed);
        if (iy >= 0 && ix >= 0) {
          float t = __ldg(xb + ((size_t)iy * p.W + ix) * p.Cin + a.c[u]);
          if (has_norm) t = apply_act(fmaf(t, s_scale[a.c[u]], s_shift[a.c[u]]), p.in.act);
          v[u] = t;
        }
      }
    }
  }
  return make_float4(v[0], v[1], v[2], v[3]);
}

template <typename AccT>
__global__ void __launch_bounds__(256) conv_wgrad_kernel(const WgradParams p) {
  constexpr int TK = 128, TN = 64, MS = 16;
  __shared__ __align__(16) float As[2][MS][TK + 4];
  __shared__ __align__(16) float Bs[2][MS][TN];
  __shared__ float s_scale[1024], s_shift[1024];
  const int tid = threadIdx.x;
  const int kt = blockIdx.x / p.n_tiles, nt = blockIdx.x - kt * p.n_tiles;
  const int HWo = p.Ho * p.Wo;
  const int K = p.kh * p.kw * p.Cin;
  const int k0 = kt * TK, n0 = nt * TN;
  const bool has_norm = p.in.scale != nullptr || p.in.stats != nullptr;
  // work items (sample, pixel chunk) of this CTA: a contiguous range, so the per-sample normalisation is reloaded rarely
  const int items = p.B * p.chunks_per_sample;
  const int item_begin = (int)((long long)items * blockIdx.y / gridDim.y), item_end = (int)((long long)items * (blockIdx.y + 1) / gridDim.y);
  const bool vec_a = (p.Cin % 4) == 0;
  const bool vec_b = (p.Cout % 4) == 0;
  // load roles: pixel slot pm; A: two float4 slots (k offsets kq, kq + 64); B: one float4 slot (co offset nq)
  const int pm = tid >> 4, kq = (tid & 15) * 4, nq = (tid & 15) * 4;
  WgradALoad a0, a1;
  wgrad_decode(p, K, k0 + kq, a0);
  wgrad_decode(p, K, k0 + kq + 64, a1);
  const int ty = tid >> 4, tx = tid & 15;        // outputs: k rows ty*4 + {0..3} and 64 + ty*4 + {0..3}; co columns tx*4 + {0..3}
  AccT acc[8][4];
#pragma unroll
  for (int i = 0; i < 8; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = (AccT)0;
  float bsum[4] = {0.f, 0.f, 0.f, 0.f};

  int b_loaded = -1;
  for (int item = item_begin; item < item_end; ++item) {
    const int b = item / p.chunks_per_sample;
    const int chunk = item - b * p.chunks_per_sample;
    const int m_begin = chunk * p.pix_per_chunk;
    const int m_end = min(HWo, m_begin + p.pix_per_chunk);
    if (m_begin >= m_end) continue;
    if (has_norm && b != b_loaded) {
      __syncthreads();
      nnk::norm_to_smem(p.in, b, p.Cin, s_scale, s_shift, tid, 256);
      b_loaded = b;
    }
    __syncthreads();                                  // scale / shift visible; the stage buffers of the previous item are free
    const float* xb = p.x + (size_t)b * p.H * p.W * p.Cin;
    const float* dyb = p.dy + (size_t)b * HWo * p.Cout;

    auto load_stage = [&](int mb, float4& ra0, float4& ra1, float4& rb) {
      const int m = mb + pm;
      ra0 = ra1 = rb = make_float4(0.f, 0.f, 0.f, 0.f);
      if (m < m_end) {
        const int oy = m / p.Wo, ox = m - oy * p.Wo;
        ra0 = wgrad_gather(p, xb, a0, oy, ox, vec_a, has_norm, s_scale, s_shift);
        ra1 = wgrad_gather(p, xb, a1, oy, ox, vec_a, has_norm, s_scale, s_shift);
        const int n = n0 + nq;
        const float* dr = dyb + (size_t)m * p.Cout + n;
        if (vec_b && n + 3 < p.Cout) {
          rb = __ldg(reinterpret_cast<const float4*>(dr));
        } else {
          if (n + 0 < p.Cout) rb.x = __ldg(dr + 0);
          if (n + 1 < p.Cout) rb.y = __ldg(dr + 1);
          if (n + 2 < p.Cout) rb.z = __ldg(dr + 2);
          if (n + 3 < p.Cout) rb.w = __ldg(dr + 3);
        }
      }
    };
    float4 ra0, ra1, rb;
    load_stage(m_begin, ra0, ra1, rb);
    *reinterpret_cast<float4*>(&As[0][pm][kq]) = ra0;
    *reinterpret_cast<float4*>(&As[0][pm][kq + 64]) = ra1;
    *reinterpret_cast<float4*>(&Bs[0][pm][nq]) = rb;
    __syncthreads();
    int buf = 0;
    for (int mb = m_begin; mb < m_end; mb += MS) {
      const bool more = mb + MS < m_end;
      if (more) load_stage(mb + MS, ra0, ra1, rb);   // in flight while this stage is multiplied
#pragma unroll 4
      for (int mm = 0; mm < MS; ++mm) {
        const float4 al = *reinterpret_cast<const float4*>(&As[buf][mm][ty * 4]);
        const float4 ah = *reinterpret_cast<const float4*>(&As[buf][mm][64 + ty * 4]);
        const float4 bv = *reinterpret_cast<const float4*>(&Bs[buf][mm][tx * 4]);
        const float av[8] = {al.x, al.y, al.z, al.w, ah.x, ah.y, ah.z, ah.w}, bw[4] = {bv.x, bv.y, bv.z, bv.w};
#pragma unroll
        for (int i = 0; i < 8; ++i)
#pragma unroll
          for (int j = 0; j < 4; ++j) acc[i][j] = fma((AccT)av[i], (AccT)bw[j], acc[i][j]);
        if (ty == 0) {
#pragma unroll
          for (int j = 0; j < 4; ++j) bsum[j] += bw[j];
        }
      }
      if (more) {
        *reinterpret_cast<float4*>(&As[buf ^ 1][pm][kq]) = ra0;
        *reinterpret_cast<float4*>(&As[buf ^ 1][pm][kq + 64]) = ra1;
        *reinterpret_cast<float4*>(&Bs[buf ^ 1][pm][nq]) = rb;
      }
      __syncthreads();
      buf ^= 1;
    }
  }   // work items
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int k = k0 + (i < 4 ? ty * 4 + i : 64 + ty * 4 + (i - 4));
    if (k >= K) continue;
    const int tap = k / p.Cin, ci = k - tap * p.Cin;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int co = n0 + tx * 4 + j;
      if (co < p.Cout) {
        float* dst = p.dw + (long long)co * p.s_co + (long long)ci * p.s_ci + (long long)tap * p.s_tap;
        atomicAdd(dst, (float)acc[i][j]);   // fire-and-forget reduction at L2 (a read-modify-write of these scattered addresses is slower)
      }
    }
  }
  if (p.dbias && kt == 0 && ty == 0) {
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int co = n0 + tx * 4 + j;
      if (co < p.Cout) atomicAdd(p.dbias + co, bsum[j]);
    }
  }
}

// Cout <= 4 layers (generator head 7x7 -> 1, PatchGAN heads 4x4 -> 1): dW[k][co] = sum_m A[m][k] * dy[m][co].
// grid = (ceil(K / 256), B * chunks); each thread owns one k, walks the chunk's pixels (dy broadcast from shared memory).
__global__ void __launch_bounds__(256) conv_wgrad_small_cout_kernel(const WgradParams p) {
  __shared__ float s_scale[1024], s_shift[1024];
  __shared__ float s_dy[256][4];
  const int tid = threadIdx.x;
  const int b = blockIdx.y / p.chunks_per_sample;
  const int chunk = blockIdx.y - b * p.chunks_per_sample;
  const int HWo = p.Ho * p.Wo;
  const int m_begin = chunk * p.pix_per_chunk;
  const int m_end = min(HWo, m_begin + p.pix_per_chunk);
  const int K = p.kh * p.kw * p.Cin;
  const bool has_norm = p.in.scale != nullptr || p.in.stats != nullptr;
  if (has_norm) nnk::norm_to_smem(p.in, b, p.Cin, s_scale, s_shift, tid, 256);
  const int k = blockIdx.x * 256 + tid;
  int ky = 0, kx = 0, c = -1;
  if (k < K) { const int tap = k / p.Cin; c = k - tap * p.Cin; ky = tap / p.kw; kx = tap - ky * p.kw; }
  __syncthreads();
  const float a_s = (has_norm && c >= 0) ? s_scale[c] : 1.f, a_t = (has_norm && c >= 0) ? s_shift[c] : 0.f;
  const float* xb = p.x + (size_t)b * p.H * p.W * p.Cin;
  const float* dyb = p.dy + (size_t)b * HWo * p.Cout;
  float acc[4] = {0.f, 0.f, 0.f, 0.f}, bsum[4] = {0.f, 0.f, 0.f, 0.f};
  for (int mb = m_begin; mb < m_end; mb += 256) {
    const int mload = mb + tid;
#pragma unroll
    for (int j = 0; j < 4; ++j) s_dy[tid][j] = (mload < m_end && j < p.Cout) ? __ldg(dyb + (size_t)mload * p.Cout + j) : 0.f;
    __syncthreads();
    const int cnt = min(256, m_end - mb);
    if (c >= 0) {
      int oy = mb / p.Wo, ox = mb - oy * p.Wo;
      for (int t = 0; t < cnt; ++t) {
        const int iy = in_coord(oy, ky, p.H, p.stride, p.pad, p.pad_mode, p.transposed);
        const int ix = in_coord(ox, kx, p.W, p.stride, p.pad, p.pad_mode, p.transposed);
        if (iy >= 0 && ix >= 0) {
          float v = __ldg(xb + ((size_t)iy * p.W + ix) * p.Cin + c);
          if (has_norm) v = apply_act(fmaf(v, a_s, a_t), p.in.act);
#pragma unroll
          for (int j = 0; j < 4; ++j) acc[j] = fmaf(v, s_dy[t][j], acc[j]);
        }
        if (++ox == p.Wo) { ox = 0; ++oy; }
      }
    }
    if (p.dbias && blockIdx.x == 0 && tid < 4) {
      for (int t = 0; t < cnt; ++t) bsum[0] += s_dy[t][tid];
    }
    __syncthreads();
  }
  if (c >= 0) {
    const int tap = k / p.Cin;
    for (int j = 0; j < p.Cout; ++j) atomicAdd(p.dw + (long long)j * p.s_co + (long long)c * p.s_ci + (long long)tap * p.s_tap, acc[j]);
  }
  if (p.dbias && blockIdx.x == 0 && tid < p.Cout) atomicAdd(p.dbias + tid, bsum[0]);
}

// Cout == 1, stride 1, Cin % 4 == 0 (generator head 7x7, PatchGAN heads 4x4) with register reuse along the row.  The kernel above walks
// its pixels with one scalar load and ~15 instructions per multiply-add (187 us on the cfg4 generator head, 6 ms on the train.sh one).
// Here a thread owns (tap row ky, channel quad) for a segment of kSegW output columns of one output row: it walks the padded input
// columns t of that segment once, loading x[iy(ky)][ix(t)] (float4, normalised + activated on the fly) and keeping the last KW values
// of dy in registers, so that one load feeds KW * 4 multiply-adds:  dW[ky][kx][c] += x[t][c] * dy[t - kx].
// Partial sums meet in shared memory per CTA, then one global atomic per element and CTA.  grid.x = ceil(units * G * kh / 256),
// unit = (sample, output row, column segment), G = Cin / 4; dynamic shared memory kh * KW * Cin floats.
constexpr int kWgSegW = 64;
template <int KW>
__global__ void __launch_bounds__(256) conv_wgrad_cout1_kernel(const WgradParams p) {
  extern __shared__ float s_part[];              // [kh][KW][Cin]
  const int G = p.Cin / 4, per_unit = G * p.kh;
  const int nelem = p.kh * KW * p.Cin;
  for (int i = threadIdx.x; i < nelem; i += 256) s_part[i] = 0.f;
  __syncthreads();
  const int segs = (p.Wo + kWgSegW - 1) / kWgSegW;
  const long long units = (long long)p.B * p.Ho * segs;
  const long long gid = (long long)blockIdx.x * 256 + threadIdx.x;
  const long long unit = gid / per_unit;
  if (unit < units) {
    const int idx = (int)(gid - unit * per_unit);
    const int cq = idx % G, ky = idx / G, c = cq * 4;
    const int seg = (int)(unit % segs);
    const long long row = unit / segs;
    const int oy = (int)(row % p.Ho), b = (int)(row / p.Ho);
    const int x0 = seg * kWgSegW, x1 = min(p.Wo, x0 + kWgSegW);
    const float* dyr = p.dy + ((size_t)b * p.Ho + oy) * p.Wo;
    if (p.dbias && idx == 0) {
      float sb = 0.f;
      for (int ox = x0; ox < x1; ++ox) sb += __ldg(dyr + ox);
      atomicAdd(p.dbias, sb);
    }
    const int iy = in_coord(oy, ky, p.H, 1, p.pad, p.pad_mode, 0);
    if (iy >= 0) {
      // deferred normalisation of the thread's four channels (sample b)
      float4 sc = make_float4(1.f, 1.f, 1.f, 1.f), sh = make_float4(0.f, 0.f, 0.f, 0.f);
      const bool has_norm = p.in.scale != nullptr || p.in.stats != nullptr;
      if (p.in.scale) {
        const size_t o = (p.in.per_sample ? (size_t)b * p.Cin : 0) + c;
        sc = __ldg(reinterpret_cast<const float4*>(p.in.scale + o));
        sh = __ldg(reinterpret_cast<const float4*>(p.in.shift + o));
      } else if (p.in.stats) {
        float scv[4], shv[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          const size_t o = 2 * ((size_t)b * p.Cin + c + u);
          const double mean = p.in.stats[o] / (double)p.in.count;
          double var = p.in.stats[o + 1] / (double)p.in.count - mean * mean;
          var = var < 0.0 ? 0.0 : var;
          const double rstd = rsqrt(var + (double)p.in.eps);
          scv[u] = (float)rstd; shv[u] = (float)(-mean * rstd);
        }
        sc = make_float4(scv[0], scv[1], scv[2], scv[3]); sh = make_float4(shv[0], shv[1], shv[2], shv[3]);
      }
      const float slope = !has_norm ? 1.f : (p.in.act == kActRelu ? 0.f : (p.in.act == kActLeaky ? 0.2f : 1.f));
      const float* xrow = p.x + (((size_t)b * p.H + iy) * p.W) * p.Cin + c;
      float acc[KW][4];
      float dyw[KW];
#pragma unroll
      for (int k = 0; k < KW; ++k) { dyw[k] = 0.f; acc[k][0] = acc[k][1] = acc[k][2] = acc[k][3] = 0.f; }
#pragma unroll 2
      for (int t = x0; t < x1 + KW - 1; ++t) {           // padded column t feeds the outputs ox = t - kx inside [x0, x1)
#pragma unroll
        for (int k = KW - 1; k > 0; --k) dyw[k] = dyw[k - 1];
        dyw[0] = t < x1 ? __ldg(dyr + t) : 0.f;
        const int ix = in_coord(t, 0, p.W, 1, p.pad, p.pad_mode, 0);
        if (ix >= 0) {
          const float4 q = __ldg(reinterpret_cast<const float4*>(xrow + (size_t)ix * p.Cin));
          float e0 = fmaf(q.x, sc.x, sh.x), e1 = fmaf(q.y, sc.y, sh.y), e2 = fmaf(q.z, sc.z, sh.z), e3 = fmaf(q.w, sc.w, sh.w);
          e0 = fmaxf(e0, slope * e0); e1 = fmaxf(e1, slope * e1); e2 = fmaxf(e2, slope * e2); e3 = fmaxf(e3, slope * e3);
#pragma unroll
          for (int k = 0; k < KW; ++k) {
            acc[k][0] = fmaf(e0, dyw[k], acc[k][0]); acc[k][1] = fmaf(e1, dyw[k], acc[k][1]);
            acc[k][2] = fmaf(e2, dyw[k], acc[k][2]); acc[k][3] = fmaf(e3, dyw[k], acc[k][3]);
          }
        }
      }
#pragma unroll
      for (int k = 0; k < KW; ++k)
#pragma unroll
        for (int u = 0; u < 4; ++u) atomicAdd(&s_part[(ky * KW + k) * p.Cin + c + u], acc[k][u]);
    }
  }
  __syncthreads();
  for (int i = threadIdx.x; i < nelem; i += 256) {
    const float v = s_part[i];
    if (v != 0.f) {
      const int tap = i / p.Cin, ci = i - tap * p.Cin;
      atomicAdd(p.dw + (long long)ci * p.s_ci + (long long)tap * p.s_tap, v);
    }
  }
}

// ------------------------------------------------------------------------------------------------
// Backward of v = act(norm(x)), norm = InstanceNorm2d(affine=False) (mode 0) or train-mode BatchNorm2d (mode 1):
//   pre = (x - mean) * rstd * gamma + beta,  g = dv * act'(pre),  xhat = (x - mean) * rstd
//   dx = gamma * rstd * (g - mean(g) - xhat * mean(g * xhat))       (means over the plane / over batch x plane)
// Pass 1 reduces (sum g, sum g*xhat) per (b, c); pass 2 applies.
// ------------------------------------------------------------------------------------------------
struct NormBwdParams {
  const float* x; const float* dv; float* dx;
  const double* stats; double count; float eps; int mode;
  const float* gamma; const float* beta;
  int act;
  double* red;                 // [B][C][2]
  float* dgamma; float* dbeta; // BatchNorm parameter gradients (accumulated), nullable
  int B, HW, C, stats_B;       // stats_B: batch size the forward statistics were taken over (BatchNorm; >= B)
  // fold_pad > 0: `dv` is the gradient of the ReflectionPad2d(fold_pad)-ed view, [B][H + 2p][W + 2p][C] (the raw output of the dgrad
  // convolution); the fold back onto [H][W] (reflect_fold_kernel) happens while loading -- one launch less on the dgrad chain
  int fold_pad, H, W;
};

// float4 of dv for (sample b, pixel pix, channels c .. c+3), folded when the gradient arrives in the padded geometry
__device__ __forceinline__ float4 load_dv4(const NormBwdParams& p, int b, int pix, int c) {
  if (p.fold_pad == 0) return __ldg(reinterpret_cast<const float4*>(p.dv + ((size_t)b * p.HW + pix) * p.C + c));
  const int P = p.fold_pad, Hp = p.H + 2 * P, Wp = p.W + 2 * P;
  const int iy = pix / p.W, ix = pix - iy * p.W;
  int ys[3], xs[3], ny = 0, nx = 0;
  ys[ny++] = iy;
  if (iy >= 1 && iy <= P) ys[ny++] = -iy;
  if (p.H - 1 - iy >= 1 && p.H - 1 - iy <= P) ys[ny++] = 2 * (p.H - 1) - iy;
  xs[nx++] = ix;
  if (ix >= 1 && ix <= P) xs[nx++] = -ix;
  if (p.W - 1 - ix >= 1 && p.W - 1 - ix <= P) xs[nx++] = 2 * (p.W - 1) - ix;
  float4 s = make_float4(0.f, 0.f, 0.f, 0.f);
  for (int a = 0; a < ny; ++a)
    for (int q = 0; q < nx; ++q) {
      const float4 t = __ldg(reinterpret_cast<const float4*>(p.dv + (((size_t)b * Hp + ys[a] + P) * Wp + xs[q] + P) * p.C + c));
      s.x += t.x; s.y += t.y; s.z += t.z; s.w += t.w;
    }
  return s;
}

// mean / rstd (and the forward affine) of sample b into shared memory
__device__ __forceinline__ void norm_bwd_coeffs(const NormBwdParams& p, int b, float* s_mean, float* s_rstd, float* s_gam, float* s_bet) {
  for (int c = threadIdx.x; c < p.C; c += blockDim.x) {
    double s = 0, q = 0, n = p.count;
    if (p.mode == 0) { s = p.stats[2 * ((size_t)b * p.C + c)]; q = p.stats[2 * ((size_t)b * p.C + c) + 1]; }
    else {
      for (int bb = 0; bb < p.stats_B; ++bb) { s += p.stats[2 * ((size_t)bb * p.C + c)]; q += p.stats[2 * ((size_t)bb * p.C + c) + 1]; }
      n = p.count * p.stats_B;
    }
    const double mean = s / n;
    double var = q / n - mean * mean;
    if (var < 0) var = 0;
    s_mean[c] = (float)mean;
    s_rstd[c] = (float)(rsqrt(var + (double)p.eps));
    s_gam[c] = (p.mode == 1 && p.gamma) ? __ldg(p.gamma + c) : 1.f;
    s_bet[c] = (p.mode == 1 && p.beta) ? __ldg(p.beta + c) : 0.f;
  }
}

__global__ void __launch_bounds__(256) norm_bwd_reduce_kernel(const NormBwdParams p) {
  __shared__ float s_mean[1024], s_rstd[1024], s_gam[1024], s_bet[1024];
  __shared__ double s_acc[2][1024];   // fp64: the apply pass subtracts mean(g) from g, and g is often nearly constant over a plane
  const int b = blockIdx.y;
  norm_bwd_coeffs(p, b, s_mean, s_rstd, s_gam, s_bet);
  for (int c = threadIdx.x; c < p.C; c += 256) { s_acc[0][c] = 0.0; s_acc[1][c] = 0.0; }
  __syncthreads();
  const int groups = p.C / 4;                    // float4 channel groups; C <= 1024 -> groups <= 256
  const int cg = threadIdx.x % groups, prow = threadIdx.x / groups, pstep = 256 / groups;
  double sg[4] = {0.0, 0.0, 0.0, 0.0}, sq[4] = {0.0, 0.0, 0.0, 0.0};
  if (prow < pstep) {
    const size_t base = (size_t)b * p.HW * p.C;
    for (int pix = blockIdx.x * pstep + prow; pix < p.HW; pix += gridDim.x * pstep) {
      const size_t e = base + (size_t)pix * p.C + cg * 4;
      const float4 xv = __ldg(reinterpret_cast<const float4*>(p.x + e));
      const float4 dv = load_dv4(p, b, pix, cg * 4);
      const float xs[4] = {xv.x, xv.y, xv.z, xv.w}, ds[4] = {dv.x, dv.y, dv.z, dv.w};
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const int c = cg * 4 + u;
        const float xhat = (xs[u] - s_mean[c]) * s_rstd[c];
        const float g = ds[u] * act_grad_from_pre(fmaf(xhat, s_gam[c], s_bet[c]), p.act);
        sg[u] += (double)g; sq[u] = fma((double)g, (double)xhat, sq[u]);
      }
    }
#pragma unroll
    for (int u = 0; u < 4; ++u) { atomicAdd(&s_acc[0][cg * 4 + u], sg[u]); atomicAdd(&s_acc[1][cg * 4 + u], sq[u]); }
  }
  __syncthreads();
  for (int c = threadIdx.x; c < p.C; c += 256) {
    double* r = p.red + 2 * ((size_t)b * p.C + c);
    atomicAdd(r, s_acc[0][c]);
    atomicAdd(r + 1, s_acc[1][c]);
  }
}

__global__ void __launch_bounds__(256) norm_bwd_apply_kernel(const NormBwdParams p) {
  __shared__ float s_mean[1024], s_rstd[1024], s_gam[1024], s_bet[1024];
  __shared__ float s_mg[1024], s_mgx[1024];
  const int b = blockIdx.y;
  norm_bwd_coeffs(p, b, s_mean, s_rstd, s_gam, s_bet);
  for (int c = threadIdx.x; c < p.C; c += 256) {
    double sg = 0, sq = 0, n = p.count;
    if (p.mode == 0) { sg = p.red[2 * ((size_t)b * p.C + c)]; sq = p.red[2 * ((size_t)b * p.C + c) + 1]; }
    else {
      for (int bb = 0; bb < p.B; ++bb) { sg += p.red[2 * ((size_t)bb * p.C + c)]; sq += p.red[2 * ((size_t)bb * p.C + c) + 1]; }
      n = p.count * p.B;
      if (blockIdx.x == 0 && b == 0) {
        if (p.dgamma) atomicAdd(p.dgamma + c, (float)sq);
        if (p.dbeta) atomicAdd(p.dbeta + c, (float)sg);
      }
    }
    s_mg[c] = (float)(sg / n);
    s_mgx[c] = (float)(sq / n);
  }
  __syncthreads();
  const size_t per_sample4 = (size_t)p.HW * p.C / 4;
  const size_t base = (size_t)b * p.HW * p.C;
  for (size_t i = (size_t)blockIdx.x * 256 + threadIdx.x; i < per_sample4; i += (size_t)gridDim.x * 256) {
    const size_t e = i * 4;
    const int c0 = (int)(e % p.C);
    const float4 xv = __ldg(reinterpret_cast<const float4*>(p.x + base + e));
    const float4 dv = load_dv4(p, b, (int)(e / p.C), c0);
    const float xs[4] = {xv.x, xv.y, xv.z, xv.w}, ds[4] = {dv.x, dv.y, dv.z, dv.w};
    float o[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const int c = c0 + u;
      const float xhat = (xs[u] - s_mean[c]) * s_rstd[c];
      const float g = ds[u] * act_grad_from_pre(fmaf(xhat, s_gam[c], s_bet[c]), p.act);
      o[u] = s_gam[c] * s_rstd[c] * (g - s_mg[c] - xhat * s_mgx[c]);
    }
    *reinterpret_cast<float4*>(p.dx + base + e) = make_float4(o[0], o[1], o[2], o[3]);
  }
}

// InstanceNorm2d backward in ONE launch: a CTA owns 8 channels of one sample (one 32-byte sector per pixel), reduces
// (sum g, sum g*xhat) over the plane, then applies -- the second read of x / dv hits L1 / L2.  grid = (C / 8, B).
__global__ void __launch_bounds__(256) instnorm_bwd_fused_kernel(const NormBwdParams p) {
  __shared__ double s_red[8][2][8];     // [warp][sum g | sum g*xhat][channel of the group]
  __shared__ float s_m[2][8];
  const int b = blockIdx.y, c0 = blockIdx.x * 8;
  const int cl = threadIdx.x & 1, pl = threadIdx.x >> 1;          // float4 half of the group, pixel lane
  const int cb = c0 + cl * 4;
  __shared__ float s_stat[2][8];      // mean / rstd of the CTA's 8 channels: derived once (8 threads), not by all 256 in fp64
  if (threadIdx.x < 8) {
    const double s = p.stats[2 * ((size_t)b * p.C + c0 + threadIdx.x)], q = p.stats[2 * ((size_t)b * p.C + c0 + threadIdx.x) + 1];
    const double m = s / p.count;
    double var = q / p.count - m * m;
    if (var < 0) var = 0;
    s_stat[0][threadIdx.x] = (float)m;
    s_stat[1][threadIdx.x] = (float)(rsqrt(var + (double)p.eps));
  }
  __syncthreads();
  float mean[4], rstd[4];
#pragma unroll
  for (int u = 0; u < 4; ++u) { mean[u] = s_stat[0][cl * 4 + u]; rstd[u] = s_stat[1][cl * 4 + u]; }
  const size_t base = (size_t)b * p.HW * p.C + cb;
  double sg[4] = {0.0, 0.0, 0.0, 0.0}, sq[4] = {0.0, 0.0, 0.0, 0.0};
  for (int pix = pl; pix < p.HW; pix += 128) {
    const float4 xv = __ldg(reinterpret_cast<const float4*>(p.x + base + (size_t)pix * p.C));
    const float4 dv = load_dv4(p, b, pix, cb);
    const float xs[4] = {xv.x, xv.y, xv.z, xv.w}, ds[4] = {dv.x, dv.y, dv.z, dv.w};
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const float xhat = (xs[u] - mean[u]) * rstd[u];
      const float g = ds[u] * act_grad_from_pre(xhat, p.act);
      sg[u] += (double)g; sq[u] = fma((double)g, (double)xhat, sq[u]);
    }
  }
#pragma unroll
  for (int u = 0; u < 4; ++u) {
#pragma unroll
    for (int o = 16; o >= 2; o >>= 1) { sg[u] += __shfl_xor_sync(0xffffffffu, sg[u], o); sq[u] += __shfl_xor_sync(0xffffffffu, sq[u], o); }
  }
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (lane < 2) {
#pragma unroll
    for (int u = 0; u < 4; ++u) { s_red[warp][0][lane * 4 + u] = sg[u]; s_red[warp][1][lane * 4 + u] = sq[u]; }
  }
  __syncthreads();
  if (threadIdx.x < 16) {
    const int which = threadIdx.x >> 3, ch = threadIdx.x & 7;
    double t = 0;
#pragma unroll
    for (int w = 0; w < 8; ++w) t += s_red[w][which][ch];
    s_m[which][ch] = (float)(t / p.count);
  }
  __syncthreads();
  float mg[4], mgx[4];
#pragma unroll
  for (int u = 0; u < 4; ++u) { mg[u] = s_m[0][cl * 4 + u]; mgx[u] = s_m[1][cl * 4 + u]; }
  for (int pix = pl; pix < p.HW; pix += 128) {
    const float4 xv = __ldg(reinterpret_cast<const float4*>(p.x + base + (size_t)pix * p.C));
    const float4 dv = load_dv4(p, b, pix, cb);
    const float xs[4] = {xv.x, xv.y, xv.z, xv.w}, ds[4] = {dv.x, dv.y, dv.z, dv.w};
    float o[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const float xhat = (xs[u] - mean[u]) * rstd[u];
      const float g = ds[u] * act_grad_from_pre(xhat, p.act);
      o[u] = rstd[u] * (g - mg[u] - xhat * mgx[u]);
    }
    *reinterpret_cast<float4*>(p.dx + base + (size_t)pix * p.C) = make_float4(o[0], o[1], o[2], o[3]);
  }
}

// (sum, sumsq) per (sample, channel) of a materialised NHWC tensor (the sum of the two branches of ConvResBlock /
// InterpolateUpsample, networks.py:387-417, which an InstanceNorm2d follows).  grid = (chunks, B); stats += (zero it first).
__global__ void __launch_bounds__(256) plane_stats_kernel(const float* __restrict__ x, int HW, int C, double* __restrict__ stats) {
  __shared__ double s_acc[2][1024];
  const int b = blockIdx.y;
  for (int c = threadIdx.x; c < C; c += 256) { s_acc[0][c] = 0.0; s_acc[1][c] = 0.0; }
  __syncthreads();
  const int groups = C / 4;
  const int cg = threadIdx.x % groups, prow = threadIdx.x / groups, pstep = 256 / groups;
  double sg[4] = {0.0, 0.0, 0.0, 0.0}, sq[4] = {0.0, 0.0, 0.0, 0.0};
  if (prow < pstep) {
    const size_t base = (size_t)b * HW * C;
    for (int pix = blockIdx.x * pstep + prow; pix < HW; pix += gridDim.x * pstep) {
      const float4 v = __ldg(reinterpret_cast<const float4*>(x + base + (size_t)pix * C + cg * 4));
      const double vs[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
      for (int u = 0; u < 4; ++u) { sg[u] += vs[u]; sq[u] = fma(vs[u], vs[u], sq[u]); }
    }
#pragma unroll
    for (int u = 0; u < 4; ++u) { atomicAdd(&s_acc[0][cg * 4 + u], sg[u]); atomicAdd(&s_acc[1][cg * 4 + u], sq[u]); }
  }
  __syncthreads();
  for (int c = threadIdx.x; c < C; c += 256) {
    double* r = stats + 2 * ((size_t)b * C + c);
    atomicAdd(r, s_acc[0][c]);
    atomicAdd(r + 1, s_acc[1][c]);
  }
}

// F.interpolate(scale_factor=2.0, mode="nearest") on NHWC (networks.py:396) and its backward (sum of the 2x2 block)
__global__ void upsample_nearest2x_kernel(const float* __restrict__ x, float* __restrict__ y, int B, int H, int W, int C4) {
  const size_t total = (size_t)B * 2 * H * 2 * W * C4;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
    const int c = (int)(i % C4);
    size_t r = i / C4;
    const int ox = (int)(r % (2 * W)); r /= 2 * W;
    const int oy = (int)(r % (2 * H));
    const int b = (int)(r / (2 * H));
    reinterpret_cast<float4*>(y)[i] = __ldg(reinterpret_cast<const float4*>(x) + (((size_t)b * H + (oy >> 1)) * W + (ox >> 1)) * C4 + c);
  }
}
__global__ void upsample_nearest2x_bwd_kernel(const float* __restrict__ dy, float* __restrict__ dx, int B, int H, int W, int C4) {
  const size_t total = (size_t)B * H * W * C4;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
    const int c = (int)(i % C4);
    size_t r = i / C4;
    const int ix = (int)(r % W); r /= W;
    const int iy = (int)(r % H);
    const int b = (int)(r / H);
    const float4* src = reinterpret_cast<const float4*>(dy);
    const size_t row0 = (((size_t)b * 2 * H + 2 * iy) * 2 * W + 2 * ix) * C4 + c, row1 = row0 + (size_t)2 * W * C4;
    const float4 a = __ldg(src + row0), b4 = __ldg(src + row0 + C4), c4 = __ldg(src + row1), d4 = __ldg(src + row1 + C4);
    reinterpret_cast<float4*>(dx)[i] = make_float4(a.x + b4.x + c4.x + d4.x, a.y + b4.y + c4.y + d4.y, a.z + b4.z + c4.z + d4.z, a.w + b4.w + c4.w + d4.w);
  }
}

// g = dy * act'(y)  from the ACTIVATED value y (epilogue activations; also plain ReLU / LeakyReLU views)
__global__ void act_bwd_kernel(const float* __restrict__ dy, const float* __restrict__ y, float* __restrict__ g, size_t n, int act) {
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x)
    g[i] = dy[i] * act_grad_from_out(y[i], act);
}

// y = a + b
__global__ void add_kernel(const float* __restrict__ a, const float* __restrict__ b, float* __restrict__ y, size_t n) {
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) y[i] = a[i] + b[i];
}

// nn.ReflectionPad2d backward: dpad [B][H+2p][W+2p][C] -> dx [B][H][W][C] (every padded position has exactly one source)
__global__ void reflect_fold_kernel(const float* __restrict__ dpad, float* __restrict__ dx, int B, int H, int W, int C, int pad) {
  const size_t total = (size_t)B * H * W * C;
  const int Hp = H + 2 * pad, Wp = W + 2 * pad;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
    const int c = (int)(i % C);
    size_t r = i / C;
    const int ix = (int)(r % W); r /= W;
    const int iy = (int)(r % H);
    const int b = (int)(r / H);
    int ys[3], xs[3], ny = 0, nx = 0;
    ys[ny++] = iy;
    if (iy >= 1 && iy <= pad) ys[ny++] = -iy;
    if (H - 1 - iy >= 1 && H - 1 - iy <= pad) ys[ny++] = 2 * (H - 1) - iy;
    xs[nx++] = ix;
    if (ix >= 1 && ix <= pad) xs[nx++] = -ix;
    if (W - 1 - ix >= 1 && W - 1 - ix <= pad) xs[nx++] = 2 * (W - 1) - ix;
    float s = 0.f;
    for (int a = 0; a < ny; ++a)
      for (int q = 0; q < nx; ++q) s += __ldg(dpad + (((size_t)b * Hp + ys[a] + pad) * Wp + xs[q] + pad) * C + c);
    dx[i] = s;
  }
}

// the same fold fused with the accumulation into an existing gradient of the unpadded tensor: dx = fold(dpad) + other
__global__ void reflect_fold_add_kernel(const float* __restrict__ dpad, const float* __restrict__ other, float* __restrict__ dx, int B, int H, int W,
                                        int C4, int pad) {
  const size_t total = (size_t)B * H * W * C4;
  const int Hp = H + 2 * pad, Wp = W + 2 * pad;
  const float4* d4 = reinterpret_cast<const float4*>(dpad);
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
    const int c = (int)(i % C4);
    size_t r = i / C4;
    const int ix = (int)(r % W); r /= W;
    const int iy = (int)(r % H);
    const int b = (int)(r / H);
    int ys[3], xs[3], ny = 0, nx = 0;
    ys[ny++] = iy;
    if (iy >= 1 && iy <= pad) ys[ny++] = -iy;
    if (H - 1 - iy >= 1 && H - 1 - iy <= pad) ys[ny++] = 2 * (H - 1) - iy;
    xs[nx++] = ix;
    if (ix >= 1 && ix <= pad) xs[nx++] = -ix;
    if (W - 1 - ix >= 1 && W - 1 - ix <= pad) xs[nx++] = 2 * (W - 1) - ix;
    float4 s = __ldg(reinterpret_cast<const float4*>(other) + i);
    for (int a = 0; a < ny; ++a)
      for (int q = 0; q < nx; ++q) {
        const float4 t = __ldg(d4 + (((size_t)b * Hp + ys[a] + pad) * Wp + xs[q] + pad) * C4 + c);
        s.x += t.x; s.y += t.y; s.z += t.z; s.w += t.w;
      }
    reinterpret_cast<float4*>(dx)[i] = s;
  }
}

// AvgPool2d(3, 2, 1, count_include_pad=False) backward
__global__ void avgpool3s2_bwd_kernel(const float* __restrict__ dy, float* __restrict__ dx, int B, int H, int W, int C, int Ho, int Wo) {
  const size_t total = (size_t)B * H * W * C;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
    const int c = (int)(i % C);
    size_t r = i / C;
    const int ix = (int)(r % W); r /= W;
    const int iy = (int)(r % H);
    const int b = (int)(r / H);
    float s = 0.f;
    for (int oy = (iy - 1 + 1) / 2; oy <= (iy + 1) / 2; ++oy) {      // windows [2oy-1, 2oy+1] containing iy
      if (oy < 0 || oy >= Ho || 2 * oy - 1 > iy || 2 * oy + 1 < iy) continue;
      const int ny = min(2 * oy + 1, H - 1) - max(2 * oy - 1, 0) + 1;
      for (int ox = ix / 2; ox <= (ix + 1) / 2; ++ox) {
        if (ox < 0 || ox >= Wo || 2 * ox - 1 > ix || 2 * ox + 1 < ix) continue;
        const int nx = min(2 * ox + 1, W - 1) - max(2 * ox - 1, 0) + 1;
        s += __ldg(dy + (((size_t)b * Ho + oy) * Wo + ox) * C + c) / (float)(ny * nx);
      }
    }
    dx[i] = s;
  }
}

// ------------------------------------------------------------------------------------------------
// BoTNet attention backward.  One CTA per (sample, head); K+emb, V and the column accumulators dV, dKp in shared
// memory; one warp per query row recomputes the softmax row.  dqkv [B][L][3C]; demb_h / demb_w accumulated.
// ------------------------------------------------------------------------------------------------
struct AttnBwdParams {
  const float* qkv; const float* emb_h; const float* emb_w; const float* dout;
  float* dqkv; float* demb_h; float* demb_w;
  int B, Hh, Ww, heads, d; float scale;
};

// NPASS: when the column accumulators dKp / dV do not fit in shared memory next to K and V (128 tokens x 128 channels), the head
// dimension is covered in NPASS slices: every pass recomputes the softmax rows (cheap) and accumulates d / NPASS channels.
template <int KPL, int NPASS>
__global__ void __launch_bounds__(256) attention_bwd_kernel(const AttnBwdParams p) {
  extern __shared__ float sm[];
  const int L = p.Hh * p.Ww, d = p.d, C = p.heads * d;
  const int dslice = d / NPASS;
  const int b = blockIdx.x / p.heads, h = blockIdx.x % p.heads;
  float* Kp = sm;                             // [L][d+1]
  float* V = Kp + (size_t)L * (d + 1);        // [L][d+1]
  float* dKp = V + (size_t)L * (d + 1);       // [L][dslice]
  float* dV = dKp + (size_t)L * dslice;       // [L][dslice]
  float* rows = dV + (size_t)L * dslice;      // [8 warps][2][d]: scaled q row, dO row
  const float* base = p.qkv + (size_t)b * L * 3 * C;
  for (int i = threadIdx.x; i < L * d; i += 256) {
    const int j = i / d, dd = i - j * d;
    const int y = j / p.Ww, x = j - y * p.Ww;
    const float* tok = base + (size_t)j * 3 * C + h * d + dd;
    Kp[j * (d + 1) + dd] = __ldg(tok + C) + __ldg(p.emb_h + y * d + dd) + __ldg(p.emb_w + x * d + dd);
    V[j * (d + 1) + dd] = __ldg(tok + 2 * C);
  }
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  float* q = rows + warp * 2 * d;
  float* go = q + d;
  const int dpl = d / 32;                     // channels per lane
  const int tpp = dpl / NPASS;                // ... handled per pass (host guarantees dpl % NPASS == 0)
#pragma unroll 1
  for (int pass = 0; pass < NPASS; ++pass) {
    __syncthreads();                          // K / V staged (pass 0); previous pass written out
    for (int i = threadIdx.x; i < L * dslice; i += 256) { dKp[i] = 0.f; dV[i] = 0.f; }
    __syncthreads();
    for (int i = warp; i < L; i += 8) {
      for (int dd = lane; dd < d; dd += 32) {
        q[dd] = __ldg(base + (size_t)i * 3 * C + h * d + dd) * p.scale;
        go[dd] = __ldg(p.dout + ((size_t)b * L + i) * C + h * d + dd);
      }
      __syncwarp();
      float sc[KPL], dp[KPL];
      float mx = -INFINITY;
#pragma unroll
      for (int u = 0; u < KPL; ++u) {
        const int j = lane + 32 * u;
        float a = -INFINITY, t = 0.f;
        if (j < L) {
          a = 0.f;
          const float* kr = Kp + j * (d + 1);
          const float* vr = V + j * (d + 1);
          for (int dd = 0; dd < d; ++dd) { a = fmaf(q[dd], kr[dd], a); t = fmaf(go[dd], vr[dd], t); }
        }
        sc[u] = a; dp[u] = t;
        mx = fmaxf(mx, a);
      }
#pragma unroll
      for (int o = 16; o; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
      float sum = 0.f;
#pragma unroll
      for (int u = 0; u < KPL; ++u) { sc[u] = (lane + 32 * u < L) ? __expf(sc[u] - mx) : 0.f; sum += sc[u]; }
#pragma unroll
      for (int o = 16; o; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
      const float inv = 1.f / sum;
      float dot = 0.f;
#pragma unroll
      for (int u = 0; u < KPL; ++u) { sc[u] *= inv; dot += sc[u] * dp[u]; }
#pragma unroll
      for (int o = 16; o; o >>= 1) dot += __shfl_xor_sync(0xffffffffu, dot, o);
      float ds[KPL];
#pragma unroll
      for (int u = 0; u < KPL; ++u) ds[u] = sc[u] * (dp[u] - dot);
      // dq_i = scale * sum_j ds_ij Kp_j (pass 0, all channels) ; dV_j += p_ij dO_i ; dKp_j += ds_ij q_i (this pass's channel slice)
      float dq[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
      for (int u = 0; u < KPL; ++u) {
        for (int l = 0; l < 32; ++l) {
          const int j = l + 32 * u;
          if (j >= L) break;
          const float pj = __shfl_sync(0xffffffffu, sc[u], l);
          const float dj = __shfl_sync(0xffffffffu, ds[u], l);
          if (pass == 0) {
            for (int t = 0; t < dpl; ++t) dq[t] = fmaf(dj, Kp[j * (d + 1) + lane + 32 * t], dq[t]);
          }
          for (int t = 0; t < tpp; ++t) {
            const int dd = lane + 32 * (pass * tpp + t);       // channel of the head
            const int ds_i = lane + 32 * t;                    // ... within this pass's slice
            atomicAdd(&dV[j * dslice + ds_i], pj * go[dd]);
            atomicAdd(&dKp[j * dslice + ds_i], dj * q[dd]);
          }
        }
      }
      if (pass == 0) {
        for (int t = 0; t < dpl; ++t) p.dqkv[((size_t)b * L + i) * 3 * C + h * d + lane + 32 * t] = dq[t] * p.scale;
      }
      __syncwarp();
    }
    __syncthreads();
    for (int i = threadIdx.x; i < L * dslice; i += 256) {
      const int j = i / dslice, si = i - j * dslice;
      const int dd = (si & 31) + 32 * (pass * tpp + (si >> 5));
      float* tok = p.dqkv + ((size_t)b * L + j) * 3 * C + h * d + dd;
      tok[C] = dKp[i];
      tok[2 * C] = dV[i];
      const int y = j / p.Ww, x = j - y * p.Ww;
      if (p.demb_h) atomicAdd(p.demb_h + y * d + dd, dKp[i]);
      if (p.demb_w) atomicAdd(p.demb_w + x * d + dd, dKp[i]);
    }
  }
}

// ------------------------------------------------------------------------------------------------
// Losses.  Forward kernels add  coef * sum(...)  into a double slot; backward kernels write (or accumulate)
// coef * (*gscale) * d/dx.  `gscale` is the upstream gradient of the 0-dim loss tensor (device scalar, nullable = 1).
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ double block_sum_256(double v) {
  __shared__ double s_part[8];
#pragma unroll
  for (int o = 16; o; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  if ((threadIdx.x & 31) == 0) s_part[threadIdx.x >> 5] = v;
  __syncthreads();
  double t = 0;
  if (threadIdx.x == 0) for (int w = 0; w < 8; ++w) t += s_part[w];
  return t;
}

__global__ void __launch_bounds__(256) mse_const_fwd_kernel(const float* __restrict__ x, size_t n, float target, double coef, double* slot) {
  float s = 0.f;
  for (size_t i = (size_t)blockIdx.x * 256 + threadIdx.x; i < n; i += (size_t)gridDim.x * 256) { const float e = x[i] - target; s = fmaf(e, e, s); }
  const double t = block_sum_256((double)s);
  if (threadIdx.x == 0) atomicAdd(slot, coef * t);
}
__global__ void mse_const_bwd_kernel(const float* __restrict__ x, size_t n, float target, float coef, const float* __restrict__ gscale,
                                     float* __restrict__ g, int accumulate) {
  const float k = 2.f * coef * (gscale ? __ldg(gscale) : 1.f);
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
    const float v = k * (x[i] - target);
    g[i] = accumulate ? g[i] + v : v;
  }
}
// nn.BCELoss against a constant target (GANLoss with --no_lsgan, networks.py:107-108): mean of -(t log p + (1 - t) log(1 - p)), the logs
// clamped at -100 and the gradient (p - t) / max(p (1 - p), 1e-12) exactly like torch (binary_cross_entropy / _backward)
__global__ void __launch_bounds__(256) bce_const_fwd_kernel(const float* __restrict__ x, size_t n, float target, double coef, double* slot) {
  float s = 0.f;
  for (size_t i = (size_t)blockIdx.x * 256 + threadIdx.x; i < n; i += (size_t)gridDim.x * 256) {
    const float p = x[i];
    s -= target * fmaxf(logf(p), -100.f) + (1.f - target) * fmaxf(log1pf(-p), -100.f);
  }
  const double t = block_sum_256((double)s);
  if (threadIdx.x == 0) atomicAdd(slot, coef * t);
}
__global__ void bce_const_bwd_kernel(const float* __restrict__ x, size_t n, float target, float coef, const float* __restrict__ gscale,
                                     float* __restrict__ g, int accumulate) {
  const float k = coef * (gscale ? __ldg(gscale) : 1.f);
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
    const float p = x[i];
    const float v = k * (p - target) / fmaxf((1.f - p) * p, 1e-12f);
    g[i] = accumulate ? g[i] + v : v;
  }
}
__global__ void __launch_bounds__(256) l1_pair_fwd_kernel(const float* __restrict__ a, const float* __restrict__ b, size_t n, double coef, double* slot) {
  float s = 0.f;
  for (size_t i = (size_t)blockIdx.x * 256 + threadIdx.x; i < n; i += (size_t)gridDim.x * 256) s += fabsf(a[i] - b[i]);
  const double t = block_sum_256((double)s);
  if (threadIdx.x == 0) atomicAdd(slot, coef * t);
}
__global__ void l1_pair_bwd_kernel(const float* __restrict__ a, const float* __restrict__ b, size_t n, float coef, const float* __restrict__ gscale,
                                   float* __restrict__ g, int accumulate) {
  const float k = coef * (gscale ? __ldg(gscale) : 1.f);
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
    const float e = a[i] - b[i];
    const float v = e > 0.f ? k : (e < 0.f ? -k : 0.f);
    g[i] = accumulate ? g[i] + v : v;
  }
}
// All loss reductions of a step in ONE launch, and all their gradient seeds in another (cfg4: 9 GAN terms + 12 feature-matching terms
// were 21 + 21 launches of a few microseconds each on the dependent chain between the discriminator forward and the sweeps).
// The table travels by value in the kernel parameters (captured by value in a CUDA graph: no host table to keep alive).
constexpr int kLossItems = 24;
struct LossItem {
  const float* a; const float* b;   // kind 0 / 2: a = predictions (b unused); kind 1: L1 pair (a, b)
  float* g;                         // backward: gradient w.r.t. a (written, not accumulated)
  long long n;
  double coef;                      // forward: loss += coef * sum; backward: d/da scaled by coef * (*gscale)
  const float* gscale;              // backward: upstream gradient of the 0-dim loss tensor (device scalar, nullable = 1)
  float target;                     // kind 0 / 2
  int kind;                         // 0 MSE against a constant, 1 L1 between two tensors, 2 BCE against a constant (on probabilities)
  int slot;                         // forward: index into the fp64 accumulator
};
struct LossTable { LossItem it[kLossItems]; int n; };

__global__ void __launch_bounds__(256) multi_loss_fwd_kernel(const LossTable tab, double* __restrict__ acc) {
  const LossItem& L = tab.it[blockIdx.y];
  float s = 0.f;
  for (size_t i = (size_t)blockIdx.x * 256 + threadIdx.x; i < (size_t)L.n; i += (size_t)gridDim.x * 256) {
    if (L.kind == 0) { const float e = L.a[i] - L.target; s = fmaf(e, e, s); }
    else if (L.kind == 1) s += fabsf(L.a[i] - L.b[i]);
    else { const float p = L.a[i]; s -= L.target * fmaxf(logf(p), -100.f) + (1.f - L.target) * fmaxf(log1pf(-p), -100.f); }
  }
  const double t = block_sum_256((double)s);
  if (threadIdx.x == 0 && (size_t)blockIdx.x * 256 < (size_t)L.n) atomicAdd(acc + L.slot, L.coef * t);
}
__global__ void __launch_bounds__(256) multi_loss_bwd_kernel(const LossTable tab) {
  const LossItem& L = tab.it[blockIdx.y];
  const float k = (float)L.coef * (L.gscale ? __ldg(L.gscale) : 1.f);
  for (size_t i = (size_t)blockIdx.x * 256 + threadIdx.x; i < (size_t)L.n; i += (size_t)gridDim.x * 256) {
    float v;
    if (L.kind == 0) v = 2.f * k * (L.a[i] - L.target);
    else if (L.kind == 1) { const float e = L.a[i] - L.b[i]; v = e > 0.f ? k : (e < 0.f ? -k : 0.f); }
    else { const float p = L.a[i]; v = k * (p - L.target) / fmaxf((1.f - p) * p, 1e-12f); }
    L.g[i] = v;
  }
}
__global__ void f64_to_f32_kernel(const double* __restrict__ a, float* __restrict__ y, int n) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) y[i] = (float)a[i];
}

// cat(lr, s, |s|*2 + lo) as NHWC [rows][nbins][3]   (lr rows may sit in a wider tensor: row stride in elements)
__global__ void disc_input_fwd_kernel(const float* __restrict__ lr, int64_t lr_clip_stride, const float* __restrict__ s, float* __restrict__ out,
                                      int64_t clips, int64_t per_clip, float lo) {
  const size_t total = (size_t)clips * per_clip;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
    const size_t b = i / per_clip, r = i - b * per_clip;
    const float v = s[i];
    out[3 * i + 0] = lr[b * lr_clip_stride + r];
    out[3 * i + 1] = v;
    out[3 * i + 2] = fabsf(v) * 2.f + lo;
  }
}
// ds = g[.,1] + 2 sign(s) g[.,2]
__global__ void disc_input_bwd_kernel(const float* __restrict__ g, const float* __restrict__ s, float* __restrict__ ds, size_t n) {
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
    const float v = s[i];
    const float sg = v > 0.f ? 1.f : (v < 0.f ? -1.f : 0.f);
    ds[i] = g[3 * i + 1] + 2.f * sg * g[3 * i + 2];
  }
}

// ------------------------------------------------------------------------------------------------
// Kernel-side weight images of EVERY convolution of a model in ONE launch (after each optimiser step): from the
// parameter's own layout to  [K][N] (direct kernels)  and / or the tcgen05 kernel's pre-swizzled TF32 hi|lo image
// (conv_umma.cuh), for the forward GEMM and for the input-gradient GEMM (transposed geometry / flipped taps).
//   element (k, n):  tap = k / Kch, kc = k % Kch;  src[kc*s_kch + n*s_n + (flip ? taps-1-tap : tap)*s_tap]
// (s_tap = 1 for the reference's parameter layout; Cin*Cout for the [kh][kw][Cin][Cout] storage order of optim.FlatBucket)
// ------------------------------------------------------------------------------------------------
struct PackDesc {
  const float* src; float* dst_kn; float* dst_umma;
  int K, N, Kch, taps, flip, kchunks;
  long long s_kch, s_n, s_tap;
  long long work_begin;          // prefix sum of kchunks*32*N over the descriptors
};
constexpr unsigned kTf32MaskPack = 0xFFFFE000u;

__global__ void __launch_bounds__(256) pack_weights_multi_kernel(const PackDesc* __restrict__ descs, int n_desc, long long total) {
  constexpr int kPer = 8;                       // a block walks 256*8 consecutive elements: the descriptor lookup is reused
  PackDesc d;
  long long d_end = -1;                         // [d.work_begin, d_end) = range of the cached descriptor
  d.work_begin = 0;
  for (long long c0 = (long long)blockIdx.x * 256 * kPer; c0 < total; c0 += (long long)gridDim.x * 256 * kPer) {
#pragma unroll 1
    for (int jj = 0; jj < kPer; ++jj) {
      const long long i = c0 + threadIdx.x + 256 * jj;
      if (i >= total) break;
      if (i >= d_end || i < d.work_begin) {
        int lo = 0, hi = n_desc - 1;
        while (lo < hi) {                       // last descriptor with work_begin <= i
          const int mid = (lo + hi + 1) >> 1;
          if (descs[mid].work_begin <= i) lo = mid; else hi = mid - 1;
        }
        d = descs[lo];
        d_end = d.work_begin + (long long)d.kchunks * 32 * d.N;
      }
      const long long r = i - d.work_begin;
      int n, k;
      if (d.dst_umma && !d.dst_kn) {            // tcgen05 image only: a warp writes one 128-byte row (32 consecutive k of one n)
        const int kk = (int)(r & 31);
        const long long q = r >> 5;
        n = (int)(q % d.N);
        k = (int)(q / d.N) * 32 + kk;
      } else {                                  // [K][N] image: n fastest
        n = (int)(r % d.N);
        k = (int)(r / d.N);
      }
      float v = 0.f;
      if (k < d.K) {
        const int tap = k / d.Kch, kc = k - tap * d.Kch;
        v = __ldg(d.src + kc * d.s_kch + n * d.s_n + (long long)(d.flip ? d.taps - 1 - tap : tap) * d.s_tap);
        if (d.dst_kn) d.dst_kn[(size_t)k * d.N + n] = v;
      }
      if (d.dst_umma) {
        const float hi_v = __uint_as_float(__float_as_uint(v) & kTf32MaskPack);
        const float lo_v = __uint_as_float((__float_as_uint(v - hi_v) + 0x1000u) & kTf32MaskPack);   // nearest TF32 of the remainder
        const int kcn = k >> 5, kk = k & 31;
        const int piece = (kk >> 2) ^ (n & 7);
        const int np = (d.N + 31) / 32 * 32;     // image rows per (chunk, part): N padded to 32 (padding rows stay zero)
        const size_t dst = (((size_t)kcn * 2) * np + n) * 32 + piece * 4 + (kk & 3);
        d.dst_umma[dst] = hi_v;
        d.dst_umma[dst + (size_t)np * 32] = lo_v;
      }
    }
  }
}

// ------------------------------------------------------------------------------------------------
// Long-form generation (generate_audio.py): segmenting a whole clip and overlap-adding the generated segments.
//   segment_gather: AudioTestDataset.seg_pad_audio (data/audio_dataset.py:153-167): zero-pad (ov, seg*ceil(L/seg) - L + ov),
//                   unfold(size = seg, step = seg - ov)
//   segment_ola:    generate_audio.py:40-53: halve the first / last ov samples of every segment, fold with stride seg - ov,
//                   crop ov at both ends (ov = 0: plain concatenation)
// ------------------------------------------------------------------------------------------------
__global__ void segment_gather_kernel(const float* __restrict__ audio, long long L, float* __restrict__ out, long long n_seg, int seg, int step,
                                      int ov) {
  const long long total = n_seg * seg;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const long long s = i / seg;
    const int j = (int)(i - s * seg);
    const long long t = s * step + j - ov;          // position in the un-padded clip
    out[i] = (t >= 0 && t < L) ? audio[t] : 0.f;
  }
}

template <typename T>
__global__ void segment_ola_kernel(const T* __restrict__ x, T* __restrict__ out, long long out_len, long long n_seg, int seg, int step, int ov,
                                   int crop_begin) {
  for (long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x; t < out_len; t += (long long)gridDim.x * blockDim.x) {
    const long long q = t + crop_begin;             // position in the folded signal (crop_begin = ov for a whole clip)
    long long s_hi = q / step;
    if (s_hi >= n_seg) s_hi = n_seg - 1;
    T acc = (T)0;
    for (long long s = s_hi; s >= 0; --s) {         // segments covering q, in fold order (ascending s summed last-to-first is the same sum of <= 2 terms)
      const long long j = q - s * step;
      if (j >= seg) break;
      T v = x[s * seg + j];
      if (ov > 0 && (j < ov || j >= seg - ov)) v *= (T)0.5;
      acc += v;
    }
    out[t] = acc;
  }
}

// ------------------------------------------------------------------------------------------------
// Evaluation metrics (util/util.py:132-177 compute_matrics): MSE, SNR of sr and lr against hr, and the log-spectral distance
// over a 2*n_fft STFT with the kbdwin(2*win) window (torchaudio.functional.spectrogram, power 2, centre = reflect padding).
//   metrics_rows_kernel: per row (last dim) sums  sum (sr-hr)^2, sum hr^2, sum (lr-hr)^2      -> rows[r][3] (double)
//   lsd_frames_kernel:   one CTA per (row, frame): hr and sr frames go through ONE complex FFT (z = hr + i sr, the two real
//                        spectra are separated afterwards), |X|^2 -> log10(. + 1e-6) -> sqrt(mean_k diff^2) -> acc += (double)
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) metrics_rows_kernel(const float* __restrict__ hr, const float* __restrict__ lr, const float* __restrict__ sr,
                                                           long long T, double* __restrict__ rows) {
  const long long r = blockIdx.y;
  const float* h = hr + r * T; const float* l = lr + r * T; const float* s = sr + r * T;
  double a = 0, b = 0, c = 0;
  for (long long i = (long long)blockIdx.x * 256 + threadIdx.x; i < T; i += (long long)gridDim.x * 256) {
    const double hv = h[i], d1 = (double)s[i] - hv, d2 = (double)l[i] - hv;
    a = fma(d1, d1, a); b = fma(hv, hv, b); c = fma(d2, d2, c);
  }
  const double ta = block_sum_256(a);
  __syncthreads();
  const double tb = block_sum_256(b);
  __syncthreads();
  const double tc = block_sum_256(c);
  if (threadIdx.x == 0) { atomicAdd(rows + 3 * r, ta); atomicAdd(rows + 3 * r + 1, tb); atomicAdd(rows + 3 * r + 2, tc); }
}

template <int NF>      // FFT size (2 * n_fft of the model: 1024)
__global__ void __launch_bounds__(256) lsd_frames_kernel(const float* __restrict__ hr, const float* __restrict__ sr, long long T, int hop, int frames,
                                                         const float* __restrict__ window, int center, double* __restrict__ acc) {
  __shared__ float2 z[NF];
  __shared__ float2 tw[NF / 2];
  const int row = blockIdx.y, fr = blockIdx.x;
  const float* h = hr + (long long)row * T; const float* s = sr + (long long)row * T;
  const long long start = (long long)fr * hop - (center ? NF / 2 : 0);
  constexpr int LOG = NF == 2048 ? 11 : (NF == 1024 ? 10 : (NF == 512 ? 9 : 8));
  for (int i = threadIdx.x; i < NF / 2; i += 256) {
    float sn, cs;
    sincospif(-2.f * (float)i / (float)NF, &sn, &cs);
    tw[i] = make_float2(cs, sn);
  }
  for (int i = threadIdx.x; i < NF; i += 256) {       // bit-reversed load of the windowed, reflect-padded frames
    long long t = start + i;
    if (t < 0) t = -t;
    if (t >= T) t = 2 * (T - 1) - t;
    const float w = __ldg(window + i);
    const unsigned j = __brev((unsigned)i) >> (32 - LOG);
    z[j] = make_float2(w * __ldg(h + t), w * __ldg(s + t));
  }
  __syncthreads();
#pragma unroll 1
  for (int st = 0; st < LOG; ++st) {
    const int half = 1 << st;
    for (int bfly = threadIdx.x; bfly < NF / 2; bfly += 256) {
      const int grp = bfly >> st, pos = bfly & (half - 1);
      const int i0 = (grp << (st + 1)) + pos, i1 = i0 + half;
      const float2 w = tw[pos << (LOG - 1 - st)];
      const float2 a = z[i0], b = z[i1];
      const float2 t = make_float2(b.x * w.x - b.y * w.y, b.x * w.y + b.y * w.x);
      z[i0] = make_float2(a.x + t.x, a.y + t.y);
      z[i1] = make_float2(a.x - t.x, a.y - t.y);
    }
    __syncthreads();
  }
  float part = 0.f;
  for (int k = threadIdx.x; k <= NF / 2; k += 256) {
    const float2 a = z[k], b = z[(NF - k) & (NF - 1)];
    // X_hr = (Z[k] + conj(Z[N-k])) / 2,  X_sr = (Z[k] - conj(Z[N-k])) / (2i)
    const float hre = 0.5f * (a.x + b.x), him = 0.5f * (a.y - b.y);
    const float sre = 0.5f * (a.y + b.y), sim = -0.5f * (a.x - b.x);
    const float ph = hre * hre + him * him, ps = sre * sre + sim * sim;
    const float d = log10f(ph + 1e-6f) - log10f(ps + 1e-6f);
    part = fmaf(d, d, part);
  }
  const double tot = block_sum_256((double)part);
  if (threadIdx.x == 0) atomicAdd(acc, sqrt(tot / (double)(NF / 2 + 1)));
}

// Tiled form for the tensor-core images of layers with Kch % 32 == 0 and N % 32 == 0 (almost all parameters): a CTA takes a
// 32 (n) x 32 (channel) x taps block of one descriptor, reads it in contiguous runs of the parameter layout, transposes through
// shared memory and writes whole 128-byte rows of the image (hi and lo) -- every global access coalesced.
// tile_begin = prefix sum of (N/32)*(Kch/32) over the descriptors.
constexpr int kPackMaxTaps = 49;
constexpr int kPackTapStride = 32 * 33 + 1;   // [tap][kc][33] with the tap stride odd: the 9 taps of one (n, kc) land in different banks
__global__ void __launch_bounds__(256) pack_weights_tiled_kernel(const PackDesc* __restrict__ descs, const long long* __restrict__ tile_begin,
                                                                 int n_desc) {
  extern __shared__ float sm_pack[];            // [taps][32 (kc)][33] (+1 per tap)
  const long long tile = blockIdx.x;
  int lo = 0, hi = n_desc - 1;
  while (lo < hi) {
    const int mid = (lo + hi + 1) >> 1;
    if (tile_begin[mid] <= tile) lo = mid; else hi = mid - 1;
  }
  const PackDesc d = descs[lo];
  const int t = (int)(tile - tile_begin[lo]);
  const int kc_blocks = d.Kch / 32;
  const int nb = t / kc_blocks, kb = t - nb * kc_blocks;
  const int n0 = nb * 32, kc0 = kb * 32;
  const int taps = d.taps;
  const int total = 32 * 32 * taps;
  const bool n_outer = d.s_n > d.s_kch;         // which of (n, kc) has the larger stride in the parameter layout
  if (d.s_tap == 1) {                           // reference parameter layout: the taps of one (n, kc) are contiguous
    for (int idx = threadIdx.x; idx < total; idx += 256) {
      const int tap = idx % taps;
      const int ab = idx / taps;
      const int inner = ab & 31, outer = ab >> 5;
      const int n = n_outer ? outer : inner, kc = n_outer ? inner : outer;
      const float v = __ldg(d.src + (long long)(kc0 + kc) * d.s_kch + (long long)(n0 + n) * d.s_n + tap);
      sm_pack[tap * kPackTapStride + kc * 33 + n] = v;
    }
  } else {                                      // [tap][ci][co] storage: 32 consecutive values of the unit-stride index per warp
    for (int idx = threadIdx.x; idx < total; idx += 256) {
      const int inner = idx & 31, outer = (idx >> 5) & 31, tap = idx >> 10;
      const int n = n_outer ? outer : inner, kc = n_outer ? inner : outer;
      const float v = __ldg(d.src + (long long)(kc0 + kc) * d.s_kch + (long long)(n0 + n) * d.s_n + (long long)tap * d.s_tap);
      sm_pack[tap * kPackTapStride + kc * 33 + n] = v;
    }
  }
  __syncthreads();
  // rows of the image: (tap_dst, n) -> 32 consecutive k = tap_dst*Kch + kc0 .. +31; a warp writes one row (lane = kc)
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int r = warp; r < taps * 32; r += 8) {
    const int tap_dst = r >> 5, n = r & 31;
    const int tap_src = d.flip ? taps - 1 - tap_dst : tap_dst;
    const float v = sm_pack[tap_src * kPackTapStride + lane * 33 + n];
    const float hi_v = __uint_as_float(__float_as_uint(v) & kTf32MaskPack);
    const float lo_v = __uint_as_float((__float_as_uint(v - hi_v) + 0x1000u) & kTf32MaskPack);
    const int k = tap_dst * d.Kch + kc0 + lane;
    const int kcn = k >> 5, kk = k & 31, ng = n0 + n;
    const int piece = (kk >> 2) ^ (ng & 7);
    const size_t dst = (((size_t)kcn * 2) * d.N + ng) * 32 + piece * 4 + (kk & 3);
    d.dst_umma[dst] = hi_v;
    d.dst_umma[dst + (size_t)d.N * 32] = lo_v;
  }
}

// ------------------------------------------------------------------------------------------------
// Audio2MDCT.normalize / denormalize as stand-alone calls (pix2pixHD_model.py:83-137, arcsinh / raw branch with abs_norm): the
// fused transform kernels do this in their epilogue / prologue; these serve callers that hold a spectrogram already.
// Computed in fp64 like the reference (its spectrograms are fp64); ln10 is the fp32 constant of :100,133.
// ------------------------------------------------------------------------------------------------
struct SpecNorm { int mode; double gain, src_lo, src_hi, norm_lo, norm_hi; };
template <typename TI, typename TO>
__global__ void spectro_normalize_kernel(const TI* __restrict__ x, TO* __restrict__ y, size_t n, SpecNorm p) {
  const double ln10 = 2.3025851249694824;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
    double s = (double)x[i];
    if (p.mode == 1) s = asinh(p.gain * s) / ln10;
    s = (s - p.src_lo) / (p.src_hi - p.src_lo);
    y[i] = (TO)(s * (p.norm_hi - p.norm_lo) + p.norm_lo);
  }
}
template <typename TI>
__global__ void spectro_denormalize_kernel(const TI* __restrict__ x, double* __restrict__ y, size_t n, SpecNorm p) {
  const double ln10 = 2.3025851249694824;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
    double s = ((double)x[i] - p.norm_lo) / (p.norm_hi - p.norm_lo);
    s = s * (p.src_hi - p.src_lo) + p.src_lo;
    y[i] = p.mode == 1 ? sinh(s * ln10) / p.gain : s;
  }
}

// ------------------------------------------------------------------------------------------------
// torchaudio.functional.resample (sinc_interp_hann polyphase FIR), the data-preparation step in front of the hot path
// (data/audio_dataset.py:66-71: HR -> LR -> HR on CPU workers in the reference):
//   y[r][q*new + ph] = sum_k x_pad[r][q*orig + k] * table[ph][k],  x_pad = x zero-padded by `width` on the left, width + orig right
// One thread per output sample; the [new][K] table sits in shared memory when it fits.
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) resample_fir_kernel(const float* __restrict__ x, long long L, const float* __restrict__ table, int K,
                                                           int orig, int nw, int width, float* __restrict__ y, long long target, int rows,
                                                           int table_in_smem) {
  extern __shared__ float s_tab[];
  if (table_in_smem) {
    for (int i = threadIdx.x; i < nw * K; i += 256) s_tab[i] = __ldg(table + i);
    __syncthreads();
  }
  const float* tab = table_in_smem ? s_tab : table;
  const long long total = (long long)rows * target;
  for (long long i = (long long)blockIdx.x * 256 + threadIdx.x; i < total; i += (long long)gridDim.x * 256) {
    const long long r = i / target, j = i - r * target;
    const long long q = j / nw;
    const int ph = (int)(j - q * nw);
    const float* xr = x + r * L;
    const long long base = q * orig - width;
    const float* t = tab + (size_t)ph * K;
    float acc = 0.f;
    for (int k = 0; k < K; ++k) {
      const long long xi = base + k;
      if (xi >= 0 && xi < L) acc = fmaf(__ldg(xr + xi), t[k], acc);
    }
    y[i] = acc;
  }
}

// torch.optim.Adam (no weight decay, no amsgrad), fp32, one flat buffer.  g is pre-scaled by grad_scale (1/world).
// ------------------------------------------------------------------------------------------------
// AudioDataset.__getitem__ noise injection (data/audio_dataset.py:72-78):
//   noise -= mean(noise); noise *= sqrt(sum(lr^2) / segment_length / 10^(snr/10)) / std(noise) (unbiased); lr += noise
// add_noise_sums_kernel: (sum n, sum n^2, sum lr^2) in fp64 -> acc[3]; add_noise_apply_kernel: the affine on every sample.
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) add_noise_sums_kernel(const float* __restrict__ lr, const float* __restrict__ noise, long long n, double* acc) {
  double s0 = 0.0, s1 = 0.0, s2 = 0.0;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const double v = noise[i], x = lr[i];
    s0 += v; s1 = fma(v, v, s1); s2 = fma(x, x, s2);
  }
  __shared__ double sh[3][256];
  sh[0][threadIdx.x] = s0; sh[1][threadIdx.x] = s1; sh[2][threadIdx.x] = s2;
  __syncthreads();
  for (int k = 128; k > 0; k >>= 1) {
    if (threadIdx.x < k) { sh[0][threadIdx.x] += sh[0][threadIdx.x + k]; sh[1][threadIdx.x] += sh[1][threadIdx.x + k]; sh[2][threadIdx.x] += sh[2][threadIdx.x + k]; }
    __syncthreads();
  }
  if (threadIdx.x < 3) atomicAdd(acc + threadIdx.x, sh[threadIdx.x][0]);
}
__global__ void add_noise_apply_kernel(const float* __restrict__ lr, const float* __restrict__ noise, float* __restrict__ out, long long n,
                                       const double* __restrict__ acc, double segment_length, double snr_db) {
  const double mean = acc[0] / (double)n;
  const double var = (acc[1] - (double)n * mean * mean) / (double)(n - 1);                // torch.std: unbiased
  const double noise_var = acc[2] / segment_length / pow(10.0, snr_db / 10.0);
  // the reference works in fp32: noise - mean, sqrt(noise_var) / std * noise, lr + noise (data/audio_dataset.py:73-78)
  const float meanf = (float)mean, scale = sqrtf((float)noise_var) / (float)sqrt(var);
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
    out[i] = lr[i] + scale * (noise[i] - meanf);
}

struct AdamParams {
  float* p; const float* g; float* m; float* v; size_t n;
  float lr, beta1, beta2, eps, grad_scale, bias_c1, bias_c2_sqrt;
  const long long* step_dev;   // optional: 1-based step counter in device memory (CUDA-graph replays), overrides bias_c*
};
__global__ void counter_inc_kernel(long long* c) { *c += 1; }
__global__ void __launch_bounds__(256) adam_flat_kernel(AdamParams a) {
  if (a.step_dev) {
    const double t = (double)*a.step_dev;
    a.bias_c1 = (float)(1.0 - pow((double)a.beta1, t));
    a.bias_c2_sqrt = (float)sqrt(1.0 - pow((double)a.beta2, t));
  }
  const float step_size = a.lr / a.bias_c1;
  const size_t n4 = a.n / 4;
  for (size_t i = (size_t)blockIdx.x * 256 + threadIdx.x; i < n4; i += (size_t)gridDim.x * 256) {
    float4 p = reinterpret_cast<float4*>(a.p)[i];
    const float4 g = reinterpret_cast<const float4*>(a.g)[i];
    float4 m = reinterpret_cast<float4*>(a.m)[i];
    float4 v = reinterpret_cast<float4*>(a.v)[i];
    float pp[4] = {p.x, p.y, p.z, p.w}, gg[4] = {g.x, g.y, g.z, g.w}, mm[4] = {m.x, m.y, m.z, m.w}, vv[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const float gr = gg[u] * a.grad_scale;
      mm[u] = mm[u] + (gr - mm[u]) * (1.f - a.beta1);               // exp_avg.lerp_(grad, 1 - beta1)
      vv[u] = vv[u] * a.beta2 + (1.f - a.beta2) * gr * gr;          // exp_avg_sq.mul_(beta2).addcmul_(grad, grad, 1 - beta2)
      const float denom = sqrtf(vv[u]) / a.bias_c2_sqrt + a.eps;
      pp[u] = pp[u] - step_size * (mm[u] / denom);
    }
    reinterpret_cast<float4*>(a.p)[i] = make_float4(pp[0], pp[1], pp[2], pp[3]);
    reinterpret_cast<float4*>(a.m)[i] = make_float4(mm[0], mm[1], mm[2], mm[3]);
    reinterpret_cast<float4*>(a.v)[i] = make_float4(vv[0], vv[1], vv[2], vv[3]);
  }
  if (blockIdx.x == 0 && threadIdx.x == 0) {
    for (size_t i = n4 * 4; i < a.n; ++i) {
      const float gr = a.g[i] * a.grad_scale;
      const float m = a.m[i] + (gr - a.m[i]) * (1.f - a.beta1);
      const float v = a.v[i] * a.beta2 + (1.f - a.beta2) * gr * gr;
      a.m[i] = m; a.v[i] = v;
      a.p[i] = a.p[i] - step_size * (m / (sqrtf(v) / a.bias_c2_sqrt + a.eps));
    }
  }
}

}  // namespace trk
