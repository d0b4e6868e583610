// mdct_kernels.cuh -- fused MDCT4 / IMDCT4 kernels for sm_100a (n_fft 512, hop 256, 256 bins).
//
// Replaces on the device (reference file:line):
//   forward : models/mdct.py:392-425 (pad, unfold, window, pre-twiddle, 512-pt Z2Z FFT, post-twiddle)
//             + models/pix2pixHD_model.py:96-100,115-123 (arcsinh compress, abs-norm affine)
//             + :400-402 (second channel |s|*2+lo)                       -> ONE kernel, one HBM pass
//   inverse : models/pix2pixHD_model.py:127-133 (denormalise, sinh expand)
//             + models/mdct.py:457-489 (pre-twiddle, FFT, post-twiddle, window, fold/overlap-add, crop)
//
// Kernel shape (both directions): persistent CTAs of 5 warps.  Warps 0-3 compute (8 threads per
// frame, 4 frames per warp, 16 frames per tile); warp 4 is the loader: while the compute warps
// work on tile i from one shared-memory buffer it stages tile i+1 into the other one with 16-byte
// coalesced global loads, de-interleaving even/odd samples so that the stride-2 TDAC gather is
// bank-conflict free.  One __syncthreads per tile.  HBM traffic is the algorithmic minimum: every
// input sample is read once (+1/16 tile overlap, served by L2) and every output written once.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <type_traits>

#include "mdct_core.cuh"

namespace mdctk {

constexpr int kComputeWarps = 4;
constexpr int kThreads = (kComputeWarps + 1) * 32;
constexpr int kTileRows = kFramesPerTile + 1;                  // 17 blocks of 256 samples feed 16 frames
constexpr int kFwdBufFloats = 2 * kTileRows * kRowPad;         // E region + O region
constexpr int kInvFramesOut = kFramesPerTile - 1;              // 15 complete output blocks per tile

struct FwdParams {
  const float* audio; int64_t audio_stride; int64_t T;
  int64_t B, F; int64_t tiles_per_clip; int64_t ntiles;
  const void* tabT; const float* tabW;
  void* out; int64_t out_clip_stride; int64_t out_chan_stride; int channels;
  NormParams np;
};

struct InvParams {
  const void* spec; int64_t spec_clip_stride;   // elements; frames are contiguous rows of 256
  int64_t B, F; int64_t tiles_per_clip; int64_t ntiles;
  const void* tabT; const float* window;
  void* out; int64_t out_clip_stride; int64_t out_len;   // samples written per clip (<= (F-1)*256)
  NormParams np; float inv_a, inv_b;                       // s_src = s*inv_a + inv_b
};

// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ float fast_asinh_scaled(float y, float c1) {
  // sign(y) * log2(|y| + sqrt(y^2+1)) * c1 ; abs error ~1e-8 * c1-scale (see DESIGN.md, K1 epilogue)
  const float ay = fabsf(y);
  float r;
  asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(fmaf(ay, ay, 1.0f)));
  float l;
  asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(l) : "f"(ay + r));
  return copysignf(l * c1, y);
}
__device__ __forceinline__ float fast_sinh(float t) {
  // (2^(t*log2e) - 2^(-t*log2e)) / 2
  const float u = t * 1.4426950408889634f;
  float p, q;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(p) : "f"(u));
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(q) : "f"(-u));
  return 0.5f * (p - q);
}

template <typename R> struct Vec2;
template <> struct Vec2<float> { using type = float2; };
template <> struct Vec2<double> { using type = double2; };

// ================================================================================================
// Forward
// ================================================================================================
__device__ __forceinline__ void fwd_stage_tile(const FwdParams& p, int64_t tile, float* buf, int lane) {
  const int64_t b = tile / p.tiles_per_clip;
  const int64_t t0 = (tile - b * p.tiles_per_clip) * kFramesPerTile;
  const float* __restrict__ src = p.audio + b * p.audio_stride;
  const bool aligned = ((reinterpret_cast<uintptr_t>(src) & 15) == 0);
  float* E = buf;
  float* O = buf + kTileRows * kRowPad;
  constexpr int kVecs = kTileRows * 64;   // float4 per tile
  constexpr int kBatch = 8;
#pragma unroll 1
  for (int base = 0; base < kVecs; base += 32 * kBatch) {
    float4 v[kBatch];
#pragma unroll
    for (int u = 0; u < kBatch; ++u) {
      const int idx = base + u * 32 + lane;
      v[u] = make_float4(0.f, 0.f, 0.f, 0.f);
      if (idx < kVecs) {
        const int bk = idx >> 6, q = idx & 63;
        const int64_t s = (t0 + bk - 1) * kHop + 4 * q;
        if (s >= 0 && s + 3 < p.T && aligned) {
          v[u] = __ldg(reinterpret_cast<const float4*>(src + s));
        } else if (s + 3 >= 0 && s < p.T) {
          if (s + 0 >= 0 && s + 0 < p.T) v[u].x = __ldg(src + s + 0);
          if (s + 1 >= 0 && s + 1 < p.T) v[u].y = __ldg(src + s + 1);
          if (s + 2 >= 0 && s + 2 < p.T) v[u].z = __ldg(src + s + 2);
          if (s + 3 >= 0 && s + 3 < p.T) v[u].w = __ldg(src + s + 3);
        }
      }
    }
#pragma unroll
    for (int u = 0; u < kBatch; ++u) {
      const int idx = base + u * 32 + lane;
      if (idx < kVecs) {
        const int bk = idx >> 6, q = idx & 63;
        *reinterpret_cast<float2*>(E + bk * kRowPad + 2 * q) = make_float2(v[u].x, v[u].z);
        *reinterpret_cast<float2*>(O + bk * kRowPad + 2 * q) = make_float2(v[u].y, v[u].w);
      }
    }
  }
}

// EPI: 0 = raw coefficients (OutT = R), 1 = fused compress + abs-norm (OutT = float, 1 or 2 channels)
template <typename R, int EPI>
__global__ void __launch_bounds__(kThreads, sizeof(R) == 4 ? 3 : 1) mdct4_fwd_kernel(const FwdParams p) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  float* tilebuf = reinterpret_cast<float*>(smem_raw);                         // [2][kFwdBufFloats]
  cx<R>* xch_all = reinterpret_cast<cx<R>*>(tilebuf + 2 * kFwdBufFloats);      // [16][kXchStride]
  using OutT = typename std::conditional<EPI == 0, R, float>::type;

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int g = lane >> 3, j = lane & 7;
  const int f = (warp & 3) * 4 + g;
  const bool loader = (warp == kComputeWarps);

  ThreadTab<R> tt;
  WinTab wt;
  if (!loader) {
    const R scale = (EPI == 1 && p.np.mode == 1) ? (R)p.np.gain : (R)1;
    load_T<R>(reinterpret_cast<const R*>(p.tabT), j, scale, tt);
    load_W(p.tabW, j, wt);
  }
  const float c1 = (float)(0.6931471805599453 / kLn10F32) * p.np.aff_a;   // log2 -> ln -> /ln10_f32 -> affine

  int64_t tile = blockIdx.x;
  int buf = 0;
  if (loader && tile < p.ntiles) fwd_stage_tile(p, tile, tilebuf, lane);
  __syncthreads();
  for (; tile < p.ntiles; tile += gridDim.x) {
    const int64_t next = tile + gridDim.x;
    if (loader) {
      if (next < p.ntiles) fwd_stage_tile(p, next, tilebuf + (buf ^ 1) * kFwdBufFloats, lane);
    } else {
      const int64_t b = tile / p.tiles_per_clip;
      const int64_t t = (tile - b * p.tiles_per_clip) * kFramesPerTile + f;
      const float* E = tilebuf + buf * kFwdBufFloats;
      const float* O = E + kTileRows * kRowPad;
      cx<R>* xch = xch_all + f * kXchStride;
      {
        cx<R> v[16];
        fwd_gather<R>(E + f * kRowPad, O + f * kRowPad, E + (f + 1) * kRowPad, O + (f + 1) * kRowPad, j, wt, v);
        pass1<R>(v, tt, j, xch);
      }
      __syncwarp();
      cx<R> y[2][8];
      pass2<R>(xch, j, y);
      __syncwarp();   // the exchange slots are rewritten by the next tile's pass 1
      if (t < p.F) {
        OutT* row = reinterpret_cast<OutT*>(p.out) + b * p.out_clip_stride + t * kBins;
#pragma unroll
        for (int h = 0; h < 2; ++h) {
#pragma unroll
          for (int k2 = 0; k2 < 8; ++k2) {
            int col; R d0, d1;
            out_pair<R>(y, j, h, k2, col, d0, d1);
            if (EPI == 0) {
              typename Vec2<OutT>::type o; o.x = (OutT)d0; o.y = (OutT)d1;
              *reinterpret_cast<typename Vec2<OutT>::type*>(row + col) = o;
            } else {
              float s0, s1;
              if (sizeof(R) == 8) {   // "exact" flavour: fp64 core and library asinh, rounded once to fp32
                if (p.np.mode == 1) {
                  s0 = (float)(asinh((double)d0) / kLn10F32 * (double)p.np.aff_a + (double)p.np.aff_b);
                  s1 = (float)(asinh((double)d1) / kLn10F32 * (double)p.np.aff_a + (double)p.np.aff_b);
                } else {
                  s0 = (float)((double)d0 * (double)p.np.aff_a + (double)p.np.aff_b);
                  s1 = (float)((double)d1 * (double)p.np.aff_a + (double)p.np.aff_b);
                }
              } else if (p.np.mode == 1) {
                s0 = fast_asinh_scaled((float)d0, c1) + p.np.aff_b;
                s1 = fast_asinh_scaled((float)d1, c1) + p.np.aff_b;
              } else {
                s0 = fmaf((float)d0, p.np.aff_a, p.np.aff_b);
                s1 = fmaf((float)d1, p.np.aff_a, p.np.aff_b);
              }
              *reinterpret_cast<float2*>(reinterpret_cast<float*>(row) + col) = make_float2(s0, s1);
              if (p.channels == 2)
                *reinterpret_cast<float2*>(reinterpret_cast<float*>(row) + p.out_chan_stride + col) =
                    make_float2(fmaf(fabsf(s0), 2.0f, p.np.lo), fmaf(fabsf(s1), 2.0f, p.np.lo));
            }
          }
        }
      }
    }
    __syncthreads();
    buf ^= 1;
  }
}

template <typename R> constexpr size_t fwd_smem_bytes() {
  return 2 * kFwdBufFloats * sizeof(float) + kFramesPerTile * kXchStride * sizeof(cx<R>);
}

// ================================================================================================
// Inverse
// ================================================================================================
// Stage 16 coefficient rows (frames t0 .. t0+15) de-interleaved: Xe[f][n] = X[2n], Xo[f][n] = X[2n+1].
template <typename R, typename S>
__device__ __forceinline__ void inv_stage_tile(const InvParams& p, int64_t tile, R* buf, int lane) {
  const int64_t b = tile / p.tiles_per_clip;
  const int64_t t0 = (tile - b * p.tiles_per_clip) * kInvFramesOut;
  const S* __restrict__ src = reinterpret_cast<const S*>(p.spec) + b * p.spec_clip_stride;
  const bool aligned = ((reinterpret_cast<uintptr_t>(src) & 15) == 0);
  R* Xe = buf;
  R* Xo = buf + kFramesPerTile * kRowPad;
  constexpr int kQuads = kFramesPerTile * 64;   // groups of 4 coefficients
  constexpr int kBatch = 8;
#pragma unroll 1
  for (int base = 0; base < kQuads; base += 32 * kBatch) {
    S v[kBatch][4];
#pragma unroll
    for (int u = 0; u < kBatch; ++u) {
      const int idx = base + u * 32 + lane;
      const int fr = idx >> 6, q = idx & 63;
      const int64_t t = t0 + fr;
      v[u][0] = v[u][1] = v[u][2] = v[u][3] = (S)0;
      if (t < p.F) {
        const S* s = src + t * kBins + 4 * q;
        if (aligned) {
          if (sizeof(S) == 4) {
            const float4 w = __ldg(reinterpret_cast<const float4*>(s));
            v[u][0] = (S)w.x; v[u][1] = (S)w.y; v[u][2] = (S)w.z; v[u][3] = (S)w.w;
          } else {
            const double2 w0 = __ldg(reinterpret_cast<const double2*>(s));
            const double2 w1 = __ldg(reinterpret_cast<const double2*>(s) + 1);
            v[u][0] = (S)w0.x; v[u][1] = (S)w0.y; v[u][2] = (S)w1.x; v[u][3] = (S)w1.y;
          }
        } else {
          v[u][0] = __ldg(s); v[u][1] = __ldg(s + 1); v[u][2] = __ldg(s + 2); v[u][3] = __ldg(s + 3);
        }
      }
    }
#pragma unroll
    for (int u = 0; u < kBatch; ++u) {
      const int idx = base + u * 32 + lane;
      const int fr = idx >> 6, q = idx & 63;
      typename Vec2<R>::type e, o;
      e.x = (R)v[u][0]; e.y = (R)v[u][2]; o.x = (R)v[u][1]; o.y = (R)v[u][3];
      *reinterpret_cast<typename Vec2<R>::type*>(Xe + fr * kRowPad + 2 * q) = e;
      *reinterpret_cast<typename Vec2<R>::type*>(Xo + fr * kRowPad + 2 * q) = o;
    }
  }
}

template <typename R> constexpr size_t inv_smem_bytes() {
  return (2 * 2 * kFramesPerTile * kRowPad + kFramesPerTile * kURow) * sizeof(R) +
         kFramesPerTile * kXchStride * sizeof(cx<R>);
}

// PRO: 0 = raw coefficients in, 1 = fused denormalise + expand (sinh) on the way in.
template <typename R, typename S, typename OutT, int PRO>
__global__ void __launch_bounds__(kThreads, sizeof(R) == 4 ? 3 : 1) imdct4_inv_kernel(const InvParams p) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  constexpr int kBufElems = 2 * kFramesPerTile * kRowPad;
  R* tilebuf = reinterpret_cast<R*>(smem_raw);                   // [2][kBufElems]
  R* Ubuf = tilebuf + 2 * kBufElems;                             // [16][kURow]
  cx<R>* xch_all = reinterpret_cast<cx<R>*>(Ubuf + kFramesPerTile * kURow);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int g = lane >> 3, j = lane & 7;
  const int f = (warp & 3) * 4 + g;
  const bool loader = (warp == kComputeWarps);

  ThreadTab<R> tt;
  // output-phase mapping: thread c handles samples i4..i4+3 of every other output block
  const int c = threadIdx.x & 127;
  const int i4 = (c & 63) * 4, half = c >> 6;
  float w0[4], w1[4];
  if (!loader) {
    load_T<R>(reinterpret_cast<const R*>(p.tabT), j, (R)1, tt);
#pragma unroll
    for (int u = 0; u < 4; ++u) { w0[u] = __ldg(p.window + i4 + u); w1[u] = __ldg(p.window + 256 + i4 + u); }
  }
  const R sc = (R)(4.0 / 512.0);
  const R inv_gain = (R)1 / (R)p.np.gain;

  int64_t tile = blockIdx.x;
  int buf = 0;
  if (loader && tile < p.ntiles) inv_stage_tile<R, S>(p, tile, tilebuf, lane);
  __syncthreads();
  for (; tile < p.ntiles; tile += gridDim.x) {
    const int64_t next = tile + gridDim.x;
    if (loader) {
      if (next < p.ntiles) inv_stage_tile<R, S>(p, next, tilebuf + (buf ^ 1) * kBufElems, lane);
    } else {
      const int64_t b = tile / p.tiles_per_clip;
      const int64_t t0 = (tile - b * p.tiles_per_clip) * kInvFramesOut;
      const R* Xe = tilebuf + buf * kBufElems + f * kRowPad;
      const R* Xo = Xe + kFramesPerTile * kRowPad;
      cx<R>* xch = xch_all + f * kXchStride;
      {
        cx<R> v[16];
#pragma unroll
        for (int r = 0; r < 16; ++r) {
          const int n = j + 8 * r;
          R a = Xe[n], bb = Xo[127 - n];
          if (PRO == 1) {
            if (sizeof(R) == 8) {
              a = (R)((double)a * (double)p.inv_a + (double)p.inv_b);
              bb = (R)((double)bb * (double)p.inv_a + (double)p.inv_b);
              if (p.np.mode == 1) { a = (R)(sinh((double)a * kLn10F32)) * inv_gain; bb = (R)(sinh((double)bb * kLn10F32)) * inv_gain; }
            } else {
              float af = fmaf((float)a, p.inv_a, p.inv_b), bf = fmaf((float)bb, p.inv_a, p.inv_b);
              if (p.np.mode == 1) { af = fast_sinh(af * (float)kLn10F32) * (float)inv_gain; bf = fast_sinh(bf * (float)kLn10F32) * (float)inv_gain; }
              a = (R)af; bb = (R)bf;
            }
          }
          cx<R> u{a, bb};
          v[r] = (r == 0) ? u : cmulc(u, rho_re<R>(r), rho_im<R>(r));
        }
        pass1<R>(v, tt, j, xch);
      }
      __syncwarp();
      {
        cx<R> y[2][8];
        pass2<R>(xch, j, y);
        R* Urow = Ubuf + f * kURow;
#pragma unroll
        for (int h = 0; h < 2; ++h)
#pragma unroll
          for (int k2 = 0; k2 < 8; ++k2) {
            int col; R d0, d1;
            out_pair<R>(y, j, h, k2, col, d0, d1);
            typename Vec2<R>::type o; o.x = d0; o.y = d1;
            *reinterpret_cast<typename Vec2<R>::type*>(Urow + col) = o;
          }
      }
      asm volatile("bar.sync 1, 128;" ::: "memory");   // the 4 compute warps: U rows complete
      // ---- window + overlap-add + crop: out block q = first half of frame q+1 + second half of frame q
      OutT* dst = reinterpret_cast<OutT*>(p.out) + b * p.out_clip_stride;
      const bool dst_aligned = ((reinterpret_cast<uintptr_t>(dst) & 15) == 0);
#pragma unroll 1
      for (int fo = half; fo < kInvFramesOut; fo += 2) {
        const int64_t q = t0 + fo;
        if (q + 1 >= p.F) break;
        const R* Ua = Ubuf + (fo + 1) * kURow;   // frame q+1 -> first half
        const R* Ub = Ubuf + fo * kURow;         // frame q   -> second half
        R o[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          const int i = i4 + u;
          const R a = unfold_first<R>(Ua, i) * (R)w0[u];
          const R bb = unfold_second<R>(Ub, i) * (R)w1[u];
          o[u] = (a + bb) * sc;
        }
        const int64_t s = q * kHop + i4;
        if (s + 3 < p.out_len && dst_aligned && sizeof(OutT) == 4) {
          *reinterpret_cast<float4*>(reinterpret_cast<float*>(dst) + s) = make_float4((float)o[0], (float)o[1], (float)o[2], (float)o[3]);
        } else if (s + 3 < p.out_len && dst_aligned) {
          double2* d2 = reinterpret_cast<double2*>(reinterpret_cast<double*>(dst) + s);
          d2[0] = make_double2((double)o[0], (double)o[1]);
          d2[1] = make_double2((double)o[2], (double)o[3]);
        } else {
#pragma unroll
          for (int u = 0; u < 4; ++u)
            if (s + u < p.out_len) dst[s + u] = (OutT)o[u];
        }
      }
    }
    __syncthreads();
    buf ^= 1;
  }
}

}  // namespace mdctk
