// mdct_kernels.cuh -- fused MDCT4 / IMDCT4 kernels for sm_100a (n_fft 512, hop 256, 256 bins).
//
// Replaces on the device (reference file:line):
//   forward : models/mdct.py:392-425 (pad, unfold, window, pre-twiddle, 512-pt Z2Z FFT, post-twiddle)
//             + models/pix2pixHD_model.py:96-100,115-123 (arcsinh compress, abs-norm affine)
//             + :400-402 (second channel |s|*2+lo)                       -> ONE kernel, one HBM pass
//   inverse : models/pix2pixHD_model.py:127-133 (denormalise, sinh expand)
//             + models/mdct.py:457-489 (pre-twiddle, FFT, post-twiddle, window, fold/overlap-add, crop)
//
// Kernel shape (both directions): persistent CTAs of `ft/4` warps (ft = frames per tile, 4..16, chosen
// by the host so that ft divides the clip's frame count well).  8 threads own one frame, 4 frames per
// warp.  Input tiles (ft+1 rows of 256 samples forward / ft coefficient rows inverse) are brought in by
// the TMA engine -- one `cp.async.bulk` per 1 KB row, issued by one lane, completing on an mbarrier --
// into a 2-stage ring, so the read stream needs no registers and stays kStages-1 tiles ahead of the
// math; ragged / unaligned / zero-padded rows are filled by warp 0 with plain stores and published by
// the same mbarrier.  The stride-2 TDAC gather reads the raw rows directly (row pitch 264 elements ->
// 2-way bank conflicts on an LSU pipe that is < 20 % busy).  There is NO block-wide barrier in the tile
// loop: stage reuse is a full/empty mbarrier pair per stage (warps arrive on `empty` after their gather,
// warp 0 refills one iteration later), and the inverse's overlap-add is software-pipelined behind two
// split-phase mbarriers (arrive early, wait late), so warps of a CTA drift freely.
// HBM traffic is the algorithmic minimum: every input element is read once (+1/ft row overlap, served
// by L2) and every output written once.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <type_traits>

#include "mdct_core.cuh"

namespace mdctk {

// Per-core-type launch shape.  fp32 core: 2-stage input ring, tiles of up to 16 frames (128 threads).  fp64 core (168 registers,
// 16-byte exchange slots): ONE input stage and tiles of at most 12 frames (96 threads), so that 4 (forward) / 3 (inverse) CTAs
// share an SM -- the fp64 kernels are latency bound (ncu r02c: 2.25 warps per scheduler, 5.5 cycles per issued instruction),
// co-resident CTAs hide the tile load instead of a second stage.
template <typename R> struct KCfg;
template <> struct KCfg<float> { static constexpr int kStages = 2, kMaxFt = 16, kMaxFtInv = 16, kMinBlocksFwd = 5, kMinBlocksInv = 4; };
// r02j sweep on B200 (8192 x 8192, tools/build_variants.sh): the fp64 kernels run at a fixed rate PER WARP (throughput follows the number of
// resident warps: 12 / SM at 168 registers), a second input stage costs warps and loses; the inverse, whose warps meet at two mbarriers
// per tile, gains 20 % from 8-frame tiles (64-thread CTAs, 4 per SM): 0.353 -> 0.285 ms.
#ifndef MDCT_F64_STAGES      // tuning knobs of the fp64 core (tools/build_variants.sh sweeps them)
#define MDCT_F64_STAGES 1
#define MDCT_F64_MAXFT 12
#define MDCT_F64_MAXFT_INV 8
#define MDCT_F64_MINB_FWD 4
#define MDCT_F64_MINB_INV 4
#endif
template <> struct KCfg<double> {
  static constexpr int kStages = MDCT_F64_STAGES, kMaxFt = MDCT_F64_MAXFT, kMaxFtInv = MDCT_F64_MAXFT_INV, kMinBlocksFwd = MDCT_F64_MINB_FWD,
                       kMinBlocksInv = MDCT_F64_MINB_INV;
};

struct FwdParams {
  const float* audio; int64_t audio_stride; int64_t T;
  int64_t B, F; int64_t tiles_per_clip; int64_t ntiles; int ft;
  const void* tabT; const float* tabW;
  void* out; int64_t out_clip_stride; int64_t out_chan_stride; int channels;
  NormParams np;
};

struct InvParams {
  const void* spec; int64_t spec_clip_stride;   // elements; frames are contiguous rows of 256
  int64_t B, F; int64_t tiles_per_clip; int64_t ntiles; int ft;
  const void* tabT; const float* window;
  void* out; int64_t out_clip_stride; int64_t out_len;   // samples written per clip (<= (F-1)*256)
  NormParams np; float inv_a, inv_b;                       // s_src = s*inv_a + inv_b
};

// ---- mbarrier + bulk-copy (TMA, 1-D) primitives ---------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_fence_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.release.cta.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.release.cta.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  const uint32_t a = smem_u32(bar);
  uint32_t done;
  do {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.acquire.cta.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(done)
        : "r"(a), "r"(parity)
        : "memory");
  } while (!done);
}
// Returns true in exactly one warp per use of `cnt`: the one whose lane 0 incremented it last.
__device__ __forceinline__ bool warp_is_last(uint32_t* cnt, int nwarps, int lane) {
  uint32_t old = 0;
  if (lane == 0) old = atomicAdd(cnt, 1u);
  old = __shfl_sync(0xffffffffu, old, 0);
  return old == (uint32_t)(nwarps - 1);
}
// global -> shared bulk copy; `bytes` multiple of 16, both addresses 16-byte aligned
__device__ __forceinline__ void bulk_g2s(void* dst_smem, const void* src_gmem, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst_smem)),
               "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}

// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ float fast_asinh_scaled(float y, float c1) {
  // sign(y) * log2(|y| + sqrt(y^2+1)) * c1
  const float ay = fabsf(y);
  float r;
  asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(fmaf(ay, ay, 1.0f)));
  float l;
  asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(l) : "f"(ay + r));
  return copysignf(l * c1, y);
}
__device__ __forceinline__ float fast_asinh_affine(float y, float c1, float b) {
  // sign(y) * log2(|y| + sqrt(y^2+1)) * c1 + b, the sign carried by the scale so that one FFMA finishes it
  const float ay = fabsf(y);
  float r;
  asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(fmaf(ay, ay, 1.0f)));
  float l;
  asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(l) : "f"(ay + r));
  return fmaf(l, copysignf(c1, y), b);
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

__host__ __device__ constexpr size_t align16(size_t x) { return (x + 15) & ~(size_t)15; }

// ================================================================================================
// Forward
// ================================================================================================
// Stage the ft+1 sample blocks of one tile (warp 0, all lanes).  Row r holds clip samples
// [(t0+r-1)*256, (t0+r)*256); samples outside [0, T) are the reference's zero padding (mdct.py:403).
__device__ __forceinline__ void fwd_produce(const FwdParams& p, int64_t tile, float* stage, uint64_t* bar, int lane) {
  const int64_t b = tile / p.tiles_per_clip;
  const int64_t t0 = (tile - b * p.tiles_per_clip) * p.ft;
  const float* __restrict__ src = p.audio + b * p.audio_stride;
  const bool aligned = ((reinterpret_cast<uintptr_t>(src) & 15) == 0);
  const int rows = p.ft + 1;
  // rows [r_lo, r_hi) lie completely inside [0, T) -> TMA; the others (clip edges) are filled by hand
  int r_lo = 0, r_hi = 0;
  if (aligned) {
    r_lo = (t0 == 0) ? 1 : 0;
    const int64_t full_blocks = p.T / kHop - (t0 - 1);   // rows r < full_blocks end inside the clip
    r_hi = (int)(full_blocks < rows ? (full_blocks > r_lo ? full_blocks : r_lo) : rows);
  }
  if (r_hi - r_lo < rows) {
#pragma unroll 1
    for (int r = 0; r < rows; ++r) {
      if (r >= r_lo && r < r_hi) continue;
      const int64_t s0 = (t0 + r - 1) * kHop;
      float* row = stage + r * kRawPitch;
#pragma unroll
      for (int u = 0; u < kHop / 32; ++u) {
        const int64_t s = s0 + lane + 32 * u;
        row[lane + 32 * u] = (s >= 0 && s < p.T) ? __ldg(src + s) : 0.f;
      }
    }
    __syncwarp();
  }
  if (lane == 0) {
    mbar_arrive_expect_tx(bar, (uint32_t)(r_hi - r_lo) * (kHop * 4));
    const float* g = src + (t0 + r_lo - 1) * kHop;
    float* d = stage + r_lo * kRawPitch;
#pragma unroll 1
    for (int r = r_lo; r < r_hi; ++r, g += kHop, d += kRawPitch) bulk_g2s(d, g, kHop * 4, bar);
  }
}

constexpr int kTabTElems = 8 * 16;       // thread twiddles, [k1][j]
constexpr int kTabWFloats = 8 * 16 * 2;  // window pairs, [r][j][2]
template <typename R> __host__ __device__ constexpr size_t fwd_smem_bytes(int ft) {
  return align16((size_t)KCfg<R>::kStages * (ft + 1) * kRawPitch * sizeof(float)) + align16((size_t)ft * kXchStride * sizeof(cx<R>)) +
         kTabTElems * sizeof(cx<R>) + kTabWFloats * sizeof(R) + 64;
}
// Stage the plan tables in shared memory: sT[k1*8 + j] = T[j][k1] * scale, sW[(r*8 + j)*2 + {0,1}] = (wE, wO)[j][r]
template <typename R>
__device__ __forceinline__ void stage_tables(const R* __restrict__ tabT, const float* __restrict__ tabW, R scale, cx<R>* sT, R* sW) {
  for (int i = threadIdx.x; i < kTabTElems; i += blockDim.x) {
    const int k1 = i >> 3, j = i & 7;
    sT[i] = cx<R>{tabT[(j * 16 + k1) * 2 + 0] * scale, tabT[(j * 16 + k1) * 2 + 1] * scale};
  }
  if (sW) {
    for (int i = threadIdx.x; i < kTabTElems; i += blockDim.x) {
      const int r = i >> 3, j = i & 7;
      sW[2 * i] = (R)tabW[(j * 16 + r) * 2 + 0];         // converted once per CTA (the fp64 core would convert per use)
      sW[2 * i + 1] = (R)tabW[(j * 16 + r) * 2 + 1];
    }
  }
}

// EPI: 0 = raw coefficients (OutT = R or float), 1 = fused compress + abs-norm (OutT = float, 1 or 2 channels).
// EXACT (fp64 core only): the reference's fp32-rounded window products and library asinh (bit-faithful flavour);
// otherwise products are formed in R and the compress epilogue is the fast fp32 one.
template <typename R, int EPI, typename OutT, bool EXACT>
__global__ void __launch_bounds__(8 * KCfg<R>::kMaxFt, KCfg<R>::kMinBlocksFwd) mdct4_fwd_kernel(const FwdParams p) {
  constexpr int kStages = KCfg<R>::kStages;
  extern __shared__ __align__(128) unsigned char smem_raw[];
  const int stage_floats = (p.ft + 1) * kRawPitch;
  float* raw = reinterpret_cast<float*>(smem_raw);
  cx<R>* xch_all = reinterpret_cast<cx<R>*>(smem_raw + align16((size_t)kStages * stage_floats * sizeof(float)));
  cx<R>* sT = reinterpret_cast<cx<R>*>(reinterpret_cast<unsigned char*>(xch_all) + align16((size_t)p.ft * kXchStride * sizeof(cx<R>)));
  R* sW = reinterpret_cast<R*>(sT + kTabTElems);
  uint64_t* full = reinterpret_cast<uint64_t*>(sW + kTabWFloats);
  uint64_t* empty = full + kStages;
  uint32_t* cnt = reinterpret_cast<uint32_t*>(empty + kStages);
  static_assert(EPI == 0 || std::is_same<OutT, float>::value, "the fused epilogue writes fp32");

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int nwarps = blockDim.x >> 5;
  const int g = lane >> 3, j = lane & 7;
  const int f = warp * 4 + g;

  if (threadIdx.x == 0) {
#pragma unroll
    for (int s = 0; s < kStages; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], nwarps); cnt[s] = 0; }
    mbar_fence_init();
  }
  stage_tables<R>(reinterpret_cast<const R*>(p.tabT), p.tabW, (EPI == 1 && p.np.mode == 1) ? (R)p.np.gain : (R)1, sT, sW);
  __syncthreads();

  int64_t tile = blockIdx.x;
  if (warp == 0) {
#pragma unroll 1
    for (int s = 0; s < kStages; ++s) {
      const int64_t tl = tile + (int64_t)s * gridDim.x;
      if (tl < p.ntiles) fwd_produce(p, tl, raw + s * stage_floats, &full[s], lane);
    }
  }

  const SmemT<R> tt{sT + j};
  const SmemW<R> wt{sW + 2 * j};
  const float c1 = (float)(0.6931471805599453 / kLn10F32) * p.np.aff_a;   // log2 -> ln -> /ln10_f32 -> affine
  cx<R>* const xch = xch_all + f * kXchStride;

  uint32_t it = 0;
#pragma unroll 1
  for (; tile < p.ntiles; tile += gridDim.x, ++it) {
    const int s = it % kStages;
    const uint32_t parity = (it / kStages) & 1;
    const int64_t b = tile / p.tiles_per_clip;
    const int64_t t = (tile - b * p.tiles_per_clip) * p.ft + f;
    const bool active = (f < p.ft) && (t < p.F);
    cx<R> v[16];
    mbar_wait(&full[s], parity);
    if (active) {
      const float* row0 = raw + s * stage_floats + f * kRawPitch;
      fwd_gather<R, !EXACT>(row0, row0 + kRawPitch, j, wt, v);
    }
    __syncwarp();
    {   // this warp is done with stage s; the LAST warp to get here refills it with the tile kStages ahead
      const bool last = warp_is_last(&cnt[s], nwarps, lane);
      if (lane == 0) mbar_arrive(&empty[s]);
      if (last) {
        const int64_t next = tile + (int64_t)kStages * gridDim.x;
        if (lane == 0) cnt[s] = 0;
        if (next < p.ntiles) {
          mbar_wait(&empty[s], parity);   // acquire: every warp's reads of the stage are complete
          fwd_produce(p, next, raw + s * stage_floats, &full[s], lane);
        }
      }
    }
    if (active) pass1<R>(v, tt, j, xch);
    __syncwarp();
    if (active) {
      cx<R> y[2][8];
      pass2<R>(xch, j, y);
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
            if (EXACT) {   // "exact" flavour: fp64 core and library asinh, rounded once to fp32
              if (p.np.mode == 1) {
                s0 = (float)(asinh((double)d0) / kLn10F32 * (double)p.np.aff_a + (double)p.np.aff_b);
                s1 = (float)(asinh((double)d1) / kLn10F32 * (double)p.np.aff_a + (double)p.np.aff_b);
              } else {
                s0 = (float)((double)d0 * (double)p.np.aff_a + (double)p.np.aff_b);
                s1 = (float)((double)d1 * (double)p.np.aff_a + (double)p.np.aff_b);
              }
            } else if (p.np.mode == 1) {
              s0 = fast_asinh_affine((float)d0, c1, p.np.aff_b);
              s1 = fast_asinh_affine((float)d1, c1, p.np.aff_b);
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
    __syncwarp();   // the exchange slots are rewritten by the next tile's pass 1
  }
}

// ================================================================================================
// Inverse
// ================================================================================================
// Stage the ft coefficient rows (frames t0 .. t0+ft-1) of one tile; rows past F are left untouched
// (their frames are computed on stale data and never reach an output).
template <typename S>
__device__ __forceinline__ void inv_produce(const InvParams& p, int64_t tile, S* stage, uint64_t* bar, int lane) {
  const int64_t b = tile / p.tiles_per_clip;
  const int64_t t0 = (tile - b * p.tiles_per_clip) * (p.ft - 1);
  const S* __restrict__ src = reinterpret_cast<const S*>(p.spec) + b * p.spec_clip_stride;
  const bool aligned = ((reinterpret_cast<uintptr_t>(src) & 15) == 0);
  int nbulk = 0;
  if (!aligned) {
#pragma unroll 1
    for (int r = 0; r < p.ft; ++r) {
      if (t0 + r >= p.F) break;
      S* row = stage + r * kRawPitch;
#pragma unroll
      for (int u = 0; u < kBins / 32; ++u) row[lane + 32 * u] = __ldg(src + (t0 + r) * kBins + lane + 32 * u);
    }
  } else {
    const int64_t left = p.F - t0;
    nbulk = (int)(left < p.ft ? left : p.ft);
  }
  __syncwarp();
  if (lane == 0) {
    mbar_arrive_expect_tx(bar, (uint32_t)nbulk * (uint32_t)(kBins * sizeof(S)));
#pragma unroll 1
    for (int r = 0; r < nbulk; ++r) bulk_g2s(stage + r * kRawPitch, src + (t0 + r) * kBins, kBins * sizeof(S), bar);
  }
}

#ifndef MDCT_U_F32
#define MDCT_U_F32 0
#endif
// Element type of the inverse's U rows (DCT-IV results awaiting window + overlap-add).  MDCT_U_F32: the mixed flavour (fp64 core, fp32
// tensors) parks them as fp32 -- 8.7 KB less shared memory per 8-frame tile (one more CTA per SM) and an fp32 output phase.
template <typename R, typename OutT, bool EXACT> struct UType { using type = R; };
#if MDCT_U_F32
template <> struct UType<double, float, false> { using type = float; };
#endif
template <typename R, typename S, typename U = R> __host__ __device__ constexpr size_t inv_smem_bytes(int ft) {
  return align16((size_t)KCfg<R>::kStages * ft * kRawPitch * sizeof(S)) + align16((size_t)ft * kXchStride * sizeof(cx<R>)) +
         align16((size_t)ft * kURow * sizeof(U)) + 512 * sizeof(float) + kTabTElems * sizeof(cx<R>) + 64;
}

template <typename R> __device__ __forceinline__ void load4(const R* p, R* o);
template <> __device__ __forceinline__ void load4<float>(const float* p, float* o) {
  const float4 v = *reinterpret_cast<const float4*>(p);
  o[0] = v.x; o[1] = v.y; o[2] = v.z; o[3] = v.w;
}
template <> __device__ __forceinline__ void load4<double>(const double* p, double* o) {
  const double2 a = *reinterpret_cast<const double2*>(p), b = *reinterpret_cast<const double2*>(p + 2);
  o[0] = a.x; o[1] = a.y; o[2] = b.x; o[3] = b.y;
}

// Window + overlap-add + crop of one finished tile: out block q = first half of frame q+1 + second
// half of frame q (mdct.py:473-488), all warps cooperating, 4 samples per thread per step.
template <typename R, typename OutT>
__device__ __forceinline__ void inv_output_phase(const InvParams& p, int64_t tile, const R* Ubuf, const float* wsm) {
  const int fout = p.ft - 1;
  const int64_t b = tile / p.tiles_per_clip;
  const int64_t t0 = (tile - b * p.tiles_per_clip) * fout;
  OutT* dst = reinterpret_cast<OutT*>(p.out) + b * p.out_clip_stride;
  const bool dst_aligned = ((reinterpret_cast<uintptr_t>(dst) & 15) == 0);
#pragma unroll 1
  for (int idx = threadIdx.x; idx < fout * 64; idx += blockDim.x) {
    const int fo = idx >> 6, i4 = (idx & 63) * 4;
    const int64_t q = t0 + fo;
    if (q + 1 >= p.F) break;   // idx is monotone in fo for a given thread
    const R* Ua = Ubuf + (fo + 1) * kURow;   // frame q+1 -> first half
    const R* Ub = Ubuf + fo * kURow;         // frame q   -> second half
    R ua[4], ub[4], o[4];
    const float4 w0 = *reinterpret_cast<const float4*>(wsm + i4);         // window pre-scaled by 4/n_fft
    const float4 w1 = *reinterpret_cast<const float4*>(wsm + 256 + i4);
    if (i4 < 128) {   // first half: U[128+i], second half: -U[127-i]
      load4<R>(Ua + 128 + i4, ua);
      load4<R>(Ub + 124 - i4, ub);
      o[0] = ua[0] * (R)w0.x - ub[3] * (R)w1.x;
      o[1] = ua[1] * (R)w0.y - ub[2] * (R)w1.y;
      o[2] = ua[2] * (R)w0.z - ub[1] * (R)w1.z;
      o[3] = ua[3] * (R)w0.w - ub[0] * (R)w1.w;
    } else {          // first half: -U[383-i], second half: -U[i-128]
      load4<R>(Ua + 380 - i4, ua);
      load4<R>(Ub + i4 - 128, ub);
      o[0] = -(ua[3] * (R)w0.x + ub[0] * (R)w1.x);
      o[1] = -(ua[2] * (R)w0.y + ub[1] * (R)w1.y);
      o[2] = -(ua[1] * (R)w0.z + ub[2] * (R)w1.z);
      o[3] = -(ua[0] * (R)w0.w + ub[3] * (R)w1.w);
    }
    const int64_t sidx = q * kHop + i4;
    if (sidx + 3 < p.out_len && dst_aligned) {
      if (sizeof(OutT) == 4) {
        *reinterpret_cast<float4*>(reinterpret_cast<float*>(dst) + sidx) = make_float4((float)o[0], (float)o[1], (float)o[2], (float)o[3]);
      } else {
        double2* d2 = reinterpret_cast<double2*>(reinterpret_cast<double*>(dst) + sidx);
        d2[0] = make_double2((double)o[0], (double)o[1]);
        d2[1] = make_double2((double)o[2], (double)o[3]);
      }
    } else {
#pragma unroll
      for (int u = 0; u < 4; ++u)
        if (sidx + u < p.out_len) dst[sidx + u] = (OutT)o[u];
    }
  }
}

// PRO: 0 = raw coefficients in, 1 = fused denormalise + expand (sinh) on the way in.
//
// Software pipeline of one warp (tile i, stage s):
//   [warp 0: refill the stage of tile i-1]  wait full[s] -> gather(i) -> arrive empty[s]
//   wait udone(i-1) -> overlap-add + store tile i-1 -> arrive odone(i-1)
//   pass 1 / pass 2 of tile i (private exchange slice) -> wait odone(i-1) -> write U rows(i) -> arrive udone(i)
// so both cross-warp dependencies (U rows complete / U rows free) are split-phase with real work in between.
template <typename R, typename S, typename OutT, int PRO, bool EXACT>
__global__ void __launch_bounds__(8 * KCfg<R>::kMaxFtInv, KCfg<R>::kMinBlocksInv) imdct4_inv_kernel(const InvParams p) {
  constexpr int kStages = KCfg<R>::kStages;
  extern __shared__ __align__(128) unsigned char smem_raw[];
  const int stage_elems = p.ft * kRawPitch;
  unsigned char* sp_ = smem_raw;
  S* raw = reinterpret_cast<S*>(sp_);                 sp_ += align16((size_t)kStages * stage_elems * sizeof(S));
  using UT = typename UType<R, OutT, EXACT>::type;
  cx<R>* xch_all = reinterpret_cast<cx<R>*>(sp_);     sp_ += align16((size_t)p.ft * kXchStride * sizeof(cx<R>));
  UT* Ubuf = reinterpret_cast<UT*>(sp_);              sp_ += align16((size_t)p.ft * kURow * sizeof(UT));
  float* wsm = reinterpret_cast<float*>(sp_);         sp_ += 512 * sizeof(float);
  cx<R>* sT = reinterpret_cast<cx<R>*>(sp_);          sp_ += kTabTElems * sizeof(cx<R>);
  uint64_t* full = reinterpret_cast<uint64_t*>(sp_);
  uint64_t* empty = full + kStages;
  uint64_t* udone = empty + kStages;
  uint64_t* odone = udone + 1;
  uint32_t* cnt = reinterpret_cast<uint32_t*>(odone + 1);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int nwarps = blockDim.x >> 5;
  const int g = lane >> 3, j = lane & 7;
  const int f = warp * 4 + g;
  const int fout = p.ft - 1;   // output blocks per tile

  if (threadIdx.x == 0) {
#pragma unroll
    for (int s = 0; s < kStages; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], nwarps); cnt[s] = 0; }
    mbar_init(udone, nwarps);
    mbar_init(odone, nwarps);
    mbar_fence_init();
  }
  for (int i = threadIdx.x; i < 512; i += blockDim.x) wsm[i] = __ldg(p.window + i) * (float)(4.0 / 512.0);   // exact (power of 2)
  // fp32 fused prologue: sinh(x ln10)/gain = (2^u - 2^-u) * (0.5/gain), u = (s*inv_a + inv_b) * ln10_f32 * log2(e);
  // the constant factor rides on the thread twiddles
  const bool fast_sinh_path = (PRO == 1 && !EXACT && p.np.mode == 1);
  stage_tables<R>(reinterpret_cast<const R*>(p.tabT), nullptr, fast_sinh_path ? (R)(0.5f / p.np.gain) : (R)1, sT, nullptr);
  __syncthreads();

  int64_t tile = blockIdx.x;
  if (warp == 0) {
#pragma unroll 1
    for (int s = 0; s < kStages; ++s) {
      const int64_t tl = tile + (int64_t)s * gridDim.x;
      if (tl < p.ntiles) inv_produce<S>(p, tl, raw + s * stage_elems, &full[s], lane);
    }
  }

  const float kexp = (float)(kLn10F32 * 1.4426950408889634);
  const float ka = fast_sinh_path ? p.inv_a * kexp : p.inv_a, kb = fast_sinh_path ? p.inv_b * kexp : p.inv_b;
  const SmemT<R> tt{sT + j};
  const R inv_gain = (R)1 / (R)p.np.gain;
  cx<R>* const xch = xch_all + f * kXchStride;
  UT* const Urow = Ubuf + f * kURow;

  uint32_t it = 0;
#pragma unroll 1
  for (; tile < p.ntiles; tile += gridDim.x, ++it) {
    const int s = it % kStages;
    const uint32_t parity = (it / kStages) & 1;
    const int64_t b = tile / p.tiles_per_clip;
    const int64_t t0 = (tile - b * p.tiles_per_clip) * fout;
    const bool active = (f < p.ft) && (t0 + f < p.F);
    cx<R> v[16];
    mbar_wait(&full[s], parity);
    if (active) {
      const S* row = raw + s * stage_elems + f * kRawPitch;
#pragma unroll
      for (int r = 0; r < 16; ++r) {
        R a = (R)row[2 * j + 16 * r], bb = (R)row[255 - 2 * j - 16 * r];   // X[2n], X[255-2n], n = j + 8r
        if (PRO == 1) {
          if (EXACT) {
            a = (R)((double)a * (double)p.inv_a + (double)p.inv_b);
            bb = (R)((double)bb * (double)p.inv_a + (double)p.inv_b);
            if (p.np.mode == 1) { a = (R)(sinh((double)a * kLn10F32)) * inv_gain; bb = (R)(sinh((double)bb * kLn10F32)) * inv_gain; }
          } else {
            float af = fmaf((float)a, ka, kb), bf = fmaf((float)bb, ka, kb);
            if (p.np.mode == 1) {
              float pa, qa, pb, qb;
              asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(pa) : "f"(af));
              asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(qa) : "f"(-af));
              asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(pb) : "f"(bf));
              asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(qb) : "f"(-bf));
              af = pa - qa; bf = pb - qb;
            }
            a = (R)af; bb = (R)bf;
          }
        }
        cx<R> u{a, bb};
        v[r] = (r == 0) ? u : cmulc(u, rho_re<R>(r), rho_im<R>(r));
      }
    }
    __syncwarp();
    {   // this warp is done with stage s; the LAST warp to get here refills it with the tile kStages ahead
      const bool last = warp_is_last(&cnt[s], nwarps, lane);
      if (lane == 0) mbar_arrive(&empty[s]);
      if (last) {
        const int64_t next = tile + (int64_t)kStages * gridDim.x;
        if (lane == 0) cnt[s] = 0;
        if (next < p.ntiles) {
          mbar_wait(&empty[s], parity);
          inv_produce<S>(p, next, raw + s * stage_elems, &full[s], lane);
        }
      }
    }
    if (active) pass1<R>(v, tt, j, xch);
    __syncwarp();
    if (it >= 1) {
      mbar_wait(udone, (it - 1) & 1);                               // U rows of tile it-1 complete in every warp
      inv_output_phase<UT, OutT>(p, tile - gridDim.x, Ubuf, wsm);
      __syncwarp();
      if (lane == 0) mbar_arrive(odone);
    }
    cx<R> y[2][8];
    if (active) pass2<R>(xch, j, y);
    if (it >= 1) mbar_wait(odone, (it - 1) & 1);                    // every warp has finished reading U rows of tile it-1
    if (active) {
#pragma unroll
      for (int h = 0; h < 2; ++h)
#pragma unroll
        for (int k2 = 0; k2 < 8; ++k2) {
          int col; R d0, d1;
          out_pair<R>(y, j, h, k2, col, d0, d1);
          typename Vec2<UT>::type o; o.x = (UT)d0; o.y = (UT)d1;
          *reinterpret_cast<typename Vec2<UT>::type*>(Urow + col) = o;
        }
    }
    __syncwarp();
    if (lane == 0) mbar_arrive(udone);
  }
  if (it >= 1) {   // drain: the last tile's overlap-add
    mbar_wait(udone, (it - 1) & 1);
    inv_output_phase<UT, OutT>(p, tile - gridDim.x, Ubuf, wsm);
  }
}

}  // namespace mdctk
