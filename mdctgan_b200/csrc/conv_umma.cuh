// conv_umma.cuh -- tcgen05 implicit-GEMM convolution for sm_100a (the generator / discriminator hot loop).
//
// Reference layers covered: every nn.Conv2d / nn.ConvTranspose2d of models/networks.py whose channel counts
// are tensor-core shaped (Cin % 4 == 0, Cout % 32 == 0): the ResnetBlock 3x3 convolutions (:440-457, > 80 % of
// the generator FLOPs), the stride-2 down / transposed up layers (:327-330, :347-350) and the PatchGAN 4x4
// layers (:649-670).  The Cin = 2 stem and the Cout = 1 heads stay on the direct kernels of nn_kernels.cuh.
//
// GEMM view:  D[M = B*Ho*Wo pixels, N = Cout] = A[M, K = kh*kw*Cin] * W[K, N]
//   * A is never materialised: eight producer warps gather it from the NHWC input (reflection / zero padding,
//     stride, ConvTranspose2d tap arithmetic), apply the producer layer's deferred InstanceNorm / BatchNorm
//     affine + ReLU / LeakyReLU on the fly, and write it straight into shared memory in the UMMA canonical
//     K-major SWIZZLE_128B layout (one 128-byte row = 32 consecutive input channels of one tap of one pixel).
//   * W arrives by TMA bulk copies (cp.async.bulk -> mbarrier complete_tx) from a pre-packed image that is
//     already in that shared-memory layout ([k-chunk][hi|lo][Cout][32 floats, 16-byte pieces XOR-swizzled]).
//   * One thread issues tcgen05.mma.kind::tf32 (M = 128, N = BN, K = 8) with the accumulator in TMEM.
//     fp32 parity (the reference's CPU path is true fp32; the waveform bar is 1e-3 rel-L2 through a sinh)
//     is kept with the 3xTF32 split: x = hi + lo, both exactly representable in TF32, and
//     D += A_hi*W_hi + A_lo*W_hi + A_hi*W_lo  (relative error ~2^-21 per product, fp32 accumulation).
//   * mbarrier full/empty ring (4 stages) between producers / weight loader and the MMA thread;
//     tcgen05.commit releases stages and publishes the accumulator to the epilogue.
//   * Split-K over a thread-block CLUSTER: at the reference's shapes M is tiny (batch 4 x 4x32 pixels = 512 rows,
//     4 M-tiles x 4 N-tiles), so `splits` (<= 8) CTAs of one cluster each take a slice of K, park their partial
//     tile in shared memory, and after one cluster barrier every CTA reduces its share of the rows over
//     distributed shared memory in rank order (deterministic), adds the bias, takes the InstanceNorm /
//     BatchNorm (sum, sumsq) statistics, applies the epilogue activation and stores NHWC.
#pragma once
#include <cooperative_groups.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include "nn_kernels.cuh"

namespace umma {
namespace cg = cooperative_groups;

constexpr int kBM = 128;              // UMMA M: output pixels per CTA tile
constexpr int kKC = 32;               // K elements per chunk = one 128-byte swizzle row of fp32
constexpr int kProducerWarps = 8;     // gather + epilogue warps
constexpr int kThreads = (kProducerWarps + 2) * 32;   // + MMA warp + weight-loader warp
constexpr int kMaxCin = 1024;
constexpr uint32_t kTf32Mask = 0xFFFFE000u;           // sign + 8 exponent + 10 mantissa bits

struct ConvUmmaParams {
  const float* x; int B, H, W, Cin;
  const float* wp;      // packed weights: [kchunks][2 (hi, lo)][Cout][32], 16-byte pieces swizzled by (n & 7)
  const float* bias;    // [Cout] or null
  float* y; int Ho, Wo, Cout;
  int kh, kw, stride, pad, pad_mode, transposed;
  nnk::InputNorm in;
  int act;
  double* stats;        // [B][Cout][2] or null
  int K, kchunks, splits;
  int classes;          // 1, or stride^2 output parity classes of a ConvTranspose2d (see TileGeom)
  int m_total, m_tiles; // output pixels / 128-row tiles PER CLASS
};

// ---- PTX primitives ------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void fence_barrier_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.release.cta.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.release.cta.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
// Bounded wait: a protocol bug must surface as a trapped kernel (an error code at the C ABI), never as a hung GPU.
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  const uint32_t a = smem_u32(bar);
  uint32_t done = 0;
  long long t0 = 0;
  for (uint32_t spin = 0;; ++spin) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.acquire.cta.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(done)
        : "r"(a), "r"(parity)
        : "memory");
    if (done) return;
    if ((spin & 0x3FFu) == 0x3FFu) {
      const long long t = clock64();
      if (t0 == 0) t0 = t;
      else if (t - t0 > 4000000000LL) __trap();   // ~2 s at 1.9 GHz
    }
  }
}
__device__ __forceinline__ void bulk_g2s(void* dst_smem, const void* src_gmem, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst_smem)),
               "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}
__device__ __forceinline__ void fence_proxy_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

__device__ __forceinline__ void tmem_alloc(uint32_t* slot_smem, uint32_t ncols) {   // whole warp
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(slot_smem)), "r"(ncols) : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {      // whole warp
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
// D[tmem] (+)= A[smem] * B[smem]^T, both operands K-major, TF32 inputs, fp32 accumulation
__device__ __forceinline__ void umma_tf32(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(tmem_d), "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
      : "memory");
}
// arrive on an mbarrier once every previously issued tcgen05.mma of this thread has completed
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
// 32 lanes x 16 consecutive fp32 columns: thread `lane` of the warp receives row (lane_base + lane)
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, float* v) {
  uint32_t r[16];
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]),
        "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr)
      : "memory");
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
  for (int i = 0; i < 16; ++i) v[i] = __uint_as_float(r[i]);
}

// Shared-memory matrix descriptor (cute::UMMA::SmemDescriptor bit layout): K-major operand, SWIZZLE_128B,
// 8-row groups 1024 bytes apart.  Stepping K by 8 TF32 elements inside the 128-byte row = start address + 32 bytes.
__device__ __forceinline__ uint64_t make_desc_k_sw128(uint32_t saddr) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr & 0x3FFFFu) >> 4);   // [0,14)  start address >> 4
  d |= (uint64_t)1 << 16;                     // [16,30) leading byte offset >> 4 (ignored for swizzled K-major)
  d |= (uint64_t)(1024 >> 4) << 32;           // [32,46) stride byte offset >> 4
  d |= (uint64_t)1 << 46;                     // [46,48) descriptor version (Blackwell)
  d |= (uint64_t)2 << 61;                     // [61,64) SWIZZLE_128B
  return d;
}
// Instruction descriptor (cute::UMMA::InstrDescriptor): D fp32, A/B TF32, both K-major, N = BN, M = 128
__host__ __device__ constexpr uint32_t make_idesc_tf32(int bn) {
  return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(bn >> 3) << 17) | ((uint32_t)(kBM >> 4) << 24);
}

// nearest TF32 (10-bit mantissa) value, as an fp32 whose low 13 bits are zero: exact in the tensor core whatever
// rounding the hardware applies to fp32 containers.  x = hi + lo with lo = tf32(x - hi): |x - hi - lo| <= 2^-23 |x|.
__host__ __device__ __forceinline__ float tf32_part(float v) {
#ifdef __CUDA_ARCH__
  return __uint_as_float((__float_as_uint(v) + 0x1000u) & kTf32Mask);
#else
  union { float f; uint32_t u; } c; c.f = v; c.u = (c.u + 0x1000u) & kTf32Mask; return c.f;
#endif
}

template <int BN, bool SPLIT3>
struct Cfg {
  static constexpr int kParts = SPLIT3 ? 2 : 1;
  static constexpr int kABytes = kBM * 128;
  static constexpr int kBBytes = BN * 128;
  static constexpr int kStageBytes = kParts * (kABytes + kBBytes);
  static constexpr int kStages = (192 * 1024 / kStageBytes) > 6 ? 6 : (192 * 1024 / kStageBytes);
  static constexpr int kPitch = BN + 4;                                   // staging row pitch (floats)
  static constexpr int kStagingBytes = ((kBM * kPitch * 4 + 1023) / 1024) * 1024;
  static constexpr int kRedBytes = 2 * 1024 * 4;                          // [2][RP][BN] floats, RP*BN = 1024
  static constexpr int kTmemCols = BN < 32 ? 32 : BN;
  static constexpr int kSmemBytes = 1024 + kStages * kStageBytes + 2 * kMaxCin * 4 + 256;
  static_assert(kStagingBytes + kRedBytes <= kStages * kStageBytes, "staging must fit in the pipeline buffers");
  static_assert(BN == 32 || BN == 64 || BN == 128, "BN");
};

// Which output pixels a CTA tile covers.  Ordinary convolutions: rows are the flattened (b, oy, ox) pixels.
// ConvTranspose2d with stride s: the s*s output parity classes ((oy + pad) % s, (ox + pad) % s) each see only the
// taps ky = py + s*i, kx = px + s*j (for k3 s2 p1: 1, 2, 2 or 4 of the 9 taps), so tiles are formed per class and
// the K loop walks the live taps only -- 4x fewer chunks than zero-filling the dead ones.
struct TileGeom {
  int cs, py, px, offy, offx;   // class stride (1 = no classes), class parities, first oy / ox of the class
  int nkx, cpt;                 // live taps along x; K chunks per tap (Cin / 32)
  int hw, woc;                  // pixels per sample in this class; class-local row width
  int kchunks;                  // K chunks this tile walks
};

__device__ __forceinline__ TileGeom make_geom(const ConvUmmaParams& p, int cls) {
  TileGeom g;
  if (p.classes > 1) {
    g.cs = p.stride; g.py = cls / g.cs; g.px = cls - g.py * g.cs;
    g.offy = ((g.py - p.pad) % g.cs + g.cs) % g.cs;
    g.offx = ((g.px - p.pad) % g.cs + g.cs) % g.cs;
    const int nky = (p.kh - g.py + g.cs - 1) / g.cs;
    g.nkx = (p.kw - g.px + g.cs - 1) / g.cs;
    g.cpt = p.Cin / kKC;
    g.woc = p.Wo / g.cs; g.hw = (p.Ho / g.cs) * g.woc;
    g.kchunks = nky * g.nkx * g.cpt;
  } else {
    g.cs = 1; g.py = g.px = g.offy = g.offx = 0; g.nkx = p.kw; g.cpt = 0;
    g.woc = p.Wo; g.hw = p.Ho * p.Wo; g.kchunks = p.kchunks;
  }
  return g;
}

template <int BN, bool SPLIT3>
__global__ void __launch_bounds__(kThreads, 1) conv2d_umma_kernel(const ConvUmmaParams p) {
  using C = Cfg<BN, SPLIT3>;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);   // SWIZZLE_128B atoms need 1024-byte alignment
  float* s_scale = reinterpret_cast<float*>(smem + C::kStages * C::kStageBytes);
  float* s_shift = s_scale + kMaxCin;
  uint64_t* full = reinterpret_cast<uint64_t*>(s_shift + kMaxCin);
  uint64_t* empty = full + C::kStages;
  uint64_t* tmem_full = empty + C::kStages;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tmem_full + 1);

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  int bx = blockIdx.x;
  const int mt = bx % p.m_tiles; bx /= p.m_tiles;
  const int cls = bx % p.classes, nt = bx / p.classes;
  const TileGeom tg = make_geom(p, cls);
  const int split = blockIdx.y;                       // == rank of this CTA in its cluster (cluster = (1, splits, 1))
  const int m0 = mt * kBM, n0 = nt * BN;
  const int kc_begin = (int)((long long)tg.kchunks * split / p.splits);
  const int kc_end = (int)((long long)tg.kchunks * (split + 1) / p.splits);
  const int nk = kc_end - kc_begin;

  if (warp == kProducerWarps) {
    if (lane == 0) {
      for (int s = 0; s < C::kStages; ++s) { mbar_init(&full[s], kProducerWarps + 1); mbar_init(&empty[s], 1); }
      mbar_init(tmem_full, 1);
      fence_barrier_init();
    }
    __syncwarp();
    tmem_alloc(tmem_slot, C::kTmemCols);
  }
  // The deferred normalisation of the producing layer.  When the whole tile belongs to one sample (or the affine is
  // per channel only) its scale / shift live in shared memory -- taken ready-made, or computed here from the
  // producer's raw (sum, sumsq) statistics (InstanceNorm2d: no separate finalize launch) -- else fetched per row.
  const bool has_norm = p.in.scale != nullptr || p.in.stats != nullptr;
  const int b_first = m0 / tg.hw;
  const int m_last = (m0 + kBM < p.m_total ? m0 + kBM : p.m_total) - 1;
  const bool one_sample = (m_last / tg.hw) == b_first;
  const bool norm_in_smem = has_norm && (one_sample || !p.in.per_sample);
  if (norm_in_smem) {
    const size_t off = p.in.per_sample ? (size_t)b_first * p.Cin : 0;
    if (p.in.stats) {
      const double inv_n = 1.0 / (double)p.in.count;
      for (int c = tid; c < p.Cin; c += kThreads) {
        const double mean = p.in.stats[2 * (off + c)] * inv_n;
        double var = p.in.stats[2 * (off + c) + 1] * inv_n - mean * mean;
        var = var < 0.0 ? 0.0 : var;
        const double rstd = 1.0 / sqrt(var + (double)p.in.eps);
        s_scale[c] = (float)rstd;
        s_shift[c] = (float)(-mean * rstd);
      }
    } else {
      for (int c = tid; c < p.Cin; c += kThreads) { s_scale[c] = __ldg(p.in.scale + off + c); s_shift[c] = __ldg(p.in.shift + off + c); }
    }
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp < kProducerWarps) {
    // ================= A producers: implicit im2col gather -> normalise -> TF32 hi/lo -> swizzled smem =========
    const int j = tid & 7;        // 16-byte piece (4 channels) of the 128-byte row
    const int rb = tid >> 3;      // rows rb, rb+32, rb+64, rb+96
    int oy[4], ox[4], bb[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int g = m0 + rb + 32 * i;
      if (g < p.m_total) {
        bb[i] = g / tg.hw;
        const int pix = g - bb[i] * tg.hw;
        const int oyc = pix / tg.woc;
        oy[i] = oyc * tg.cs + tg.offy;
        ox[i] = (pix - oyc * tg.woc) * tg.cs + tg.offx;
      } else {
        bb[i] = -1; oy[i] = 0; ox[i] = 0;
      }
    }
    // The gather is latency bound (one L2 round trip per chunk), so the loads of kGroup chunks are issued before
    // the first of them is consumed; the stage's `empty` barrier is only needed before the shared-memory stores.
    constexpr int kGroup = C::kStages < 4 ? C::kStages : 4;
    for (int it0 = 0; it0 < nk; it0 += kGroup) {
      float4 v[kGroup][4];
      int cc[kGroup];
      uint32_t okmask = 0;
#pragma unroll
      for (int gi = 0; gi < kGroup; ++gi) {
        const int it = it0 + gi;
        cc[gi] = -1;
#pragma unroll
        for (int i = 0; i < 4; ++i) v[gi][i] = make_float4(0.f, 0.f, 0.f, 0.f);
        if (it < nk) {
          const int kc = kc_begin + it;
          int ky, kx, c;
          bool kvalid = true;
          if (p.classes > 1) {
            const int t = kc / tg.cpt;
            c = (kc - t * tg.cpt) * kKC + j * 4;
            const int ty = t / tg.nkx;
            ky = tg.py + tg.cs * ty;
            kx = tg.px + tg.cs * (t - ty * tg.nkx);
          } else {
            const int k = kc * kKC + j * 4;
            kvalid = k < p.K;
            const int tap = k / p.Cin;
            c = k - tap * p.Cin;
            ky = tap / p.kw;
            kx = tap - ky * p.kw;
          }
          if (kvalid) {
            cc[gi] = c;
#pragma unroll
            for (int i = 0; i < 4; ++i) {
              if (bb[i] >= 0) {
                const int iy = nnk::in_coord(oy[i], ky, p.H, p.stride, p.pad, p.pad_mode, p.transposed);
                const int ix = nnk::in_coord(ox[i], kx, p.W, p.stride, p.pad, p.pad_mode, p.transposed);
                if (iy >= 0 && ix >= 0) {
                  okmask |= 1u << (gi * 4 + i);
                  v[gi][i] = __ldg(reinterpret_cast<const float4*>(p.x + (((size_t)bb[i] * p.H + iy) * p.W + ix) * p.Cin + c));
                }
              }
            }
          }
        }
      }
#pragma unroll
      for (int gi = 0; gi < kGroup; ++gi) {
        const int it = it0 + gi;
        if (it < nk) {
          const int s = it % C::kStages;
          const uint32_t ph = (uint32_t)(it / C::kStages) & 1u;
          mbar_wait(&empty[s], ph ^ 1u);
          const int c = cc[gi];
          float4 sc = make_float4(1.f, 1.f, 1.f, 1.f), sh = make_float4(0.f, 0.f, 0.f, 0.f);
          if (norm_in_smem && c >= 0) {
            sc = *reinterpret_cast<const float4*>(s_scale + c);
            sh = *reinterpret_cast<const float4*>(s_shift + c);
          }
          uint8_t* a_hi = smem + s * C::kStageBytes;
          uint8_t* a_lo = a_hi + C::kABytes;
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            float e[4] = {v[gi][i].x, v[gi][i].y, v[gi][i].z, v[gi][i].w};
            if (okmask & (1u << (gi * 4 + i))) {
              if (has_norm) {
                if (!norm_in_smem) {
                  sc = __ldg(reinterpret_cast<const float4*>(p.in.scale + (size_t)bb[i] * p.Cin + c));
                  sh = __ldg(reinterpret_cast<const float4*>(p.in.shift + (size_t)bb[i] * p.Cin + c));
                }
                e[0] = fmaf(e[0], sc.x, sh.x); e[1] = fmaf(e[1], sc.y, sh.y); e[2] = fmaf(e[2], sc.z, sh.z); e[3] = fmaf(e[3], sc.w, sh.w);
              }
#pragma unroll
              for (int u = 0; u < 4; ++u) e[u] = nnk::apply_act(e[u], p.in.act);
            }
            const int r = rb + 32 * i;
            const uint32_t off = (uint32_t)(r >> 3) * 1024u + (uint32_t)(r & 7) * 128u + (uint32_t)((j ^ (r & 7)) << 4);
            const float4 hi = make_float4(tf32_part(e[0]), tf32_part(e[1]), tf32_part(e[2]), tf32_part(e[3]));
            *reinterpret_cast<float4*>(a_hi + off) = hi;
            if (SPLIT3) {
              *reinterpret_cast<float4*>(a_lo + off) =
                  make_float4(tf32_part(e[0] - hi.x), tf32_part(e[1] - hi.y), tf32_part(e[2] - hi.z), tf32_part(e[3] - hi.w));
            }
          }
          fence_proxy_async_smem();   // generic-proxy stores -> visible to the tensor core's async-proxy reads
          __syncwarp();
          if (lane == 0) mbar_arrive(&full[s]);
        }
      }
    }
    // ================= epilogue part 1: accumulator TMEM -> registers -> staging tile in shared memory =========
    mbar_wait(tmem_full, 0);
    tc_fence_after();
    float* stage_out = reinterpret_cast<float*>(smem);
    const int q = warp & 3, half = warp >> 2;          // warp w may only touch TMEM lanes 32*(w%4) .. +31
    const int row = q * 32 + lane;
    constexpr int kColsPerWarp = BN / 2;
#pragma unroll
    for (int c0 = 0; c0 < kColsPerWarp; c0 += 16) {
      float a[16];
      tmem_ld16(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(half * kColsPerWarp + c0), a);
      float* dst = stage_out + row * C::kPitch + half * kColsPerWarp + c0;
#pragma unroll
      for (int u = 0; u < 16; u += 4) *reinterpret_cast<float4*>(dst + u) = make_float4(a[u], a[u + 1], a[u + 2], a[u + 3]);
    }
    tc_fence_before();
  } else if (warp == kProducerWarps) {
    // ================= MMA issuer (one thread) =================
    if (lane == 0) {
      constexpr uint32_t idesc = make_idesc_tf32(BN);
      for (int it = 0; it < nk; ++it) {
        const int s = it % C::kStages;
        const uint32_t ph = (uint32_t)(it / C::kStages) & 1u;
        mbar_wait(&full[s], ph);
        tc_fence_after();
        const uint32_t a_hi = smem_u32(smem + s * C::kStageBytes);
        const uint32_t a_lo = a_hi + C::kABytes;
        const uint32_t b_hi = a_hi + C::kParts * C::kABytes;
        const uint32_t b_lo = b_hi + C::kBBytes;
#pragma unroll
        for (int k4 = 0; k4 < kKC / 8; ++k4) {
          const uint64_t da = make_desc_k_sw128(a_hi + k4 * 32), db = make_desc_k_sw128(b_hi + k4 * 32);
          umma_tf32(tmem_base, da, db, idesc, (it | k4) != 0 ? 1u : 0u);
          if (SPLIT3) {
            umma_tf32(tmem_base, make_desc_k_sw128(a_lo + k4 * 32), db, idesc, 1u);
            umma_tf32(tmem_base, da, make_desc_k_sw128(b_lo + k4 * 32), idesc, 1u);
          }
        }
        umma_commit(&empty[s]);     // stage reusable once these MMAs have read it
      }
      umma_commit(tmem_full);       // accumulator complete
    }
    __syncwarp();
  } else {
    // ================= weight loader: TMA bulk copies of the pre-swizzled tile image =================
    if (lane == 0) {
      for (int it = 0; it < nk; ++it) {
        const int s = it % C::kStages;
        const uint32_t ph = (uint32_t)(it / C::kStages) & 1u;
        mbar_wait(&empty[s], ph ^ 1u);
        mbar_arrive_expect_tx(&full[s], (uint32_t)(C::kParts * C::kBBytes));
        uint8_t* b_hi = smem + s * C::kStageBytes + C::kParts * C::kABytes;
        int wkc = kc_begin + it;    // chunk index in the packed weights: (ky*kw + kx) * Cin/32 + channel chunk
        if (p.classes > 1) {
          const int t = wkc / tg.cpt, ty = t / tg.nkx;
          wkc = ((tg.py + tg.cs * ty) * p.kw + tg.px + tg.cs * (t - ty * tg.nkx)) * tg.cpt + (wkc - t * tg.cpt);
        }
        const float* src = p.wp + (((size_t)wkc * 2) * p.Cout + n0) * kKC;
        bulk_g2s(b_hi, src, C::kBBytes, &full[s]);
        if (SPLIT3) bulk_g2s(b_hi + C::kBBytes, src + (size_t)p.Cout * kKC, C::kBBytes, &full[s]);
      }
    }
    __syncwarp();
  }

  __syncthreads();   // staging tile complete; every tcgen05 operation of this CTA has retired
  if (warp == kProducerWarps) {
    tc_fence_after();
    tmem_dealloc(tmem_base, C::kTmemCols);
  }
  cg::cluster_group cluster = cg::this_cluster();
  if (p.splits > 1) cluster.sync();   // all partial tiles of the cluster are parked

  // ================= epilogue part 2: split-K reduction over DSMEM, bias, statistics, activation, store =======
  float* stage_out = reinterpret_cast<float*>(smem);
  float* red = reinterpret_cast<float*>(smem + C::kStagingBytes);
  constexpr int CQ = BN / 4;          // float4 columns per row
  constexpr int RP = 256 / CQ;        // rows per pass over 256 threads
  const int r_begin = kBM * split / p.splits, r_end = kBM * (split + 1) / p.splits;
  const int cq = tid % CQ, rg = tid / CQ;
  float ssum[4] = {0.f, 0.f, 0.f, 0.f}, ssq[4] = {0.f, 0.f, 0.f, 0.f};
  if (tid < 256) {
    const int n = n0 + cq * 4;
    float4 bias = make_float4(0.f, 0.f, 0.f, 0.f);
    if (p.bias) bias = __ldg(reinterpret_cast<const float4*>(p.bias + n));
    const float* peer[8];
#pragma unroll
    for (int s = 0; s < 8; ++s) peer[s] = (p.splits > 1 && s < p.splits) ? cluster.map_shared_rank(stage_out, s) : stage_out;
    int cur_b = -1;
    for (int r = r_begin + rg; r < r_end; r += RP) {
      const int g = m0 + r;
      if (g >= p.m_total) break;
      float4 t[8];
#pragma unroll
      for (int s = 0; s < 8; ++s)
        t[s] = s < p.splits ? *reinterpret_cast<const float4*>(peer[s] + r * C::kPitch + cq * 4) : make_float4(0.f, 0.f, 0.f, 0.f);
      float4 acc = bias;
#pragma unroll
      for (int s = 0; s < 8; ++s) { acc.x += t[s].x; acc.y += t[s].y; acc.z += t[s].z; acc.w += t[s].w; }   // rank order: deterministic
      const int b = g / tg.hw;
      if (p.stats) {
        if (!one_sample && b != cur_b) {   // tile spans samples (planes smaller than 128 pixels): flush per sample
          if (cur_b >= 0) {
#pragma unroll
            for (int u = 0; u < 4; ++u) {
              double* st = p.stats + ((size_t)cur_b * p.Cout + n + u) * 2;
              atomicAdd(st, (double)ssum[u]); atomicAdd(st + 1, (double)ssq[u]);
              ssum[u] = 0.f; ssq[u] = 0.f;
            }
          }
          cur_b = b;
        }
        ssum[0] += acc.x; ssq[0] += acc.x * acc.x; ssum[1] += acc.y; ssq[1] += acc.y * acc.y;
        ssum[2] += acc.z; ssq[2] += acc.z * acc.z; ssum[3] += acc.w; ssq[3] += acc.w * acc.w;
      }
      acc.x = nnk::apply_act(acc.x, p.act); acc.y = nnk::apply_act(acc.y, p.act);
      acc.z = nnk::apply_act(acc.z, p.act); acc.w = nnk::apply_act(acc.w, p.act);
      const int pix = g - b * tg.hw;
      const int oyc = pix / tg.woc;
      const int oyo = oyc * tg.cs + tg.offy, oxo = (pix - oyc * tg.woc) * tg.cs + tg.offx;
      *reinterpret_cast<float4*>(p.y + (((size_t)b * p.Ho + oyo) * p.Wo + oxo) * p.Cout + n) = acc;
    }
    if (p.stats && !one_sample && cur_b >= 0) {
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        double* st = p.stats + ((size_t)cur_b * p.Cout + n + u) * 2;
        atomicAdd(st, (double)ssum[u]); atomicAdd(st + 1, (double)ssq[u]);
      }
    }
  }
  if (p.stats && one_sample) {        // block-uniform branch
    if (tid < 256) {
#pragma unroll
      for (int u = 0; u < 4; ++u) { red[rg * BN + cq * 4 + u] = ssum[u]; red[1024 + rg * BN + cq * 4 + u] = ssq[u]; }
    }
    __syncthreads();
    if (tid < BN) {
      float s = 0.f, q = 0.f;
#pragma unroll
      for (int r = 0; r < RP; ++r) { s += red[r * BN + tid]; q += red[1024 + r * BN + tid]; }
      double* st = p.stats + ((size_t)b_first * p.Cout + n0 + tid) * 2;
      atomicAdd(st, (double)s);
      atomicAdd(st + 1, (double)q);
    }
  }
  if (p.splits > 1) cluster.sync();   // no CTA may exit while a peer still reads its shared memory
}

// [K][Cout] fp32 (nn_ops.pack_conv_weight layout) -> the kernel's shared-memory image, TF32 hi / lo parts
__global__ void pack_weight_umma_kernel(const float* __restrict__ w, float* __restrict__ out, int K, int Cout, int kchunks) {
  const size_t total = (size_t)kchunks * kKC * Cout;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
    const int n = (int)(i % Cout);
    const int k = (int)(i / Cout);
    const float v = k < K ? __ldg(w + (size_t)k * Cout + n) : 0.f;
    const float hi = tf32_part(v);
    const float lo = tf32_part(v - hi);
    const int kc = k >> 5, kk = k & 31;
    const int piece = (kk >> 2) ^ (n & 7);
    const size_t dst = (((size_t)kc * 2) * Cout + n) * kKC + piece * 4 + (kk & 3);
    out[dst] = hi;
    out[dst + (size_t)Cout * kKC] = lo;
  }
}

}  // namespace umma
