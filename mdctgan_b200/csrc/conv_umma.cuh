// conv_umma.cuh -- tcgen05 implicit-GEMM convolution for sm_100a (the generator / discriminator hot loop).
//
// Reference layers covered: every nn.Conv2d / nn.ConvTranspose2d of models/networks.py whose channel counts
// are tensor-core shaped (Cin % 4 == 0, Cout % 4 == 0; the packed weight image pads Cout to a multiple of 32 with zero rows, the
// epilogue masks them -- the reference's train.sh recipe has 56 / 112 channels): the ResnetBlock 3x3 convolutions (:440-457, > 80 % of
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
constexpr int kProducerWarps = 16;    // gather + epilogue warps (the gather is issue bound: 4 warps per scheduler)
constexpr int kProducerThreads = kProducerWarps * 32;
constexpr int kRows = kBM * 8 / kProducerThreads;   // rows of the A tile per producer thread (8 threads per 128-byte row)
static_assert(kRows == 2, "validity queue packs 2 bits per chunk");
constexpr int kThreads = (kProducerWarps + 2) * 32;   // + MMA warp + weight-loader warp
constexpr int kMaxCin = 1024;
constexpr int kMaxSplits = 16;       // split-K ranks = CTAs of one cluster (8 portable; 16 with cudaFuncAttributeNonPortableClusterSizeAllowed)
constexpr int kNormTab = 4096;        // floats per deferred-normalisation table (scale, shift): samples in a tile x Cin
constexpr uint32_t kTf32Mask = 0xFFFFE000u;           // sign + 8 exponent + 10 mantissa bits

struct ConvUmmaParams {
  const float* x; int B, H, W, Cin;
  const float* wp;      // packed weights: [kchunks][2 (hi, lo)][CoutP][32], 16-byte pieces swizzled by (n & 7); CoutP = Cout rounded up to 32
  const float* bias;    // [Cout] or null
  float* y; int Ho, Wo, Cout;
  int CoutP;            // rows per (chunk, part) of the packed weight image: Cout rounded up to a multiple of 32
  int kh, kw, stride, pad, pad_mode, transposed;
  nnk::InputNorm in;
  int act;
  double* stats;        // [B][Cout][2] or null
  int K, kchunks, splits;
  int classes;          // 1, or stride^2 output parity classes of a ConvTranspose2d (see TileGeom)
  int m_tiles;          // 128-pixel tiles per (sample, class); span: per class, over the pixels of ALL samples back to back
  int span;             // 1: tiles run over the flattened (sample, pixel) index -- planes smaller than a tile share one
  long long* trace;     // debug: per-CTA phase timestamps [CTA][16] (clock64; slot 0 = globaltimer), or null
};

__device__ __forceinline__ void trace_mark(const ConvUmmaParams& p, int slot, bool who) {
  if (p.trace && who) {
    long long t;
    if (slot == 0) asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    else t = clock64();
    p.trace[((size_t)blockIdx.y * gridDim.x + blockIdx.x) * 16 + slot] = t;
  }
}

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
static __device__ __noinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
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
// statistics live in global memory: a plain reduction, not the generic-address atomic (which carries a shared-memory CAS path)
__device__ __forceinline__ void red_add_f64(double* gptr, double v) {
  asm volatile("red.global.add.f64 [%0], %1;" ::"l"(__cvta_generic_to_global(gptr)), "d"(v) : "memory");
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

// 3xTF32 split x = hi + lo.  hi keeps the top 10 mantissa bits (low 13 bits cleared: exact in the tensor core whatever
// rounding the hardware applies to fp32 containers); lo = x - hi is exact in fp32 (<= 13 significant bits) and is
// consumed as TF32 by the tensor core (its truncation of lo leaves a residual <= 2^-21 |x|).  Rounding lo to the nearest TF32
// in the gather was measured (r02): generator output 6.0e-5 instead of 6.3e-5 from the fp64 truth (fp32 FFMA kernels: 3.5e-5) for
// +35 % time in the issue-bound gather loop -- the residual is dominated by the tensor core's fp32 accumulation, not by lo.
__host__ __device__ __forceinline__ float tf32_hi(float v) {
#ifdef __CUDA_ARCH__
  return __uint_as_float(__float_as_uint(v) & kTf32Mask);
#else
  union { float f; uint32_t u; } c; c.f = v; c.u &= kTf32Mask; return c.f;
#endif
}

__device__ __forceinline__ float tf32_rn(float v) { return __uint_as_float((__float_as_uint(v) + 0x1000u) & kTf32Mask); }

template <int BN, bool SPLIT3>
struct Cfg {
  static constexpr int kParts = SPLIT3 ? 2 : 1;
  static constexpr int kABytes = kBM * 128;
  static constexpr int kBBytes = BN * 128;
  static constexpr int kStageBytes = kParts * (kABytes + kBBytes);
  static constexpr int kStages = (192 * 1024 / kStageBytes) >= 8 ? 8 : (192 * 1024 / kStageBytes);   // 3 for BN = 128 with the hi/lo split
  static constexpr int kPitch = BN + 4;                                   // staging row pitch (floats)
  static constexpr int kStagingBytes = ((kBM * kPitch * 4 + 1023) / 1024) * 1024;
  static constexpr int kRedBytes = 2 * kProducerThreads * 4 * 8;          // [2][RP][BN] doubles, RP*BN = 4 * producer threads
  static constexpr int kTmemCols = BN < 32 ? 32 : BN;
  static constexpr int kSmemBytes = 1024 + kStages * kStageBytes + 2 * kNormTab * 4 + 256;
  static_assert(kSmemBytes <= 227 * 1024, "shared memory");
  static_assert(kStagingBytes + kRedBytes <= kStages * kStageBytes, "staging must fit in the pipeline buffers");
  static_assert(BN == 32 || BN == 64 || BN == 128, "BN");
};

// Which output pixels a CTA tile covers: 128 consecutive rows of the flattened (sample, pixel) index of one class.  Planes of
// 128 pixels and more are tiled per sample; smaller planes (the 2x16 bottleneck of the reference's generator: 32 pixels)
// share a tile across samples (`span`), with the per-sample InstanceNorm affine of the input looked up per row and the
// per-sample statistics of the output reduced sample by sample in the epilogue.  Ordinary convolutions: consecutive pixels of
// the sample.  ConvTranspose2d with stride s: the s*s output parity classes ((oy + pad) % s, (ox + pad) % s) each
// see only the taps ky = py + s*i, kx = px + s*j (for k3 s2 p1: 1, 2, 2 or 4 of the 9 taps), so tiles are formed per
// class and the K loop walks the live taps only -- 4x fewer chunks than zero-filling the dead ones.
// Input coordinate of class-local output (oyc, oxc) at class-local tap (ky, kx):  iy = oyb + sgn*ky,
//   convolution:           oyb = oyc*stride - pad,          sgn = +1
//   transposed, classes:   oyb = oyc + (offy + pad - py)/s, sgn = -1   (oy = oyc*s + offy, real tap = py + s*ky)
struct TileGeom {
  int cs, py, px, offy, offx;   // class stride (1 = no classes), class parities, first oy / ox of the class
  int nkx, cpt, ntaps;          // live taps along x; K chunks per tap (Cin / 32, classes only); live taps in total
  int hw, woc;                  // pixels per sample in this class; class-local row width
  int kchunks;                  // K chunks this tile walks
  int sgn, mul, addy, addx;     // oyb = oyc*mul + addy, oxb = oxc*mul + addx
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
    // ragged classes (Ho or Wo not a multiple of the stride): class (py, px) owns the outputs oy = offy + cs*i < Ho, ox = offx + cs*j < Wo
    g.woc = (p.Wo - g.offx + g.cs - 1) / g.cs; g.hw = ((p.Ho - g.offy + g.cs - 1) / g.cs) * g.woc;
    g.ntaps = nky * g.nkx;
    g.kchunks = g.ntaps * g.cpt;
    g.sgn = -1; g.mul = 1; g.addy = (g.offy + p.pad - g.py) / g.cs; g.addx = (g.offx + p.pad - g.px) / g.cs;
  } else {
    g.cs = 1; g.py = g.px = g.offy = g.offx = 0; g.nkx = p.kw; g.cpt = 0; g.ntaps = p.kh * p.kw;
    g.woc = p.Wo; g.hw = p.Ho * p.Wo; g.kchunks = p.kchunks;
    if (p.transposed) { g.sgn = -1; g.mul = 1; g.addy = g.addx = p.pad; }          // stride 1 only (host-checked)
    else { g.sgn = 1; g.mul = p.stride; g.addy = g.addx = -p.pad; }
  }
  return g;
}

__device__ __forceinline__ void cp_async16_zfill(uint32_t dst_smem, const void* src, uint32_t src_bytes) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst_smem), "l"(src), "r"(src_bytes) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }
__device__ __forceinline__ float4 lds128(uint32_t a) {
  float4 v;
  asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(a) : "memory");
  return v;
}
__device__ __forceinline__ void sts128(uint32_t a, float4 v) {
  asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(a), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
}

template <int BN, bool SPLIT3>
__global__ void __launch_bounds__(kThreads, 1) conv2d_umma_kernel(const ConvUmmaParams p) {
  using C = Cfg<BN, SPLIT3>;
  static_assert(C::kStages >= 2 && C::kStages <= 8, "stage count");
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);   // SWIZZLE_128B atoms need 1024-byte alignment
  const uint32_t smem_a = smem_u32(smem);
  float* s_scale = reinterpret_cast<float*>(smem + C::kStages * C::kStageBytes);
  float* s_shift = s_scale + kNormTab;
  uint64_t* full = reinterpret_cast<uint64_t*>(s_shift + kNormTab);
  uint64_t* empty = full + C::kStages;
  uint64_t* tmem_full = empty + C::kStages;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tmem_full + 1);

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  trace_mark(p, 0, tid == 0);
  trace_mark(p, 1, tid == 0);
  int bx = blockIdx.x;
  const int mt = bx % p.m_tiles; bx /= p.m_tiles;
  int b_tile = 0;
  if (!p.span) { b_tile = bx % p.B; bx /= p.B; }
  const int cls = bx % p.classes, nt = bx / p.classes;
  const TileGeom tg = make_geom(p, cls);
  const int split = blockIdx.y;                       // == rank of this CTA in its cluster (cluster = (1, splits, 1))
  // rows of the tile = flattened (sample, class-local pixel) indices [m0, m0 + 128) below row_limit
  const int m0 = (p.span ? 0 : b_tile * tg.hw) + mt * kBM, n0 = nt * BN;
  const int row_limit = p.span ? p.B * tg.hw : (b_tile + 1) * tg.hw;
  const int b_first = m0 / tg.hw;
  const int b_last = min((min(m0 + kBM, row_limit) - 1) / tg.hw, p.B - 1);      // samples the tile touches (b_last < b_first: empty tile)
  const int kc_begin = (int)((long long)tg.kchunks * split / p.splits);
  const int kc_end = (int)((long long)tg.kchunks * (split + 1) / p.splits);
  const int nk = kc_end - kc_begin;

  if (warp == kProducerWarps) {
    if (lane == 0) {
#pragma unroll 1
      for (int s = 0; s < C::kStages; ++s) { mbar_init(&full[s], kProducerWarps + 1); mbar_init(&empty[s], 1); }
      mbar_init(tmem_full, 1);
      fence_barrier_init();
    }
    __syncwarp();
    tmem_alloc(tmem_slot, C::kTmemCols);
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  trace_mark(p, 2, tid == 0);
  const bool has_norm = p.in.stats != nullptr || p.in.scale != nullptr;
  const int tab_rows = (has_norm && p.in.per_sample) ? (b_last - b_first + 1) : 1;   // rows of the normalisation tables

  if (warp < kProducerWarps) {
    // ================= A producers: implicit im2col gather -> normalise -> TF32 hi/lo -> swizzled smem =========
    // The gather is instruction-issue and instruction-fetch bound (short kernel, cold I-cache on every SM), so the
    // loop is rolled and small: cp.async (LDGSTS, zero-fill for padding) drops each thread's 16-byte pieces of chunk
    // `it` straight into their swizzled slot of the stage, kAhead chunks ahead; the same thread later reads its own
    // pieces back, applies  v = max(e, slope*e), e = x*scale + shift  (slope 1 / 0 / 0.2 = none / ReLU / LeakyReLU)
    // and rewrites them in place as TF32 hi (+ lo in the second buffer).  Row offsets are recomputed per tap only.
    constexpr int kAhead = C::kStages - 1;
    constexpr int kRowStep = kProducerThreads / 8;
    const int j = tid & 7;        // 16-byte piece (4 channels) of the 128-byte row
    const int rb = tid >> 3;      // rows rb, rb + 64
    const uint32_t row_off = (uint32_t)(rb >> 3) * 1024u + (uint32_t)(rb & 7) * 128u + (uint32_t)((j ^ (rb & 7)) << 4);
    const float slope = p.in.act == nnk::kActRelu ? 0.f : (p.in.act == nnk::kActLeaky ? 0.2f : 1.f);
    const float* xb = p.x;
    int oyb[kRows], oxb[kRows], xoff[kRows], toff[kRows];      // toff: offset of the row's sample in the normalisation table
    bool rvalid[kRows];
#pragma unroll
    for (int i = 0; i < kRows; ++i) {
      const int row = m0 + rb + kRowStep * i;
      rvalid[i] = row < row_limit;
      const int bs = rvalid[i] ? row / tg.hw : b_first;
      const int pix = row - bs * tg.hw;
      xoff[i] = bs * p.H * p.W * p.Cin;
      toff[i] = tab_rows > 1 ? (bs - b_first) * p.Cin : 0;
      const int oyc = pix / tg.woc;
      oyb[i] = oyc * tg.mul + tg.addy;
      oxb[i] = (pix - oyc * tg.woc) * tg.mul + tg.addx;
    }
    // issue-side walker over the tile's K space (live taps x Cin), and a second channel walker for the process side
    int k0 = kc_begin * kKC + j * 4;
    int tap = k0 / p.Cin;
    int c = k0 - tap * p.Cin;
    int ky = tap / tg.nkx, kx = tap - ky * tg.nkx;
    int c2 = c;
    bool tap_dirty = true;
    int rowoff[kRows];
    uint32_t okq = 0;             // validity bits of the chunks in flight, 2 per chunk, newest in the low bits
    // issue side of the pipeline: the cp.async gathers of chunk `it` (called kAhead chunks ahead of the transform)
    auto issue_chunk = [&](int it) {
        // ---- issue chunk it
      const int s = it % C::kStages;
      mbar_wait(&empty[s], ((uint32_t)(it / C::kStages) & 1u) ^ 1u);
      if (tap_dirty) {
        tap_dirty = false;
#pragma unroll
        for (int i = 0; i < kRows; ++i) {
          int iy = oyb[i] + tg.sgn * ky, ix = oxb[i] + tg.sgn * kx;
          bool ok = rvalid[i] && tap < tg.ntaps;
          if (p.pad_mode == nnk::kPadReflect) {
            iy = iy < 0 ? -iy : iy; iy = iy >= p.H ? 2 * p.H - 2 - iy : iy;
            ix = ix < 0 ? -ix : ix; ix = ix >= p.W ? 2 * p.W - 2 - ix : ix;
          } else {
            ok = ok && (unsigned)iy < (unsigned)p.H && (unsigned)ix < (unsigned)p.W;
          }
          rowoff[i] = ok ? xoff[i] + (iy * p.W + ix) * p.Cin : -1;
        }
      }
      const uint32_t dst = smem_a + s * C::kStageBytes + row_off;
      uint32_t bits = 0;
#pragma unroll
      for (int i = 0; i < kRows; ++i) {
        const bool ok = rowoff[i] >= 0;
        bits |= ok ? (1u << i) : 0u;
        cp_async16_zfill(dst + i * (kRowStep * 128), ok ? xb + rowoff[i] + c : xb, ok ? 16u : 0u);
      }
      cp_async_commit();
      okq = (okq << 2) | bits;
      c += kKC;
      while (c >= p.Cin) {
        c -= p.Cin; ++tap; tap_dirty = true;
        if (++kx == tg.nkx) { kx = 0; ++ky; }
      }
    };
#pragma unroll 1
    for (int it = 0; it < kAhead; ++it)
      if (it < nk) issue_chunk(it);
    {
      // The deferred normalisation of the producing layer -> scale / shift tables in shared memory (one row per sample the tile
      // touches for a per-sample norm, else one): ready-made, derived here from the producer's raw (sum, sumsq) statistics
      // (InstanceNorm2d: no separate finalize launch), or the identity.  Filled by the producer warps AFTER their first
      // gathers (and the first weight copies) are in flight -- only the transform below needs it.
      const int n_tab = tab_rows * p.Cin;
      const size_t off = p.in.per_sample ? (size_t)b_first * p.Cin : 0;
      if (p.in.stats) {
        // four entries per thread and pass: the eight statistics loads are issued together (an L2 round trip each: rolled one by one
        // they were 4 us of this kernel's prologue on a 4 x 512-channel table)
        const double inv_n = 1.0 / (double)p.in.count;
        const double2* st2 = reinterpret_cast<const double2*>(p.in.stats) + off;
#pragma unroll 1
        for (int c0 = tid; c0 < n_tab; c0 += 4 * kProducerThreads) {
          double2 v[4];
#pragma unroll
          for (int u = 0; u < 4; ++u) {
            const int cc = c0 + u * kProducerThreads;
            v[u] = cc < n_tab ? __ldg(st2 + cc) : make_double2(0.0, 1.0);
          }
#pragma unroll
          for (int u = 0; u < 4; ++u) {
            const int cc = c0 + u * kProducerThreads;
            if (cc < n_tab) {      // (warp-uniform: the fp64 divide / square root of padding entries would cost a microsecond on small tables)
              const double mean = v[u].x * inv_n;
              double var = v[u].y * inv_n - mean * mean;
              var = var < 0.0 ? 0.0 : var;
              const double rstd = rsqrt(var + (double)p.in.eps);
              s_scale[cc] = (float)rstd; s_shift[cc] = (float)(-mean * rstd);
            }
          }
        }
      } else if (p.in.scale) {
#pragma unroll 1
        for (int cc = tid; cc < n_tab; cc += kProducerThreads) { s_scale[cc] = __ldg(p.in.scale + off + cc); s_shift[cc] = __ldg(p.in.shift + off + cc); }
      } else {
#pragma unroll 1
        for (int cc = tid; cc < p.Cin; cc += kProducerThreads) { s_scale[cc] = 1.f; s_shift[cc] = 0.f; }
      }
      asm volatile("bar.sync 1, %0;" ::"n"(kProducerThreads) : "memory");      // producer warps only
    }
#pragma unroll 1
    for (int it = kAhead; it < nk + kAhead; ++it) {
      {
        // ---- process chunk q = it - kAhead
        const int q = it - kAhead;
        if (it < nk) cp_async_wait<kAhead - 1>(); else cp_async_wait<0>();
        const int newer = (it < nk ? it : nk) - 1 - q;          // chunks issued after q
        const uint32_t ok2 = (okq >> (2 * newer)) & 3u;
        const int s = q % C::kStages;
        const uint32_t a_hi = smem_a + s * C::kStageBytes + row_off;
#pragma unroll
        for (int i = 0; i < kRows; ++i) {
          float4 e = make_float4(0.f, 0.f, 0.f, 0.f);
          if (ok2 & (1u << i)) {
            const float4 sc = *reinterpret_cast<const float4*>(s_scale + toff[i] + c2);
            const float4 sh = *reinterpret_cast<const float4*>(s_shift + toff[i] + c2);
            e = lds128(a_hi + i * (kRowStep * 128));
            e.x = fmaf(e.x, sc.x, sh.x); e.y = fmaf(e.y, sc.y, sh.y); e.z = fmaf(e.z, sc.z, sh.z); e.w = fmaf(e.w, sc.w, sh.w);
            e.x = fmaxf(e.x, slope * e.x); e.y = fmaxf(e.y, slope * e.y); e.z = fmaxf(e.z, slope * e.z); e.w = fmaxf(e.w, slope * e.w);
          }
          if (SPLIT3) {
            const float4 hi = make_float4(tf32_hi(e.x), tf32_hi(e.y), tf32_hi(e.z), tf32_hi(e.w));
            sts128(a_hi + i * (kRowStep * 128), hi);
            sts128(a_hi + C::kABytes + i * (kRowStep * 128), make_float4(e.x - hi.x, e.y - hi.y, e.z - hi.z, e.w - hi.w));
          } else {
            sts128(a_hi + i * (kRowStep * 128), make_float4(tf32_rn(e.x), tf32_rn(e.y), tf32_rn(e.z), tf32_rn(e.w)));
          }
        }
        c2 += kKC;
        while (c2 >= p.Cin) c2 -= p.Cin;
        fence_proxy_async_smem();   // generic-proxy stores -> visible to the tensor core's async-proxy reads
        __syncwarp();
        if (lane == 0) mbar_arrive(&full[s]);
        if (q == 0) trace_mark(p, 3, tid == 0);
      }
      if (it < nk) issue_chunk(it);
    }
    trace_mark(p, 4, tid == 0);
    // ================= epilogue part 1: accumulator TMEM -> registers -> staging tile in shared memory =========
    mbar_wait(tmem_full, 0);
    tc_fence_after();
    trace_mark(p, 5, tid == 0);
    float* stage_out = reinterpret_cast<float*>(smem);
    const int q4 = warp & 3;                            // warp w may only touch TMEM lanes 32*(w%4) .. +31
    const int row = q4 * 32 + lane;
#pragma unroll 1
    for (int c0 = (warp >> 2) * 16; c0 < BN; c0 += (kProducerWarps / 4) * 16) {   // 16-column slices round-robin over the 4 warps of a lane quarter
      float a[16];
      tmem_ld16(tmem_base + ((uint32_t)(q4 * 32) << 16) + (uint32_t)c0, a);
      float* dst = stage_out + row * C::kPitch + c0;
#pragma unroll
      for (int u = 0; u < 16; u += 4) *reinterpret_cast<float4*>(dst + u) = make_float4(a[u], a[u + 1], a[u + 2], a[u + 3]);
    }
    tc_fence_before();
  } else if (warp == kProducerWarps) {
    // ================= MMA issuer (one thread) =================
    if (lane == 0) {
      constexpr uint32_t idesc = make_idesc_tf32(BN);
#pragma unroll 1
      for (int it = 0; it < nk; ++it) {
        const int s = it % C::kStages;
        mbar_wait(&full[s], (uint32_t)(it / C::kStages) & 1u);
        tc_fence_after();
        const uint32_t a_hi = smem_a + s * C::kStageBytes;
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
      trace_mark(p, 11, true);
    }
    __syncwarp();
  } else {
    // ================= weight loader: TMA bulk copies of the pre-swizzled tile image =================
    if (lane == 0) {
#pragma unroll 1
      for (int it = 0; it < nk; ++it) {
        const int s = it % C::kStages;
        mbar_wait(&empty[s], ((uint32_t)(it / C::kStages) & 1u) ^ 1u);
        mbar_arrive_expect_tx(&full[s], (uint32_t)(C::kParts * C::kBBytes));
        uint8_t* b_hi = smem + s * C::kStageBytes + C::kParts * C::kABytes;
        int wkc = kc_begin + it;    // chunk index in the packed weights: (ky*kw + kx) * Cin/32 + channel chunk
        if (p.classes > 1) {
          const int t = wkc / tg.cpt, ty = t / tg.nkx;
          wkc = ((tg.py + tg.cs * ty) * p.kw + tg.px + tg.cs * (t - ty * tg.nkx)) * tg.cpt + (wkc - t * tg.cpt);
        }
        const float* src = p.wp + (((size_t)wkc * 2) * p.CoutP + n0) * kKC;
        bulk_g2s(b_hi, src, C::kBBytes, &full[s]);
        if (SPLIT3) bulk_g2s(b_hi + C::kBBytes, src + (size_t)p.CoutP * kKC, C::kBBytes, &full[s]);
      }
    }
    __syncwarp();
  }

  __syncthreads();   // staging tile complete; every tcgen05 operation of this CTA has retired
  trace_mark(p, 6, tid == 0);
  if (warp == kProducerWarps) {
    tc_fence_after();
    tmem_dealloc(tmem_base, C::kTmemCols);
  }
  cg::cluster_group cluster = cg::this_cluster();
  if (p.splits > 1) cluster.sync();   // all partial tiles of the cluster are parked
  trace_mark(p, 7, tid == 0);

  // ================= epilogue part 2: split-K reduction over DSMEM, bias, statistics, activation, store =======
  float* stage_out = reinterpret_cast<float*>(smem);
  double* red = reinterpret_cast<double*>(smem + C::kStagingBytes);   // fp64: E[x^2] - mean^2 must survive |mean| >> std
  constexpr int CQ = BN / 4;                  // float4 columns per row
  constexpr int RP = kProducerThreads / CQ;   // rows per pass over the producer threads
  constexpr int kRedHalf = kProducerThreads * 4;
  const int r_begin = kBM * split / p.splits, r_end = kBM * (split + 1) / p.splits;
  const int cq = tid % CQ, rg = tid / CQ;
  const int n = n0 + cq * 4;
  float4 bias = make_float4(0.f, 0.f, 0.f, 0.f);
  const bool col_ok = n < p.Cout;             // Cout % 4 == 0: a column quad is all-valid or all-padding
  if (p.bias && tid < kProducerThreads && col_ok) bias = __ldg(reinterpret_cast<const float4*>(p.bias + n));
  // split-K reduction first, for every row this thread owns (rows r_begin + rg + i*RP), all remote loads in flight together
  constexpr int kMaxRows = kBM / RP;
  constexpr bool kHoist = kMaxRows <= 4;      // BN = 128: 8 rows per thread would spill; it keeps the loads in the sample loop
  float4 accv[kHoist ? kMaxRows : 1];
  if (kHoist && tid < kProducerThreads) {
#pragma unroll
    for (int i = 0; i < (kHoist ? kMaxRows : 1); ++i) {
      const int r = r_begin + rg + i * RP;
      float4 acc = bias;
      if (r < r_end && m0 + r < row_limit) {
        if (p.splits == 1) {                      // (block-uniform) no reduction: the common case on large planes
          const float4 t = *reinterpret_cast<const float4*>(stage_out + r * C::kPitch + cq * 4);
          acc.x += t.x; acc.y += t.y; acc.z += t.z; acc.w += t.w;
        } else {
#pragma unroll
          for (int s = 0; s < kMaxSplits; ++s) {  // rank order: deterministic
            if (s < p.splits) {
              const float4 t = *reinterpret_cast<const float4*>(cluster.map_shared_rank(stage_out, s) + r * C::kPitch + cq * 4);
              acc.x += t.x; acc.y += t.y; acc.z += t.z; acc.w += t.w;
            }
          }
        }
      }
      accv[i] = acc;
    }
  }
  // then one pass per sample the tile touches: its rows are stored, its (sum, sumsq) go out
  // (only the samples that own rows of THIS rank's share [r_begin, r_end): on a spanning tile that is usually one of several)
  const int r_last = min(r_end, row_limit - m0) - 1;
  const int sb_first = r_last >= r_begin ? max(b_first, (m0 + r_begin) / tg.hw) : 1, sb_last = r_last >= r_begin ? min(b_last, (m0 + r_last) / tg.hw) : 0;
#pragma unroll 1
  for (int sb = sb_first; sb <= sb_last; ++sb) {
    const int lo = max(r_begin, sb * tg.hw - m0), hi = min(r_end, min((sb + 1) * tg.hw, row_limit) - m0);
    if (tid < kProducerThreads) {
      double ssum[4] = {0.0, 0.0, 0.0, 0.0}, ssq[4] = {0.0, 0.0, 0.0, 0.0};
      auto finish_row = [&](int r, float4 acc) {      // statistics, activation, store of one reduced row of sample sb
        const int pix = m0 + r - sb * tg.hw;
        if (p.stats) {
          const double a0 = acc.x, a1 = acc.y, a2 = acc.z, a3 = acc.w;
          ssum[0] += a0; ssq[0] = fma(a0, a0, ssq[0]); ssum[1] += a1; ssq[1] = fma(a1, a1, ssq[1]);
          ssum[2] += a2; ssq[2] = fma(a2, a2, ssq[2]); ssum[3] += a3; ssq[3] = fma(a3, a3, ssq[3]);
        }
        acc.x = nnk::apply_act(acc.x, p.act); acc.y = nnk::apply_act(acc.y, p.act);
        acc.z = nnk::apply_act(acc.z, p.act); acc.w = nnk::apply_act(acc.w, p.act);
        const int oyc = pix / tg.woc;
        const int oyo = oyc * tg.cs + tg.offy, oxo = (pix - oyc * tg.woc) * tg.cs + tg.offx;
        if (col_ok) *reinterpret_cast<float4*>(p.y + (((size_t)sb * p.Ho + oyo) * p.Wo + oxo) * p.Cout + n) = acc;
      };
      if (kHoist) {
#pragma unroll
        for (int i = 0; i < (kHoist ? kMaxRows : 1); ++i) {
          const int r = r_begin + rg + i * RP;
          if (r >= lo && r < hi) finish_row(r, accv[i]);
        }
      } else {
#pragma unroll 1
        for (int r = lo + ((rg - (lo - r_begin)) % RP + RP) % RP; r < hi; r += RP) {      // my rows (r = r_begin + rg mod RP) inside [lo, hi)
          float4 acc = bias;
          if (p.splits == 1) {
            const float4 t = *reinterpret_cast<const float4*>(stage_out + r * C::kPitch + cq * 4);
            acc.x += t.x; acc.y += t.y; acc.z += t.z; acc.w += t.w;
          } else {
#pragma unroll
            for (int s = 0; s < kMaxSplits; ++s) {  // rank order: deterministic
              if (s < p.splits) {
                const float4 t = *reinterpret_cast<const float4*>(cluster.map_shared_rank(stage_out, s) + r * C::kPitch + cq * 4);
                acc.x += t.x; acc.y += t.y; acc.z += t.z; acc.w += t.w;
              }
            }
          }
          finish_row(r, acc);
        }
      }
      if (p.stats) {
        // lanes with the same column quad (lane % CQ) hold different row groups: fold them inside the warp first, so that the
        // block-level step below adds one partial per WARP (16) instead of one per row group (up to 64)
#pragma unroll
        for (int off = CQ; off < 32; off <<= 1) {
#pragma unroll
          for (int u = 0; u < 4; ++u) {
            ssum[u] += __shfl_xor_sync(0xffffffffu, ssum[u], off);
            ssq[u] += __shfl_xor_sync(0xffffffffu, ssq[u], off);
          }
        }
        if (lane < CQ) {
#pragma unroll
          for (int u = 0; u < 4; ++u) { red[warp * BN + cq * 4 + u] = ssum[u]; red[kRedHalf + warp * BN + cq * 4 + u] = ssq[u]; }
        }
      }
    }
    if (p.stats) {        // block-uniform branch
      __syncthreads();
      if (tid < BN && hi > lo && n0 + tid < p.Cout) {
        double s = 0.0, q = 0.0;
#pragma unroll
        for (int r = 0; r < kProducerWarps; ++r) { s += red[r * BN + tid]; q += red[kRedHalf + r * BN + tid]; }
        double* st = p.stats + ((size_t)sb * p.Cout + n0 + tid) * 2;
        red_add_f64(st, s);
        red_add_f64(st + 1, q);
      }
      if (sb < sb_last) __syncthreads();   // `red` is rewritten by the next sample's pass
    }
  }
  trace_mark(p, 8, tid == 0);
  trace_mark(p, 9, tid == 0);
  if (p.splits > 1) cluster.sync();   // no CTA may exit while a peer still reads its shared memory
  trace_mark(p, 10, tid == 0);
}

// [K][Cout] fp32 (nn_ops.pack_conv_weight layout) -> the kernel's shared-memory image, TF32 hi / lo parts
static __global__ void pack_weight_umma_kernel(const float* __restrict__ w, float* __restrict__ out, int K, int Cout, int CoutP, int kchunks) {
  const size_t total = (size_t)kchunks * kKC * Cout;      // rows Cout .. CoutP-1 of the image stay zero (the caller clears it)
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
    const int n = (int)(i % Cout);
    const int k = (int)(i / Cout);
    const float v = k < K ? __ldg(w + (size_t)k * Cout + n) : 0.f;
    const float hi = tf32_hi(v);
    const float lo = __uint_as_float((__float_as_uint(v - hi) + 0x1000u) & kTf32Mask);   // nearest TF32 of the remainder
    const int kc = k >> 5, kk = k & 31;
    const int piece = (kk >> 2) ^ (n & 7);
    const size_t dst = (((size_t)kc * 2) * CoutP + n) * kKC + piece * 4 + (kk & 3);
    out[dst] = hi;
    out[dst + (size_t)CoutP * kKC] = lo;
  }
}

}  // namespace umma
