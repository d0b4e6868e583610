// wgrad_umma.cuh -- tcgen05 weight-gradient kernel of nn.Conv2d / nn.ConvTranspose2d for sm_100a.
//
// What autograd derives for the reference's convolutions (models/networks.py:440-457 ResnetBlock, :327-350 down / up
// layers, :649-670 PatchGAN; driven from train.py:185,197):
//
//     dW[k, co] = sum_p A[p, k] * dY[p, co],   k = tap*Cin + ci,  p = (sample, output pixel)
//
// GEMM view: M = k (128 per tile), N = co (BN per tile), reduction over ALL pixels of ALL samples.  Both operands arrive
// with the NON-reduction index contiguous in memory -- the NHWC gather gives [p][32 consecutive ci of one tap] (128 bytes),
// the gradient is [p][co] -- which is the UMMA "MN-major" operand form.  For 32-bit (TF32) operands the only MN-major shared-
// memory layout the tensor core accepts is SWIZZLE_128B_BASE32B (descriptor layout type 1; cute Layout_MN_SW128_32B_Atom:
// Swizzle<2,5,2> o ((8,n),(4,k)) in 16-byte units): one 512-byte atom = 4 reduction rows (pixels) x 128 bytes (32 MN-contiguous
// fp32), 32-byte pieces XOR-swizzled by the row -- a gathered pixel row of 32 channels lands as one swizzled 128-byte row, no
// transpose anywhere.  (The plain SWIZZLE_128B layout with the MN-major bits set computes zeros: measured.)  One
// tcgen05.mma.kind::tf32 (M = 128, N = BN, K = 8) consumes two atoms (8 pixels); fp32-class results come from the same
// 3xTF32 split as the forward path (hi*hi + lo*hi + hi*lo, fp32 accumulation in TMEM).
//
//   * 16 producer warps gather BOTH operands with cp.async (zero-fill for padding / ragged ends) into a 3-4 stage ring,
//     then apply the producer layer's deferred normalisation + activation to A (explicit per-sample or per-channel
//     scale / shift, read through L1), split hi / lo in place, and publish the stage on an mbarrier;
//   * one thread issues the MMAs; tcgen05.commit releases stages and finally publishes the accumulator;
//   * epilogue: TMEM -> registers -> red.global.add straight into the gradient buffer through (s_co, s_ci, s_tap) element
//     strides (16-byte vector reductions when co is the contiguous index), so the reduction over pixels can be split over
//     gridDim.y CTAs without a second pass; dbias = column sums of dY taken by the m-tile-0 CTAs on the way.
#pragma once
#include "conv_umma.cuh"

namespace umma {

struct WgradUmmaParams {
  const float* x; int B, H, W, Cin;
  nnk::InputNorm in;        // explicit scale / shift (or none); the raw-statistics form is resolved by the caller
  const float* dy; int Ho, Wo, Cout;
  int kh, kw, stride, pad, pad_mode, transposed;
  float* dw; long long s_co, s_ci, s_tap;
  float* dbias;
  int K;                    // kh*kw*Cin
  int n_tiles;              // ceil(Cout / BN)
  int P;                    // B*Ho*Wo: reduction length
  int chunks;               // ceil(P / 32)
};

constexpr int kWgPix = 32;                       // pixels per pipeline stage = 4 MMA k-steps
constexpr int kWgThreads = (kProducerWarps + 1) * 32;

template <int NBLK, bool SPLIT3>
struct WgCfg {
  static constexpr int BN = 32 * NBLK;
  static constexpr int kNBlk = NBLK;
  static constexpr int kParts = SPLIT3 ? 2 : 1;
  static constexpr int kABytes = kWgPix * 4 * 128;                 // [8 pixel quads][4 MN blocks][4 rows][128 B]
  static constexpr int kBBytes = kWgPix * kNBlk * 128;
  static constexpr int kStageBytes = kParts * (kABytes + kBBytes);
  static constexpr int kStages = (200 * 1024 / kStageBytes) >= 4 ? 4 : (200 * 1024 / kStageBytes);
  static constexpr int kTmemCols = BN <= 32 ? 32 : (BN <= 64 ? 64 : (BN <= 128 ? 128 : 256));      // power of two >= BN
  static constexpr int kSmemBytes = 1024 + kStages * kStageBytes + BN * 4 + 256;
  static_assert(NBLK >= 1 && NBLK <= 8, "N tile = 1..8 blocks of 32 output channels (UMMA N <= 256)");
  static_assert(kStages >= 2, "stages");
};

// MN-major 32-bit operand, SWIZZLE_128B_BASE32B: 32-element (128-byte) MN blocks `lbo` bytes apart, 4-row reduction atoms `sbo`
// bytes apart
__device__ __forceinline__ uint64_t make_desc_mn_sw128(uint32_t saddr, uint32_t lbo, uint32_t sbo) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr & 0x3FFFFu) >> 4);
  d |= (uint64_t)(lbo >> 4) << 16;
  d |= (uint64_t)(sbo >> 4) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)1 << 61;     // SWIZZLE_128B_BASE32B
  return d;
}
// byte offset of the 16-byte piece c16 of row `pl` (0..31) of MN block j in a stage whose pixel quads are `quad_bytes` apart
__device__ __forceinline__ uint32_t mn_piece_off(int pl, int j, int c16, uint32_t quad_bytes) {
  return (uint32_t)(pl >> 2) * quad_bytes + (uint32_t)j * 512u + (uint32_t)(pl & 3) * 128u +
         (uint32_t)((((c16 >> 1) ^ (pl & 3)) << 5) | ((c16 & 1) << 4));
}
// D fp32, A / B TF32, both MN-major (bits 15 / 16), N = BN, M = 128
__host__ __device__ constexpr uint32_t make_idesc_tf32_mn(int bn) { return make_idesc_tf32(bn) | (1u << 15) | (1u << 16); }

__device__ __forceinline__ void red_add_v4(float* gptr, float a, float b, float c, float d) {
  asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(__cvta_generic_to_global(gptr)), "f"(a), "f"(b), "f"(c), "f"(d) : "memory");
}
__device__ __forceinline__ void red_add_f32(float* gptr, float a) {
  asm volatile("red.global.add.f32 [%0], %1;" ::"l"(__cvta_generic_to_global(gptr)), "f"(a) : "memory");
}

template <int NBLK, bool SPLIT3>
__global__ void __launch_bounds__(kWgThreads, 1) conv_wgrad_umma_kernel(const WgradUmmaParams p) {
  using C = WgCfg<NBLK, SPLIT3>;
  constexpr int BN = C::BN;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  const uint32_t smem_a = smem_u32(smem);
  float* s_bias = reinterpret_cast<float*>(smem + C::kStages * C::kStageBytes);
  uint64_t* full = reinterpret_cast<uint64_t*>(s_bias + BN);
  uint64_t* empty = full + C::kStages;
  uint64_t* tmem_full = empty + C::kStages;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tmem_full + 1);

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int mt = blockIdx.x / p.n_tiles, nt = blockIdx.x - mt * p.n_tiles;
  const int m0 = mt * kBM, n0 = nt * BN;
  const int c_begin = (int)((long long)p.chunks * blockIdx.y / gridDim.y);
  const int c_end = (int)((long long)p.chunks * (blockIdx.y + 1) / gridDim.y);
  const int nk = c_end - c_begin;
  const bool want_bias = p.dbias != nullptr && mt == 0;

  if (warp == kProducerWarps) {
    if (lane == 0) {
#pragma unroll 1
      for (int s = 0; s < C::kStages; ++s) { mbar_init(&full[s], kProducerWarps); mbar_init(&empty[s], 1); }
      mbar_init(tmem_full, 1);
      fence_barrier_init();
    }
    __syncwarp();
    tmem_alloc(tmem_slot, C::kTmemCols);
  }
  if (tid < BN) s_bias[tid] = 0.f;
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  if (nk <= 0) {               // more splits than chunks (host avoids it)
    if (warp == kProducerWarps) tmem_dealloc(tmem_base, C::kTmemCols);
    return;
  }

  if (warp < kProducerWarps) {
    // ================= producers: gather A (im2col rows) and B (dY rows) -> transform -> TF32 hi / lo, swizzled =================
    constexpr int kAhead = C::kStages - 1;
    constexpr int kBPer = (8 * BN + kProducerThreads - 1) / kProducerThreads;      // B pieces per thread per stage: ceil(BN / 64)
    const int c16 = tid & 7;
    // A: piece (p_local = (tid >> 5) + 16 i, MN block j = (tid >> 3) & 3, 16-byte piece c16): its 4 consecutive k = tap*Cin + ci never
    // straddle a tap (Cin % 4 == 0); tap and first channel are thread constants
    const int ja = (tid >> 3) & 3;
    const int kpiece = m0 + 32 * ja + 4 * c16;
    const bool a_live = kpiece < p.K;
    const int tap = a_live ? kpiece / p.Cin : 0;
    const int ci0 = a_live ? kpiece - tap * p.Cin : 0;
    const int ky = tap / p.kw, kx = tap - ky * p.kw;
    const int HWo = p.Ho * p.Wo;
    const bool has_norm = p.in.scale != nullptr;
    const float slope = p.in.act == nnk::kActRelu ? 0.f : (p.in.act == nnk::kActLeaky ? 0.2f : 1.f);
    float4 sc = make_float4(1.f, 1.f, 1.f, 1.f), sh = make_float4(0.f, 0.f, 0.f, 0.f);
    if (has_norm && !p.in.per_sample && a_live) {
      sc = __ldg(reinterpret_cast<const float4*>(p.in.scale + ci0));
      sh = __ldg(reinterpret_cast<const float4*>(p.in.shift + ci0));
    }
    // B: piece id = tid + 512 i: c16 = id & 7, block jb = (id >> 3) % kNBlk, p_local = id / (8 kNBlk)
    float bsum[kBPer][4];          // column sums of dY, per B piece of this thread (pieces of one thread sit in different N blocks)
#pragma unroll
    for (int i = 0; i < kBPer; ++i) bsum[i][0] = bsum[i][1] = bsum[i][2] = bsum[i][3] = 0.f;
    uint32_t okq = 0;             // validity bits of the A pieces of the chunks in flight, 2 per chunk, newest in the low bits

#pragma unroll 1
    for (int it = 0; it < nk + kAhead; ++it) {
      if (it >= kAhead) {
        // ---- process chunk q
        const int q = it - kAhead;
        if (it < nk) cp_async_wait<kAhead - 1>(); else cp_async_wait<0>();
        const int newer = (it < nk ? it : nk) - 1 - q;
        const uint32_t ok2 = (okq >> (2 * newer)) & 3u;
        const int s = q % C::kStages;
        const uint32_t st_a = smem_a + s * C::kStageBytes;
        const uint32_t st_b = st_a + C::kParts * C::kABytes;
#pragma unroll
        for (int i = 0; i < 2; ++i) {
          const int pl = (tid >> 5) + 16 * i;
          const uint32_t off = mn_piece_off(pl, ja, c16, 2048u);
          float4 e = make_float4(0.f, 0.f, 0.f, 0.f);
          if (ok2 & (1u << i)) {
            e = lds128(st_a + off);
            if (has_norm) {
              if (p.in.per_sample) {
                const size_t o = (size_t)(((c_begin + q) * kWgPix + pl) / HWo) * p.Cin + ci0;
                sc = __ldg(reinterpret_cast<const float4*>(p.in.scale + o));
                sh = __ldg(reinterpret_cast<const float4*>(p.in.shift + o));
              }
              e.x = fmaf(e.x, sc.x, sh.x); e.y = fmaf(e.y, sc.y, sh.y); e.z = fmaf(e.z, sc.z, sh.z); e.w = fmaf(e.w, sc.w, sh.w);
            }
            e.x = fmaxf(e.x, slope * e.x); e.y = fmaxf(e.y, slope * e.y); e.z = fmaxf(e.z, slope * e.z); e.w = fmaxf(e.w, slope * e.w);
          }
          if (SPLIT3) {
            const float4 hi = make_float4(tf32_hi(e.x), tf32_hi(e.y), tf32_hi(e.z), tf32_hi(e.w));
            sts128(st_a + off, hi);
            sts128(st_a + C::kABytes + off, make_float4(e.x - hi.x, e.y - hi.y, e.z - hi.z, e.w - hi.w));
          } else {
            sts128(st_a + off, make_float4(tf32_rn(e.x), tf32_rn(e.y), tf32_rn(e.z), tf32_rn(e.w)));
          }
        }
#pragma unroll
        for (int i = 0; i < kBPer; ++i) {
          const int id = tid + kProducerThreads * i;
          if (id < 8 * BN) {
            const int jb = (id >> 3) % C::kNBlk, pl = id / (8 * C::kNBlk);
            const uint32_t off = mn_piece_off(pl, jb, c16, (uint32_t)(C::kNBlk * 512));
            const float4 e = lds128(st_b + off);          // zero-filled beyond P
            if (want_bias) { bsum[i][0] += e.x; bsum[i][1] += e.y; bsum[i][2] += e.z; bsum[i][3] += e.w; }
            if (SPLIT3) {
              const float4 hi = make_float4(tf32_hi(e.x), tf32_hi(e.y), tf32_hi(e.z), tf32_hi(e.w));
              sts128(st_b + off, hi);
              sts128(st_b + C::kBBytes + off, make_float4(e.x - hi.x, e.y - hi.y, e.z - hi.z, e.w - hi.w));
            } else {
              sts128(st_b + off, make_float4(tf32_rn(e.x), tf32_rn(e.y), tf32_rn(e.z), tf32_rn(e.w)));
            }
          }
        }
        fence_proxy_async_smem();
        __syncwarp();
        if (lane == 0) mbar_arrive(&full[s]);
      }
      if (it < nk) {
        // ---- issue chunk it
        const int s = it % C::kStages;
        mbar_wait(&empty[s], ((uint32_t)(it / C::kStages) & 1u) ^ 1u);
        const uint32_t st_a = smem_a + s * C::kStageBytes;
        const uint32_t st_b = st_a + C::kParts * C::kABytes;
        const int pbase = (c_begin + it) * kWgPix;
        uint32_t bits = 0;
#pragma unroll
        for (int i = 0; i < 2; ++i) {
          const int pl = (tid >> 5) + 16 * i;
          const int pp = pbase + pl;
          const uint32_t dst = st_a + mn_piece_off(pl, ja, c16, 2048u);
          const float* src = p.x;
          bool ok = a_live && pp < p.P;
          if (ok) {
            const int b = pp / HWo;
            const int r = pp - b * HWo;
            const int oy = r / p.Wo, ox = r - oy * p.Wo;
            const int iy = nnk::in_coord(oy, ky, p.H, p.stride, p.pad, p.pad_mode, p.transposed);
            const int ix = nnk::in_coord(ox, kx, p.W, p.stride, p.pad, p.pad_mode, p.transposed);
            ok = iy >= 0 && ix >= 0;
            if (ok) src = p.x + (((size_t)b * p.H + iy) * p.W + ix) * p.Cin + ci0;
          }
          bits |= ok ? (1u << i) : 0u;
          cp_async16_zfill(dst, src, ok ? 16u : 0u);
        }
#pragma unroll
        for (int i = 0; i < kBPer; ++i) {
          const int id = tid + kProducerThreads * i;
          if (id < 8 * BN) {
            const int jb = (id >> 3) % C::kNBlk, pl = id / (8 * C::kNBlk);
            const int pp = pbase + pl;
            const uint32_t dst = st_b + mn_piece_off(pl, jb, c16, (uint32_t)(C::kNBlk * 512));
            const int co = n0 + 32 * jb + 4 * c16;
            const bool ok = pp < p.P && co < p.Cout;          // Cout % 4 == 0: a piece is all-valid or all-padding
            cp_async16_zfill(dst, ok ? p.dy + (size_t)pp * p.Cout + co : p.dy, ok ? 16u : 0u);
          }
        }
        cp_async_commit();
        okq = (okq << 2) | bits;
      }
    }
    // ---- dbias: column sums of dY (m-tile 0 only), block-reduced through shared memory
    if (want_bias) {
#pragma unroll
      for (int i = 0; i < kBPer; ++i) {
        const int id = tid + kProducerThreads * i;
        if (id < 8 * BN) {
          const int jb = (id >> 3) % C::kNBlk;
#pragma unroll
          for (int u = 0; u < 4; ++u) atomicAdd(&s_bias[jb * 32 + c16 * 4 + u], bsum[i][u]);
        }
      }
    }
    // ================= epilogue: accumulator TMEM -> registers -> red.global.add into dW =================
    mbar_wait(tmem_full, 0);
    tc_fence_after();
    const int q4 = warp & 3;
    const int k = m0 + q4 * 32 + lane;
    const bool row_ok = k < p.K;
    const int tp = row_ok ? k / p.Cin : 0;
    const int ci = row_ok ? k - tp * p.Cin : 0;
    float* rowp = p.dw + (long long)ci * p.s_ci + (long long)tp * p.s_tap + (long long)n0 * p.s_co;
#pragma unroll 1
    for (int c0 = (warp >> 2) * 16; c0 < BN; c0 += (kProducerWarps / 4) * 16) {
      float a[16];
      tmem_ld16(tmem_base + ((uint32_t)(q4 * 32) << 16) + (uint32_t)c0, a);      // whole warp (sync.aligned)
      if (row_ok) {
        if (p.s_co == 1) {
#pragma unroll
          for (int u = 0; u < 16; u += 4)
            if (n0 + c0 + u < p.Cout) red_add_v4(rowp + c0 + u, a[u], a[u + 1], a[u + 2], a[u + 3]);
        } else {
#pragma unroll
          for (int u = 0; u < 16; ++u)
            if (n0 + c0 + u < p.Cout) red_add_f32(rowp + (long long)(c0 + u) * p.s_co, a[u]);
        }
      }
    }
    tc_fence_before();
  } else {
    // ================= MMA issuer (one thread) =================
    if (lane == 0) {
      constexpr uint32_t idesc = make_idesc_tf32_mn(BN);
      constexpr uint32_t kSboB = C::kNBlk * 512;        // bytes between the 4-pixel atoms of B; a k-step (8 pixels) is two of them
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
        for (int g = 0; g < kWgPix / 8; ++g) {
          const uint64_t da = make_desc_mn_sw128(a_hi + g * 4096, 512, 2048), db = make_desc_mn_sw128(b_hi + g * 2 * kSboB, 512, kSboB);
          umma_tf32(tmem_base, da, db, idesc, (it | g) != 0 ? 1u : 0u);
          if (SPLIT3) {
            umma_tf32(tmem_base, make_desc_mn_sw128(a_lo + g * 4096, 512, 2048), db, idesc, 1u);
            umma_tf32(tmem_base, da, make_desc_mn_sw128(b_lo + g * 2 * kSboB, 512, kSboB), idesc, 1u);
          }
        }
        umma_commit(&empty[s]);
      }
      umma_commit(tmem_full);
    }
    __syncwarp();
  }

  __syncthreads();
  if (warp == kProducerWarps) {
    tc_fence_after();
    tmem_dealloc(tmem_base, C::kTmemCols);
  }
  if (want_bias && tid < BN && n0 + tid < p.Cout) red_add_f32(p.dbias + n0 + tid, s_bias[tid]);
}

}  // namespace umma
