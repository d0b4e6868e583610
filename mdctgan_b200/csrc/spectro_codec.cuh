// spectro_codec.cuh -- the secondary spectrogram encodings of Audio2MDCT (reference models/pix2pixHD_model.py:83-163), the ones the
// fused transform kernels do not carry: the dB magnitude encoding (default options: neither --arcsinh_transform nor --raw_mdct),
// --explicit_encoding (two dB channels of the alpha-mixed positive / negative parts) and the per-sample min / max normalisation
// (no --abs_norm) of any encoding.  Off the hot path (every shipped script trains with arcsinh + abs_norm), so these are plain
// element-wise fp64 kernels around the raw MDCT4 / IMDCT4 launches: encode (+ per-plane min / max) -> affine; affine^-1 -> decode.
// The arithmetic is fp64 like the reference's (its spectrograms are fp64); min / max are rounded to fp32 where the reference does (:112-115).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace codec {

constexpr int kModeRaw = 0, kModeArcsinh = 1, kModeDb = 2, kModeExplicit = 3;

struct EncodeParams {
  const void* spec; int spec_f64;      // raw MDCT coefficients [B][plane], fp32 or fp64
  long long B, plane;
  int mode; double gain, alpha, min_value;
  double* enc;                         // [B][C][plane], C = 2 for explicit, else 1
  float* sign;                         // [B][plane] torch.sign(spectro) (pix2pixHD_model.py:36) or null
  unsigned* minmax_key;                // [B][C][2] order-preserving keys of the fp32 (min, max) of every plane, or null
};

// torchaudio.functional.amplitude_to_DB(x, multiplier = 20, amin, db_multiplier = 1): 20 log10(max(x, amin)) - 20
__device__ __forceinline__ double amp_to_db(double x, double amin) { return 20.0 * log10(x > amin ? x : amin) - 20.0; }

// monotone map float -> unsigned (atomicMin / atomicMax on the keys order like the floats)
__device__ __forceinline__ unsigned f2key(float f) { const unsigned u = __float_as_uint(f); return (u & 0x80000000u) ? ~u : (u | 0x80000000u); }
__host__ __device__ __forceinline__ float key2f(unsigned k) {
  const unsigned u = (k & 0x80000000u) ? (k & 0x7FFFFFFFu) : ~k;
#ifdef __CUDA_ARCH__
  return __uint_as_float(u);
#else
  union { unsigned u; float f; } c; c.u = u; return c.f;
#endif
}

__device__ __forceinline__ void encode_one(const EncodeParams& p, double s, double* e0, double* e1) {
  *e1 = 0.0;
  if (p.mode == kModeExplicit) {            // :84-95: neg = (|s| - s)/2, pos = s + neg
    const double neg = 0.5 * (fabs(s) - s), pos = s + neg;
    *e0 = amp_to_db(p.alpha * pos + (1.0 - p.alpha) * neg, p.min_value);
    *e1 = amp_to_db((1.0 - p.alpha) * pos + p.alpha * neg, p.min_value);
  } else if (p.mode == kModeArcsinh) {      // :96-100 (ln10 is the fp32 constant)
    *e0 = asinh(p.gain * s) / 2.3025851249694824;
  } else if (p.mode == kModeRaw) {
    *e0 = s;
  } else {                                  // :104-106
    *e0 = amp_to_db(fabs(s) + p.min_value, p.min_value);
  }
}

// grid (chunks, B): block-level min / max, one atomic pair per channel per CTA
__global__ void __launch_bounds__(256) spectro_encode_kernel(const EncodeParams p) {
  const long long b = blockIdx.y;
  const int C = p.mode == kModeExplicit ? 2 : 1;
  float mn[2] = {INFINITY, INFINITY}, mx[2] = {-INFINITY, -INFINITY};
  for (long long i = (long long)blockIdx.x * 256 + threadIdx.x; i < p.plane; i += (long long)gridDim.x * 256) {
    const double s = p.spec_f64 ? reinterpret_cast<const double*>(p.spec)[b * p.plane + i] : (double)reinterpret_cast<const float*>(p.spec)[b * p.plane + i];
    double e[2];
    encode_one(p, s, &e[0], &e[1]);
    for (int c = 0; c < C; ++c) {
      p.enc[(b * C + c) * p.plane + i] = e[c];
      const float f = (float)e[c];
      mn[c] = fminf(mn[c], f); mx[c] = fmaxf(mx[c], f);
    }
    if (p.sign) p.sign[b * p.plane + i] = s > 0.0 ? 1.f : (s < 0.0 ? -1.f : 0.f);
  }
  if (p.minmax_key) {
    __shared__ float red[2][2][8];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    for (int c = 0; c < C; ++c) {
      float a = mn[c], z = mx[c];
      for (int off = 16; off; off >>= 1) { a = fminf(a, __shfl_xor_sync(0xffffffffu, a, off)); z = fmaxf(z, __shfl_xor_sync(0xffffffffu, z, off)); }
      if (lane == 0) { red[c][0][warp] = a; red[c][1][warp] = z; }
    }
    __syncthreads();
    if (threadIdx.x < 2 * C) {
      const int c = threadIdx.x >> 1, which = threadIdx.x & 1;
      float v = red[c][which][0];
      for (int w = 1; w < 8; ++w) v = which ? fmaxf(v, red[c][which][w]) : fminf(v, red[c][which][w]);
      unsigned* dst = p.minmax_key + ((size_t)b * C + c) * 2 + which;
      if (which) atomicMax(dst, f2key(v)); else atomicMin(dst, f2key(v));
    }
  }
}

// keys -> fp32 (min, max) in place (the buffer is then read as float)
__global__ void minmax_keys_to_float_kernel(unsigned* keys, int n) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) keys[i] = __float_as_uint(key2f(keys[i]));
}
__global__ void minmax_keys_init_kernel(unsigned* keys, int n) {      // (min, max) pairs: +inf key, -inf key
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) keys[i] = (i & 1) ? 0u : 0xFFFFFFFFu;
}

struct AffineParams {
  long long planes, plane;             // planes = B * C
  const float* minmax;                 // [planes][2] fp32 (min, max) or null -> (src_lo, src_hi)
  double src_lo, src_hi, norm_lo, norm_hi;
};

// :116-123: (x - min) / (max - min) * (hi - lo) + lo, fp64, rounded to fp32 once (to_spectro returns .float())
__global__ void __launch_bounds__(256) spectro_affine_kernel(const double* __restrict__ enc, float* __restrict__ out, const AffineParams p) {
  const long long pl = blockIdx.y;
  // audio_min / audio_max are fp32 tensors in the reference: their difference is an fp32 subtraction, the rest promotes to fp64
  const float lo_f = p.minmax ? p.minmax[2 * pl] : (float)p.src_lo, hi_f = p.minmax ? p.minmax[2 * pl + 1] : (float)p.src_hi;
  const double lo = (double)lo_f, span = (double)__fsub_rn(hi_f, lo_f);
  for (long long i = (long long)blockIdx.x * 256 + threadIdx.x; i < p.plane; i += (long long)gridDim.x * 256) {
    double s = (enc[pl * p.plane + i] - lo) / span;
    out[pl * p.plane + i] = (float)(s * (p.norm_hi - p.norm_lo) + p.norm_lo);
  }
}

struct DecodeParams {
  const float* s;                      // normalised spectrogram [B][C][plane] fp32
  long long B, plane;
  int mode; double gain, alpha, min_value;
  const float* minmax;                 // [B][C][2] or null
  double src_lo, src_hi, norm_lo, norm_hi;
  const float* pha;                    // [B][plane] sign / pseudo-phase multiplier (dB mode, :150-157) or null
  double* out;                         // raw MDCT coefficients [B][plane] fp64
};

// denormalize (:127-137) + the channel recombination / phase product of to_audio (:142-157)
__global__ void __launch_bounds__(256) spectro_decode_kernel(const DecodeParams p) {
  const long long b = blockIdx.y;
  const int C = p.mode == kModeExplicit ? 2 : 1;
  for (long long i = (long long)blockIdx.x * 256 + threadIdx.x; i < p.plane; i += (long long)gridDim.x * 256) {
    double d[2] = {0.0, 0.0};
    for (int c = 0; c < C; ++c) {
      const long long pl = b * C + c;
      const float lo_f = p.minmax ? p.minmax[2 * pl] : (float)p.src_lo, hi_f = p.minmax ? p.minmax[2 * pl + 1] : (float)p.src_hi;
      double x = ((double)p.s[pl * p.plane + i] - p.norm_lo) / (p.norm_hi - p.norm_lo);
      x = x * (double)__fsub_rn(hi_f, lo_f) + (double)lo_f;      // (max - min): an fp32 subtraction in the reference (:130)
      if (p.mode == kModeArcsinh) x = sinh(x * 2.3025851249694824) / p.gain;
      else if (p.mode != kModeRaw) x = 10.0 * pow(pow(10.0, 0.1 * x), 0.5) - p.min_value;      // aF.DB_to_amplitude(x, 10.0, 0.5) - min_value
      d[c] = x;
    }
    double v = d[0];
    if (p.mode == kModeExplicit) v = (d[0] - d[1]) / (2.0 * p.alpha - 1.0);
    else if (p.mode == kModeDb && p.pha) v *= (double)p.pha[b * p.plane + i];
    p.out[b * p.plane + i] = v;
  }
}

}  // namespace codec
