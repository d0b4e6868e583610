// mdct_core.cuh -- arithmetic core of the fused MDCT4 / IMDCT4 kernels (n_fft = 512, 256 bins).
//
// What is computed (reference semantics: models/mdct.py:392-425 forward, :457-489 inverse; the
// reference evaluates the lapped DCT-IV with a zero-padded 512-point complex128 FFT):
//
//     X[t,k] = sum_{m<512} w[m] x_pad[256 t + m] cos(pi/256 (m + 1/2 + 128)(k + 1/2))
//
// How it is computed here (not the reference's formulation): TDAC fold of the 512 windowed
// samples to 256, then the 256-point DCT-IV as ONE 128-point complex FFT with pre/post twiddles,
// the FFT split 16 x 8 so that 8 threads own one frame:
//     pass 1: thread j (0..7) owns points n = j + 8r, r = 0..15 -> radix-16 butterfly in registers
//     exchange through shared memory (swizzled, conflict free)
//     pass 2: thread a (0..7) owns k1 in {a, 15-a} -> two radix-8 butterflies; the pairing makes
//             every thread hold ADJACENT output bins (X[2k], X[2k+1]) -> 8-byte coalesced stores.
// All thread-independent twiddles are compile-time constants (mdct_consts.h); the thread-dependent
// ones (T[j][k1], window) come from the plan tables and live in registers for the kernel's life.
//
// Everything here is __host__ __device__ so that tests/emu/ can run the very same arithmetic and
// index maps on the CPU, thread by thread, against the oracle (test infrastructure; the product
// never runs this on the host).
#pragma once
#include "mdct_consts.h"

namespace mdctk {

constexpr int kBins = 256;      // N  = n_fft / 2
constexpr int kFft = 128;       // N/2 complex points
constexpr int kHop = 256;
constexpr int kWin = 512;
constexpr int kMaxFramesPerTile = 16;  // 4 warps x 4 frames
constexpr int kRawPitch = 264;  // elements per raw 256-element row in shared memory (+8: frames shift 8 banks)
constexpr int kXchStride = 145; // complex slots per frame in the exchange buffer: 16 rows of 9 (+1 pad) + 1
constexpr int kURow = 272;      // inverse: floats per U row (256 + 16)

template <typename R> struct cx { R re, im; };
template <typename R> MDCT_HD cx<R> operator+(cx<R> a, cx<R> b) { return {a.re + b.re, a.im + b.im}; }
template <typename R> MDCT_HD cx<R> operator-(cx<R> a, cx<R> b) { return {a.re - b.re, a.im - b.im}; }
template <typename R> MDCT_HD cx<R> cmul(cx<R> a, cx<R> b) {
  return {a.re * b.re - a.im * b.im, a.re * b.im + a.im * b.re};
}
template <typename R> MDCT_HD cx<R> cmulc(cx<R> a, R br, R bi) {
  return {a.re * br - a.im * bi, a.re * bi + a.im * br};
}
template <typename R> MDCT_HD cx<R> mul_mi(cx<R> a) { return {a.im, -a.re}; }   // a * (-i)

// exact fp32 product, never contracted into an FMA: the reference rounds w*x to fp32
// (models/mdct.py:410 multiplies fp32 frames by the fp32 window) before anything else happens.
MDCT_HD float fmul32(float a, float b) {
#ifdef __CUDA_ARCH__
  return __fmul_rn(a, b);
#else
  volatile float p = a * b;
  return p;
#endif
}

// ---- radix-4 / 8 / 16 forward DFTs (kernel exp(-2 pi i nk/N)), in place -------------------------
template <typename R> MDCT_HD void dft4(cx<R>& a0, cx<R>& a1, cx<R>& a2, cx<R>& a3) {
  cx<R> t0 = a0 + a2, t1 = a0 - a2, t2 = a1 + a3, t3 = mul_mi(a1 - a3);
  a0 = t0 + t2; a1 = t1 + t3; a2 = t0 - t2; a3 = t1 - t3;
}

// multiply by W16^m (compile-time m)
template <int M, typename R> MDCT_HD cx<R> mul_w16(cx<R> a) {
  if (M == 0) return a;
  if (M == 4) return mul_mi(a);
  if (M == 2) { const R c = (R)0.70710678118654752440; return {(a.re + a.im) * c, (a.im - a.re) * c}; }
  if (M == 6) { const R c = (R)0.70710678118654752440; return {(a.im - a.re) * c, -(a.re + a.im) * c}; }
  return cmulc(a, w16_re<R>(M), w16_im<R>(M));
}

// x[16] in natural order -> x[4*k1 + k2] = X[k1 + 4*k2]
template <typename R> MDCT_HD void dft16(cx<R>* x) {
  dft4(x[0], x[4], x[8], x[12]);
  dft4(x[1], x[5], x[9], x[13]);
  dft4(x[2], x[6], x[10], x[14]);
  dft4(x[3], x[7], x[11], x[15]);
  // x[n2 + 4*k1] *= W16^(n2*k1)
  x[5] = mul_w16<1>(x[5]);  x[9] = mul_w16<2>(x[9]);   x[13] = mul_w16<3>(x[13]);
  x[6] = mul_w16<2>(x[6]);  x[10] = mul_w16<4>(x[10]); x[14] = mul_w16<6>(x[14]);
  x[7] = mul_w16<3>(x[7]);  x[11] = mul_w16<6>(x[11]); x[15] = mul_w16<9>(x[15]);
  dft4(x[0], x[1], x[2], x[3]);
  dft4(x[4], x[5], x[6], x[7]);
  dft4(x[8], x[9], x[10], x[11]);
  dft4(x[12], x[13], x[14], x[15]);
}
MDCT_HD constexpr int dft16_k(int p) { return (p >> 2) + 4 * (p & 3); }

// x[8] in natural order -> x[4*k1 + k2] = X[k1 + 2*k2]
template <typename R> MDCT_HD void dft8(cx<R>* x) {
  for (int n2 = 0; n2 < 4; ++n2) { cx<R> s = x[n2] + x[n2 + 4], d = x[n2] - x[n2 + 4]; x[n2] = s; x[n2 + 4] = d; }
  x[5] = mul_w16<2>(x[5]); x[6] = mul_w16<4>(x[6]); x[7] = mul_w16<6>(x[7]);   // W8^n2 = W16^(2 n2)
  dft4(x[0], x[1], x[2], x[3]);
  dft4(x[4], x[5], x[6], x[7]);
}
MDCT_HD constexpr int dft8_k(int p) { return (p >> 2) + 2 * (p & 3); }

// ---- per-thread tables (registers) ---------------------------------------------------------------
// T[k1]  = exp(-i pi ((j+1/8) + (k1+1/8) + 4 j k1) / 256) * scale     (pass-1 thread j)
// wE/wO  = window at the even / odd sample positions the thread's 16 points read (forward only)
template <typename R> struct ThreadTab {
  cx<R> T[16];
  MDCT_HD cx<R> operator()(int k1) const { return T[k1]; }
};
struct WinTab {
  float wE[16], wO[16];
  MDCT_HD float e(int r) const { return wE[r]; }
  MDCT_HD float o(int r) const { return wO[r]; }
};
// The same tables read on the fly from shared memory (kernels; keeps 64 / 32 registers free per thread):
// sT[k1*8 + j] (complex, pre-scaled) and sW[r*8 + j] = (wE, wO) -- the 8 lanes of a frame read consecutive slots.
template <typename R> struct SmemT {
  const cx<R>* base;   // + j
  MDCT_HD cx<R> operator()(int k1) const { return base[k1 * 8]; }
};
template <typename R> struct SmemW {
  const R* base;       // + 2*j; already converted to the core type
  MDCT_HD R e(int r) const { return base[r * 16]; }
  MDCT_HD R o(int r) const { return base[r * 16 + 1]; }
};

// Plan table layout in global memory (built once on the host in fp64, see capi.cu):
//   tabT_f32 [8][16][2] float, tabT_f64 [8][16][2] double, tabW [8][16][2] float (wE, wO)
template <typename R> MDCT_HD void load_T(const R* __restrict__ tabT, int j, R scale, ThreadTab<R>& t) {
#pragma unroll
  for (int k1 = 0; k1 < 16; ++k1) {
    t.T[k1].re = tabT[(j * 16 + k1) * 2 + 0] * scale;
    t.T[k1].im = tabT[(j * 16 + k1) * 2 + 1] * scale;
  }
}
MDCT_HD void load_W(const float* __restrict__ tabW, int j, WinTab& w) {
#pragma unroll
  for (int r = 0; r < 16; ++r) {
    w.wE[r] = tabW[(j * 16 + r) * 2 + 0];
    w.wO[r] = tabW[(j * 16 + r) * 2 + 1];
  }
}

// ---- forward pass-1 input: window + TDAC fold + r-dependent pre-twiddle -------------------------
// row0 / row1: the 256 raw samples of block t / block t+1 (frame t covers padded samples
// [256 t, 256 t + 512) = block t-1 .. t in clip coordinates; the caller passes the two rows).
// Even sample E[e] = row[2e], odd sample O[o] = row[2o+1].
// NATIVE = true: products formed in R (fp32 flavour: contractable into FMAs; fp64 core on fp32 data: exact, since
// 24 + 24 mantissa bits fit a double).  NATIVE = false keeps the reference's fp32-rounded products
// (mdct.py:410), which the bit-faithful fp64 flavour needs for 1e-13 parity.
template <typename R, bool NATIVE, typename WT>
MDCT_HD void fwd_gather(const float* row0, const float* row1, int j, const WT& w, cx<R>* v) {
#pragma unroll
  for (int r = 0; r < 16; ++r) {
    // n = j + 8r; e = 2*((n+64) mod 128), o = 2*((63-n) mod 128)+1, written wrap-free so that every
    // access is lane-base + immediate
    const int e = (r < 8) ? 2 * j + 16 * r + 128 : 2 * j + 16 * r - 128;
    const int o = (r < 8) ? 127 - 2 * j - 16 * r : 383 - 2 * j - 16 * r;
    R ue, uo;
    const auto we = w.e(r), wo = w.o(r);
    if (NATIVE) {
      const R wE = (R)we, wO = (R)wo;
      if (r < 8) {   // n < 64
        ue = -(wE * (R)row1[o]) - (wO * (R)row1[e]);
        uo = (wO * (R)row0[o]) - (wE * (R)row0[e]);
      } else {
        ue = (wE * (R)row0[e]) - (wO * (R)row0[o]);
        uo = -(wO * (R)row1[e]) - (wE * (R)row1[o]);
      }
    } else {
      if (r < 8) {
        ue = -(R)fmul32((float)we, row1[o]) - (R)fmul32((float)wo, row1[e]);
        uo = (R)fmul32((float)wo, row0[o]) - (R)fmul32((float)we, row0[e]);
      } else {
        ue = (R)fmul32((float)we, row0[e]) - (R)fmul32((float)wo, row0[o]);
        uo = -(R)fmul32((float)wo, row1[e]) - (R)fmul32((float)we, row1[o]);
      }
    }
    v[r] = (r == 0) ? cx<R>{ue, uo} : cmulc(cx<R>{ue, uo}, rho_re<R>(r), rho_im<R>(r));
  }
}

// ---- inverse pass-1 input: ue[n] = X[2n], uo[n] = X[255-2n], from the raw coefficient row ---------
template <typename R, typename S>
MDCT_HD void inv_gather(const S* row, int j, cx<R>* v) {
#pragma unroll
  for (int r = 0; r < 16; ++r) {
    const int n = j + 8 * r;
    cx<R> u{(R)row[2 * n], (R)row[255 - 2 * n]};
    v[r] = (r == 0) ? u : cmulc(u, rho_re<R>(r), rho_im<R>(r));
  }
}

// ---- pass 1 butterfly + thread twiddle + swizzled store to the exchange buffer -----------------
// Exchange layout: slot(k1, j) = 9*k1 + j.  Pass 1 (lane j, compile-time k1) and pass 2 (lane a reads rows
// k1 = a and 15-a, compile-time j) both address it as lane-base + immediate, and the 18-word row pitch
// with the 290-word frame pitch makes the 64-bit pass-2 loads bank-conflict free across a half warp.
MDCT_HD int xch_slot(int k1, int j) { return k1 * 9 + j; }

template <typename R, typename TT> MDCT_HD void pass1(cx<R>* v, const TT& t, int j, cx<R>* xch) {
  dft16(v);
#pragma unroll
  for (int p = 0; p < 16; ++p) {
    const int k1 = dft16_k(p);
    xch[xch_slot(k1, j)] = cmul(v[p], t(k1));
  }
}

// ---- pass 2: two radix-8 butterflies; y[h][k2] is bin k = k1 + 16 k2, k1 = h ? 15-a : a -----------
template <typename R> MDCT_HD void pass2(const cx<R>* xch, int a, cx<R> (*y)[8]) {
#pragma unroll
  for (int h = 0; h < 2; ++h) {
    const int k1 = h ? 15 - a : a;
    cx<R> u[8];
#pragma unroll
    for (int jj = 0; jj < 8; ++jj) u[jj] = xch[xch_slot(k1, jj)];
    dft8(u);
#pragma unroll
    for (int p = 0; p < 8; ++p) {
      const int k2 = dft8_k(p);
      y[h][k2] = (k2 == 0) ? u[p] : cmulc(u[p], pik_re<R>(k2), pik_im<R>(k2));
    }
  }
}

// Output pairing: the DCT-IV result D satisfies D[2k] = Re y_k, D[255-2k] = -Im y_k, so thread a holds
//   pair (D[c], D[c+1]) at column c = 2a + 32 k2      = ( Re y[0][k2], -Im y[1][7-k2] )
//   pair (D[c], D[c+1]) at column c = 30 - 2a + 32 k2 = ( Re y[1][k2], -Im y[0][7-k2] )
template <typename R> MDCT_HD void out_pair(const cx<R> (*y)[8], int a, int h, int k2, int& col, R& d0, R& d1) {
  if (h == 0) { col = 2 * a + 32 * k2; d0 = y[0][k2].re; d1 = -y[1][7 - k2].im; }
  else        { col = 30 - 2 * a + 32 * k2; d0 = y[1][k2].re; d1 = -y[0][7 - k2].im; }
}

// ---- inverse: TDAC unfold of the DCT-IV result U (256 values per frame) --------------------------
// first half  (m <  256) of frame t: i <128: U[128+i]   else -U[383-i]
// second half (m >= 256) of frame t: i <128: -U[127-i]  else -U[i-128]
template <typename R> MDCT_HD R unfold_first(const R* U, int i) { return i < 128 ? U[128 + i] : -U[383 - i]; }
template <typename R> MDCT_HD R unfold_second(const R* U, int i) { return i < 128 ? -U[127 - i] : -U[i - 128]; }

// ---- fused compress / expand (Audio2MDCT.normalize / denormalize, pix2pixHD_model.py:83-137) -----
// mode 0: raw coefficients; mode 1: arcsinh(gain*X)/ln10_f32 (gain is folded into T by the caller);
// followed by the abs-norm affine  s -> (s - smin)/(smax-smin)*(hi-lo)+lo  = s*aff_a + aff_b.
struct NormParams {
  int mode;        // 0 raw_mdct, 1 arcsinh
  float gain;      // arcsinh_gain (mode 1)
  float aff_a;     // (hi-lo)/(smax-smin)
  float aff_b;     // lo - smin*aff_a
  float lo;        // norm_range[0]  (second channel: |s|*2 + lo)
};
constexpr double kLn10F32 = 2.3025851249694824;   // float32(log(10)), pix2pixHD_model.py:100,133

}  // namespace mdctk
