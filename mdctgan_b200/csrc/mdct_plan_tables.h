// mdct_plan_tables.h -- host-side construction of the per-thread tables of the DCT-IV core
// (plain C++, fp64 math, no CUDA).  Used by capi.cu (uploaded into the plan) and by tests/emu/.
#pragma once
#include <cmath>
#include <vector>

namespace mdctk {

struct PlanTablesHost {
  std::vector<float> T32;    // [8][16][2]
  std::vector<double> T64;   // [8][16][2]
  std::vector<float> W;      // [8][16][2]  (wE, wO) per (j, r)
};

// T[j][k1] = exp(-i*pi*((j+1/8) + (k1+1/8) + 4*j*k1)/256)
//   = tau_j (thread part of the pre-twiddle) * W128^(j*k1) (inter-pass twiddle) * sigma_k1 (k1 part of
//   the post-twiddle); the r / k2 parts are the compile-time constants rho_r / pi_k2.
inline void build_plan_tables(const float* window512, PlanTablesHost& t) {
  const double pi = 3.14159265358979323846264338327950288;
  t.T32.resize(8 * 16 * 2);
  t.T64.resize(8 * 16 * 2);
  t.W.resize(8 * 16 * 2);
  for (int j = 0; j < 8; ++j) {
    for (int k1 = 0; k1 < 16; ++k1) {
      const double num = (j + 0.125) + (k1 + 0.125) + 4.0 * j * k1;   // exact in fp64
      const double ang = -pi * std::fmod(num, 512.0) / 256.0;
      const double c = std::cos(ang), s = std::sin(ang);
      t.T64[(j * 16 + k1) * 2 + 0] = c;
      t.T64[(j * 16 + k1) * 2 + 1] = s;
      t.T32[(j * 16 + k1) * 2 + 0] = (float)c;
      t.T32[(j * 16 + k1) * 2 + 1] = (float)s;
    }
    for (int r = 0; r < 16; ++r) {
      const int n = j + 8 * r;
      int ie, io;   // window index met by the even-position / odd-position sample of the first half
      if (n < 64) { ie = 128 + 2 * n; io = 127 - 2 * n; }
      else        { ie = 2 * n - 128; io = 383 - 2 * n; }
      t.W[(j * 16 + r) * 2 + 0] = window512[ie];
      t.W[(j * 16 + r) * 2 + 1] = window512[io];
    }
  }
}

// Synthesis window of the fp32 flavour: w_s[n] = w[n] / (w[n]^2 + w[n+N]^2).  For a symmetric analysis
// window this satisfies the TDAC conditions exactly (alias terms cancel because w_s/w is symmetric about
// the half-window centre; the overlap terms sum to 1), so the fp32 rounding of kbdwin -- the reference's
// whole round-trip error floor, 1.4 eps*peak -- no longer reaches the output.  It differs from
// the reference's synthesis window (= w) by <= 2.4e-7 relative.  The fp64 flavour keeps w (bit-faithful).
inline void build_synthesis_window(const float* w, int n, std::vector<float>& ws) {
  ws.resize(n);
  const int h = n / 2;
  for (int m = 0; m < n; ++m) {
    const double a = w[m], b = w[(m + h) % n];
    ws[m] = (float)(a / (a * a + b * b));
  }
}

// The fold assumes w[m] == w[511-m] (true for every window the reference builds: kbdwin is
// cat(half, flip(half)), util/util.py:186).  Returns false for an asymmetric window.
inline bool window_is_symmetric(const float* w, int n) {
  for (int m = 0; m < n / 2; ++m)
    if (w[m] != w[n - 1 - m]) return false;
  return true;
}

}  // namespace mdctk
