// tests/emu/emu_mdct.cpp -- TEST INFRASTRUCTURE ONLY.
// Runs the arithmetic core of the CUDA kernels (mdctgan_b200/csrc/mdct_core.cuh: the very same
// templates, index maps, swizzles and twiddle tables) on the CPU, one emulated thread at a time,
// so that index/twiddle mistakes are caught in the GPU-less build container.  Built by
// tests/test_emu_core.py with g++ and driven through ctypes; never part of the product.
#include <cstdint>
#include <cstring>
#include <vector>

#include "../../mdctgan_b200/csrc/mdct_core.cuh"
#include "../../mdctgan_b200/csrc/mdct_plan_tables.h"

using namespace mdctk;

template <typename R> static const R* tabT(const PlanTablesHost& t);
template <> const float* tabT<float>(const PlanTablesHost& t) { return t.T32.data(); }
template <> const double* tabT<double>(const PlanTablesHost& t) { return t.T64.data(); }

// one frame through pass1 / exchange / pass2; D[256] = DCT-IV result in natural order
template <typename R, typename Gather>
static void frame_core(const PlanTablesHost& tabs, R scale, Gather gather, R* D) {
  std::vector<cx<R>> xch(kXchStride);
  for (int j = 0; j < 8; ++j) {
    ThreadTab<R> tt;
    load_T<R>(tabT<R>(tabs), j, scale, tt);
    cx<R> v[16];
    gather(j, v);
    pass1<R>(v, tt, j, xch.data());
  }
  for (int a = 0; a < 8; ++a) {
    cx<R> y[2][8];
    pass2<R>(xch.data(), a, y);
    for (int h = 0; h < 2; ++h)
      for (int k2 = 0; k2 < 8; ++k2) {
        int col; R d0, d1;
        out_pair<R>(y, a, h, k2, col, d0, d1);
        D[col] = d0; D[col + 1] = d1;
      }
  }
}

// NATIVE / ROUND32: see mdct_core.cuh fwd_gather; ROUND32 = the kernel's fp32 store of the coefficients
template <typename R, bool NATIVE, bool ROUND32>
static void fwd_impl(const float* x, int64_t T, int64_t F, const float* window, double* out) {
  PlanTablesHost tabs;
  build_plan_tables(window, tabs);
  // raw padded rows exactly as the kernel stages them: row b holds samples x[256(b-1) .. 256b)
  std::vector<float> rows((F + 1) * kRawPitch, 0.f);
  for (int64_t b = 0; b <= F; ++b)
    for (int p = 0; p < 256; ++p) {
      int64_t s = 256 * (b - 1) + p;
      rows[b * kRawPitch + p] = (s >= 0 && s < T) ? x[s] : 0.f;
    }
  std::vector<R> D(256);
  for (int64_t t = 0; t < F; ++t) {
    const float *row0 = &rows[t * kRawPitch], *row1 = &rows[(t + 1) * kRawPitch];
    frame_core<R>(tabs, (R)1, [&](int j, cx<R>* v) {
      WinTab w; load_W(tabs.W.data(), j, w);
      fwd_gather<R, NATIVE>(row0, row1, j, w, v);
    }, D.data());
    for (int k = 0; k < 256; ++k) out[t * 256 + k] = ROUND32 ? (double)(float)D[k] : (double)D[k];
  }
}

template <typename R, bool SYN, bool ROUND32>
static void inv_impl(const double* spec, int64_t F, const float* window_in, double* audio /*(F-1)*256*/) {
  PlanTablesHost tabs;
  build_plan_tables(window_in, tabs);
  std::vector<float> wsyn;
  build_synthesis_window(window_in, 512, wsyn);
  const float* window = SYN ? wsyn.data() : window_in;   // same choice as capi.cu launch_inv
  std::vector<R> U(F * kURow);
  for (int64_t t = 0; t < F; ++t) {
    frame_core<R>(tabs, (R)1, [&](int j, cx<R>* v) { inv_gather<R, double>(&spec[t * 256], j, v); }, &U[t * kURow]);
  }
  const R sc = (R)(4.0 / 512.0);
  for (int64_t q = 0; q + 1 < F; ++q)
    for (int i = 0; i < 256; ++i) {
      R a = unfold_first<R>(&U[(q + 1) * kURow], i) * (R)window[i];
      R b = unfold_second<R>(&U[q * kURow], i) * (R)window[256 + i];
      audio[q * 256 + i] = ROUND32 ? (double)(float)((a + b) * sc) : (double)((a + b) * sc);
    }
}

extern "C" {
int emu_window_symmetric(const float* w) { return window_is_symmetric(w, 512) ? 1 : 0; }
// flavour: 0 = fp32 core, 1 = fp64 core / fp64 I/O (bit-faithful), 2 = "mixed": fp64 core on fp32 I/O
void emu_mdct_fwd(const float* x, int64_t T, int64_t F, const float* window, double* out, int flavour) {
  if (flavour == 1) fwd_impl<double, false, false>(x, T, F, window, out);
  else if (flavour == 2) fwd_impl<double, true, true>(x, T, F, window, out);
  else if (flavour == 3) fwd_impl<double, false, true>(x, T, F, window, out);
  else fwd_impl<float, true, false>(x, T, F, window, out);
}
void emu_imdct(const double* spec, int64_t F, const float* window, double* audio, int flavour) {
  if (flavour == 1) inv_impl<double, false, false>(spec, F, window, audio);
  else if (flavour == 2 || flavour == 3) inv_impl<double, true, true>(spec, F, window, audio);
  else inv_impl<float, true, false>(spec, F, window, audio);
}
}
