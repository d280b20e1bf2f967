// Shared device/host helpers for the NLT B200 kernels.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdarg.h>
#include "nlt_b200.h"

namespace nlt {

// ---- thread-local error string (nlt_last_error) -----------------------------
extern thread_local char g_err[512];
int set_err(int code, const char* fmt, ...);
#define NLT_CHECK_ARG(cond, ...) \
  do { if (!(cond)) return nlt::set_err(NLT_ERR_INVALID, __VA_ARGS__); } while (0)
extern unsigned long long g_launches;   // kernels launched by this library (diagnostic only)
#define NLT_CUDA_LAUNCH_CHECK(what) \
  do { __atomic_add_fetch(&nlt::g_launches, 1ull, __ATOMIC_RELAXED); cudaError_t e_ = cudaGetLastError(); \
       if (e_ != cudaSuccess) return nlt::set_err(NLT_ERR_CUDA, "%s: %s", what, cudaGetErrorString(e_)); } while (0)

// ---- division by an invariant (Granlund-Montgomery), n < 2^31 ----------------
struct FastDiv {
  uint32_t d, m, l;
};
inline FastDiv make_fastdiv(uint32_t d) {
  FastDiv f;
  f.d = d ? d : 1;
  uint32_t l = 0;
  while ((1ull << l) < f.d) ++l;
  f.l = l;
  f.m = (uint32_t)((((1ull << 32) * ((1ull << l) - f.d)) / f.d) + 1);
  return f;
}
__device__ __forceinline__ uint32_t fdiv(uint32_t n, const FastDiv& f) {
  return (__umulhi(n, f.m) + n) >> f.l;
}

// ---- activation ---------------------------------------------------------------
// forward (nlt/networks/elements.py:69-78)
__device__ __forceinline__ float act_fwd(float x, int act) {
  switch (act) {
    case NLT_ACT_RELU: return x > 0.f ? x : 0.f;
    case NLT_ACT_LEAKYRELU: return x > 0.f ? x : 0.3f * x;
    case NLT_ACT_ELU: return x > 0.f ? x : expm1f(x);
    default: return x;
  }
}
// derivative recovered from the saved OUTPUT y (no pre-activation is stored)
__device__ __forceinline__ float act_bwd_from_y(float y, int act) {
  switch (act) {
    case NLT_ACT_RELU: return y > 0.f ? 1.f : 0.f;
    case NLT_ACT_LEAKYRELU: return y > 0.f ? 1.f : 0.3f;
    case NLT_ACT_ELU: return y > 0.f ? 1.f : y + 1.f;
    default: return 1.f;
  }
}

// One axis of the (phase-decomposed) gather map.
//   output coordinate  o = o0 + os * t      t in [0, nt)
//   tap index          d = d0 + ds * u      u in [0, nu)
//   input coordinate   i = t * it + u * iu + i0   (valid iff 0 <= i < n_in)
struct AxisMap {
  int o0, os, nt;
  int it, iu, i0;
  int d0, ds, nu;
  int n_in;
};

struct Seg {
  const float* ptr;
  const float* sub;
  int C;      // channels of this source
  int coff;   // channel offset inside the virtual concat (weight c index)
  int bcast;  // batch-broadcast
  int vec;    // 1: C % 4 == 0 and pointers 16B aligned -> float4 path
};

// Kernel-side description of one phase of a generalised convolution.
struct GConvK {
  AxisMap ay, ax;
  int N, Hin, Win, Hout, Wout, kw;
  int nseg;
  Seg seg[NLT_MAX_SEG];
  int Cout;
  const float* w;
  long long wt, wc, wn;
  uint32_t M;  // lattice pixels in this phase = N * ay.nt * ax.nt
  FastDiv div_x, div_yx;
};

__device__ __forceinline__ void decode_pixel(const GConvK& g, uint32_t m, int& n, int& ty, int& tx) {
  uint32_t q = fdiv(m, g.div_yx);
  uint32_t r = m - q * g.div_yx.d;
  uint32_t y = fdiv(r, g.div_x);
  n = (int)q;
  ty = (int)y;
  tx = (int)(r - y * g.div_x.d);
}

// build the per-phase kernel descriptors of a public descriptor (host)
int build_phases(const nlt_gconv_desc* d, GConvK* out, int* nphase);

}  // namespace nlt
