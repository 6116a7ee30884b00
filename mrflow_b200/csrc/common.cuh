// Shared helpers for libmrflow_b200 (sm_100a only).
//
// Build note: every translation unit is compiled with -fmad=false, so `a*b+c` written with plain
// operators is NOT contracted into an FMA -- the pointwise kernels rely on that to reproduce the
// reference's (numpy / OpenCV, no FMA) rounding bit for bit.  Where an FMA is wanted it is
// written explicitly as fma()/fmaf().
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <atomic>
#include "../../include/mrflow_b200.h"

namespace mrf {

extern thread_local char g_err[512];
extern std::atomic<long long> g_launches;

int fail(int code, const char* fmt, ...);
int check_launch(const char* what);

inline cudaStream_t S(mrf_stream_t s) { return (cudaStream_t)s; }
inline unsigned cdiv(long long a, long long b) { return (unsigned)((a + b - 1) / b); }

struct Mat3 { double m[9]; };
struct Pt2  { double x, y; };

inline Mat3 mat3(const double* p) { Mat3 r; for (int i = 0; i < 9; ++i) r.m[i] = p[i]; return r; }
inline Pt2  pt2(const double* p)  { Pt2 r; r.x = p[0]; r.y = p[1]; return r; }

constexpr int kSMs = 148;   // B200

// ---------------------------------------------------------------------------------------------
// cv2.perspectiveTransform on one fp32 point with an fp64 matrix (OpenCV: fp64 arithmetic,
// w = 1/w then multiply, one rounding to fp32, |w| <= FLT_EPSILON -> 0).
// Reference call sites: utils/flow_homography.py:15,36.
__device__ __forceinline__ void persp_f32(float xf, float yf, const Mat3& M, float& X, float& Y) {
  const double x = (double)xf, y = (double)yf;
  double w = (x * M.m[6] + y * M.m[7]) + M.m[8];
  if (fabs(w) > 1.1920928955078125e-07) {
    w = 1.0 / w;
    X = (float)(((x * M.m[0] + y * M.m[1]) + M.m[2]) * w);
    Y = (float)(((x * M.m[3] + y * M.m[4]) + M.m[5]) * w);
  } else {
    X = 0.f; Y = 0.f;
  }
}

// ---------------------------------------------------------------------------------------------
// cv2.remap coordinate quantisation: cvRound(coord*32) with x86 "integer indefinite" for NaN /
// out-of-range, integer part saturated to int16, 5 fraction bits.
__device__ __forceinline__ void cv_quantise(float c, int& ipart, int& frac) {
  const float v = c * 32.0f;
  const int s = (fabsf(v) < 2147483648.0f) ? __float2int_rn(v) : (int)0x80000000;
  ipart = max(-32768, min(32767, s >> 5));
  frac = s & 31;
}

// OpenCV interpolateCubic(t = frac/32), A = -0.75, plain fp32 ops in OpenCV's order.
__device__ __forceinline__ void cubic_weights(int frac, float w[4]) {
  const float A = -0.75f;
  const float t = (float)frac * 0.03125f;
  const float t1 = t + 1.0f;
  w[0] = ((A * t1 - 5.0f * A) * t1 + 8.0f * A) * t1 - 4.0f * A;
  w[1] = ((A + 2.0f) * t - (A + 3.0f)) * t * t + 1.0f;
  const float u = 1.0f - t;
  w[2] = ((A + 2.0f) * u - (A + 3.0f)) * u * u + 1.0f;
  w[3] = 1.0f - w[0] - w[1] - w[2];
}

__device__ __forceinline__ void linear_weights(int frac, float w[2]) {
  const float t = (float)frac * 0.03125f;
  w[0] = 1.0f - t;
  w[1] = t;
}

// Exact median of N (odd) values by forgetful selection: keep N/2+2 candidates, repeatedly move their minimum
// and maximum to the ends (neither can be the median), replace the minimum by the next unseen value and drop
// the maximum; three values remain, the middle one is the median.  ~1.5*len compare-exchanges per round
// instead of the N^2/4 of a partial selection sort.  fetch(k) returns the k-th value (k is a compile-time
// constant after unrolling, so the candidates stay in registers).
__device__ __forceinline__ void sort2(double& a, double& b) { const double lo = fmin(a, b), hi = fmax(a, b); a = lo; b = hi; }
template <int N, class Fetch>
__device__ __forceinline__ double median_forgetful(Fetch fetch) {
  static_assert(N % 2 == 1 && N >= 3, "odd window");
  constexpr int R = N / 2 + 2;
  double v[R];
#pragma unroll
  for (int k = 0; k < R; ++k) v[k] = fetch(k);
#pragma unroll
  for (int len = R; len >= 3; --len) {
#pragma unroll
    for (int i = 0; i < len / 2; ++i) sort2(v[i], v[len - 1 - i]);
#pragma unroll
    for (int i = 1; i < (len + 1) / 2; ++i) sort2(v[0], v[i]);             // minimum -> v[0]
#pragma unroll
    for (int i = len / 2; i < len - 1; ++i) sort2(v[i], v[len - 1]);       // maximum -> v[len-1]
    if (len > 3) v[0] = fetch(N - (len - 3));                              // next unseen value replaces the minimum
  }
  return v[1];
}

}  // namespace mrf

#define MRF_CHECK_ARG(cond) do { if (!(cond)) return mrf::fail(MRF_E_BADARG, "%s: bad argument: %s", __func__, #cond); } while (0)
