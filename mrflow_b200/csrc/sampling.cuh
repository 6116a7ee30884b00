// cv2.remap-exact sampling of packed fp32 texels (INTER_CUBIC / INTER_LINEAR, BORDER_REPLICATE).
//
// Reference: structure_refinement/interpolation.py:76-80 (interp_lin -> cv2.remap on fp32 data and
// fp32 maps).  The arithmetic specification is oracle/cvemul.py::remap, which is pinned bit-exact
// against OpenCV 4.13: coordinates quantised to 1/32 px, fp32 weights, 2-D weight = fp32(wy*wx),
// every product and every sum rounded separately (this file is compiled with -fmad=false), and
// OpenCV's two summation orders -- row sums then rows for a footprint strictly inside the image,
// tap by tap for a footprint that touches the border.
#pragma once
#include "common.cuh"

namespace mrf {

struct F3 { float x, y, z; };

// Fetch functor contract: F3 operator()(int yy, int xx) with (yy, xx) already clamped to the image.

template <class Fetch>
__device__ __forceinline__ F3 sample_cubic(const Fetch& tex, int img_h, int img_w, float xf, float yf) {
  int ix, iy, fx, fy;
  cv_quantise(xf, ix, fx);
  cv_quantise(yf, iy, fy);
  float wx[4], wy[4];
  cubic_weights(fx, wx);
  cubic_weights(fy, wy);
  const int ix0 = ix - 1, iy0 = iy - 1;
  const bool inside = (ix0 >= 0) && (ix0 < img_w - 3) && (iy0 >= 0) && (iy0 < img_h - 3);
  F3 acc;
  if (inside) {
    // row sums left to right, then the four rows
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      F3 row;
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const float wgt = wy[j] * wx[i];
        const F3 t = tex(iy0 + j, ix0 + i);
        const float px = t.x * wgt, py = t.y * wgt, pz = t.z * wgt;
        if (i == 0) { row.x = px; row.y = py; row.z = pz; }
        else        { row.x = row.x + px; row.y = row.y + py; row.z = row.z + pz; }
      }
      if (j == 0) acc = row;
      else { acc.x = acc.x + row.x; acc.y = acc.y + row.y; acc.z = acc.z + row.z; }
    }
  } else {
    // border path: one running sum over the sixteen clamped taps
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int yy = max(0, min(img_h - 1, iy0 + j));
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const int xx = max(0, min(img_w - 1, ix0 + i));
        const float wgt = wy[j] * wx[i];
        const F3 t = tex(yy, xx);
        const float px = t.x * wgt, py = t.y * wgt, pz = t.z * wgt;
        if (i == 0 && j == 0) { acc.x = px; acc.y = py; acc.z = pz; }
        else { acc.x = acc.x + px; acc.y = acc.y + py; acc.z = acc.z + pz; }
      }
    }
  }
  return acc;
}

template <class Fetch>
__device__ __forceinline__ F3 sample_linear(const Fetch& tex, int img_h, int img_w, float xf, float yf) {
  int ix, iy, fx, fy;
  cv_quantise(xf, ix, fx);
  cv_quantise(yf, iy, fy);
  float wx[2], wy[2];
  linear_weights(fx, wx);
  linear_weights(fy, wy);
  F3 acc;
#pragma unroll
  for (int j = 0; j < 2; ++j) {
    const int yy = max(0, min(img_h - 1, iy + j));
#pragma unroll
    for (int i = 0; i < 2; ++i) {
      const int xx = max(0, min(img_w - 1, ix + i));
      const float wgt = wy[j] * wx[i];
      const F3 t = tex(yy, xx);
      const float px = t.x * wgt, py = t.y * wgt, pz = t.z * wgt;
      if (i == 0 && j == 0) { acc.x = px; acc.y = py; acc.z = pz; }
      else { acc.x = acc.x + px; acc.y = acc.y + py; acc.z = acc.z + pz; }
    }
  }
  return acc;
}

// Homography step of interpolate_through_H.py:30-40: fp64, numpy operation order
// ((xn*h0 + yn*h1) + h2) / ((xn*h6 + yn*h7) + h8); mask on the fp64 coordinates.
__device__ __forceinline__ void through_H(const Mat3& H, double xn, double yn, int img_h, int img_w,
                                          double& X, double& Y, bool& inside) {
  const double den = (xn * H.m[6] + yn * H.m[7]) + H.m[8];
  X = ((xn * H.m[0] + yn * H.m[1]) + H.m[2]) / den;
  Y = ((xn * H.m[3] + yn * H.m[4]) + H.m[5]) / den;
  inside = (X >= 0.0) && (X <= (double)(img_w - 1)) && (Y >= 0.0) && (Y <= (double)(img_h - 1));
}

}  // namespace mrf
