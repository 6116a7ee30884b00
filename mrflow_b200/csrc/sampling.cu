// Channel construction (a11 prologue) and the stand-alone sampler (a8).
//
//  * mrf_prepare_channels: structure_refinement/optimize_structure_single_scale.py:350-371 with
//    variational_dataterm_grayscale=1 -- gray = cv2.cvtColor(fp32 RGB, RGB2GRAY)*255 (fp64),
//    np.gradient, cv2.GaussianBlur(3x3, sigma=-1).  OpenCV's arithmetic is restated bit-exactly in
//    oracle/cvemul.py (rgb2gray32, gaussian_blur3); the kernels below follow that specification.
//  * mrf_warp_sample: structure_refinement/interpolate_through_H.py:7-51 (compute_derivs=False).
//
// One thread per output element, coalesced along x; the channel kernel stages its tile in shared memory, so HBM
// traffic is the compulsory read-RGB / write-channels.
#include "sampling.cuh"

namespace mrf {

__device__ __forceinline__ int reflect101(int i, int n) {
  if (i < 0) i = -i;
  if (i >= n) i = 2 * n - 2 - i;
  return i;
}

// gray, gradients and blur of one 32 x 8 tile in ONE kernel: the tile's gray values (2-pixel halo) are formed from
// RGB in shared memory -- cv2.cvtColor(fp32, COLOR_RGB2GRAY) = fma(B, 0.114f, fma(R, 0.299f, G*0.587f)), then
// .astype('float64') * 255.0 (:352) --, the two np.gradient planes (1-pixel halo) from those, and the 3x3 Gaussian
// (cv2.GaussianBlur(fp64, (3,3), sigmaX=-1): row pass (p[x-1]*.25 + p[x]*.5) + p[x+1]*.25, column pass
// r[y]*.5 + (r[y-1] + r[y+1])*.25, BORDER_REFLECT_101) from those.  The reflected blur taps of a tile at the frame
// border fall inside the tile's own gradient halo.  One launch per frame, no gray plane in HBM; the operations and
// their order are oracle/cvemul.py's, so the result is bit-identical to OpenCV's.
constexpr int kChTX = 32, kChTY = 8;
template <typename T>
__global__ void __launch_bounds__(kChTX * kChTY)
k_channels(const T* __restrict__ rgb, int h, int w, double* __restrict__ ch64, float4* __restrict__ texels) {
  __shared__ double s_gray[kChTY + 4][kChTX + 4];
  __shared__ double s_gy[kChTY + 2][kChTX + 2];
  __shared__ double s_gx[kChTY + 2][kChTX + 2];
  const int x0 = blockIdx.x * kChTX, y0 = blockIdx.y * kChTY;
  for (int k = threadIdx.x; k < (kChTY + 4) * (kChTX + 4); k += kChTX * kChTY) {
    const int r = k / (kChTX + 4), c = k - r * (kChTX + 4);
    const int Y = y0 - 2 + r, X = x0 - 2 + c;
    if (Y >= 0 && Y < h && X >= 0 && X < w) {
      const long long i = (long long)Y * w + X;
      // I.astype('float32') when T = double; cv2.cvtColor(fp32, COLOR_RGB2GRAY); .astype('float64') * 255.0
      const float rr = (float)rgb[3 * i], gg = (float)rgb[3 * i + 1], bb = (float)rgb[3 * i + 2];
      const float yv = fmaf(bb, 0.114f, fmaf(rr, 0.299f, gg * 0.587f));
      s_gray[r][c] = (double)yv * 255.0;
    }
  }
  __syncthreads();
  for (int k = threadIdx.x; k < (kChTY + 2) * (kChTX + 2); k += kChTX * kChTY) {
    const int r = k / (kChTX + 2), c = k - r * (kChTX + 2);
    const int Y = y0 - 1 + r, X = x0 - 1 + c;
    if (Y >= 0 && Y < h && X >= 0 && X < w) {
      const int gr = r + 1, gc = c + 1;                                  // the same pixel in s_gray
      // np.gradient, unit spacing, edge_order = 1
      double dy, dx;
      if (Y == 0) dy = s_gray[gr + 1][gc] - s_gray[gr][gc];
      else if (Y == h - 1) dy = s_gray[gr][gc] - s_gray[gr - 1][gc];
      else dy = (s_gray[gr + 1][gc] - s_gray[gr - 1][gc]) / 2.0;
      if (X == 0) dx = s_gray[gr][gc + 1] - s_gray[gr][gc];
      else if (X == w - 1) dx = s_gray[gr][gc] - s_gray[gr][gc - 1];
      else dx = (s_gray[gr][gc + 1] - s_gray[gr][gc - 1]) / 2.0;
      s_gy[r][c] = dy; s_gx[r][c] = dx;
    }
  }
  __syncthreads();
  const int x = x0 + (threadIdx.x & 31), y = y0 + (threadIdx.x >> 5);
  if (x >= w || y >= h) return;
  // cv2.GaussianBlur(fp64, (3,3), sigmaX=-1): [1/4,1/2,1/4] separable, BORDER_REFLECT_101
  const int ca = reflect101(x - 1, w) - (x0 - 1), cb = x - (x0 - 1), cc = reflect101(x + 1, w) - (x0 - 1);
  double ry[3], rx[3];
#pragma unroll
  for (int j = 0; j < 3; ++j) {
    const int rj = reflect101(y + j - 1, h) - (y0 - 1);
    ry[j] = (s_gy[rj][ca] * 0.25 + s_gy[rj][cb] * 0.5) + s_gy[rj][cc] * 0.25;
    rx[j] = (s_gx[rj][ca] * 0.25 + s_gx[rj][cb] * 0.5) + s_gx[rj][cc] * 0.25;
  }
  const double gy = ry[1] * 0.5 + (ry[0] + ry[2]) * 0.25;
  const double gx = rx[1] * 0.5 + (rx[0] + rx[2]) * 0.25;
  const long long i = (long long)y * w + x;
  const double g = s_gray[(threadIdx.x >> 5) + 2][(threadIdx.x & 31) + 2];
  if (ch64) { ch64[3 * i] = g; ch64[3 * i + 1] = gy; ch64[3 * i + 2] = gx; }
  if (texels) texels[i] = make_float4((float)g, (float)gy, (float)gx, 0.0f);
}

// ------------------------------------------------------------------------------------- a8
struct FetchGlobalC {
  const float* img; int w; int C; int c;
  __device__ __forceinline__ F3 operator()(int yy, int xx) const {
    F3 r; r.x = img[((long long)yy * w + xx) * C + c]; r.y = 0.f; r.z = 0.f; return r;
  }
};
struct FetchGlobal3 {
  const float* img; int w; int C;
  __device__ __forceinline__ F3 operator()(int yy, int xx) const {
    const float* p = img + ((long long)yy * w + xx) * C;
    F3 r; r.x = p[0]; r.y = p[1]; r.z = p[2]; return r;
  }
};

template <int INTERP>
__global__ void __launch_bounds__(256)
k_warp_sample(const float* __restrict__ img, int img_h, int img_w, int C, bool useH, Mat3 H,
              const double* __restrict__ xn, const double* __restrict__ yn, size_t n,
              float* __restrict__ J, double* __restrict__ mask) {
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
    double X = xn[i], Y = yn[i];
    bool inside;
    if (useH) through_H(H, X, Y, img_h, img_w, X, Y, inside);
    else inside = (X >= 0.0) && (X <= (double)(img_w - 1)) && (Y >= 0.0) && (Y <= (double)(img_h - 1));
    if (mask) mask[i] = inside ? 1.0 : 0.0;
    if (!J) continue;
    const float xf = (float)X, yf = (float)Y;
    if (C >= 3) {
      FetchGlobal3 f{img, img_w, C};
      const F3 v = (INTERP == MRF_INTERP_CUBIC) ? sample_cubic(f, img_h, img_w, xf, yf)
                                                 : sample_linear(f, img_h, img_w, xf, yf);
      J[i * C] = v.x; J[i * C + 1] = v.y; J[i * C + 2] = v.z;
      if (C == 4) {
        FetchGlobalC f1{img, img_w, C, 3};
        const F3 u = (INTERP == MRF_INTERP_CUBIC) ? sample_cubic(f1, img_h, img_w, xf, yf)
                                                   : sample_linear(f1, img_h, img_w, xf, yf);
        J[i * C + 3] = u.x;
      }
    } else {
      for (int c = 0; c < C; ++c) {
        FetchGlobalC f1{img, img_w, C, c};
        const F3 u = (INTERP == MRF_INTERP_CUBIC) ? sample_cubic(f1, img_h, img_w, xf, yf)
                                                   : sample_linear(f1, img_h, img_w, xf, yf);
        J[i * C + c] = u.x;
      }
    }
  }
}

static inline unsigned grid_for(long long n, int threads = 256) {
  long long b = (n + threads - 1) / threads;
  const long long cap = (long long)kSMs * 16;
  if (b > cap) b = cap;
  if (b < 1) b = 1;
  return (unsigned)b;
}

}  // namespace mrf

using namespace mrf;

extern "C" {

int mrf_prepare_channels(const void* rgb, int rgb_is_f64, int h, int w, double* ch64, float* texels,
                         double* gray_scratch, mrf_stream_t stream) {
  MRF_CHECK_ARG(rgb && h >= 2 && w >= 2 && (ch64 || texels));
  (void)gray_scratch;   /* kept in the signature; the gray plane now lives in shared memory only */
  dim3 grid(cdiv(w, kChTX), cdiv(h, kChTY));
  if (rgb_is_f64) k_channels<double><<<grid, kChTX * kChTY, 0, S(stream)>>>((const double*)rgb, h, w, ch64, (float4*)texels);
  else            k_channels<float><<<grid, kChTX * kChTY, 0, S(stream)>>>((const float*)rgb, h, w, ch64, (float4*)texels);
  return check_launch("mrf_prepare_channels/channels");
}

int mrf_warp_sample(const float* img, int img_h, int img_w, int C, const double* H_host, const double* xn,
                    const double* yn, size_t n, int interp, float* J, double* mask, mrf_stream_t stream) {
  MRF_CHECK_ARG(img && xn && yn && img_h > 0 && img_w > 0 && C >= 1 && C <= 4 && (J || mask));
  MRF_CHECK_ARG(interp == MRF_INTERP_CUBIC || interp == MRF_INTERP_LINEAR);
  if (n == 0) return 0;
  Mat3 H;
  for (int i = 0; i < 9; ++i) H.m[i] = H_host ? H_host[i] : (i % 4 == 0 ? 1.0 : 0.0);
  if (interp == MRF_INTERP_CUBIC)
    k_warp_sample<MRF_INTERP_CUBIC><<<grid_for((long long)n), 256, 0, S(stream)>>>(
        img, img_h, img_w, C, H_host != nullptr, H, xn, yn, n, J, mask);
  else
    k_warp_sample<MRF_INTERP_LINEAR><<<grid_for((long long)n), 256, 0, S(stream)>>>(
        img, img_h, img_w, C, H_host != nullptr, H, xn, yn, n, J, mask);
  return check_launch("mrf_warp_sample");
}

}  // extern "C"
