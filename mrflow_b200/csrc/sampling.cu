// Channel construction (a11 prologue) and the stand-alone sampler (a8).
//
//  * mrf_prepare_channels: structure_refinement/optimize_structure_single_scale.py:350-371 with
//    variational_dataterm_grayscale=1 -- gray = cv2.cvtColor(fp32 RGB, RGB2GRAY)*255 (fp64),
//    np.gradient, cv2.GaussianBlur(3x3, sigma=-1).  OpenCV's arithmetic is restated bit-exactly in
//    oracle/cvemul.py (rgb2gray32, gaussian_blur3); the kernels below follow that specification.
//  * mrf_warp_sample: structure_refinement/interpolate_through_H.py:7-51 (compute_derivs=False).
//
// Both are streaming kernels (one thread per output element, coalesced along x); the gray plane
// and the gradient stencil are served from L1/L2, so HBM traffic is the compulsory
// read-RGB / write-channels.
#include "sampling.cuh"

namespace mrf {

// cv2.cvtColor(fp32, COLOR_RGB2GRAY): fma(B, 0.114f, fma(R, 0.299f, G*0.587f)); then
// .astype('float64') * 255.0  (optimize_structure_single_scale.py:352)
template <typename T>
__global__ void __launch_bounds__(256)
k_gray(const T* __restrict__ rgb, long long n, double* __restrict__ gray) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n;
       i += (long long)gridDim.x * blockDim.x) {
    // I.astype('float32') when T = double
    const float r = (float)rgb[3 * i], g = (float)rgb[3 * i + 1], b = (float)rgb[3 * i + 2];
    const float y = fmaf(b, 0.114f, fmaf(r, 0.299f, g * 0.587f));
    gray[i] = (double)y * 255.0;
  }
}

__device__ __forceinline__ int reflect101(int i, int n) {
  if (i < 0) i = -i;
  if (i >= n) i = 2 * n - 2 - i;
  return i;
}

// np.gradient along one axis, unit spacing, edge_order=1.
__device__ __forceinline__ double grad_y(const double* __restrict__ g, int h, int w, int y, int x) {
  if (y == 0) return g[(long long)w + x] - g[x];
  if (y == h - 1) return g[(long long)(h - 1) * w + x] - g[(long long)(h - 2) * w + x];
  return (g[(long long)(y + 1) * w + x] - g[(long long)(y - 1) * w + x]) / 2.0;
}
__device__ __forceinline__ double grad_x(const double* __restrict__ g, int h, int w, int y, int x) {
  const double* row = g + (long long)y * w;
  if (x == 0) return row[1] - row[0];
  if (x == w - 1) return row[w - 1] - row[w - 2];
  return (row[x + 1] - row[x - 1]) / 2.0;
}

// cv2.GaussianBlur(fp64, (3,3), sigmaX=-1): separable [1/4,1/2,1/4], BORDER_REFLECT_101; row pass
// (p[x-1]*.25 + p[x]*.5) + p[x+1]*.25, column pass r[y]*.5 + (r[y-1] + r[y+1])*.25.
template <bool DY>
__device__ __forceinline__ double blur3_of_grad(const double* __restrict__ g, int h, int w, int y, int x) {
  double r[3];
#pragma unroll
  for (int j = 0; j < 3; ++j) {
    const int yy = reflect101(y + j - 1, h);
    const int xa = reflect101(x - 1, w), xc = reflect101(x + 1, w);
    const double a = DY ? grad_y(g, h, w, yy, xa) : grad_x(g, h, w, yy, xa);
    const double b = DY ? grad_y(g, h, w, yy, x) : grad_x(g, h, w, yy, x);
    const double c = DY ? grad_y(g, h, w, yy, xc) : grad_x(g, h, w, yy, xc);
    r[j] = (a * 0.25 + b * 0.5) + c * 0.25;
  }
  return r[1] * 0.5 + (r[0] + r[2]) * 0.25;
}

__global__ void __launch_bounds__(256)
k_channels(const double* __restrict__ gray, int h, int w, double* __restrict__ ch64,
           float4* __restrict__ texels) {
  const int x = blockIdx.x * 32 + (threadIdx.x & 31);
  const int y = blockIdx.y * 8 + (threadIdx.x >> 5);
  if (x >= w || y >= h) return;
  const long long i = (long long)y * w + x;
  const double g = gray[i];
  const double gy = blur3_of_grad<true>(gray, h, w, y, x);
  const double gx = blur3_of_grad<false>(gray, h, w, y, x);
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
  MRF_CHECK_ARG(rgb && gray_scratch && h >= 2 && w >= 2 && (ch64 || texels));
  const long long n = (long long)h * w;
  if (rgb_is_f64) k_gray<double><<<grid_for(n), 256, 0, S(stream)>>>((const double*)rgb, n, gray_scratch);
  else            k_gray<float><<<grid_for(n), 256, 0, S(stream)>>>((const float*)rgb, n, gray_scratch);
  int rc = check_launch("mrf_prepare_channels/gray");
  if (rc) return rc;
  dim3 grid(cdiv(w, 32), cdiv(h, 8));
  k_channels<<<grid, 256, 0, S(stream)>>>(gray_scratch, h, w, ch64, (float4*)texels);
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
