// Element-wise helpers around the path: the robust functions of utils/robust_functions.py:26-70
// applied to device arrays (for callers such as optimize_structure_single_scale.py:289,302 that hold
// the generated closures), the square root used by their MAD auto-sigma (:23-29), and the 3x3
// erosion of the thresholded rigidity at pipeline/compute_structure.py:44.  Streaming, HBM-bound.
#include "common.cuh"
#include <math.h>
#include <string.h>

namespace mrf {

// kind = MRF_RHO_*, deriv = 0 -> rho(x), 1 -> d rho / d x  (x is the squared residual).
// p1: sigma (or eps); p2: Charbonnier exponent a (0.5 -> sqrt form) / sigma^3 for Geman-McClure'.
__global__ void __launch_bounds__(256)
k_robust_apply(const double* __restrict__ x, size_t n, int kind, int deriv, double p1, double p2,
               double* __restrict__ out) {
  const double s2 = p1 * p1;
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
    const double v = x[i];
    double r;
    switch (kind) {
      case MRF_RHO_LORENTZIAN:
        r = deriv ? 1.0 / (s2 + v) : log(1.0 + v / s2); break;                          // :52-53
      case MRF_RHO_LORENTZIAN_AUTO:
        r = deriv ? p1 / (v + 2.0 * s2) : p1 * log(1.0 + (0.5 * v) / s2); break;       // :58-59
      case MRF_RHO_CHARBONNIER:
        if (p2 == 0.5) r = deriv ? 1.0 / (2.0 * sqrt(v + s2)) : sqrt(v + s2);           // :33-34
        else           r = deriv ? p2 * pow(v + s2, p2 - 1.0) : pow(v + s2, p2);        // :41-42
        break;
      case MRF_RHO_GEMAN_MCCLURE: {
        const double d = v + s2;
        r = deriv ? p2 / (d * d) : (p1 * v) / d; break;                                 // :64-65
      }
      case MRF_RHO_QUADRATIC:
        r = deriv ? 1.0 : v; break;                                                     // :46-47
      default:
        r = sqrt(v); break;                                                             // MRF_RHO_NONE: sqrt(x) for _mad
    }
    out[i] = r;
  }
}

// cv2.erode(u8, ones(3,3)) with OpenCV's default border (+inf: outside pixels never lower the min)
__global__ void __launch_bounds__(256)
k_erode3(const uint8_t* __restrict__ src, int h, int w, uint8_t* __restrict__ dst) {
  const int x = blockIdx.x * 32 + (threadIdx.x & 31);
  const int y = blockIdx.y * 8 + (threadIdx.x >> 5);
  if (x >= w || y >= h) return;
  uint8_t m = 255;
#pragma unroll
  for (int dy = -1; dy <= 1; ++dy) {
    const int yy = y + dy;
    if (yy < 0 || yy >= h) continue;
#pragma unroll
    for (int dx = -1; dx <= 1; ++dx) {
      const int xx = x + dx;
      if (xx < 0 || xx >= w) continue;
      const uint8_t v = src[(long long)yy * w + xx];
      m = v < m ? v : m;
    }
  }
  dst[(long long)y * w + x] = m;
}

// pipeline/optimize_variational_refinement.py:20-42: cv2.medianBlur(uint8 0/1 map, 3) (BORDER_REPLICATE:
// the median of nine 0/1 values is 1 iff at least five are set), then the occlusion-aware merge of the
// two normed structures.  bool * fp64 products are written as multiplications by 0.0 / 1.0 so that
// inf / NaN propagate exactly as in numpy.
__device__ __forceinline__ bool median3_bit(const uint8_t* __restrict__ m, int h, int w, int x, int y) {
  int cnt = 0;
#pragma unroll
  for (int dy = -1; dy <= 1; ++dy) {
    const int yy = max(0, min(h - 1, y + dy));
#pragma unroll
    for (int dx = -1; dx <= 1; ++dx) cnt += m[(long long)yy * w + max(0, min(w - 1, x + dx))] ? 1 : 0;
  }
  return cnt >= 5;
}

__global__ void __launch_bounds__(256)
k_merge_structures(const double* __restrict__ S0, const double* __restrict__ S1, const uint8_t* __restrict__ occ0,
                   const uint8_t* __restrict__ occ1, int h, int w, double mu0, double mu1, int reasoning,
                   double* __restrict__ out, uint8_t* __restrict__ occ0_blur, uint8_t* __restrict__ occ1_blur) {
  const int x = blockIdx.x * 32 + (threadIdx.x & 31);
  const int y = blockIdx.y * 8 + (threadIdx.x >> 5);
  if (x >= w || y >= h) return;
  const long long i = (long long)y * w + x;
  const bool ob = median3_bit(occ0, h, w, x, y), of = median3_bit(occ1, h, w, x, y);   // :23-24
  if (occ0_blur) occ0_blur[i] = ob;
  if (occ1_blur) occ1_blur[i] = of;
  bool fwd_only = !ob && of, bwd_only = !of && ob;                                     // :28-29
  bool none = !fwd_only && !bwd_only;                                                  // :30
  if (reasoning == 0) { bwd_only = true; none = false; }                              // :32-37
  const double s0 = S0[i] * mu1 / mu0, s1 = S1[i];                                     // :40-41
  const double a = (fwd_only ? 1.0 : 0.0) * s0, b = (bwd_only ? 1.0 : 0.0) * s1;
  const double c = ((none ? 1.0 : 0.0) * (s0 + s1)) / 2.0;
  out[i] = (a + b) + c;                                                                // :42
}

}  // namespace mrf

using namespace mrf;

extern "C" {

int mrf_robust_apply_f64(const double* x, size_t n, int kind, int deriv, double p1, double p2, double* out,
                         mrf_stream_t stream) {
  MRF_CHECK_ARG(x && out && kind >= MRF_RHO_LORENTZIAN && kind <= MRF_RHO_NONE);
  if (n == 0) return 0;
  long long b = (long long)((n + 255) / 256);
  if (b > kSMs * 16) b = kSMs * 16;
  k_robust_apply<<<(unsigned)b, 256, 0, S(stream)>>>(x, n, kind, deriv, p1, p2, out);
  return check_launch("mrf_robust_apply_f64");
}

int mrf_peer_alloc(size_t bytes, void** ptr_out, unsigned char* handle64_out) {
  MRF_CHECK_ARG(bytes > 0 && ptr_out && handle64_out);
  static_assert(sizeof(cudaIpcMemHandle_t) == 64, "IPC handle size");
  void* p = nullptr;
  cudaError_t e = cudaMalloc(&p, bytes);
  if (e != cudaSuccess) return fail((int)e, "mrf_peer_alloc: %s", cudaGetErrorString(e));
  cudaIpcMemHandle_t h;
  e = cudaIpcGetMemHandle(&h, p);
  if (e != cudaSuccess) { cudaFree(p); return fail((int)e, "mrf_peer_alloc/handle: %s", cudaGetErrorString(e)); }
  memcpy(handle64_out, &h, 64);
  *ptr_out = p;
  return 0;
}

// Rows [y0, y0 + rows) of every label plane of a device [L][frame_h][w] volume into the same rows of a host
// volume of the same shape (pinned memory): one pitched DMA (cudaMemcpy2DAsync), so a row chunk can travel
// over PCIe while the next chunk is being computed.
int mrf_download_rows(void* dst_host, const void* src_dev, size_t elem_bytes, int n_planes, int frame_h, int w, int y0,
                      int rows, mrf_stream_t stream) {
  MRF_CHECK_ARG(dst_host && src_dev && elem_bytes > 0 && n_planes > 0 && frame_h > 0 && w > 0 && y0 >= 0 && rows > 0 &&
                y0 + rows <= frame_h);
  const size_t pitch = (size_t)frame_h * w * elem_bytes, off = (size_t)y0 * w * elem_bytes, width = (size_t)rows * w * elem_bytes;
  cudaError_t e = cudaMemcpy2DAsync((char*)dst_host + off, pitch, (const char*)src_dev + off, pitch, width, (size_t)n_planes,
                                    cudaMemcpyDeviceToHost, S(stream));
  if (e != cudaSuccess) return fail((int)e, "mrf_download_rows: %s", cudaGetErrorString(e));
  return 0;
}

int mrf_peer_open(const unsigned char* handle64, void** ptr_out) {
  MRF_CHECK_ARG(handle64 && ptr_out);
  cudaIpcMemHandle_t h;
  memcpy(&h, handle64, 64);
  void* p = nullptr;
  cudaError_t e = cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess);
  if (e != cudaSuccess) return fail((int)e, "mrf_peer_open: %s", cudaGetErrorString(e));
  *ptr_out = p;
  return 0;
}

int mrf_peer_close(void* ptr) {
  MRF_CHECK_ARG(ptr);
  cudaError_t e = cudaIpcCloseMemHandle(ptr);
  if (e != cudaSuccess) return fail((int)e, "mrf_peer_close: %s", cudaGetErrorString(e));
  return 0;
}

int mrf_peer_free(void* ptr) {
  MRF_CHECK_ARG(ptr);
  cudaError_t e = cudaFree(ptr);
  if (e != cudaSuccess) return fail((int)e, "mrf_peer_free: %s", cudaGetErrorString(e));
  return 0;
}

int mrf_merge_structures(const double* S0, const double* S1, const uint8_t* occ_bwd, const uint8_t* occ_fwd, int h,
                         int w, double mu0, double mu1, int occlusion_reasoning, double* out, uint8_t* occ_bwd_blur,
                         uint8_t* occ_fwd_blur, mrf_stream_t stream) {
  MRF_CHECK_ARG(S0 && S1 && occ_bwd && occ_fwd && out && h > 0 && w > 0);
  MRF_CHECK_ARG(occ_bwd_blur != occ_bwd && occ_fwd_blur != occ_fwd);
  dim3 grid(cdiv(w, 32), cdiv(h, 8));
  k_merge_structures<<<grid, 256, 0, S(stream)>>>(S0, S1, occ_bwd, occ_fwd, h, w, mu0, mu1, occlusion_reasoning, out,
                                                  occ_bwd_blur, occ_fwd_blur);
  return check_launch("mrf_merge_structures");
}

int mrf_erode3x3_u8(const uint8_t* src, int h, int w, uint8_t* dst, mrf_stream_t stream) {
  MRF_CHECK_ARG(src && dst && h > 0 && w > 0 && src != dst);
  dim3 grid(cdiv(w, 32), cdiv(h, 8));
  k_erode3<<<grid, 256, 0, S(stream)>>>(src, h, w, dst);
  return check_launch("mrf_erode3x3_u8");
}

}  // extern "C"
