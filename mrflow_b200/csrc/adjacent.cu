// The two stages either side of the plane+parallax path:
//   * rigidity/occlusions.py:10-74  computeOcclusionsFromConsistency -- producer of the path's
//     `occlusions` input: two bilinear cv2.remap calls (default BORDER_CONSTANT) per direction, an
//     fp32 consistency error, threshold, and the invalid-region bookkeeping;
//   * rigidity/inference.py:140-167 get_image_weights -- the contrast-sensitive 8-neighbourhood
//     weights the MRF consumer (extern/eff_trws) takes next to the int32 unaries;
//   * utils/flow_io.py:33-48 flow_read -- the .flo payload (interleaved u, v) split into the two fp32 planes the
//     path takes, so a flow file goes host -> HBM as it lies on disk.
// Streaming kernels, one thread per pixel, coalesced along x.
#include "sampling.cuh"

namespace mrf {

// cv2.remap(single-channel fp32, INTER_LINEAR, BORDER_CONSTANT 0): oracle/cvemul.py::remap(border="constant")
__device__ __forceinline__ float sample_linear_const(const float* __restrict__ img, int h, int w, int ix, int iy,
                                                     const float wx[2], const float wy[2]) {
  float acc = 0.f;
#pragma unroll
  for (int j = 0; j < 2; ++j) {
#pragma unroll
    for (int i = 0; i < 2; ++i) {
      const int yy = iy + j, xx = ix + i;
      const bool in = (yy >= 0) && (yy < h) && (xx >= 0) && (xx < w);
      const float t = in ? img[(long long)yy * w + xx] : 0.0f;
      const float p = t * (wy[j] * wx[i]);
      acc = (i == 0 && j == 0) ? p : acc + p;
    }
  }
  return acc;
}

// one direction of getWarpedError (:29-45): error (fp32) and invalid flag
__device__ __forceinline__ void warped_error(const float* __restrict__ uf, const float* __restrict__ vf,
                                             const float* __restrict__ ub, const float* __restrict__ vb, int h, int w,
                                             int x, int y, long long i, float& err, bool& invalid) {
  const float fu = uf[i], fv = vf[i];
  const double xs = (double)x + (double)fu, ys = (double)y + (double)fv;     // int64 + fp32 -> fp64
  int ix, iy, fx, fy;
  cv_quantise((float)xs, ix, fx);
  cv_quantise((float)ys, iy, fy);
  float wx[2], wy[2];
  linear_weights(fx, wx);
  linear_weights(fy, wy);
  const float uw = sample_linear_const(ub, h, w, ix, iy, wx, wy);
  const float vw = sample_linear_const(vb, h, w, ix, iy, wx, wy);
  invalid = !((ys >= 0.0) && (ys < (double)h) && (xs >= 0.0) && (xs < (double)w));   // :41
  const float a = uw + fu, b = vw + fv;
  err = sqrtf(a * a + b * b);                                                         // :42
}

__global__ void __launch_bounds__(256)
k_flow_consistency(const float* __restrict__ uf0, const float* __restrict__ vf0, const float* __restrict__ ub0,
                   const float* __restrict__ vb0, const float* __restrict__ uf1, const float* __restrict__ vf1,
                   const float* __restrict__ ub1, const float* __restrict__ vb1, int h, int w, float thr,
                   uint8_t* __restrict__ occ_b, uint8_t* __restrict__ occ_f, uint8_t* __restrict__ occ_both) {
  const int x = blockIdx.x * 32 + (threadIdx.x & 31);
  const int y = blockIdx.y * 8 + (threadIdx.x >> 5);
  if (x >= w || y >= h) return;
  const long long i = (long long)y * w + x;
  float eb, ef; bool ib, i_f;
  warped_error(uf0, vf0, ub0, vb0, h, w, x, y, i, eb, ib);
  warped_error(uf1, vf1, ub1, vb1, h, w, x, y, i, ef, i_f);
  bool ob = (eb > thr) || ib;          // :62-63  (fp32 error against the threshold as fp32)
  bool of = (ef > thr) || i_f;
  if (i_f && !ib) ob = false;          // :66-70
  if (ib && !i_f) of = false;
  occ_b[i] = ob; occ_f[i] = of;
  if (occ_both) occ_both[i] = ob && of;
}

// ------------------------------------------------------------------------------- image weights
// the four neighbour differences of inference.py:146-153 at one pixel (mean over the channels)
__device__ __forceinline__ void neighbour_diffs(const double* __restrict__ I, int h, int w, int c, int x, int y,
                                                double d[4]) {
  const double* p = I + ((long long)y * w + x) * c;
  double s0 = 0.0, s1 = 0.0, s2 = 0.0, s3 = 0.0;
  const bool dn = y + 1 < h, rt = x + 1 < w, up = y >= 1;
  for (int k = 0; k < c; ++k) {
    const double v = p[k];
    if (dn) s0 = s0 + (p[(long long)w * c + k] - v);                  // np.diff(I, axis=0)
    if (rt) s1 = s1 + (p[c + k] - v);                                 // np.diff(I, axis=1)
    if (up && rt) s2 = s2 + (v - p[-(long long)w * c + c + k]);       // I[1:, :-1] - I[:-1, 1:]
    if (dn && rt) s3 = s3 + (v - p[(long long)w * c + c + k]);        // I[:-1, :-1] - I[1:, 1:]
  }
  const double cd = (double)c;
  d[0] = dn ? s0 / cd : 0.0;            // vertical   (diff_0)
  d[1] = rt ? s1 / cd : 0.0;            // horizontal (diff_1)
  d[2] = (up && rt) ? s2 / cd : 0.0;    // ne
  d[3] = (dn && rt) ? s3 / cd : 0.0;    // se
}

constexpr int kIwBlocks = 1024;

__global__ void __launch_bounds__(256)
k_image_weight_sums(const double* __restrict__ I, int h, int w, int c, double* __restrict__ partial) {
  __shared__ double sh[4][8];
  const long long n = (long long)h * w;
  double acc[4] = {0.0, 0.0, 0.0, 0.0};
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const int y = (int)(i / w), x = (int)(i - (long long)y * w);
    double d[4];
    neighbour_diffs(I, h, w, c, x, y, d);
#pragma unroll
    for (int k = 0; k < 4; ++k) acc[k] += d[k] * d[k];
  }
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    for (int o = 16; o > 0; o >>= 1) acc[k] += __shfl_down_sync(0xffffffffu, acc[k], o);
    if (lane == 0) sh[k][wid] = acc[k];
  }
  __syncthreads();
  if (threadIdx.x < 4) {
    double s = 0.0;
    for (int j = 0; j < 8; ++j) s += sh[threadIdx.x][j];
    partial[threadIdx.x * gridDim.x + blockIdx.x] = s;
  }
}

// beta_k = 1 / (2 * max(1e-6, mean(d_k^2)))  :155-159
__global__ void __launch_bounds__(kIwBlocks)
k_image_weight_betas(const double* __restrict__ partial, int nblocks, double n_px, double* __restrict__ beta) {
  __shared__ double sh[32];
  for (int k = 0; k < 4; ++k) {
    double v = (threadIdx.x < nblocks) ? partial[k * nblocks + threadIdx.x] : 0.0;
    for (int o = 16; o > 0; o >>= 1) v += __shfl_down_sync(0xffffffffu, v, o);
    if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = v;
    __syncthreads();
    if (threadIdx.x < 32) {
      v = (threadIdx.x < (blockDim.x >> 5)) ? sh[threadIdx.x] : 0.0;
      for (int o = 16; o > 0; o >>= 1) v += __shfl_down_sync(0xffffffffu, v, o);
      if (threadIdx.x == 0) beta[k] = 1.0 / (2.0 * fmax(1e-6, v / n_px));
    }
    __syncthreads();
  }
}

// w = exp(-d^2 * beta).astype(float32)  :161-165 ; output order horizontal, vertical, ne, se (:167)
__global__ void __launch_bounds__(256)
k_image_weights(const double* __restrict__ I, int h, int w, int c, const double* __restrict__ beta,
                float* __restrict__ w_h, float* __restrict__ w_v, float* __restrict__ w_ne, float* __restrict__ w_se) {
  const int x = blockIdx.x * 32 + (threadIdx.x & 31);
  const int y = blockIdx.y * 8 + (threadIdx.x >> 5);
  if (x >= w || y >= h) return;
  double d[4];
  neighbour_diffs(I, h, w, c, x, y, d);
  const long long i = (long long)y * w + x;
  w_v[i] = (float)exp(-(d[0] * d[0]) * beta[0]);
  w_h[i] = (float)exp(-(d[1] * d[1]) * beta[1]);
  w_ne[i] = (float)exp(-(d[2] * d[2]) * beta[2]);
  w_se[i] = (float)exp(-(d[3] * d[3]) * beta[3]);
}


// utils/flow_io.py:44-46: tmp = payload.reshape(h, 2 w); u = tmp[:, 0::2]; v = tmp[:, 1::2].  The payload starts 12 bytes
// into the file, so it is only 4-byte aligned: one fp32 pair per thread as two 4-byte loads (the pair's 8 bytes are
// contiguous, a warp reads 256 contiguous bytes), planes written coalesced.
__global__ void __launch_bounds__(256)
k_flo_decode(const float* __restrict__ payload, long long n, float* __restrict__ u, float* __restrict__ v) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    u[i] = payload[2 * i];
    v[i] = payload[2 * i + 1];
  }
}

}  // namespace mrf

using namespace mrf;

extern "C" {

int mrf_flow_consistency(const float* uf0, const float* vf0, const float* ub0, const float* vb0, const float* uf1,
                         const float* vf1, const float* ub1, const float* vb1, int h, int w, double threshold,
                         uint8_t* occ_backward, uint8_t* occ_forward, uint8_t* occ_both, mrf_stream_t stream) {
  MRF_CHECK_ARG(uf0 && vf0 && ub0 && vb0 && uf1 && vf1 && ub1 && vb1 && occ_backward && occ_forward && h > 0 && w > 0);
  dim3 grid(cdiv(w, 32), cdiv(h, 8));
  k_flow_consistency<<<grid, 256, 0, S(stream)>>>(uf0, vf0, ub0, vb0, uf1, vf1, ub1, vb1, h, w, (float)threshold,
                                                  occ_backward, occ_forward, occ_both);
  return check_launch("mrf_flow_consistency");
}

size_t mrf_image_weights_scratch_bytes(void) { return sizeof(double) * (4 * kIwBlocks + 8); }

int mrf_image_weights(const double* I, int h, int w, int c, float* w_horizontal, float* w_vertical, float* w_ne,
                      float* w_se, void* scratch, mrf_stream_t stream) {
  MRF_CHECK_ARG(I && w_horizontal && w_vertical && w_ne && w_se && scratch && h > 0 && w > 0 && c >= 1 && c <= 4);
  double* partial = (double*)scratch;
  double* beta = partial + 4 * kIwBlocks;
  k_image_weight_sums<<<kIwBlocks, 256, 0, S(stream)>>>(I, h, w, c, partial);
  int rc = check_launch("mrf_image_weights/sums");
  if (rc) return rc;
  k_image_weight_betas<<<1, kIwBlocks, 0, S(stream)>>>(partial, kIwBlocks, (double)h * (double)w, beta);
  rc = check_launch("mrf_image_weights/betas");
  if (rc) return rc;
  dim3 grid(cdiv(w, 32), cdiv(h, 8));
  k_image_weights<<<grid, 256, 0, S(stream)>>>(I, h, w, c, beta, w_horizontal, w_vertical, w_ne, w_se);
  return check_launch("mrf_image_weights/exp");
}

int mrf_flo_decode(const void* file_dev, size_t file_bytes, int h, int w, float* u, float* v, mrf_stream_t stream) {
  MRF_CHECK_ARG(file_dev && u && v && h > 0 && w > 0);
  MRF_CHECK_ARG(((uintptr_t)file_dev & 3) == 0);
  const long long n = (long long)h * w;
  MRF_CHECK_ARG(file_bytes >= 12 + (size_t)n * 8);                       /* header + h*w (u, v) pairs */
  const float* payload = reinterpret_cast<const float*>(static_cast<const char*>(file_dev) + 12);
  const long long b = (n + 255) / 256;
  k_flo_decode<<<(unsigned)(b < (long long)kSMs * 8 ? b : (long long)kSMs * 8), 256, 0, S(stream)>>>(payload, n, u, v);
  return check_launch("mrf_flo_decode");
}

}  // extern "C"
