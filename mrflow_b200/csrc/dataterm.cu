// The linearised data term of the variational refinement:
//   * compute_derivs_through_H  structure_refinement/interpolate_through_H.py:53-97  (a9): image
//     derivatives after a homography by four cv2.warpPerspective calls (fp64 image, INTER_LINEAR on a
//     1/32-px fixed-point grid, BORDER_REPLICATE) and fp64 finite-difference scalings;
//   * computeDataTerm  structure_refinement/optimize_structure_single_scale.py:186-305: per-pixel warp of
//     the neighbour frame and of its derivative images, residual Iz, directional derivative, rho', the
//     diagonal / right-hand side of the linear system, two 5x5 median filters, and the energy.
// The arithmetic specification is oracle/cvemul.py (warp_perspective, perspective_transform64, remap),
// pinned bit-exact against OpenCV 4.13, and oracle/hotpath.py::compute_data_term, pinned against the
// reference's own computeDataTerm.  Streaming / gather kernels, one thread per pixel; this stage runs
// once per outer iteration (not per label), so it is two orders of magnitude lighter than the cost volume.
#include "sampling.cuh"

namespace mrf {

struct Mat3x4 { Mat3 m[4]; };

// cv2.warpPerspective(fp64, INTER_LINEAR, BORDER_REPLICATE) at one destination pixel; Mi = OpenCV's own
// inverse of the matrix, bw0 = OpenCV's block width (the x coordinate is split at block boundaries).
__device__ __forceinline__ void warp_persp_px(const double* __restrict__ I, int h, int w, int C, const Mat3& Mi, int bw0,
                                              int x, int y, double out[4]) {
  const int bxi = (x / bw0) * bw0;
  const double bx = (double)bxi, x1 = (double)(x - bxi), yy = (double)y;
  const double X0 = (Mi.m[0] * bx + Mi.m[1] * yy) + Mi.m[2];
  const double Y0 = (Mi.m[3] * bx + Mi.m[4] * yy) + Mi.m[5];
  const double W0 = (Mi.m[6] * bx + Mi.m[7] * yy) + Mi.m[8];
  double W = W0 + Mi.m[6] * x1;
  W = (W != 0.0) ? 32.0 / W : 0.0;
  const double fX = fmax(-2147483648.0, fmin(2147483647.0, (X0 + Mi.m[0] * x1) * W));
  const double fY = fmax(-2147483648.0, fmin(2147483647.0, (Y0 + Mi.m[3] * x1) * W));
  // saturate_cast<int>(double) = cvRound; NaN (0 * inf) follows the x86 "integer indefinite"
  const int X = (fX == fX) ? __double2int_rn(fX) : (int)0x80000000;
  const int Y = (fY == fY) ? __double2int_rn(fY) : (int)0x80000000;
  const int sx = max(-32768, min(32767, X >> 5)), sy = max(-32768, min(32767, Y >> 5));
  float wx[2], wy[2];
  linear_weights(X & 31, wx);
  linear_weights(Y & 31, wy);
  for (int c = 0; c < C; ++c) out[c] = 0.0;
#pragma unroll
  for (int j = 0; j < 2; ++j) {
    const int yc = max(0, min(h - 1, sy + j));
#pragma unroll
    for (int i = 0; i < 2; ++i) {
      const int xc = max(0, min(w - 1, sx + i));
      const double wgt = (double)(wy[j] * wx[i]);
      const double* p = I + ((long long)yc * w + xc) * C;
      for (int c = 0; c < C; ++c) {
        const double t = p[c] * wgt;
        out[c] = (i == 0 && j == 0) ? t : out[c] + t;
      }
    }
  }
}

// cv2.perspectiveTransform on an fp64 point (OpenCV 4.13's fp64 kernel is FMA-contracted:
// fma(x, m0, y*m1) + m2), then the distance to the original point (np.linalg.norm, :81-84)
__device__ __forceinline__ double displaced_norm(const Mat3& M, double x, double y) {
  const double wv = fma(x, M.m[6], y * M.m[7]) + M.m[8];
  const double iw = (fabs(wv) > 2.220446049250313e-16) ? 1.0 / wv : 0.0;
  const double px = (fma(x, M.m[0], y * M.m[1]) + M.m[2]) * iw;
  const double py = (fma(x, M.m[3], y * M.m[4]) + M.m[5]) * iw;
  const double ax = x - px, ay = y - py;
  return sqrt(ax * ax + ay * ay);
}

__global__ void __launch_bounds__(256)
k_derivs_through_H(const double* __restrict__ I, int h, int w, int C, Mat3x4 M, Mat3x4 Mi, int bw0,
                   double* __restrict__ dx, double* __restrict__ dy) {
  const int x = blockIdx.x * 32 + (threadIdx.x & 31);
  const int y = blockIdx.y * 8 + (threadIdx.x >> 5);
  if (x >= w || y >= h) return;
  const long long i = ((long long)y * w + x) * C;
  double Iw[4][4], d[4];
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    warp_persp_px(I, h, w, C, Mi.m[k], bw0, x, y, Iw[k]);
    d[k] = displaced_norm(M.m[k], (double)x, (double)y);
  }
  for (int c = 0; c < C; ++c) {
    const double v = I[i + c];
    dx[i + c] = 0.5 * ((Iw[0][c] - v) * d[0] + (v - Iw[1][c]) * d[1]);     // :95
    dy[i + c] = 0.5 * ((Iw[2][c] - v) * d[2] + (v - Iw[3][c]) * d[3]);     // :96
  }
}

// fp64 [h,w,3] -> packed fp32 texels [h,w,4] (the I.astype('float32') of interpolation.py:76)
__global__ void __launch_bounds__(256)
k_pack_texels(const double* __restrict__ src, long long n, float4* __restrict__ dst) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
    dst[i] = make_float4((float)src[3 * i], (float)src[3 * i + 1], (float)src[3 * i + 2], 0.f);
}

struct FetchTex {
  const float4* t; int w;
  __device__ __forceinline__ F3 operator()(int yy, int xx) const {
    const float4 v = __ldg(t + ((long long)yy * w + xx));
    F3 r; r.x = v.x; r.y = v.y; r.z = v.z; return r;
  }
};

// computeDataTerm :203-286: residual Iz, directional derivative dI2w_dphi and dr/da per pixel
__global__ void __launch_bounds__(256)
k_data_term_residual(const double* __restrict__ A, const double* __restrict__ ref_ch, const float4* __restrict__ nb,
                     const float4* __restrict__ nb_dx, const float4* __restrict__ nb_dy,
                     const double* __restrict__ ref_dx, const double* __restrict__ ref_dy,
                     const uint8_t* __restrict__ nonrigid, const uint8_t* __restrict__ occluded, int h, int w,
                     Mat3 H, Pt2 q, double mu, double B, double cw0, double cw1, double cw2,
                     double* __restrict__ Iz_out, double* __restrict__ dphi_out, double* __restrict__ drda_out) {
  const int x = blockIdx.x * 32 + (threadIdx.x & 31);
  const int y = blockIdx.y * 8 + (threadIdx.x >> 5);
  if (x >= w || y >= h) return;
  const long long i = (long long)y * w + x;
  const double xd = (double)x, yd = (double)y;
  const double vx = q.x - xd, vy = q.y - yd;
  const double dist = sqrt(vx * vx + vy * vy);
  const double nrm = fmax(0.5, dist);
  const double t1 = vx / nrm, t2 = vy / nrm;
  const double n = dist / mu;
  const double a = A[i];
  const double den = (B * a) / mu - 1.0;
  const double r = ((a * B) * n) / den;                       // :221
  const double drda = -(n * B) / (den * den);                 // :228
  const double xn = xd + r * t1, yn = yd + r * t2;            // :236-237
  double X, Y; bool inside;
  through_H(H, xn, yn, h, w, X, Y, inside);
  const float xf = (float)X, yf = (float)Y;
  FetchTex f0{nb, w}, f1{nb_dx, w}, f2{nb_dy, w};
  const F3 J = sample_cubic(f0, h, w, xf, yf);                // I2w
  F3 gx = sample_cubic(f1, h, w, xf, yf);                     // dI2w_dx
  F3 gy = sample_cubic(f2, h, w, xf, yf);                     // dI2w_dy
  if (!inside) { gx.x = gx.y = gx.z = 0.f; gy.x = gy.y = gy.z = 0.f; }     // :246-251
  // derivatives of I1 through the identity, sampled on the grid = the fp32 values themselves (:255-259),
  // summed in fp32 and halved where the warped sample is inside the frame (:261-269)
  const double div = inside ? 2.0 : 1.0;
  const double* rx = ref_dx + 3 * i; const double* ry = ref_dy + 3 * i;
  const double ax0 = (double)(gx.x + (float)rx[0]) / div, ax1 = (double)(gx.y + (float)rx[1]) / div,
               ax2 = (double)(gx.z + (float)rx[2]) / div;
  const double ay0 = (double)(gy.x + (float)ry[0]) / div, ay1 = (double)(gy.y + (float)ry[1]) / div,
               ay2 = (double)(gy.z + (float)ry[2]) / div;
  // the in-place division writes fp32 back (:265-269)
  const double bx0 = (double)(float)ax0, bx1 = (double)(float)ax1, bx2 = (double)(float)ax2;
  const double by0 = (double)(float)ay0, by1 = (double)(float)ay1, by2 = (double)(float)ay2;
  const double dxm = ((bx0 * cw0 + bx1 * cw1) + bx2 * cw2) / 3.0;          // :275
  const double dym = ((by0 * cw0 + by1 * cw1) + by2 * cw2) / 3.0;          // :276
  const double* c = ref_ch + 3 * i;
  double Iz = ((((double)J.x - c[0]) * cw0 + ((double)J.y - c[1]) * cw1) + ((double)J.z - c[2]) * cw2) / 3.0;   // :279
  const double dphi = t1 * dxm + t2 * dym;                    // :284
  if (!inside) Iz = 0.0;                                      // :287
  if (occluded && occluded[i]) Iz = 0.0;                      // :290
  if (nonrigid && nonrigid[i]) Iz = 0.0;                      // :291
  Iz_out[i] = Iz; dphi_out[i] = dphi; drda_out[i] = drda;
}

// computeDataTerm :294-302: psi = rho'(Iz^2); diag = psi*dphi^2*drda^2; b = psi*Iz*dphi*drda; E = sum rho(Iz^2)
constexpr int kDtBlocks = 1024;
__global__ void __launch_bounds__(256)
k_data_term_system(const double* __restrict__ Iz, const double* __restrict__ dphi, const double* __restrict__ drda,
                   long long n, int kind, double p1, const double* __restrict__ dev_median, double* __restrict__ diag,
                   double* __restrict__ b, double* __restrict__ partial) {
  __shared__ double sh[8];
  if (dev_median) {                            // auto-sigma from a device-resident median of |Iz| (robust_functions.py:23-29)
    const double m = *dev_median;
    p1 = (m != m) ? m : fmax(1e-6, 1.4826 * m);
  }
  const double s2 = p1 * p1;
  double acc = 0.0;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const double z = Iz[i], v = z * z;
    double psi, ev;
    switch (kind) {
      case MRF_RHO_LORENTZIAN:      psi = 1.0 / (s2 + v);          ev = log(1.0 + v / s2); break;
      case MRF_RHO_LORENTZIAN_AUTO: psi = p1 / (v + 2.0 * s2);     ev = p1 * log(1.0 + (0.5 * v) / s2); break;
      case MRF_RHO_CHARBONNIER:     psi = 1.0 / (2.0 * sqrt(v + s2)); ev = sqrt(v + s2); break;
      case MRF_RHO_GEMAN_MCCLURE:   { const double d = v + s2; psi = (p1 * p1 * p1) / (d * d); ev = (p1 * v) / d; break; }
      default:                      psi = 1.0; ev = v; break;
    }
    const double dp = dphi[i], da = drda[i];
    diag[i] = (psi * (dp * dp)) * (da * da);                   // psi_data * dI2w_dphi**2 * dr_da**2
    b[i] = ((psi * z) * dp) * da;                              // psi_data * Iz * dI2w_dphi * dr_da
    acc += ev;
  }
  for (int o = 16; o > 0; o >>= 1) acc += __shfl_down_sync(0xffffffffu, acc, o);
  if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x == 0) { double s = 0.0; for (int k = 0; k < 8; ++k) s += sh[k]; partial[blockIdx.x] = s; }
}

__global__ void __launch_bounds__(kDtBlocks)
k_sum_partials(const double* __restrict__ partial, int n, double* __restrict__ out) {
  __shared__ double sh[32];
  double v = (threadIdx.x < n) ? partial[threadIdx.x] : 0.0;
  for (int o = 16; o > 0; o >>= 1) v += __shfl_down_sync(0xffffffffu, v, o);
  if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = v;
  __syncthreads();
  if (threadIdx.x < 32) {
    v = (threadIdx.x < (blockDim.x >> 5)) ? sh[threadIdx.x] : 0.0;
    for (int o = 16; o > 0; o >>= 1) v += __shfl_down_sync(0xffffffffu, v, o);
    if (threadIdx.x == 0) out[0] = v;
  }
}

// scipy.ndimage.median_filter(size=5), mode='reflect' (d c b a | a b c d | d c b a)
__device__ __forceinline__ int reflect_sym(int i, int n) {
  if (n == 1) return 0;
  const int p = 2 * n;
  i %= p; if (i < 0) i += p;
  return i < n ? i : p - 1 - i;
}
__global__ void __launch_bounds__(256)
k_median5x5(const double* __restrict__ src, int h, int w, double* __restrict__ dst) {
  const int x = blockIdx.x * 32 + (threadIdx.x & 31);
  const int y = blockIdx.y * 8 + (threadIdx.x >> 5);
  if (x >= w || y >= h) return;
  int xs[5]; long long ys[5];
#pragma unroll
  for (int k = 0; k < 5; ++k) { xs[k] = reflect_sym(x + k - 2, w); ys[k] = (long long)reflect_sym(y + k - 2, h) * w; }
  const double med = median_forgetful<25>([&](int k) -> double { return src[ys[k / 5] + xs[k % 5]]; });
  dst[(long long)y * w + x] = med;
}

}  // namespace mrf

using namespace mrf;

extern "C" {

int mrf_derivs_through_H(const double* I, int h, int w, int C, const double* M4_host, const double* Minv4_host, int bw0,
                         double* dx, double* dy, mrf_stream_t stream) {
  MRF_CHECK_ARG(I && dx && dy && M4_host && Minv4_host && h > 0 && w > 0 && C >= 1 && C <= 4 && bw0 > 0);
  Mat3x4 M, Mi;
  for (int k = 0; k < 4; ++k) { M.m[k] = mat3(M4_host + 9 * k); Mi.m[k] = mat3(Minv4_host + 9 * k); }
  dim3 grid(cdiv(w, 32), cdiv(h, 8));
  k_derivs_through_H<<<grid, 256, 0, S(stream)>>>(I, h, w, C, M, Mi, bw0, dx, dy);
  return check_launch("mrf_derivs_through_H");
}

int mrf_pack_texels(const double* src, int h, int w, float* texels, mrf_stream_t stream) {
  MRF_CHECK_ARG(src && texels && h > 0 && w > 0 && (((uintptr_t)texels) & 15) == 0);
  const long long n = (long long)h * w;
  long long b = (n + 255) / 256; if (b > kSMs * 16) b = kSMs * 16;
  k_pack_texels<<<(unsigned)b, 256, 0, S(stream)>>>(src, n, (float4*)texels);
  return check_launch("mrf_pack_texels");
}

int mrf_data_term_residual(const double* A, const double* ref_ch, const float* nb_texels, const float* nb_dx_texels,
                           const float* nb_dy_texels, const double* ref_dx, const double* ref_dy,
                           const uint8_t* nonrigid, const uint8_t* occluded, int h, int w, const double* H_host,
                           const double* q_host, double mu, double B, const double* cw_host, double* Iz, double* dphi,
                           double* dr_da, mrf_stream_t stream) {
  MRF_CHECK_ARG(A && ref_ch && nb_texels && nb_dx_texels && nb_dy_texels && ref_dx && ref_dy && H_host && q_host && cw_host);
  MRF_CHECK_ARG(Iz && dphi && dr_da && h > 0 && w > 0);
  dim3 grid(cdiv(w, 32), cdiv(h, 8));
  k_data_term_residual<<<grid, 256, 0, S(stream)>>>(A, ref_ch, (const float4*)nb_texels, (const float4*)nb_dx_texels,
                                                    (const float4*)nb_dy_texels, ref_dx, ref_dy, nonrigid, occluded, h, w,
                                                    mat3(H_host), pt2(q_host), mu, B, cw_host[0], cw_host[1], cw_host[2], Iz,
                                                    dphi, dr_da);
  return check_launch("mrf_data_term_residual");
}

size_t mrf_data_term_scratch_bytes(void) { return sizeof(double) * kDtBlocks; }

int mrf_data_term_system(const double* Iz, const double* dphi, const double* dr_da, size_t n, int rho, double rho_param, const double* dev_median,
                         double* diag, double* b, double* energy, void* scratch, mrf_stream_t stream) {
  MRF_CHECK_ARG(Iz && dphi && dr_da && diag && b && energy && scratch && n > 0);
  MRF_CHECK_ARG(rho >= MRF_RHO_LORENTZIAN && rho <= MRF_RHO_LORENTZIAN_AUTO);
  k_data_term_system<<<kDtBlocks, 256, 0, S(stream)>>>(Iz, dphi, dr_da, (long long)n, rho, rho_param, dev_median, diag, b,
                                                       (double*)scratch);
  int rc = check_launch("mrf_data_term_system");
  if (rc) return rc;
  k_sum_partials<<<1, kDtBlocks, 0, S(stream)>>>((const double*)scratch, kDtBlocks, energy);
  return check_launch("mrf_data_term_system/energy");
}

int mrf_median5x5_f64(const double* src, int h, int w, double* dst, mrf_stream_t stream) {
  MRF_CHECK_ARG(src && dst && src != dst && h > 0 && w > 0);
  dim3 grid(cdiv(w, 32), cdiv(h, 8));
  k_median5x5<<<grid, 256, 0, S(stream)>>>(src, h, w, dst);
  return check_launch("mrf_median5x5_f64");
}

}  // extern "C"
