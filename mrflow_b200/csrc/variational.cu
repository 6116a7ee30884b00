// The variational structure refinement around the data term:
// structure_refinement/optimize_structure_single_scale.py:310-652 without its scipy.sparse matrices.
//   * edge weights        utils/edge_weights.py:38-55 (computeWeightsLocallyNormalized)
//   * smoothness terms    :505-550 -- first / second differences with replicated borders
//                         (structure_refinement/construct_matrices.py:14-38), robust arguments
//   * the linear operator :580-595 -- A = Z (D + l1 G'P1G + l2 G2'P2G2 + lc Pc) Z + 0.01 I applied
//                         matrix-free as a 5x5-footprint stencil (one thread per pixel, fp64)
//   * MINRES              scipy.sparse.linalg.minres (the solver the reference calls, :86,:593) as one persistent
//                         cooperative kernel: operator, Lanczos step, solution update, the scalar recurrences
//                         and the stopping tests all on the device, two grid barriers per iteration
//   * update              clip_dA :157-183, clip to +-1 :612, 7x7 median of A + dA :636
// Unlike the cost volume this stage is iterative and latency-bound (two grid-wide reductions per MINRES
// iteration, ~100 iterations per solve, 5 solves per frame); DESIGN.md section 4.5 has the measured breakdown.
#include "common.cuh"
#include <cooperative_groups.h>
#include <math.h>
#include <stdlib.h>

namespace mrf {

constexpr int kVarBlocks = 1024;

// utils/robust_functions.py:23-29 from the median of sqrt(x): max(1e-6, 1.4826 * median), NaN kept
__device__ __forceinline__ double auto_sigma(double m) { return (m != m) ? m : fmax(1e-6, 1.4826 * m); }

__device__ __forceinline__ double block_sum256(double v, double* sh) {
  for (int o = 16; o > 0; o >>= 1) v += __shfl_down_sync(0xffffffffu, v, o);
  if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = v;
  __syncthreads();
  double r = 0.0;
  if (threadIdx.x == 0) for (int k = 0; k < (int)(blockDim.x >> 5); ++k) r += sh[k];
  __syncthreads();
  return r;      // valid in thread 0
}

__global__ void __launch_bounds__(kVarBlocks)
k_var_sum_final(const double* __restrict__ partial, int n, int nvals, double* __restrict__ out) {
  __shared__ double sh[32];
  for (int k = 0; k < nvals; ++k) {
    double v = (threadIdx.x < n) ? partial[k * n + threadIdx.x] : 0.0;
    v = block_sum256(v, sh);
    if (threadIdx.x == 0) out[k] = v;
  }
}

// ------------------------------------------------------------------------------- edge weights
__device__ __forceinline__ int refl101(int i, int n) {
  if (n == 1) return 0;
  while (i < 0 || i >= n) { if (i < 0) i = -i; if (i >= n) i = 2 * n - 2 - i; }
  return i;
}
// cv2.cvtColor(I.astype('float32'), COLOR_RGB2GRAY) (:381-383), kept in fp32
template <typename T>
__global__ void __launch_bounds__(256)
k_gray32(const T* __restrict__ rgb, long long n, float* __restrict__ gray) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const float r = (float)rgb[3 * i], g = (float)rgb[3 * i + 1], b = (float)rgb[3 * i + 2];
    gray[i] = fmaf(b, 0.114f, fmaf(r, 0.299f, g * 0.587f));
  }
}
// np.gradient of an fp32 image, squared (fp32): gxsq, gysq
__global__ void __launch_bounds__(256)
k_grad_sq32(const float* __restrict__ I, int h, int w, float* __restrict__ gxsq, float* __restrict__ gysq) {
  const int x = blockIdx.x * 32 + (threadIdx.x & 31), y = blockIdx.y * 8 + (threadIdx.x >> 5);
  if (x >= w || y >= h) return;
  const long long i = (long long)y * w + x;
  float gx, gy;
  if (w == 1) gx = 0.f; else if (x == 0) gx = I[i + 1] - I[i]; else if (x == w - 1) gx = I[i] - I[i - 1]; else gx = (I[i + 1] - I[i - 1]) / 2.0f;
  if (h == 1) gy = 0.f; else if (y == 0) gy = I[i + w] - I[i]; else if (y == h - 1) gy = I[i] - I[i - w]; else gy = (I[i + w] - I[i - w]) / 2.0f;
  gxsq[i] = gx * gx; gysq[i] = gy * gy;
}
// cv2.blur(fp32, (r, r)) = normalised box filter, BORDER_REFLECT_101, sums in fp64, one rounding to fp32;
// horizontal pass to fp64, vertical pass + weight exp(-g * 1.0 / (2 * max(1e-6, mean)))  (fp32, :49-50)
__global__ void __launch_bounds__(256)
k_box_rows(const float* __restrict__ a, const float* __restrict__ b, int h, int w, int radius, double* __restrict__ ra,
           double* __restrict__ rb) {
  const int x = blockIdx.x * 32 + (threadIdx.x & 31), y = blockIdx.y * 8 + (threadIdx.x >> 5);
  if (x >= w || y >= h) return;
  const int half = radius / 2;
  double sa = 0.0, sb = 0.0;
  for (int k = -half; k <= half; ++k) {
    const long long j = (long long)y * w + refl101(x + k, w);
    sa += (double)a[j]; sb += (double)b[j];
  }
  ra[(long long)y * w + x] = sa; rb[(long long)y * w + x] = sb;
}
__global__ void __launch_bounds__(256)
k_box_cols_weights(const double* __restrict__ ra, const double* __restrict__ rb, const float* __restrict__ gxsq,
                   const float* __restrict__ gysq, int h, int w, int radius, float* __restrict__ wx, float* __restrict__ wy) {
  const int x = blockIdx.x * 32 + (threadIdx.x & 31), y = blockIdx.y * 8 + (threadIdx.x >> 5);
  if (x >= w || y >= h) return;
  const int half = radius / 2;
  double sa = 0.0, sb = 0.0;
  for (int k = -half; k <= half; ++k) {
    const long long j = (long long)refl101(y + k, h) * w + x;
    sa += ra[j]; sb += rb[j];
  }
  const double scale = 1.0 / ((double)radius * (double)radius);
  const float ma = (float)(sa * scale), mb = (float)(sb * scale);
  const long long i = (long long)y * w + x;
  wx[i] = expf((-gxsq[i] * 1.0f) / (2.0f * fmaxf(1e-6f, ma)));
  wy[i] = expf((-gysq[i] * 1.0f) / (2.0f * fmaxf(1e-6f, mb)));
}
// weights[eroded == 0] = 0; weights = max(weights, min_weight)  (:388-392)
__global__ void __launch_bounds__(256)
k_mask_weights(float* __restrict__ wx, float* __restrict__ wy, const uint8_t* __restrict__ eroded, long long n, float min_w) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    float a = wx[i], b = wy[i];
    if (!eroded[i]) { a = 0.f; b = 0.f; }
    wx[i] = fmaxf(a, min_w); wy[i] = fmaxf(b, min_w);
  }
}

// ------------------------------------------------------------------------------- differences
__device__ __forceinline__ double dfx(const double* __restrict__ v, int w, int x, long long i) { return (x + 1 < w) ? v[i + 1] - v[i] : 0.0; }
__device__ __forceinline__ double dfy(const double* __restrict__ v, int h, int w, int y, long long i) { return (y + 1 < h) ? v[i + w] - v[i] : 0.0; }
__device__ __forceinline__ double dxx(const double* __restrict__ v, int w, int x, long long i) {
  return (v[x > 0 ? i - 1 : i] + v[x + 1 < w ? i + 1 : i]) - 2.0 * v[i];
}
__device__ __forceinline__ double dyy(const double* __restrict__ v, int h, int w, int y, long long i) {
  return (v[y > 0 ? i - w : i] + v[y + 1 < h ? i + w : i]) - 2.0 * v[i];
}

// arg1 = afac_sq * (wx*gx^2 + wy*gy^2), arg2 = afac_sq * (wx*dxx^2 + wy*dyy^2)   (:508-511, :541-547)
__global__ void __launch_bounds__(256)
k_smooth_args(const double* __restrict__ A, const float* __restrict__ wx, const float* __restrict__ wy, int h, int w,
              double afac_sq, double* __restrict__ arg1, double* __restrict__ arg2, double* __restrict__ root1,
              double* __restrict__ root2) {
  const int x = blockIdx.x * 32 + (threadIdx.x & 31), y = blockIdx.y * 8 + (threadIdx.x >> 5);
  if (x >= w || y >= h) return;
  const long long i = (long long)y * w + x;
  const double a = (double)wx[i], b = (double)wy[i];
  const double gx = dfx(A, w, x, i), gy = dfy(A, h, w, y, i);
  const double sx = dxx(A, w, x, i), sy = dyy(A, h, w, y, i);
  const double a1 = afac_sq * (a * (gx * gx) + b * (gy * gy));
  const double a2 = afac_sq * (a * (sx * sx) + b * (sy * sy));
  arg1[i] = a1; arg2[i] = a2;
  root1[i] = sqrt(a1); root2[i] = sqrt(a2);      // the samples of the auto-sigma median (robust_functions.py:23-29)
}

// Lorentzian auto-sigma weights of the two smoothness terms and their energies (:511,:547,:560-561):
// p1x = wx*psi1, p1y = wy*psi1, p2x = psi2*wx, p2y = psi2*wy ; partial sums of rho(arg1), rho(arg2)
__global__ void __launch_bounds__(256)
k_smooth_weights(const double* __restrict__ arg1, const double* __restrict__ arg2, const float* __restrict__ wx,
                 const float* __restrict__ wy, long long n, double afac_sq, double s1, double s2,
                 const double* __restrict__ dev_medians, double* __restrict__ p1x, double* __restrict__ p1y,
                 double* __restrict__ p2x, double* __restrict__ p2y, double* __restrict__ partial) {
  __shared__ double sh[8];
  if (dev_medians) { s1 = auto_sigma(dev_medians[0]); s2 = auto_sigma(dev_medians[1]); }
  double e1 = 0.0, e2 = 0.0;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const double a1 = arg1[i], a2 = arg2[i];
    const double psi1 = afac_sq * (s1 / (a1 + 2.0 * (s1 * s1)));
    const double psi2 = afac_sq * (s2 / (a2 + 2.0 * (s2 * s2)));
    const double a = (double)wx[i], b = (double)wy[i];
    p1x[i] = a * psi1; p1y[i] = b * psi1; p2x[i] = psi2 * a; p2y[i] = psi2 * b;
    e1 += s1 * log(1.0 + (0.5 * a1) / (s1 * s1));
    e2 += s2 * log(1.0 + (0.5 * a2) / (s2 * s2));
  }
  e1 = block_sum256(e1, sh);
  e2 = block_sum256(e2, sh);
  if (threadIdx.x == 0) { partial[blockIdx.x] = e1; partial[gridDim.x + blockIdx.x] = e2; }
}

// data-term merge (:448-461) and consistency residual (:476-490)
__global__ void __launch_bounds__(256)
k_var_data_merge(const double* __restrict__ diag_f, const double* __restrict__ b_f, const double* __restrict__ diag_b,
                 const double* __restrict__ b_b, const uint8_t* __restrict__ occ_f, const uint8_t* __restrict__ occ_b,
                 const double* __restrict__ A, const double* __restrict__ A_bwd_ref, const double* __restrict__ A_fwd_ref,
                 long long n, double s_bwd, double s_ref, double afac_sq, double* __restrict__ D, double* __restrict__ bd,
                 double* __restrict__ Izc, double* __restrict__ argc, double* __restrict__ rootc) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const bool of = occ_f[i] != 0, ob = occ_b[i] != 0;
    D[i] = of ? diag_b[i] * s_bwd : diag_f[i];
    bd[i] = of ? b_b[i] * s_bwd : b_f[i];
    const double a = A[i], rb = A_bwd_ref[i] * s_ref, rf = A_fwd_ref[i];
    const double hi = fmax(rf, rb), lo = fmin(rf, rb);
    double z = (a - hi) * (a > hi ? 1.0 : 0.0) + (a - lo) * (a < lo ? 1.0 : 0.0);
    if (ob) z = a - rf;
    if (of) z = a - rb;
    Izc[i] = z;
    const double ac = afac_sq * (z * z);
    argc[i] = ac;
    rootc[i] = sqrt(ac);
  }
}

// Charbonnier (auto eps) weight and energy of the consistency term (:491-497)
__global__ void __launch_bounds__(256)
k_cons_weights(const double* __restrict__ argc, long long n, double afac_sq, double eps, const double* __restrict__ dev_median,
               double* __restrict__ pc, double* __restrict__ partial) {
  __shared__ double sh[8];
  if (dev_median) eps = auto_sigma(*dev_median);
  double e = 0.0;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const double r = sqrt(argc[i] + eps * eps);
    pc[i] = afac_sq * (1.0 / (2.0 * r));
    e += r;
  }
  e = block_sum256(e, sh);
  if (threadIdx.x == 0) partial[blockIdx.x] = e;
}

// ------------------------------------------------------------------------------- the operator
struct VarOp {
  const uint8_t* rigid;        // Z = diag(rigidity == 1), may be null (identity)
  const double* D;             // data diagonal, may be null
  const double *p1x, *p1y, *p2x, *p2y;
  const double* pc;            // consistency diagonal, may be null
  double l1, l2, lc, eps;
  int h, w;
};

__device__ __forceinline__ double zval(const VarOp& o, const double* __restrict__ x, long long i) {
  return (!o.rigid || o.rigid[i]) ? x[i] : 0.0;
}
// G'P1G v and G2'P2G2 v at pixel (x, y), v = Z x gathered on the fly
// the ten smoothness weights a pixel's row of the operator touches (centre, left / up, right / down)
struct SmoothCoef { double p1x_c, p1x_l, p1y_c, p1y_u, p2x_l, p2x_c, p2x_r, p2y_u, p2y_c, p2y_d; };
__device__ __forceinline__ SmoothCoef load_coef(const VarOp& o, int x, int y) {
  const int h = o.h, w = o.w;
  const long long i = (long long)y * w + x;
  SmoothCoef c;
  c.p1x_c = o.p1x[i]; c.p1y_c = o.p1y[i]; c.p2x_c = o.p2x[i]; c.p2y_c = o.p2y[i];
  c.p1x_l = x > 0 ? o.p1x[i - 1] : 0.0;     c.p2x_l = x > 0 ? o.p2x[i - 1] : 0.0;
  c.p1y_u = y > 0 ? o.p1y[i - w] : 0.0;     c.p2y_u = y > 0 ? o.p2y[i - w] : 0.0;
  c.p2x_r = x + 1 < w ? o.p2x[i + 1] : 0.0; c.p2y_d = y + 1 < h ? o.p2y[i + w] : 0.0;
  return c;
}
// INTERIOR: every neighbour within two pixels is inside the frame, so the border cases fold away at compile time
template <bool INTERIOR, class Acc>
__device__ __forceinline__ double smooth_apply_c(const VarOp& o, int x, int y, const SmoothCoef& c, Acc V) {
  const int h = o.h, w = o.w;
  auto DX = [&](int yy, int xx) -> double { return (INTERIOR || xx + 1 < w) ? V(yy, xx + 1) - V(yy, xx) : 0.0; };
  auto DY = [&](int yy, int xx) -> double { return (INTERIOR || yy + 1 < h) ? V(yy + 1, xx) - V(yy, xx) : 0.0; };
  auto DXX = [&](int yy, int xx) -> double {
    return (V(yy, (INTERIOR || xx > 0) ? xx - 1 : xx) + V(yy, (INTERIOR || xx + 1 < w) ? xx + 1 : xx)) - 2.0 * V(yy, xx);
  };
  auto DYY = [&](int yy, int xx) -> double {
    return (V((INTERIOR || yy > 0) ? yy - 1 : yy, xx) + V((INTERIOR || yy + 1 < h) ? yy + 1 : yy, xx)) - 2.0 * V(yy, xx);
  };
  // first order: (G' u)(p) = -ux(p) + ux(p - ex) - uy(p) + uy(p - ey),  u = P1 G v
  double s1 = 0.0;
  s1 -= c.p1x_c * DX(y, x);
  if (INTERIOR || x > 0) s1 += c.p1x_l * DX(y, x - 1);
  s1 -= c.p1y_c * DY(y, x);
  if (INTERIOR || y > 0) s1 += c.p1y_u * DY(y - 1, x);
  // second order: (Dxx' u)(p) = u(p-1) + u(p+1) - 2u(p), the replicated ends folding back onto p
  double s2 = 0.0;
  {
    const double uc = c.p2x_c * DXX(y, x);
    double t = -2.0 * uc;
    if (!INTERIOR && x == 0) t += uc;
    if (!INTERIOR && x == w - 1) t += uc;
    if (INTERIOR || x > 0) t += c.p2x_l * DXX(y, x - 1);
    if (INTERIOR || x + 1 < w) t += c.p2x_r * DXX(y, x + 1);
    s2 += t;
  }
  {
    const double uc = c.p2y_c * DYY(y, x);
    double t = -2.0 * uc;
    if (!INTERIOR && y == 0) t += uc;
    if (!INTERIOR && y == h - 1) t += uc;
    if (INTERIOR || y > 0) t += c.p2y_u * DYY(y - 1, x);
    if (INTERIOR || y + 1 < h) t += c.p2y_d * DYY(y + 1, x);
    s2 += t;
  }
  return o.l1 * s1 + o.l2 * s2;
}
template <class Acc>
__device__ __forceinline__ double smooth_apply_f(const VarOp& o, int x, int y, Acc V) {
  return smooth_apply_c<false>(o, x, y, load_coef(o, x, y), V);
}
__device__ __forceinline__ double smooth_apply(const VarOp& o, const double* __restrict__ xv, int x, int y) {
  const int w = o.w;
  return smooth_apply_f(o, x, y, [&](int yy, int xx) -> double { return zval(o, xv, (long long)yy * w + xx); });
}

__global__ void __launch_bounds__(256)
k_var_apply(VarOp o, const double* __restrict__ x, double* __restrict__ y) {
  const int px = blockIdx.x * 32 + (threadIdx.x & 31), py = blockIdx.y * 8 + (threadIdx.x >> 5);
  if (px >= o.w || py >= o.h) return;
  const long long i = (long long)py * o.w + px;
  double r = 0.0;
  if (!o.rigid || o.rigid[i]) {
    const double v = x[i];
    r = smooth_apply(o, x, px, py);
    if (o.D) r += o.D[i] * v;
    if (o.pc) r += o.lc * (o.pc[i] * v);
  }
  y[i] = r + o.eps * x[i];
}

// b = Z (-bd - (l1 A1 + l2 A2) A - lc * pc * Izc)   (:581,:588)
__global__ void __launch_bounds__(256)
k_var_rhs(const double* __restrict__ bd, const double* __restrict__ sA, const double* __restrict__ pc,
          const double* __restrict__ Izc, const uint8_t* __restrict__ rigid, long long n, double lc, double* __restrict__ b) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    double v = -bd[i] - sA[i];
    if (pc) v -= lc * (pc[i] * Izc[i]);
    b[i] = rigid[i] ? v : 0.0;
  }
}

// ------------------------------------------------------------------------------- MINRES, device-resident
// scipy.sparse.linalg.minres (x0 = 0, no preconditioner, shift 0) as ONE persistent cooperative kernel: the
// grid (as many blocks as are co-resident) walks the whole solve, two grid-wide barriers per iteration.
//   phase A: v = y / beta (tile of v staged in shared memory);  t = Op v  (- (beta/oldb) r1 from the 2nd
//            iteration);  partial <v, t>                                                  -- barrier 1 --
//   phase B: t -= (alfa/beta) r2;  (r1, r2) <- (r2, t) by buffer rotation;  partial |t|^2  -- barrier 2 --
//   phase C: w' = (v - oldeps w1 - delta w2) / gamma;  x += phi w';  partial |x|^2   (no barrier: its sum is
//            read after barrier 1 of the next iteration, v is double-buffered so phase A may run ahead)
// After a barrier every block adds the per-block partials in index order and advances the scalar
// recurrences itself -- identical arithmetic in every block, so no broadcast and no host round trip.  The
// 256-byte state block makes the solve resumable (the host may bound the iterations per launch).
enum {
  MS_BETA1 = 0, MS_OLDB, MS_BETA, MS_DBAR, MS_EPSLN, MS_PHIBAR, MS_RHS1, MS_RHS2, MS_TNORM2, MS_GMAX, MS_GMIN, MS_CS, MS_SN,
  MS_ALFA, MS_OLDEPS, MS_DELTA, MS_DENOM, MS_PHI, MS_ITN, MS_ISTOP, MS_ANORM, MS_ACOND, MS_RNORM, MS_YNORM, MS_ROOT, MS_GBAR,
  MS_GAMMA, MS_TEST1, MS_TEST2, MS_NS_A, MS_NS_BC, MS_NS_SYNC, MS_WORDS = 32
};
static_assert(MS_ITN == MRF_MINRES_ITN && MS_ISTOP == MRF_MINRES_ISTOP && MS_ANORM == MRF_MINRES_ANORM &&
              MS_ACOND == MRF_MINRES_ACOND && MS_RNORM == MRF_MINRES_RNORM && MS_YNORM == MRF_MINRES_YNORM &&
              MS_WORDS == MRF_MINRES_STATE_WORDS, "state layout is part of the ABI");
constexpr double kEps = 2.220446049250313e-16;
constexpr double kDblMax = 1.7976931348623157e308;

struct MinresScalars {
  double s[MS_WORDS];
  __device__ void load(const double* __restrict__ g) { for (int k = 0; k < MS_WORDS; ++k) s[k] = g[k]; }
  __device__ void store(double* __restrict__ g) const { for (int k = 0; k < MS_WORDS; ++k) g[k] = s[k]; }
  // after <r2, y> of the Lanczos step: new beta, the previous plane rotation applied, the next one formed
  __device__ void lanczos(double alfa, double beta_sq) {
    const double itn = s[MS_ITN] + 1.0;
    const double oldb = s[MS_BETA];
    const double beta = sqrt(beta_sq);
    // a non-finite system never satisfies a stopping test (scipy would grind through all 5n iterations on
    // NaNs): stop at once with istop 7; the caller then marks the whole solution non-finite, as scipy's is
    if (!(alfa - alfa == 0.0) || !(beta_sq - beta_sq == 0.0)) s[MS_ISTOP] = 7.0;
    s[MS_ALFA] = alfa; s[MS_OLDB] = oldb; s[MS_BETA] = beta;
    s[MS_TNORM2] = s[MS_TNORM2] + ((alfa * alfa + oldb * oldb) + beta * beta);
    if (itn == 1.0 && beta / s[MS_BETA1] <= 10.0 * kEps) s[MS_ISTOP] = -1.0;
    const double cs = s[MS_CS], sn = s[MS_SN], dbar0 = s[MS_DBAR];
    s[MS_OLDEPS] = s[MS_EPSLN];
    const double delta = cs * dbar0 + sn * alfa;
    const double gbar = sn * dbar0 - cs * alfa;
    const double epsln = sn * beta;
    const double dbar = -cs * beta;
    const double root = sqrt(gbar * gbar + dbar * dbar);
    double gamma = sqrt(gbar * gbar + beta * beta);
    gamma = fmax(gamma, kEps);
    const double phibar0 = s[MS_PHIBAR];
    const double cs1 = gbar / gamma, sn1 = beta / gamma;
    s[MS_DELTA] = delta; s[MS_GBAR] = gbar; s[MS_EPSLN] = epsln; s[MS_DBAR] = dbar; s[MS_ROOT] = root;
    s[MS_GAMMA] = gamma; s[MS_CS] = cs1; s[MS_SN] = sn1;
    s[MS_PHI] = cs1 * phibar0; s[MS_PHIBAR] = sn1 * phibar0;
    s[MS_DENOM] = 1.0 / gamma;
  }
  // after |x|^2: norms and stopping tests of the iteration
  __device__ void tests(double ynorm_sq, double rtol, double maxiter) {
    const double itn = s[MS_ITN] + 1.0;
    s[MS_ITN] = itn;
    const double gamma = s[MS_GAMMA];
    const double gmax = fmax(s[MS_GMAX], gamma), gmin = fmin(s[MS_GMIN], gamma);
    s[MS_GMAX] = gmax; s[MS_GMIN] = gmin;
    const double z = s[MS_RHS1] / gamma;
    s[MS_RHS1] = s[MS_RHS2] - s[MS_DELTA] * z;
    s[MS_RHS2] = -s[MS_EPSLN] * z;
    const double Anorm = sqrt(s[MS_TNORM2]);
    const double ynorm = sqrt(ynorm_sq);
    const double epsx = (Anorm * ynorm) * kEps;
    const double rnorm = s[MS_PHIBAR];
    const double inf = __longlong_as_double(0x7FF0000000000000ll);
    const double test1 = (ynorm == 0.0 || Anorm == 0.0) ? inf : rnorm / (Anorm * ynorm);
    const double test2 = (Anorm == 0.0) ? inf : s[MS_ROOT] / Anorm;
    const double Acond = gmax / gmin;
    double istop = s[MS_ISTOP];
    if (istop == 0.0) {
      if (1.0 + test2 <= 1.0) istop = 2.0;
      if (1.0 + test1 <= 1.0) istop = 1.0;
      if (itn >= maxiter) istop = 6.0;
      if (Acond >= 0.1 / kEps) istop = 4.0;
      if (epsx >= s[MS_BETA1]) istop = 3.0;
      if (test2 <= rtol) istop = 2.0;
      if (test1 <= rtol) istop = 1.0;
    }
    s[MS_ISTOP] = istop; s[MS_ANORM] = Anorm; s[MS_ACOND] = Acond; s[MS_RNORM] = rnorm; s[MS_YNORM] = ynorm;
    s[MS_TEST1] = test1; s[MS_TEST2] = test2;
  }
};

// every block adds the G partials in the same order -> the same bits everywhere; result in all threads
__device__ __forceinline__ double sum_partials_all(const double* partial, unsigned G, double* sh, double* sh_out) {
  double acc = 0.0;
  for (unsigned k = threadIdx.x; k < G; k += blockDim.x) acc += __ldcg(partial + k);
  acc = block_sum256(acc, sh);
  if (threadIdx.x == 0) *sh_out = acc;
  __syncthreads();
  const double r = *sh_out;
  __syncthreads();
  return r;
}

// two reductions at once (their loads overlap): sums of pa[0..G) and pb[0..G), same fixed order as above
__device__ __forceinline__ void sum_partials_all2(const double* pa, const double* pb, unsigned G, double* sh, double* sh_out2,
                                                  double& ra, double& rb) {
  double a = 0.0, b = 0.0;
  for (unsigned k = threadIdx.x; k < G; k += blockDim.x) { a += __ldcg(pa + k); b += __ldcg(pb + k); }
  a = block_sum256(a, sh);
  b = block_sum256(b, sh);
  if (threadIdx.x == 0) { sh_out2[0] = a; sh_out2[1] = b; }
  __syncthreads();
  ra = sh_out2[0]; rb = sh_out2[1];
  __syncthreads();
}

__global__ void __launch_bounds__(256)
k_minres_init(const double* __restrict__ b, long long n, double* __restrict__ r1, double* __restrict__ r2, double* __restrict__ w,
              double* __restrict__ w2, double* __restrict__ x, double* __restrict__ partial) {
  __shared__ double sh[8];
  double acc = 0.0;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const double v = b[i];
    r1[i] = v; r2[i] = v; w[i] = 0.0; w2[i] = 0.0; x[i] = 0.0;
    acc += v * v;
  }
  acc = block_sum256(acc, sh);
  if (threadIdx.x == 0) partial[blockIdx.x] = acc;
}
__global__ void __launch_bounds__(kVarBlocks)
k_minres_init_state(const double* __restrict__ partial, int nblocks, double* __restrict__ state) {
  __shared__ double sh[32];
  double v = (threadIdx.x < nblocks) ? partial[threadIdx.x] : 0.0;
  v = block_sum256(v, sh);
  if (threadIdx.x == 0) {
    for (int k = 0; k < MS_WORDS; ++k) state[k] = 0.0;
    const double beta1 = sqrt(v);
    state[MS_BETA1] = beta1; state[MS_BETA] = beta1; state[MS_PHIBAR] = beta1; state[MS_RHS1] = beta1;
    state[MS_GMIN] = kDblMax; state[MS_CS] = -1.0;
    if (v == 0.0) state[MS_ISTOP] = 100.0;            // b = 0: x = 0 is the answer (scipy returns before iterating)
    if (!(v - v == 0.0)) state[MS_ISTOP] = 7.0;       // non-finite right-hand side
  }
}

constexpr int kTileW = 32, kTileH = 8, kHalo = 2;
__device__ __forceinline__ unsigned long long gtime() { unsigned long long t; asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t)); return t; }
constexpr int kBatch = 3;                      // elements per thread whose loads are issued together in phases B, C

__global__ void __launch_bounds__(256, 2)
k_minres_persistent(VarOp o, double* r3, double* w2buf, double* v2buf, double* x, double* state, double* partial, int itn_done,
                    int n_iters, double rtol, double maxiter) {
  namespace cg = cooperative_groups;
  cg::grid_group grid = cg::this_grid();
  __shared__ MinresScalars m;                  // every block keeps its own copy, advanced by its thread 0
  __shared__ double tile[kTileH + 2 * kHalo][kTileW + 2 * kHalo + 1];
  __shared__ double sh[8];
  __shared__ double sh_out;
  __shared__ double sh_out2[2];
  const unsigned G = gridDim.x, bid = blockIdx.x;
  const bool prof = (bid == 0 && threadIdx.x == 0);       // block 0 accounts its own time per phase (ns) in the state
  unsigned long long nsA = 0, nsBC = 0, nsSync = 0;
  const int h = o.h, w = o.w;
  const long long n = (long long)h * w;
  const long long stride = (long long)G * 256;
  const int tiles_x = (w + kTileW - 1) / kTileW, tiles_y = (h + kTileH - 1) / kTileH;
  const int ntiles = tiles_x * tiles_y;
  double* pa = partial; double* pb = partial + G; double* pcn = partial + 2 * (size_t)G;
  if (threadIdx.x == 0) m.load(state);
  __syncthreads();
  bool pending = false;                        // |x|^2 of the previous iteration not yet reduced
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  for (int k = itn_done; k < itn_done + n_iters; ++k) {
    if (m.s[MS_ISTOP] != 0.0) break;
    double* r1 = r3 + (size_t)(k % 3) * n;
    double* r2 = r3 + (size_t)((k + 1) % 3) * n;
    double* t = r3 + (size_t)((k + 2) % 3) * n;
    double* w2_ = w2buf + (size_t)(k % 2) * n;
    double* w1_ = w2buf + (size_t)((k + 1) % 2) * n;
    double* v = v2buf + (size_t)(k % 2) * n;
    // ---- phase A
    const unsigned long long tA = prof ? gtime() : 0ull;
    {
      const double beta = m.s[MS_BETA], oldb = m.s[MS_OLDB];
      const bool later = k >= 1;               // (MS_ITN still lags by one here: the previous tests are pending)
      const double sc = 1.0 / beta;
      const double c = later ? beta / oldb : 0.0;
      double acc = 0.0;
      constexpr int kTW = kTileW + 2 * kHalo, kTE = (kTileH + 2 * kHalo) * kTW;       // 36 x 12 staged values
      for (int tl = bid; tl < ntiles; tl += G) {
        const int x0 = (tl % tiles_x) * kTileW, y0 = (tl / tiles_x) * kTileH;
        const bool interior = x0 >= kHalo && y0 >= kHalo && x0 + kTileW + kHalo <= w && y0 + kTileH + kHalo <= h;
        // issue every load of this tile at once: the (up to two) staged values of this thread with their masks,
        // and the thread's own coefficients -- one round trip to L2 instead of four
        double sv[2]; unsigned char sm[2]; bool in[2];
#pragma unroll
        for (int u = 0; u < 2; ++u) {
          const int e = threadIdx.x + u * 256;
          const int ry = e / kTW, rx = e % kTW;
          const int gy = y0 + ry - kHalo, gx = x0 + rx - kHalo;
          in[u] = e < kTE && (interior || (gx >= 0 && gx < w && gy >= 0 && gy < h));
          const long long j = in[u] ? (long long)gy * w + gx : 0;
          sv[u] = in[u] ? r2[j] : 0.0;
          sm[u] = (in[u] && o.rigid) ? o.rigid[j] : (unsigned char)1;
        }
        const int px = x0 + tx, py = y0 + ty;
        const bool own = px < w && py < h;
        const long long i = own ? (long long)py * w + px : 0;
        SmoothCoef cf = {};
        double Di = 0.0, pci = 0.0, r1i = 0.0;
        if (own) {
          cf = load_coef(o, px, py);
          if (o.D) Di = o.D[i];
          if (o.pc) pci = o.pc[i];
          if (later) r1i = r1[i];
        }
#pragma unroll
        for (int u = 0; u < 2; ++u) {
          const int e = threadIdx.x + u * 256;
          if (e < kTE) tile[e / kTW][e % kTW] = (in[u] && sm[u]) ? sc * sv[u] : 0.0;     // v = s*y as scipy forms it, masked by Z
        }
        __syncthreads();
        if (own) {
          const double vi = sc * r2[i];
          double r = 0.0;
          if (!o.rigid || o.rigid[i]) {
            auto Vt = [&](int yy, int xx) -> double { return tile[yy - y0 + kHalo][xx - x0 + kHalo]; };
            r = interior ? smooth_apply_c<true>(o, px, py, cf, Vt) : smooth_apply_c<false>(o, px, py, cf, Vt);
            if (o.D) r += Di * vi;
            if (o.pc) r += o.lc * (pci * vi);
          }
          r = r + o.eps * vi;
          if (later) r = r - c * r1i;
          v[i] = vi; t[i] = r;
          acc += vi * r;
        }
        __syncthreads();
      }
      acc = block_sum256(acc, sh);
      if (threadIdx.x == 0) pa[bid] = acc;
    }
    const unsigned long long tA1 = prof ? gtime() : 0ull;
    grid.sync();                                                        // barrier 1
    const unsigned long long tB0 = prof ? gtime() : 0ull;
    double alfa, ysq;
    sum_partials_all2(pa, pcn, G, sh, sh_out2, alfa, ysq);              // (|x|^2 partials are only meaningful when pending)
    if (pending) {
      if (threadIdx.x == 0) m.tests(ysq, rtol, maxiter);
      __syncthreads();
      pending = false;
      if (m.s[MS_ISTOP] != 0.0) break;                                  // the previous iteration converged; x is final
    }
    // ---- phase B
    {
      const double c2 = alfa / m.s[MS_BETA];
      double acc = 0.0;
      for (long long base = bid * 256ll + threadIdx.x; base < n; base += kBatch * stride) {
        double a[kBatch], b[kBatch];
#pragma unroll
        for (int u = 0; u < kBatch; ++u) {
          const long long i = base + u * stride;
          if (i < n) { a[u] = t[i]; b[u] = r2[i]; }
        }
#pragma unroll
        for (int u = 0; u < kBatch; ++u) {
          const long long i = base + u * stride;
          if (i < n) {
            const double y = a[u] - c2 * b[u];
            t[i] = y;
            acc += y * y;
          }
        }
      }
      acc = block_sum256(acc, sh);
      if (threadIdx.x == 0) pb[bid] = acc;
    }
    const unsigned long long tB1 = prof ? gtime() : 0ull;
    grid.sync();                                                        // barrier 2
    const unsigned long long tC0 = prof ? gtime() : 0ull;
    {
      const double bsq = sum_partials_all(pb, G, sh, &sh_out);
      if (threadIdx.x == 0) m.lanczos(alfa, bsq);
      __syncthreads();
    }
    // ---- phase C
    {
      const double oldeps = m.s[MS_OLDEPS], delta = m.s[MS_DELTA], denom = m.s[MS_DENOM], phi = m.s[MS_PHI];
      double acc = 0.0;
      for (long long base = bid * 256ll + threadIdx.x; base < n; base += kBatch * stride) {
        double a[kBatch], b[kBatch], cw[kBatch], d[kBatch];
#pragma unroll
        for (int u = 0; u < kBatch; ++u) {
          const long long i = base + u * stride;
          if (i < n) { a[u] = v[i]; b[u] = w1_[i]; cw[u] = w2_[i]; d[u] = x[i]; }
        }
#pragma unroll
        for (int u = 0; u < kBatch; ++u) {
          const long long i = base + u * stride;
          if (i < n) {
            const double wn = ((a[u] - oldeps * b[u]) - delta * cw[u]) * denom;       // the new w takes w1's place
            w1_[i] = wn;
            const double xn = d[u] + phi * wn;
            x[i] = xn;
            acc += xn * xn;
          }
        }
      }
      acc = block_sum256(acc, sh);
      if (threadIdx.x == 0) pcn[bid] = acc;
      pending = true;
    }
    if (prof) { const unsigned long long tC1 = gtime(); nsA += tA1 - tA; nsSync += (tB0 - tA1) + (tC0 - tB1); nsBC += (tB1 - tB0) + (tC1 - tC0); }
  }
  if (pending) {                                                        // uniform across the grid
    grid.sync();
    const double ysq = sum_partials_all(pcn, G, sh, &sh_out);
    if (threadIdx.x == 0) m.tests(ysq, rtol, maxiter);
    __syncthreads();
  }
  if (bid == 0 && threadIdx.x == 0) {
    m.s[MS_NS_A] += (double)nsA; m.s[MS_NS_BC] += (double)nsBC; m.s[MS_NS_SYNC] += (double)nsSync;
    m.store(state);
  }
}

// out = s * y
__global__ void __launch_bounds__(256)
k_scale(const double* __restrict__ y, long long n, double s, double* __restrict__ out) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) out[i] = s * y[i];
}

// ------------------------------------------------------------------------------- update
// NaN / inf scrub (:600-606), clip_dA (:157-183, threshold 1.0), clip to +-1 (:612); out = A + dA
__global__ void __launch_bounds__(256)
k_var_clip(const double* __restrict__ da, const double* __restrict__ A, int h, int w, Pt2 q, double mu, double B, double thr,
           double* __restrict__ out) {
  const int x = blockIdx.x * 32 + (threadIdx.x & 31), y = blockIdx.y * 8 + (threadIdx.x >> 5);
  if (x >= w || y >= h) return;
  const long long i = (long long)y * w + x;
  double d = da[i];
  if (d != d || isinf(d)) d = 0.0;
  const double vx = q.x - (double)x, vy = q.y - (double)y;
  const double n = sqrt(vx * vx + vy * vy) / mu;
  const double a = A[i];
  const double num = -(((mu * mu) * thr - ((2.0 * mu) * thr) * a * B) + (thr * (a * a)) * (B * B));
  const double dmax = num / (((thr * a) * (B * B)) + ((-(mu * mu)) * n - mu * thr) * B);
  const double dmin = num / (((thr * a) * (B * B)) + ((mu * mu) * n - mu * thr) * B);
  d = fmax(dmin, fmin(d, dmax));         // np.maximum(dA_min, np.minimum(dA, dA_max)); NaN bounds propagate like numpy's
  if (dmin != dmin || dmax != dmax) d = __longlong_as_double(0x7FF8000000000000ll);
  d = fmin(fmax(d, -1.0), 1.0);
  if (d != d) d = __longlong_as_double(0x7FF8000000000000ll);
  out[i] = a + d;
}

__device__ __forceinline__ int reflect_sym7(int i, int n) {
  if (n == 1) return 0;
  const int p = 2 * n;
  i %= p; if (i < 0) i += p;
  return i < n ? i : p - 1 - i;
}
// A_new = median_filter(A + dA, size=7)  (scipy 'reflect'); the caller then has dA = A_new - A, A = A + dA (:636,:650)
__global__ void __launch_bounds__(128)
k_median7x7(const double* __restrict__ src, int h, int w, double* __restrict__ dst) {
  const int x = blockIdx.x * 32 + (threadIdx.x & 31), y = blockIdx.y * 4 + (threadIdx.x >> 5);
  if (x >= w || y >= h) return;
  int xs[7]; long long ys[7];
#pragma unroll
  for (int k = 0; k < 7; ++k) { xs[k] = reflect_sym7(x + k - 3, w); ys[k] = (long long)reflect_sym7(y + k - 3, h) * w; }
  const double med = median_forgetful<49>([&](int k) -> double { return src[ys[k / 7] + xs[k % 7]]; });
  dst[(long long)y * w + x] = med;
}
// A <- A + (med - A)   (the reference forms da = median - A and then A + da: two roundings, kept)
__global__ void __launch_bounds__(256)
k_var_advance(double* __restrict__ A, const double* __restrict__ med, long long n) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const double a = A[i];
    const double da = med[i] - a;
    A[i] = a + da;
  }
}

static inline unsigned lin_grid(long long n) {
  long long b = (n + 255) / 256;
  if (b > kVarBlocks) b = kVarBlocks;
  return (unsigned)(b < 1 ? 1 : b);
}

}  // namespace mrf

using namespace mrf;

extern "C" {

size_t mrf_var_scratch_bytes(void) { return sizeof(double) * (2 * kVarBlocks + 8); }

int mrf_gray32(const void* rgb, int rgb_is_f64, size_t n, float* gray32, mrf_stream_t stream) {
  MRF_CHECK_ARG(rgb && gray32 && n > 0);
  const unsigned g = (unsigned)(((long long)n + 255) / 256 > 4096 ? 4096 : ((long long)n + 255) / 256);
  if (rgb_is_f64) k_gray32<double><<<g, 256, 0, S(stream)>>>((const double*)rgb, (long long)n, gray32);
  else            k_gray32<float><<<g, 256, 0, S(stream)>>>((const float*)rgb, (long long)n, gray32);
  return check_launch("mrf_gray32");
}

int mrf_edge_weights_local(const float* gray32, const uint8_t* rigidity_eroded, int h, int w, int norm_radius, double min_weight,
                           float* wx, float* wy, float* tmp_a, float* tmp_b, double* tmp_ra, double* tmp_rb,
                           mrf_stream_t stream) {
  MRF_CHECK_ARG(gray32 && wx && wy && tmp_a && tmp_b && tmp_ra && tmp_rb && h > 0 && w > 0 && norm_radius > 0 && (norm_radius & 1));
  dim3 grid(cdiv(w, 32), cdiv(h, 8));
  k_grad_sq32<<<grid, 256, 0, S(stream)>>>(gray32, h, w, tmp_a, tmp_b);
  int rc = check_launch("mrf_edge_weights_local/grad");
  if (rc) return rc;
  k_box_rows<<<grid, 256, 0, S(stream)>>>(tmp_a, tmp_b, h, w, norm_radius, tmp_ra, tmp_rb);
  rc = check_launch("mrf_edge_weights_local/rows");
  if (rc) return rc;
  k_box_cols_weights<<<grid, 256, 0, S(stream)>>>(tmp_ra, tmp_rb, tmp_a, tmp_b, h, w, norm_radius, wx, wy);
  rc = check_launch("mrf_edge_weights_local/cols");
  if (rc || !rigidity_eroded) return rc;
  const long long n = (long long)h * w;
  k_mask_weights<<<lin_grid(n), 256, 0, S(stream)>>>(wx, wy, rigidity_eroded, n, (float)min_weight);
  return check_launch("mrf_edge_weights_local/mask");
}

int mrf_var_smooth_args(const double* A, const float* wx, const float* wy, int h, int w, double afac_sq, double* arg1,
                        double* arg2, double* root1, double* root2, mrf_stream_t stream) {
  MRF_CHECK_ARG(A && wx && wy && arg1 && arg2 && root1 && root2 && h > 0 && w > 0);
  dim3 grid(cdiv(w, 32), cdiv(h, 8));
  k_smooth_args<<<grid, 256, 0, S(stream)>>>(A, wx, wy, h, w, afac_sq, arg1, arg2, root1, root2);
  return check_launch("mrf_var_smooth_args");
}

int mrf_var_smooth_weights(const double* arg1, const double* arg2, const float* wx, const float* wy, size_t n, double afac_sq,
                           double sigma1, double sigma2, const double* dev_medians2, double* p1x, double* p1y, double* p2x,
                           double* p2y, double* energies2, void* scratch, mrf_stream_t stream) {
  MRF_CHECK_ARG(arg1 && arg2 && wx && wy && p1x && p1y && p2x && p2y && energies2 && scratch && n > 0);
  const unsigned g = lin_grid((long long)n);
  k_smooth_weights<<<g, 256, 0, S(stream)>>>(arg1, arg2, wx, wy, (long long)n, afac_sq, sigma1, sigma2, dev_medians2, p1x, p1y, p2x,
                                             p2y, (double*)scratch);
  int rc = check_launch("mrf_var_smooth_weights");
  if (rc) return rc;
  k_var_sum_final<<<1, kVarBlocks, 0, S(stream)>>>((const double*)scratch, (int)g, 2, energies2);
  return check_launch("mrf_var_smooth_weights/energies");
}

int mrf_var_data_merge(const double* diag_f, const double* b_f, const double* diag_b, const double* b_b, const uint8_t* occ_f,
                       const uint8_t* occ_b, const double* A, const double* A_bwd_ref, const double* A_fwd_ref, size_t n,
                       double scale_bwd, double scale_ref, double afac_sq, double* D, double* bd, double* Iz_c, double* arg_c,
                       double* root_c, mrf_stream_t stream) {
  MRF_CHECK_ARG(diag_f && b_f && diag_b && b_b && occ_f && occ_b && A && A_bwd_ref && A_fwd_ref && D && bd && Iz_c && arg_c && root_c && n > 0);
  k_var_data_merge<<<lin_grid((long long)n), 256, 0, S(stream)>>>(diag_f, b_f, diag_b, b_b, occ_f, occ_b, A, A_bwd_ref, A_fwd_ref,
                                                                  (long long)n, scale_bwd, scale_ref, afac_sq, D, bd, Iz_c, arg_c, root_c);
  return check_launch("mrf_var_data_merge");
}

int mrf_var_consistency_weights(const double* arg_c, size_t n, double afac_sq, double eps, const double* dev_median, double* pc,
                                double* energy, void* scratch, mrf_stream_t stream) {
  MRF_CHECK_ARG(arg_c && pc && energy && scratch && n > 0);
  const unsigned g = lin_grid((long long)n);
  k_cons_weights<<<g, 256, 0, S(stream)>>>(arg_c, (long long)n, afac_sq, eps, dev_median, pc, (double*)scratch);
  int rc = check_launch("mrf_var_consistency_weights");
  if (rc) return rc;
  k_var_sum_final<<<1, kVarBlocks, 0, S(stream)>>>((const double*)scratch, (int)g, 1, energy);
  return check_launch("mrf_var_consistency_weights/energy");
}

int mrf_var_apply(const double* x, const uint8_t* rigid, const double* D, const double* p1x, const double* p1y, const double* p2x,
                  const double* p2y, const double* pc, double l1, double l2, double lc, double eps, int h, int w, double* y,
                  mrf_stream_t stream) {
  MRF_CHECK_ARG(x && y && x != y && p1x && p1y && p2x && p2y && h > 0 && w > 0);
  VarOp o;
  o.rigid = rigid; o.D = D; o.p1x = p1x; o.p1y = p1y; o.p2x = p2x; o.p2y = p2y; o.pc = pc;
  o.l1 = l1; o.l2 = l2; o.lc = lc; o.eps = eps; o.h = h; o.w = w;
  dim3 grid(cdiv(w, 32), cdiv(h, 8));
  k_var_apply<<<grid, 256, 0, S(stream)>>>(o, x, y);
  return check_launch("mrf_var_apply");
}

int mrf_var_rhs(const double* bd, const double* smooth_A, const double* pc, const double* Iz_c, const uint8_t* rigid, size_t n,
                double lc, double* b, mrf_stream_t stream) {
  MRF_CHECK_ARG(bd && smooth_A && rigid && b && n > 0 && (!pc || Iz_c));
  k_var_rhs<<<lin_grid((long long)n), 256, 0, S(stream)>>>(bd, smooth_A, pc, Iz_c, rigid, (long long)n, lc, b);
  return check_launch("mrf_var_rhs");
}

int mrf_vec_scale(const double* y, size_t n, double s, double* out, mrf_stream_t stream) {
  MRF_CHECK_ARG(y && out && n > 0);
  k_scale<<<lin_grid((long long)n), 256, 0, S(stream)>>>(y, (long long)n, s, out);
  return check_launch("mrf_vec_scale");
}

static int persistent_grid(int* blocks_out) {
  // the co-resident grid is a property of the CURRENT device: cached per device, not per process
  static int cached_by_dev[64] = {0};
  int dev = 0;
  cudaError_t e0 = cudaGetDevice(&dev);
  if (e0 != cudaSuccess) return fail((int)e0, "mrf_minres: %s", cudaGetErrorString(e0));
  int scratch_slot = 0;
  int& cached = (dev >= 0 && dev < 64) ? cached_by_dev[dev] : scratch_slot;
  if (!cached) {
    int sms = 0, per_sm = 0;
    cudaError_t e = cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    if (e == cudaSuccess) e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_minres_persistent, 256, 0);
    if (e != cudaSuccess) return fail((int)e, "mrf_minres: occupancy query: %s", cudaGetErrorString(e));
    if (per_sm < 1 || sms < 1) return fail(MRF_E_BADARG, "mrf_minres: kernel does not fit an SM");
    int cap = 2;                                // resident blocks per SM used (MRF_MINRES_BLOCKS_PER_SM overrides)
    if (const char* env = getenv("MRF_MINRES_BLOCKS_PER_SM")) { const int v = atoi(env); if (v >= 1) cap = v; }
    if (per_sm > cap) per_sm = cap;
    cached = sms * per_sm;
  }
  *blocks_out = cached;
  return 0;
}

size_t mrf_minres_scratch_bytes(int h, int w) {
  (void)h; (void)w;
  return sizeof(double) * 4 * 2048;             // per-block partials of the three reductions (grid <= 2048) + init
}

int mrf_minres_init(const double* b, size_t n, double* r3, double* w2, double* x, double* state, void* scratch,
                    mrf_stream_t stream) {
  MRF_CHECK_ARG(b && r3 && w2 && x && state && scratch && n > 0);
  double* partial = (double*)scratch;
  const unsigned g = lin_grid((long long)n);
  // r1 = r3[0], r2 = r3[1]; both w buffers start at zero
  k_minres_init<<<g, 256, 0, S(stream)>>>(b, (long long)n, r3, r3 + n, w2, w2 + n, x, partial);
  int rc = check_launch("mrf_minres_init");
  if (rc) return rc;
  k_minres_init_state<<<1, kVarBlocks, 0, S(stream)>>>(partial, (int)g, state);
  return check_launch("mrf_minres_init/state");
}

int mrf_minres_iterate(const uint8_t* rigid, const double* D, const double* p1x, const double* p1y, const double* p2x,
                       const double* p2y, const double* pc, double l1, double l2, double lc, double eps, int h, int w,
                       double* r3, double* w2, double* v2, double* x, double* state, void* scratch, int itn_done, int n_iters,
                       double rtol, long long maxiter, mrf_stream_t stream) {
  MRF_CHECK_ARG(p1x && p1y && p2x && p2y && r3 && w2 && v2 && x && state && scratch && h > 0 && w > 0 && itn_done >= 0 && n_iters > 0);
  VarOp o;
  o.rigid = rigid; o.D = D; o.p1x = p1x; o.p1y = p1y; o.p2x = p2x; o.p2y = p2y; o.pc = pc;
  o.l1 = l1; o.l2 = l2; o.lc = lc; o.eps = eps; o.h = h; o.w = w;
  int blocks = 0;
  int rc = persistent_grid(&blocks);
  if (rc) return rc;
  if (blocks > 2048) blocks = 2048;
  double* partial = (double*)scratch;
  double maxit = (double)maxiter;
  void* args[] = {&o, &r3, &w2, &v2, &x, &state, &partial, &itn_done, &n_iters, &rtol, &maxit};
  cudaError_t e = cudaLaunchCooperativeKernel((const void*)k_minres_persistent, dim3(blocks), dim3(256), args, 0, S(stream));
  if (e != cudaSuccess) return fail((int)e, "mrf_minres_iterate: cooperative launch: %s", cudaGetErrorString(e));
  return check_launch("mrf_minres_iterate");
}

int mrf_var_clip_update(const double* da, const double* A, int h, int w, const double* q_host, double mu, double B,
                        double threshold, double* A_plus_da, mrf_stream_t stream) {
  MRF_CHECK_ARG(da && A && A_plus_da && q_host && h > 0 && w > 0);
  dim3 grid(cdiv(w, 32), cdiv(h, 8));
  k_var_clip<<<grid, 256, 0, S(stream)>>>(da, A, h, w, pt2(q_host), mu, B, threshold, A_plus_da);
  return check_launch("mrf_var_clip_update");
}

int mrf_median7x7_f64(const double* src, int h, int w, double* dst, mrf_stream_t stream) {
  MRF_CHECK_ARG(src && dst && src != dst && h > 0 && w > 0);
  dim3 grid(cdiv(w, 32), cdiv(h, 4));
  k_median7x7<<<grid, 128, 0, S(stream)>>>(src, h, w, dst);
  return check_launch("mrf_median7x7_f64");
}

int mrf_var_advance(double* A, const double* median_of_A_plus_da, size_t n, mrf_stream_t stream) {
  MRF_CHECK_ARG(A && median_of_A_plus_da && n > 0);
  k_var_advance<<<lin_grid((long long)n), 256, 0, S(stream)>>>(A, median_of_A_plus_da, (long long)n);
  return check_launch("mrf_var_advance");
}

}  // extern "C"
