// Pointwise plane+parallax geometry kernels (SURVEY.md section 8a rows a1-a7 and the dense parts of
// a4).  All of them are streaming, one thread per pixel, coalesced along x; they are HBM-bound
// with a few dozen fp64 flops per pixel.  Arithmetic follows the reference operation by
// operation (compiled with -fmad=false) so results are bit-identical wherever the reference only
// uses IEEE +,-,*,/,sqrt; transcendental functions (atan2, sin, exp, I0e) agree to a few ulp.
#include "peersync.cuh"
#include <cooperative_groups.h>
#include <stdarg.h>
#include <math.h>

namespace mrf {

thread_local char g_err[512] = {0};
std::atomic<long long> g_launches{0};

int fail(int code, const char* fmt, ...) {
  va_list ap; va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
  return code;
}

int check_launch(const char* what) {
  g_launches.fetch_add(1, std::memory_order_relaxed);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return fail((int)e, "%s: %s", what, cudaGetErrorString(e));
  return 0;
}

// ------------------------------------------------------------------------------- a1 + a2
// utils/flow_homography.py:4-16 + pipeline/compute_structure.py:517-537
__global__ void __launch_bounds__(256)
k_flow_to_parallax(const float* __restrict__ u, const float* __restrict__ v, int rows, int w, int y0,
                   Mat3 Hinv, Pt2 q, float* __restrict__ u_res, float* __restrict__ v_res,
                   double* __restrict__ parallax, double* __restrict__ dists) {
  const long long n = (long long)rows * w;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n;
       i += (long long)gridDim.x * blockDim.x) {
    const int yl = (int)(i / w);
    const int x = (int)(i - (long long)yl * w);
    const int y = yl + y0;
    // x + u: int64 + fp32 -> fp64 in numpy, then .astype('float32')
    const float x2 = (float)((double)x + (double)u[i]);
    const float y2 = (float)((double)y + (double)v[i]);
    float X, Y;
    persp_f32(x2, y2, Hinv, X, Y);
    const float ur = X - (float)x;   // fp32 - fp32
    const float vr = Y - (float)y;
    if (u_res) u_res[i] = ur;
    if (v_res) v_res[i] = vr;
    if (parallax || dists) {
      const double fx = q.x - (double)x;
      const double fy = q.y - (double)y;
      const double d = sqrt(fx * fx + fy * fy);
      if (dists) dists[i] = d;
      if (parallax) {
        const double dn = fmax(d, 1e-3);
        parallax[i] = (double)ur * (fx / dn) + (double)vr * (fy / dn);
      }
    }
  }
}

// a2 on its own: pipeline/compute_structure.py:517-537 (flow2parallax) for a given residual flow
// (fp64, an fp32 residual promoted): parallax, unit vectors to the epipole [rows,w,2], distances
__global__ void __launch_bounds__(256)
k_flow2parallax(const double* __restrict__ u, const double* __restrict__ v, int rows, int w, int y0, Pt2 q,
                double* __restrict__ parallax, double* __restrict__ foe_vectors, double* __restrict__ dists) {
  const long long n = (long long)rows * w;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n;
       i += (long long)gridDim.x * blockDim.x) {
    const int yl = (int)(i / w);
    const int x = (int)(i - (long long)yl * w);
    const double fx = q.x - (double)x, fy = q.y - (double)(yl + y0);
    const double d = sqrt(fx * fx + fy * fy);
    const double dn = fmax(d, 1e-3);
    const double nx = fx / dn, ny = fy / dn;
    if (parallax) parallax[i] = u[i] * nx + v[i] * ny;
    if (foe_vectors) { foe_vectors[2 * i] = nx; foe_vectors[2 * i + 1] = ny; }
    if (dists) dists[i] = d;
  }
}

// ------------------------------------------------------------------------------- mu
// pipeline/compute_structure.py:84 -- sum of ||p-q|| over the band; fixed grid, fixed tree ->
// deterministic.  (numpy's pairwise sum rounds differently by O(1e-16) relative.)
constexpr int kRedBlocks = 1024;
constexpr int kRedThreads = 256;

__device__ __forceinline__ double block_sum(double v, double* sh) {
  for (int o = 16; o > 0; o >>= 1) v += __shfl_down_sync(0xffffffffu, v, o);
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  if (lane == 0) sh[wid] = v;
  __syncthreads();
  const int nw = blockDim.x >> 5;
  v = (threadIdx.x < nw) ? sh[threadIdx.x] : 0.0;
  if (wid == 0) for (int o = 16; o > 0; o >>= 1) v += __shfl_down_sync(0xffffffffu, v, o);
  __syncthreads();
  return v;
}

__global__ void __launch_bounds__(kRedThreads)
k_dist_sum_partial(int rows, int w, int y0, Pt2 q, double* __restrict__ partial) {
  __shared__ double sh[32];
  const long long n = (long long)rows * w;
  double acc = 0.0;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n;
       i += (long long)gridDim.x * blockDim.x) {
    const int yl = (int)(i / w);
    const int x = (int)(i - (long long)yl * w);
    const double fx = q.x - (double)x, fy = q.y - (double)(yl + y0);
    acc += sqrt(fx * fx + fy * fy);
  }
  acc = block_sum(acc, sh);
  if (threadIdx.x == 0) partial[blockIdx.x] = acc;
}

__global__ void __launch_bounds__(kRedBlocks)
k_sum_final(const double* __restrict__ partial, int n, double* __restrict__ out, double divide_by) {
  __shared__ double sh[32];
  double v = (threadIdx.x < n) ? partial[threadIdx.x] : 0.0;
  v = block_sum(v, sh);
  if (threadIdx.x == 0) out[0] = (divide_by != 0.0) ? v / divide_by : v;   // .mean() = sum / n
}

// ------------------------------------------------------------------------------- a3
// pipeline/compute_structure.py:25-33
__global__ void __launch_bounds__(256)
k_parallax_to_structure(const double* __restrict__ parallax, int rows, int w, int y0, Pt2 q, double B,
                        double mu, double* __restrict__ structure, const double* __restrict__ st, int frame) {
  if (st) { mu = st[MRF_STATE_MU + frame]; B = st[MRF_STATE_B + frame]; }
  const long long n = (long long)rows * w;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n;
       i += (long long)gridDim.x * blockDim.x) {
    const int yl = (int)(i / w);
    const int x = (int)(i - (long long)yl * w);
    const double dx = (double)x - q.x, dy = (double)(yl + y0) - q.y;
    const double dst = sqrt(dx * dx + dy * dy);
    const double p = parallax[i];
    structure[i] = p / (B * (p / mu - dst / mu));
  }
}

// ------------------------------------------------------------------------------- a6
// pipeline/structure2flow.py:14-38, :72-73; utils/flow_homography.py:20-37
__global__ void __launch_bounds__(256)
k_structure_to_flow(const double* __restrict__ structure, int rows, int w, int y0, Pt2 q, double mu,
                    Mat3 H, double b, const uint8_t* __restrict__ nonrigid,
                    const float* __restrict__ init_u, const float* __restrict__ init_v,
                    float* __restrict__ u, float* __restrict__ v, const double* __restrict__ st, int frame) {
  if (st) { mu = st[MRF_STATE_MU + frame]; b = st[MRF_STATE_B + frame]; }
  const long long n = (long long)rows * w;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n;
       i += (long long)gridDim.x * blockDim.x) {
    if (nonrigid && nonrigid[i]) { u[i] = init_u[i]; v[i] = init_v[i]; continue; }
    const int yl = (int)(i / w);
    const int x = (int)(i - (long long)yl * w);
    const int y = yl + y0;
    const double dx = (double)x - q.x, dy = (double)y - q.y;
    const double d = sqrt(dx * dx + dy * dy);
    const double nn = d / mu;
    const double bs = b * structure[i];
    const double par = (bs * nn) / (bs / mu - 1.0);
    const double dm = fmax(1e-3, d);
    const double ur = ((q.x - (double)x) / dm) * par;
    const double vr = ((q.y - (double)y) / dm) * par;
    const float x2 = (float)((double)x + ur);
    const float y2 = (float)((double)y + vr);
    float X, Y;
    persp_f32(x2, y2, H, X, Y);
    u[i] = X - (float)x;
    v[i] = Y - (float)y;
  }
}

// utils/flow_homography.py:20-37 on its own (get_full_flow): residual flow fp64 (an fp32 input
// promoted, as numpy promotes int64 + fp32) -> full flow fp32
__global__ void __launch_bounds__(256)
k_full_flow(const double* __restrict__ ur, const double* __restrict__ vr, int rows, int w, int y0, Mat3 H,
            float* __restrict__ u, float* __restrict__ v) {
  const long long n = (long long)rows * w;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n;
       i += (long long)gridDim.x * blockDim.x) {
    const int yl = (int)(i / w);
    const int x = (int)(i - (long long)yl * w);
    const int y = yl + y0;
    const float x2 = (float)((double)x + ur[i]);
    const float y2 = (float)((double)y + vr[i]);
    float X, Y;
    persp_f32(x2, y2, H, X, Y);
    u[i] = X - (float)x;
    v[i] = Y - (float)y;
  }
}

// ------------------------------------------------------------------------------- a7
// Exponentially scaled modified Bessel function I0e(x) = exp(-|x|) I0(x), fp64.
// scipy.special.ive(0, x) (the reference's call, rigidity/inference.py:54) is restated by its
// published definition: ascending series for x <= 30, Hankel asymptotic expansion above.
__device__ double i0e_f64(double x) {
  x = fabs(x);
  if (x <= 30.0) {
    const double q = 0.25 * x * x;
    double term = 1.0, sum = 1.0;
    for (int k = 1; k < 200; ++k) {
      term *= q / ((double)k * (double)k);
      sum += term;
      if (term < 1e-17 * sum) break;
    }
    return sum * exp(-x);
  }
  const double inv = 1.0 / (8.0 * x);
  double term = 1.0, sum = 1.0;
  for (int k = 1; k < 24; ++k) {
    const double o = (double)(2 * k - 1);
    term *= o * o * inv / (double)k;
    sum += term;
    if (term < 1e-17 * sum) break;
  }
  return sum / sqrt(6.283185307179586476925 * x);
}

// rigidity/inference.py:28-59 with the reference's dtypes (u_res, v_res fp32 => dist, ang_uv and
// the Bessel argument/result are fp32 quantities; the rest is fp64).
__device__ __forceinline__ double unary_rigid(float ur, float vr, int x, int y, Pt2 q, double sigma,
                                              double prior_r, double prior_n) {
  const float dist = sqrtf(ur * ur + vr * vr);
  const double ang_foe = atan2(q.y - (double)y, q.x - (double)x);
  // numpy's fp32 arctan2 is libm/SVML dependent (1-4 ulp); use the correctly rounded value
  const float ang_uv = (float)atan2((double)vr, (double)ur);
  const double ang = (double)ang_uv - ang_foe;
  const double dfl = (double)dist * sin(ang);
  const double p = 1.0 / sqrt(2.0 * 3.141592653589793 * (sigma * sigma)) *
                   exp(-(dfl * dfl) / (2.0 * (sigma * sigma)));
  const float arg = (dist * dist) / (float)((sigma * sigma) * 4.0);
  const float bes = (float)i0e_f64((double)arg);
  const double c = sqrt(2.0 * 3.141592653589793) / sigma * (double)bes;
  return prior_r * p / (prior_r * p + prior_n * c / (2.0 * 3.141592653589793));
}

__global__ void __launch_bounds__(256)
k_rigidity_unaries(const float* __restrict__ ur, const float* __restrict__ vr, int rows, int w, int y0,
                   Pt2 q, double sigma, double pr, double pn, double* __restrict__ prob) {
  const long long n = (long long)rows * w;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n;
       i += (long long)gridDim.x * blockDim.x) {
    const int yl = (int)(i / w);
    const int x = (int)(i - (long long)yl * w);
    prob[i] = unary_rigid(ur[i], vr[i], x, yl + y0, q, sigma, pr, pn);
  }
}

// pipeline/compute_structure.py:104-153
__global__ void __launch_bounds__(256)
k_rigidity_direction(const float* __restrict__ ur0, const float* __restrict__ vr0,
                     const float* __restrict__ ur1, const float* __restrict__ vr1,
                     const uint8_t* __restrict__ occ0, const uint8_t* __restrict__ occ1,
                     const double* __restrict__ p_cnn, int rows, int w, int y0, Pt2 q0, Pt2 q1,
                     double sigma, double wc, double* __restrict__ p_dir, uint8_t* __restrict__ p_thr,
                     int* __restrict__ thr_count) {
  const long long n = (long long)rows * w;
  int mine = 0;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n;
       i += (long long)gridDim.x * blockDim.x) {
    const int yl = (int)(i / w);
    const int x = (int)(i - (long long)yl * w);
    const int y = yl + y0;
    const float a0 = ur0[i], b0 = vr0[i], a1 = ur1[i], b1 = vr1[i];
    double pb = unary_rigid(a0, b0, x, y, q0, sigma, 0.5, 0.5);
    double pf = unary_rigid(a1, b1, x, y, q1, sigma, 0.5, 0.5);
    if (sqrtf(a0 * a0 + b0 * b0) < 0.1f) pb = 0.6;      // :112-113,:128 (fp32 compare)
    if (sqrtf(a1 * a1 + b1 * b1) < 0.1f) pf = 0.6;
    const bool ob = occ0 && occ0[i], of = occ1 && occ1[i];
    if (of) pf = 0.5;                                   // :135-136
    if (ob) pb = 0.5;
    double pd = (pb + pf) / 2.0;                         // :140
    if (ob) pd = 0.25 + 0.5 * pf;                        // :142
    if (of) pd = 0.25 + 0.5 * pb;                        // :144
    if (ob && of) pd = 0.5;                              // :146
    if (p_dir) p_dir[i] = pd;
    if (p_thr) {
      const bool t = (wc * p_cnn[i] + (1.0 - wc) * pd) > 0.5;                // :152-153
      p_thr[i] = t ? 1 : 0;
      mine += t ? 1 : 0;
    }
  }
  if (thr_count) {                                                          // p_thr.sum()  :164
    for (int o = 16; o > 0; o >>= 1) mine += __shfl_down_sync(0xffffffffu, mine, o);
    if ((threadIdx.x & 31) == 0 && mine) atomicAdd(thr_count, mine);
  }
}

// pipeline/compute_structure.py:172-221
__global__ void __launch_bounds__(256)
k_structure_candidates(const double* __restrict__ par0, const double* __restrict__ par1,
                       const uint8_t* __restrict__ p_thr, int rows, int w, int y0, Pt2 q0, Pt2 q1,
                       double mu0, double mu1, double B_fwd, int want, uint8_t* __restrict__ pv,
                       double* __restrict__ cand, int* __restrict__ count, const double* __restrict__ st) {
  if (st) { mu0 = st[MRF_STATE_MU]; mu1 = st[MRF_STATE_MU + 1]; if (B_fwd < 0.0) B_fwd = st[MRF_STATE_B + 1]; }
  const long long n = (long long)rows * w;
  for (long long i0 = blockIdx.x * (long long)blockDim.x; i0 < n; i0 += (long long)gridDim.x * blockDim.x) {
    const long long i = i0 + threadIdx.x;
    bool valid = false;
    double val = 0.0;
    if (i < n) {
      const double p0 = par0[i], p1 = par1[i];
      valid = (fabs(p0) > 0.1) && (fabs(p1) > 0.1) && (p_thr[i] != 0);
      if (pv) pv[i] = valid ? 1 : 0;
      if (valid && cand) {
        const int yl = (int)(i / w);
        const int x = (int)(i - (long long)yl * w);
        const int y = yl + y0;
        const double dx1 = (double)x - q1.x, dy1 = (double)y - q1.y;
        const double d1 = sqrt(dx1 * dx1 + dy1 * dy1);
        if (want == 1) {
          // parallax2normedstructure(par1, q1, 1.0, mu1)  :187
          val = p1 / (1.0 * (p1 / mu1 - d1 / mu1));
        } else {
          // :209-219; dists_normed = foe_dists/mu with foe_dists from flow2parallax (q - x)
          const double fx0 = q0.x - (double)x, fy0 = q0.y - (double)y;
          const double dn0 = sqrt(fx0 * fx0 + fy0 * fy0) / mu0;
          const double fx1 = q1.x - (double)x, fy1 = q1.y - (double)y;
          const double dn1 = sqrt(fx1 * fx1 + fy1 * fy1) / mu1;
          const double target = p1 / (B_fwd * (p1 / mu1 - dn1));
          const double templ = mu1 / mu0 * p0 / (p0 / mu0 - dn0);
          val = templ / target;
        }
      }
    }
    if (cand) {
      // warp-aggregated compaction (order is irrelevant for order statistics)
      const unsigned m = __ballot_sync(0xffffffffu, valid);
      const int lane = threadIdx.x & 31;
      int base = 0;
      if (lane == 0 && m) base = atomicAdd(count, __popc(m));
      base = __shfl_sync(0xffffffffu, base, 0);
      if (valid) cand[base + __popc(m & ((1u << lane) - 1u))] = val;
    }
  }
}

// pipeline/compute_structure.py:374-394, rigidity/inference.py:113
__global__ void __launch_bounds__(256)
k_rigidity_structure(const double* __restrict__ S0, const double* __restrict__ S1,
                     const double* __restrict__ p_dir, const double* __restrict__ p_cnn,
                     const uint8_t* __restrict__ occ0, const uint8_t* __restrict__ occ1, long long n,
                     double mu0, double mu1, double ss, double wc, double* __restrict__ un64,
                     int32_t* __restrict__ un32, const double* __restrict__ st) {
  if (st) { mu0 = st[MRF_STATE_MU]; mu1 = st[MRF_STATE_MU + 1]; }
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n;
       i += (long long)gridDim.x * blockDim.x) {
    const double ad = S0[i] * mu1 / mu0 - S1[i];
    const double t = ad / ss;
    double ps = exp(-(t * t));
    const bool bad = (occ0 && occ0[i]) || (occ1 && occ1[i]);
    const double pd = p_dir[i];
    if (bad) ps = 0.5;
    double pm = pd * ps;
    if (bad) pm = 0.25 + 0.5 * pd;
    const double pr = wc * p_cnn[i] + (1.0 - wc) * pm;
    const double u0 = 1.0 - pr, u1 = pr;
    if (un64) { un64[2 * i] = u0; un64[2 * i + 1] = u1; }
    if (un32) { un32[2 * i] = (int32_t)((-u0) * 1000.0); un32[2 * i + 1] = (int32_t)((-u1) * 1000.0); }
  }
}

__global__ void __launch_bounds__(256)
k_abs_dev(const double* __restrict__ x, size_t n, double c, double* __restrict__ out, const int* __restrict__ n_dev,
          const double* __restrict__ c_dev) {
  if (n_dev) n = (size_t)max(*n_dev, 0);
  if (c_dev) c = *c_dev;
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x)
    out[i] = fabs(x[i] - c);
}

// B_fwd and B_b of pipeline/compute_structure.py:183-223 from device-resident pieces, so that the stage needs
// no host round trip: mode 1 -> st[B+1] = B_fwd = 1.426 * MAD (:190), or 1.0 with fewer than 100 valid pixels
// (:183-184) or scale_structure outside {1} (:198); mode 2 -> st[B] = B_b = the median ratio (:221), or -B_fwd
// with fewer than 100 valid pixels (:206)
__global__ void k_scales_finalize(double* __restrict__ st, const int* __restrict__ n_valid, const double* __restrict__ med,
                                  int mode, int scale_structure) {
  const bool few = *n_valid < 100;
  if (mode == 1) st[MRF_STATE_B + 1] = (few || scale_structure != 1) ? 1.0 : 1.426 * med[0];
  else st[MRF_STATE_B] = few ? -st[MRF_STATE_B + 1] : med[0];
}

__global__ void k_set_f64(double* p, double v) { p[0] = v; }

static inline unsigned grid_for(long long n, int threads = 256) {
  long long b = (n + threads - 1) / threads;
  const long long cap = (long long)kSMs * 16;
  if (b > cap) b = cap;
  if (b < 1) b = 1;
  return (unsigned)b;
}


// ---------------------------------------------------------------------------------------------
// Fused forms of the pre-pass for the device-resident path (same per-pixel arithmetic as the
// kernels above; fewer, fatter launches).
#ifndef MRF_CAND_PX
#define MRF_CAND_PX 1024   // pixels per block of the candidates kernel (2048: 13 us more per single-triplet step, same throughput)
#endif
constexpr int kFrontBlocks = 592;   // 4 x 148
// blocks of the candidates kernel: each zeroes and flushes an 8192-bin shared histogram, so four fat blocks
// per SM at most
static inline unsigned front_cand_blocks(long long n) {
  long long b = (n + MRF_CAND_PX - 1) / MRF_CAND_PX;
  if (b > 4 * kSMs) b = 4 * kSMs;
  return (unsigned)(b < 1 ? 1 : b);
}

// one pixel, one frame: residual flow -> parallax (a1 + a2); d = ||p - q||
__device__ __forceinline__ double front_parallax_px(float u, float v, int x, int y, const Mat3& Hinv, const Pt2 q, double& d) {
  const float x2 = (float)((double)x + (double)u);
  const float y2 = (float)((double)y + (double)v);
  float X, Y;
  persp_f32(x2, y2, Hinv, X, Y);
  const float ur = X - (float)x, vr = Y - (float)y;
  const double fx = q.x - (double)x, fy = q.y - (double)y;
  d = sqrt(fx * fx + fy * fy);
  const double dn = fmax(d, 1e-3);
  return (double)ur * (fx / dn) + (double)vr * (fy / dn);
}

// one pixel: the candidate ratio template / target of compute_structure.py:209-219 (B_fwd fixed)
__device__ __forceinline__ double front_candidate_px(double p0, double p1, int x, int y, const Pt2 q0, const Pt2 q1, double mu0,
                                                     double mu1, double B_fwd) {
  // dists_normed = foe_dists/mu with foe_dists from flow2parallax (q - x)
  const double fx0 = q0.x - (double)x, fy0 = q0.y - (double)y;
  const double dn0 = sqrt(fx0 * fx0 + fy0 * fy0) / mu0;
  const double fx1 = q1.x - (double)x, fy1 = q1.y - (double)y;
  const double dn1 = sqrt(fx1 * fx1 + fy1 * fy1) / mu1;
  const double target = p1 / (B_fwd * (p1 / mu1 - dn1));
  const double templ = mu1 / mu0 * p0 / (p0 / mu0 - dn0);
  return templ / target;
}

// both frames: residual flow -> parallax (a1 + a2) and the block partial sums behind mu
__global__ void __launch_bounds__(256)
k_front_parallax(const float* __restrict__ u0, const float* __restrict__ v0, const float* __restrict__ u1,
                 const float* __restrict__ v1, int rows, int w, int y0, Mat3 Hinv0, Mat3 Hinv1, Pt2 q0, Pt2 q1,
                 double* __restrict__ par0, double* __restrict__ par1, double* __restrict__ partial) {
  __shared__ double sh[32];
  const long long n = (long long)rows * w;
  double acc0 = 0.0, acc1 = 0.0;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n;
       i += (long long)gridDim.x * blockDim.x) {
    const int yl = (int)(i / w);
    const int x = (int)(i - (long long)yl * w);
    const int y = yl + y0;
#pragma unroll
    for (int f = 0; f < 2; ++f) {
      double d;
      (f ? par1 : par0)[i] = front_parallax_px((f ? u1 : u0)[i], (f ? v1 : v0)[i], x, y, f ? Hinv1 : Hinv0, f ? q1 : q0, d);
      if (f) acc1 += d; else acc0 += d;
    }
  }
  acc0 = block_sum(acc0, sh);
  acc1 = block_sum(acc1, sh);
  if (threadIdx.x == 0) { partial[blockIdx.x] = acc0; partial[gridDim.x + blockIdx.x] = acc1; }
}

__global__ void __launch_bounds__(1024)
k_front_mu(const double* __restrict__ partial, int nblocks, double n_px, double* __restrict__ st) {
  __shared__ double sh[32];
  for (int f = 0; f < 2; ++f) {
    double v = (threadIdx.x < nblocks) ? partial[f * nblocks + threadIdx.x] : 0.0;
    v = block_sum(v, sh);
    if (threadIdx.x == 0) st[MRF_STATE_MU + f] = (n_px != 0.0) ? v / n_px : v;   // n_px = 0: leave the band's sum
  }
}

// mu of both frames from the 2 x kRedBlocks block partials, in k_front_mu's summation tree (32 warp sums of
// 32 consecutive partials each, then one warp over those), evaluated by a 256-thread block -- every block
// of the candidates kernel gets the same bits without a separate launch
__device__ __forceinline__ void front_mu_from_partials(const double* __restrict__ partial, double n_px, double* s_mu,
                                                       double* sh) {
  static_assert(kRedBlocks == 1024, "summation tree written for 1024 partials");
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  for (int f = 0; f < 2; ++f) {
    for (int g = wid; g < 32; g += 8) {
      double v = partial[f * kRedBlocks + 32 * g + lane];
      for (int o = 16; o > 0; o >>= 1) v += __shfl_down_sync(0xffffffffu, v, o);
      if (lane == 0) sh[g] = v;
    }
    __syncthreads();
    if (wid == 0) {
      double v = sh[lane];
      for (int o = 16; o > 0; o >>= 1) v += __shfl_down_sync(0xffffffffu, v, o);
      if (lane == 0) s_mu[f] = v / n_px;
    }
    __syncthreads();
  }
}

// pipeline/compute_structure.py:172-221 with B_fwd fixed (build_cost_volume.py:116-131): valid pixels,
// the compacted candidate ratios template/target, and -- fused -- the first pass of the exact median over
// them (select13.cuh).  partial != nullptr: mu is formed here from the parallax kernel's block partials
// (whole frame on one GPU) and block 0 writes it to the state; otherwise mu is read from the state.
__global__ void __launch_bounds__(256)
k_front_candidates13(const double* __restrict__ par0, const double* __restrict__ par1,
                     const uint8_t* __restrict__ p_thr, int rows, int w, int y0, Pt2 q0, Pt2 q1, double B_fwd,
                     double* __restrict__ cand, int* __restrict__ count, double* __restrict__ st,
                     const double* __restrict__ partial, double n_px, void* sel, const PeerArg pa) {
  __shared__ unsigned int sh[kS13Bins];
  __shared__ double s_mu[2];
  __shared__ double s_red[32];
  s13_zero(sh);
  if (partial) {
    front_mu_from_partials(partial, n_px, s_mu, s_red);
    if (blockIdx.x == 0 && threadIdx.x == 0) { st[MRF_STATE_MU] = s_mu[0]; st[MRF_STATE_MU + 1] = s_mu[1]; }
  } else {
    if (threadIdx.x < 2) s_mu[threadIdx.x] = st[MRF_STATE_MU + threadIdx.x];
  }
  __syncthreads();
  const double mu0 = s_mu[0], mu1 = s_mu[1];
  const long long n = (long long)rows * w;
  for (long long i0 = blockIdx.x * (long long)blockDim.x; i0 < n; i0 += (long long)gridDim.x * blockDim.x) {
    const long long i = i0 + threadIdx.x;
    bool valid = false;
    double val = 0.0;
    if (i < n) {
      const double p0 = par0[i], p1 = par1[i];
      valid = (fabs(p0) > 0.1) && (fabs(p1) > 0.1) && (p_thr[i] != 0);
      if (valid) {
        const int yl = (int)(i / w);
        const int x = (int)(i - (long long)yl * w);
        const int y = yl + y0;
        val = front_candidate_px(p0, p1, x, y, q0, q1, mu0, mu1, B_fwd);          // :209-219
      }
    }
    // warp-aggregated compaction (order is irrelevant for order statistics)
    const unsigned m = __ballot_sync(0xffffffffu, valid);
    const int lane = threadIdx.x & 31;
    int base = 0;
    if (lane == 0 && m) base = atomicAdd(count, __popc(m));
    base = __shfl_sync(0xffffffffu, base, 0);
    if (valid) cand[base + __popc(m & ((1u << lane) - 1u))] = val;
    s13_count(sh, valid, (unsigned)(order_key(val) >> s13_shift(0)));
  }
  __syncthreads();
  s13_flush(sh, s13_hist(sel, 0));                // this rank's histogram; the next kernel sums it over the ranks
}

// both frames: parallax -> normed structure (a3) and block partial min / max / NaN of the
// range-masked structure (build_cost_volume.py:154-169)
struct Box4 { int x0, x1, y0, y1, on; };
// sweep of this block over the band: parallax -> normed structure, running min / max / NaN of the range-masked
// structure; the block's result goes to part[(f * nblocks + block) * 3 ..]
__device__ __forceinline__ void front_structures_block(const double* __restrict__ par0, const double* __restrict__ par1, int rows,
                                                       int w, int y0, const Pt2 q0, const Pt2 q1, double B_fwd, double B0,
                                                       double mu0, double mu1, const uint8_t* __restrict__ zero_mask,
                                                       const Box4& b0, const Box4& b1, double* __restrict__ S0,
                                                       double* __restrict__ S1, double* __restrict__ part, unsigned block,
                                                       unsigned nblocks, double (*smin)[8], double (*smax)[8], int (*snan)[8]) {
  const long long n = (long long)rows * w;
  double lo[2] = {INFINITY, INFINITY}, hi[2] = {-INFINITY, -INFINITY};
  int nan[2] = {0, 0};
  for (long long i = block * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)nblocks * blockDim.x) {
    const int yl = (int)(i / w);
    const int x = (int)(i - (long long)yl * w);
    const int y = yl + y0;
    const bool zm = zero_mask && zero_mask[i];
#pragma unroll
    for (int f = 0; f < 2; ++f) {
      const Pt2 q = f ? q1 : q0; const double B = f ? B_fwd : B0; const double mu = f ? mu1 : mu0;
      const Box4& b = f ? b1 : b0;
      const double dx = (double)x - q.x, dy = (double)y - q.y;
      const double dst = sqrt(dx * dx + dy * dy);
      const double p = (f ? par1 : par0)[i];
      double v = p / (B * (p / mu - dst / mu));
      (f ? S1 : S0)[i] = v;
      if (zm) v = 0.0;
      if (b.on && x >= b.x0 && x < b.x1 && y >= b.y0 && y < b.y1) v = 0.0;
      if (v != v) nan[f] = 1;
      lo[f] = fmin(lo[f], v); hi[f] = fmax(hi[f], v);
    }
  }
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
#pragma unroll
  for (int f = 0; f < 2; ++f) {
    for (int o = 16; o > 0; o >>= 1) {
      lo[f] = fmin(lo[f], __shfl_down_sync(0xffffffffu, lo[f], o));
      hi[f] = fmax(hi[f], __shfl_down_sync(0xffffffffu, hi[f], o));
      nan[f] |= __shfl_down_sync(0xffffffffu, nan[f], o);
    }
    if (lane == 0) { smin[f][wid] = lo[f]; smax[f][wid] = hi[f]; snan[f][wid] = nan[f]; }
  }
  __syncthreads();
  if (threadIdx.x < 2) {
    const int f = threadIdx.x;
    double a = smin[f][0], b = smax[f][0]; int c = snan[f][0];
    for (int k = 1; k < 8; ++k) { a = fmin(a, smin[f][k]); b = fmax(b, smax[f][k]); c |= snan[f][k]; }
    double* o = part + ((size_t)f * nblocks + block) * 3;
    o[0] = a; o[1] = b; o[2] = (double)c;
  }
}

__global__ void __launch_bounds__(256)
k_front_structures(const double* __restrict__ par0, const double* __restrict__ par1, int rows, int w, int y0, Pt2 q0,
                   Pt2 q1, double B_fwd, const uint8_t* __restrict__ zero_mask, Box4 b0, Box4 b1,
                   double* __restrict__ S0, double* __restrict__ S1, double* __restrict__ st,
                   double* __restrict__ part, const void* __restrict__ sel,
                   const unsigned long long* __restrict__ gathered, int world, const PeerArg pa) {
  __shared__ double smin[2][8], smax[2][8];
  __shared__ int snan[2][8];
  int gstride = 3;
  if (pa.world > 0) {          // row bands over peer memory: wait for every rank's successor words
    peer_phase_wait(pa, kS13Passes);
    gathered = &pa.p[pa.rank]->gather[0][0][0];
    world = pa.world; gstride = 8;
  }
  // B_b = median(template / target)  compute_structure.py:221 -- the same bits in every block
  const double mu0 = st[MRF_STATE_MU], mu1 = st[MRF_STATE_MU + 1], B0 = s13_median(sel, gathered, world, gstride);
  if (blockIdx.x == 0 && threadIdx.x == 0) { st[MRF_STATE_B] = B0; st[MRF_STATE_B + 1] = B_fwd; }
  front_structures_block(par0, par1, rows, w, y0, q0, q1, B_fwd, B0, mu0, mu1, zero_mask, b0, b1, S0, S1, part, blockIdx.x,
                         gridDim.x, smin, smax, snan);
}

// final (NaN-ignoring) min / max of both frames and the NaN flags -> state; row bands all-reduce these
// (MIN, MAX, MAX) before mrf_label_range_dev
__global__ void __launch_bounds__(256)
k_front_minmax(const double* __restrict__ part, int nblocks, double* __restrict__ st, const PeerArg pa) {
  __shared__ double smin[8], smax[8];
  __shared__ int snan[8];
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  for (int f = 0; f < 2; ++f) {
    double lo = INFINITY, hi = -INFINITY; int nan = 0;
    for (int i = threadIdx.x; i < nblocks; i += blockDim.x) {
      const double* o = part + ((size_t)f * nblocks + i) * 3;
      lo = fmin(lo, o[0]); hi = fmax(hi, o[1]); if (o[2] != 0.0) nan = 1;
    }
    for (int o = 16; o > 0; o >>= 1) {
      lo = fmin(lo, __shfl_down_sync(0xffffffffu, lo, o));
      hi = fmax(hi, __shfl_down_sync(0xffffffffu, hi, o));
      nan |= __shfl_down_sync(0xffffffffu, nan, o);
    }
    if (lane == 0) { smin[wid] = lo; smax[wid] = hi; snan[wid] = nan; }
    __syncthreads();
    if (threadIdx.x == 0) {
      for (int k = 1; k < 8; ++k) { lo = fmin(lo, smin[k]); hi = fmax(hi, smax[k]); nan |= snan[k]; }
      st[MRF_STATE_MINMAX + 2 * f] = lo;
      st[MRF_STATE_MINMAX + 2 * f + 1] = hi;
      st[MRF_STATE_NANFLAG + f] = (double)nan;
    }
    __syncthreads();
  }
  // row bands over peer memory: this (single-block) kernel sends the band's six range words to every rank
  if (pa.world > 0) peer_phase_done(pa, kS13Passes + 1, reinterpret_cast<const unsigned long long*>(st + MRF_STATE_MINMAX), 6, 1);
}

// final min/max of both frames -> state, then np.linspace(1.5*min, 1.5*max, L) (see k_label_range); one block
__device__ __forceinline__ void front_labels_block(const double* __restrict__ part, int nblocks, double* __restrict__ st, int L,
                                                   double* smin, double* smax, int* snan) {
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  for (int f = 0; f < 2; ++f) {
    double lo = INFINITY, hi = -INFINITY; int nan = 0;
    for (int i = threadIdx.x; i < nblocks; i += blockDim.x) {
      const double* o = part + ((size_t)f * nblocks + i) * 3;
      lo = fmin(lo, o[0]); hi = fmax(hi, o[1]); if (o[2] != 0.0) nan = 1;
    }
    for (int o = 16; o > 0; o >>= 1) {
      lo = fmin(lo, __shfl_down_sync(0xffffffffu, lo, o));
      hi = fmax(hi, __shfl_down_sync(0xffffffffu, hi, o));
      nan |= __shfl_down_sync(0xffffffffu, nan, o);
    }
    if (lane == 0) { smin[wid] = lo; smax[wid] = hi; snan[wid] = nan; }
    __syncthreads();
    if (threadIdx.x == 0) {
      for (int k = 1; k < 8; ++k) { lo = fmin(lo, smin[k]); hi = fmax(hi, smax[k]); nan |= snan[k]; }
      const double qnan = __longlong_as_double(0x7FF8000000000000ll);
      st[MRF_STATE_MINMAX + 2 * f] = nan ? qnan : lo;
      st[MRF_STATE_MINMAX + 2 * f + 1] = nan ? qnan : hi;
    }
    __syncthreads();
  }
  const int i = threadIdx.x;
  if (i >= L) return;
  const double lo0 = st[MRF_STATE_MINMAX], hi0 = st[MRF_STATE_MINMAX + 1];
  const double lo1 = st[MRF_STATE_MINMAX + 2], hi1 = st[MRF_STATE_MINMAX + 3];
  const bool nan = (lo0 != lo0) || (lo1 != lo1) || (hi0 != hi0) || (hi1 != hi1);
  const double qnan = __longlong_as_double(0x7FF8000000000000ll);
  const double start = nan ? qnan : 1.5 * fmin(lo0, lo1);
  const double stop = nan ? qnan : 1.5 * fmax(hi0, hi1);
  const double div = (double)(L - 1);
  const double delta = stop - start;
  const double step = delta / div;
  double yv = (step == 0.0) ? ((double)i / div) * delta : (double)i * step;
  yv = yv + start;
  if (i == L - 1) yv = stop;
  st[MRF_STATE_LABELS + i] = yv;
}

__global__ void __launch_bounds__(256)
k_front_labels(const double* __restrict__ part, int nblocks, double* __restrict__ st, int L) {
  __shared__ double smin[8], smax[8];
  __shared__ int snan[8];
  front_labels_block(part, nblocks, st, L, smin, smax, snan);
}

// ---------------------------------------------------------------------------------------------
// The whole pre-pass of a frame on one GPU as ONE persistent cooperative kernel (grid barriers instead of
// launches): parallax of both frames and the sums behind mu | mu, valid pixels, compacted candidate ratios and the
// first radix pass of their exact median | further radix passes only until the class of the median holds at most
// kFusedCap keys | gather of that class (and the smallest key above it) | rank selection of the two middle order
// statistics in shared memory, B_b, normed structures, range min / max | label set.  Same per-pixel arithmetic and
// the same summation tree for mu as the stepwise kernels (and as mrf_front_mu_whole for row bands), so the state it
// leaves is bit-identical to theirs.
constexpr int kFusedCap = 1024;
struct FrontFusedArgs {
  const float *u0, *v0, *u1, *v1;
  const uint8_t *rigid, *zero_mask;
  int rows, w, L;
  Mat3 Hinv0, Hinv1;
  Pt2 q0, q1;
  double B_fwd;
  Box4 b0, b1;
  double *par0, *par1, *S0, *S1, *cand;
  int* count;
  double* st;
  double* partial;            // 4096 doubles: 2 x kRedBlocks distance sums, later 6 x gridDim range partials
  void* sel;                  // select13 scratch (zeroed); st[7] = gather counter
  unsigned long long* gbuf;   // kFusedCap keys
};

__global__ void __launch_bounds__(256, 3)
k_front_fused(const __grid_constant__ FrontFusedArgs a) {
  namespace cg = cooperative_groups;
  cg::grid_group grid = cg::this_grid();
  __shared__ unsigned int sh[kS13Bins];
  __shared__ unsigned long long s_tmp[16];
  __shared__ unsigned long long s_keys[kFusedCap];
  __shared__ double s_mu[2];
  __shared__ double s_red[32];
  __shared__ double smin[2][8], smax[2][8];
  __shared__ int snan[2][8];
  const unsigned G = gridDim.x;
  const long long n = (long long)a.rows * a.w;
  unsigned long long* st13 = s13_st(a.sel);

  // ---- phase 1: parallax, in the 1024 virtual blocks of k_front_parallax (same partial sums, same bits of mu)
  for (unsigned v = blockIdx.x; v < (unsigned)kRedBlocks; v += G) {
    double acc0 = 0.0, acc1 = 0.0;
    for (long long i = v * (long long)kRedThreads + threadIdx.x; i < n; i += (long long)kRedBlocks * kRedThreads) {
      const int y = (int)(i / a.w);
      const int x = (int)(i - (long long)y * a.w);
      double d0, d1;
      a.par0[i] = front_parallax_px(a.u0[i], a.v0[i], x, y, a.Hinv0, a.q0, d0);
      a.par1[i] = front_parallax_px(a.u1[i], a.v1[i], x, y, a.Hinv1, a.q1, d1);
      acc0 += d0; acc1 += d1;
    }
    acc0 = block_sum(acc0, s_red);
    acc1 = block_sum(acc1, s_red);
    if (threadIdx.x == 0) { a.partial[v] = acc0; a.partial[kRedBlocks + v] = acc1; }
  }
  s13_zero(sh);
  grid.sync();

  // ---- phase 2: mu, candidates, first radix pass
  front_mu_from_partials(a.partial, (double)n, s_mu, s_red);
  if (blockIdx.x == 0 && threadIdx.x == 0) { a.st[MRF_STATE_MU] = s_mu[0]; a.st[MRF_STATE_MU + 1] = s_mu[1]; }
  const double mu0 = s_mu[0], mu1 = s_mu[1];
  unsigned int my_nan = 0;
  for (long long i0 = blockIdx.x * (long long)blockDim.x; i0 < n; i0 += (long long)G * blockDim.x) {
    const long long i = i0 + threadIdx.x;
    bool valid = false;
    double val = 0.0;
    if (i < n) {
      const double p0 = a.par0[i], p1 = a.par1[i];
      valid = (fabs(p0) > 0.1) && (fabs(p1) > 0.1) && (a.rigid[i] != 0);
      if (valid) {
        const int y = (int)(i / a.w);
        const int x = (int)(i - (long long)y * a.w);
        val = front_candidate_px(p0, p1, x, y, a.q0, a.q1, mu0, mu1, a.B_fwd);
        if (val != val) ++my_nan;
      }
    }
    const unsigned m = __ballot_sync(0xffffffffu, valid);
    const int lane = threadIdx.x & 31;
    int base = 0;
    if (lane == 0 && m) base = atomicAdd(a.count, __popc(m));
    base = __shfl_sync(0xffffffffu, base, 0);
    if (valid) a.cand[base + __popc(m & ((1u << lane) - 1u))] = val;
    s13_count(sh, valid, (unsigned)(order_key(val) >> s13_shift(0)));
  }
  for (int o = 16; o > 0; o >>= 1) my_nan += __shfl_down_sync(0xffffffffu, my_nan, o);
  if ((threadIdx.x & 31) == 0 && my_nan) atomicAdd(&st13[3], (unsigned long long)my_nan);
  __syncthreads();
  s13_flush(sh, s13_hist(a.sel, 0));
  grid.sync();

  // ---- radix passes until the median's class is small
  const size_t nc = (size_t)max(*a.count, 0);
  const size_t stride = (size_t)G * blockDim.x;
  const size_t warp_first = blockIdx.x * (size_t)blockDim.x + threadIdx.x - (threadIdx.x & 31);
  unsigned long long prefix = 0, k_rem = 0, n_total = 0, class_count = 0;
  int pass = 0;
  while (true) {
    s13_resolve_digit(s13_hist(a.sel, pass), pass, prefix, k_rem, n_total, class_count, s_tmp);
    ++pass;
    if (class_count <= (unsigned long long)kFusedCap || pass == kS13Passes) break;
    s13_zero(sh);
    __syncthreads();
    const int shift = s13_shift(pass);
    const unsigned long long himask = s13_himask(pass);
    for (size_t base = warp_first; base < nc; base += 4 * stride) {
      unsigned long long kk[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const size_t i = base + u * stride + (threadIdx.x & 31);
        kk[u] = (i < nc) ? order_key(a.cand[i]) : 0ull;
      }
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const size_t i = base + u * stride + (threadIdx.x & 31);
        s13_count(sh, (i < nc) && ((kk[u] & himask) == prefix), (unsigned)((kk[u] >> shift) & (kS13Bins - 1)));
      }
    }
    __syncthreads();
    s13_flush(sh, s13_hist(a.sel, pass));
    grid.sync();
  }

  // ---- gather the class of the median (if it fits) and the smallest key above the class
  {
    const unsigned long long himask = s13_himask(pass);       // pass == 5: every bit
    const bool fits = class_count <= (unsigned long long)kFusedCap;
    unsigned long long min_above = 0xFFFFFFFFFFFFFFFFull;
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < nc; i += stride) {
      const unsigned long long k = order_key(a.cand[i]);
      const unsigned long long kp = k & himask;
      if (kp == prefix) {
        if (fits) a.gbuf[atomicAdd(&st13[7], 1ull)] = k;
      } else if (kp > prefix && k < min_above) {
        min_above = k;
      }
    }
    for (int o = 16; o > 0; o >>= 1) {
      const unsigned long long m = __shfl_down_sync(0xffffffffu, min_above, o);
      min_above = m < min_above ? m : min_above;
    }
    if ((threadIdx.x & 31) == 0 && min_above != 0xFFFFFFFFFFFFFFFFull) atomicMax(&st13[5], ~min_above);   // inverted: 0 = none
  }
  grid.sync();

  // ---- the two middle order statistics (every block, redundantly), B_b, structures, range partials
  {
    const unsigned long long none = 0xFFFFFFFFFFFFFFFFull;
    const unsigned long long above = ~st13[5];                // smallest key above the class (all ones: none)
    if (threadIdx.x == 0) { s_tmp[11] = none; s_tmp[12] = none; }
    const int c = (class_count <= (unsigned long long)kFusedCap) ? (int)class_count : 0;
    for (int j = threadIdx.x; j < c; j += blockDim.x) s_keys[j] = a.gbuf[j];
    __syncthreads();
    if (c > 0) {
      for (int j = threadIdx.x; j < c; j += blockDim.x) {
        const unsigned long long kj = s_keys[j];
        unsigned long long r = 0;
        for (int t = 0; t < c; ++t) { const unsigned long long kt = s_keys[t]; r += (kt < kj) || (kt == kj && t < j); }
        if (r == k_rem) s_tmp[11] = kj;
        if (r == k_rem + 1) s_tmp[12] = kj;
      }
    } else if (threadIdx.x == 0 && class_count > 0) {         // more than kFusedCap copies of one key
      s_tmp[11] = prefix;
      if (k_rem + 1 < class_count) s_tmp[12] = prefix;
    }
    __syncthreads();
    const unsigned long long ka = s_tmp[11];
    const unsigned long long kb = (s_tmp[12] != none) ? s_tmp[12] : above;
    const double va = key_value(ka), vb = key_value(kb);
    double B0 = (n_total % 2 == 1) ? va : (0.0 + va + vb) / 2.0;     // numpy.median
    if (n_total == 0 || st13[3] != 0) B0 = __longlong_as_double(0x7FF8000000000000ll);
    if (blockIdx.x == 0 && threadIdx.x == 0) {
      a.st[MRF_STATE_B] = B0; a.st[MRF_STATE_B + 1] = a.B_fwd;
      st13[0] = ka; st13[6] = n_total;
    }
    front_structures_block(a.par0, a.par1, a.rows, a.w, 0, a.q0, a.q1, a.B_fwd, B0, mu0, mu1, a.zero_mask, a.b0, a.b1, a.S0, a.S1,
                           a.partial, blockIdx.x, G, smin, smax, snan);
  }
  grid.sync();
  if (blockIdx.x == 0) front_labels_block(a.partial, (int)G, a.st, a.L, smin[0], smax[0], snan[0]);
}

}  // namespace mrf

using namespace mrf;

extern "C" {

int mrf_version(void) { return 100; }
const char* mrf_last_error(void) { return g_err; }
long long mrf_launch_count(void) { return g_launches.load(); }
void mrf_reset_launch_count(void) { g_launches.store(0); }

int mrf_flow_to_parallax(const float* u, const float* v, int rows, int w, int y0, const double* Hinv_host,
                         const double* q_host, float* u_res, float* v_res, double* parallax, double* dists,
                         mrf_stream_t stream) {
  MRF_CHECK_ARG(u && v && rows > 0 && w > 0 && Hinv_host && q_host);
  const long long n = (long long)rows * w;
  k_flow_to_parallax<<<grid_for(n), 256, 0, S(stream)>>>(u, v, rows, w, y0, mat3(Hinv_host), pt2(q_host),
                                                        u_res, v_res, parallax, dists);
  return check_launch("mrf_flow_to_parallax");
}

int mrf_flow2parallax(const double* u, const double* v, int rows, int w, int y0, const double* q_host,
                      double* parallax, double* foe_vectors, double* dists, mrf_stream_t stream) {
  MRF_CHECK_ARG(u && v && rows > 0 && w > 0 && q_host && (parallax || foe_vectors || dists));
  k_flow2parallax<<<grid_for((long long)rows * w), 256, 0, S(stream)>>>(u, v, rows, w, y0, pt2(q_host), parallax,
                                                                       foe_vectors, dists);
  return check_launch("mrf_flow2parallax");
}

size_t mrf_reduce_scratch_bytes(void) { return sizeof(double) * kRedBlocks * 2; }

int mrf_epipole_dist_sum(int rows, int w, int y0, const double* q_host, double* out_sum, void* scratch,
                         mrf_stream_t stream) {
  MRF_CHECK_ARG(rows > 0 && w > 0 && q_host && out_sum && scratch);
  k_dist_sum_partial<<<kRedBlocks, kRedThreads, 0, S(stream)>>>(rows, w, y0, pt2(q_host), (double*)scratch);
  int rc = check_launch("mrf_epipole_dist_sum/partial");
  if (rc) return rc;
  k_sum_final<<<1, kRedBlocks, 0, S(stream)>>>((const double*)scratch, kRedBlocks, out_sum, 0.0);
  return check_launch("mrf_epipole_dist_sum/final");
}

int mrf_parallax_to_structure(const double* parallax, int rows, int w, int y0, const double* q_host, double B,
                              double mu, double* structure, mrf_stream_t stream) {
  MRF_CHECK_ARG(parallax && structure && rows > 0 && w > 0 && q_host);
  k_parallax_to_structure<<<grid_for((long long)rows * w), 256, 0, S(stream)>>>(parallax, rows, w, y0,
                                                                               pt2(q_host), B, mu, structure, nullptr, 0);
  return check_launch("mrf_parallax_to_structure");
}

int mrf_structure_to_flow(const double* structure, int rows, int w, int y0, const double* q_host, double mu,
                          const double* H_host, double b, const uint8_t* nonrigid, const float* init_u,
                          const float* init_v, float* u, float* v, const double* dev_state, int state_frame,
                          mrf_stream_t stream) {
  MRF_CHECK_ARG(structure && u && v && rows > 0 && w > 0 && q_host && H_host);
  MRF_CHECK_ARG(!nonrigid || (init_u && init_v));
  MRF_CHECK_ARG(!dev_state || state_frame == 0 || state_frame == 1);
  k_structure_to_flow<<<grid_for((long long)rows * w), 256, 0, S(stream)>>>(
      structure, rows, w, y0, pt2(q_host), mu, mat3(H_host), b, nonrigid, init_u, init_v, u, v, dev_state,
      state_frame);
  return check_launch("mrf_structure_to_flow");
}

int mrf_full_flow(const double* u_res, const double* v_res, int rows, int w, int y0, const double* H_host,
                  float* u, float* v, mrf_stream_t stream) {
  MRF_CHECK_ARG(u_res && v_res && u && v && rows > 0 && w > 0 && H_host);
  k_full_flow<<<grid_for((long long)rows * w), 256, 0, S(stream)>>>(u_res, v_res, rows, w, y0, mat3(H_host), u, v);
  return check_launch("mrf_full_flow");
}

int mrf_rigidity_unaries(const float* u_res, const float* v_res, int rows, int w, int y0, const double* q_host,
                         double sigma, double prior_rigid, double prior_nonrigid, double* prob,
                         mrf_stream_t stream) {
  MRF_CHECK_ARG(u_res && v_res && prob && rows > 0 && w > 0 && q_host && sigma > 0);
  k_rigidity_unaries<<<grid_for((long long)rows * w), 256, 0, S(stream)>>>(
      u_res, v_res, rows, w, y0, pt2(q_host), sigma, prior_rigid, prior_nonrigid, prob);
  return check_launch("mrf_rigidity_unaries");
}

int mrf_rigidity_direction(const float* ures0, const float* vres0, const float* ures1, const float* vres1,
                           const uint8_t* occ0, const uint8_t* occ1, const double* p_cnn, int rows, int w,
                           int y0, const double* q0_host, const double* q1_host, double sigma_direction,
                           double weight_cnn, double* p_dir, uint8_t* p_thr, int* thr_count,
                           mrf_stream_t stream) {
  MRF_CHECK_ARG(ures0 && vres0 && ures1 && vres1 && rows > 0 && w > 0 && q0_host && q1_host);
  MRF_CHECK_ARG(!p_thr || p_cnn);
  MRF_CHECK_ARG(!thr_count || p_thr);
  if (thr_count) {
    cudaError_t e = cudaMemsetAsync(thr_count, 0, sizeof(int), S(stream));
    if (e != cudaSuccess) return fail((int)e, "mrf_rigidity_direction: %s", cudaGetErrorString(e));
  }
  k_rigidity_direction<<<grid_for((long long)rows * w), 256, 0, S(stream)>>>(
      ures0, vres0, ures1, vres1, occ0, occ1, p_cnn, rows, w, y0, pt2(q0_host), pt2(q1_host),
      sigma_direction, weight_cnn, p_dir, p_thr, thr_count);
  return check_launch("mrf_rigidity_direction");
}

int mrf_structure_candidates(const double* par0, const double* par1, const uint8_t* p_thr, int rows, int w,
                             int y0, const double* q0_host, const double* q1_host, double mu0, double mu1,
                             double B_fwd, int want, uint8_t* pv, double* cand, int* count,
                             mrf_stream_t stream) {
  MRF_CHECK_ARG(par0 && par1 && p_thr && rows > 0 && w > 0 && q0_host && q1_host);
  MRF_CHECK_ARG(!cand || (count && (want == 1 || want == 2)));
  if (count) {
    cudaError_t e = cudaMemsetAsync(count, 0, sizeof(int), S(stream));
    if (e != cudaSuccess) return fail((int)e, "mrf_structure_candidates: %s", cudaGetErrorString(e));
  }
  k_structure_candidates<<<grid_for((long long)rows * w), 256, 0, S(stream)>>>(
      par0, par1, p_thr, rows, w, y0, pt2(q0_host), pt2(q1_host), mu0, mu1, B_fwd, want, pv, cand, count, nullptr);
  return check_launch("mrf_structure_candidates");
}

int mrf_rigidity_structure(const double* S0, const double* S1, const double* p_dir, const double* p_cnn,
                           const uint8_t* occ0, const uint8_t* occ1, int rows, int w, double mu0, double mu1,
                           double sigma_structure, double weight_cnn, double* unaries_f64,
                           int32_t* unaries_i32, mrf_stream_t stream) {
  MRF_CHECK_ARG(S0 && S1 && p_dir && p_cnn && rows > 0 && w > 0 && (unaries_f64 || unaries_i32));
  const long long n = (long long)rows * w;
  k_rigidity_structure<<<grid_for(n), 256, 0, S(stream)>>>(S0, S1, p_dir, p_cnn, occ0, occ1, n, mu0, mu1,
                                                          sigma_structure, weight_cnn, unaries_f64,
                                                          unaries_i32, nullptr);
  return check_launch("mrf_rigidity_structure");
}

// ---- the same stage pieces with the whole-frame scalars (mu, B) read from the device state instead of the host:
// pipeline/compute_structure.py:172-246,374-394 without a host round trip (mrflow_b200.planeparallax.structure_pass)
int mrf_structure_candidates_dev(const double* par0, const double* par1, const uint8_t* p_thr, int rows, int w, int y0,
                                 const double* q0_host, const double* q1_host, const double* dev_state, int want,
                                 uint8_t* pv, double* cand, int* count, mrf_stream_t stream) {
  MRF_CHECK_ARG(par0 && par1 && p_thr && rows > 0 && w > 0 && q0_host && q1_host && dev_state && cand && count);
  MRF_CHECK_ARG(want == 1 || want == 2);
  cudaError_t e = cudaMemsetAsync(count, 0, sizeof(int), S(stream));
  if (e != cudaSuccess) return fail((int)e, "mrf_structure_candidates_dev: %s", cudaGetErrorString(e));
  // B_fwd < 0 asks the kernel to read it (like mu) from the state
  k_structure_candidates<<<grid_for((long long)rows * w), 256, 0, S(stream)>>>(
      par0, par1, p_thr, rows, w, y0, pt2(q0_host), pt2(q1_host), 0.0, 0.0, -1.0, want, pv, cand, count, dev_state);
  return check_launch("mrf_structure_candidates_dev");
}
int mrf_abs_deviation_dev(const double* x, const int* n_dev, size_t capacity, const double* c_dev, double* out,
                          mrf_stream_t stream) {
  MRF_CHECK_ARG(x && n_dev && c_dev && out && capacity > 0);
  k_abs_dev<<<grid_for((long long)capacity), 256, 0, S(stream)>>>(x, capacity, 0.0, out, n_dev, c_dev);
  return check_launch("mrf_abs_deviation_dev");
}
int mrf_scales_finalize(double* dev_state, const int* n_valid_dev, const double* median_dev, int mode, int scale_structure,
                        mrf_stream_t stream) {
  MRF_CHECK_ARG(dev_state && n_valid_dev && median_dev && (mode == 1 || mode == 2));
  k_scales_finalize<<<1, 1, 0, S(stream)>>>(dev_state, n_valid_dev, median_dev, mode, scale_structure);
  return check_launch("mrf_scales_finalize");
}
int mrf_parallax_to_structure_dev(const double* parallax, int rows, int w, int y0, const double* q_host,
                                  const double* dev_state, int frame, double* structure, mrf_stream_t stream) {
  MRF_CHECK_ARG(parallax && structure && rows > 0 && w > 0 && q_host && dev_state && (frame == 0 || frame == 1));
  k_parallax_to_structure<<<grid_for((long long)rows * w), 256, 0, S(stream)>>>(parallax, rows, w, y0, pt2(q_host), 0.0, 0.0,
                                                                               structure, dev_state, frame);
  return check_launch("mrf_parallax_to_structure_dev");
}
int mrf_rigidity_structure_dev(const double* S0, const double* S1, const double* p_dir, const double* p_cnn,
                               const uint8_t* occ0, const uint8_t* occ1, int rows, int w, const double* dev_state,
                               double sigma_structure, double weight_cnn, double* unaries_f64, int32_t* unaries_i32,
                               mrf_stream_t stream) {
  MRF_CHECK_ARG(S0 && S1 && p_dir && p_cnn && rows > 0 && w > 0 && dev_state && (unaries_f64 || unaries_i32));
  const long long n = (long long)rows * w;
  k_rigidity_structure<<<grid_for(n), 256, 0, S(stream)>>>(S0, S1, p_dir, p_cnn, occ0, occ1, n, 0.0, 0.0, sigma_structure,
                                                          weight_cnn, unaries_f64, unaries_i32, dev_state);
  return check_launch("mrf_rigidity_structure_dev");
}

int mrf_abs_deviation_f64(const double* x, size_t n, double c, double* out, mrf_stream_t stream) {
  MRF_CHECK_ARG(x && out);
  if (n == 0) return 0;
  k_abs_dev<<<grid_for((long long)n), 256, 0, S(stream)>>>(x, n, c, out, nullptr, nullptr);
  return check_launch("mrf_abs_deviation_f64");
}


// ---------------------------------------------------------------------------------------------
// The pre-pass of build_cost_volume with every whole-frame scalar kept on the device
// (pipeline/build_cost_volume.py:78-192 with the live conventions of compute_structure.py): no host
// synchronisation between the flows and the label set, so the whole step can be captured in a CUDA
// graph.  Launch sequence only; the arithmetic is that of the kernels above.
// co-resident blocks of the fused pre-pass (cooperative launch): occupancy x SM count of the CURRENT device
static int front_fused_grid(int* blocks_out) {
  static int cached[64] = {0};
  int dev = 0;
  cudaError_t e = cudaGetDevice(&dev);
  if (e != cudaSuccess) return fail((int)e, "front: %s", cudaGetErrorString(e));
  if (dev < 0 || dev >= 64 || cached[dev] == 0) {
    int per_sm = 0, sms = 0;
    e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_front_fused, 256, 0);
    if (e == cudaSuccess) e = cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    if (e != cudaSuccess) return fail((int)e, "front: occupancy query: %s", cudaGetErrorString(e));
    int b = per_sm * sms;
    if (b < 1) return fail(MRF_E_SCRATCH, "front: the fused pre-pass kernel cannot be resident");
    if (b > 2 * sms) b = 2 * sms;          // two blocks per SM: the grid barriers cost more than extra blocks give
    if (dev >= 0 && dev < 64) cached[dev] = b;
    *blocks_out = b;
    return 0;
  }
  *blocks_out = cached[dev];
  return 0;
}

static size_t front_sel_bytes() { return (mrf_select_scratch_bytes(0) + 255) / 256 * 256; }
size_t mrf_front_scratch_bytes(void) { return sizeof(double) * 4096 + front_sel_bytes() + sizeof(unsigned long long) * kFusedCap + 64; }

int mrf_cost_volume_front(const mrf_front_params* p, const float* u0, const float* v0, const float* u1,
                          const float* v1, const uint8_t* rigid_mask, const uint8_t* zero_mask,
                          double* par0, double* par1, double* S0, double* S1, double* cand, int* count,
                          double* dev_state, void* scratch, mrf_stream_t stream) {
  MRF_CHECK_ARG(p && u0 && v0 && u1 && v1 && rigid_mask && par0 && par1 && S0 && S1 && cand && count && dev_state && scratch);
  MRF_CHECK_ARG(p->rows > 0 && p->w > 0 && p->y0 == 0 && p->frame_h == p->rows);   /* whole frame on one GPU */
  MRF_CHECK_ARG(p->n_labels >= 2 && p->n_labels <= 256);
  const long long n = (long long)p->rows * p->w;
  long long fb = (n + 255) / 256;
  if (fb > kFrontBlocks) fb = kFrontBlocks;
  double* partial = (double*)scratch;                                    // 2 * kRedBlocks doubles, then 6 * fb
  void* sel = (char*)scratch + 4096 * sizeof(double);
  cudaError_t e = cudaMemsetAsync(sel, 0, kS13Bytes, S(stream));
  if (e == cudaSuccess) e = cudaMemsetAsync(count, 0, sizeof(int), S(stream));
  if (e != cudaSuccess) return fail((int)e, "front: %s", cudaGetErrorString(e));
  Box4 bx[2];
  for (int f = 0; f < 2; ++f) {
    bx[f].x0 = p->box[f][0]; bx[f].x1 = p->box[f][1]; bx[f].y0 = p->box[f][2]; bx[f].y1 = p->box[f][3];
    bx[f].on = p->have_box[f];
  }
  if (p->fused) {
    // one persistent cooperative kernel (grid barriers instead of nine launches)
    int blocks = 0;
    int rc = front_fused_grid(&blocks);
    if (rc) return rc;
    FrontFusedArgs fa;
    fa.u0 = u0; fa.v0 = v0; fa.u1 = u1; fa.v1 = v1; fa.rigid = rigid_mask; fa.zero_mask = zero_mask;
    fa.rows = p->rows; fa.w = p->w; fa.L = p->n_labels;
    fa.Hinv0 = mat3(p->Hinv0); fa.Hinv1 = mat3(p->Hinv1); fa.q0 = pt2(p->q0); fa.q1 = pt2(p->q1);
    fa.B_fwd = p->B_fwd; fa.b0 = bx[0]; fa.b1 = bx[1];
    fa.par0 = par0; fa.par1 = par1; fa.S0 = S0; fa.S1 = S1; fa.cand = cand; fa.count = count; fa.st = dev_state;
    fa.partial = partial; fa.sel = sel;
    fa.gbuf = (unsigned long long*)((char*)sel + front_sel_bytes());
    void* args[] = {&fa};
    e = cudaLaunchCooperativeKernel((const void*)k_front_fused, dim3((unsigned)blocks), dim3(256), args, 0, S(stream));
    if (e != cudaSuccess) return fail((int)e, "front: cooperative launch: %s", cudaGetErrorString(e));
    return check_launch("front/fused");
  }
  // separate launches (the default; what row bands run too, with collectives in between)
  // same grid and summation tree as mrf_epipole_dist_sum, so mu has the same bits on both routes
  k_front_parallax<<<kRedBlocks, kRedThreads, 0, S(stream)>>>(u0, v0, u1, v1, p->rows, p->w, 0, mat3(p->Hinv0),
                                                              mat3(p->Hinv1), pt2(p->q0), pt2(p->q1), par0, par1, partial);
  int rc = check_launch("front/parallax");
  if (rc) return rc;
  // mu, pv and the candidate ratios of :209-219 with B_fwd fixed (build_cost_volume.py:116-131), first select pass
  k_front_candidates13<<<front_cand_blocks(n), 256, 0, S(stream)>>>(par0, par1, rigid_mask, p->rows, p->w, 0, pt2(p->q0),
                                                                    pt2(p->q1), p->B_fwd, cand, count, dev_state, partial,
                                                                    (double)n, sel, PeerArg{});
  rc = check_launch("front/candidates");
  if (rc) return rc;
  for (int pass = 1; pass < kS13Passes && !rc; ++pass) rc = mrf_select13_hist_dev(cand, count, (size_t)n, pass, sel, nullptr, stream);
  if (!rc) rc = mrf_select13_successor_dev(cand, count, (size_t)n, sel, nullptr, stream);
  if (rc) return rc;
  k_front_structures<<<(unsigned)fb, 256, 0, S(stream)>>>(par0, par1, p->rows, p->w, 0, pt2(p->q0), pt2(p->q1), p->B_fwd,
                                                          zero_mask, bx[0], bx[1], S0, S1, dev_state, partial, sel, nullptr, 0, PeerArg{});
  rc = check_launch("front/structures");
  if (rc) return rc;
  k_front_labels<<<1, 256, 0, S(stream)>>>(partial, (int)fb, dev_state, p->n_labels);
  return check_launch("front/labels");
}


// ---- the same pre-pass in pieces for row bands: between the pieces the five 8192-bin
// histograms of the exact median, its three successor words and the six range words are exchanged -- by NCCL
// collectives the caller enqueues on the stream, or, with a peer context, by the kernels themselves over NVLink
// peer memory (peersync.cuh).  mu needs no exchange (mrf_front_mu_whole).  No host round trip either way.
int mrf_front_mu_whole(const mrf_front_params* p, double* dev_state, void* scratch, mrf_stream_t stream) {
  MRF_CHECK_ARG(p && dev_state && scratch && p->frame_h > 0 && p->w > 0);
  double* partial = (double*)scratch;
  for (int f = 0; f < 2; ++f) {
    k_dist_sum_partial<<<kRedBlocks, kRedThreads, 0, S(stream)>>>(p->frame_h, p->w, 0, pt2(f ? p->q1 : p->q0),
                                                                  partial + f * kRedBlocks);
    const int rc = check_launch("front_mu_whole/partial");
    if (rc) return rc;
  }
  k_front_mu<<<1, kRedBlocks, 0, S(stream)>>>(partial, kRedBlocks, (double)p->frame_h * (double)p->w, dev_state);
  return check_launch("front_mu_whole/mu");
}

int mrf_front_parallax_band(const mrf_front_params* p, const float* u0, const float* v0, const float* u1,
                            const float* v1, double* par0, double* par1, void* scratch, mrf_stream_t stream) {
  MRF_CHECK_ARG(p && u0 && v0 && u1 && v1 && par0 && par1 && scratch && p->rows > 0 && p->w > 0);
  double* partial = (double*)scratch;        // the band's own distance sums: not used (mu comes from mrf_front_mu_whole)
  k_front_parallax<<<kRedBlocks, kRedThreads, 0, S(stream)>>>(u0, v0, u1, v1, p->rows, p->w, p->y0, mat3(p->Hinv0),
                                                              mat3(p->Hinv1), pt2(p->q0), pt2(p->q1), par0, par1, partial);
  return check_launch("front_band/parallax");
}

int mrf_front_candidates_band(const mrf_front_params* p, const double* par0, const double* par1,
                              const uint8_t* rigid_mask, double* dev_state, double* cand, int* count,
                              void* select_scratch, const mrf_peer_sync* peer, mrf_stream_t stream) {
  MRF_CHECK_ARG(p && par0 && par1 && rigid_mask && dev_state && cand && count && select_scratch);
  PeerArg pa;
  if (peer_arg(peer, &pa)) return MRF_E_BADARG;
  cudaError_t e = cudaMemsetAsync(select_scratch, 0, kS13Bytes, S(stream));
  if (e == cudaSuccess) e = cudaMemsetAsync(count, 0, sizeof(int), S(stream));
  if (e != cudaSuccess) return fail((int)e, "front_band: %s", cudaGetErrorString(e));
  const long long n = (long long)p->rows * p->w;
  k_front_candidates13<<<front_cand_blocks(n), 256, 0, S(stream)>>>(par0, par1, rigid_mask, p->rows, p->w, p->y0,
                                                                    pt2(p->q0), pt2(p->q1), p->B_fwd, cand, count,
                                                                    dev_state, nullptr, 0.0, select_scratch, pa);
  return check_launch("front_band/candidates");
}

int mrf_front_structures_band(const mrf_front_params* p, const double* par0, const double* par1,
                              const uint8_t* zero_mask, double* S0, double* S1, double* dev_state,
                              const void* select_scratch, const unsigned long long* gathered_successor, int world,
                              void* scratch, const mrf_peer_sync* peer, mrf_stream_t stream) {
  MRF_CHECK_ARG(p && par0 && par1 && S0 && S1 && dev_state && scratch && select_scratch);
  MRF_CHECK_ARG(!gathered_successor || world >= 1);
  PeerArg pa;
  if (peer_arg(peer, &pa)) return MRF_E_BADARG;
  const long long n = (long long)p->rows * p->w;
  long long fb = (n + 255) / 256;
  if (fb > kFrontBlocks) fb = kFrontBlocks;
  Box4 bx[2];
  for (int f = 0; f < 2; ++f) {
    bx[f].x0 = p->box[f][0]; bx[f].x1 = p->box[f][1]; bx[f].y0 = p->box[f][2]; bx[f].y1 = p->box[f][3];
    bx[f].on = p->have_box[f];
  }
  double* partial = (double*)scratch;
  k_front_structures<<<(unsigned)fb, 256, 0, S(stream)>>>(par0, par1, p->rows, p->w, p->y0, pt2(p->q0), pt2(p->q1),
                                                          p->B_fwd, zero_mask, bx[0], bx[1], S0, S1, dev_state, partial,
                                                          select_scratch, gathered_successor, world, pa);
  int rc = check_launch("front_band/structures");
  if (rc) return rc;
  k_front_minmax<<<1, 256, 0, S(stream)>>>(partial, (int)fb, dev_state, pa);
  return check_launch("front_band/minmax");
}

}  // extern "C"
