// The per-label data cost of the plane+parallax model -- the job of the extension that is missing
// from the reference (pipeline/build_cost_volume.py:209-219), re-specified as
// one evaluation of the residual of computeDataTerm
// (structure_refinement/optimize_structure_single_scale.py:210-286) per label:
//
//   geometry :210-235   vx = qx-x, vy = qy-y, nrm = max(0.5, |v|), n = |v|/mu,
//                       r = (A*B*n)/(B*A/mu - 1), (xn, yn) = (x, y) + r*(v/nrm)          [fp64]
//   warp     :238       interpolate_through_H.py:30-42: (X, Y) = H(xn, yn) in fp64, inside-frame mask
//                       on the fp64 coordinates, cv2.remap(fp32 image, fp32 coords, CUBIC, REPLICATE)
//   residual :269-286   Iz = mean_c((J_c - I1_c) * cw_c); 0 where outside / occluded / non-rigid
//   cost     :302       rho(Iz^2), utils/robust_functions.py:26-70; stored as fp32
//
// One thread owns one reference pixel and walks the labels, so the label-major [L][rows][w] volume
// is written with fully coalesced 128-byte warp stores and the per-pixel min/argmin is a register
// reduction.  A block owns a 32x8 pixel tile.  In the staged variant the block first bounds the
// footprint of its tile over the whole label range (the sample points of one pixel lie on the
// image of its epipolar segment under H), copies that window of the neighbour frame's packed
// texels (gray, d/dy, d/dx, 0) into shared memory with one TMA bulk copy per window row
// (cp.async.bulk global->shared completing on an mbarrier), and gathers the 16 bicubic taps with
// LDS.128.  A sample whose footprint touches the image border (cv2.remap's clamped, running-sum
// order) takes its taps from the window too (sample_window_border); one whose footprint leaves the
// staged window (labels near the pole of the parallax formula, very large parallax) takes the
// global-memory gather path, which is also the whole kernel in the unstaged variant -- results are
// bit-identical either way.
//
// Arithmetic is written operation by operation in the reference's order and compiled with
// -fmad=false, so the fp64 coordinates, the 1/32-px bins, the fp32 tap sums and the fp64 residual
// are the oracle's bit for bit; only log() (relative error < 2^-43 before the fp32 rounding; one
// implementation per sampler, log_ge1) can differ.  Divisions by per-label / per-launch constants and the two projective divisions that share
// a divisor are evaluated with a reciprocal and exact FMA remainder corrections (div_rc), which
// returns the correctly rounded quotient -- the same bits as a division -- at a fifth of the cost.
#include "sampling.cuh"

namespace mrf {

#ifndef MRF_TY
#define MRF_TY 8          // tile rows (warps per block); experiments: -DMRF_TY=7 with MRF_VARIANT_OCC4
#endif
#ifndef MRF_GROUP_LINEAR
#define MRF_GROUP_LINEAR 2   // labels per iteration of the grouped loop, bilinear sampling (experiments: 3, 4)
#endif
constexpr int kTX = 32, kTY = MRF_TY, kThreads = kTX * kTY;
constexpr int kMaxLabels = 256;
constexpr int kWinCap = 3072;      // staged window capacity in texels (48 KB)
constexpr int kWinCap4 = 2560;     // ... with four blocks per SM (40 KB)
constexpr int kWinMaxW = 384;      // widest staged row

struct CVArgs {
  int rows, w, y0, frame_h;
  int nb_rows, nb_y0;
  Mat3 H;
  Pt2 q;
  double mu, B, label_scale;
  int L, rho;
  double rho_param;
  double cw0, cw1, cw2;
  int layout;
  float i32_scale;
  const double* ref_ch;
  const float4* nb;
  const double* labels;
  const uint8_t* nonrigid;
  const uint8_t* occluded;
  void* cost;
  float* min_cost;
  int32_t* argmin;
  int* halo_miss;
  long long plane_stride;   // elements between label planes of the LHW volume
  const double* state;      // device-resident scalars (mu, B, labels) or nullptr
  int frame;
  int win_cap;              // staged window capacity in texels (<= kWinCap; smaller when 4 blocks share an SM)
  double wm1, hm1;          // (double)(w - 1), (double)(frame_h - 1)
  double s2, rs2;           // rho_param^2 and RN(1 / s2)
};

// utils/robust_functions.py:26-70, argument already squared
__device__ __forceinline__ double rho_eval(int kind, double x, double p) {
  switch (kind) {
    case MRF_RHO_LORENTZIAN:      return log(1.0 + x / (p * p));                 // :52
    case MRF_RHO_CHARBONNIER:     return sqrt(x + p * p);                        // :33 (a = 0.5)
    case MRF_RHO_GEMAN_MCCLURE:   return (p * x) / (x + p * p);                  // :64
    case MRF_RHO_QUADRATIC:       return x;                                      // :46
    case MRF_RHO_LORENTZIAN_AUTO: return p * log(1.0 + (0.5 * x) / (p * p));     // :58
    default:                      return x;
  }
}

struct FetchBand {   // neighbour texels in global memory, band rows nb_y0 .. nb_y0+nb_rows-1
  const float4* nb; int w; int nb_y0;
  __device__ __forceinline__ F3 operator()(int yy, int xx) const {
    const float4 t = __ldg(nb + ((long long)(yy - nb_y0) * w + xx));
    F3 r; r.x = t.x; r.y = t.y; r.z = t.z; return r;
  }
};
template <int INTERP, class Fetch>
__device__ __forceinline__ F3 sample(const Fetch& f, int h, int w, float xf, float yf) {
  if (INTERP == MRF_INTERP_CUBIC) return sample_cubic(f, h, w, xf, yf);
  return sample_linear(f, h, w, xf, yf);
}

template <int INTERP>
__device__ __noinline__ F3 sample_band(const float4* nb, int w, int nb_y0, int frame_h, float xf, float yf) {
  FetchBand f{nb, w, nb_y0};
  return sample<INTERP>(f, frame_h, w, xf, yf);
}

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void lds128(uint32_t addr, float& a, float& b, float& c, float& d) {
  asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(a), "=f"(b), "=f"(c), "=f"(d) : "r"(addr));
}

__device__ __forceinline__ void mbar_init(uint64_t* bar, int count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "WAIT_%=:\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
      "@p bra DONE_%=;\n\t"
      "bra WAIT_%=;\n\t"
      "DONE_%=:\n\t}" ::"r"(smem_u32(bar)), "r"(parity) : "memory");
}
// TMA bulk copy global -> shared, completion counted in bytes on the mbarrier (SASS: UBLKCP)
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
               ::"r"(smem_u32(dst)), "l"(src), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}

// ---------------------------------------------------------------------------------------------
// Exact fp64 division with a known reciprocal.  For rd = RN(1/d) and no over/underflow,
// q0 = a*rd is within 2 ulp of a/d, one remainder correction makes it faithful, and by Markstein's
// theorem the second correction (exact remainder of a faithful quotient, fma with RN(1/d)) returns
// RN(a/d) -- the same bits as the reference's `/`, in 5 fp64 instructions instead of a division
// subroutine.  tests/test_exact_division.py replays the sequence in rational arithmetic.
__device__ __forceinline__ double div_rc(double a, double d, double rd) {
  double q = a * rd;
  double r = fma(-d, q, a);
  q = fma(r, rd, q);
  r = fma(-d, q, a);
  return fma(r, rd, q);
}
// a / 3 (the channel mean, `.mean(axis=2)`), correctly rounded with ONE correction.  With c = RN(1/3) = (1/3)(1 - 2^-54):
// q0 = RN(a c) is within 1.5 ulp of q* = a/3, the remainder r = a - 3 q0 is exact (a small multiple of ulp(q0)), and
// q0 + r c = q* - (q* - q0) 2^-54, i.e. q* displaced by less than 2^-54 ulp.  a/3 is never closer than ulp/6 to a rounding
// boundary (3 x a 54-bit midpoint is odd in a bit position below ulp(a)), so the final rounding returns RN(a/3).
// tests/test_exact_division.py replays the sequence in rational arithmetic.
__device__ __forceinline__ double div3(double a, double third) {
  const double q = a * third;
  const double r = fma(-3.0, q, a);
  return fma(r, third, q);
}
// exponent of a double within [2^-100, 2^100] (and finite, non-zero): the range in which the
// sequences above cannot over/underflow for this kernel's operands
__device__ __forceinline__ bool tame(double v) {
  const unsigned ex = ((unsigned)__double2hiint(v) >> 20) & 0x7ffu;
  return (ex - 923u) <= 200u;
}

// The bicubic kernel's log(u), u >= 1 (the kernel is bound by shared-memory wavefronts there, and the table form's
// scattered load costs more than its fp64 instructions save: 0.310 against 0.301 ms per frame).  Argument
// reduction u = m*2^e with m in [sqrt(1/2), sqrt(2)), s = (m-1)/(m+1), log m = 2 atanh(s) as an
// odd polynomial through s^15; relative error < 2^-43, i.e. the fp32 cost the kernel stores is the
// correctly rounded one except for about one sample in 2^19 (then off by one fp32 ulp, 6e-8
// relative, against the 1e-5 bar).  numpy's own log is only specified to < 1 ulp of fp64 either.
__constant__ double kLn2 = 0.6931471805599453094;
__constant__ double kThird = 1.0 / 3.0;
__constant__ double kAtanhC[7] = {1.0 / 3.0, 1.0 / 5.0, 1.0 / 7.0, 1.0 / 9.0, 1.0 / 11.0, 1.0 / 13.0, 1.0 / 15.0};

__device__ __forceinline__ double log_ge1_series(double u) {
  int hi = __double2hiint(u);
  const int lo = __double2loint(u);
  const bool special = hi >= 0x7ff00000;                // inf / NaN: returned as they are
  int e = (hi >> 20) - 1023;
  hi = (hi & 0x000fffff) | 0x3ff00000;
  if (hi >= 0x3ff6a09e) { hi -= 0x00100000; e += 1; }   // m >= sqrt(2): halve
  const double m = __hiloint2double(hi, lo);
  const double f = m - 1.0;
  const double d = 2.0 + f;
  double y;
  asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(d));
  double t = fma(-d, y, 1.0);
  y = fma(y, t, y);
  t = fma(-d, y, 1.0);
  y = fma(y, t, y);
  const double sv = f * y;
  const double s2 = sv * sv;
  double pl = fma(s2, kAtanhC[6], kAtanhC[5]);
  pl = fma(s2, pl, kAtanhC[4]);
  pl = fma(s2, pl, kAtanhC[3]);
  pl = fma(s2, pl, kAtanhC[2]);
  pl = fma(s2, pl, kAtanhC[1]);
  pl = fma(s2, pl, kAtanhC[0]);
  pl = pl * s2;
  const double two_s = sv + sv;
  const double res = fma((double)e, kLn2, fma(two_s, pl, two_s));
  return special ? u : res;
}


// The bilinear kernel's log(u), u >= 1 (that kernel is bound by the fp64 pipe: -5 % with this form;
// utils/robust_functions.py:52,58).  u = m * 2^e with m in [1, 2);
// the top seven mantissa bits pick (RN(1/c), RN(log c)) for the midpoint c of m's 1/128-wide interval from a 2 KB table
// in shared memory (log_table.inc; the first interval uses c = 1, so that r = m - 1 is exact and the relative
// accuracy holds as u -> 1); r = m/c - 1 as one fma, |r| < 2^-7; log m = log c + log1p(r), log1p as its Taylor
// polynomial through r^6 (next term < 2^-44.8 of the result).  Relative error < 2^-44, i.e. the fp32 cost the kernel
// stores is the correctly rounded one except for about one sample in 2^20 (then off by one fp32 ulp, 6e-8 relative,
// against the 1e-5 bar).  numpy's own log is only specified to < 1 ulp of fp64 either.  Nine fp64 instructions and
// one shared-memory load; the argument-reduced atanh series it replaces took 21 and a reciprocal.
__device__ const double kLogTab[128][2] = {
#include "log_table.inc"
};
__constant__ double kLog1pC[4] = {-1.0 / 6.0, 1.0 / 5.0, -1.0 / 4.0, 1.0 / 3.0};

__device__ __forceinline__ double log_ge1_table(double u, uint32_t tab_s) {
  const int hi = __double2hiint(u);
  const int lo = __double2loint(u);
  const bool special = hi >= 0x7ff00000;                // inf / NaN: returned as they are
  const int e = (hi >> 20) - 1023;
  const double m = __hiloint2double((hi & 0x000fffff) | 0x3ff00000, lo);
  double rc, lc;
  asm volatile("ld.shared.v2.f64 {%0, %1}, [%2];" : "=d"(rc), "=d"(lc) : "r"(tab_s + (((uint32_t)hi >> 13) & 127u) * 16u));
  const double r = fma(m, rc, -1.0);
  double q = fma(r, kLog1pC[0], kLog1pC[1]);
  q = fma(r, q, kLog1pC[2]);
  q = fma(r, q, kLog1pC[3]);
  q = fma(r, q, -0.5);
  const double t = fma(r * r, q, r);                    // log1p(r)
  const double res = fma((double)e, kLn2, lc + t);
  return special ? u : res;
}

// one log per sampler: every variant of a sampler (staged, gather, the cold paths) uses the same one
template <int INTERP>
__device__ __forceinline__ double log_ge1(double u, uint32_t tab_s) {
  if constexpr (INTERP == MRF_INTERP_LINEAR) return log_ge1_table(u, tab_s);
  else return log_ge1_series(u);
}

// Reciprocal of a tame double, correctly rounded: hardware seed (MUFU.RCP64H, ~2^-23) and three
// Newton steps.  The third step applied to an estimate already within an ulp returns RN(1/d) unless
// d's significand is all ones (Markstein; such a divisor is sent down the plain-division path).
__device__ __forceinline__ double rcp_rn_tame(double d) {
  double y;
  asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(d));
  double e = fma(-d, y, 1.0);
  y = fma(y, e, y);
  e = fma(-d, y, 1.0);
  y = fma(y, e, y);
  e = fma(-d, y, 1.0);
  return fma(y, e, y);
}
// tame exponent and significand not all ones
__device__ __forceinline__ bool tame_divisor(double v) {
  const unsigned hi = (unsigned)__double2hiint(v);
  const bool ones = ((hi & 0xfffffu) == 0xfffffu) && ((unsigned)__double2loint(v) == 0xffffffffu);
  // |v| in [2^-100, 2^101): biased exponent 923 .. 1123, tested on the high word with the sign cleared
  return (((hi & 0x7fffffffu) - 0x39b00000u) < 0x0c900000u) && !ones;
}

// One sample evaluated with plain IEEE divisions and the library log -- the arithmetic of the
// reference written naively.  Cold path: taken only for samples whose operands leave the range in
// which the reciprocal sequences are proven exact (labels at the pole of the parallax formula,
// degenerate homographies), a handful per frame at most.
template <int INTERP>
__device__ __noinline__ float slow_sample_cost(const CVArgs& a, double ab, double dl, double n, double xd, double yd,
                                               double t1, double t2, double c0, double c1, double c2) {
  const double r = (ab * n) / dl;
  const double xn = xd + r * t1, yn = yd + r * t2;
  double X, Y; bool inside;
  through_H(a.H, xn, yn, a.frame_h, a.w, X, Y, inside);
  double cst = rho_eval(a.rho, 0.0, a.rho_param);
  if (inside) {
    const float xf = (float)X, yf = (float)Y;
    int iy, fy; cv_quantise(yf, iy, fy);
    const int lo = (INTERP == MRF_INTERP_CUBIC) ? -1 : 0, hi = (INTERP == MRF_INTERP_CUBIC) ? 2 : 1;
    const int ya = max(0, min(a.frame_h - 1, iy + lo)), yb = max(0, min(a.frame_h - 1, iy + hi));
    F3 J;
    if (ya < a.nb_y0 || yb > a.nb_y0 + a.nb_rows - 1) {
      if (a.halo_miss) atomicAdd(a.halo_miss, 1);
      J.x = J.y = J.z = 0.f;
    } else {
      FetchBand f{a.nb, a.w, a.nb_y0};
      J = sample<INTERP>(f, a.frame_h, a.w, xf, yf);
    }
    const double e0 = ((double)J.x - c0) * a.cw0;
    const double e1 = ((double)J.y - c1) * a.cw1;
    const double e2 = ((double)J.z - c2) * a.cw2;
    const double Iz = ((e0 + e1) + e2) / 3.0;
    cst = (a.rho == MRF_RHO_NONE) ? Iz : rho_eval(a.rho, Iz * Iz, a.rho_param);
  }
  return (float)cst;
}

// A sample with exact coordinates whose footprint touches the image border or leaves the staged
// window: global-memory gather (for row-band inputs with the halo check), then the same residual / rho.
struct WinRef { const int* s_win; uint32_t win_s, pitch, wt_s, log_s; };   // the staged window (s_win = nullptr: none), the log table
template <int INTERP> __device__ __noinline__ F3 sample_window_border(uint32_t win_s, uint32_t pitch, uint32_t wt_s, int w, int frame_h, int sx, int sy);
template <int INTERP> __device__ __forceinline__ bool border_in_window(const int* s_win, int w, int frame_h, int sx, int sy);

template <int INTERP>
__device__ __noinline__ float band_sample_cost(const CVArgs& a, float xf, float yf, double c0, double c1, double c2,
                                               const WinRef wr) {
  int iy, fy; cv_quantise(yf, iy, fy);
  const int lo = (INTERP == MRF_INTERP_CUBIC) ? -1 : 0, hi = (INTERP == MRF_INTERP_CUBIC) ? 2 : 1;
  const int ya = max(0, min(a.frame_h - 1, iy + lo)), yb = max(0, min(a.frame_h - 1, iy + hi));
  F3 J;
  // an in-frame sample (the only kind that comes here): |coord * 32| < 2^31, the quantisation cannot saturate
  const int sx = __float2int_rn(xf * 32.0f), sy = __float2int_rn(yf * 32.0f);
  if (wr.s_win && border_in_window<INTERP>(wr.s_win, a.w, a.frame_h, sx, sy)) {
    J = sample_window_border<INTERP>(wr.win_s, wr.pitch, wr.wt_s, a.w, a.frame_h, sx, sy);   // image border, staged taps
  } else if (ya < a.nb_y0 || yb > a.nb_y0 + a.nb_rows - 1) {
    if (a.halo_miss) atomicAdd(a.halo_miss, 1);
    J.x = J.y = J.z = 0.f;
  } else {
    FetchBand f{a.nb, a.w, a.nb_y0};
    J = sample<INTERP>(f, a.frame_h, a.w, xf, yf);
  }
  const double e0 = ((double)J.x - c0) * a.cw0;
  const double e1 = ((double)J.y - c1) * a.cw1;
  const double e2 = ((double)J.z - c2) * a.cw2;
  const double Iz = div3((e0 + e1) + e2, kThird);
  double cst;
  if (a.rho == MRF_RHO_LORENTZIAN && tame(a.s2)) cst = log_ge1<INTERP>(1.0 + div_rc(Iz * Iz, a.s2, a.rs2), wr.log_s);
  else cst = (a.rho == MRF_RHO_NONE) ? Iz : rho_eval(a.rho, Iz * Iz, a.rho_param);
  return (float)cst;
}


// ---------------------------------------------------------------------------------------------
// Pieces of the grouped label loop (FORM 2): the loop handles the labels in pairs, and each phase of
// the pair -- fp64 geometry, the tap gathers, the fp64 residual / rho tail -- is one branch-free
// straight line over both labels, so that the two dependent fp64 chains of a phase interleave
// (the single-label forms leave the fp64 pipe's latency exposed: `wait` was the top stall reason).
// Every operation and its order are those of the other forms; results are bit-identical.
struct SampleCode {
  uint32_t p0;      // shared-memory address of the footprint's first texel (a harmless address when !fast)
  uint32_t wx, wy;  // addresses of the two weight vectors in the 1/32-px table
  bool fast;        // exact sequences proven, inside the frame, footprint inside the staged window, pixel alive
  bool cold;        // alive, not fast, and the sample counts (inexact operands, or inside the frame)
};
struct WinGeom { int fx_lo, fy_lo, fx_span, fy_span; uint32_t win_s, win_base, pitch, wt_s; };

template <int INTERP>
__device__ __forceinline__ void sample_code(const CVArgs& a, uint32_t lab_addr, double n_fast, double xk, double yk,
                                            double t1, double t2, const WinGeom& g, bool dead, SampleCode& s) {
  constexpr int kOff = (INTERP == MRF_INTERP_CUBIC) ? 1 : 0;
  double ab, dl, rdl;
  asm volatile("ld.shared.v2.f64 {%0, %1}, [%2];" : "=d"(ab), "=d"(dl) : "r"(lab_addr));
  asm volatile("ld.shared.f64 %0, [%1];" : "=d"(rdl) : "r"(lab_addr + 16u));
  const double r = div_rc(ab * n_fast, dl, rdl);             // r = A*B*n / (B*A/mu - 1)  :221
  const double xn = xk + r * t1, yn = yk + r * t2;           // :224-235
  const double den = (xn * a.H.m[6] + yn * a.H.m[7]) + a.H.m[8];   // interpolate_through_H.py:31-35
  const double Xn = (xn * a.H.m[0] + yn * a.H.m[1]) + a.H.m[2];
  const double Yn = (xn * a.H.m[3] + yn * a.H.m[4]) + a.H.m[5];
  const double rd = rcp_rn_tame(den);
  const double X = div_rc(Xn, den, rd);
  const double Y = div_rc(Yn, den, rd);
  const bool exact = tame_divisor(den);                      // else: NaN / wild operands somewhere above
  const bool inside = (X >= 0.0) && (X <= a.wm1) && (Y >= 0.0) && (Y <= a.hm1);   // :39
  const float xf = (float)X, yf = (float)Y;                  // interpolation.py:77-78
  const int sx = __float2int_rn(xf * 32.0f), sy = __float2int_rn(yf * 32.0f);      // cv2.remap's 1/32-px bins
  const int ix0 = (sx >> 5) - kOff, iy0 = (sy >> 5) - kOff;
  const bool inwin = (unsigned)(ix0 - g.fx_lo) <= (unsigned)g.fx_span && (unsigned)(iy0 - g.fy_lo) <= (unsigned)g.fy_span;
  s.fast = exact && inside && inwin && !dead;
  s.cold = !dead && !s.fast && (!exact || inside);
  s.p0 = s.fast ? (g.win_s + (uint32_t)iy0 * g.pitch + (uint32_t)ix0 * 16u) : g.win_base;
  s.wx = g.wt_s + (uint32_t)(sx & 31) * 16u;
  s.wy = g.wt_s + (uint32_t)(sy & 31) * 16u;
}

// the taps of one sample out of the staged window: OpenCV's order (bicubic: row sums left to right, then the
// rows; bilinear: one running sum), every product and sum rounded separately
template <int INTERP>
__device__ __forceinline__ F3 window_taps(const SampleCode& s, uint32_t pitch) {
  constexpr int kTaps = (INTERP == MRF_INTERP_CUBIC) ? 4 : 2;
  float wxa[4], wya[4];
  lds128(s.wx, wxa[0], wxa[1], wxa[2], wxa[3]);
  lds128(s.wy, wya[0], wya[1], wya[2], wya[3]);
  F3 J;
#pragma unroll
  for (int j = 0; j < kTaps; ++j) {
    F3 row;
#pragma unroll
    for (int i = 0; i < kTaps; ++i) {
      const float wgt = wya[j] * wxa[i];
      float tx, ty, tz, tw;
      lds128(s.p0 + (uint32_t)j * pitch + (uint32_t)i * 16u, tx, ty, tz, tw);
      const float px = tx * wgt, py = ty * wgt, pz = tz * wgt;
      if (INTERP == MRF_INTERP_CUBIC) {
        if (i == 0) { row.x = px; row.y = py; row.z = pz; }
        else        { row.x = row.x + px; row.y = row.y + py; row.z = row.z + pz; }
      } else {
        if (i == 0 && j == 0) { J.x = px; J.y = py; J.z = pz; }
        else { J.x = J.x + px; J.y = J.y + py; J.z = J.z + pz; }
      }
    }
    if (INTERP == MRF_INTERP_CUBIC) {
      if (j == 0) J = row;
      else { J.x = J.x + row.x; J.y = J.y + row.y; J.z = J.z + row.z; }
    }
  }
  return J;
}

// A sample the grouped loop could not take from the window: geometry again with the same sequences (same
// bits), then the plain-division path or the global-memory gather.  Cold.
template <int INTERP>
__device__ __noinline__ float cold_cost(const CVArgs& a, uint32_t lab_addr, double n, double xk, double yk, double t1,
                                        double t2, uint32_t ref_s, float cost_dead, const WinRef wr) {
  double ab, dl, rdl, r0, r1, r2;
  asm volatile("ld.shared.v2.f64 {%0, %1}, [%2];" : "=d"(ab), "=d"(dl) : "r"(lab_addr));
  asm volatile("ld.shared.f64 %0, [%1];" : "=d"(rdl) : "r"(lab_addr + 16u));
  asm volatile("ld.shared.f64 %0, [%1];" : "=d"(r0) : "r"(ref_s));
  asm volatile("ld.shared.f64 %0, [%1];" : "=d"(r1) : "r"(ref_s + (uint32_t)(kThreads * 8)));
  asm volatile("ld.shared.f64 %0, [%1];" : "=d"(r2) : "r"(ref_s + (uint32_t)(kThreads * 16)));
  const double n_fast = ((n == 0.0) || tame(n)) ? n : __longlong_as_double(0x7FF8000000000000ll);
  const double r = div_rc(ab * n_fast, dl, rdl);
  const double xn = xk + r * t1, yn = yk + r * t2;
  const double den = (xn * a.H.m[6] + yn * a.H.m[7]) + a.H.m[8];
  if (!tame_divisor(den)) return slow_sample_cost<INTERP>(a, ab, dl, n, xk, yk, t1, t2, r0, r1, r2);
  const double Xn = (xn * a.H.m[0] + yn * a.H.m[1]) + a.H.m[2];
  const double Yn = (xn * a.H.m[3] + yn * a.H.m[4]) + a.H.m[5];
  const double rd = rcp_rn_tame(den);
  const double X = div_rc(Xn, den, rd);
  const double Y = div_rc(Yn, den, rd);
  const bool inside = (X >= 0.0) && (X <= a.wm1) && (Y >= 0.0) && (Y <= a.hm1);
  if (!inside) return cost_dead;
  return band_sample_cost<INTERP>(a, (float)X, (float)Y, r0, r1, r2, wr);
}

template <int N> struct IntC { static constexpr int value = N; };

// A sample whose footprint touches the image border (cv2.remap's BORDER_REPLICATE path: taps clamped to the frame, ONE
// running sum over the taps in row-major order -- not the row sums of an interior footprint) taken from the staged
// window: the clamped taps of such a sample lie inside the window whenever the sample itself does.  The pixels of
// the frame's outermost columns and rows produce these for most labels (0.4 % of the samples, but one lane is enough
// to send its warp here: 4.5 % of the warp-samples at Sintel size), and the global-memory gather they used to take
// cost 12 % of the kernel (profiles/r4n_ablation_study.jsonl).  win_s: shared address of texel (0, 0).
template <int INTERP>
__device__ __noinline__ F3 sample_window_border(uint32_t win_s, uint32_t pitch, uint32_t wt_s, int w, int frame_h, int sx, int sy) {
  constexpr int kTaps = (INTERP == MRF_INTERP_CUBIC) ? 4 : 2;
  constexpr int kOff = (INTERP == MRF_INTERP_CUBIC) ? 1 : 0;
  const int ix0 = (sx >> 5) - kOff, iy0 = (sy >> 5) - kOff;
  float wxa[4], wya[4];
  lds128(wt_s + (uint32_t)(sx & 31) * 16u, wxa[0], wxa[1], wxa[2], wxa[3]);
  lds128(wt_s + (uint32_t)(sy & 31) * 16u, wya[0], wya[1], wya[2], wya[3]);
  F3 acc;
#pragma unroll
  for (int j = 0; j < kTaps; ++j) {
    const int yy = max(0, min(frame_h - 1, iy0 + j));
    const uint32_t row = win_s + (uint32_t)yy * pitch;
#pragma unroll
    for (int i = 0; i < kTaps; ++i) {
      const int xx = max(0, min(w - 1, ix0 + i));
      const float wgt = wya[j] * wxa[i];
      float tx, ty, tz, tw;
      lds128(row + (uint32_t)xx * 16u, tx, ty, tz, tw);
      const float px = tx * wgt, py = ty * wgt, pz = tz * wgt;
      if (i == 0 && j == 0) { acc.x = px; acc.y = py; acc.z = pz; }
      else { acc.x = acc.x + px; acc.y = acc.y + py; acc.z = acc.z + pz; }
    }
  }
  return acc;
}
// the clamped footprint of (sx, sy) lies inside the staged window s_win = (wx0, wy0, ww, wh)
template <int INTERP>
__device__ __forceinline__ bool border_in_window(const int* s_win, int w, int frame_h, int sx, int sy) {
  constexpr int kTaps = (INTERP == MRF_INTERP_CUBIC) ? 4 : 2;
  constexpr int kOff = (INTERP == MRF_INTERP_CUBIC) ? 1 : 0;
  const int ix0 = (sx >> 5) - kOff, iy0 = (sy >> 5) - kOff;
  const int cx0 = max(0, min(w - 1, ix0)), cx1 = max(0, min(w - 1, ix0 + kTaps - 1));
  const int cy0 = max(0, min(frame_h - 1, iy0)), cy1 = max(0, min(frame_h - 1, iy0 + kTaps - 1));
  const int wx0 = s_win[0], wy0 = s_win[1], ww = s_win[2], wh = s_win[3];
  return cx0 >= wx0 && cx1 < wx0 + ww && cy0 >= wy0 && cy1 < wy0 + wh;
}

struct __align__(16) LabelRec { double ab, den, rden, safe; };   // safe: 1.0 / 0.0

template <int INTERP, bool STAGED, int LAYOUT, int FORM>
__device__ __forceinline__ void cost_volume_block(const CVArgs& a, unsigned char* dyn_smem) {
  __shared__ LabelRec s_lab[kMaxLabels];    // A*B, B*A/mu - 1, RN(1/that), exact-sequence-safe flag
  __shared__ __align__(16) float4 s_wt[32]; // OpenCV's interpolation table: weights at frac/32
  __shared__ __align__(16) double s_log[INTERP == MRF_INTERP_LINEAR ? 128 : 1][2];   // log_ge1_table's table (bilinear only)
  __shared__ double s_c[2];                 // min / max over labels of A*B / (B*A/mu - 1)
  __shared__ double s_ref[3][kThreads];     // this thread's reference-frame channels (register relief)
  __shared__ float s_red[kTY][4];
  __shared__ int s_win[4];                  // wx0, wy0, ww, wh
  __shared__ __align__(8) uint64_t s_bar;

  const int lane = threadIdx.x & 31, wrp = threadIdx.x >> 5;
  const int x = blockIdx.x * kTX + lane;
  const int yl = blockIdx.y * kTY + wrp;
  const bool valid = (x < a.w) && (yl < a.rows);
  const int y = yl + a.y0;
  const int L = a.L;

  // whole-frame scalars: from the host, or from the device state the pre-pass left in HBM
  double g_mu = a.mu, g_B = a.B, g_scale = a.label_scale;
  const double* g_labels = a.labels;
  if (a.state) {
    g_mu = a.state[MRF_STATE_MU + a.frame];
    g_B = a.state[MRF_STATE_B + a.frame];
    g_scale = (a.frame == 0) ? a.state[MRF_STATE_MU] / a.state[MRF_STATE_MU + 1] : 1.0;   // :435-436
    g_labels = a.state + MRF_STATE_LABELS;
  }
  // per-label scalars: a = label*scale (:435-436), A*B and B*A/mu - 1 (:221)
  for (int l = threadIdx.x; l < L; l += kThreads) {
    const double al = g_labels[l] * g_scale;
    const double ab = al * g_B;
    const double den = ab / g_mu - 1.0;
    // a label for which the reciprocal sequence is not proven exact (pole of the formula, huge or
    // non-finite operands) gets rden = NaN: r, den and the window test below then fail on their own
    // and the sample goes down the plain-division path
    LabelRec rec;
    const bool safe = tame(den) && (ab == 0.0 || tame(ab));
    rec.ab = ab; rec.den = den;
    rec.rden = safe ? 1.0 / den : __longlong_as_double(0x7FF8000000000000ll);
    rec.safe = safe ? 1.0 : 0.0;
    s_lab[l] = rec;
  }
  if (INTERP == MRF_INTERP_LINEAR && threadIdx.x >= 64 && threadIdx.x < 192) {
    const int k = threadIdx.x - 64;
    s_log[k][0] = kLogTab[k][0]; s_log[k][1] = kLogTab[k][1];
  }
  if (threadIdx.x >= 32 && threadIdx.x < 64) {
    const int fr = threadIdx.x - 32;
    float wv[4] = {0.f, 0.f, 0.f, 0.f};
    if (INTERP == MRF_INTERP_CUBIC) cubic_weights(fr, wv); else linear_weights(fr, wv);
    s_wt[fr] = make_float4(wv[0], wv[1], wv[2], wv[3]);
  }
  if (STAGED && threadIdx.x == 0) mbar_init(&s_bar, 1);
  __syncthreads();
  if (STAGED && wrp == 0) {
    double cmin = 1.0e300, cmax = -1.0e300;
    for (int l = lane; l < L; l += 32) {
      const double c = s_lab[l].ab / s_lab[l].den;
      cmin = fmin(cmin, c); cmax = fmax(cmax, c);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      cmin = fmin(cmin, __shfl_xor_sync(0xffffffffu, cmin, o));
      cmax = fmax(cmax, __shfl_xor_sync(0xffffffffu, cmax, o));
    }
    if (lane == 0) { s_c[0] = cmin; s_c[1] = cmax; }
  }

  // per-pixel geometry :210-218
  const double xd = (double)x, yd = (double)y;
  const double vx = a.q.x - xd, vy = a.q.y - yd;
  const double dist = sqrt(vx * vx + vy * vy);
  const double nrm = fmax(0.5, dist);
  const double t1 = vx / nrm, t2 = vy / nrm;
  const double n = dist / g_mu;

  const long long pix = (long long)yl * a.w + x;
  bool dead = !valid;
  double c0 = 0.0, c1 = 0.0, c2 = 0.0;
  if (valid) {
    if (a.nonrigid && a.nonrigid[pix]) dead = true;    // Iz[rigidity==0] = 0   :286
    if (a.occluded && a.occluded[pix]) dead = true;    // Iz[occlusions==1] = 0 :285
    c0 = a.ref_ch[3 * pix]; c1 = a.ref_ch[3 * pix + 1]; c2 = a.ref_ch[3 * pix + 2];
  }
  s_ref[0][threadIdx.x] = c0; s_ref[1][threadIdx.x] = c1; s_ref[2][threadIdx.x] = c2;   // read back by the owner only

  int wx0 = 0, wy0 = 0, ww = 0, wh = 0;
  if (STAGED) {
    // Bound the tile's footprint: r ranges over n*[cmin, cmax]; H maps the segment to a segment.
    float bx0 = 3.0e38f, bx1 = -3.0e38f, by0 = 3.0e38f, by1 = -3.0e38f;
    __syncthreads();
    if (!dead) {
#pragma unroll
      for (int e = 0; e < 2; ++e) {
        const double r = s_c[e] * n;
        const double xn = xd + r * t1, yn = yd + r * t2;
        double X, Y; bool in;
        through_H(a.H, xn, yn, a.frame_h, a.w, X, Y, in);
        const float Xf = fminf(fmaxf((float)X, -1.0e6f), 1.0e6f);   // NaN -> -1e6 .. harmless
        const float Yf = fminf(fmaxf((float)Y, -1.0e6f), 1.0e6f);
        bx0 = fminf(bx0, Xf); bx1 = fmaxf(bx1, Xf);
        by0 = fminf(by0, Yf); by1 = fmaxf(by1, Yf);
      }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      bx0 = fminf(bx0, __shfl_xor_sync(0xffffffffu, bx0, o));
      bx1 = fmaxf(bx1, __shfl_xor_sync(0xffffffffu, bx1, o));
      by0 = fminf(by0, __shfl_xor_sync(0xffffffffu, by0, o));
      by1 = fmaxf(by1, __shfl_xor_sync(0xffffffffu, by1, o));
    }
    if (lane == 0) { s_red[wrp][0] = bx0; s_red[wrp][1] = bx1; s_red[wrp][2] = by0; s_red[wrp][3] = by1; }
    __syncthreads();
    if (threadIdx.x == 0) {
      for (int k = 1; k < kTY; ++k) {
        s_red[0][0] = fminf(s_red[0][0], s_red[k][0]); s_red[0][1] = fmaxf(s_red[0][1], s_red[k][1]);
        s_red[0][2] = fminf(s_red[0][2], s_red[k][2]); s_red[0][3] = fmaxf(s_red[0][3], s_red[k][3]);
      }
      int X0 = 0, Y0 = 0, W = 0, Hh = 0;
      if (s_red[0][0] <= s_red[0][1]) {
        const int ylo = max(0, a.nb_y0), yhi = min(a.frame_h, a.nb_y0 + a.nb_rows) - 1;
        int x0 = max(0, (int)floorf(s_red[0][0]) - 2), x1 = min(a.w - 1, (int)ceilf(s_red[0][1]) + 3);
        int y0 = max(ylo, (int)floorf(s_red[0][2]) - 2), y1 = min(yhi, (int)ceilf(s_red[0][3]) + 3);
        if (x1 >= x0 && y1 >= y0) {
          W = x1 - x0 + 1; Hh = y1 - y0 + 1;
          if (W > kWinMaxW) { x0 += (W - kWinMaxW) / 2; W = kWinMaxW; }
          const int hmax = a.win_cap / W;
          if (Hh > hmax) { y0 += (Hh - hmax) / 2; Hh = hmax; }
          X0 = x0; Y0 = y0;
        }
      }
      s_win[0] = X0; s_win[1] = Y0; s_win[2] = W; s_win[3] = Hh;
      if (W > 0) mbar_expect_tx(&s_bar, (uint32_t)(W * Hh) * 16u);
    }
    __syncthreads();
    wx0 = s_win[0]; wy0 = s_win[1]; ww = s_win[2]; wh = s_win[3];
    if (ww > 0) {
      for (int j = threadIdx.x; j < wh; j += kThreads)
        bulk_g2s(dyn_smem + (size_t)j * ww * 16, a.nb + ((long long)(wy0 + j - a.nb_y0) * a.w + wx0),
                 (uint32_t)ww * 16u, &s_bar);
      mbar_wait(&s_bar, 0);
    }
  }
  if (!valid) return;

  // first-tap positions whose whole footprint is inside the staged window AND strictly inside the
  // image (OpenCV's row-sum order): ix0 in [fx_lo, fx_lo + fx_span], iy0 likewise
  constexpr int kTaps = (INTERP == MRF_INTERP_CUBIC) ? 4 : 2;
  constexpr int kOff = (INTERP == MRF_INTERP_CUBIC) ? 1 : 0;
  int fx_lo = wx0, fy_lo = wy0;
  int fx_span = min(wx0 + ww, a.w) - kTaps - wx0;            // < 0: nothing staged
  int fy_span = min(wy0 + wh, a.frame_h) - kTaps - wy0;
  if (!STAGED || fx_span < 0 || fy_span < 0) { fx_lo = 0x7fff0000; fx_span = 0; fy_span = 0; }   // never matches
  uint32_t win_base = smem_u32(dyn_smem);                    // any valid footprint address for unused samples
  uint32_t win_s = win_base - (uint32_t)(wy0 * ww + wx0) * 16u;             // address of texel (0, 0)
  uint32_t wt_s = smem_u32(s_wt);
  const uint32_t log_s = smem_u32(s_log);
  uint32_t lab_s = smem_u32(s_lab);
  uint32_t ref_s = smem_u32(&s_ref[0][threadIdx.x]);
  const uint32_t pitch = (uint32_t)ww * 16u;
  // keep the shared-window addresses in registers (the compiler would otherwise rebuild them from
  // the CTA id on every iteration)
  if constexpr (FORM == 2) asm volatile("" : "+r"(win_s), "+r"(wt_s), "+r"(lab_s), "+r"(ref_s), "+r"(win_base));
  else asm volatile("" : "+r"(win_s), "+r"(wt_s), "+r"(lab_s), "+r"(ref_s));      // win_base: only the grouped form's stand-in address
  double xk = xd, yk = yd;                                  // pinned copies: not rebuilt from the thread id
  asm volatile("" : "+d"(xk), "+d"(yk));
  const float cost_dead = (float)rho_eval(a.rho, 0.0, a.rho_param);
  const bool lor_fast = (a.rho == MRF_RHO_LORENTZIAN) && tame(a.s2);
  const bool lor_unit = lor_fast && (a.s2 == 1.0);          // sigma = 1 (the default)
  const bool unit_w12 = (a.cw1 == 1.0) && (a.cw2 == 1.0);   // the gradient channels' weights (:371)
  // a pixel whose n is not tame poisons its own fast path the same way
  const double n_fast = ((n == 0.0) || tame(n)) ? n : __longlong_as_double(0x7FF8000000000000ll);
  float best = 0.f; int best_l = 0;
  // this pixel's first cost element; without a volume the stores land on the pixel's own min / argmin
  // slot (rewritten after the loop), so the loop needs no null test
  char* cost_p;
  uint32_t cost_step = (LAYOUT == MRF_LAYOUT_LHW_F32) ? (uint32_t)(a.plane_stride * 4) : 4u;
  if (a.cost) cost_p = (LAYOUT == MRF_LAYOUT_LHW_F32) ? (char*)((float*)a.cost + pix)
                                                      : (char*)((int32_t*)a.cost + (size_t)pix * L);
  else { cost_p = a.min_cost ? (char*)(a.min_cost + pix) : (char*)(a.argmin + pix); cost_step = 0; }
  uint32_t cost_off = 0;                                    // host side guarantees L * step < 4 GiB

  // Two forms of the label loop with identical arithmetic.  Grouped (FORM 2; bilinear): labels two at a time,
  // every phase one branch-free straight line over both, so the compiler overlaps the fp64 chains with each
  // other and with the gathers (bilinear, 4 taps: 0.207 -> 0.190 ms per frame at Sintel size).  Branching
  // (FORM 0; bicubic): one label at a time, branches round the sampling.  The bicubic kernel is co-limited by
  // instruction issue and shared-memory wavefronts (16 x LDS.128 per sample); the grouped form executes 6 %
  // fewer instructions there and halves the `wait` stalls, but its back-to-back gathers of two labels raise
  // the shared-memory stalls by more: 0.363 against 0.330 ms (profiles/r4a_kernel_forms.jsonl, r4b_*).
  if constexpr (FORM == 2) {
    // Grouped form: labels two at a time, three branch-free phases per pair (see sample_code above).
    const unsigned full = __activemask();
    const bool warp_dead = __all_sync(full, dead);            // e.g. inside a non-rigid region: nothing to sample
    WinGeom wg;
    wg.fx_lo = fx_lo; wg.fy_lo = fy_lo; wg.fx_span = fx_span; wg.fy_span = fy_span;
    wg.win_s = win_s; wg.win_base = win_base; wg.pitch = pitch; wg.wt_s = wt_s;
    auto body = [&](auto gc, int l) {
      constexpr int G = decltype(gc)::value;
      SampleCode sc[G];
#pragma unroll
      for (int g = 0; g < G; ++g)
        sample_code<INTERP>(a, lab_s + (uint32_t)(l + g) * 32u, n_fast, xk, yk, t1, t2, wg, dead, sc[g]);
      F3 J[G];
#pragma unroll
      for (int g = 0; g < G; ++g) J[g] = window_taps<INTERP>(sc[g], pitch);
      double r0, r1, r2;                                      // reference-frame channels of this pixel
      asm volatile("ld.shared.f64 %0, [%1];" : "=d"(r0) : "r"(ref_s));
      asm volatile("ld.shared.f64 %0, [%1];" : "=d"(r1) : "r"(ref_s + (uint32_t)(kThreads * 8)));
      asm volatile("ld.shared.f64 %0, [%1];" : "=d"(r2) : "r"(ref_s + (uint32_t)(kThreads * 16)));
      double Iz[G], cst[G];
#pragma unroll
      for (int g = 0; g < G; ++g) {
        const double e0 = ((double)J[g].x - r0) * a.cw0;      // :274
        double e1 = (double)J[g].y - r1, e2 = (double)J[g].z - r2;
        if (!unit_w12) { e1 = e1 * a.cw1; e2 = e2 * a.cw2; }  // x * 1.0 == x: the default weights (0.01, 1, 1) skip two products
        Iz[g] = div3((e0 + e1) + e2, kThird);          // .mean(axis=2): sum / 3.0, exact
      }
      if (lor_unit) {
#pragma unroll
        for (int g = 0; g < G; ++g) cst[g] = log_ge1<INTERP>(1.0 + Iz[g] * Iz[g], log_s);                        // :52, sigma = 1
      } else if (lor_fast) {
#pragma unroll
        for (int g = 0; g < G; ++g) cst[g] = log_ge1<INTERP>(1.0 + div_rc(Iz[g] * Iz[g], a.s2, a.rs2), log_s);   // :52
      } else {
#pragma unroll
        for (int g = 0; g < G; ++g)
          cst[g] = (a.rho == MRF_RHO_NONE) ? Iz[g] : rho_eval(a.rho, Iz[g] * Iz[g], a.rho_param);  // :302
      }
      float cf[G];
      bool any_cold = false;
#pragma unroll
      for (int g = 0; g < G; ++g) { cf[g] = sc[g].fast ? (float)cst[g] : cost_dead; any_cold = any_cold || sc[g].cold; }
      if (any_cold) {                                         // rare: pole / degenerate operands, window or border miss
#pragma unroll
        for (int g = 0; g < G; ++g)
          if (sc[g].cold) cf[g] = cold_cost<INTERP>(a, lab_s + (uint32_t)(l + g) * 32u, n, xk, yk, t1, t2, ref_s, cost_dead,
                                                    WinRef{STAGED ? s_win : nullptr, win_s, pitch, wt_s, log_s});
      }
#pragma unroll
      for (int g = 0; g < G; ++g) {
        if (LAYOUT == MRF_LAYOUT_LHW_F32) __stcs((float*)(cost_p + cost_off), cf[g]);
        else *(int32_t*)(cost_p + cost_off) = (int32_t)(cf[g] * a.i32_scale);
        cost_off += cost_step;
        if ((g == 0 && l == 0) || cf[g] < best) { best = cf[g]; best_l = l + g; }   // np.argmin: first minimum
      }
    };
    if (warp_dead) {
      for (int l = 0; l < L; ++l, cost_off += cost_step) {
        if (LAYOUT == MRF_LAYOUT_LHW_F32) __stcs((float*)(cost_p + cost_off), cost_dead);
        else *(int32_t*)(cost_p + cost_off) = (int32_t)(cost_dead * a.i32_scale);
      }
      best = cost_dead; best_l = 0;
    } else {
      constexpr int G = (INTERP == MRF_INTERP_LINEAR) ? MRF_GROUP_LINEAR : 2;
      int l = 0;
      for (; l + G <= L; l += G) body(IntC<G>{}, l);
      for (; l < L; ++l) body(IntC<1>{}, l);
    }
  } else {
    for (int l = 0; l < L; ++l, cost_off += cost_step) {
      float cf = cost_dead;
      if (!dead) {
        double ab, dl, rdl;
        asm volatile("ld.shared.v2.f64 {%0, %1}, [%2];" : "=d"(ab), "=d"(dl) : "r"(lab_s + (uint32_t)l * 32u));
        asm volatile("ld.shared.f64 %0, [%1];" : "=d"(rdl) : "r"(lab_s + (uint32_t)l * 32u + 16u));
        // r = A*B*n / (B*A/mu - 1)  :221
        const double r = div_rc(ab * n_fast, dl, rdl);
        const double xn = xk + r * t1, yn = yk + r * t2;      // :224-235
        // interpolate_through_H.py:31-35
        const double den = (xn * a.H.m[6] + yn * a.H.m[7]) + a.H.m[8];
        const double Xn = (xn * a.H.m[0] + yn * a.H.m[1]) + a.H.m[2];
        const double Yn = (xn * a.H.m[3] + yn * a.H.m[4]) + a.H.m[5];
        const double rd = rcp_rn_tame(den);
        const double X = div_rc(Xn, den, rd);
        const double Y = div_rc(Yn, den, rd);
        // every condition under which the sequences above are not proven exact
        const bool exact = tame_divisor(den);
        const bool inside = (X >= 0.0) && (X <= a.wm1) && (Y >= 0.0) && (Y <= a.hm1);   // :39
        if (!exact) {
          cf = slow_sample_cost<INTERP>(a, ab, dl, n, xk, yk, t1, t2, s_ref[0][threadIdx.x], s_ref[1][threadIdx.x],
                                        s_ref[2][threadIdx.x]);
        } else if (inside) {
          const float xf = (float)X, yf = (float)Y;           // interpolation.py:77-78
          // cv2.remap's quantisation; inside the frame |coord*32| < 2^31 and the int16 saturation
          // cannot trigger for any position the window test below accepts
          const int sx = __float2int_rn(xf * 32.0f), sy = __float2int_rn(yf * 32.0f);
          const int ix0 = (sx >> 5) - kOff, iy0 = (sy >> 5) - kOff;
          F3 J;
          if ((unsigned)(ix0 - fx_lo) <= (unsigned)fx_span && (unsigned)(iy0 - fy_lo) <= (unsigned)fy_span) {
            const uint32_t p0 = win_s + (uint32_t)iy0 * pitch + (uint32_t)ix0 * 16u;
            float wxa[4], wya[4];
            lds128(wt_s + (uint32_t)(sx & 31) * 16u, wxa[0], wxa[1], wxa[2], wxa[3]);
            lds128(wt_s + (uint32_t)(sy & 31) * 16u, wya[0], wya[1], wya[2], wya[3]);
  #pragma unroll
            for (int j = 0; j < kTaps; ++j) {
              F3 row;
  #pragma unroll
              for (int i = 0; i < kTaps; ++i) {
                const float wgt = wya[j] * wxa[i];
                float tx, ty, tz, tw;
                lds128(p0 + (uint32_t)j * pitch + (uint32_t)i * 16u, tx, ty, tz, tw);
                const float px = tx * wgt, py = ty * wgt, pz = tz * wgt;
                if (INTERP == MRF_INTERP_CUBIC) {               // row sums, then rows
                  if (i == 0) { row.x = px; row.y = py; row.z = pz; }
                  else        { row.x = row.x + px; row.y = row.y + py; row.z = row.z + pz; }
                } else {                                        // linear: one running sum
                  if (i == 0 && j == 0) { J.x = px; J.y = py; J.z = pz; }
                  else { J.x = J.x + px; J.y = J.y + py; J.z = J.z + pz; }
                }
              }
              if (INTERP == MRF_INTERP_CUBIC) {
                if (j == 0) J = row;
                else { J.x = J.x + row.x; J.y = J.y + row.y; J.z = J.z + row.z; }
              }
            }
          } else if (STAGED && border_in_window<INTERP>(s_win, a.w, a.frame_h, sx, sy)) {
            // footprint at the image border, clamped taps inside the staged window
            J = sample_window_border<INTERP>(win_s, pitch, wt_s, a.w, a.frame_h, sx, sy);
          } else {
            // footprint leaves the staged window: gather from global
            // memory; for row-band inputs check the halo (taps of an in-frame sample: rows iy-1..iy+2)
            const int iy = sy >> 5;
            const int ya = max(0, min(a.frame_h - 1, iy - 1)), yb = max(0, min(a.frame_h - 1, iy + 2));
            if (ya < a.nb_y0 || yb > a.nb_y0 + a.nb_rows - 1) {
              if (a.halo_miss) atomicAdd(a.halo_miss, 1);
              J.x = J.y = J.z = 0.f;
            } else {
              J = sample_band<INTERP>(a.nb, a.w, a.nb_y0, a.frame_h, xf, yf);
            }
          }
          double r0, r1, r2;                                  // reference-frame channels of this pixel
          asm volatile("ld.shared.f64 %0, [%1];" : "=d"(r0) : "r"(ref_s));
          asm volatile("ld.shared.f64 %0, [%1];" : "=d"(r1) : "r"(ref_s + (uint32_t)(kThreads * 8)));
          asm volatile("ld.shared.f64 %0, [%1];" : "=d"(r2) : "r"(ref_s + (uint32_t)(kThreads * 16)));
          const double e0 = ((double)J.x - r0) * a.cw0;       // :274
          double e1 = (double)J.y - r1, e2 = (double)J.z - r2;
          if (!unit_w12) { e1 = e1 * a.cw1; e2 = e2 * a.cw2; } // x * 1.0 == x: the default weights (0.01, 1, 1) skip two products
          const double Iz = div3((e0 + e1) + e2, kThird);   // .mean(axis=2): sum / 3.0, exact
          double cst;
          if (lor_unit) {
            cst = log_ge1<INTERP>(1.0 + Iz * Iz, log_s);                     // sigma = 1: x / sigma**2 == x exactly  :52
          } else if (lor_fast) {
            const double z = div_rc(Iz * Iz, a.s2, a.rs2);    // x / sigma**2  :52
            cst = log_ge1<INTERP>(1.0 + z, log_s);
          } else {
            cst = (a.rho == MRF_RHO_NONE) ? Iz : rho_eval(a.rho, Iz * Iz, a.rho_param);   // :302
          }
          cf = (float)cst;
        }
      }
      if (LAYOUT == MRF_LAYOUT_LHW_F32) __stcs((float*)(cost_p + cost_off), cf);
      else *(int32_t*)(cost_p + cost_off) = (int32_t)(cf * a.i32_scale);
      if (l == 0 || cf < best) { best = cf; best_l = l; }     // np.argmin: first minimum
    }
  }
  if (a.min_cost) a.min_cost[pix] = best;
  if (a.argmin) a.argmin[pix] = best_l;
}

template <int INTERP, bool STAGED, int OCC, int LAYOUT, int FORM>
__global__ void __launch_bounds__(kThreads, OCC)
k_cost_volume(const __grid_constant__ CVArgs a0, const __grid_constant__ CVArgs a1) {
  // blockIdx.z selects the neighbour frame: both frames of a triplet share one grid, so the tail of one
  // frame's blocks overlaps the head of the other's (grid.z = 1: a0 only)
  extern __shared__ __align__(128) unsigned char dyn_smem[];
  if constexpr (FORM == 2) {
    // one copy of the body per argument block: every per-frame scalar (H, the frame bounds, the channel
    // weights) is then a constant-bank operand of the instruction that uses it -- no loads, no registers
    if (blockIdx.z == 0) cost_volume_block<INTERP, STAGED, LAYOUT, FORM>(a0, dyn_smem);
    else cost_volume_block<INTERP, STAGED, LAYOUT, FORM>(a1, dyn_smem);
  } else {
    cost_volume_block<INTERP, STAGED, LAYOUT, FORM>((blockIdx.z == 0) ? a0 : a1, dyn_smem);
  }
}

// ---------------------------------------------------------------------------------------------
// Auto-sigma mode of the cost volume (utils/robust_functions.py:55-58): the Lorentzian's sigma is
// max(1e-6, 1.4826 * median |Iz|) over the WHOLE residual plane of a label, so a plane needs its fp64
// residual first, then an exact median, then rho.  Two plain kernels per label (the reference arithmetic
// written naively: IEEE divisions, library log); a second mode, not the tuned path.
template <int INTERP>
__global__ void __launch_bounds__(256)
k_residual_plane(const __grid_constant__ CVArgs a, int label, double* __restrict__ Iz_out, double* __restrict__ abs_out) {
  const long long n = (long long)a.rows * a.w;
  double g_mu = a.mu, g_B = a.B, g_scale = a.label_scale;
  const double* g_labels = a.labels;
  if (a.state) {
    g_mu = a.state[MRF_STATE_MU + a.frame];
    g_B = a.state[MRF_STATE_B + a.frame];
    g_scale = (a.frame == 0) ? a.state[MRF_STATE_MU] / a.state[MRF_STATE_MU + 1] : 1.0;
    g_labels = a.state + MRF_STATE_LABELS;
  }
  const double al = g_labels[label] * g_scale;
  const double ab = al * g_B;
  const double den_l = ab / g_mu - 1.0;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const int yl = (int)(i / a.w);
    const int x = (int)(i - (long long)yl * a.w);
    const double xd = (double)x, yd = (double)(yl + a.y0);
    double Iz = 0.0;
    const bool dead = (a.nonrigid && a.nonrigid[i]) || (a.occluded && a.occluded[i]);
    if (!dead) {
      const double vx = a.q.x - xd, vy = a.q.y - yd;                         // :210-218
      const double dist = sqrt(vx * vx + vy * vy);
      const double nrm = fmax(0.5, dist);
      const double r = (ab * (dist / g_mu)) / den_l;                         // :221
      const double xn = xd + r * (vx / nrm), yn = yd + r * (vy / nrm);
      double X, Y; bool inside;
      through_H(a.H, xn, yn, a.frame_h, a.w, X, Y, inside);
      if (inside) {
        const float xf = (float)X, yf = (float)Y;
        int iy, fy; cv_quantise(yf, iy, fy);
        const int lo = (INTERP == MRF_INTERP_CUBIC) ? -1 : 0, hi = (INTERP == MRF_INTERP_CUBIC) ? 2 : 1;
        const int ya = max(0, min(a.frame_h - 1, iy + lo)), yb = max(0, min(a.frame_h - 1, iy + hi));
        F3 J;
        if (ya < a.nb_y0 || yb > a.nb_y0 + a.nb_rows - 1) {
          if (a.halo_miss) atomicAdd(a.halo_miss, 1);
          J.x = J.y = J.z = 0.f;
        } else {
          FetchBand f{a.nb, a.w, a.nb_y0};
          J = sample<INTERP>(f, a.frame_h, a.w, xf, yf);
        }
        const double e0 = ((double)J.x - a.ref_ch[3 * i]) * a.cw0;           // :274
        const double e1 = ((double)J.y - a.ref_ch[3 * i + 1]) * a.cw1;
        const double e2 = ((double)J.z - a.ref_ch[3 * i + 2]) * a.cw2;
        Iz = ((e0 + e1) + e2) / 3.0;
      }
    }
    Iz_out[i] = Iz;
    abs_out[i] = fabs(Iz);
  }
}

// cost = rho(Iz^2) with sigma = max(1e-6, 1.4826 * median) read from the device; fp32 plane, running min / argmin
__global__ void __launch_bounds__(256)
k_auto_rho_plane(const double* __restrict__ Iz, long long n, const double* __restrict__ median_abs, int label,
                 float* __restrict__ cost_plane, float* __restrict__ min_cost, int32_t* __restrict__ argmin) {
  const double sg = fmax(1e-6, 1.4826 * median_abs[0]);                      // robust_functions.py:29,55
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const double x = Iz[i] * Iz[i];
    const float cf = (float)(sg * log(1.0 + (0.5 * x) / (sg * sg)));        // :58
    if (cost_plane) cost_plane[i] = cf;
    if (min_cost && (label == 0 || cf < min_cost[i])) {                      // np.argmin: first minimum
      min_cost[i] = cf;
      if (argmin) argmin[i] = label;
    }
  }
}

}  // namespace mrf

using namespace mrf;

struct CVHost {            // the device-pointer half of one frame's arguments
  const float* nb_texels; const uint8_t* occluded; void* cost; float* min_cost; int32_t* argmin;
};

static int cv_fill(const mrf_costvol_params* p, const double* ref_ch, const double* labels, const uint8_t* nonrigid,
                   const CVHost& h, int* halo_miss, const double* dev_state, CVArgs* out) {
  MRF_CHECK_ARG(p && ref_ch && h.nb_texels && (labels || dev_state));
  MRF_CHECK_ARG(!dev_state || p->state_frame == 0 || p->state_frame == 1);
  MRF_CHECK_ARG(p->rows > 0 && p->w > 0 && p->frame_h >= p->y0 + p->rows && p->y0 >= 0);
  MRF_CHECK_ARG(p->nb_rows > 0 && p->nb_y0 >= 0 && p->nb_y0 + p->nb_rows <= p->frame_h);
  MRF_CHECK_ARG(p->n_labels >= 1 && p->n_labels <= kMaxLabels);
  MRF_CHECK_ARG(p->interp == MRF_INTERP_CUBIC || p->interp == MRF_INTERP_LINEAR);
  MRF_CHECK_ARG(p->rho >= MRF_RHO_LORENTZIAN && p->rho <= MRF_RHO_NONE);
  MRF_CHECK_ARG(p->layout == MRF_LAYOUT_LHW_F32 || p->layout == MRF_LAYOUT_HWL_I32);
  MRF_CHECK_ARG(h.cost || h.min_cost || h.argmin);
  MRF_CHECK_ARG(((uintptr_t)h.nb_texels & 15) == 0);
  const long long plane_stride = p->cost_plane_stride > 0 ? p->cost_plane_stride : (long long)p->rows * p->w;
  MRF_CHECK_ARG(plane_stride >= (long long)p->rows * p->w);
  MRF_CHECK_ARG((double)plane_stride * 4.0 * p->n_labels < 4294967296.0);   /* 32-bit store offsets */
  CVArgs& a = *out;
  a.rows = p->rows; a.w = p->w; a.y0 = p->y0; a.frame_h = p->frame_h;
  a.nb_rows = p->nb_rows; a.nb_y0 = p->nb_y0;
  a.H = mat3(p->H); a.q = pt2(p->q);
  a.mu = p->mu; a.B = p->B; a.label_scale = p->label_scale;
  a.L = p->n_labels; a.rho = p->rho; a.rho_param = p->rho_param;
  a.cw0 = p->cw[0]; a.cw1 = p->cw[1]; a.cw2 = p->cw[2];
  a.layout = p->layout; a.i32_scale = (float)p->i32_scale;
  a.ref_ch = ref_ch; a.nb = (const float4*)h.nb_texels; a.labels = labels;
  a.nonrigid = nonrigid; a.occluded = h.occluded;
  a.cost = h.cost; a.min_cost = h.min_cost; a.argmin = h.argmin; a.halo_miss = halo_miss;
  a.plane_stride = plane_stride;
  a.state = dev_state; a.frame = p->state_frame;
  a.wm1 = (double)(p->w - 1); a.hm1 = (double)(p->frame_h - 1);
  a.s2 = p->rho_param * p->rho_param; a.rs2 = 1.0 / a.s2;
  const int occ = (p->variant & MRF_VARIANT_OCC2) ? 2 : ((p->variant & MRF_VARIANT_OCC4) ? 4 : 3);
  a.win_cap = (occ == 4) ? kWinCap4 : kWinCap;
  return 0;
}

// launch for one frame (nz = 1) or both frames of a triplet in one grid (nz = 2; p = the first frame's params:
// shape, sampler, layout and variant are common to both)
static int cv_launch(const mrf_costvol_params* p, const CVArgs& a, const CVArgs& b, int nz, mrf_stream_t stream) {
  const dim3 grid(cdiv(p->w, kTX), cdiv(p->rows, kTY), (unsigned)nz);
  const bool staged = ((p->variant & 1) != MRF_VARIANT_GATHER);
  const int occ = (p->variant & MRF_VARIANT_OCC2) ? 2 : ((p->variant & MRF_VARIANT_OCC4) ? 4 : 3);
  // the grouped form takes its samples from the staged window; it is the default for bilinear sampling
  // (MRF_VARIANT_SINGLE selects the one-label form) and opt-in for bicubic (MRF_VARIANT_GROUPED)
  const bool grouped = (p->interp == MRF_INTERP_LINEAR) ? !(p->variant & MRF_VARIANT_SINGLE) : (p->variant & MRF_VARIANT_GROUPED) != 0;
  const int form = (grouped && staged) ? 2 : 0;
  const size_t smem = staged ? (size_t)a.win_cap * 16 : 0;
  cudaError_t e = cudaSuccess;
#define MRF_LAUNCH(I, ST, OC, LY, WFv)                                                                         \
  do {                                                                                                         \
    /* the opt-in is per device and per instantiation: remember it per device (one bit each) */               \
    static std::atomic<unsigned long long> attr_set{0};                                                       \
    int dev_ = 0;                                                                                              \
    if (smem) cudaGetDevice(&dev_);                                                                            \
    if (smem && !((attr_set.load() >> (dev_ & 63)) & 1ull)) {                                                  \
      e = cudaFuncSetAttribute(k_cost_volume<I, ST, OC, LY, WFv>, cudaFuncAttributeMaxDynamicSharedMemorySize, \
                               (int)smem);                                                                     \
      if (e == cudaSuccess) attr_set.fetch_or(1ull << (dev_ & 63));                                            \
    }                                                                                                          \
    if (e == cudaSuccess) k_cost_volume<I, ST, OC, LY, WFv><<<grid, kThreads, smem, S(stream)>>>(a, b);        \
  } while (0)
#define MRF_LAUNCH_LY(I, ST, OC, WFv) do { if (p->layout == MRF_LAYOUT_LHW_F32) MRF_LAUNCH(I, ST, OC, MRF_LAYOUT_LHW_F32, WFv); else MRF_LAUNCH(I, ST, OC, MRF_LAYOUT_HWL_I32, WFv); } while (0)
#define MRF_LAUNCH_ST(I, OC, WFv) do { if (staged) MRF_LAUNCH_LY(I, true, OC, WFv); else MRF_LAUNCH_LY(I, false, OC, WFv); } while (0)
#define MRF_LAUNCH_OC(I, WFv) do { if (occ == 2) MRF_LAUNCH_ST(I, 2, WFv); else if (occ == 4) MRF_LAUNCH_ST(I, 4, WFv); else MRF_LAUNCH_ST(I, 3, WFv); } while (0)
#define MRF_LAUNCH_G(I) do { if (occ == 2) MRF_LAUNCH_LY(I, true, 2, 2); else if (occ == 4) MRF_LAUNCH_LY(I, true, 4, 2); else MRF_LAUNCH_LY(I, true, 3, 2); } while (0)
  if (p->interp == MRF_INTERP_CUBIC) {
    if (form == 2) MRF_LAUNCH_G(MRF_INTERP_CUBIC); else MRF_LAUNCH_OC(MRF_INTERP_CUBIC, 0);
  } else {
    if (form == 2) MRF_LAUNCH_G(MRF_INTERP_LINEAR); else MRF_LAUNCH_OC(MRF_INTERP_LINEAR, 0);
  }
#undef MRF_LAUNCH_G
#undef MRF_LAUNCH_LY
#undef MRF_LAUNCH_OC
#undef MRF_LAUNCH_ST
#undef MRF_LAUNCH
  if (e != cudaSuccess) return fail((int)e, "mrf_cost_volume: %s", cudaGetErrorString(e));
  return check_launch("mrf_cost_volume");
}

extern "C" int mrf_cost_volume(const mrf_costvol_params* p, const double* ref_ch, const float* nb_texels,
                               const double* labels, const uint8_t* nonrigid, const uint8_t* occluded,
                               void* cost, float* min_cost, int32_t* argmin, int* halo_miss,
                               const double* dev_state, mrf_stream_t stream) {
  CVArgs a;
  const CVHost h{nb_texels, occluded, cost, min_cost, argmin};
  const int rc = cv_fill(p, ref_ch, labels, nonrigid, h, halo_miss, dev_state, &a);
  if (rc) return rc;
  return cv_launch(p, a, a, 1, stream);
}

extern "C" int mrf_cost_volume_pair(const mrf_costvol_params* p0, const mrf_costvol_params* p1, const double* ref_ch,
                                    const float* nb_texels0, const float* nb_texels1, const double* labels,
                                    const uint8_t* nonrigid, const uint8_t* occluded0, const uint8_t* occluded1,
                                    void* cost0, void* cost1, float* min_cost0, float* min_cost1, int32_t* argmin0,
                                    int32_t* argmin1, int* halo_miss, const double* dev_state, mrf_stream_t stream) {
  MRF_CHECK_ARG(p0 && p1);
  MRF_CHECK_ARG(p0->rows == p1->rows && p0->w == p1->w && p0->y0 == p1->y0 && p0->frame_h == p1->frame_h);
  MRF_CHECK_ARG(p0->interp == p1->interp && p0->layout == p1->layout && p0->variant == p1->variant && p0->n_labels == p1->n_labels);
  CVArgs a, b;
  const CVHost h0{nb_texels0, occluded0, cost0, min_cost0, argmin0}, h1{nb_texels1, occluded1, cost1, min_cost1, argmin1};
  int rc = cv_fill(p0, ref_ch, labels, nonrigid, h0, halo_miss, dev_state, &a);
  if (!rc) rc = cv_fill(p1, ref_ch, labels, nonrigid, h1, halo_miss, dev_state, &b);
  if (rc) return rc;
  return cv_launch(p0, a, b, 2, stream);
}

// One label plane of the auto-sigma cost volume: residual (fp64) + |residual| into scratch planes.
extern "C" int mrf_cost_residual_plane(const mrf_costvol_params* p, const double* ref_ch, const float* nb_texels,
                                       const double* labels, const uint8_t* nonrigid, const uint8_t* occluded, int label,
                                       double* residual, double* abs_residual, int* halo_miss, const double* dev_state,
                                       mrf_stream_t stream) {
  MRF_CHECK_ARG(residual && abs_residual && p && label >= 0 && label < p->n_labels);
  CVArgs a;
  const CVHost h{nb_texels, occluded, residual, nullptr, nullptr};
  const int rc = cv_fill(p, ref_ch, labels, nonrigid, h, halo_miss, dev_state, &a);
  if (rc) return rc;
  const long long n = (long long)p->rows * p->w;
  const unsigned blocks = (unsigned)((n + 255) / 256 < (long long)kSMs * 8 ? (n + 255) / 256 : (long long)kSMs * 8);
  if (p->interp == MRF_INTERP_CUBIC) k_residual_plane<MRF_INTERP_CUBIC><<<blocks, 256, 0, S(stream)>>>(a, label, residual, abs_residual);
  else k_residual_plane<MRF_INTERP_LINEAR><<<blocks, 256, 0, S(stream)>>>(a, label, residual, abs_residual);
  return check_launch("mrf_cost_residual_plane");
}

// Lorentzian with sigma = max(1e-6, 1.4826 * *median_abs_dev) applied to a residual plane; updates the running
// min / argmin over labels (call with label = 0, 1, 2, ... in order)
extern "C" int mrf_cost_auto_rho_plane(const double* residual, size_t n, const double* median_abs_dev, int label,
                                       float* cost_plane, float* min_cost, int32_t* argmin, mrf_stream_t stream) {
  MRF_CHECK_ARG(residual && median_abs_dev && n > 0 && label >= 0 && (cost_plane || min_cost));
  MRF_CHECK_ARG(!argmin || min_cost);
  const unsigned blocks = (unsigned)((n + 255) / 256 < (size_t)kSMs * 8 ? (n + 255) / 256 : (size_t)kSMs * 8);
  k_auto_rho_plane<<<blocks, 256, 0, S(stream)>>>(residual, (long long)n, median_abs_dev, label, cost_plane, min_cost, argmin);
  return check_launch("mrf_cost_auto_rho_plane");
}
