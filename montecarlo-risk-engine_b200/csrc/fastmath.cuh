// Branch-free FP64 elementary functions for the fused Monte Carlo kernels.
//
// Why not libdevice: ncu on the first version of irc_main_kernel showed the kernel
// issue-bound (27 % FP64 instructions): libdevice's exp/log/sincospi carry special-case
// branches (no interleaving of independent paths across BSSY/BSYNC) and materialise every
// polynomial coefficient as a 64-bit immediate (2 UMOV per DFMA).  These versions
//   * are straight-line code (select instead of branch), so the two paths a thread owns
//     interleave and hide each other's DFMA latency,
//   * assume the argument ranges the Monte Carlo actually produces (documented per
//     function) instead of handling NaN / Inf / denormals,
//   * stay within ~2 ulp (tests/test_fastmath_gpu.py), far inside the 1e-10 parity budget.
#pragma once
#include <cstdint>

namespace mcre {

#ifndef MCRE_HOST_EMU   // (tests/host/fastmath_host.cpp compiles this header for the CPU with its own models of these two)
__device__ __forceinline__ double fm_rcp_approx(double x) {
  double y;
  asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(x));
  return y;
}
__device__ __forceinline__ double fm_rsqrt_approx(double x) {
  double y;
  asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(x));
  return y;
}
#endif

// a / b for normal, finite b (|b| in [1e-300, 1e300]): MUFU seed + 2 Newton steps + residual fix.
__device__ __forceinline__ double fm_div(double a, double b) {
  double y = fm_rcp_approx(b);
  double e = fma(-b, y, 1.0);
  y = fma(y, e, y);
  e = fma(-b, y, 1.0);
  y = fma(y, e, y);
  double q = a * y;
  return fma(fma(-b, q, a), y, q);
}

// sqrt(x) for x > 0 (normal range): no zero guard.  y = MUFU.RSQ64H(x) reads the high word only (relative
// error e <= 2^-20); g = x y = sqrt(x) (1 + e), r = 1 - g y = -2e - e^2 and
// sqrt(x) = g (1 - r)^(-1/2) = g + g r (1/2 + 3/8 r) + O(r^3): 5 FP64 instructions, 2^-60 relative truncation,
// the rounding of g is absorbed by r.
__device__ __forceinline__ double fm_sqrt_pos(double x) {
  const double y = fm_rsqrt_approx(x);
  const double g = x * y;
  const double r = fma(-g, y, 1.0);
  return fma(g * r, fma(r, 0.375, 0.5), g);
}

// sqrt(x) for x >= 0 (0 -> 0), normal range.
__device__ __forceinline__ double fm_sqrt(double x) {
  const double g = fm_sqrt_pos(x > 0.0 ? x : 1.0);
  return x > 0.0 ? g : 0.0;
}


// exp(x) for |x| <= 700.  Cody-Waite reduction, degree-13 Taylor on |r| <= ln2/2
// (truncation 4e-18 relative), scaling by two exact powers of two.
__device__ __forceinline__ double fm_exp(double x) {
  const double L2E = 1.4426950408889634, LN2_HI = 6.93147180369123816490e-01, LN2_LO = 1.90821492927058770002e-10;
  const double MAGIC = 6755399441055744.0;  // 1.5 * 2^52
  double t = fma(x, L2E, MAGIC);
  const int n = __double2loint(t);
  t -= MAGIC;
  double r = fma(t, -LN2_HI, x);
  r = fma(t, -LN2_LO, r);
  double p = 1.6059043836821613e-10;           // 1/13!
  p = fma(p, r, 2.08767569878681e-09);         // 1/12!
  p = fma(p, r, 2.505210838544172e-08);        // 1/11!
  p = fma(p, r, 2.755731922398589e-07);        // 1/10!
  p = fma(p, r, 2.7557319223985893e-06);       // 1/9!
  p = fma(p, r, 2.48015873015873e-05);         // 1/8!
  p = fma(p, r, 1.984126984126984e-04);        // 1/7!
  p = fma(p, r, 1.3888888888888889e-03);       // 1/6!
  p = fma(p, r, 8.333333333333333e-03);        // 1/5!
  p = fma(p, r, 4.1666666666666664e-02);       // 1/4!
  p = fma(p, r, 1.6666666666666666e-01);       // 1/3!
  p = fma(p, r, 0.5);
  p = fma(p, r, 1.0);
  p = fma(p, r, 1.0);
  // 2^n = 2^(n/2) * 2^(n - n/2): both factors normal for |n| <= 1020
  const int n1 = n >> 1, n2 = n - n1;
  const double s1 = __hiloint2double((n1 + 1023) << 20, 0), s2 = __hiloint2double((n2 + 1023) << 20, 0);
  return p * s1 * s2;
}

// log(u) for u in [2^-60, 2): mantissa/exponent split, s = (m-1)/(m+1), atanh series to s^21.
__device__ __forceinline__ double fm_log(double u) {
  int hi = __double2hiint(u);
  const int lo = __double2loint(u);
  int e = (hi >> 20) - 1023;
  hi = (hi & 0x000fffff) | 0x3ff00000;       // m in [1, 2)
  const bool big = hi >= 0x3ff6a09f;          // m > sqrt(2) -> m/2, e+1 (keeps |s| <= 0.1716)
  hi = big ? hi - 0x00100000 : hi;
  e = big ? e + 1 : e;
  const double m = __hiloint2double(hi, lo);
  const double f = m - 1.0;
  const double s = fm_div(f, m + 1.0);
  const double z = s * s;
  double p = 4.7619047619047616e-02;            // 1/21
  p = fma(p, z, 5.2631578947368418e-02);        // 1/19
  p = fma(p, z, 5.8823529411764705e-02);        // 1/17
  p = fma(p, z, 6.6666666666666666e-02);        // 1/15
  p = fma(p, z, 7.6923076923076927e-02);        // 1/13
  p = fma(p, z, 9.0909090909090912e-02);        // 1/11
  p = fma(p, z, 1.1111111111111111e-01);        // 1/9
  p = fma(p, z, 1.4285714285714285e-01);        // 1/7
  p = fma(p, z, 2.0000000000000001e-01);        // 1/5
  p = fma(p, z, 3.3333333333333331e-01);        // 1/3
  // log m = 2 s + 2 s z p
  const double two_s = s + s;
  const double lm = fma(two_s * z, p, two_s);
  const double ed = (double)e;
  const double LN2_HI = 6.93147180369123816490e-01, LN2_LO = 1.90821492927058770002e-10;
  return fma(ed, LN2_HI, fma(ed, LN2_LO, lm));
}

// (sin(2 pi u), cos(2 pi u)) for u in [0, 1): reduce a = 2u to a quadrant of width 1/2,
// Taylor in (pi r) with |pi r| <= pi/4 (sin to x^17, cos to x^18), then swap / negate.
__device__ __forceinline__ void fm_sincos2pi(double u, double &sn, double &cs) {
  const double a = u + u;                       // [0, 2)
  const double MAGIC = 6755399441055744.0;
  double t = fma(a, 2.0, MAGIC);
  const int q = __double2loint(t);             // nearest integer to 2a: 0..4
  t -= MAGIC;
  const double r = fma(t, -0.5, a);             // [-1/4, 1/4], exact
  const double x = r * 3.141592653589793116 + r * 1.2246467991473532e-16;
  const double z = x * x;
  double ps = 2.8114572543455206e-15;           // 1/17!
  ps = fma(ps, z, -7.6471637318198164e-13);     // -1/15!
  ps = fma(ps, z, 1.6059043836821613e-10);      // 1/13!
  ps = fma(ps, z, -2.505210838544172e-08);      // -1/11!
  ps = fma(ps, z, 2.7557319223985893e-06);      // 1/9!
  ps = fma(ps, z, -1.984126984126984e-04);      // -1/7!
  ps = fma(ps, z, 8.333333333333333e-03);       // 1/5!
  ps = fma(ps, z, -1.6666666666666666e-01);     // -1/3!
  const double s0 = fma(x * z, ps, x);
  double pc = 1.5619206968586225e-16;           // 1/18!
  pc = fma(pc, z, -4.7794773323873853e-14);     // -1/16!
  pc = fma(pc, z, 1.1470745597729725e-11);      // 1/14!
  pc = fma(pc, z, -2.08767569878681e-09);       // -1/12!
  pc = fma(pc, z, 2.755731922398589e-07);       // 1/10!
  pc = fma(pc, z, -2.48015873015873e-05);       // -1/8!
  pc = fma(pc, z, 1.3888888888888889e-03);      // 1/6!
  pc = fma(pc, z, -4.1666666666666664e-02);     // -1/4!
  pc = fma(pc, z, 0.5);
  const double c0 = fma(-z, pc, 1.0);
  // angle = q*pi/2 + x : rotate by quadrant
  const bool swap = q & 1;
  const double sa = swap ? c0 : s0, ca = swap ? s0 : c0;
  sn = (q & 2) ? -sa : sa;
  cs = ((q + 1) & 2) ? -ca : ca;
}


#if defined(MCRE_FAST_MATH) && MCRE_FAST_MATH >= 2
// =====================================================================================
// Table-driven variants.  History: the v1 kernel spent ~100 of its 148 FP64 instructions per
// path-step in the long Taylor polynomials above and 77 instructions materialising their
// coefficients (profiles/r01_irc_main_v1_*); round 1 reduced the arguments against small tables
// in SHARED memory.  Round 2 (profiles/r02_issue_mix_probe.log): an FP64 instruction holds the
// issue port of its scheduler for ~2.2 cycles and integer instructions do NOT issue in its shadow,
// so every FP64 instruction removed is worth two integer ones.  The tables were doubled / quadrupled
// (14 KB per block, built by fm_tables_init at kernel start with libdevice) so that each polynomial
// loses one or two degrees, and every non-immediate coefficient sits in __constant__ memory:
//   exp    : 2^(n/256)              256 doubles    9 FP64 instructions
//   -2 log : (-2/c_j, -2 log c_j)   256 pairs     10  (the Box-Muller radius wants -2 log u)
//   sincos : (sin, cos)(2 pi n/512) 512 pairs     15 for the fused Box-Muller rotation
//   sqrt   : MUFU seed + second-order correction   5
// Every kernel that calls these must call fm_tables_init() first.
// =====================================================================================
struct FmShared {
  double exp2t[256];
  double2 logt[256];
  double2 sct[512];
};
static __shared__ FmShared s_fm;

static __constant__ double FM_C[32] = {
    // exp: 1/24, 1/6, 256/ln2, ln2/256 hi (21 trailing zero bits), ln2/256 lo   (index 0..4), pad
    4.1666666666666664e-02, 1.6666666666666666e-01, 369.3299304675746, 0.00270760617331689, 7.453964567463233e-13,
    0.0, 0.0, 0.0,
    // -2 log1p(-g/2) = g + g^2 (1/4 + g/12 + g^2/32 + g^3/80 + g^4/192): 1/192, 1/80, 1/32, 1/12   (8..11);
    // -2 ln2 (12); 2^52 + 1023 (13); 1 - 2^-53 (14); 1.5 2^52 - 512 (15)
    0.005208333333333333, 0.0125, 0.03125, 0.08333333333333333, -1.3862943611198906, 4503599627371519.0,
    0.99999999999999988898, 6755399441055232.0,
    // sin(2 pi x) = x (S1 + z (S3 + z S5)), cos(2 pi x) - 1 = z (C2 + z C4), z = x^2, |x| <= 2^-10 turns   (16..20)
    6.283185307179586, -41.34170224039976, 81.60524927607506, -19.739208802178716, 64.9393940226683,
    // exp(x), |x| <= 2^-6: 1/720, 1/120 (21, 22); 1e-12 (23)
    1.3888888888888889e-03, 8.3333333333333332e-03, 1e-12,
    0.0, 0.0, 0.0, 0.0, 0.0, 0.0, 0.0, 0.0};

// Builds the shared tables (all threads of the block; ends with a barrier).
__device__ __forceinline__ void fm_tables_init() {
  for (int i = threadIdx.x; i < 256; i += blockDim.x) {
    s_fm.exp2t[i] = exp2((double)i * (1.0 / 256.0));
    // interval j of the mantissa m in [1,2): centre c_j, except the two intervals that touch
    // u = 1 (j = 0: c = 1, j = 255: c = 2) so that log(u) keeps its relative accuracy there.
    // Intervals with m >= 1 + 106/256 (~sqrt 2) are folded down by a factor 2 (exponent + 1).
    double c = 1.0 + ((double)i + 0.5) * (1.0 / 256.0);
    if (i == 0) c = 1.0;
    if (i == 255) c = 2.0;
    const double rc = 1.0 / c;
    const double L = (i == 0 || i == 255) ? 0.0 : -log(i >= 106 ? rc * 2.0 : rc);
    s_fm.logt[i] = make_double2(-2.0 * rc, -2.0 * L);
  }
  for (int i = threadIdx.x; i < 512; i += blockDim.x) {
    double sn, cs;
    sincospi((double)i * (1.0 / 256.0), &sn, &cs);
    s_fm.sct[i] = make_double2(sn, cs);
  }
  __syncthreads();
}

// ---- lock-step ("PP-wide") forms --------------------------------------------------------
// The FP64 pipe of this GPU only reaches its 2-cycle issue rate when ONE warp issues several
// independent DFMAs back to back (tools/probes/dfma_latency.cu: dependent chain 8.1 cycles;
// 1 chain/warp -> 3.0 cycles per instruction however many warps are resident, 2 chains -> 2.5,
// 4 chains -> 2.17).  So every function below evaluates PP independent arguments with the
// operations interleaved in source order; the scalar entry points are the PP = 1 instances.
#define MCRE_VP _Pragma("unroll") for (int p = 0; p < PP; ++p)

// exp(x) for |x| <= 700: n = round(256 x / ln2), exp(x) = 2^(n>>8) * T[n&255] * exp(r), |r| <= ln2/512
// (r^5/120 < 4e-17: the series stops at r^4).
template <int PP>
__device__ __forceinline__ void fm_exp_tv(const double (&x)[PP], double (&out)[PP]) {
  const double MAGIC = 6755399441055744.0;  // 1.5 * 2^52
  double t[PP], r[PP], T[PP], q[PP], tr[PP];
  int n[PP];
  MCRE_VP t[p] = fma(x[p], FM_C[2], MAGIC);
  MCRE_VP n[p] = __double2loint(t[p]);
  MCRE_VP t[p] -= MAGIC;
  MCRE_VP T[p] = s_fm.exp2t[n[p] & 255];
  MCRE_VP r[p] = fma(t[p], -FM_C[3], x[p]);
  MCRE_VP r[p] = fma(t[p], -FM_C[4], r[p]);
  MCRE_VP q[p] = fma(r[p], FM_C[0], FM_C[1]);
  MCRE_VP q[p] = fma(q[p], r[p], 0.5);
  MCRE_VP q[p] = fma(q[p], r[p], 1.0);          // 1 + r/2 + r^2/6 + r^3/24
  MCRE_VP tr[p] = T[p] * r[p];
  MCRE_VP q[p] = fma(tr[p], q[p], T[p]);        // T (1 + r q)
  // scale by 2^(n>>8) in the exponent field (result stays normal for |x| <= 700)
  MCRE_VP out[p] = __hiloint2double(__double2hiint(q[p]) + ((n[p] >> 8) << 20), __double2loint(q[p]));
}
__device__ __forceinline__ double fm_exp_t(double x) {
  const double a[1] = {x};
  double o[1];
  fm_exp_tv<1>(a, o);
  return o[0];
}

// exp(x) for |x| <= 2^-6 without table or range reduction (Taylor to x^6: 2^-42 / 5040 relative).
template <int PP>
__device__ __forceinline__ void fm_exp_smallv(const double (&x)[PP], double (&out)[PP]) {
  double q[PP];
  MCRE_VP q[p] = fma(x[p], FM_C[21], FM_C[22]);
  MCRE_VP q[p] = fma(q[p], x[p], FM_C[0]);
  MCRE_VP q[p] = fma(q[p], x[p], FM_C[1]);
  MCRE_VP q[p] = fma(q[p], x[p], 0.5);
  MCRE_VP q[p] = fma(q[p], x[p], 1.0);
  MCRE_VP out[p] = fma(q[p], x[p], 1.0);
}

__device__ __forceinline__ double fm_exp_small(double x) {
  const double a[1] = {x};
  double o[1];
  fm_exp_smallv<1>(a, o);
  return o[0];
}

// -2 log(u) for u in [2^-60, 2): u = 2^e m, m / c_j - 1 = -g/2 with the table holding -2/c_j, so
//   -2 log u = e (-2 ln2) - 2 log c_j + (g + g^2/4 + g^3/12 + g^4/32 + g^5/80 + g^6/192),  |g| <= 2^-8
// (2^-7 in the interval that touches 1 from above, u in [1, 1 + 2^-8): uniforms never get there; the general
// logarithm, ABOVE_ONE = true, carries the g^7/448 term for it).
template <int PP, bool ABOVE_ONE = false>
__device__ __forceinline__ void fm_neg2log_tv(const double (&u)[PP], double (&out)[PP]) {
  int hi[PP], E[PP];
  double m[PP], g[PP], q[PP], gg[PP], ed[PP], a[PP];
  double2 tb[PP];
  MCRE_VP hi[p] = __double2hiint(u[p]);
  MCRE_VP tb[p] = s_fm.logt[(hi[p] >> 12) & 255];
  MCRE_VP E[p] = (hi[p] + 0x96000) >> 20;              // biased exponent, +1 when m >= 1 + 106/256
  MCRE_VP m[p] = __hiloint2double((hi[p] & 0x000fffff) | 0x3ff00000, __double2loint(u[p]));
  // (double)(E - 1023) through the 2^52 bias: one FP64 add
  MCRE_VP ed[p] = __hiloint2double(0x43300000, E[p]) - FM_C[13];
  MCRE_VP g[p] = fma(m[p], tb[p].x, 2.0);              // -2 (m / c_j - 1)
  if (ABOVE_ONE) {
    MCRE_VP q[p] = fma(g[p], 0.002232142857142857, FM_C[8]);
    MCRE_VP q[p] = fma(q[p], g[p], FM_C[9]);
  } else {
    MCRE_VP q[p] = fma(g[p], FM_C[8], FM_C[9]);
  }
  MCRE_VP gg[p] = g[p] * g[p];
  MCRE_VP q[p] = fma(q[p], g[p], FM_C[10]);
  // e (-2 ln2) - 2 log c_j: |e| <= 60, so the rounding of ln2 costs at most 60 * 2^-53 absolute -
  // below half an ulp of the result whenever e != 0; for e = 0 the term vanishes
  MCRE_VP a[p] = fma(ed[p], FM_C[12], tb[p].y);
  MCRE_VP q[p] = fma(q[p], g[p], FM_C[11]);
  MCRE_VP q[p] = fma(q[p], g[p], 0.25);
  MCRE_VP q[p] = fma(gg[p], q[p], g[p]);
  MCRE_VP out[p] = a[p] + q[p];
}
// log(u) for u in [2^-60, 2)
template <int PP>
__device__ __forceinline__ void fm_log_tv(const double (&u)[PP], double (&out)[PP]) {
  fm_neg2log_tv<PP, true>(u, out);
  MCRE_VP out[p] *= -0.5;
}
__device__ __forceinline__ double fm_log_t(double u) {
  const double a[1] = {u};
  double o[1];
  fm_log_tv<1>(a, o);
  return o[0];
}

// sqrt(x) for x > 0 (normal range), PP-wide.  y = MUFU.RSQ64H(x) reads the high word only: relative error
// e <= 2^-20; with g = x y = sqrt(x) (1 + e) and r = 1 - g y = -2e - e^2:  sqrt(x) = g (1 - r)^(-1/2)
// = g + g r (1/2 + 3/8 r) + O(r^3) (2^-60 relative).  The rounding of g is absorbed by r.
template <int PP>
__device__ __forceinline__ void fm_sqrt_posv(const double (&x)[PP], double (&out)[PP]) {
  double y[PP], g[PP], r[PP], q[PP], t[PP];
  MCRE_VP y[p] = fm_rsqrt_approx(x[p]);
  MCRE_VP g[p] = x[p] * y[p];
  MCRE_VP r[p] = fma(-g[p], y[p], 1.0);
  MCRE_VP q[p] = fma(r[p], 0.375, 0.5);
  MCRE_VP t[p] = g[p] * r[p];
  MCRE_VP out[p] = fma(t[p], q[p], g[p]);
}

// (sin(2 pi u), cos(2 pi u)) for u in [0, 1): n = round(512 u), angle = 2 pi (n/512 + x), |x| <= 2^-10.
template <int PP>
__device__ __forceinline__ void fm_sincos2pi_tv(const double (&u)[PP], double (&sn)[PP], double (&cs)[PP]) {
  const double MAGIC = 6755399441055744.0;
  double t[PP], x[PP], z[PP], ps[PP], pc[PP], a[PP], b[PP];
  double2 tb[PP];
  MCRE_VP t[p] = fma(u[p], 512.0, MAGIC);
  MCRE_VP tb[p] = s_fm.sct[__double2loint(t[p]) & 511];
  MCRE_VP t[p] -= MAGIC;
  MCRE_VP x[p] = fma(t[p], -0.001953125, u[p]);        // exact, [-2^-10, 2^-10]
  MCRE_VP z[p] = x[p] * x[p];
  MCRE_VP ps[p] = fma(z[p], FM_C[18], FM_C[17]);
  MCRE_VP pc[p] = fma(z[p], FM_C[20], FM_C[19]);
  MCRE_VP ps[p] = fma(ps[p], z[p], FM_C[16]);
  MCRE_VP pc[p] = z[p] * pc[p];                        // cos - 1
  MCRE_VP ps[p] = x[p] * ps[p];                        // sin
  // rotate: sin(a + x) = S + (C sx + S cm),  cos(a + x) = C + (C cm - S sx)
  MCRE_VP a[p] = tb[p].x * pc[p];
  MCRE_VP b[p] = tb[p].y * pc[p];
  MCRE_VP a[p] = fma(tb[p].y, ps[p], a[p]);
  MCRE_VP b[p] = fma(-tb[p].x, ps[p], b[p]);
  MCRE_VP sn[p] = tb[p].x + a[p];
  MCRE_VP cs[p] = tb[p].y + b[p];
}
// Box-Muller tail: (rad cos(2 pi v), rad sin(2 pi v)) with v = d - 1 given as the mantissa double
// d in [1, 2) (no subtraction needed: the reduction works on d directly), the rotation against the
// table entry fused with the scaling by rad:  rad cos(a + x) = RC + (RC cm - RS sx), RC = rad C ...
template <int PP>
__device__ __forceinline__ void fm_polar_tv(const double (&d)[PP], const double (&rad)[PP], double (&zc)[PP],
                                            double (&zs)[PP]) {
  double t[PP], x[PP], z[PP], ps[PP], pc[PP], rc[PP], rs[PP];
  double2 tb[PP];
  MCRE_VP t[p] = fma(d[p], 512.0, FM_C[15]);            // 1.5 2^52 + round(512 v)
  MCRE_VP tb[p] = s_fm.sct[__double2loint(t[p]) & 511];
  MCRE_VP t[p] -= FM_C[15];                             // round(512 v) + 512
  MCRE_VP x[p] = fma(t[p], -0.001953125, d[p]);        // v - n/512, exact, [-2^-10, 2^-10]
  MCRE_VP rc[p] = rad[p] * tb[p].y;
  MCRE_VP rs[p] = rad[p] * tb[p].x;
  MCRE_VP z[p] = x[p] * x[p];
  MCRE_VP ps[p] = fma(z[p], FM_C[18], FM_C[17]);
  MCRE_VP pc[p] = fma(z[p], FM_C[20], FM_C[19]);
  MCRE_VP ps[p] = fma(ps[p], z[p], FM_C[16]);
  MCRE_VP pc[p] = z[p] * pc[p];                        // cos - 1
  MCRE_VP ps[p] = x[p] * ps[p];                        // sin
  MCRE_VP zc[p] = fma(rc[p], pc[p], rc[p]);
  MCRE_VP zs[p] = fma(rs[p], pc[p], rs[p]);
  MCRE_VP zc[p] = fma(-rs[p], ps[p], zc[p]);
  MCRE_VP zs[p] = fma(rc[p], ps[p], zs[p]);
}

__device__ __forceinline__ void fm_sincos2pi_t(double u, double &sn, double &cs) {
  const double a[1] = {u};
  double s[1], c[1];
  fm_sincos2pi_tv<1>(a, s, c);
  sn = s[0]; cs = c[0];
}
#endif

}  // namespace mcre
