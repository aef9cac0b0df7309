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

// sqrt(x) for x >= 0 (0 -> 0), normal range.
__device__ __forceinline__ double fm_sqrt(double x) {
  const double xs = x > 0.0 ? x : 1.0;
  double y = fm_rsqrt_approx(xs);
  double g = xs * y, h = 0.5 * y;
  double r = fma(-h, g, 0.5);
  g = fma(g, r, g); h = fma(h, r, h);
  r = fma(-h, g, 0.5);
  g = fma(g, r, g); h = fma(h, r, h);
  g = fma(fma(-g, g, xs), h, g);
  return x > 0.0 ? g : 0.0;
}

// sqrt(x) for x > 0 (normal range): no zero guard.  MUFU seed (2^-22) + two coupled
// Goldschmidt iterations; the last step is the residual correction g + (x - g^2) h with the
// first-iteration h (relative error 2^-44: enough for a correction of relative size 2^-44).
__device__ __forceinline__ double fm_sqrt_pos(double x) {
  const double y = fm_rsqrt_approx(x);
  double g = x * y, h = 0.5 * y;
  const double r = fma(-h, g, 0.5);
  g = fma(g, r, g); h = fma(h, r, h);
  return fma(fma(-g, g, x), h, g);
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
// Table-driven variants (second ncu pass, profiles/r01_irc_main_v1_*): the v1 kernel spent
// ~100 of its 148 FP64 instructions per path-step in the long Taylor polynomials above and
// 77 instructions materialising their coefficients.  Here the argument is reduced against
// small tables in SHARED memory (4.6 KB per block, built by fm_tables_init at kernel start
// with libdevice), so the polynomials shrink to degree 5-7, and the coefficients sit in
// __constant__ memory (one LDCU.128 per two coefficients instead of four UMOV).
//   exp   : 2^(n/64)            64 doubles     19 -> 11 FP64 instructions
//   log   : (1/c_j, log c_j)    128 pairs      25 -> 13
//   sincos: (sin, cos)(2 pi n/128) 128 pairs   24 -> 14
// Every kernel that calls these must call fm_tables_init() first.
// =====================================================================================
struct FmShared {
  double exp2t[64];
  double2 logt[128];
  double2 sct[128];
};
static __shared__ FmShared s_fm;

static __constant__ double FM_C[24] = {
    // exp: 1/120, 1/24, 1/6, 1/2   (index 0..3)
    8.3333333333333332e-03, 4.1666666666666664e-02, 1.6666666666666666e-01, 0.5,
    // exp reduction: 64/ln2, ln2/64 hi, ln2/64 lo, pad   (4..7)
    92.33248261689366, 0.01083042469326756, 2.9815858271643302e-12, 0.0,
    // log1p(f)/f - 1: -1/2, 1/3, -1/4, 1/5, -1/6, 1/7, -1/8, pad   (8..15)
    -0.5, 3.3333333333333331e-01, -0.25, 0.2, -1.6666666666666666e-01, 1.4285714285714285e-01, -0.125, 0.0,
    // sin: -1/6, 1/120, -1/5040 ; cos: -1/2, 1/24, -1/720 ; 2 pi ; pad   (16..23)
    -1.6666666666666666e-01, 8.3333333333333332e-03, -1.984126984126984e-04,
    -0.5, 4.1666666666666664e-02, -1.3888888888888889e-03, 6.283185307179586, 0.0};

// Builds the shared tables (all threads of the block; ends with a barrier).
__device__ __forceinline__ void fm_tables_init() {
  for (int i = threadIdx.x; i < 64; i += blockDim.x) s_fm.exp2t[i] = exp2((double)i * (1.0 / 64.0));
  for (int i = threadIdx.x; i < 128; i += blockDim.x) {
    // interval j of the mantissa m in [1,2): centre c_j, except the two intervals that touch
    // u = 1 (j = 0: c = 1, j = 127: c = 2) so that log(u) keeps its relative accuracy there.
    // Intervals with m >= 1 + 53/128 (~sqrt 2) are folded down by a factor 2 (exponent + 1).
    double c = 1.0 + ((double)i + 0.5) * (1.0 / 128.0);
    if (i == 0) c = 1.0;
    if (i == 127) c = 2.0;
    const double rc = 1.0 / c;
    double L = (i == 0 || i == 127) ? 0.0 : -log(i >= 53 ? rc * 2.0 : rc);
    s_fm.logt[i] = make_double2(rc, L);
    double sn, cs;
    sincospi((double)i * (1.0 / 64.0), &sn, &cs);
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

// exp(x) for |x| <= 700: n = round(64 x / ln2), exp(x) = 2^(n>>6) * T[n&63] * exp(r), |r| <= ln2/128.
template <int PP>
__device__ __forceinline__ void fm_exp_tv(const double (&x)[PP], double (&out)[PP]) {
  const double MAGIC = 6755399441055744.0;  // 1.5 * 2^52
  double t[PP], r[PP], T[PP], q[PP], tr[PP];
  int n[PP];
  MCRE_VP t[p] = fma(x[p], FM_C[4], MAGIC);
  MCRE_VP n[p] = __double2loint(t[p]);
  MCRE_VP t[p] -= MAGIC;
  MCRE_VP T[p] = s_fm.exp2t[n[p] & 63];
  MCRE_VP r[p] = fma(t[p], -FM_C[5], x[p]);
  MCRE_VP r[p] = fma(t[p], -FM_C[6], r[p]);
  MCRE_VP q[p] = fma(r[p], FM_C[0], FM_C[1]);
  MCRE_VP q[p] = fma(q[p], r[p], FM_C[2]);
  MCRE_VP q[p] = fma(q[p], r[p], FM_C[3]);
  MCRE_VP q[p] = fma(q[p], r[p], 1.0);          // 1 + r/2 + r^2/6 + r^3/24 + r^4/120
  MCRE_VP tr[p] = T[p] * r[p];
  MCRE_VP q[p] = fma(tr[p], q[p], T[p]);        // T (1 + r q)
  // scale by 2^(n>>6) in the exponent field (result stays normal for |x| <= 700)
  MCRE_VP out[p] = __hiloint2double(__double2hiint(q[p]) + ((n[p] >> 6) << 20), __double2loint(q[p]));
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
  MCRE_VP q[p] = fma(x[p], 1.3888888888888889e-03, 8.3333333333333332e-03);
  MCRE_VP q[p] = fma(q[p], x[p], 4.1666666666666664e-02);
  MCRE_VP q[p] = fma(q[p], x[p], 1.6666666666666666e-01);
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

// log(u) for u in [2^-60, 2).
template <int PP>
__device__ __forceinline__ void fm_log_tv(const double (&u)[PP], double (&out)[PP]) {
  const double LN2 = 6.931471805599453094e-01;
  int hi[PP], e[PP];
  double m[PP], f[PP], q[PP], ff[PP], ed[PP], a[PP];
  double2 tb[PP];
  MCRE_VP hi[p] = __double2hiint(u[p]);
  MCRE_VP tb[p] = s_fm.logt[(hi[p] >> 13) & 127];
  MCRE_VP e[p] = ((hi[p] + 0x96000) >> 20) - 1023;     // exponent, +1 when m >= 1 + 53/128
  MCRE_VP m[p] = __hiloint2double((hi[p] & 0x000fffff) | 0x3ff00000, __double2loint(u[p]));
  // (double)e through the 2^52 + 2^31 bias: one integer op + one FP64 add
  MCRE_VP ed[p] = __hiloint2double(0x43300000, e[p] ^ 0x80000000) - 4503601774854144.0;
  MCRE_VP f[p] = fma(m[p], tb[p].x, -1.0);             // m / c_j - 1, |f| <= 1/128
  MCRE_VP q[p] = fma(f[p], FM_C[14], FM_C[13]);
  MCRE_VP ff[p] = f[p] * f[p];
  MCRE_VP q[p] = fma(q[p], f[p], FM_C[12]);
  // e ln2 + log c_j: |e| <= 60, so the rounding of ln2 costs at most 60 * 2^-54 absolute -
  // below half an ulp of the result whenever e != 0; for e = 0 the term vanishes
  MCRE_VP a[p] = fma(ed[p], LN2, tb[p].y);
  MCRE_VP q[p] = fma(q[p], f[p], FM_C[11]);
  MCRE_VP q[p] = fma(q[p], f[p], FM_C[10]);
  MCRE_VP q[p] = fma(q[p], f[p], FM_C[9]);
  MCRE_VP q[p] = fma(q[p], f[p], FM_C[8]);             // -1/2 + f/3 - f^2/4 ...
  MCRE_VP q[p] = fma(ff[p], q[p], f[p]);               // log1p(f)
  MCRE_VP out[p] = a[p] + q[p];
}
__device__ __forceinline__ double fm_log_t(double u) {
  const double a[1] = {u};
  double o[1];
  fm_log_tv<1>(a, o);
  return o[0];
}

// sqrt(x) for x > 0, PP-wide (same arithmetic as fm_sqrt_pos).
template <int PP>
__device__ __forceinline__ void fm_sqrt_posv(const double (&x)[PP], double (&out)[PP]) {
  double y[PP], g[PP], h[PP], r[PP];
  MCRE_VP y[p] = fm_rsqrt_approx(x[p]);
  MCRE_VP g[p] = x[p] * y[p];
  MCRE_VP h[p] = 0.5 * y[p];
  MCRE_VP r[p] = fma(-h[p], g[p], 0.5);
  MCRE_VP g[p] = fma(g[p], r[p], g[p]);
  MCRE_VP h[p] = fma(h[p], r[p], h[p]);
  MCRE_VP r[p] = fma(-g[p], g[p], x[p]);
  MCRE_VP out[p] = fma(r[p], h[p], g[p]);
}

// (sin(2 pi u), cos(2 pi u)) for u in [0, 1): n = round(128 u), angle = 2 pi n/128 + x, |x| <= pi/128.
template <int PP>
__device__ __forceinline__ void fm_sincos2pi_tv(const double (&u)[PP], double (&sn)[PP], double (&cs)[PP]) {
  const double MAGIC = 6755399441055744.0;
  double t[PP], x[PP], z[PP], ps[PP], pc[PP], xz[PP], a[PP], b[PP];
  double2 tb[PP];
  MCRE_VP t[p] = fma(u[p], 128.0, MAGIC);
  MCRE_VP tb[p] = s_fm.sct[__double2loint(t[p]) & 127];
  MCRE_VP t[p] -= MAGIC;
  MCRE_VP x[p] = fma(t[p], -0.0078125, u[p]);          // exact, [-1/256, 1/256]
  MCRE_VP x[p] = x[p] * FM_C[22];
  MCRE_VP z[p] = x[p] * x[p];
  MCRE_VP ps[p] = fma(z[p], FM_C[18], FM_C[17]);
  MCRE_VP pc[p] = fma(z[p], FM_C[21], FM_C[20]);
  MCRE_VP xz[p] = x[p] * z[p];
  MCRE_VP ps[p] = fma(ps[p], z[p], FM_C[16]);
  MCRE_VP pc[p] = fma(pc[p], z[p], FM_C[19]);
  MCRE_VP ps[p] = fma(xz[p], ps[p], x[p]);             // sin x
  MCRE_VP pc[p] = z[p] * pc[p];                        // cos x - 1
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
  const double MAGIC = 6755399441055744.0;
  double t[PP], x[PP], z[PP], ps[PP], pc[PP], xz[PP], rc[PP], rs[PP];
  double2 tb[PP];
  MCRE_VP t[p] = fma(d[p], 128.0, MAGIC - 128.0);       // MAGIC + round(128 v)
  MCRE_VP tb[p] = s_fm.sct[__double2loint(t[p]) & 127];
  MCRE_VP t[p] -= MAGIC - 128.0;                        // round(128 v) + 128
  MCRE_VP x[p] = fma(t[p], -0.0078125, d[p]);          // v - n/128, exact, [-1/256, 1/256]
  MCRE_VP x[p] = x[p] * FM_C[22];
  MCRE_VP rc[p] = rad[p] * tb[p].y;
  MCRE_VP rs[p] = rad[p] * tb[p].x;
  MCRE_VP z[p] = x[p] * x[p];
  MCRE_VP ps[p] = fma(z[p], FM_C[18], FM_C[17]);
  MCRE_VP pc[p] = fma(z[p], FM_C[21], FM_C[20]);
  MCRE_VP xz[p] = x[p] * z[p];
  MCRE_VP ps[p] = fma(ps[p], z[p], FM_C[16]);
  MCRE_VP pc[p] = fma(pc[p], z[p], FM_C[19]);
  MCRE_VP ps[p] = fma(xz[p], ps[p], x[p]);             // sin x
  MCRE_VP pc[p] = z[p] * pc[p];                        // cos x - 1
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
