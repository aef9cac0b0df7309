// Forward-mode dual numbers for pathwise sensitivities in registers.
//
// The kernels are written once against a scalar type R: R = double (values only) or
// R = Dual<NT> (value + NT tangents w.r.t. the model parameters).  This is the
// hand-written replacement of torch.autograd on the hot path
// (src/controller/controller.py:609-627): metric profiles have many more outputs
// (one per exposure date) than the model has parameters, so tangents are propagated
// forward with the path instead of taping it.
#pragma once
#if defined(MCRE_FAST_MATH) && MCRE_FAST_MATH >= 2
#include "fastmath.cuh"
#define MCRE_EXP(x) fm_exp_t(x)
#define MCRE_LOG(x) fm_log_t(x)
#define MCRE_SQRT(x) fm_sqrt(x)
#elif defined(MCRE_FAST_MATH)
#include "fastmath.cuh"
#define MCRE_EXP(x) fm_exp(x)
#define MCRE_LOG(x) fm_log(x)
#define MCRE_SQRT(x) fm_sqrt(x)
#else
#define MCRE_EXP(x) exp(x)
#define MCRE_LOG(x) log(x)
#define MCRE_SQRT(x) sqrt(x)
#endif

namespace mcre {

template <int N>
struct Dual {
  double v;
  double d[N];
};

template <int N> struct RealOf { typedef Dual<N> type; };
template <> struct RealOf<0> { typedef double type; };

// ---- plain double overloads -----------------------------------------------------
__device__ __forceinline__ double val(double x) { return x; }
__device__ __forceinline__ double tan_of(double, int) { return 0.0; }
__device__ __forceinline__ double r_exp(double x) { return MCRE_EXP(x); }
__device__ __forceinline__ double r_log(double x) { return MCRE_LOG(x); }
__device__ __forceinline__ double r_sqrt(double x) { return MCRE_SQRT(x); }
#if defined(MCRE_FAST_MATH) && MCRE_FAST_MATH >= 2
__device__ __forceinline__ double r_sqrt_pos(double x) { return fm_sqrt_pos(x); }
__device__ __forceinline__ double r_exp_small(double x) { return fm_exp_small(x); }
#else
__device__ __forceinline__ double r_sqrt_pos(double x) { return MCRE_SQRT(x); }
__device__ __forceinline__ double r_exp_small(double x) { return MCRE_EXP(x); }
#endif
#ifdef MCRE_FAST_MATH
#define MCRE_RCP(x) fm_div(1.0, (x))
__device__ __forceinline__ double r_div(double a, double b) { return fm_div(a, b); }
#else
#define MCRE_RCP(x) (1.0 / (x))
__device__ __forceinline__ double r_div(double a, double b) { return a / b; }
#endif
__device__ __forceinline__ double r_relu(double x) { return fmax(x, 0.0); }
__device__ __forceinline__ double r_max(double x, double c) { return fmax(x, c); }
__device__ __forceinline__ double r_mask(double x, bool keep) { return keep ? x : 0.0; }
__device__ __forceinline__ double r_const(double c, const double *) { return c; }
template <typename R> __device__ __forceinline__ R r_zero();
template <> __device__ __forceinline__ double r_zero<double>() { return 0.0; }
// load a (1+NT)-strided constant
template <typename R> __device__ __forceinline__ R r_load(const double *p, int idx);
template <> __device__ __forceinline__ double r_load<double>(const double *p, int idx) { return __ldg(p + idx); }

// ---- Dual<N> ----------------------------------------------------------------------
#define MCRE_DUAL_LOOP for (int i = 0; i < N; ++i)
template <int N> __device__ __forceinline__ double val(const Dual<N> &x) { return x.v; }
// makes x the independent variable of tangent slot `slot` (slot < 0: a constant)
__device__ __forceinline__ void r_seed(double &, int) {}
template <int N> __device__ __forceinline__ void r_seed(Dual<N> &x, int slot) {
#pragma unroll
  for (int i = 0; i < N; ++i) x.d[i] = i == slot ? 1.0 : 0.0;
}
template <int N> __device__ __forceinline__ double tan_of(const Dual<N> &x, int i) { return x.d[i]; }

template <int N> __device__ __forceinline__ Dual<N> operator+(const Dual<N> &a, const Dual<N> &b) {
  Dual<N> r; r.v = a.v + b.v;
#pragma unroll
  MCRE_DUAL_LOOP r.d[i] = a.d[i] + b.d[i];
  return r;
}
template <int N> __device__ __forceinline__ Dual<N> operator-(const Dual<N> &a, const Dual<N> &b) {
  Dual<N> r; r.v = a.v - b.v;
#pragma unroll
  MCRE_DUAL_LOOP r.d[i] = a.d[i] - b.d[i];
  return r;
}
template <int N> __device__ __forceinline__ Dual<N> operator-(const Dual<N> &a) {
  Dual<N> r; r.v = -a.v;
#pragma unroll
  MCRE_DUAL_LOOP r.d[i] = -a.d[i];
  return r;
}
template <int N> __device__ __forceinline__ Dual<N> operator*(const Dual<N> &a, const Dual<N> &b) {
  Dual<N> r; r.v = a.v * b.v;
#pragma unroll
  MCRE_DUAL_LOOP r.d[i] = a.d[i] * b.v + a.v * b.d[i];
  return r;
}
template <int N> __device__ __forceinline__ Dual<N> operator/(const Dual<N> &a, const Dual<N> &b) {
  Dual<N> r; double inv = MCRE_RCP(b.v); r.v = a.v * inv;
#pragma unroll
  MCRE_DUAL_LOOP r.d[i] = (a.d[i] - r.v * b.d[i]) * inv;
  return r;
}
template <int N> __device__ __forceinline__ Dual<N> r_div(const Dual<N> &a, const Dual<N> &b) { return a / b; }
template <int N> __device__ __forceinline__ Dual<N> r_div(const Dual<N> &a, double b) { return a * MCRE_RCP(b); }
template <int N> __device__ __forceinline__ Dual<N> operator+(const Dual<N> &a, double b) { Dual<N> r = a; r.v += b; return r; }
template <int N> __device__ __forceinline__ Dual<N> operator+(double b, const Dual<N> &a) { Dual<N> r = a; r.v += b; return r; }
template <int N> __device__ __forceinline__ Dual<N> operator-(const Dual<N> &a, double b) { Dual<N> r = a; r.v -= b; return r; }
template <int N> __device__ __forceinline__ Dual<N> operator-(double b, const Dual<N> &a) { Dual<N> r = -a; r.v += b; return r; }
template <int N> __device__ __forceinline__ Dual<N> operator*(const Dual<N> &a, double b) {
  Dual<N> r; r.v = a.v * b;
#pragma unroll
  MCRE_DUAL_LOOP r.d[i] = a.d[i] * b;
  return r;
}
template <int N> __device__ __forceinline__ Dual<N> operator*(double b, const Dual<N> &a) { return a * b; }
template <int N> __device__ __forceinline__ Dual<N> operator/(const Dual<N> &a, double b) { return a * (1.0 / b); }
template <int N> __device__ __forceinline__ Dual<N> operator/(double a, const Dual<N> &b) {
  Dual<N> r; double inv = MCRE_RCP(b.v); r.v = a * inv; double s = -r.v * inv;
#pragma unroll
  MCRE_DUAL_LOOP r.d[i] = s * b.d[i];
  return r;
}
template <int N> __device__ __forceinline__ Dual<N> r_div(double a, const Dual<N> &b) { return a / b; }
template <int N> __device__ __forceinline__ Dual<N> &operator+=(Dual<N> &a, const Dual<N> &b) { a = a + b; return a; }
template <int N> __device__ __forceinline__ Dual<N> &operator+=(Dual<N> &a, double b) { a.v += b; return a; }

template <int N> __device__ __forceinline__ Dual<N> r_exp(const Dual<N> &x) {
  Dual<N> r; r.v = MCRE_EXP(x.v);
#pragma unroll
  MCRE_DUAL_LOOP r.d[i] = r.v * x.d[i];
  return r;
}
template <int N> __device__ __forceinline__ Dual<N> r_log(const Dual<N> &x) {
  Dual<N> r; r.v = MCRE_LOG(x.v); double inv = MCRE_RCP(x.v);
#pragma unroll
  MCRE_DUAL_LOOP r.d[i] = inv * x.d[i];
  return r;
}
// sqrt with torch's convention d sqrt(0) = inf * 0 -> we use 0 at exactly 0 (the clamp in front kills it)
template <int N> __device__ __forceinline__ Dual<N> r_sqrt(const Dual<N> &x) {
  Dual<N> r; r.v = MCRE_SQRT(x.v); double s = r.v > 0.0 ? 0.5 / r.v : 0.0;
#pragma unroll
  MCRE_DUAL_LOOP r.d[i] = s * x.d[i];
  return r;
}
template <int N> __device__ __forceinline__ Dual<N> r_sqrt_pos(const Dual<N> &x) {
  Dual<N> r; r.v = r_sqrt_pos(x.v); double s = 0.5 / r.v;
#pragma unroll
  MCRE_DUAL_LOOP r.d[i] = s * x.d[i];
  return r;
}
template <int N> __device__ __forceinline__ Dual<N> r_exp_small(const Dual<N> &x) {
  Dual<N> r; r.v = r_exp_small(x.v);
#pragma unroll
  MCRE_DUAL_LOOP r.d[i] = r.v * x.d[i];
  return r;
}
// relu / clamp-from-below: subgradient 1 where x > bound else 0 (torch.clamp / relu backward)
template <int N> __device__ __forceinline__ Dual<N> r_relu(const Dual<N> &x) {
  Dual<N> r; bool on = x.v > 0.0; r.v = on ? x.v : 0.0;
#pragma unroll
  MCRE_DUAL_LOOP r.d[i] = on ? x.d[i] : 0.0;
  return r;
}
template <int N> __device__ __forceinline__ Dual<N> r_max(const Dual<N> &x, double c) {
  Dual<N> r; bool on = x.v >= c; r.v = on ? x.v : c;
#pragma unroll
  MCRE_DUAL_LOOP r.d[i] = on ? x.d[i] : 0.0;
  return r;
}
template <int N> __device__ __forceinline__ Dual<N> r_mask(const Dual<N> &x, bool keep) {
  Dual<N> r; r.v = keep ? x.v : 0.0;
#pragma unroll
  MCRE_DUAL_LOOP r.d[i] = keep ? x.d[i] : 0.0;
  return r;
}
template <int N> __device__ __forceinline__ Dual<N> dual_zero() {
  Dual<N> r; r.v = 0.0;
#pragma unroll
  MCRE_DUAL_LOOP r.d[i] = 0.0;
  return r;
}
template <int N> __device__ __forceinline__ Dual<N> dual_load(const double *p, int idx) {
  Dual<N> r; const double *q = p + (size_t)idx * (N + 1); r.v = __ldg(q);
#pragma unroll
  MCRE_DUAL_LOOP r.d[i] = __ldg(q + 1 + i);
  return r;
}
#undef MCRE_DUAL_LOOP

// generic front-ends selected on the scalar type
template <typename R> struct RealTraits;
template <> struct RealTraits<double> {
  static const int NT = 0;
  __device__ static __forceinline__ double zero() { return 0.0; }
  __device__ static __forceinline__ double load(const double *p, int idx) { return __ldg(p + idx); }
  __device__ static __forceinline__ double lift(double c) { return c; }
};
template <int N> struct RealTraits<Dual<N> > {
  static const int NT = N;
  __device__ static __forceinline__ Dual<N> zero() { return dual_zero<N>(); }
  __device__ static __forceinline__ Dual<N> load(const double *p, int idx) { return dual_load<N>(p, idx); }
  __device__ static __forceinline__ Dual<N> lift(double c) { Dual<N> r = dual_zero<N>(); r.v = c; return r; }
};

// x with its value replaced (tangents kept): used where lanes share a group-reduced value
// but each keeps the tangents of its own contribution.
__device__ __forceinline__ double r_with_value(double, double v) { return v; }
template <int N> __device__ __forceinline__ Dual<N> r_with_value(const Dual<N> &x, double v) { Dual<N> r = x; r.v = v; return r; }
// independent variable #k of the tangent space
template <typename R> struct RealVar;
template <> struct RealVar<double> { __device__ static __forceinline__ double make(double v, int) { return v; } };
template <int N> struct RealVar<Dual<N> > {
  __device__ static __forceinline__ Dual<N> make(double v, int k) {
    Dual<N> r; r.v = v;
#pragma unroll
    for (int i = 0; i < N; ++i) r.d[i] = (i == k) ? 1.0 : 0.0;
    return r;
  }
};

// Symmetric linear fuzzy indicator (src/maths/maths.py:3-9): hard 1{x>0} or
// clamp((x+eps)/(2 eps), 0, 1) with slope 1/(2 eps) inside the band.
__device__ __forceinline__ double r_fuzzy(double x, bool fuzzy, double eps) {
  if (!fuzzy) return x > 0.0 ? 1.0 : 0.0;
  return fmin(fmax((x + eps) / (2.0 * eps), 0.0), 1.0);
}
template <int N> __device__ __forceinline__ Dual<N> r_fuzzy(const Dual<N> &x, bool fuzzy, double eps) {
  Dual<N> r;
  if (!fuzzy) { r = dual_zero<N>(); r.v = x.v > 0.0 ? 1.0 : 0.0; return r; }
  double t = (x.v + eps) / (2.0 * eps);
  bool in = t >= 0.0 && t <= 1.0;
  r.v = fmin(fmax(t, 0.0), 1.0);
  double s = in ? 1.0 / (2.0 * eps) : 0.0;
#pragma unroll
  for (int i = 0; i < N; ++i) r.d[i] = s * x.d[i];
  return r;
}

}  // namespace mcre
