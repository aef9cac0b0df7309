// Second-order forward-mode numbers: value, N first and N (N + 1) / 2 second derivatives with respect to the lane-local
// parameters, propagated with the path.  The hand-written replacement of the reference's double backward
// (src/controller/controller.py:631-648: torch.autograd.grad of every first derivative with create_graph=True): a
// kernel written against the scalar type R yields pathwise Hessians when R = Dual2<N>.
//
// The piecewise-linear functions (relu, clamp, the fuzzy indicator, masks) have zero curvature, exactly like their torch
// double backward, so a Hessian estimated this way carries the products' smooth curvature only (the reference's
// convention; a kink contributes nothing).
//
// The kernels see the second derivatives as further "tangent" components: RealTraits<Dual2<N>>::NT = N + N (N + 1) / 2,
// tan_of(x, k) = d[k] for k < N and the packed upper triangle h[(i, j), i <= j] behind it (row-major: 00 01 02 11 12 22
// for N = 3).
#pragma once
#include "dual.cuh"

namespace mcre {

template <int N>
struct Dual2 {
  static const int M = N * (N + 1) / 2;
  double v;
  double d[N];
  double h[N * (N + 1) / 2];
};

#define MCRE_D2_PAIRS(body)                         \
  {                                                 \
    int ij = 0;                                     \
    _Pragma("unroll") for (int i = 0; i < N; ++i)   \
    _Pragma("unroll") for (int j = i; j < N; ++j) { \
      body;                                         \
      ++ij;                                         \
    }                                               \
  }

template <> struct RealOf<9> { typedef Dual2<3> type; };

template <int N> __device__ __forceinline__ double val(const Dual2<N> &x) { return x.v; }
template <int N> __device__ __forceinline__ double tan_of(const Dual2<N> &x, int k) { return k < N ? x.d[k] : x.h[k - N]; }
template <int N> __device__ __forceinline__ Dual2<N> dual2_zero() {
  Dual2<N> r; r.v = 0.0;
#pragma unroll
  for (int i = 0; i < N; ++i) r.d[i] = 0.0;
#pragma unroll
  for (int i = 0; i < Dual2<N>::M; ++i) r.h[i] = 0.0;
  return r;
}
template <int N> __device__ __forceinline__ void r_seed(Dual2<N> &x, int slot) {
  const double v = x.v; x = dual2_zero<N>(); x.v = v;
#pragma unroll
  for (int i = 0; i < N; ++i) x.d[i] = i == slot ? 1.0 : 0.0;
}
// f(x) with f' = f1, f'' = f2 at x.v
template <int N> __device__ __forceinline__ Dual2<N> d2_chain(const Dual2<N> &x, double f, double f1, double f2) {
  Dual2<N> r; r.v = f;
#pragma unroll
  for (int i = 0; i < N; ++i) r.d[i] = f1 * x.d[i];
  MCRE_D2_PAIRS(r.h[ij] = f1 * x.h[ij] + f2 * (x.d[i] * x.d[j]))
  return r;
}
// x where `on`, else the constant c (zero derivatives)
template <int N> __device__ __forceinline__ Dual2<N> d2_select(const Dual2<N> &x, bool on, double c) {
  Dual2<N> r; r.v = on ? x.v : c;
#pragma unroll
  for (int i = 0; i < N; ++i) r.d[i] = on ? x.d[i] : 0.0;
#pragma unroll
  for (int i = 0; i < Dual2<N>::M; ++i) r.h[i] = on ? x.h[i] : 0.0;
  return r;
}

template <int N> __device__ __forceinline__ Dual2<N> operator+(const Dual2<N> &a, const Dual2<N> &b) {
  Dual2<N> r; r.v = a.v + b.v;
#pragma unroll
  for (int i = 0; i < N; ++i) r.d[i] = a.d[i] + b.d[i];
#pragma unroll
  for (int i = 0; i < Dual2<N>::M; ++i) r.h[i] = a.h[i] + b.h[i];
  return r;
}
template <int N> __device__ __forceinline__ Dual2<N> operator-(const Dual2<N> &a, const Dual2<N> &b) {
  Dual2<N> r; r.v = a.v - b.v;
#pragma unroll
  for (int i = 0; i < N; ++i) r.d[i] = a.d[i] - b.d[i];
#pragma unroll
  for (int i = 0; i < Dual2<N>::M; ++i) r.h[i] = a.h[i] - b.h[i];
  return r;
}
template <int N> __device__ __forceinline__ Dual2<N> operator-(const Dual2<N> &a) {
  Dual2<N> r; r.v = -a.v;
#pragma unroll
  for (int i = 0; i < N; ++i) r.d[i] = -a.d[i];
#pragma unroll
  for (int i = 0; i < Dual2<N>::M; ++i) r.h[i] = -a.h[i];
  return r;
}
template <int N> __device__ __forceinline__ Dual2<N> operator*(const Dual2<N> &a, const Dual2<N> &b) {
  Dual2<N> r; r.v = a.v * b.v;
#pragma unroll
  for (int i = 0; i < N; ++i) r.d[i] = a.d[i] * b.v + a.v * b.d[i];
  MCRE_D2_PAIRS(r.h[ij] = a.h[ij] * b.v + a.v * b.h[ij] + (a.d[i] * b.d[j] + a.d[j] * b.d[i]))
  return r;
}
template <int N> __device__ __forceinline__ Dual2<N> operator*(const Dual2<N> &a, double b) {
  Dual2<N> r; r.v = a.v * b;
#pragma unroll
  for (int i = 0; i < N; ++i) r.d[i] = a.d[i] * b;
#pragma unroll
  for (int i = 0; i < Dual2<N>::M; ++i) r.h[i] = a.h[i] * b;
  return r;
}
template <int N> __device__ __forceinline__ Dual2<N> operator*(double b, const Dual2<N> &a) { return a * b; }
template <int N> __device__ __forceinline__ Dual2<N> d2_recip(const Dual2<N> &b) {
  const double inv = MCRE_RCP(b.v), inv2 = inv * inv;
  return d2_chain(b, inv, -inv2, 2.0 * inv2 * inv);
}
template <int N> __device__ __forceinline__ Dual2<N> operator/(const Dual2<N> &a, const Dual2<N> &b) { return a * d2_recip(b); }
template <int N> __device__ __forceinline__ Dual2<N> operator/(const Dual2<N> &a, double b) { return a * (1.0 / b); }
template <int N> __device__ __forceinline__ Dual2<N> operator/(double a, const Dual2<N> &b) { return d2_recip(b) * a; }
template <int N> __device__ __forceinline__ Dual2<N> r_div(const Dual2<N> &a, const Dual2<N> &b) { return a / b; }
template <int N> __device__ __forceinline__ Dual2<N> r_div(const Dual2<N> &a, double b) { return a * MCRE_RCP(b); }
template <int N> __device__ __forceinline__ Dual2<N> r_div(double a, const Dual2<N> &b) { return a / b; }
template <int N> __device__ __forceinline__ Dual2<N> operator+(const Dual2<N> &a, double b) { Dual2<N> r = a; r.v += b; return r; }
template <int N> __device__ __forceinline__ Dual2<N> operator+(double b, const Dual2<N> &a) { Dual2<N> r = a; r.v += b; return r; }
template <int N> __device__ __forceinline__ Dual2<N> operator-(const Dual2<N> &a, double b) { Dual2<N> r = a; r.v -= b; return r; }
template <int N> __device__ __forceinline__ Dual2<N> operator-(double b, const Dual2<N> &a) { Dual2<N> r = -a; r.v += b; return r; }
template <int N> __device__ __forceinline__ Dual2<N> &operator+=(Dual2<N> &a, const Dual2<N> &b) { a = a + b; return a; }
template <int N> __device__ __forceinline__ Dual2<N> &operator+=(Dual2<N> &a, double b) { a.v += b; return a; }

template <int N> __device__ __forceinline__ Dual2<N> r_exp(const Dual2<N> &x) { const double e = MCRE_EXP(x.v); return d2_chain(x, e, e, e); }
template <int N> __device__ __forceinline__ Dual2<N> r_exp_small(const Dual2<N> &x) { const double e = r_exp_small(x.v); return d2_chain(x, e, e, e); }
template <int N> __device__ __forceinline__ Dual2<N> r_log(const Dual2<N> &x) {
  const double inv = MCRE_RCP(x.v);
  return d2_chain(x, MCRE_LOG(x.v), inv, -inv * inv);
}
template <int N> __device__ __forceinline__ Dual2<N> r_sqrt(const Dual2<N> &x) {
  const double s = MCRE_SQRT(x.v), f1 = s > 0.0 ? 0.5 / s : 0.0;
  return d2_chain(x, s, f1, s > 0.0 ? -0.5 * f1 / x.v : 0.0);
}
template <int N> __device__ __forceinline__ Dual2<N> r_sqrt_pos(const Dual2<N> &x) {
  const double s = r_sqrt_pos(x.v), f1 = 0.5 / s;
  return d2_chain(x, s, f1, -0.5 * f1 / x.v);
}
template <int N> __device__ __forceinline__ Dual2<N> r_relu(const Dual2<N> &x) { return d2_select(x, x.v > 0.0, 0.0); }
template <int N> __device__ __forceinline__ Dual2<N> r_max(const Dual2<N> &x, double c) { return d2_select(x, x.v >= c, c); }
template <int N> __device__ __forceinline__ Dual2<N> r_mask(const Dual2<N> &x, bool keep) { return d2_select(x, keep, 0.0); }
template <int N> __device__ __forceinline__ Dual2<N> r_with_value(const Dual2<N> &x, double v) { Dual2<N> r = x; r.v = v; return r; }
template <int N> __device__ __forceinline__ Dual2<N> r_fuzzy(const Dual2<N> &x, bool fuzzy, double eps) {
  if (!fuzzy) { Dual2<N> r = dual2_zero<N>(); r.v = x.v > 0.0 ? 1.0 : 0.0; return r; }
  const double t = (x.v + eps) / (2.0 * eps);
  const bool in = t >= 0.0 && t <= 1.0;
  return d2_chain(x, fmin(fmax(t, 0.0), 1.0), in ? 1.0 / (2.0 * eps) : 0.0, 0.0);
}

template <int N> struct RealTraits<Dual2<N> > {
  static const int NT = N + Dual2<N>::M;
  __device__ static __forceinline__ Dual2<N> zero() { return dual2_zero<N>(); }
  __device__ static __forceinline__ Dual2<N> load(const double *p, int idx) {
    Dual2<N> r; const double *q = p + (size_t)idx * (NT + 1); r.v = __ldg(q);
#pragma unroll
    for (int i = 0; i < N; ++i) r.d[i] = __ldg(q + 1 + i);
#pragma unroll
    for (int i = 0; i < Dual2<N>::M; ++i) r.h[i] = __ldg(q + 1 + N + i);
    return r;
  }
  __device__ static __forceinline__ Dual2<N> lift(double c) { Dual2<N> r = dual2_zero<N>(); r.v = c; return r; }
};
template <int N> struct RealVar<Dual2<N> > {
  __device__ static __forceinline__ Dual2<N> make(double v, int k) {
    Dual2<N> r = dual2_zero<N>(); r.v = v;
#pragma unroll
    for (int i = 0; i < N; ++i) r.d[i] = (i == k) ? 1.0 : 0.0;
    return r;
  }
};

#undef MCRE_D2_PAIRS
}  // namespace mcre
