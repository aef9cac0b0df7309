// Heston Euler and Andersen-QE sub-steps, templated on the scalar type (double or Dual<N>)
// so the path generator, the fused equity kernel and its tangent builds share one body.
#pragma once
#include "dual.cuh"

namespace mcre {

// Heston Euler / QE (src/models/heston.py:99-121, 161-253); templated on the scalar so the
// equity kernels reuse it with tangents.
template <typename R>
__device__ __forceinline__ void heston_qe_step(const R &sigma, const R &rate, const R &rho, const R &kappa,
                                               const R &theta, double dt, bool smooth, double zS, double zV, double u,
                                               R &logS, R &v) {
  const double eps = 1e-12;
  const R e = r_exp(-(kappa * dt));
  const R m = theta + (v - theta) * e;
  const R s2 = v * sigma * sigma * e * (1.0 - e) / kappa + theta * sigma * sigma * (1.0 - e) * (1.0 - e) / (2.0 * kappa);
  const R psi = s2 / (m * m + eps);
  const R invpsi = 1.0 / (psi + eps);
  const R tq = r_max(2.0 * invpsi - 1.0, 0.0);
  const R b2 = r_max(2.0 * invpsi - 1.0 + r_sqrt(2.0 * invpsi) * r_sqrt(tq), 0.0);
  const R b = r_sqrt(b2);
  const R a = m / (1.0 + b2);
  const R bz = b + zV;
  const R v1 = a * bz * bz;
  R p = (psi - 1.0) / (psi + 1.0);
  {  // clamp(p, 0, 1-1e-6) with pass-through gradient inside the band
    const double pv = val(p);
    if (pv < 0.0) p = RealTraits<R>::lift(0.0);
    else if (pv > 1.0 - 1e-6) p = RealTraits<R>::lift(1.0 - 1e-6);
  }
  const R beta = (1.0 - p) / (m + eps);
  const double one_minus_u = fmax(1.0 - u, eps);
  const R v_tail = r_log(r_max(1.0 - p, eps) / one_minus_u) / (beta + eps);
  const R v2 = r_fuzzy(u - p, smooth, 0.3) * v_tail;
  const R wq = r_fuzzy(psi - 1.5, smooth, 0.5);
  const R vn = (1.0 - wq) * v1 + wq * v2;
  const R K0 = -(rho * kappa * theta / sigma) * dt;
  const R K1 = (kappa * rho / sigma - 0.5) * dt - rho / sigma;
  const R K2 = rho / sigma;
  const R K3 = (1.0 - rho * rho) * dt;
  const R var_int = r_max(K3 * v + 0.0 * vn, 0.0);
  const R vol = r_sqrt(r_max(var_int, eps));
  logS = logS + rate * dt + K0 + K1 * v + K2 * vn + vol * zS;
  v = vn;
}
// ---- QE step with the step-size-dependent constants hoisted ---------------------------------
// Everything in heston.py:123-159 that depends only on the parameters and dt (the decay e, the two
// variance-moment coefficients, K0..K3, the drift) is computed once per distinct dt; the per-step
// work is then the psi-dependent part.  Outside the fuzzy band (psi <= 1 with smoothing, psi <= 1.5
// without) the exponential branch has weight exactly 0 and zero derivative, so it is skipped when
// the whole warp agrees (uniform branch) - its value is finite, so 0 * v2 = 0 exactly.
template <typename R>
struct QeStepConst {
  double dt;
  R e, c1, c2, drift, K1, K2, K3;
};
template <typename R>
__device__ __forceinline__ void heston_qe_prepare(const R &sigma, const R &rate, const R &rho, const R &kappa,
                                                  const R &theta, double dt, QeStepConst<R> &c) {
  c.dt = dt;
  c.e = r_exp(-(kappa * dt));
  const R s2k = r_div(sigma * sigma, kappa), ome = 1.0 - c.e;
  c.c1 = s2k * c.e * ome;                       // v coefficient of s^2 (heston.py:123-143)
  c.c2 = theta * s2k * ome * ome * 0.5;
  const R ros = r_div(rho, sigma);
  c.drift = rate * dt - ros * kappa * theta * dt;   // r dt + K0
  c.K1 = (kappa * ros - 0.5) * dt - ros;
  c.K2 = ros;
  c.K3 = (1.0 - rho * rho) * dt;
}
template <typename R>
__device__ __forceinline__ void heston_qe_step_c(const QeStepConst<R> &c, const R &theta, bool smooth, double zS,
                                                 double zV, double u, R &logS, R &v) {
  const double eps = 1e-12;
  const R m = theta + (v - theta) * c.e;
  const R s2 = v * c.c1 + c.c2;
  const R psi = r_div(s2, m * m + eps);
  const R invpsi = r_div(1.0, psi + eps);
  const R t2 = 2.0 * invpsi - 1.0;
  const R tq = r_max(t2, 0.0);
  const R b2 = r_max(t2 + r_sqrt(2.0 * invpsi) * r_sqrt(tq), 0.0);
  const R b = r_sqrt(b2);
  const R a = r_div(m, 1.0 + b2);
  const R bz = b + zV;
  const R v1 = a * bz * bz;
  const R wq = r_fuzzy(psi - 1.5, smooth, 0.5);
  R vn = (1.0 - wq) * v1;
  const bool need_tail = smooth ? (val(psi) >= 1.0) : (val(psi) > 1.5);
  if (__any_sync(0xffffffffu, need_tail)) {
    R p = r_div(psi - 1.0, psi + 1.0);
    {  // clamp(p, 0, 1-1e-6) with pass-through gradient inside the band
      const double pv = val(p);
      if (pv < 0.0) p = RealTraits<R>::lift(0.0);
      else if (pv > 1.0 - 1e-6) p = RealTraits<R>::lift(1.0 - 1e-6);
    }
    const R beta = r_div(1.0 - p, m + eps);
    const double one_minus_u = fmax(1.0 - u, eps);
    const R v_tail = r_div(r_log(r_div(r_max(1.0 - p, eps), one_minus_u)), beta + eps);
    const R v2 = r_fuzzy(u - p, smooth, 0.3) * v_tail;
    vn = vn + wq * v2;
  }
  const R var_int = r_max(c.K3 * v + 0.0 * vn, 0.0);
  const R vol = r_sqrt(r_max(var_int, eps));
  logS = logS + c.drift + c.K1 * v + c.K2 * vn + vol * zS;
  v = vn;
}

template <typename R>
__device__ __forceinline__ void heston_euler_step(const R &sigma, const R &rate, const R &kappa, const R &theta,
                                                  double dt, double sq, const R &w0, const R &w1, R &logS, R &v) {
  const R vp = r_sqrt(r_max(v, 0.0));
  logS = logS + (rate - 0.5 * v) * dt + vp * sq * w0;
  v = r_max(v + kappa * (theta - v) * dt + sigma * vp * sq * w1, 0.0);
}

}  // namespace mcre
