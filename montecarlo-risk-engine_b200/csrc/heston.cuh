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
template <typename R>
__device__ __forceinline__ void heston_euler_step(const R &sigma, const R &rate, const R &kappa, const R &theta,
                                                  double dt, double sq, const R &w0, const R &w1, R &logS, R &v) {
  const R vp = r_sqrt(r_max(v, 0.0));
  logS = logS + (rate - 0.5 * v) * dt + vp * sq * w0;
  v = r_max(v + kappa * (theta - v) * dt + sigma * vp * sq * w1, 0.0);
}

}  // namespace mcre
