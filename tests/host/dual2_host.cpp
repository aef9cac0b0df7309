// Test infrastructure: compiles csrc/dual.cuh + csrc/dual2.cuh (first- and second-order forward-mode numbers of the
// kernels) for the HOST, so that their arithmetic can be checked against torch's double backward on a machine without a
// GPU (tests/test_dual2_host.py).  Nothing in the product links this file.
#include <cmath>
#include <cstddef>
#include <cstdint>

#define __device__
#define __host__
#define __forceinline__ inline
#define __ldg(p) (*(p))
#include "../../montecarlo-risk-engine_b200/csrc/dual2.cuh"

using namespace mcre;

// A payoff-like composite of every operation the kernels use, on R = Dual<3> or Dual2<3> seeded at x = (spot, vol, rate):
// mean over the draws z of [relu, fuzzy indicator, division, sqrt, log, reciprocal, clamp, mask, exp] x discount factor.
template <typename R>
static R composite(const double *x, const double *z, int nz) {
  typedef RealVar<R> V;
  const R s = V::make(x[0], 0), v = V::make(x[1], 1), r = V::make(x[2], 2);
  R acc = RealTraits<R>::zero();
  for (int i = 0; i < nz; ++i) {
    const R S = s * r_exp((r - 0.5 * v * v) * 2.0 + v * (1.4142135623730951 * z[i]));
    R pay = r_relu(S - 1.2) * r_fuzzy(S - 1.0, true, 0.5) / r_sqrt(S);
    pay = pay + r_log(S) * r_log(S) - 1.0 / S;
    pay = pay + r_max(S, 1.5) * r_div(S, v) + r_mask(S * S, z[i] > 0.0) - r_div(2.0, S + 3.0) + (-S) * 0.25;
    pay += r_sqrt_pos(S + 1.0) * r_exp_small(r * 0.01);
    acc += pay * r_exp(-(r * 2.0));
  }
  return acc * (1.0 / (double)nz);
}

// out: value, 3 first derivatives, 6 second derivatives (upper triangle, row-major)
extern "C" void dual2_host_second(const double *x, const double *z, int nz, double *out) {
  const Dual2<3> y = composite<Dual2<3> >(x, z, nz);
  out[0] = val(y);
  for (int k = 0; k < 9; ++k) out[1 + k] = tan_of(y, k);
}
extern "C" void dual2_host_first(const double *x, const double *z, int nz, double *out) {
  const Dual<3> y = composite<Dual<3> >(x, z, nz);
  out[0] = val(y);
  for (int k = 0; k < 3; ++k) out[1 + k] = tan_of(y, k);
}
