// Test infrastructure: compiles csrc/fastmath.cuh (the device elementary functions of the fused kernels) for the
// HOST, so that their accuracy can be checked on a machine without a GPU (tests/test_fastmath_host.py).  The shims
// below stand in for the CUDA intrinsics; MUFU.RSQ64H / MUFU.RCP64H are modelled pessimistically (they read and
// write the high 32 bits of a double only).  Nothing in the product links this file.
#include <cmath>
#include <cstdint>
#include <cstring>

#define MCRE_HOST_EMU 1
#define MCRE_FAST_MATH 2
#define __device__
#define __host__
#define __forceinline__ inline
#define __shared__
#define __constant__ const
struct double2 { double x, y; };
static inline double2 make_double2(double x, double y) { return double2{x, y}; }
static struct { int x; } threadIdx = {0}, blockDim = {1};
static inline void __syncthreads() {}
static inline double __hiloint2double(int hi, int lo) {
  uint64_t v = ((uint64_t)(uint32_t)hi << 32) | (uint32_t)lo;
  double d; memcpy(&d, &v, 8); return d;
}
static inline int __double2hiint(double d) { uint64_t v; memcpy(&v, &d, 8); return (int)(v >> 32); }
static inline int __double2loint(double d) { uint64_t v; memcpy(&v, &d, 8); return (int)(uint32_t)v; }
static inline double trunc_hi(double d) { return __hiloint2double(__double2hiint(d), 0); }
static inline void sincospi(double x, double *s, double *c) {
  const long double a = 3.14159265358979323846264338327950288L * (long double)x;
  *s = (double)sinl(a); *c = (double)cosl(a);
}
using std::fma;
namespace mcre {
static inline double fm_rcp_approx(double x) { return trunc_hi(1.0 / trunc_hi(x)); }
static inline double fm_rsqrt_approx(double x) { return trunc_hi(1.0 / std::sqrt(trunc_hi(x))); }
}
#include "../../montecarlo-risk-engine_b200/csrc/fastmath.cuh"

using namespace mcre;
extern "C" void fm_host_init() { fm_tables_init(); }
// same function numbering as csrc/fastmath_probe.cu
extern "C" void fm_host_eval(int fn, const double *x, double *y, long long n) {
  for (long long i = 0; i < n; ++i) {
    const double v = x[i];
    double s, c;
    switch (fn) {
      case 0: y[i] = fm_exp(v); break;
      case 1: y[i] = fm_log(v); break;
      case 2: y[i] = fm_sqrt(v); break;
      case 3: fm_sincos2pi(v, s, c); y[i] = s; break;
      case 4: fm_sincos2pi(v, s, c); y[i] = c; break;
      case 5: y[i] = fm_div(1.0, v); break;
      case 10: y[i] = fm_exp_t(v); break;
      case 11: y[i] = fm_log_t(v); break;
      case 12: y[i] = fm_sqrt_pos(v); break;
      case 13: fm_sincos2pi_t(v, s, c); y[i] = s; break;
      case 14: fm_sincos2pi_t(v, s, c); y[i] = c; break;
      case 15: { const double a[1] = {v}; double o[1]; fm_neg2log_tv<1>(a, o); y[i] = o[0]; break; }
      case 16: y[i] = fm_exp_small(v); break;
      case 17: case 18: {
        // Box-Muller tail on the mantissa double d = 1 + v with radius 1: (cos, sin)(2 pi v)
        const double d[1] = {1.0 + v}, r[1] = {1.0}; double zc[1], zs[1];
        fm_polar_tv<1>(d, r, zc, zs); y[i] = fn == 17 ? zc[0] : zs[0]; break;
      }
      default: y[i] = 0.0;
    }
  }
}
