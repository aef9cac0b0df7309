// Test hook: evaluates the branch-free elementary functions of fastmath.cuh on an array,
// so tests/test_fastmath_gpu.py can measure their ulp error against numpy.
#define MCRE_FAST_MATH 2
#include "common.cuh"
#include "fastmath.cuh"

namespace mcre {
__global__ void fastmath_probe_kernel(int fn, const double *x, double *y, long long n) {
  fm_tables_init();
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
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
    case 13: fm_sincos2pi_t(v, s, c); y[i] = s; break;
    case 14: fm_sincos2pi_t(v, s, c); y[i] = c; break;
    default: y[i] = 0.0;
  }
}
}  // namespace mcre

extern "C" int mcre_fastmath_probe(int32_t fn, const double *d_x, double *d_y, int64_t n, void *stream) {
  if (!d_x || !d_y || n < 0) return mcre::fail(-1, "null argument%s", "");
  if (n == 0) return 0;
  mcre::fastmath_probe_kernel<<<(unsigned)((n + 255) / 256), 256, 0, (cudaStream_t)stream>>>(fn, d_x, d_y, n);
  MCRE_LAUNCHED();
  return 0;
}
