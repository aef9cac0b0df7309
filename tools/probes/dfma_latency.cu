// Micro-probe (not part of the library): FP64 pipe issue rate and dependent-issue latency on
// this GPU, as a function of independent DFMA chains per thread and resident warps per SM
// sub-partition.  Build + run on the GPU box:
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o /tmp/dfma_latency tools/probes/dfma_latency.cu && /tmp/dfma_latency
#include <cstdio>
#include <cuda_runtime.h>

template <int CH>
__global__ void chains(double *sink, int iters, double a, double b, long long *cyc) {
  double x[CH];
#pragma unroll
  for (int i = 0; i < CH; ++i) x[i] = (double)(threadIdx.x + i) * 1e-3;
  long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int k = 0; k < 8; ++k)
#pragma unroll
      for (int i = 0; i < CH; ++i) x[i] = fma(x[i], a, b);
  }
  long long t1 = clock64();
  double s = 0.0;
#pragma unroll
  for (int i = 0; i < CH; ++i) s += x[i];
  if (s == 12345.6789) sink[0] = s;
  if (threadIdx.x == 0 && blockIdx.x == 0) cyc[0] = t1 - t0;
}

// mixed: one DFMA chain + INT chains, to see whether integer instructions issue in the shadow of FP64
template <int NI>
__global__ void mixed(double *sink, int iters, double a, double b, long long *cyc, unsigned mul, unsigned add) {
  double x[4];
  unsigned v[NI > 0 ? NI : 1];
#pragma unroll
  for (int i = 0; i < 4; ++i) x[i] = (double)(threadIdx.x + i) * 1e-3;
#pragma unroll
  for (int i = 0; i < NI; ++i) v[i] = threadIdx.x + i;
  long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int k = 0; k < 8; ++k) {
#pragma unroll
      for (int i = 0; i < 4; ++i) x[i] = fma(x[i], a, b);
#pragma unroll
      for (int i = 0; i < NI; ++i) v[i] = (v[i] * mul + add) ^ (v[i] >> 7);   // runtime constants + xorshift: not foldable
    }
  }
  long long t1 = clock64();
  double s = 0.0;
  unsigned w = 0;
#pragma unroll
  for (int i = 0; i < 4; ++i) s += x[i];
#pragma unroll
  for (int i = 0; i < NI; ++i) w ^= v[i];
  if (s == 12345.6789 || w == 0x12345u) sink[0] = s;
  if (threadIdx.x == 0 && blockIdx.x == 0) cyc[0] = t1 - t0;
}

int main() {
  double *sink; long long *cyc, h;
  cudaMalloc(&sink, 8); cudaMalloc(&cyc, 8);
  int sms; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
  const int iters = 4096;
  printf("chains warps/SMSP cycles_per_DFMA_warp_instr(per SMSP)\n");
  for (int wps = 1; wps <= 8; wps *= 2) {
#define RUN(CH)                                                                         \
  {                                                                                     \
    chains<CH><<<sms, 128 * wps>>>(sink, iters, 1.0000001, 1e-9, cyc);                   \
    cudaDeviceSynchronize();                                                            \
    chains<CH><<<sms, 128 * wps>>>(sink, iters, 1.0000001, 1e-9, cyc);                   \
    cudaDeviceSynchronize();                                                            \
    cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);                                     \
    printf("%d %d %.3f\n", CH, wps, (double)h / ((double)iters * 8 * CH * wps));        \
  }
    RUN(1) RUN(2) RUN(4) RUN(8)
  }
  printf("mixed: 4 DFMA chains + NI IMAD chains, 4 warps/SMSP: cycles per (4 DFMA + NI IMAD) group per warp\n");
#define RUNM(NI)                                                                        \
  {                                                                                     \
    mixed<NI><<<sms, 512>>>(sink, iters, 1.0000001, 1e-9, cyc, 0xD2511F53u, 0x9E3779B9u); \
    cudaDeviceSynchronize();                                                            \
    mixed<NI><<<sms, 512>>>(sink, iters, 1.0000001, 1e-9, cyc, 0xD2511F53u, 0x9E3779B9u); \
    cudaDeviceSynchronize();                                                            \
    cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);                                     \
    printf("NI=%d: %.3f cycles per group per SMSP-warp slot (DFMA-only bound 8.0)\n", NI, (double)h / ((double)iters * 8 * 4)); \
  }
  RUNM(0) RUNM(1) RUNM(2) RUNM(4) RUNM(6) RUNM(8)
  printf("(each IMAD chain step = 3 integer instructions: IMAD, SHF, LOP3)\n");
  return 0;
}
