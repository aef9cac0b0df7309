// Micro-probe (not part of the library): issue cost of FP64 instructions next to the integer
// instructions of Philox on this GPU.  Each kernel runs a loop body of NF DFMA (8 independent
// chains), NW IMAD.WIDE.U32 and NL LOP3 (8 independent chains each), interleaved, and reports
// cycles per loop body per warp scheduler (SMSP) with 1, 2 and 4 resident warps per SMSP.
// If integer work issued in the shadow of the 2-cycle DFMA, adding it would be free until the
// issue port (1 instruction / cycle) saturates; the table shows it is additive instead.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o /tmp/issue_mix tools/probes/issue_mix.cu && /tmp/issue_mix
#include <cstdio>
#include <cuda_runtime.h>

template <int NF, int NW, int NL, int NM>
__global__ void mix(double *sink, int iters, double a, double b, unsigned m, long long *cyc) {
  double x[8];
  unsigned long long w[8];
  unsigned l[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    x[i] = (double)(threadIdx.x + i) * 1e-3;
    w[i] = threadIdx.x * 7u + i;
    l[i] = threadIdx.x * 13u + i;
  }
  long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
    constexpr int N = NF > NW ? (NF > NL ? NF : NL) : (NW > NL ? NW : NL);
#pragma unroll
    for (int k = 0; k < N; ++k) {
      if (k < NF) asm volatile("fma.rn.f64 %0, %0, %1, %2;" : "+d"(x[k & 7]) : "d"(a), "d"(b));
      if (k < NW) {
        unsigned lo = (unsigned)w[k & 7];
        asm volatile("mul.wide.u32 %0, %1, %2;" : "=l"(w[k & 7]) : "r"(lo), "r"(m));
      }
      if (k < NL) asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(l[k & 7]) : "r"(m), "r"(l[(k + 1) & 7]));
      if (k < NM) asm volatile("mov.b64 %0, {%1, %2};" : "=d"(x[(k + 3) & 7]) : "r"(l[k & 7]), "r"(l[(k + 2) & 7]));
    }
  }
  long long t1 = clock64();
  double s = 0.0;
  unsigned long long q = 0;
#pragma unroll
  for (int i = 0; i < 8; ++i) { s += x[i]; q += w[i] + l[i]; }
  if (s == 12345.6789 || q == 0x1234567ull) sink[0] = s;
  if (threadIdx.x == 0 && blockIdx.x == 0) cyc[0] = t1 - t0;
}

template <int NF, int NW, int NL, int NM>
void run(const char *name, double *sink, long long *cyc, int sms) {
  const int iters = 2048;
  printf("%-34s", name);
  for (int wps = 1; wps <= 4; wps *= 2) {
    long long h = 0;
    for (int rep = 0; rep < 2; ++rep) {
      mix<NF, NW, NL, NM><<<sms, 128 * wps>>>(sink, iters, 1.0000001, 1e-9, 0xD2511F53u, cyc);
      cudaDeviceSynchronize();
    }
    cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
    printf("  %d w/SMSP: %7.2f", wps, (double)h / ((double)iters * wps));
  }
  printf("   (cycles per body per warp)\n");
}

int main() {
  double *sink; long long *cyc;
  cudaMalloc(&sink, 8); cudaMalloc(&cyc, 8);
  int sms; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
  run<64, 0, 0, 0>("64 DFMA", sink, cyc, sms);
  run<0, 64, 0, 0>("64 IMAD.WIDE", sink, cyc, sms);
  run<0, 0, 64, 0>("64 LOP3", sink, cyc, sms);
  run<0, 32, 32, 0>("32 IMAD.WIDE + 32 LOP3", sink, cyc, sms);
  run<64, 16, 0, 0>("64 DFMA + 16 IMAD.WIDE", sink, cyc, sms);
  run<64, 0, 16, 0>("64 DFMA + 16 LOP3", sink, cyc, sms);
  run<64, 0, 32, 0>("64 DFMA + 32 LOP3", sink, cyc, sms);
  run<64, 0, 64, 0>("64 DFMA + 64 LOP3", sink, cyc, sms);
  run<64, 16, 16, 0>("64 DFMA + 16 IMAD.WIDE + 16 LOP3", sink, cyc, sms);
  run<64, 20, 28, 0>("64 DFMA + 20 IMAD.WIDE + 28 LOP3", sink, cyc, sms);
  run<64, 32, 32, 0>("64 DFMA + 32 IMAD.WIDE + 32 LOP3", sink, cyc, sms);
  run<64, 0, 0, 16>("64 DFMA + 16 MOV pairs", sink, cyc, sms);
  return 0;
}
