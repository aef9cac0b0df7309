// Library-wide utilities: error state, launch accounting, the fixed-order tree reduction
// of per-chunk partial sums, and the FP64 FMA peak probe used as roofline denominator.
#include "common.cuh"

namespace mcre {
thread_local char g_err[512] = "";
std::atomic<long long> g_launches{0};

int sm_count() {
  static int cached = 0;
  if (cached == 0) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&cached, cudaDevAttrMultiProcessorCount, dev);
    if (cached <= 0) cached = 148;
  }
  return cached;
}

// Balanced pairwise sum over chunks with a binary-counter stack: element i is pushed at
// level 0 and equal-level neighbours are merged immediately.  For a power-of-two count
// this is exactly the balanced binary tree over chunk indices, so a rank that owns an
// aligned power-of-two block of chunks produces a node of the global tree and the result
// is independent of how many GPUs the chunks were spread over.
__global__ void tree_reduce_kernel(const double *partial, long long n_chunks, long long n_slots, double *out) {
  const long long slot = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (slot >= n_slots) return;
  double stack[48];
  int top = 0;
  for (long long c = 0; c < n_chunks; ++c) {
    double v = partial[c * n_slots + slot];
    long long idx = c + 1;
    // merge while the low bit of the running count is clear: (c+1) has trailing zeros
    while ((idx & 1) == 0) { v = stack[--top] + v; idx >>= 1; }
    stack[top++] = v;
  }
  double s = 0.0;
  bool first = true;
  // leftover levels (non power-of-two counts): fold from the most recent to the oldest
  while (top > 0) {
    double v = stack[--top];
    s = first ? v : v + s;
    first = false;
  }
  out[slot] = s;
}

// Register-resident DFMA chains: 16 independent accumulators per thread.
__global__ void dfma_peak_kernel(double *sink, int iters, double a, double b) {
  double x[16];
#pragma unroll
  for (int i = 0; i < 16; ++i) x[i] = (double)(threadIdx.x + i) * 1e-3;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 16; ++i) x[i] = fma(x[i], a, b);
  }
  double s = 0.0;
#pragma unroll
  for (int i = 0; i < 16; ++i) s += x[i];
  if (s == 12345.6789) sink[0] = s;  // never true; keeps the chains alive
}
}  // namespace mcre

using namespace mcre;

extern "C" int mcre_tree_reduce(const double *d_partial, int64_t n_chunks, int64_t n_slots, double *d_out,
                                void *stream) {
  if (!d_partial || !d_out) return fail(-1, "null argument%s", "");
  if (n_slots <= 0) return 0;
  const int threads = 128;
  const unsigned blocks = (unsigned)((n_slots + threads - 1) / threads);
  tree_reduce_kernel<<<blocks, threads, 0, (cudaStream_t)stream>>>(d_partial, n_chunks, n_slots, d_out);
  MCRE_LAUNCHED();
  return 0;
}

extern "C" int mcre_dfma_peak(double *tflops_out, void *stream) {
  if (!tflops_out) return fail(-1, "null argument%s", "");
  cudaStream_t st = (cudaStream_t)stream;
  double *sink = nullptr;
  MCRE_CUDA(cudaMalloc((void **)&sink, 8));
  const int threads = 256, blocks = sm_count() * 8, iters = 1 << 14;
  cudaEvent_t e0, e1;
  MCRE_CUDA(cudaEventCreate(&e0));
  MCRE_CUDA(cudaEventCreate(&e1));
  double best = 0.0;
  for (int rep = 0; rep < 6; ++rep) {
    MCRE_CUDA(cudaEventRecord(e0, st));
    dfma_peak_kernel<<<blocks, threads, 0, st>>>(sink, iters, 1.0000001, 1e-9);
    MCRE_LAUNCHED();
    MCRE_CUDA(cudaEventRecord(e1, st));
    MCRE_CUDA(cudaEventSynchronize(e1));
    float ms = 0.f;
    MCRE_CUDA(cudaEventElapsedTime(&ms, e0, e1));
    const double flops = 2.0 * 16.0 * (double)iters * threads * blocks;
    const double tf = flops / (ms * 1e-3) * 1e-12;
    if (rep > 0 && tf > best) best = tf;
  }
  cudaEventDestroy(e0);
  cudaEventDestroy(e1);
  cudaFree(sink);
  *tflops_out = best;
  return 0;
}

extern "C" int64_t mcre_launch_count(void) { return g_launches.load(); }
extern "C" const char *mcre_last_error(void) { return g_err; }
extern "C" int mcre_abi_version(void) { return MCRE_ABI_VERSION; }
