// Deterministic block reductions into shared-memory accumulators.
//
// A block owns one fixed chunk of global path ids at a time.  Per call, every thread
// contributes NV doubles; each is summed over the warp with a fixed shuffle tree, the
// per-warp sums are staged in shared memory and summed in warp order, and the block
// total is added to acc[slot..slot+NV).  The summation order never depends on timing
// or on how many GPUs share the path range, so partial sums are reproducible bit for
// bit.  The stage buffer is double-buffered by `parity`, so one
// __syncthreads() per call is enough.
#pragma once

namespace mcre {

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
  for (int off = 16; off > 0; off >>= 1) v += __shfl_xor_sync(0xffffffffu, v, off);
  return v;
}

// stage: [2][n_warps][NVMAX]; acc: block accumulator in shared memory.
template <int NV>
__device__ __forceinline__ void block_accumulate(const double (&vals)[NV], double *acc, int slot,
                                                 double *stage, int nvmax, int &parity) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
  double *st = stage + (size_t)parity * nw * nvmax;
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    double s = warp_sum(vals[i]);
    if (lane == 0) st[warp * nvmax + i] = s;
  }
  __syncthreads();
  if (threadIdx.x < NV) {
    double s = 0.0;
    for (int w = 0; w < nw; ++w) s += st[w * nvmax + threadIdx.x];
    acc[slot + threadIdx.x] += s;
  }
  parity ^= 1;
}

}  // namespace mcre
