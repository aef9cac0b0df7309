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

// Transposed variant for 128-thread blocks and few values per call (the value-only interest-rate kernel: 4 - 16 sums
// per metric date): every thread stores its NV values to shared memory, then G = 128 / NV' threads per value
// (NV' = NV rounded up to a power of two) each add 128 / G staged entries in a fixed order and finish with log2(G)
// shuffles - 2 NV + 128 / G + log2 G instructions per thread instead of the 10 NV of NV shuffle trees, and the same
// single barrier per call.  stage: [2][NV][128].  Summation order fixed => reproducible bit for bit.
template <int NV>
__device__ __forceinline__ void block_accumulate_t128(const double (&vals)[NV], double *acc, int slot, double *stage,
                                                      int &parity) {
  constexpr int NVP = NV <= 4 ? 4 : NV <= 8 ? 8 : NV <= 16 ? 16 : 32;
  constexpr int G = 128 / NVP;          // threads per value: 32, 16, 8 or 4 (a group never spans warps)
  constexpr int PER = 128 / G;          // staged entries per thread
  static_assert(NV >= 4 && NV <= 32, "block_accumulate_t128: 4..32 values per call");
  const int tid = threadIdx.x;
  double *st = stage + (size_t)parity * NV * 128;
#pragma unroll
  for (int i = 0; i < NV; ++i) st[i * 128 + tid] = vals[i];
  __syncthreads();
  const int v = tid / G, g = tid % G;
  const int vr = v < NV ? v : NV - 1;   // (idle groups of a non-power-of-two NV redo the last value: all lanes shuffle)
  double s = 0.0;
#pragma unroll
  for (int j = 0; j < PER; ++j) s += st[vr * 128 + g + j * G];
#pragma unroll
  for (int off = G / 2; off > 0; off >>= 1) s += __shfl_xor_sync(0xffffffffu, s, off, G);
  if (g == 0 && v < NV) acc[slot + v] += s;
  parity ^= 1;
}

}  // namespace mcre
