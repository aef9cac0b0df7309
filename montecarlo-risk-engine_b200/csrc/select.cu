// Exact order statistics per row by MSB-first radix select (8 bits per pass, 8 passes)
// on the order-preserving 64-bit image of IEEE doubles.
//
// Takes the place of torch.sort(e).values[q_index (+-1)] in PFEMetric
// (src/metrics/pfe_metric.py:59-71): the PFE needs three adjacent order statistics per
// exposure date, not a sorted array.  Each pass is count (histogram of the next digit
// among elements matching the prefix found so far) + scan (pick the bucket holding the
// wanted rank).  With paths sharded over GPUs the caller sums the histograms over ranks
// between the two steps; everything else is local.
#include "common.cuh"

namespace mcre {

__device__ __forceinline__ unsigned long long key_of(double x) {
  unsigned long long u = (unsigned long long)__double_as_longlong(x);
  return (u >> 63) ? ~u : (u | 0x8000000000000000ull);
}
__device__ __forceinline__ double value_of(unsigned long long k) {
  unsigned long long u = (k >> 63) ? (k & 0x7fffffffffffffffull) : ~k;
  return __longlong_as_double((long long)u);
}

constexpr int SEL_MAX_R = 4;

// grid: (blocks_x, n_rows).  hist: [n_rows][R][256].
// The first version spent 76 instructions per element (a loop over rank slots reading its tables from shared
// memory, 2.2 TB/s, issue bound - profiles/r01_other_kernels_ncu_summary.md); now the distinct prefixes of the
// row sit in registers, the match is one AND + compare against the already-found high bits, and four
// independent loads are in flight per thread.
// After the second pass the rows are compacted (select_compact_kernel): `compact` [n_rows][cap] then holds the
// elements that still match one of the row's prefixes, ccount[row] how many (> cap: the row did not fit and keeps
// reading the full data).
__global__ void __launch_bounds__(256) select_count_kernel(const double *__restrict__ values, long long row_stride,
                                                           long long n_local, int R, int pass,
                                                           const unsigned long long *__restrict__ prefix,
                                                           unsigned long long *hist, const double *__restrict__ compact,
                                                           const unsigned int *__restrict__ ccount, long long cap) {
  __shared__ unsigned int sh[SEL_MAX_R][256];
  const int row = blockIdx.y;
  for (int i = threadIdx.x; i < SEL_MAX_R * 256; i += blockDim.x) (&sh[0][0])[i] = 0u;
  // rank slots that share a prefix are counted once (adjacent ranks usually do): distinct prefixes pd[0..nd),
  // slot r reads the histogram of rep[r]
  unsigned long long pd[SEL_MAX_R];
  int rep[SEL_MAX_R];
  int nd = 0;
#pragma unroll
  for (int r = 0; r < SEL_MAX_R; ++r) {
    pd[r] = ~0ull; rep[r] = 0;
    if (r < R) {
      const unsigned long long pr = prefix[(size_t)row * R + r];
      int found = -1;
#pragma unroll
      for (int q = 0; q < SEL_MAX_R; ++q) if (q < nd && found < 0 && pd[q] == pr) found = q;
      if (found < 0) {
#pragma unroll
        for (int q = 0; q < SEL_MAX_R; ++q) if (q == nd) pd[q] = pr;
        found = nd++;
      }
      rep[r] = found;
    }
  }
  __syncthreads();
  const int shift = 56 - 8 * pass;
  // bits above the current digit (none in pass 0); the prefixes only have those bits set
  const unsigned long long high = pass == 0 ? 0ull : (~0ull << (shift + 8));
  const double *v = values + (size_t)row * row_stride;
  long long nblk = gridDim.x;
  if (compact) {
    const unsigned int cc = ccount[row];
    if ((long long)cc <= cap) {
      // a compacted row is small: four blocks read it (fewer histogram flushes), the others have nothing to do
      if (blockIdx.x >= 4) return;
      v = compact + (size_t)row * cap; n_local = cc; nblk = gridDim.x < 4 ? gridDim.x : 4;
    }
  }
  auto count = [&](double x) {
    const unsigned long long k = key_of(x);
    const unsigned int digit = (unsigned int)(k >> shift) & 0xffu;
    const unsigned long long kh = k & high;
#pragma unroll
    for (int q = 0; q < SEL_MAX_R; ++q)
      if (q < nd && kh == pd[q]) atomicAdd(&sh[q][digit], 1u);
  };
  const long long stride = nblk * blockDim.x;
  long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  for (; i + 3 * stride < n_local; i += 4 * stride) {
    const double x0 = __ldg(v + i), x1 = __ldg(v + i + stride), x2 = __ldg(v + i + 2 * stride),
                 x3 = __ldg(v + i + 3 * stride);
    count(x0); count(x1); count(x2); count(x3);
  }
  for (; i < n_local; i += stride) count(__ldg(v + i));
  __syncthreads();
  for (int j = threadIdx.x; j < R * 256; j += blockDim.x) {
    const int r = j >> 8, d = j & 255;
    int q = 0;
#pragma unroll
    for (int t = 0; t < SEL_MAX_R; ++t) if (t == r) q = rep[t];
    const unsigned int c = sh[q][d];
    if (c) atomicAdd(&hist[((size_t)row * R + r) * 256 + d], (unsigned long long)c);
  }
}

// Writes the elements of every row that match one of its prefixes (all bits above `shift` fixed) to
// compact[row][0..).  Matches are staged in shared memory and flushed with ONE global atomic per ~1000 of them:
// a returning atomic per warp on the row's counter serialised at the L2 and made the first version of this pass
// three times slower than a plain read of the data.
constexpr int SEL_STAGE = 2048;
__global__ void __launch_bounds__(256) select_compact_kernel(const double *__restrict__ values, long long row_stride,
                                                             long long n_local, int R, int shift,
                                                             const unsigned long long *__restrict__ prefix,
                                                             double *__restrict__ compact, unsigned int *ccount,
                                                             long long cap) {
  __shared__ double stage[SEL_STAGE];
  __shared__ unsigned int n_stage, base_out;
  const int row = blockIdx.y;
  unsigned long long pd[SEL_MAX_R];
#pragma unroll
  for (int r = 0; r < SEL_MAX_R; ++r) pd[r] = r < R ? prefix[(size_t)row * R + r] : prefix[(size_t)row * R];
  const unsigned long long high = ~0ull << shift;
  const double *v = values + (size_t)row * row_stride;
  double *out = compact + (size_t)row * cap;
  const long long stride = (long long)gridDim.x * blockDim.x;
  if (threadIdx.x == 0) n_stage = 0;
  __syncthreads();
  auto flush = [&]() {     // called by all threads of the block
    __syncthreads();
    const unsigned int cnt = n_stage;
    if (threadIdx.x == 0 && cnt) base_out = atomicAdd(ccount + row, cnt);
    __syncthreads();
    for (unsigned int j = threadIdx.x; j < cnt; j += blockDim.x) {
      const long long pos = (long long)base_out + j;
      if (pos < cap) out[pos] = stage[j];
    }
    __syncthreads();
    if (threadIdx.x == 0) n_stage = 0;
    __syncthreads();
  };
  // four independent loads in flight per thread; uniform trip count per block
  for (long long base = (long long)blockIdx.x * blockDim.x; base < n_local; base += 4 * stride) {
    double x[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const long long i = base + j * stride + threadIdx.x;
      x[j] = i < n_local ? __ldg(v + i) : __longlong_as_double(0x7ff8000000000001ll);   // a NaN payload no prefix holds
    }
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const long long i = base + j * stride + threadIdx.x;
      const unsigned long long kh = key_of(x[j]) & high;
      bool keep = false;
#pragma unroll
      for (int r = 0; r < SEL_MAX_R; ++r) keep = keep || kh == pd[r];
      if (keep && i < n_local) stage[atomicAdd(&n_stage, 1u)] = x[j];
    }
    // at most 4 * 256 new entries per iteration: flush while another iteration may not fit.  The decision must be
    // uniform over the block (flush() contains barriers): n_stage is read between two barriers, so no thread can
    // start the next iteration's atomicAdd before every thread holds the same value.
    __syncthreads();
    const unsigned int staged = n_stage;
    __syncthreads();
    if (staged > SEL_STAGE - 4 * 256) flush();
  }
  flush();
}

// one thread per (row, rank slot)
__global__ void select_scan_kernel(int n, int pass, const unsigned long long *hist, unsigned long long *prefix,
                                   long long *remaining) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const unsigned long long *h = hist + (size_t)i * 256;
  long long k = remaining[i];
  unsigned long long cum = 0;
  int d = 0;
  for (; d < 255; ++d) {
    if (cum + h[d] > (unsigned long long)k) break;
    cum += h[d];
  }
  remaining[i] = k - (long long)cum;
  prefix[i] |= (unsigned long long)d << (56 - 8 * pass);
}

__global__ void select_finish_kernel(int n, const unsigned long long *prefix, double *out) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) out[i] = value_of(prefix[i]);
}

// smallest local index whose value equals the row's target bit for bit (LLONG_MAX: none here)
__global__ void __launch_bounds__(256) select_locate_kernel(const double *__restrict__ values, long long row_stride,
                                                            long long n_local, const double *__restrict__ target,
                                                            long long *index) {
  const int row = blockIdx.y;
  const unsigned long long want = key_of(target[row]);
  const double *v = values + (size_t)row * row_stride;
  long long best = 0x7fffffffffffffffll;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n_local; i += (long long)gridDim.x * blockDim.x)
    if (key_of(v[i]) == want && i < best) best = i;
  if (best != 0x7fffffffffffffffll) atomicMin(index + row, best);
}

}  // namespace mcre

using namespace mcre;

extern "C" int mcre_select_locate(const double *d_values, int64_t row_stride, int64_t n_local, int32_t n_rows,
                                  const double *d_targets, int64_t *d_index, void *stream) {
  if (!d_targets || !d_index || n_rows <= 0) return fail(-1, "select locate: bad argument%s", "");
  cudaStream_t st = (cudaStream_t)stream;
  MCRE_CUDA(cudaMemsetAsync(d_index, 0x7f, (size_t)n_rows * 8, st));   // 0x7f7f...: larger than any index
  if (n_local <= 0) return 0;
  if (!d_values) return fail(-1, "select locate: null values%s", "");
  long long bx = (n_local + 256 * 8 - 1) / (256 * 8);
  const long long cap = (long long)sm_count() * 16 / n_rows + 1;
  if (bx > cap) bx = cap;
  dim3 grid((unsigned)bx, (unsigned)n_rows);
  select_locate_kernel<<<grid, 256, 0, st>>>(d_values, row_stride, n_local, d_targets, (long long *)d_index);
  MCRE_LAUNCHED();
  return 0;
}

struct mcre_select_plan {
  int n_rows, R;
  unsigned long long *prefix = nullptr;
  long long *remaining = nullptr;
  // candidates left after the first two passes (mcre_select_compact)
  double *compact = nullptr;
  unsigned int *ccount = nullptr;
  long long cap = 0;
  bool compacted = false;
};

extern "C" int mcre_select_create(int32_t n_rows, int32_t n_ranks_per_row, mcre_select_plan **out) {
  if (!out || n_rows <= 0 || n_ranks_per_row <= 0 || n_ranks_per_row > SEL_MAX_R)
    return fail(-1, "select: invalid shape%s", "");
  mcre_select_plan *p = new mcre_select_plan();
  p->n_rows = n_rows; p->R = n_ranks_per_row;
  const size_t n = (size_t)n_rows * n_ranks_per_row;
  MCRE_CUDA(cudaMalloc((void **)&p->prefix, n * 8));
  MCRE_CUDA(cudaMalloc((void **)&p->remaining, n * 8));
  *out = p;
  return 0;
}
extern "C" void mcre_select_destroy(mcre_select_plan *p) {
  if (!p) return;
  if (p->prefix) cudaFree(p->prefix);
  if (p->remaining) cudaFree(p->remaining);
  if (p->compact) cudaFree(p->compact);
  if (p->ccount) cudaFree(p->ccount);
  delete p;
}
extern "C" int mcre_select_begin(mcre_select_plan *p, const int64_t *ranks, void *stream) {
  if (!p || !ranks) return fail(-1, "null argument%s", "");
  const size_t n = (size_t)p->n_rows * p->R;
  p->compacted = false;
  MCRE_CUDA(cudaMemsetAsync(p->prefix, 0, n * 8, (cudaStream_t)stream));
  MCRE_CUDA(cudaMemcpyAsync(p->remaining, ranks, n * 8, cudaMemcpyHostToDevice, (cudaStream_t)stream));
  MCRE_CUDA(cudaStreamSynchronize((cudaStream_t)stream));  // `ranks` is pageable host memory
  return 0;
}
extern "C" int mcre_select_count(mcre_select_plan *p, const double *d_values, int64_t row_stride, int64_t n_local,
                                 int32_t pass, uint64_t *d_hist, void *stream) {
  if (!p || !d_hist || pass < 0 || pass > 7) return fail(-1, "select: bad argument%s", "");
  cudaStream_t st = (cudaStream_t)stream;
  MCRE_CUDA(cudaMemsetAsync(d_hist, 0, (size_t)p->n_rows * p->R * 256 * 8, st));
  if (n_local <= 0) return 0;
  if (!d_values) return fail(-1, "select: null values%s", "");
  long long bx = (n_local + 256 * 8 - 1) / (256 * 8);
  const long long cap = (long long)sm_count() * 16 / p->n_rows + 1;
  if (bx > cap) bx = cap;
  if (bx < 1) bx = 1;
  dim3 grid((unsigned)bx, (unsigned)p->n_rows);
  select_count_kernel<<<grid, 256, 0, st>>>(d_values, row_stride, n_local, p->R, pass, p->prefix,
                                            (unsigned long long *)d_hist, p->compacted ? p->compact : nullptr, p->ccount,
                                            p->cap);
  MCRE_LAUNCHED();
  return 0;
}
// After `passes_done` passes (>= 2: the top 16 bits separate sign, exponent and 4 mantissa bits) keep only the
// elements that can still be one of the wanted order statistics: the remaining passes read ~1/16 of the data
// or less instead of all of it (rows with more than n_local / 8 candidates, e.g. all paths equal at t = 0,
// keep reading the full row).  Local to the rank: histograms of compacted and full rows add up the same way.
extern "C" int mcre_select_compact(mcre_select_plan *p, const double *d_values, int64_t row_stride, int64_t n_local,
                                   int32_t passes_done, void *stream) {
  if (!p || passes_done < 1 || passes_done > 7) return fail(-1, "select compact: bad argument%s", "");
  if (n_local < 65536) return 0;     // not worth a pass
  if (!d_values) return fail(-1, "select: null values%s", "");
  cudaStream_t st = (cudaStream_t)stream;
  const long long cap = n_local / 8 + 1024;
  if (!p->compact || p->cap != cap) {
    if (p->compact) { cudaFree(p->compact); p->compact = nullptr; }
    if (cudaMalloc((void **)&p->compact, (size_t)p->n_rows * cap * sizeof(double)) != cudaSuccess) {
      cudaGetLastError();
      p->compact = nullptr;
      return 0;                      // no memory for the shortcut: the remaining passes read the full rows
    }
    p->cap = cap;
  }
  if (!p->ccount) MCRE_CUDA(cudaMalloc((void **)&p->ccount, (size_t)p->n_rows * sizeof(unsigned int)));
  MCRE_CUDA(cudaMemsetAsync(p->ccount, 0, (size_t)p->n_rows * sizeof(unsigned int), st));
  long long bx = (n_local + 256 * 8 - 1) / (256 * 8);
  const long long capb = (long long)sm_count() * 16 / p->n_rows + 1;
  if (bx > capb) bx = capb;
  dim3 grid((unsigned)bx, (unsigned)p->n_rows);
  select_compact_kernel<<<grid, 256, 0, st>>>(d_values, row_stride, n_local, p->R, 64 - 8 * passes_done, p->prefix,
                                              p->compact, p->ccount, cap);
  MCRE_LAUNCHED();
  p->compacted = true;
  return 0;
}

extern "C" int mcre_select_scan(mcre_select_plan *p, int32_t pass, const uint64_t *d_hist, void *stream) {
  if (!p || !d_hist) return fail(-1, "null argument%s", "");
  const int n = p->n_rows * p->R;
  select_scan_kernel<<<(n + 127) / 128, 128, 0, (cudaStream_t)stream>>>(n, pass, (const unsigned long long *)d_hist,
                                                                        p->prefix, p->remaining);
  MCRE_LAUNCHED();
  return 0;
}
extern "C" int mcre_select_finish(mcre_select_plan *p, double *d_out, void *stream) {
  if (!p || !d_out) return fail(-1, "null argument%s", "");
  const int n = p->n_rows * p->R;
  select_finish_kernel<<<(n + 127) / 128, 128, 0, (cudaStream_t)stream>>>(n, p->prefix, d_out);
  MCRE_LAUNCHED();
  return 0;
}
