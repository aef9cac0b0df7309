// Longstaff-Schwartz backward induction on spilled pre-simulation arrays.
//
// The reference regresses, for every regression date t_k (latest first), the deflated
// future cashflows of the product under the exercise policy already fitted for later dates
// (src/controller/controller.py:294-383): it rolls compute_normalized_cashflows over the
// product dates [t_next, last) for every starting state with float32 accumulators, adds
// the cached tail, and solves a tall least-squares problem in [1, x, x^2].
// For single-right exercise products (BermudanOption / AmericanOption,
// src/products/bermudan_option.py:93-188) state 0 (exercised) never carries value, so the
// roll collapses to one float32 running value V per path:
//     V <- fp32( fp32(ex ? imm_i / N_i : 0) + (ex ? 0 : V) ),   ex = imm_i > phi(x_i) . coef_i
// and the tall least squares to 8 moments per date.  One fused launch per regression date:
// HBM-bound, 5 coalesced f64 streams + the f32 value array (36 B / path / date).
#include "common.cuh"
#include "reduce.cuh"

namespace mcre {

constexpr int LSM_NV = 8;

__global__ void __launch_bounds__(256) lsm_step_kernel(const double *__restrict__ xk, const double *__restrict__ nk,
                                                       double shift_k, double scale_k, const double *__restrict__ xi,
                                                       const double *__restrict__ ni, const double *__restrict__ imm,
                                                       int has_coef, double c0, double c1, double c2, double shift_i,
                                                       double scale_i, float *__restrict__ value, long long n, int chunk,
                                                       double *__restrict__ partial) {
  __shared__ double acc[LSM_NV];
  __shared__ double stage[2 * 8 * LSM_NV];
  const long long n_chunks = (n + chunk - 1) / chunk;
  for (long long ch = blockIdx.x; ch < n_chunks; ch += gridDim.x) {
    if (threadIdx.x < LSM_NV) acc[threadIdx.x] = 0.0;
    __syncthreads();
    int parity = 0;
    // each thread first sums its own paths of the chunk (stride blockDim), the block reduces once
    double vals[LSM_NV];
#pragma unroll
    for (int i = 0; i < LSM_NV; ++i) vals[i] = 0.0;
    for (int it = 0; it < chunk; it += blockDim.x) {
      const long long p = ch * chunk + it + threadIdx.x;
      if (it + (int)threadIdx.x < chunk && p < n) {
        float v = value[p];
        if (imm) {
          const double im = imm[p];
          double cont = 0.0;
          if (has_coef) {
            const double u = (xi[p] - shift_i) * scale_i;
            cont = c0 + u * (c1 + u * c2);
          }
          const bool ex = im > cont;
          // float32 step value updated with a float64 cashflow, then float32 + float32
          // (controller.py:330-349)
          const float step = (float)(ex ? im / ni[p] : 0.0);
          v = step + (ex ? 0.0f : v);
          value[p] = v;
        }
        const double u = (xk[p] - shift_k) * scale_k;
        const double y = nk[p] * (double)v;   // numeraire * total cashflows (controller.py:368)
        const double u2 = u * u;
        vals[0] += 1.0; vals[1] += u; vals[2] += u2; vals[3] += u2 * u; vals[4] += u2 * u2;
        vals[5] += y; vals[6] += y * u; vals[7] += y * u2;
      }
    }
    block_accumulate<LSM_NV>(vals, acc, 0, stage, LSM_NV, parity);
    __syncthreads();
    if (threadIdx.x < LSM_NV) partial[ch * LSM_NV + threadIdx.x] = acc[threadIdx.x];
    __syncthreads();
  }
}

// Gathers the regression / exercise inputs of an equity exercise product from materialised paths
// [n_paths][n_dates][state_dim]: one thread per path, date-major outputs (coalesced writes).
__global__ void __launch_bounds__(256) lsm_prepare_equity_kernel(const double *__restrict__ paths, long long n, int n_dates,
                                                                 int state_dim, int n_reg, const int *reg_date,
                                                                 const double *reg_num, int n_ex, const int *ex_date,
                                                                 int x_col, int x_is_log, int n_under, const int *ucol,
                                                                 const double *uw, const int *ulog, double strike,
                                                                 double sign, double *x, double *num, double *imm) {
  const long long p = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= n) return;
  const double *row = paths + (size_t)p * n_dates * state_dim;
  for (int k = 0; k < n_reg; ++k) {
    const double v = row[(size_t)reg_date[k] * state_dim + x_col];
    x[(size_t)k * n + p] = x_is_log ? exp(v) : v;
    num[(size_t)k * n + p] = reg_num[k];
  }
  for (int i = 0; i < n_ex; ++i) {
    const double *st = row + (size_t)ex_date[i] * state_dim;
    double U = 0.0;
    for (int j = 0; j < n_under; ++j) {
      const double v = st[ucol[j]];
      U += uw[j] * (ulog[j] ? exp(v) : v);
    }
    imm[(size_t)i * n + p] = fmax((U - strike) * sign, 0.0);
  }
}

}  // namespace mcre

using namespace mcre;

extern "C" int mcre_lsm_prepare_equity(const double *d_paths, int64_t n_paths, int32_t n_dates, int32_t state_dim,
                                       int32_t n_reg, const int32_t *reg_date, const double *reg_numeraire, int32_t n_ex,
                                       const int32_t *ex_date, int32_t x_col, int32_t x_is_log, int32_t n_under,
                                       const int32_t *under_col, const double *under_w, const int32_t *under_is_log,
                                       double strike, double sign, double *d_x, double *d_n, double *d_imm, void *stream) {
  if (!d_paths || !reg_date || !reg_numeraire || !ex_date || !under_col || !under_w || !under_is_log || !d_x || !d_n || !d_imm)
    return fail(-1, "null argument%s", "");
  if (n_paths <= 0) return 0;
  cudaStream_t st = (cudaStream_t)stream;
  DevArray<int> rd, ed, uc, ul;
  DevArray<double> rn, uw;
  int rc = rd.upload(reg_date, n_reg);
  if (!rc) rc = ed.upload(ex_date, n_ex);
  if (!rc) rc = uc.upload(under_col, n_under);
  if (!rc) rc = ul.upload(under_is_log, n_under);
  if (!rc) rc = rn.upload(reg_numeraire, n_reg);
  if (!rc) rc = uw.upload(under_w, n_under);
  if (!rc) {
    lsm_prepare_equity_kernel<<<(unsigned)((n_paths + 255) / 256), 256, 0, st>>>(
        d_paths, n_paths, n_dates, state_dim, n_reg, rd.p, rn.p, n_ex, ed.p, x_col, x_is_log, n_under, uc.p, uw.p, ul.p,
        strike, sign, d_x, d_n, d_imm);
    g_launches.fetch_add(1, std::memory_order_relaxed);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) rc = cuda_fail(e, "kernel launch");
    else cudaStreamSynchronize(st);   // the temporary index tables are freed below
  }
  rd.release(); ed.release(); uc.release(); ul.release(); rn.release(); uw.release();
  return rc;
}

extern "C" int mcre_lsm_step(const double *d_xk, const double *d_nk, double shift_k, double scale_k, const double *d_xi,
                             const double *d_ni, const double *d_imm, const double *coef_i, double shift_i,
                             double scale_i, float *d_value, int64_t n, int32_t chunk_paths, double *d_partial,
                             double *d_moments, void *stream) {
  if (!d_xk || !d_nk || !d_value || !d_partial || !d_moments) return fail(-1, "null argument%s", "");
  if (d_imm && (!d_xi || !d_ni)) return fail(-1, "lsm: exercise update needs x_i and N_i%s", "");
  if (chunk_paths <= 0 || chunk_paths % 256 != 0) return fail(-2, "lsm: chunk_paths must be a positive multiple of 256%s", "");
  cudaStream_t st = (cudaStream_t)stream;
  if (n <= 0) {   // a rank without pre-simulation paths contributes zero moments
    MCRE_CUDA(cudaMemsetAsync(d_moments, 0, LSM_NV * sizeof(double), st));
    return 0;
  }
  const long long n_chunks = (n + chunk_paths - 1) / chunk_paths;
  long long grid = (long long)sm_count() * 8;
  if (grid > n_chunks) grid = n_chunks;
  lsm_step_kernel<<<(unsigned)grid, 256, 0, st>>>(d_xk, d_nk, shift_k, scale_k, d_xi, d_ni, d_imm, coef_i != nullptr,
                                                  coef_i ? coef_i[0] : 0.0, coef_i ? coef_i[1] : 0.0,
                                                  coef_i ? coef_i[2] : 0.0, shift_i, scale_i, d_value, n, chunk_paths,
                                                  d_partial);
  MCRE_LAUNCHED();
  return mcre_tree_reduce(d_partial, n_chunks, LSM_NV, d_moments, stream);
}
