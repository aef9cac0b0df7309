// Generic path generator: every model of the reference, any ModelConfig composition,
// paths materialised as [n_paths][n_dates][state_dim].
//
// This is the compatibility seam for MonteCarloEngine.generate_paths()
// (src/engine/engine.py:27-123) and the direct model-level parity check; the fused
// kernels never materialise paths.  One thread per path, state in a small local array,
// runtime dispatch on the model kind (uniform across the warp).
#include "common.cuh"
#include "philox.cuh"
#include "dual.cuh"
#include "heston.cuh"
#include "launch.cuh"

namespace mcre {

constexpr int PATHS_MAX_DIM = 16;

struct PathsDev {
  int n_models;
  const int *kind, *nassets, *flags, *param_off, *state_off, *noise_off;
  const double *params;
  int scheme, noise_dim, state_dim, n_sub, n_dates, n_pre_dates;
  const double *step_dt, *step_t1;
  const int *step_date, *step_chol;
  const double *chol, *step_aux, *init_state;
};

__device__ __forceinline__ void step_bs(const double *p, int scheme, double dt, double sq, double *s, const double *w) {
  const double sigma = p[1], rate = p[2];
  if (scheme == MCRE_SCHEME_ANALYTICAL) s[0] = s[0] * exp(rate * dt + (w[0] - 0.5 * dt * sigma * sigma));
  else s[0] = s[0] + (rate * s[0] * dt + sigma * s[0] * sq * w[0]);
}
__device__ __forceinline__ void step_bsm(const double *p, int na, int scheme, double dt, double sq, double *s,
                                         const double *w) {
  const double rate = p[2 * na];
  for (int i = 0; i < na; ++i) {
    const double sig = p[na + i];
    if (scheme == MCRE_SCHEME_ANALYTICAL) s[i] = s[i] * exp((rate - 0.5 * sig * sig) * dt + w[i]);
    else s[i] = s[i] + (rate * s[i] * dt + sig * s[i] * sq * w[i]);
  }
}
__device__ __forceinline__ void step_vasicek(const double *p, int scheme, double dt, double sq, double theta_t,
                                             double *s, const double *w) {
  const double sigma = p[1], a = p[3];
  const double r = s[0];
  s[1] = s[1] + r * dt;
  if (scheme == MCRE_SCHEME_ANALYTICAL) s[0] = (theta_t + (r - theta_t) * exp(-a * dt)) + w[0];
  else s[0] = r + a * (theta_t - r) * dt + sigma * sq * w[0];
}
__device__ __forceinline__ void step_cirpp(const double *p, bool deterministic, double dt, double sq, const double *aux,
                                           double *s, const double *w) {
  if (deterministic) {
    s[1] = s[1] + aux[1] * dt;
    s[0] = aux[2];
    return;
  }
  const double kappa = p[0], theta = p[1], sigma = p[2];
  const double y = s[0];
  const double yn = y + kappa * (theta - y) * dt + sigma * sqrt(fmax(y, 0.0)) * sq * w[0];
  s[1] = s[1] + (y + aux[0]) * dt;
  s[0] = fmax(yn, 1e-12);
}
__device__ __forceinline__ void step_schwartz(const double *p, int scheme, double dt, double sq, const double *aux,
                                              double *s, const double *w) {
  const double kappa = p[1], ss = p[2], mu = p[3], sl = p[4];
  double x, y;
  if (scheme == MCRE_SCHEME_ANALYTICAL) {
    const double xm = fabs(kappa) <= 1e-12 ? s[1] : s[1] * exp(-kappa * dt);
    x = xm + w[0];
    y = s[2] + mu * dt + w[1];
  } else {
    x = s[1] - kappa * s[1] * dt + ss * sq * w[0];
    y = s[2] + mu * dt + sl * sq * w[1];
  }
  s[0] = aux[0] + x + y;
  s[1] = x; s[2] = y;
}
__global__ void __launch_bounds__(128) paths_kernel(PathsDev P, RngDev rng, long long path_begin, long long n_paths,
                                                    double *out) {
  const long long lpath = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (lpath >= n_paths) return;
  const long long gpath = path_begin + lpath;
  double s[PATHS_MAX_DIM], z[PATHS_MAX_DIM], w[PATHS_MAX_DIM];
  const int D = P.state_dim, d = P.noise_dim;
  for (int i = 0; i < D; ++i) s[i] = P.init_state[i];
  NormalStream ns; ns.init(rng, (unsigned long long)gpath);
  double *o = out + (size_t)lpath * P.n_dates * D;
  for (int di = 0; di < P.n_pre_dates; ++di)
    for (int i = 0; i < D; ++i) o[(size_t)di * D + i] = s[i];
  for (int is = 0; is < P.n_sub; ++is) {
    if (rng.mode == MCRE_RNG_INJECT) {
      const double *zp = rng.z + ((size_t)is * rng.n_total + gpath) * d;
      for (int j = 0; j < d; ++j) z[j] = zp[j];
    } else {
      for (int j = 0; j < d; ++j) z[j] = ns.next();
    }
    const double *L = P.chol + (size_t)P.step_chol[is] * d * d;
    for (int i = 0; i < d; ++i) {
      double acc = 0.0;
      for (int j = 0; j <= i; ++j) acc += L[i * d + j] * z[j];
      w[i] = acc;
    }
    const double dt = P.step_dt[is], sq = sqrt(dt);
    for (int m = 0; m < P.n_models; ++m) {
      const double *p = P.params + P.param_off[m];
      double *sm = s + P.state_off[m];
      const double *wm = w + P.noise_off[m];
      const double *aux = P.step_aux + ((size_t)is * P.n_models + m) * 4;
      switch (P.kind[m]) {
        case MCRE_MODEL_BS: step_bs(p, P.scheme, dt, sq, sm, wm); break;
        case MCRE_MODEL_BSM: step_bsm(p, P.nassets[m], P.scheme, dt, sq, sm, wm); break;
        case MCRE_MODEL_VASICEK: step_vasicek(p, P.scheme, dt, sq, aux[3], sm, wm); break;
        case MCRE_MODEL_CIRPP: step_cirpp(p, (P.flags[m] & 1) != 0, dt, sq, aux, sm, wm); break;
        case MCRE_MODEL_SCHWARTZ2F: step_schwartz(p, P.scheme, dt, sq, aux, sm, wm); break;
        case MCRE_MODEL_HESTON: {
          // params: spot, sigma, rate, rho, kappa, theta, v0
          if (P.scheme == MCRE_SCHEME_QE) {
            const double u = rng.mode == MCRE_RNG_INJECT ? rng.u[(size_t)is * rng.n_total + gpath]
                                                         : ns.uniform((uint32_t)is);
            heston_qe_step<double>(p[1], p[2], p[3], p[4], p[5], dt, (P.flags[m] & 2) != 0, wm[0], wm[1], u, sm[0], sm[1]);
          } else {
            heston_euler_step<double>(p[1], p[2], p[4], p[5], dt, sq, wm[0], wm[1], sm[0], sm[1]);
          }
          break;
        }
        default: break;
      }
    }
    const int di = P.step_date[is];
    if (di >= 0)
      for (int i = 0; i < D; ++i) o[(size_t)di * D + i] = s[i];
  }
}

// Correlated joint noise w = L z of every (sub-step, path), materialised: [n_sub][n_paths][dim].  z is the Philox
// stream (normal # is * dim + j of the path) or the injected reference stream.  Used by hybrid books whose model
// families run in separate fused kernels on slices of one joint draw (mcre/hybrid.py).
__global__ void __launch_bounds__(128) correlated_normals_kernel(RngDev rng, int n_sub, int dim, const double *__restrict__ chol,
                                                                 long long n_paths, double *__restrict__ out) {
  const long long gp = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (gp >= n_paths) return;
  NormalStream ns; ns.init(rng, (unsigned long long)gp);
  double z[PATHS_MAX_DIM];
  for (int is = 0; is < n_sub; ++is) {
    if (rng.mode == MCRE_RNG_INJECT) {
      const double *zp = rng.z + ((size_t)is * rng.n_total + gp) * dim;
      for (int j = 0; j < dim; ++j) z[j] = zp[j];
    } else {
      for (int j = 0; j < dim; ++j) z[j] = ns.next();
    }
    double *o = out + ((size_t)is * n_paths + gp) * dim;
    for (int i = 0; i < dim; ++i) {
      double acc = 0.0;
      for (int j = 0; j <= i; ++j) acc += chol[i * dim + j] * z[j];
      o[i] = acc;
    }
  }
}

}  // namespace mcre

using namespace mcre;

extern "C" int mcre_correlated_normals(const mcre_rng *rng, int32_t n_sub, int32_t dim, const double *chol,
                                       int64_t n_paths, double *d_out, void *stream) {
  if (!rng || !chol || !d_out) return fail(-1, "null argument%s", "");
  if (dim < 1 || dim > PATHS_MAX_DIM || n_sub < 0) return fail(-2, "correlated normals: 1..16 noise dimensions%s", "");
  if (rng->mode == MCRE_RNG_INJECT && (!rng->d_z || rng->n_paths_total < n_paths))
    return fail(-1, "inject mode without (enough) normals%s", "");
  if (n_paths <= 0 || n_sub == 0) return 0;
  DevArray<double> L;
  int rc = L.upload(chol, (size_t)dim * dim);
  if (rc) return rc;
  const unsigned blocks = (unsigned)((n_paths + 127) / 128);
  correlated_normals_kernel<<<blocks, 128, 0, (cudaStream_t)stream>>>(make_rng(rng), n_sub, dim, L.p, n_paths, d_out);
  g_launches.fetch_add(1);
  cudaError_t e = cudaGetLastError();
  if (e == cudaSuccess) e = cudaStreamSynchronize((cudaStream_t)stream);   // the factor is freed below
  L.release();
  return e == cudaSuccess ? 0 : cuda_fail(e, "correlated normals kernel");
}

extern "C" int mcre_generate_paths(const mcre_paths_desc *c, const mcre_rng *rng, const mcre_shard *shard,
                                   double *d_paths, void *stream) {
  if (!c || !rng || !shard || !d_paths) return fail(-1, "null argument%s", "");
  if (c->state_dim > PATHS_MAX_DIM || c->noise_dim > PATHS_MAX_DIM)
    return fail(-1, "paths: at most 16 state / noise dimensions%s", "");
  if (rng->mode == MCRE_RNG_INJECT && !rng->d_z) return fail(-1, "inject mode without normals%s", "");
  if (rng->mode == MCRE_RNG_INJECT && c->scheme == MCRE_SCHEME_QE && !rng->d_u)
    return fail(-1, "inject mode: QE needs uniforms%s", "");
  const int nm = c->n_models;
  std::vector<int> poff(nm), soff(nm), noff(nm);
  int po = 0, so = 0, no = 0, n_chol = 0;
  for (int m = 0; m < nm; ++m) {
    poff[m] = po; soff[m] = so; noff[m] = no;
    const int na = c->model_nassets[m];
    switch (c->model_kind[m]) {
      case MCRE_MODEL_BS: po += 3; so += 1; no += 1; break;
      case MCRE_MODEL_BSM: po += 2 * na + 1; so += na; no += na; break;
      case MCRE_MODEL_HESTON: po += 7; so += 2; no += 2; break;
      case MCRE_MODEL_VASICEK: po += 4; so += 2; no += 1; break;
      case MCRE_MODEL_CIRPP: po += 4; so += 2; no += 1; break;
      case MCRE_MODEL_SCHWARTZ2F: po += 6; so += 3; no += 2; break;
      default: return fail(-1, "paths: unknown model kind%s", "");
    }
  }
  if (so != c->state_dim || no != c->noise_dim) return fail(-1, "paths: state / noise dimensions do not add up%s", "");
  for (int s = 0; s < c->n_sub; ++s) if (c->step_chol[s] + 1 > n_chol) n_chol = c->step_chol[s] + 1;
  DevArray<int> kind, nassets, flags, dpoff, dsoff, dnoff, step_date, step_chol;
  DevArray<double> params, step_dt, step_t1, chol, aux, init;
  int rc = 0;
#define UP(a, h, n) if (!rc) rc = a.upload(h, (size_t)(n))
  UP(kind, c->model_kind, nm); UP(nassets, c->model_nassets, nm); UP(flags, c->model_flags, nm);
  UP(dpoff, poff.data(), nm); UP(dsoff, soff.data(), nm); UP(dnoff, noff.data(), nm);
  UP(params, c->model_params, po); UP(step_dt, c->step_dt, c->n_sub); UP(step_t1, c->step_t1, c->n_sub);
  UP(step_date, c->step_date, c->n_sub); UP(step_chol, c->step_chol, c->n_sub);
  UP(chol, c->chol, (size_t)n_chol * c->noise_dim * c->noise_dim);
  UP(aux, c->step_aux, (size_t)c->n_sub * nm * 4); UP(init, c->init_state, c->state_dim);
#undef UP
  if (!rc && shard->n_paths > 0) {
    PathsDev P;
    P.n_models = nm; P.kind = kind.p; P.nassets = nassets.p; P.flags = flags.p; P.param_off = dpoff.p;
    P.state_off = dsoff.p; P.noise_off = dnoff.p; P.params = params.p; P.scheme = c->scheme;
    P.noise_dim = c->noise_dim; P.state_dim = c->state_dim; P.n_sub = c->n_sub; P.n_dates = c->n_dates;
    P.n_pre_dates = c->n_pre_dates; P.step_dt = step_dt.p; P.step_t1 = step_t1.p; P.step_date = step_date.p;
    P.step_chol = step_chol.p; P.chol = chol.p; P.step_aux = aux.p; P.init_state = init.p;
    RngDev r;
    r.mode = rng->mode; r.k0 = (uint32_t)rng->seed; r.k1 = (uint32_t)rng->stream; r.z = rng->d_z; r.u = rng->d_u;
    r.n_total = rng->n_paths_total;
    const int threads = 128;
    const unsigned blocks = (unsigned)((shard->n_paths + threads - 1) / threads);
    paths_kernel<<<blocks, threads, 0, (cudaStream_t)stream>>>(P, r, shard->path_begin, shard->n_paths, d_paths);
    g_launches.fetch_add(1);
    cudaError_t e = cudaGetLastError();
    if (e == cudaSuccess) e = cudaStreamSynchronize((cudaStream_t)stream);  // tables are freed below
    if (e != cudaSuccess) rc = cuda_fail(e, "paths kernel");
  }
  kind.release(); nassets.release(); flags.release(); dpoff.release(); dsoff.release(); dnoff.release();
  step_date.release(); step_chol.release(); params.release(); step_dt.release(); step_t1.release(); chol.release();
  aux.release(); init.release();
  return rc;
}
