// Interest-rate / credit family: fused path generation + cashflows + exposure + metrics.
//
// One thread owns one path for its whole life: the Vasicek (and CIR++) state lives in
// registers, normals come from Philox in registers (or from the reference's injected
// draws), and at every simulation date the thread evaluates the netted cashflows, the
// regression-proxy exposure, threshold / MPoR collateral and the metric integrands.
// Only block-reduced sums ever reach HBM.  See include/mcre.h for what each entry point
// replaces in the reference.
#include "common.cuh"
#include "philox.cuh"
#include "dual.cuh"
#include "reduce.cuh"

namespace mcre {

struct IrcDev {
  int nt, scheme, has_cir, cir_det, vas_noise, cir_noise;
  const double *vas, *cir, *cir_init, *chol;
  int n_sub, n_dates, n_pre_dates;
  const double *step_dt; const int *step_date; const double *step_vas, *step_cir;
  const int *date_flags, *date_expo, *date_metric, *date_reg, *date_float_off;
  const double *float_coef, *float_inv_tau;
  int n_float;
  int n_sets, n_expo, n_metric, acc_flags;
  const double *set_fix, *set_float, *set_threshold; const int *set_flags, *set_lag;
  double *expo_coef;  // mutable: uploaded after the regression solve
  const double *expo_basis, *cva_coef;
  double lgd;
  int n_units, n_reg;
  const double *unit_fix, *unit_float; const double *reg_basis;
};

struct ShardDev {
  long long path_begin, n_paths;
  int chunk;
};

// ---- per-path model state ---------------------------------------------------------
template <typename R>
struct IrcState {
  R r, logB, y, logBl;
};

template <typename R, bool CIR>
struct IrcParams {
  R r0, sigma, theta, a;        // Vasicek
  R kappa, ctheta, csigma, y0;  // CIR++
  R L10, L11;                   // Cholesky rows used by the second noise column
  R L00;
};

// One sub-step of the joint model (src/models/vasicek.py:52-112, cirpp.py:155-198,
// model_config.py:223-276).  z0/z1 are the independent draws; the correlated noise is
// z @ L^T with L the lower Cholesky factor (model.py:46-48).
template <typename R, bool CIR, int SCHEME>
__device__ __forceinline__ void irc_step(const IrcDev &P, const IrcParams<R, CIR> &mp, IrcState<R> &s, int is,
                                         double z0, double z1) {
  typedef RealTraits<R> T;
  const double dt = __ldg(P.step_dt + is);
  const double sq = sqrt(dt);
  R w0 = mp.L00 * z0;
  R w1 = T::zero();
  if (CIR) w1 = mp.L10 * z0 + mp.L11 * z1;
  const R wv = (CIR && P.vas_noise == 1) ? w1 : w0;
  // numeraire integral uses the pre-step rate (left Riemann sum)
  s.logB = s.logB + s.r * dt;
  if (SCHEME == MCRE_SCHEME_ANALYTICAL) {
    R decay = T::load(P.step_vas, is * 2 + 0), nstd = T::load(P.step_vas, is * 2 + 1);
    // exact OU transition; the 1x1 Cholesky factor of the step covariance is nstd (vasicek.py:52-86)
    s.r = mp.theta + (s.r - mp.theta) * decay + nstd * z0;
  } else {
    s.r = s.r + mp.a * (mp.theta - s.r) * dt + mp.sigma * sq * wv;
  }
  if (CIR) {
    const R wc = (P.cir_noise == 1) ? w1 : w0;
    if (P.cir_det) {
      R lam1 = T::load(P.step_cir, is * 2 + 0), lam2 = T::load(P.step_cir, is * 2 + 1);
      s.logBl = s.logBl + lam1 * dt;
      s.y = lam2;
    } else {
      R psi = T::load(P.step_cir, is * 2 + 0);
      R ypos = r_relu(s.y);
      R yn = s.y + mp.kappa * (mp.ctheta - s.y) * dt + mp.csigma * r_sqrt(ypos) * sq * wc;
      s.logBl = s.logBl + (s.y + psi) * dt;
      s.y = r_max(yn, 1e-12);
    }
  }
}

template <typename R, bool CIR>
__device__ __forceinline__ void irc_load_params(const IrcDev &P, IrcParams<R, CIR> &mp) {
  typedef RealTraits<R> T;
  mp.r0 = T::load(P.vas, 0); mp.sigma = T::load(P.vas, 1); mp.theta = T::load(P.vas, 2); mp.a = T::load(P.vas, 3);
  mp.L00 = T::load(P.chol, 0);
  if (CIR) {
    mp.kappa = T::load(P.cir, 0); mp.ctheta = T::load(P.cir, 1); mp.csigma = T::load(P.cir, 2);
    mp.y0 = T::load(P.cir_init, 0);
    mp.L10 = T::load(P.chol, 2); mp.L11 = T::load(P.chol, 3);
  } else {
    mp.kappa = mp.ctheta = mp.csigma = mp.y0 = mp.L10 = mp.L11 = T::zero();
  }
}

template <typename R, bool CIR>
__device__ __forceinline__ void irc_draw(const RngDev &rng, NormalStream &ns, int is, long long lpath,
                                         long long gpath, double &z0, double &z1) {
  if (rng.mode == MCRE_RNG_INJECT) {
    const int d = CIR ? 2 : 1;
    const double *p = rng.z + ((size_t)is * rng.n_total + gpath) * d;
    z0 = p[0];
    z1 = CIR ? p[1] : 0.0;
  } else {
    if (CIR) ns.next2(z0, z1);
    else { z0 = ns.next(); z1 = 0.0; }
  }
}

// threshold dead-band (src/products/netting_set.py:48-72)
template <typename R>
__device__ __forceinline__ R apply_threshold(const R &x, double h) {
  const double v = val(x);
  if (v > h) return x - h;
  if (v < -h) return x + h;
  return RealTraits<R>::zero();
}

// =====================================================================================
// Main simulation kernel
// slot layout: [n_metric][NS][4+2NT] = pos, pos^2, neg, neg^2, d pos[NT], d neg[NT]
//              then [NS][4+2NT]      = pv, pv^2, cva, cva^2, d pv[NT], d cva[NT]
// Value slots hold sum(x - c) and sum((x - c)^2) with c = shift[slot], the value global
// path 0 takes (written by a one-path "pilot" launch of this same kernel).  Shifting by
// a sample of the distribution keeps the variance formula free of cancellation and makes
// degenerate dates (all paths equal, e.g. t = 0) give an exact zero Monte Carlo error.
// =====================================================================================
template <int NT, int NS, bool CIR, int SCHEME>
__global__ void __launch_bounds__(256) irc_main_kernel(IrcDev P, RngDev rng, ShardDev sh, double *partial,
                                                       double *spill, double *shift, int pilot) {
  typedef typename RealOf<NT>::type R;
  typedef RealTraits<R> T;
  constexpr int NV = 4 + 2 * NT;        // values per (set, date)
  constexpr int NVB = NS * NV;          // values per block_accumulate call
  extern __shared__ double smem[];
  const int nw = blockDim.x >> 5;
  const int n_slots = (P.n_metric + 1) * NVB;
  double *acc = smem;                   // [n_slots]
  double *stage = smem + n_slots;       // [2][nw][NVB]
  const long long n_chunks = (sh.n_paths + sh.chunk - 1) / sh.chunk;

  IrcParams<R, CIR> mp;
  irc_load_params<R, CIR>(P, mp);
  const int acc_flags = P.acc_flags;
  double thr[NS]; int sflags[NS];
#pragma unroll
  for (int s = 0; s < NS; ++s) {
    thr[s] = s < P.n_sets ? __ldg(P.set_threshold + s) : 0.0;
    sflags[s] = s < P.n_sets ? __ldg(P.set_flags + s) : 0;
  }

  for (long long chunk = blockIdx.x; chunk < n_chunks; chunk += gridDim.x) {
    for (int i = threadIdx.x; i < n_slots; i += blockDim.x) acc[i] = 0.0;
    __syncthreads();
    int parity = 0;
    for (int it = 0; it < sh.chunk; it += blockDim.x) {
      const long long lpath = chunk * sh.chunk + it + threadIdx.x;
      const bool live = lpath < sh.n_paths;
      const long long gpath = sh.path_begin + (live ? lpath : 0);
      NormalStream ns; ns.init(rng, (unsigned long long)gpath);
      IrcState<R> st;
      st.r = mp.r0; st.logB = T::zero(); st.y = mp.y0; st.logBl = T::zero();
      R pv[NS], cva[NS], hist[NS][MCRE_IRC_MAX_LAG];
#pragma unroll
      for (int s = 0; s < NS; ++s) {
        pv[s] = T::zero(); cva[s] = T::zero();
#pragma unroll
        for (int l = 0; l < MCRE_IRC_MAX_LAG; ++l) hist[s][l] = T::zero();
      }

      // ---- date evaluation (cashflows -> exposure -> metrics) ------------------------
      auto eval_date = [&](int di) {
        const int flags = __ldg(P.date_flags + di);
        if (!(flags & (MCRE_DATE_HAS_CASHFLOW | MCRE_DATE_HAS_EXPOSURE | MCRE_DATE_HAS_METRIC))) return;
        const R invN = r_exp(-st.logB);  // 1 / numeraire, numeraire = exp(logB) (vasicek.py:154-156)
        if ((flags & MCRE_DATE_HAS_CASHFLOW) && (acc_flags & MCRE_ACC_PV)) {
          R cf[NS];
#pragma unroll
          for (int s = 0; s < NS; ++s) cf[s] = T::lift(s < P.n_sets ? __ldg(P.set_fix + (size_t)s * P.n_dates + di) : 0.0);
          const int j0 = __ldg(P.date_float_off + di), j1 = __ldg(P.date_float_off + di + 1);
          for (int j = j0; j < j1; ++j) {
            // LIBOR from the bond price at the payment date's own short rate (bond.py:55-66)
            R alpha = T::load(P.float_coef, j * 2 + 0), B = T::load(P.float_coef, j * 2 + 1);
            R libor = (r_exp(B * st.r - alpha) - 1.0) * __ldg(P.float_inv_tau + j);
#pragma unroll
            for (int s = 0; s < NS; ++s)
              if (s < P.n_sets) cf[s] = cf[s] + libor * __ldg(P.set_float + (size_t)s * P.n_float + j);
          }
#pragma unroll
          for (int s = 0; s < NS; ++s) pv[s] = pv[s] + cf[s] * invN;
        }
        if (flags & MCRE_DATE_HAS_EXPOSURE) {
          const int e = __ldg(P.date_expo + di);
          const double shift = __ldg(P.expo_basis + e * 2), scale = __ldg(P.expo_basis + e * 2 + 1);
          const R u = (st.r - shift) * scale;
#pragma unroll
          for (int s = 0; s < NS; ++s) {
#pragma unroll
            for (int l = MCRE_IRC_MAX_LAG - 1; l > 0; --l) hist[s][l] = hist[s][l - 1];
            if (s < P.n_sets) {
              const int cb = (e * P.n_sets + s) * 3;
              R c0 = T::load(P.expo_coef, cb), c1 = T::load(P.expo_coef, cb + 1), c2 = T::load(P.expo_coef, cb + 2);
              hist[s][0] = (c0 + u * (c1 + u * c2)) * invN;  // continuation / numeraire (controller.py:438-447)
            }
          }
        }
        if (flags & MCRE_DATE_HAS_METRIC) {
          const int m = __ldg(P.date_metric + di);
          double vals[NVB];
          R surv = T::zero(), dflt = T::zero();
          const bool cva_date = (acc_flags & MCRE_ACC_CVA) && m < P.n_metric - 1;
          if (CIR && cva_date) {
            // S(0,t_k) = exp(-logB_lambda); S(t_k,t_k+1 | y) = C exp(-B y)   (cirpp.py:298-317)
            R C = T::load(P.cva_coef, m * 2), Bc = T::load(P.cva_coef, m * 2 + 1);
            surv = r_exp(-st.logBl);
            dflt = surv * (1.0 - C * r_exp(-(Bc * st.y)));
          }
#pragma unroll
          for (int s = 0; s < NS; ++s) {
            R unsec;
            if (sflags[s] & 1) {
              const int lag = s < P.n_sets ? __ldg(P.set_lag + (size_t)s * P.n_metric + m) : -1;
              R delayed = T::zero();
#pragma unroll
              for (int l = 0; l < MCRE_IRC_MAX_LAG; ++l) if (l == lag) delayed = hist[s][l];
              unsec = hist[s][0] - apply_threshold(delayed, thr[s]);
            } else {
              unsec = apply_threshold(hist[s][0], thr[s]);
            }
            const R pos = r_relu(unsec);
            const R neg = -r_relu(-unsec);
            if (cva_date && (sflags[s] & 2)) cva[s] = cva[s] + pos * dflt;
            const double keep = live ? 1.0 : 0.0;
            const int sb = m * NVB + s * NV;
            if (pilot) {
              if (threadIdx.x == 0) { shift[sb + 0] = val(pos); shift[sb + 2] = val(neg); }
            }
            const double dp = val(pos) - shift[sb + 0], dn = val(neg) - shift[sb + 2];
            vals[s * NV + 0] = keep * dp; vals[s * NV + 1] = keep * dp * dp;
            vals[s * NV + 2] = keep * dn; vals[s * NV + 3] = keep * dn * dn;
#pragma unroll
            for (int k = 0; k < NT; ++k) {
              vals[s * NV + 4 + k] = keep * tan_of(pos, k);
              vals[s * NV + 4 + NT + k] = keep * tan_of(neg, k);
            }
            if ((acc_flags & MCRE_ACC_SPILL) && live && s < P.n_sets)
              spill[((size_t)s * P.n_metric + m) * sh.n_paths + lpath] = val(unsec);
          }
          if ((acc_flags & (MCRE_ACC_POS | MCRE_ACC_NEG)) && !pilot)
            block_accumulate<NVB>(vals, acc, m * NVB, stage, NVB, parity);
        }
      };

      for (int di = 0; di < P.n_pre_dates; ++di) eval_date(di);
      for (int is = 0; is < P.n_sub; ++is) {
        double z0, z1;
        irc_draw<R, CIR>(rng, ns, is, lpath, gpath, z0, z1);
        irc_step<R, CIR, SCHEME>(P, mp, st, is, z0, z1);
        const int di = __ldg(P.step_date + is);
        if (di >= 0) eval_date(di);
      }
      // ---- per-path totals ------------------------------------------------------------
      {
        double vals[NVB];
        const double keep = live ? 1.0 : 0.0;
#pragma unroll
        for (int s = 0; s < NS; ++s) {
          const R c = cva[s] * P.lgd;
          const int sb = P.n_metric * NVB + s * NV;
          if (pilot) {
            if (threadIdx.x == 0) { shift[sb + 0] = val(pv[s]); shift[sb + 2] = val(c); }
          }
          const double dp = val(pv[s]) - shift[sb + 0], dc = val(c) - shift[sb + 2];
          vals[s * NV + 0] = keep * dp; vals[s * NV + 1] = keep * dp * dp;
          vals[s * NV + 2] = keep * dc; vals[s * NV + 3] = keep * dc * dc;
#pragma unroll
          for (int k = 0; k < NT; ++k) {
            vals[s * NV + 4 + k] = keep * tan_of(pv[s], k);
            vals[s * NV + 4 + NT + k] = keep * tan_of(c, k);
          }
        }
        if (!pilot) block_accumulate<NVB>(vals, acc, P.n_metric * NVB, stage, NVB, parity);
      }
    }
    if (pilot) return;
    __syncthreads();
    for (int i = threadIdx.x; i < n_slots; i += blockDim.x) partial[(size_t)chunk * n_slots + i] = acc[i];
    __syncthreads();
  }
}

// =====================================================================================
// Pre-simulation pass A: forward simulation, spills per regression date the explanatory
// variable, the numeraire and the FP32 window sums of each unit's discounted cashflows.
// scratch: x [n_reg][n] f64 | N [n_reg][n] f64 | W [n_units][n_reg][n] f32
// =====================================================================================
template <bool CIR, int SCHEME>
__global__ void __launch_bounds__(256) irc_presim_forward_kernel(IrcDev P, RngDev rng, ShardDev sh, double *xbuf,
                                                                 double *nbuf, float *wbuf) {
  typedef double R;
  typedef RealTraits<R> T;
  const long long lpath = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (lpath >= sh.n_paths) return;
  const long long gpath = sh.path_begin + lpath;
  const long long n = sh.n_paths;
  IrcParams<R, CIR> mp;
  irc_load_params<R, CIR>(P, mp);
  NormalStream ns; ns.init(rng, (unsigned long long)gpath);
  IrcState<R> st;
  st.r = mp.r0; st.logB = 0.0; st.y = mp.y0; st.logBl = 0.0;
  float W[MCRE_IRC_MAX_UNITS];
#pragma unroll
  for (int u = 0; u < MCRE_IRC_MAX_UNITS; ++u) W[u] = 0.0f;

  auto eval_date = [&](int di) {
    const int flags = __ldg(P.date_flags + di);
    if (!(flags & (MCRE_DATE_HAS_CASHFLOW | MCRE_DATE_HAS_REGRESSION))) return;
    const double numeraire = exp(st.logB);
    if (flags & MCRE_DATE_HAS_CASHFLOW) {
      double cf[MCRE_IRC_MAX_UNITS];
#pragma unroll
      for (int u = 0; u < MCRE_IRC_MAX_UNITS; ++u)
        cf[u] = u < P.n_units ? __ldg(P.unit_fix + (size_t)u * P.n_dates + di) : 0.0;
      const int j0 = __ldg(P.date_float_off + di), j1 = __ldg(P.date_float_off + di + 1);
      for (int j = j0; j < j1; ++j) {
        const double alpha = __ldg(P.float_coef + j * 2), B = __ldg(P.float_coef + j * 2 + 1);
        const double libor = (1.0 / exp(alpha - B * st.r) - 1.0) * __ldg(P.float_inv_tau + j);
#pragma unroll
        for (int u = 0; u < MCRE_IRC_MAX_UNITS; ++u)
          if (u < P.n_units) cf[u] += libor * __ldg(P.unit_float + (size_t)u * P.n_float + j);
      }
      // FP32 accumulator updated with an FP64 addend: W <- fp32(fp64(W) + cf)
      // (controller.py:330,341: float32 step_value += float64 cashflows)
#pragma unroll
      for (int u = 0; u < MCRE_IRC_MAX_UNITS; ++u) W[u] = (float)((double)W[u] + cf[u] / numeraire);
    }
    if (flags & MCRE_DATE_HAS_REGRESSION) {
      const int k = __ldg(P.date_reg + di);
      xbuf[(size_t)k * n + lpath] = st.r;
      nbuf[(size_t)k * n + lpath] = numeraire;
#pragma unroll
      for (int u = 0; u < MCRE_IRC_MAX_UNITS; ++u) {
        if (u < P.n_units && k > 0) wbuf[((size_t)u * P.n_reg + (k - 1)) * n + lpath] = W[u];
        W[u] = 0.0f;  // cashflows at or before the first regression date never enter a window
      }
    }
  };
  for (int di = 0; di < P.n_pre_dates; ++di) eval_date(di);
  for (int is = 0; is < P.n_sub; ++is) {
    double z0, z1;
    irc_draw<R, CIR>(rng, ns, is, lpath, gpath, z0, z1);
    irc_step<R, CIR, SCHEME>(P, mp, st, is, z0, z1);
    const int di = __ldg(P.step_date + is);
    if (di >= 0) eval_date(di);
  }
#pragma unroll
  for (int u = 0; u < MCRE_IRC_MAX_UNITS; ++u)
    if (u < P.n_units && P.n_reg > 0) wbuf[((size_t)u * P.n_reg + (P.n_reg - 1)) * n + lpath] = W[u];
}

// Pre-simulation pass B: backward FP32 suffix sums + Gram / right-hand-side moments.
// slot layout: [n_reg][5 + 3*NU] = sum u^0..u^4, then per unit sum u^0..u^2 * Y
template <int NU>
__global__ void __launch_bounds__(256) irc_presim_moments_kernel(IrcDev P, ShardDev sh, const double *xbuf,
                                                                 const double *nbuf, const float *wbuf,
                                                                 double *partial) {
  constexpr int NV = 5 + 3 * NU;
  extern __shared__ double smem[];
  const int nw = blockDim.x >> 5;
  const int n_slots = P.n_reg * NV;
  double *acc = smem;
  double *stage = smem + n_slots;
  const long long n = sh.n_paths;
  const long long n_chunks = (n + sh.chunk - 1) / sh.chunk;
  for (long long chunk = blockIdx.x; chunk < n_chunks; chunk += gridDim.x) {
    for (int i = threadIdx.x; i < n_slots; i += blockDim.x) acc[i] = 0.0;
    __syncthreads();
    int parity = 0;
    for (int it = 0; it < sh.chunk; it += blockDim.x) {
      const long long lpath = chunk * sh.chunk + it + threadIdx.x;
      const bool live = lpath < n;
      const long long p = live ? lpath : 0;
      float S[NU];
#pragma unroll
      for (int u = 0; u < NU; ++u) S[u] = 0.0f;
      for (int k = P.n_reg - 1; k >= 0; --k) {
        const double x = xbuf[(size_t)k * n + p], numeraire = nbuf[(size_t)k * n + p];
        const double uu = (x - __ldg(P.reg_basis + k * 2)) * __ldg(P.reg_basis + k * 2 + 1);
        const double keep = live ? 1.0 : 0.0;
        double vals[NV];
        vals[0] = keep; vals[1] = keep * uu; vals[2] = vals[1] * uu; vals[3] = vals[2] * uu; vals[4] = vals[3] * uu;
#pragma unroll
        for (int u = 0; u < NU; ++u) {
          // total = step_value + tail_value, both float32 (controller.py:349)
          if (u < P.n_units) S[u] = wbuf[((size_t)u * P.n_reg + k) * n + p] + S[u];
          const double Y = numeraire * (double)S[u];  // numeraire.unsqueeze(1) * total_cfs (controller.py:368)
          vals[5 + 3 * u] = keep * Y; vals[6 + 3 * u] = keep * Y * uu; vals[7 + 3 * u] = keep * Y * uu * uu;
        }
        block_accumulate<NV>(vals, acc, k * NV, stage, NV, parity);
      }
    }
    __syncthreads();
    for (int i = threadIdx.x; i < n_slots; i += blockDim.x) partial[(size_t)chunk * n_slots + i] = acc[i];
    __syncthreads();
  }
}

}  // namespace mcre

// =====================================================================================
// Host side of the C ABI
// =====================================================================================
using namespace mcre;

struct mcre_irc_plan {
  IrcDev d;
  DevArray<double> vas, cir, cir_init, chol, step_dt, step_vas, step_cir, float_coef, float_inv_tau, set_fix,
      set_float, set_threshold, expo_coef, expo_basis, cva_coef, unit_fix, unit_float, reg_basis;
  DevArray<int> step_date, date_flags, date_expo, date_metric, date_reg, date_float_off, set_flags, set_lag;
  size_t expo_coef_count = 0;
};

static RngDev make_rng(const mcre_rng *r) {
  RngDev d;
  d.mode = r->mode; d.k0 = (uint32_t)r->seed; d.k1 = (uint32_t)r->stream;
  d.z = r->d_z; d.u = r->d_u; d.n_total = r->n_paths_total;
  return d;
}
static int check_shard(const mcre_shard *s) {
  if (!s || s->n_paths < 0 || s->chunk_paths <= 0 || s->chunk_paths % 256 != 0)
    return fail(-2, "invalid shard: chunk_paths must be a positive multiple of 256%s", "");
  if (s->path_begin % s->chunk_paths != 0) return fail(-2, "invalid shard: path_begin not chunk aligned%s", "");
  return 0;
}

extern "C" int mcre_irc_create(const mcre_irc_desc *c, mcre_irc_plan **out) {
  if (!c || !out) return fail(-1, "null argument%s", "");
  if (c->nt != 0 && c->nt != 4 && c->nt != 8) return fail(-1, "irc: nt must be 0, 4 or 8%s", "");
  if (c->n_sets < 0 || c->n_sets > MCRE_IRC_MAX_SETS) return fail(-1, "irc: n_sets out of range%s", "");
  if (c->n_units < 0 || c->n_units > MCRE_IRC_MAX_UNITS) return fail(-1, "irc: n_units out of range%s", "");
  if (c->scheme != MCRE_SCHEME_EULER && c->scheme != MCRE_SCHEME_ANALYTICAL)
    return fail(-1, "irc: scheme must be EULER or ANALYTICAL%s", "");
  if (c->scheme == MCRE_SCHEME_ANALYTICAL && c->has_cir)
    return fail(-1, "irc: ANALYTICAL is not defined for the Vasicek+CIR++ hybrid (model_config.py:216-221)%s", "");
  mcre_irc_plan *p = new mcre_irc_plan();
  const int w = 1 + c->nt;
  const int n_float = c->date_float_off ? c->date_float_off[c->n_dates] : 0;
  int rc = 0;
#define UP(field, host, count) if (!rc) rc = p->field.upload(host, (size_t)(count))
  UP(vas, c->vas, 4 * w); UP(cir, c->cir, c->has_cir ? 4 * w : 0); UP(cir_init, c->cir_init, c->has_cir ? w : 0);
  UP(chol, c->chol, 4 * w);
  UP(step_dt, c->step_dt, c->n_sub); UP(step_date, c->step_date, c->n_sub);
  UP(step_vas, c->step_vas, (size_t)c->n_sub * 2 * w); UP(step_cir, c->step_cir, c->has_cir ? (size_t)c->n_sub * 2 * w : 0);
  UP(date_flags, c->date_flags, c->n_dates); UP(date_expo, c->date_expo, c->n_dates);
  UP(date_metric, c->date_metric, c->n_dates); UP(date_reg, c->date_reg, c->n_dates);
  UP(date_float_off, c->date_float_off, c->n_dates + 1);
  UP(float_coef, c->float_coef, (size_t)n_float * 2 * w); UP(float_inv_tau, c->float_inv_tau, n_float);
  UP(set_fix, c->set_fix, (size_t)c->n_sets * c->n_dates); UP(set_float, c->set_float, (size_t)c->n_sets * n_float);
  UP(set_threshold, c->set_threshold, c->n_sets); UP(set_flags, c->set_flags, c->n_sets);
  UP(set_lag, c->set_lag, (size_t)c->n_sets * c->n_metric);
  p->expo_coef_count = (size_t)c->n_expo * c->n_sets * 3 * w;
  UP(expo_coef, c->expo_coef, p->expo_coef_count); UP(expo_basis, c->expo_basis, (size_t)c->n_expo * 2);
  UP(cva_coef, c->cva_coef, (size_t)c->n_metric * 2 * w);
  UP(unit_fix, c->unit_fix, (size_t)c->n_units * c->n_dates); UP(unit_float, c->unit_float, (size_t)c->n_units * n_float);
  UP(reg_basis, c->reg_basis, (size_t)c->n_reg * 2);
#undef UP
  if (rc) { mcre_irc_destroy(p); return rc; }
  IrcDev &d = p->d;
  d.nt = c->nt; d.scheme = c->scheme; d.has_cir = c->has_cir; d.cir_det = c->cir_deterministic;
  d.vas_noise = c->vas_noise; d.cir_noise = c->cir_noise;
  d.vas = p->vas.p; d.cir = p->cir.p; d.cir_init = p->cir_init.p; d.chol = p->chol.p;
  d.n_sub = c->n_sub; d.n_dates = c->n_dates; d.n_pre_dates = c->n_pre_dates;
  d.step_dt = p->step_dt.p; d.step_date = p->step_date.p; d.step_vas = p->step_vas.p; d.step_cir = p->step_cir.p;
  d.date_flags = p->date_flags.p; d.date_expo = p->date_expo.p; d.date_metric = p->date_metric.p;
  d.date_reg = p->date_reg.p; d.date_float_off = p->date_float_off.p;
  d.float_coef = p->float_coef.p; d.float_inv_tau = p->float_inv_tau.p; d.n_float = n_float;
  d.n_sets = c->n_sets; d.n_expo = c->n_expo; d.n_metric = c->n_metric; d.acc_flags = c->acc_flags;
  d.set_fix = p->set_fix.p; d.set_float = p->set_float.p; d.set_threshold = p->set_threshold.p;
  d.set_flags = p->set_flags.p; d.set_lag = p->set_lag.p;
  d.expo_coef = p->expo_coef.p; d.expo_basis = p->expo_basis.p; d.cva_coef = p->cva_coef.p; d.lgd = c->lgd;
  d.n_units = c->n_units; d.n_reg = c->n_reg;
  d.unit_fix = p->unit_fix.p; d.unit_float = p->unit_float.p; d.reg_basis = p->reg_basis.p;
  *out = p;
  return 0;
}

extern "C" void mcre_irc_destroy(mcre_irc_plan *p) {
  if (!p) return;
  p->vas.release(); p->cir.release(); p->cir_init.release(); p->chol.release(); p->step_dt.release();
  p->step_vas.release(); p->step_cir.release(); p->float_coef.release(); p->float_inv_tau.release();
  p->set_fix.release(); p->set_float.release(); p->set_threshold.release(); p->expo_coef.release();
  p->expo_basis.release(); p->cva_coef.release(); p->unit_fix.release(); p->unit_float.release();
  p->reg_basis.release(); p->step_date.release(); p->date_flags.release(); p->date_expo.release();
  p->date_metric.release(); p->date_reg.release(); p->date_float_off.release(); p->set_flags.release();
  p->set_lag.release();
  delete p;
}

static int ns_template(int n_sets) { return n_sets <= 1 ? 1 : (n_sets <= 2 ? 2 : 4); }
static int nu_template(int n_units) { return n_units <= 1 ? 1 : (n_units <= 2 ? 2 : 4); }

extern "C" int64_t mcre_irc_main_slots(const mcre_irc_plan *p) {
  return (int64_t)(p->d.n_metric + 1) * ns_template(p->d.n_sets) * (4 + 2 * p->d.nt);
}
extern "C" int64_t mcre_irc_presim_slots(const mcre_irc_plan *p) {
  return (int64_t)p->d.n_reg * (5 + 3 * nu_template(p->d.n_units));
}
extern "C" int64_t mcre_irc_presim_scratch_bytes(const mcre_irc_plan *p, int64_t n_paths) {
  return (int64_t)p->d.n_reg * n_paths * (16 + 4 * (int64_t)p->d.n_units) + 256;
}
extern "C" int64_t mcre_irc_partial_bytes(const mcre_irc_plan *p, int64_t n_paths, int32_t chunk, int presim) {
  int64_t n_chunks = (n_paths + chunk - 1) / chunk;
  return n_chunks * (presim ? mcre_irc_presim_slots(p) : mcre_irc_main_slots(p)) * 8;
}

extern "C" int mcre_irc_set_coefficients(mcre_irc_plan *p, const double *coef, void *stream) {
  if (!p || (!coef && p->expo_coef_count)) return fail(-1, "null argument%s", "");
  if (p->expo_coef_count == 0) return 0;
  MCRE_CUDA(cudaMemcpyAsync(p->d.expo_coef, coef, p->expo_coef_count * sizeof(double), cudaMemcpyHostToDevice,
                            (cudaStream_t)stream));
  return 0;
}

template <int NT, int NS>
static int launch_main(mcre_irc_plan *p, const RngDev &rng, const ShardDev &sh, double *partial, double *spill,
                       double *shift, cudaStream_t st) {
  const IrcDev &d = p->d;
  const int threads = 256, nw = threads / 32;
  const int nvb = NS * (4 + 2 * NT);
  const size_t smem = ((size_t)(d.n_metric + 1) * nvb + 2 * nw * nvb) * sizeof(double);
  const long long n_chunks = (sh.n_paths + sh.chunk - 1) / sh.chunk;
  if (n_chunks == 0) return 0;
#define LAUNCH(CIRV, SCH)                                                                              \
  do {                                                                                                 \
    auto k = irc_main_kernel<NT, NS, CIRV, SCH>;                                                       \
    if (smem > 48 * 1024) MCRE_CUDA(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); \
    int per_sm = 1;                                                                                    \
    MCRE_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k, threads, smem));               \
    if (per_sm < 1) return fail(-3, "irc main kernel does not fit: too many metric dates x tangents%s", ""); \
    long long grid = (long long)sm_count() * per_sm;                                                   \
    if (grid > n_chunks) grid = n_chunks;                                                              \
    ShardDev pilot_sh{0, 1, sh.chunk};                                                                 \
    k<<<1, threads, smem, st>>>(d, rng, pilot_sh, partial, spill, shift, 1);                           \
    MCRE_LAUNCHED();                                                                                   \
    k<<<(unsigned)grid, threads, smem, st>>>(d, rng, sh, partial, spill, shift, 0);                    \
    MCRE_LAUNCHED();                                                                                   \
  } while (0)
  if (d.has_cir) LAUNCH(true, MCRE_SCHEME_EULER);
  else if (d.scheme == MCRE_SCHEME_ANALYTICAL) LAUNCH(false, MCRE_SCHEME_ANALYTICAL);
  else LAUNCH(false, MCRE_SCHEME_EULER);
#undef LAUNCH
  return 0;
}

extern "C" int mcre_irc_mainsim(mcre_irc_plan *p, const mcre_rng *rng, const mcre_shard *shard, double *d_partial,
                                double *d_acc, double *d_shift, double *d_spill, void *stream) {
  if (!p || !rng || !d_partial || !d_acc || !d_shift) return fail(-1, "null argument%s", "");
  int rc = check_shard(shard);
  if (rc) return rc;
  if ((p->d.acc_flags & MCRE_ACC_SPILL) && !d_spill) return fail(-1, "spill requested but d_spill is null%s", "");
  if (rng->mode == MCRE_RNG_INJECT && !rng->d_z) return fail(-1, "inject mode without normals%s", "");
  RngDev r = make_rng(rng);
  ShardDev sh{shard->path_begin, shard->n_paths, shard->chunk_paths};
  cudaStream_t st = (cudaStream_t)stream;
  const int ns = ns_template(p->d.n_sets);
  // tangent builds exist for up to 2 netting sets per launch (register budget); the host
  // splits larger books into groups and replays the same Philox streams.
  if (p->d.nt == 0) {
    rc = ns == 1 ? launch_main<0, 1>(p, r, sh, d_partial, d_spill, d_shift, st)
       : ns == 2 ? launch_main<0, 2>(p, r, sh, d_partial, d_spill, d_shift, st)
                 : launch_main<0, 4>(p, r, sh, d_partial, d_spill, d_shift, st);
  } else if (ns > 2) {
    return fail(-3, "irc: at most 2 netting sets per launch when tangents are on%s", "");
  } else if (p->d.nt == 4) {
    rc = ns == 1 ? launch_main<4, 1>(p, r, sh, d_partial, d_spill, d_shift, st)
                 : launch_main<4, 2>(p, r, sh, d_partial, d_spill, d_shift, st);
  } else {
    rc = ns == 1 ? launch_main<8, 1>(p, r, sh, d_partial, d_spill, d_shift, st)
                 : launch_main<8, 2>(p, r, sh, d_partial, d_spill, d_shift, st);
  }
  if (rc) return rc;
  const long long n_chunks = (sh.n_paths + sh.chunk - 1) / sh.chunk;
  return mcre_tree_reduce(d_partial, n_chunks, mcre_irc_main_slots(p), d_acc, stream);
}

extern "C" int mcre_irc_presim(mcre_irc_plan *p, const mcre_rng *rng, const mcre_shard *shard, void *d_scratch,
                               double *d_partial, double *d_moments, void *stream) {
  if (!p || !rng || !d_scratch || !d_partial || !d_moments) return fail(-1, "null argument%s", "");
  int rc = check_shard(shard);
  if (rc) return rc;
  if (p->d.nt != 0) return fail(-4, "irc presim: tangents through the regression are not implemented%s", "");
  if (rng->mode == MCRE_RNG_INJECT && !rng->d_z) return fail(-1, "inject mode without normals%s", "");
  const IrcDev &d = p->d;
  if (d.n_reg == 0 || shard->n_paths == 0) return 0;
  RngDev r = make_rng(rng);
  ShardDev sh{shard->path_begin, shard->n_paths, shard->chunk_paths};
  cudaStream_t st = (cudaStream_t)stream;
  const long long n = sh.n_paths;
  double *xbuf = (double *)d_scratch;
  double *nbuf = xbuf + (size_t)d.n_reg * n;
  float *wbuf = (float *)(nbuf + (size_t)d.n_reg * n);
  const int threads = 256;
  const unsigned blocks = (unsigned)((n + threads - 1) / threads);
  if (d.has_cir) irc_presim_forward_kernel<true, MCRE_SCHEME_EULER><<<blocks, threads, 0, st>>>(d, r, sh, xbuf, nbuf, wbuf);
  else if (d.scheme == MCRE_SCHEME_ANALYTICAL)
    irc_presim_forward_kernel<false, MCRE_SCHEME_ANALYTICAL><<<blocks, threads, 0, st>>>(d, r, sh, xbuf, nbuf, wbuf);
  else irc_presim_forward_kernel<false, MCRE_SCHEME_EULER><<<blocks, threads, 0, st>>>(d, r, sh, xbuf, nbuf, wbuf);
  MCRE_LAUNCHED();
  const int nu = nu_template(d.n_units);
  const int nv = 5 + 3 * nu, nw = threads / 32;
  const size_t smem = ((size_t)d.n_reg * nv + 2 * nw * nv) * sizeof(double);
  const long long n_chunks = (n + sh.chunk - 1) / sh.chunk;
#define LAUNCHB(NUV)                                                                                     \
  do {                                                                                                   \
    auto k = irc_presim_moments_kernel<NUV>;                                                             \
    if (smem > 48 * 1024) MCRE_CUDA(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); \
    int per_sm = 1;                                                                                      \
    MCRE_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k, threads, smem));                 \
    if (per_sm < 1) return fail(-3, "irc presim kernel does not fit: too many regression dates%s", "");  \
    long long grid = (long long)sm_count() * per_sm;                                                     \
    if (grid > n_chunks) grid = n_chunks;                                                                \
    k<<<(unsigned)grid, threads, smem, st>>>(d, sh, xbuf, nbuf, wbuf, d_partial);                        \
    MCRE_LAUNCHED();                                                                                     \
  } while (0)
  if (nu == 1) LAUNCHB(1); else if (nu == 2) LAUNCHB(2); else LAUNCHB(4);
#undef LAUNCHB
  return mcre_tree_reduce(d_partial, n_chunks, mcre_irc_presim_slots(p), d_moments, stream);
}
