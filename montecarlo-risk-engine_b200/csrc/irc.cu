// Interest-rate / credit family: fused path generation + cashflows + exposure + metrics.
//
// One thread owns one path for its whole life: the Vasicek (and CIR++) state lives in
// registers, normals come from Philox in registers (or from the reference's injected
// draws), and at every simulation date the thread evaluates the netted cashflows, the
// regression-proxy exposure, threshold / MPoR collateral and the metric integrands.
// Only block-reduced sums ever reach HBM.  See include/mcre.h for what each entry point
// replaces in the reference.
#define MCRE_FAST_MATH 1
#include "common.cuh"
#include "philox.cuh"
#include "dual.cuh"
#include "reduce.cuh"
#include "launch.cuh"

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
  const double *step_rec, *date_rec;  // packed records of the main kernel
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
    const R theta_t = T::load(P.step_vas, is * 2 + 0);  // mean level at t1 (constant for Vasicek)
    s.r = s.r + mp.a * (theta_t - s.r) * dt + mp.sigma * sq * wv;
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
//
// Layout choices that came out of the first ncu profile (profiles/r01_irc_main_v0_*):
//  * PP paths per thread, evaluated in lock-step: all plan loads, branches and index
//    arithmetic are shared by the PP paths and their Horner chains interleave;
//  * every per-step / per-date scalar sits in one packed record (mcre_irc_create packs
//    them), read with a handful of wide uniform loads instead of ~45 scalar loads;
//  * MODE 1 ("CVA only": one netting set, no threshold / collateral, no other metric):
//    relu(E_k) S(0,t_k) = relu(poly) exp(-(logB + logB_lambda)) - two exponentials per date.
// =====================================================================================
constexpr int STEP_HDR = 4;   // dt, sqrt(dt), bits(date index), pad
constexpr int DATE_HDR = 6;   // bits(flags|(expo+1)<<32), bits((metric+1)|float_off<<32), bits(float_cnt), pad, shift, scale

__device__ __forceinline__ int lo32(double x) { return __double2loint(x); }
__device__ __forceinline__ int hi32(double x) { return __double2hiint(x); }

template <int NT, int NS, bool CIR, int SCHEME, int PP, int MODE>
__global__ void __launch_bounds__(128, (NT == 0 ? 4 : 1)) irc_main_kernel(IrcDev P, RngDev rng, ShardDev sh,
                                                                         double *partial, double *spill,
                                                                         double *shift, int pilot) {
  typedef typename RealOf<NT>::type R;
  typedef RealTraits<R> T;
  constexpr int W = NT + 1;
  constexpr int NV = 4 + 2 * NT;        // values per (set, date)
  constexpr int NVB = NS * NV;          // values per block_accumulate call
  constexpr int SR = STEP_HDR + 4 * W;  // packed step record stride
  extern __shared__ double smem[];
  const int nw = blockDim.x >> 5;
  const int n_slots = (P.n_metric + 1) * NVB;
  double *acc = smem;                   // [n_slots]
  double *stage = smem + n_slots;       // [2][nw][NVB]
  const long long n_chunks = (sh.n_paths + sh.chunk - 1) / sh.chunk;
  const int DR = (DATE_HDR + 2 * W + 3 * W * P.n_sets + 1) & ~1;  // even: records are read as double2

  IrcParams<R, CIR> mp;
  irc_load_params<R, CIR>(P, mp);
  const int acc_flags = P.acc_flags;
  const bool vas_second = CIR && P.vas_noise == 1, cir_second = CIR && P.cir_noise == 1, cir_det = P.cir_det != 0;
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
    for (int it = 0; it < sh.chunk; it += blockDim.x * PP) {
      long long lpath[PP], gpath[PP];
      bool live[PP];
      NormalStream ns[PP];
      IrcState<R> st[PP];
      R pv[PP][NS], cva[PP][NS], hist[PP][NS][MODE == 1 ? 1 : MCRE_IRC_MAX_LAG];
#pragma unroll
      for (int p = 0; p < PP; ++p) {
        lpath[p] = chunk * sh.chunk + it + p * (int)blockDim.x + threadIdx.x;
        live[p] = lpath[p] < sh.n_paths;
        gpath[p] = sh.path_begin + (live[p] ? lpath[p] : 0);
        ns[p].init(rng, (unsigned long long)gpath[p]);
        st[p].r = mp.r0; st[p].logB = T::zero(); st[p].y = mp.y0; st[p].logBl = T::zero();
#pragma unroll
        for (int s = 0; s < NS; ++s) {
          pv[p][s] = T::zero(); cva[p][s] = T::zero();
#pragma unroll
          for (int l = 0; l < (MODE == 1 ? 1 : MCRE_IRC_MAX_LAG); ++l) hist[p][s][l] = T::zero();
        }
      }

      // ---- date evaluation (cashflows -> exposure -> metrics) ------------------------
      auto eval_date = [&](int di) {
        const double *dr = P.date_rec + (size_t)di * DR;
        const double2 h0 = __ldg((const double2 *)dr), h1 = __ldg((const double2 *)dr + 1),
                      h2 = __ldg((const double2 *)dr + 2);
        const int flags = lo32(h0.x), e = hi32(h0.x) - 1, m = lo32(h0.y) - 1;
        if (!(flags & (MCRE_DATE_HAS_CASHFLOW | MCRE_DATE_HAS_EXPOSURE | MCRE_DATE_HAS_METRIC))) return;
        const double bshift = h2.x, bscale = h2.y;
        const double *dc = dr + DATE_HDR;   // C[w], B[w], coef[set][3][w]
        if (MODE == 1) {
          // CVA-only fast path: contribution at metric dates k < n_metric-1 only
          if (!(flags & MCRE_DATE_HAS_METRIC) || m >= P.n_metric - 1) return;
          const R C = T::load(dc, 0), Bc = T::load(dc, 1);
          const R c0 = T::load(dc, 2), c1 = T::load(dc, 3), c2 = T::load(dc, 4);
#pragma unroll
          for (int p = 0; p < PP; ++p) {
            const R u = (st[p].r - bshift) * bscale;
            const R pos = r_relu(c0 + u * (c1 + u * c2));
            const R ds = r_exp(-(st[p].logB + st[p].logBl));
            const R cond = C * r_exp(-(Bc * st[p].y));
            cva[p][0] = cva[p][0] + pos * ds * (1.0 - cond);
          }
          return;
        }
        R invN[PP];
#pragma unroll
        for (int p = 0; p < PP; ++p) invN[p] = r_exp(-st[p].logB);  // 1 / numeraire (vasicek.py:154-156)
        if ((flags & MCRE_DATE_HAS_CASHFLOW) && (acc_flags & MCRE_ACC_PV)) {
          R cf[PP][NS];
#pragma unroll
          for (int s = 0; s < NS; ++s) {
            const double fx = s < P.n_sets ? __ldg(P.set_fix + (size_t)s * P.n_dates + di) : 0.0;
#pragma unroll
            for (int p = 0; p < PP; ++p) cf[p][s] = T::lift(fx);
          }
          const int j0 = hi32(h0.y), j1 = j0 + lo32(h1.x);
          for (int j = j0; j < j1; ++j) {
            // LIBOR from the bond price at the payment date's own short rate (bond.py:55-66)
            const R alpha = T::load(P.float_coef, j * 2 + 0), B = T::load(P.float_coef, j * 2 + 1);
            const double inv_tau = __ldg(P.float_inv_tau + j);
#pragma unroll
            for (int p = 0; p < PP; ++p) {
              const R libor = (r_exp(B * st[p].r - alpha) - 1.0) * inv_tau;
#pragma unroll
              for (int s = 0; s < NS; ++s)
                if (s < P.n_sets) cf[p][s] = cf[p][s] + libor * __ldg(P.set_float + (size_t)s * P.n_float + j);
            }
          }
#pragma unroll
          for (int p = 0; p < PP; ++p)
#pragma unroll
            for (int s = 0; s < NS; ++s) pv[p][s] = pv[p][s] + cf[p][s] * invN[p];
        }
        if (flags & MCRE_DATE_HAS_EXPOSURE) {
#pragma unroll
          for (int s = 0; s < NS; ++s) {
            R c0 = T::zero(), c1 = T::zero(), c2 = T::zero();
            if (s < P.n_sets) { c0 = T::load(dc, 2 + s * 3); c1 = T::load(dc, 3 + s * 3); c2 = T::load(dc, 4 + s * 3); }
#pragma unroll
            for (int p = 0; p < PP; ++p) {
#pragma unroll
              for (int l = MCRE_IRC_MAX_LAG - 1; l > 0; --l) hist[p][s][l] = hist[p][s][l - 1];
              const R u = (st[p].r - bshift) * bscale;
              hist[p][s][0] = (c0 + u * (c1 + u * c2)) * invN[p];  // continuation / numeraire (controller.py:438-447)
            }
          }
        }
        if (flags & MCRE_DATE_HAS_METRIC) {
          double vals[NVB];
#pragma unroll
          for (int i = 0; i < NVB; ++i) vals[i] = 0.0;
          const bool cva_date = (acc_flags & MCRE_ACC_CVA) && m < P.n_metric - 1;
          R dflt[PP];
#pragma unroll
          for (int p = 0; p < PP; ++p) dflt[p] = T::zero();
          if (CIR && cva_date) {
            // S(0,t_k) = exp(-logB_lambda); S(t_k,t_k+1 | y) = C exp(-B y)   (cirpp.py:298-317)
            const R C = T::load(dc, 0), Bc = T::load(dc, 1);
#pragma unroll
            for (int p = 0; p < PP; ++p) dflt[p] = r_exp(-st[p].logBl) * (1.0 - C * r_exp(-(Bc * st[p].y)));
          }
#pragma unroll
          for (int s = 0; s < NS; ++s) {
            const int lag = ((sflags[s] & 1) && s < P.n_sets) ? __ldg(P.set_lag + (size_t)s * P.n_metric + m) : -1;
            const int sb = m * NVB + s * NV;
            double sh_pos = 0.0, sh_neg = 0.0;
            if (!pilot) { sh_pos = shift[sb + 0]; sh_neg = shift[sb + 2]; }
#pragma unroll
            for (int p = 0; p < PP; ++p) {
              R unsec;
              if (sflags[s] & 1) {
                R delayed = T::zero();
#pragma unroll
                for (int l = 0; l < MCRE_IRC_MAX_LAG; ++l) if (l == lag) delayed = hist[p][s][l];
                unsec = hist[p][s][0] - apply_threshold(delayed, thr[s]);
              } else {
                unsec = apply_threshold(hist[p][s][0], thr[s]);
              }
              const R pos = r_relu(unsec);
              const R neg = -r_relu(-unsec);
              if (cva_date && (sflags[s] & 2)) cva[p][s] = cva[p][s] + pos * dflt[p];
              if (pilot && p == 0 && threadIdx.x == 0) { shift[sb + 0] = val(pos); shift[sb + 2] = val(neg); }
              const double keep = live[p] ? 1.0 : 0.0;
              const double dp = val(pos) - sh_pos, dn = val(neg) - sh_neg;
              vals[s * NV + 0] += keep * dp; vals[s * NV + 1] += keep * dp * dp;
              vals[s * NV + 2] += keep * dn; vals[s * NV + 3] += keep * dn * dn;
#pragma unroll
              for (int k = 0; k < NT; ++k) {
                vals[s * NV + 4 + k] += keep * tan_of(pos, k);
                vals[s * NV + 4 + NT + k] += keep * tan_of(neg, k);
              }
              if ((acc_flags & MCRE_ACC_SPILL) && live[p] && s < P.n_sets)
                spill[((size_t)s * P.n_metric + m) * sh.n_paths + lpath[p]] = val(unsec);
            }
          }
          if ((acc_flags & (MCRE_ACC_POS | MCRE_ACC_NEG)) && !pilot)
            block_accumulate<NVB>(vals, acc, m * NVB, stage, NVB, parity);
        }
      };

      for (int di = 0; di < P.n_pre_dates; ++di) eval_date(di);
      for (int is = 0; is < P.n_sub; ++is) {
        const double *sr = P.step_rec + (size_t)is * SR;
        const double2 g0 = __ldg((const double2 *)sr), g1 = __ldg((const double2 *)sr + 1);
        const double dt = g0.x, sq = g0.y;
        const int di = lo32(g1.x);
        const R sv0 = T::load(sr + STEP_HDR, 0), sv1 = T::load(sr + STEP_HDR, 1);
        R sc0 = T::zero(), sc1 = T::zero();
        if (CIR) { sc0 = T::load(sr + STEP_HDR, 2); sc1 = T::load(sr + STEP_HDR, 3); }
#pragma unroll
        for (int p = 0; p < PP; ++p) {
          double z0, z1;
          irc_draw<R, CIR>(rng, ns[p], is, lpath[p], gpath[p], z0, z1);
          // correlated noise z @ L^T (model.py:46-48)
          const R w0 = mp.L00 * z0;
          R w1 = T::zero();
          if (CIR) w1 = mp.L10 * z0 + mp.L11 * z1;
          IrcState<R> &s = st[p];
          s.logB = s.logB + s.r * dt;   // left Riemann sum with the pre-step rate (vasicek.py:80,107)
          if (SCHEME == MCRE_SCHEME_ANALYTICAL) {
            // exact OU transition; the 1x1 Cholesky factor of the step covariance is sv1 (vasicek.py:52-86)
            s.r = mp.theta + (s.r - mp.theta) * sv0 + sv1 * z0;
          } else {
            const R wv = vas_second ? w1 : w0;
            s.r = s.r + mp.a * (sv0 - s.r) * dt + mp.sigma * sq * wv;
          }
          if (CIR) {
            const R wc = cir_second ? w1 : w0;
            if (cir_det) {                    // cirpp.py:155-172
              s.logBl = s.logBl + sc0 * dt;
              s.y = sc1;
            } else {                          // full-truncation Euler, cirpp.py:174-198
              const R yn = s.y + mp.kappa * (mp.ctheta - s.y) * dt + mp.csigma * r_sqrt(r_relu(s.y)) * sq * wc;
              s.logBl = s.logBl + (s.y + sc0) * dt;
              s.y = r_max(yn, 1e-12);
            }
          }
        }
        if (di >= 0) eval_date(di);
      }
      // ---- per-path totals ------------------------------------------------------------
      {
        double vals[NVB];
#pragma unroll
        for (int i = 0; i < NVB; ++i) vals[i] = 0.0;
#pragma unroll
        for (int s = 0; s < NS; ++s) {
          const int sb = P.n_metric * NVB + s * NV;
          double sh_pv = 0.0, sh_cva = 0.0;
          if (!pilot) { sh_pv = shift[sb + 0]; sh_cva = shift[sb + 2]; }
#pragma unroll
          for (int p = 0; p < PP; ++p) {
            const R c = cva[p][s] * P.lgd;
            if (pilot && p == 0 && threadIdx.x == 0) { shift[sb + 0] = val(pv[p][s]); shift[sb + 2] = val(c); }
            const double keep = live[p] ? 1.0 : 0.0;
            const double dp = val(pv[p][s]) - sh_pv, dcv = val(c) - sh_cva;
            vals[s * NV + 0] += keep * dp; vals[s * NV + 1] += keep * dp * dp;
            vals[s * NV + 2] += keep * dcv; vals[s * NV + 3] += keep * dcv * dcv;
#pragma unroll
            for (int k = 0; k < NT; ++k) {
              vals[s * NV + 4 + k] += keep * tan_of(pv[p][s], k);
              vals[s * NV + 4 + NT + k] += keep * tan_of(c, k);
            }
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
  DevArray<double> step_rec, date_rec;
  std::vector<double> h_date_rec;   // host copy: coefficients are patched in after the regression solve
  std::vector<int> h_date_expo;
  int date_stride = 0;
  size_t expo_coef_count = 0;
  bool cva_only = false;
};

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
  {
    // packed per-step and per-date records of the main kernel (layout: STEP_HDR / DATE_HDR above)
    const int SR = STEP_HDR + 4 * w, DR = (DATE_HDR + 2 * w + 3 * w * c->n_sets + 1) & ~1;
    std::vector<double> srec((size_t)c->n_sub * SR + 2, 0.0);
    for (int s = 0; s < c->n_sub; ++s) {
      double *r = srec.data() + (size_t)s * SR;
      r[0] = c->step_dt[s]; r[1] = sqrt(c->step_dt[s]); r[2] = pack2(c->step_date[s], 0);
      for (int k = 0; k < 2 * w; ++k) r[STEP_HDR + k] = c->step_vas[(size_t)s * 2 * w + k];
      if (c->has_cir) for (int k = 0; k < 2 * w; ++k) r[STEP_HDR + 2 * w + k] = c->step_cir[(size_t)s * 2 * w + k];
    }
    p->h_date_rec.assign((size_t)c->n_dates * DR + 2, 0.0);
    p->h_date_expo.assign(c->date_expo, c->date_expo + c->n_dates);
    p->date_stride = DR;
    for (int di = 0; di < c->n_dates; ++di) {
      double *r = p->h_date_rec.data() + (size_t)di * DR;
      const int e = c->date_expo[di], m = c->date_metric[di];
      r[0] = pack2(c->date_flags[di], e + 1);
      r[1] = pack2(m + 1, c->date_float_off[di]);
      r[2] = pack2(c->date_float_off[di + 1] - c->date_float_off[di], 0);
      if (e >= 0) { r[4] = c->expo_basis[e * 2]; r[5] = c->expo_basis[e * 2 + 1]; }
      if (m >= 0 && c->cva_coef) for (int k = 0; k < 2 * w; ++k) r[DATE_HDR + k] = c->cva_coef[(size_t)m * 2 * w + k];
      if (e >= 0 && c->expo_coef)
        for (int k = 0; k < 3 * w * c->n_sets; ++k) r[DATE_HDR + 2 * w + k] = c->expo_coef[(size_t)e * 3 * w * c->n_sets + k];
    }
    UP(step_rec, srec.data(), srec.size());
    UP(date_rec, p->h_date_rec.data(), p->h_date_rec.size());
    // "CVA only" fast mode: one set, CVA the only accumulator, no threshold / collateral
    p->cva_only = c->has_cir && c->nt == 0 && c->n_sets == 1 && c->acc_flags == MCRE_ACC_CVA &&
                  c->set_threshold[0] == 0.0 && (c->set_flags[0] & 1) == 0 && (c->set_flags[0] & 2) != 0;
  }
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
  d.step_rec = p->step_rec.p; d.date_rec = p->date_rec.p;
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
  p->set_lag.release(); p->step_rec.release(); p->date_rec.release();
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
  const int w = 1 + p->d.nt, per_date = 3 * w * p->d.n_sets, DR = p->date_stride;
  for (int di = 0; di < p->d.n_dates; ++di) {
    const int e = p->h_date_expo[di];
    if (e < 0) continue;
    double *r = p->h_date_rec.data() + (size_t)di * DR + DATE_HDR + 2 * w;
    for (int k = 0; k < per_date; ++k) r[k] = coef[(size_t)e * per_date + k];
  }
  MCRE_CUDA(cudaMemcpyAsync((void *)p->d.date_rec, p->h_date_rec.data(), p->h_date_rec.size() * sizeof(double),
                            cudaMemcpyHostToDevice, (cudaStream_t)stream));
  MCRE_CUDA(cudaStreamSynchronize((cudaStream_t)stream));  // host staging buffers may be reused by the caller
  return 0;
}

template <int NT, int NS>
static int launch_main(mcre_irc_plan *p, const RngDev &rng, const ShardDev &sh, double *partial, double *spill,
                       double *shift, cudaStream_t st) {
  const IrcDev &d = p->d;
  constexpr int PP = NT == 0 ? 2 : 1;   // paths per thread
  const int threads = 128, nw = threads / 32;
  const int nvb = NS * (4 + 2 * NT);
  const size_t smem = ((size_t)(d.n_metric + 1) * nvb + 2 * nw * nvb) * sizeof(double);
  const long long n_chunks = (sh.n_paths + sh.chunk - 1) / sh.chunk;
  if (n_chunks == 0) return 0;
#define LAUNCH(CIRV, SCH, MODEV)                                                                       \
  do {                                                                                                 \
    auto k = irc_main_kernel<NT, NS, CIRV, SCH, PP, MODEV>;                                            \
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
  if (d.has_cir) {
    if (NT == 0 && NS == 1 && p->cva_only) LAUNCH(true, MCRE_SCHEME_EULER, (NT == 0 && NS == 1 ? 1 : 0));
    else LAUNCH(true, MCRE_SCHEME_EULER, 0);
  } else if (d.scheme == MCRE_SCHEME_ANALYTICAL) LAUNCH(false, MCRE_SCHEME_ANALYTICAL, 0);
  else LAUNCH(false, MCRE_SCHEME_EULER, 0);
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
