// Vasicek + CIR++ "CVA only" kernel: the headline path (BASELINE config 3: wrong-way-risk CVA of a payer swap,
// one netting set, no threshold / collateral, CVA the only metric, value-only build).
//
// What it replaces in the reference, per path and sub-step: MonteCarloEngine.generate_paths (engine.py:27-123) for
// ModelConfig([Vasicek, CIR++]) under EULER (vasicek.py:88-112, cirpp.py:174-198, model_config.py:223-276), the
// request resolution of numeraire / spot / survival probabilities (request_interface.py:115-130), the regression-proxy
// exposure (controller.py:438-447) and CVAMetric.evaluate (cva_metric.py:62-100).
//
// Round-2 rewrite of the MODE 1 branch of irc_main_kernel.  The probe tools/probes/issue_mix.cu showed that on this
// GPU an FP64 instruction occupies the issue port of its scheduler for ~2.2 cycles and integer instructions do not
// issue in its shadow, so the kernel is bound by (2.2 x FP64 + 1 x other) instructions per path-step.  Everything
// below is about removing instructions (84 FP64 + 96 other per path-step in round 1):
//  * every per-step scalar that does not depend on the path is folded on the host into one packed record per
//    sub-step (mcre_irc_plan::build_cva_records):
//      r'  = r (1 - a dt) + a theta_t dt + (sigma sqrt(dt) L_v.) z            2-3 FMA   (was 4-5)
//      y'  = y (1 - kappa dt) + kappa theta dt + sqrt(y) (sigma_c sqrt(dt) L_c.) z
//      A  -= (r + y) dt    the only integral the CVA integrand needs: logB + logB_lambda = -A + sum psi dt, and the
//                          deterministic shift integral sum psi dt is folded into the exposure coefficients
//  * the integrand relu(E_k) S(0,t_k) (1 - S(t_k,t_k+1 | y)) / N_k becomes
//      (E' + |E'|) exp(A) H(y),   E' = 1/2 exp(-sum psi dt) (c0 + c1 r + c2 r^2),
//      H(y) = 1 - C exp(-B y) = (1 - C) + C B y - C B^2 y^2 / 2 ... to y^5 while B y <= 2^-9 (warp vote on the high
//      word of y; otherwise the table exponential): one exponential per date instead of two;
//  * max(y', 1e-12) and the validity tests are integer compares on the high words + one warp vote (an FP64 max is a
//    DSETP and two selects); relu(x) = (x + |x|) / 2 is one DADD;
//  * Box-Muller on the shortened elementary functions of fastmath.cuh (31 FP64 instructions per normal pair);
//  * the unit of work and of deterministic summation is one PASS (128 threads x PP paths = 512 paths): per-path CVA
//    totals are block-reduced once per pass into partial[pass], passes are handed out by an atomic counter (a whole
//    4096-path chunk per hand-out left 1.15 chunks per block on a strong-scaled 8-GPU pass: 58 % of the SMs busy);
//  * the pilot launch is gone: block 0 simulates global path 0 first and publishes its value (the common shift c of
//    sum(x - c), sum((x - c)^2)); the other blocks accumulate against the first path of their own chunk and convert
//    to c when they store the chunk partial (exact algebra, fixed order, so results stay independent of timing and
//    of the number of GPUs); chunks are handed out by an atomic counter, so block 0's late start costs nothing.
#include "irc_main.cuh"

#ifndef MCRE_CVA_PP
#define MCRE_CVA_PP 4
#endif
#ifndef MCRE_CVA_MINB
#define MCRE_CVA_MINB 3
#endif

namespace mcre {

constexpr int CVA_EV_KV1 = 1;      // the short rate also loads on the second normal
constexpr int CVA_EV_KC1 = 2;      // the intensity also loads on the second normal
constexpr int CVA_EV_DATE = 4;     // a metric date k < n_metric - 1 follows the step: CVA contribution
constexpr int CVA_EV_NOSTEP = 8;   // date at the calibration date: no step, no draw
constexpr int HI_1E_12 = 0x3D719799;   // high word of 1e-12

struct CvaDev {
  const double *rec;   // [n_events][CVA_REC]
  int n_pre, n_sub;    // events: n_pre dates at the calibration date (no step), then one per sub-step
  double r0, y0, lgd;
  unsigned *sync;      // [0]: next chunk, [1]: pilot value published
  // Philox round keys (key + i * Weyl constant): kernel parameters sit in the constant bank and feed the
  // round's LOP3 directly - no key-schedule instructions in the loop
  uint32_t rk0[10], rk1[10];
  // USTEP builds: all sub-steps share their scalars (constant dt, constant mean levels): parameters, not loads
  double st[9];
  int st_flags;
};

// 1: the Philox rounds of draw block s+1 are issued inside event s (software pipeline, draw words ping-pong between
// two register sets).  Measured (profiles/r02_cva_kernel.md): ptxas still emits the rounds as one cluster next to
// long DFMA runs, 2 % slower than 0 because of the 16 extra live registers - off.
#ifndef MCRE_CVA_PF
#define MCRE_CVA_PF 0
#endif

// Ten Philox rounds for the draw block `block` of PP paths in lock-step (the paths of one pass share the high word
// of their global id: a pass never leaves its 4096-aligned chunk).
template <int PP>
__device__ __forceinline__ void cva_philox(const CvaDev &P, const uint32_t (&plo)[PP], uint32_t phi, uint32_t block,
                                           uint32_t (&c0)[PP], uint32_t (&c1)[PP], uint32_t (&c2)[PP],
                                           uint32_t (&c3)[PP]) {
  MCRE_VP { c0[p] = plo[p]; c1[p] = phi; c2[p] = block; c3[p] = 0u; }
#pragma unroll
  for (int i = 0; i < 10; ++i) {
    MCRE_VP Philox::round(c0[p], c1[p], c2[p], c3[p], P.rk0[i], P.rk1[i]);
  }
}

template <int PP, bool USTEP, bool PF>
__global__ void __launch_bounds__(128, MCRE_CVA_MINB) irc_cva_kernel(const __grid_constant__ CvaDev P, RngDev rng,
                                                                      ShardDev sh, double *__restrict__ partial,
                                                                      double *shift_tail) {
  fm_tables_init();
  __shared__ double s_stage[2][4];
  __shared__ double s_first;
  __shared__ long long s_chunk;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const long long n_chunks = (sh.n_paths + sh.chunk - 1) / sh.chunk;
  const bool inject = rng.mode == MCRE_RNG_INJECT;
  bool pilot = blockIdx.x == 0;
  bool have_shift = false;
  double gshift = 0.0;

  static_assert((PP & (PP - 1)) == 0, "a pass (128 x PP paths) must divide the 256-aligned chunk size");
  const int unit = 128 * PP;                                   // paths per pass
  const long long n_units = (sh.n_paths + unit - 1) / unit;
  while (true) {
    long long u = 0;
    if (!pilot) {
      if (tid == 0) s_chunk = (long long)atomicAdd(P.sync, 1u);
      __syncthreads();
      u = s_chunk;
      __syncthreads();
      if (u >= n_units) break;
    }
    double s1 = 0.0, s2 = 0.0, first = 0.0;
    {
      // global id of path p of this thread: g0 + 128 p (the pilot pass simulates global path 0 in every lane);
      // lanes past the end of the shard simulate ids nobody owns and are masked out of the sums
      const long long l0 = u * unit + tid;
      const long long g0 = pilot ? 0 : sh.path_begin + l0;
      const int stride = pilot ? 0 : 128;
      unsigned live = 0u;
      uint32_t plo[PP];
      MCRE_VP {
        if (!pilot && l0 + p * 128 < sh.n_paths) live |= 1u << p;
        plo[p] = (uint32_t)(g0 + p * stride);
      }
      const uint32_t phi = (uint32_t)((unsigned long long)g0 >> 32);
      double r[PP], y[PP], A[PP], cva[PP];
      MCRE_VP { r[p] = P.r0; y[p] = P.y0; A[p] = 0.0; cva[p] = 0.0; }

      // ---- CVA contribution of the metric date that closes event `ev` ---------------------------------------
      auto date_part = [&](const double *rec) {
        const double2 *rd = (const double2 *)(rec + 10);
        const double2 h0 = __ldg(rd), h1 = __ldg(rd + 1), h2 = __ldg(rd + 2), h3 = __ldg(rd + 3), h4 = __ldg(rd + 4);
        const double c0 = h0.x, c1 = h0.y, c2 = h1.x, d0 = h1.y, d1 = h2.x, d2 = h2.y, d3 = h3.x, d4 = h3.y,
                     d5 = h4.x;
        const int thr = __double2loint(h4.y);
        double e[PP], ea[PP], h[PP];
        MCRE_VP e[p] = fma(r[p], c2, c1);
        MCRE_VP e[p] = fma(r[p], e[p], c0);
        bool small = true;
        MCRE_VP small = small && (__double2hiint(y[p]) < thr);
        fm_exp_tv<PP>(A, ea);
        MCRE_VP e[p] = e[p] + fabs(e[p]);          // relu of the exposure proxy (the 1/2 sits in c0..c2)
        if (__all_sync(0xffffffffu, small)) {
          MCRE_VP h[p] = fma(y[p], d5, d4);
          MCRE_VP h[p] = fma(h[p], y[p], d3);
          MCRE_VP h[p] = fma(h[p], y[p], d2);
          MCRE_VP h[p] = fma(h[p], y[p], d1);
          MCRE_VP h[p] = fma(h[p], y[p], d0);
        } else {
          __syncwarp();                              // (a real branch, see the clamp above)
          const double2 h5 = __ldg(rd + 5);
          double xb[PP], eb[PP];
          MCRE_VP xb[p] = h5.y * y[p];             // -B y
          fm_exp_tv<PP>(xb, eb);
          // 1 - C exp(-B y)   (cirpp.py:298-317); volatile: not to be merged with the polynomial branch
          MCRE_VP asm volatile("fma.rn.f64 %0, %1, %2, %3;" : "=d"(h[p]) : "d"(-h5.x), "d"(eb[p]), "d"(1.0));
        }
        MCRE_VP e[p] = e[p] * ea[p];
        MCRE_VP cva[p] = fma(e[p], h[p], cva[p]);
      };

      // ---- one sub-step: draws (words `w` were produced one event earlier when PF), SDE step, date ----------
      auto step_event = [&](int is, uint32_t (&w0)[PP], uint32_t (&w1)[PP], uint32_t (&w2)[PP], uint32_t (&w3)[PP],
                            uint32_t (&n0)[PP], uint32_t (&n1)[PP], uint32_t (&n2)[PP], uint32_t (&n3)[PP]) {
        const double *rec = P.rec + (size_t)(P.n_pre + is) * CVA_REC;
        double ndt, om_v, c_v, kv0, kv1, om_c, c_c, kc0, kc1;
        int flags;
        if (USTEP) {
          ndt = P.st[0]; om_v = P.st[1]; c_v = P.st[2]; kv0 = P.st[3]; kv1 = P.st[4]; om_c = P.st[5]; c_c = P.st[6];
          kc0 = P.st[7]; kc1 = P.st[8];
          flags = P.st_flags | (__double2loint(__ldg(rec + 9)) & CVA_EV_DATE);
        } else {
          const double2 *rs = (const double2 *)rec;
          const double2 g0_ = __ldg(rs), g1 = __ldg(rs + 1), g2 = __ldg(rs + 2), g3 = __ldg(rs + 3), g4 = __ldg(rs + 4);
          ndt = g0_.x; om_v = g0_.y; c_v = g1.x; kv0 = g1.y; kv1 = g2.x; om_c = g2.y; c_c = g3.x; kc0 = g3.y; kc1 = g4.x;
          flags = __double2loint(g4.y);
        }
        double z0[PP], z1[PP];
        if (inject) {
          MCRE_VP {
            long long g = g0 + p * stride;
            if (!((live >> p) & 1u) && !pilot) g = sh.path_begin;
            const double *zp = rng.z + ((size_t)is * rng.n_total + g) * 2;
            z0[p] = zp[0]; z1[p] = zp[1];
          }
        } else {
          if (PF) {
            // the integer rounds of the NEXT draw block are independent of everything below: issued here so that
            // they interleave with the FP64 work of this event inside one warp
            cva_philox<PP>(P, plo, phi, (uint32_t)is + 1u, n0, n1, n2, n3);
          } else {
            cva_philox<PP>(P, plo, phi, (uint32_t)is, w0, w1, w2, w3);
          }
          NormalStreamV<PP>::box_muller(w0, w1, w2, w3, z0, z1);
        }
        // integrals with the pre-step state (left Riemann sums, vasicek.py:80,107, cirpp.py:196-197)
        MCRE_VP A[p] = fma(r[p], ndt, A[p]);
        MCRE_VP A[p] = fma(y[p], ndt, A[p]);
        double sy[PP], wn[PP], yn[PP];
        fm_sqrt_posv<PP>(y, sy);                  // y >= 1e-12 after every step and y0 > 0
        MCRE_VP r[p] = fma(r[p], om_v, c_v);
        MCRE_VP r[p] = fma(kv0, z0[p], r[p]);
        if (flags & CVA_EV_KV1) { MCRE_VP r[p] = fma(kv1, z1[p], r[p]); }
        MCRE_VP wn[p] = kc0 * z0[p];
        if (flags & CVA_EV_KC1) { MCRE_VP wn[p] = fma(kc1, z1[p], wn[p]); }
        MCRE_VP yn[p] = fma(y[p], om_c, c_c);
        MCRE_VP yn[p] = fma(sy[p], wn[p], yn[p]);
        // y' = max(yn, 1e-12) (cirpp.py:198): high word above that of 1e-12 <=> certainly larger
        bool above = true;
        MCRE_VP above = above && (__double2hiint(yn[p]) > HI_1E_12);
        if (__all_sync(0xffffffffu, above)) {
          MCRE_VP y[p] = yn[p];
        } else {
          // rare.  (__syncwarp + volatile: a real branch - the compiler otherwise predicates this side into
          // the instruction stream of every step, or folds both sides into an unconditional FP64 max)
          __syncwarp();
          MCRE_VP asm volatile("max.f64 %0, %1, %2;" : "=d"(y[p]) : "d"(yn[p]), "d"(1e-12));
        }
        if (flags & CVA_EV_DATE) date_part(rec);
      };

      for (int ev = 0; ev < P.n_pre; ++ev) {
        const double *rec = P.rec + (size_t)ev * CVA_REC;
        if (__double2loint(__ldg(rec + 9)) & CVA_EV_DATE) date_part(rec);
      }
      uint32_t qa0[PP], qa1[PP], qa2[PP], qa3[PP], qb0[PP], qb1[PP], qb2[PP], qb3[PP];
      if (PF && !inject) cva_philox<PP>(P, plo, phi, 0u, qa0, qa1, qa2, qa3);
      int is = 0;
      // two events per trip: the draw words ping-pong between two register sets (no copies)
#pragma unroll 1
      for (; is + 1 < P.n_sub; is += 2) {
        step_event(is, qa0, qa1, qa2, qa3, qb0, qb1, qb2, qb3);
        step_event(is + 1, qb0, qb1, qb2, qb3, qa0, qa1, qa2, qa3);
      }
      if (is < P.n_sub) step_event(is, qa0, qa1, qa2, qa3, qb0, qb1, qb2, qb3);

      // ---- per-path totals ----------------------------------------------------------------------
      if (pilot) {
        if (tid == 0) {
          shift_tail[0] = 0.0; shift_tail[1] = 0.0; shift_tail[2] = cva[0] * P.lgd; shift_tail[3] = 0.0;
          __threadfence();
          atomicExch(P.sync + 1, 1u);
        }
        pilot = false;
        continue;
      }
      if (tid == 0) s_first = cva[0] * P.lgd;
      __syncthreads();
      first = s_first;
      MCRE_VP {
        const double d = ((live >> p) & 1u) ? fma(cva[p], P.lgd, -first) : 0.0;
        s1 += d;
        s2 = fma(d, d, s2);
      }
    }
    s1 = warp_sum(s1);
    s2 = warp_sum(s2);
    if (lane == 0) { s_stage[0][warp] = s1; s_stage[1][warp] = s2; }
    __syncthreads();
    if (tid == 0) {
      double t1 = 0.0, t2 = 0.0;
      for (int w = 0; w < 4; ++w) { t1 += s_stage[0][w]; t2 += s_stage[1][w]; }
      if (!have_shift) {
        while (atomicAdd(P.sync + 1, 0u) == 0u) __nanosleep(100);
        __threadfence();
        gshift = *(volatile double *)(shift_tail + 2);
        have_shift = true;
      }
      // sums against the pass's first path -> sums against the common shift (global path 0):
      //   sum(x - c) = sum(x - f) + n (f - c),  sum((x - c)^2) = sum((x - f)^2) + (f - c) (2 sum(x - f) + n (f - c))
      const long long left = sh.n_paths - u * unit;
      const double n = (double)(left < unit ? left : unit);
      const double dl = first - gshift;
      double *out = partial + (size_t)u * 4;
      out[0] = 0.0; out[1] = 0.0;
      out[2] = fma(n, dl, t1);
      out[3] = fma(dl, fma(n, dl, 2.0 * t1), t2);
    }
    __syncthreads();
  }
}

}  // namespace mcre

using namespace mcre;

// Packs the event records from the plan's host copies: everything that does not depend on the regression (called
// from mcre_irc_create, the records travel with the plan arena).  Slots 10..12 (the exposure polynomial) are filled
// later, on the host (irc_cva_apply_coefficients) or on the device (irc_patch_coefficients_kernel, irc_solve.cu):
// slot 22 keeps g_k = 1/2 exp(-sum psi dt), slot 23 the date the event closes (+1).
void irc_cva_fill_static(mcre_irc_plan *p) {
  const CvaHost &h = p->cva;
  const int n_events = h.n_pre_dates + h.n_sub;
  std::vector<double> &rec = p->cva_rec_host;
  rec.assign((size_t)n_events * CVA_REC + 2, 0.0);
  const double sigma = h.vas[1], a = h.vas[3];
  const double kappa = h.cir[0], ctheta = h.cir[1], csigma = h.cir[2];
  const double *Lv = h.chol + 2 * h.vas_noise, *Lc = h.chol + 2 * h.cir_noise;   // rows of the lower Cholesky factor
  const int DR = p->date_stride;
  double psi_int = 0.0;   // sum psi(t1) dt over the sub-steps so far: the deterministic part of logB_lambda
  auto date_part = [&](double *r, int di) {
    if (di < 0) return;
    const int m = h.date_metric[di];
    if (!(h.date_flags[di] & MCRE_DATE_HAS_METRIC) || m < 0 || m >= h.n_metric - 1) return;
    const double *dr = p->h_date_rec.data() + (size_t)di * DR + DATE_HDR;   // C, B of the date
    const double C = dr[0], Bc = dr[1];
    double term = -C;           // d_j = -C (-B)^j / j!
    r[13] = 1.0 - C;
    for (int j = 1; j <= 5; ++j) { term *= -Bc / (double)j; r[13 + j] = term; }
    int thr = 0;                // high word below which B y <= 2^-9 for certain
    if (Bc > 0.0) {
      const double ymax = 0.001953125 / Bc;
      long long bits; memcpy(&bits, &ymax, 8);
      thr = (int)(bits >> 32);
    }
    r[19] = pack2(thr, 0);
    r[20] = C; r[21] = -Bc;
    r[22] = 0.5 * exp(-psi_int);
    r[23] = pack2(di + 1, 0);
    long long fb; memcpy(&fb, &r[9], 8);
    r[9] = pack2((int)(fb & 0xffffffff) | CVA_EV_DATE, 0);
  };
  for (int ev = 0; ev < n_events; ++ev) {
    double *r = rec.data() + (size_t)ev * CVA_REC;
    if (ev < h.n_pre_dates) {
      r[9] = pack2(CVA_EV_NOSTEP, 0);
      date_part(r, ev);
      continue;
    }
    const int s = ev - h.n_pre_dates;
    const double dt = h.step_dt[s], sq = sqrt(dt);
    int flags = 0;
    r[0] = -dt;
    r[1] = 1.0 - a * dt; r[2] = a * h.step_theta[s] * dt;
    r[3] = sigma * sq * Lv[0]; r[4] = sigma * sq * Lv[1];
    r[5] = 1.0 - kappa * dt; r[6] = kappa * ctheta * dt;
    r[7] = csigma * sq * Lc[0]; r[8] = csigma * sq * Lc[1];
    if (r[4] != 0.0) flags |= CVA_EV_KV1;
    if (r[8] != 0.0) flags |= CVA_EV_KC1;
    r[9] = pack2(flags, 0);
    psi_int += h.step_psi[s] * dt;
    date_part(r, h.step_date[s]);
  }
}

// Host path of the coefficients (mcre_irc_set_coefficients): the raw-basis polynomial of h_date_rec, scaled by g_k.
int irc_cva_apply_coefficients(mcre_irc_plan *p, cudaStream_t st) {
  const CvaHost &h = p->cva;
  const int n_events = h.n_pre_dates + h.n_sub;
  std::vector<double> &rec = p->cva_rec_host;
  const int DR = p->date_stride;
  for (int ev = 0; ev < n_events; ++ev) {
    double *r = rec.data() + (size_t)ev * CVA_REC;
    long long db; memcpy(&db, &r[23], 8);
    const int di = (int)(db & 0xffffffff) - 1;
    if (di < 0) continue;
    const double *dr = p->h_date_rec.data() + (size_t)di * DR + DATE_HDR;   // C, B, c0, c1, c2 (raw basis)
    r[10] = r[22] * dr[2]; r[11] = r[22] * dr[3]; r[12] = r[22] * dr[4];
  }
  MCRE_CUDA(cudaMemcpyAsync(p->cva_rec_dev, rec.data(), (size_t)n_events * CVA_REC * sizeof(double),
                            cudaMemcpyHostToDevice, st));
  MCRE_H2D((size_t)n_events * CVA_REC * sizeof(double));
  return 0;
}

long long irc_cva_units(long long n_paths) { return (n_paths + 128 * MCRE_CVA_PP - 1) / (128 * MCRE_CVA_PP); }

int irc_cva_launch(mcre_irc_plan *p, const RngDev &rng, const ShardDev &sh, double *partial, double *shift,
                   cudaStream_t st) {
  const CvaHost &h = p->cva;
  CvaDev d;
  d.rec = p->cva_rec_dev; d.n_pre = h.n_pre_dates; d.n_sub = h.n_sub;
  d.r0 = h.vas[0]; d.y0 = h.y0; d.lgd = h.lgd;
  d.sync = p->cva_sync_dev;
  for (int i = 0; i < 10; ++i) { d.rk0[i] = rng.k0 + (uint32_t)i * 0x9E3779B9u; d.rk1[i] = rng.k1 + (uint32_t)i * 0xBB67AE85u; }
  // all sub-steps alike (constant dt and mean level)?  Then their scalars travel as kernel parameters.
  const std::vector<double> &rec = p->cva_rec_host;
  bool ustep = h.n_sub > 0;
  const double *first = rec.data() + (size_t)h.n_pre_dates * CVA_REC;
  long long f0 = 0;
  if (ustep) memcpy(&f0, first + 9, 8);
  for (int s = 1; s < h.n_sub && ustep; ++s) {
    const double *r = first + (size_t)s * CVA_REC;
    long long fs; memcpy(&fs, r + 9, 8);
    ustep = memcmp(r, first, 9 * sizeof(double)) == 0 && ((fs ^ f0) & ~(long long)CVA_EV_DATE) == 0;
  }
  for (int k = 0; k < 9; ++k) d.st[k] = ustep ? first[k] : 0.0;
  d.st_flags = ustep ? (int)(f0 & ~(long long)CVA_EV_DATE) : 0;
  const long long n_chunks = (sh.n_paths + 128 * MCRE_CVA_PP - 1) / (128 * MCRE_CVA_PP);   // passes
  auto k = ustep ? irc_cva_kernel<MCRE_CVA_PP, true, MCRE_CVA_PF != 0> : irc_cva_kernel<MCRE_CVA_PP, false, MCRE_CVA_PF != 0>;
  int per_sm = 1;
  MCRE_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k, 128, 0));
  if (per_sm < 1) return fail(-3, "irc cva kernel does not fit%s", "");
  long long grid = (long long)sm_count() * per_sm;
  if (grid > n_chunks) grid = n_chunks;
  if (grid < 1) grid = 1;      // a rank without paths still publishes the common shift (global path 0)
  MCRE_CUDA(cudaMemsetAsync(d.sync, 0, 2 * sizeof(unsigned), st));
  // shift layout [n_metric + 1][4]: only the tail row (pv, -, cva, -) is used by this mode
  k<<<(unsigned)grid, 128, 0, st>>>(d, rng, sh, partial, shift + (size_t)p->d.n_metric * 4);
  MCRE_LAUNCHED();
  return 0;
}
