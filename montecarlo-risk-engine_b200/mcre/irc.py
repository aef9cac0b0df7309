"""Plan compiler + driver for the interest-rate / credit family (IRC).

Lowers (Vasicek [+ CIR++] model, bonds / swaps, netting sets, metrics, timelines) into
the flat tables of ``mcre_irc_desc`` (include/mcre.h) and drives the two fused CUDA
passes: pre-simulation moments -> host solve of the 3x3 normal equations -> main
simulation.  What it takes the place of in the reference:
  request collection        src/request_interface/request_interface.py:22-97
  _perform_regression       src/controller/controller.py:272-383
  evaluate_products         src/controller/controller.py:565-661
"""
from __future__ import annotations

import collections
import contextlib
import ctypes as C
import math
import os

import numpy as np
import torch

from common.enums import SimulationScheme
from mcre import binding as B
from mcre import runtime as RT
from mcre.dual import D, cholesky_dual
from mcre.lsm import (backward_induction, backward_induction_device, regression_tangents, solve_normal_equations, solve_normal_equations_batch,
                      to_raw_basis)
from mcre.timegrid import build_time_grid
from metrics.metric import MetricType
from models.cirpp import CIRPPModel
from models.model_config import ModelConfig
from models.vasicek import VasicekModel
from products.bermudan_option import BermudanOption
from products.bond import Bond
from products.product import OptionType
from products.swap import InterestRateSwap, IRSType

CHUNK_PATHS = 4096

# Lowered plans of runs that differ only in the inter-model correlation (the rho sweep of the wrong-way-risk CVA
# config): every table except the Cholesky factor is a function of (model parameters, schedules, timelines, metric
# list), so it is built once and reused.  Keyed on the VALUES of everything lower() reads (never on object identity);
# value-only plans of linear products on plain Vasicek (+ CIR++) models only.  MCRE_PLAN_CACHE=0 switches it off.
_PLAN_CACHE = collections.OrderedDict()
_PLAN_CACHE_MAX = 64


def plan_cache_enabled():
    return os.environ.get("MCRE_PLAN_CACHE", "1") != "0"


def _sig_bond(b):
    return ("B", float(b.startdate), float(b.maturity), float(b.notional), float(b.tenor),
            None if b.fixed_rate is None else float(b.fixed_rate), bool(b.pays_notional), tuple(b.payment_dates.tolist()))


def _sig_product(p):
    if type(p) is Bond:
        return _sig_bond(p)
    if type(p) is InterestRateSwap:
        return ("S", p.irs_type.name, _sig_bond(p.fixed_leg), _sig_bond(p.floating_leg))
    return None


def _packs(duals):
    if not duals:
        return np.zeros(0)
    if duals[0].t.shape[0] == 0:
        return np.array([d.v for d in duals], dtype=np.float64)      # value-only plan
    return np.concatenate([d.pack() for d in duals])


class LinearLeg:
    """Fixed and floating cashflow weights of one linear product on its payment dates."""

    def __init__(self):
        self.fixed = {}   # date -> amount
        self.floating = {}  # date -> {(t1, t2): weight}

    def add_bond(self, bond: Bond, sign: float):
        dates = bond.payment_dates.tolist()
        accr = bond.accrual_fractions()
        last = len(dates) - 1
        if bond.is_fixed():
            k = float(bond.fixed_rate)
            for i, (d, a) in enumerate(zip(dates, accr)):
                amt = k * a + (float(bond.notional) if (bond.pays_notional and i == last) else 0.0)
                self.fixed[d] = self.fixed.get(d, 0.0) + sign * amt
        else:
            for i, (d, a, per) in enumerate(zip(dates, accr, bond.libor_periods())):
                self.floating.setdefault(d, {})
                self.floating[d][per] = self.floating[d].get(per, 0.0) + sign * a
                if bond.pays_notional and i == last:
                    self.fixed[d] = self.fixed.get(d, 0.0) + sign * float(bond.notional)


def linear_leg_of(product):
    leg = LinearLeg()
    if isinstance(product, Bond):
        leg.add_bond(product, 1.0)
    elif isinstance(product, InterestRateSwap):
        s = 1.0 if product.irs_type == IRSType.PAYER else -1.0
        leg.add_bond(product.floating_leg, s)
        leg.add_bond(product.fixed_leg, -s)
    else:
        raise TypeError(type(product))
    return leg


def underlying_terms(und, t_obs):
    """Value of an option underlying observed at t_obs as const + sum_T w_T P(t_obs, T)
    (reference: bond.py:115-163, swap.py:129-140): the underlying is re-scheduled from the
    observation date and valued on forward zero bonds, notional included."""
    const, terms = 0.0, {}

    def add(T, w):
        nonlocal const
        if T == t_obs:
            const += w          # P(t, t) = 1
            return
        for key in terms:
            if abs(key - T) < 1e-9:
                terms[key] += w
                return
        terms[T] = w

    def add_bond(bond, sign):
        b = bond.with_startdate(t_obs)
        dates, notional = b.payment_dates.tolist(), float(b.notional)
        if b.is_fixed():
            prev = t_obs
            for d in dates:
                add(d, sign * notional * float(b.fixed_rate) * (d - prev))
                prev = d
        else:
            starts = [per[0] for per in b.libor_periods()] + [float(b.maturity)]
            for i in range(len(dates)):
                add(starts[i], sign * notional)
                add(starts[i + 1], -sign * notional)
        if b.pays_notional:
            add(dates[-1], sign * notional)

    if isinstance(und, Bond):
        add_bond(und, 1.0)
    elif isinstance(und, InterestRateSwap):
        s = 1.0 if und.irs_type == IRSType.PAYER else -1.0
        swap = und.with_startdate(t_obs)
        add_bond(swap.floating_leg, s)
        add_bond(swap.fixed_leg, -s)
    else:
        raise TypeError(type(und))
    return const, sorted(terms.items())


def is_linear(p):
    return isinstance(p, (Bond, InterestRateSwap))


def is_rate_bermudan(p):
    return isinstance(p, BermudanOption) and isinstance(p.underlying, (Bond, InterestRateSwap))


def is_rate_european(p):
    from products.european_option import EuropeanOption
    return isinstance(p, EuropeanOption) and isinstance(p.underlying, (Bond, InterestRateSwap))


def with_single_exercise_proxies(ctrl):
    """European options on bonds / swaps (european_option.py:45-68 with a rate underlying; the reference's
    pv_european_bond_option.py, ee_pfe_swaption.py) are the one-date case of the exercise units the kernels have: the
    payoff max(+-(U - K), 0) is "exercise iff positive" with no continuation after the only date, the exposure proxy before
    it is the alive-state regression, and there is none from the exercise date on - the same numbers the reference's
    one-state European product gives.  -> a view of the controller whose rate Europeans are replaced by one-date
    BermudanOption proxies (own coefficient tensors), or the controller itself if it has none."""
    import copy
    if not any(is_rate_european(p) for p in ctrl.products):
        return ctrl
    view = copy.copy(ctrl)
    view.regression_coeffs = list(ctrl.regression_coeffs)
    proxies, sets = [], []
    for ns in ctrl.netting_sets:
        ns_view = copy.copy(ns)
        ns_view.products = []
        for p in ns.products:
            if is_rate_european(p):
                q = BermudanOption(p.underlying, [float(p.exercise_date[0])], float(p.strike[0]), p.option_type,
                                   asset_id=p.asset_ids[0])
                q.name, q.product_id = p.name if p.name else "EuropeanOption", p.product_id
                # (quadratic basis of the kernels; a run that regresses with another basis is refused by the controller, a
                # PV run - nothing regressed - may carry any regression_function, pv_european_bond_option.py:53)
                q.regression_coeffs = torch.zeros((1, 2, 3), dtype=torch.float64)
                view.regression_coeffs[p.product_id] = torch.zeros((len(ctrl.exposure_timeline), 2, 3), dtype=torch.float64)
                proxies.append((p, q))
                p = q
            ns_view.products.append(p)
        sets.append(ns_view)
    view.netting_sets = sets
    view.products = [p for ns in sets for p in ns.products]
    view.requires_regression = any(view._product_requires_regression(p) for p in view.products)
    view._european_proxies = proxies
    return view


def split_model(model):
    """(vasicek, cir or None, vas_index, cir_index) or None if not an IRC model."""
    if isinstance(model, VasicekModel):
        return model, None, 0, -1
    if isinstance(model, ModelConfig):
        ms = model.models
        vas = [i for i, m in enumerate(ms) if isinstance(m, VasicekModel)]
        cir = [i for i, m in enumerate(ms) if isinstance(m, CIRPPModel)]
        if len(vas) == 1 and len(cir) <= 1 and len(vas) + len(cir) == len(ms):
            if model.id_to_model["numeraire"] != vas[0]:
                return None
            return ms[vas[0]], (ms[cir[0]] if cir else None), vas[0], (cir[0] if cir else -1)
    return None


class IrcBackend:
    @staticmethod
    def supports(ctrl):
        if split_model(ctrl.model) is None:
            return False
        return all(is_linear(p) or is_rate_bermudan(p) for p in ctrl.products)

    def __init__(self, ctrl):
        self.c = ctrl
        self.vas, self.cir, self.vas_idx, self.cir_idx = split_model(ctrl.model)
        self.has_cir = self.cir is not None
        scheme = ctrl.simulation_scheme
        if scheme == SimulationScheme.ANALYTICAL and self.has_cir:
            raise NotImplementedError("Inter covariance not implemented for the requested pair of models.")
        if scheme not in (SimulationScheme.EULER, SimulationScheme.ANALYTICAL):
            raise NotImplementedError(f"Scheme {scheme} is not defined for Vasicek / CIR++ models.")
        self.scheme = scheme
        self.nt = len(ctrl.model.model_params) if ctrl.differentiate else 0
        self._keep = []
        #: set by mcre/hybrid.py: {"ext_rate": rate of the numeraire model, "chunk": reduction chunk shared with the
        #: equity launches, "captured": [(set indices, exposure spill, pv spill, pv shift)] filled by run()}
        self.hybrid = None

    # ------------------------------------------------------------------ lowering
    def _dual_params(self):
        if getattr(self, "_dual_cache", None) is None:
            self._dual_cache = self._dual_params_uncached()
        return self._dual_cache

    def _dual_params_uncached(self):
        nt = self.nt
        if isinstance(self.c.model, ModelConfig):
            offs = self.c.model.param_offsets()
            pv = self.vas.dual_params(offs[self.vas_idx], nt)
            pc = self.cir.dual_params(offs[self.cir_idx], nt) if self.has_cir else None
        else:
            pv, pc = self.vas.dual_params(0, nt), None
        return pv, pc

    def basis_at(self, t):
        """Standardisation (shift, scale) of the explanatory variable r(t): its Vasicek mean /
        std (pure conditioning aid; the fitted values do not depend on it)."""
        cache = self.__dict__.setdefault("_basis_cache", {})
        hit = cache.get(t)
        if hit is not None:
            return hit
        pv, _ = self._dual_params()
        r0, sig, th, a = (x.v for x in pv)
        tau = t - self.vas.t0()
        if tau <= 0:
            out = (r0, 1.0)
        else:
            mean = th + (r0 - th) * math.exp(-a * tau)
            std = sig * math.sqrt((1.0 - math.exp(-2.0 * a * tau)) / (2.0 * a))
            out = (mean, 1.0 / std if std > 0 else 1.0)
        cache[t] = out
        return out

    def lower(self, set_indices, unit_products, berm_units=None, reg_times=None):
        """Build the descriptor for a group of netting sets (main) and regression units
        (pre-simulation).  `berm_units`: [(BermudanOption, set row)] (default: the Bermudan
        options of the given sets); `reg_times`: regression dates of the pre-simulation
        (default: the internal exposure dates).
        Returns (desc, tables) with `tables` keeping numpy arrays alive."""
        key = self._lower_key(set_indices, unit_products, berm_units, reg_times)
        if key is not None:
            hit = _PLAN_CACHE.get(key)
            if hit is not None:
                _PLAN_CACHE.move_to_end(key)
                return self._lower_from_cache(hit)
        out = self._lower_uncached(set_indices, unit_products, berm_units, reg_times)
        if key is not None:
            _PLAN_CACHE[key] = out
            while len(_PLAN_CACHE) > _PLAN_CACHE_MAX:
                _PLAN_CACHE.popitem(last=False)
        return out

    def _lower_key(self, set_indices, unit_products, berm_units, reg_times):
        """Hashable value signature of everything _lower_uncached reads except the inter-model correlation; None when
        the plan is not cacheable (tangents, exercise units, model extensions)."""
        c = self.c
        if not plan_cache_enabled() or self.nt != 0 or berm_units or reg_times is not None:
            return None
        if type(self.vas) is not VasicekModel or (self.has_cir and type(self.cir) is not CIRPPModel):
            return None
        sets = [c.netting_sets[i] for i in set_indices]
        prods = [p for ns in sets for p in ns.products] + list(unit_products)
        sigs = [_sig_product(p) for p in prods]
        if any(sg is None for sg in sigs):
            return None
        model_sig = (tuple(self.vas.param_values()), float(self.vas.t0()), self.vas_idx, self.cir_idx,
                     None if self.hybrid is None else float(self.hybrid["ext_rate"]))
        if self.has_cir:
            model_sig += (tuple(self.cir.param_values()), float(self.cir.t0()), tuple(self.cir.tenors.tolist()),
                          tuple(self.cir.hazard_rates.tolist()), bool(self.cir.deterministic), tuple(self.cir.asset_ids))
        rm = c.risk_metrics
        metric_sig = tuple((m.metric_type.name, getattr(m, "counterparty_id", None), getattr(m, "recovery_rate", None))
                           for m in rm.metrics)
        set_sig = tuple((float(ns.threshold), ns.margin_period_of_risk, ns.counterparty_id, len(ns.products),
                         tuple(c.netting_set_delayed_exposure_indices[i].tolist()))
                        for i, ns in zip(set_indices, sets))
        return (model_sig, self.scheme.name, int(c.num_steps), tuple(c.simulation_timeline.tolist()),
                tuple(c.exposure_timeline.tolist()), tuple(c.metric_exposure_timeline.tolist()),
                tuple(c.metric_exposure_indices.tolist()), metric_sig, set_sig, tuple(sigs), len(unit_products))

    def _cholesky(self):
        """Lower Cholesky factor of the joint correlation as [L00, L01, L10, L11] duals (model.py:50-73)."""
        nt = self.nt
        zero = D(0.0, None, nt)
        if isinstance(self.c.model, ModelConfig) and self.has_cir:
            pv, pc = self._dual_params()
            sub = [None, None]
            sub[self.vas_idx], sub[self.cir_idx] = pv, pc
            corr = self.c.model.joint_correlation(self.scheme, sub)
            L = cholesky_dual(corr)
            return [L[0][0], L[0][1], L[1][0], L[1][1]]
        one = D(1.0, None, nt)
        return [one, zero, zero, one]

    def _lower_from_cache(self, hit):
        """A cached plan with this run's Cholesky factor: copy of the descriptor, shared (read-only) tables."""
        desc, keep, info = hit
        d = B.IrcDesc.from_buffer_copy(desc)
        t = dict(keep)
        t["chol"], d.chol = B.as_dp(_packs(self._cholesky()))
        return d, t, info

    def _lower_uncached(self, set_indices, unit_products, berm_units=None, reg_times=None):
        c, nt = self.c, self.nt
        w = 1 + nt
        pv, pc = self._dual_params()
        t0 = self.vas.t0()
        grid = build_time_grid(t0, c.simulation_timeline.tolist(), c.num_steps)
        dates = grid.dates
        date_idx = grid.index_map()
        n_dates, n_sub = len(dates), grid.n_sub
        zero = D(0.0, None, nt)

        # ---- correlation / Cholesky -------------------------------------------------
        chol = self._cholesky()

        # ---- per-step model scalars (shared by the pre-simulation and main plans) ---------
        step_vas, step_cir = [], []
        cached = getattr(self, "_step_cache", None)
        if cached is not None:
            step_vas, step_cir = cached
        fast = cached is None and nt == 0 and self.has_cir and not self.cir.deterministic and self.scheme != SimulationScheme.ANALYTICAL
        if fast:
            # value-only plan: the per-step tables in one vectorised evaluation
            psis = self.cir.psi_values([x.v for x in pc], grid.t1)
            theta_t = [self.vas.mean_level(pv, t) for t in grid.t1]
            for s in range(n_sub):
                step_vas += [theta_t[s], zero]
                step_cir += [D._val(float(psis[s])), zero]
            cached = (step_vas, step_cir)
        for s in range(n_sub if cached is None else 0):
            if self.scheme == SimulationScheme.ANALYTICAL:
                decay, _ = self.vas.exact_step_constants(pv, grid.dt[s])
                _, nstd = self.vas.exact_step_constants(pv, grid.dt_nominal[s])
                step_vas += [decay, nstd]
            else:
                step_vas += [self.vas.mean_level(pv, grid.t1[s]), zero]
            if self.has_cir:
                if self.cir.deterministic:
                    step_cir += [D(self.cir.market_hazard(grid.t1[s]), None, nt),
                                 D(self.cir.market_hazard(grid.t2[s]), None, nt)]
                else:
                    step_cir += [self.cir.psi(pc, grid.t1[s]), zero]
        self._step_cache = (step_vas, step_cir)
        if self.has_cir:
            cir_init = D(self.cir.market_hazard(t0), None, nt) if self.cir.deterministic else pc[3]

        # ---- cashflow tables ---------------------------------------------------------
        sets = [c.netting_sets[i] for i in set_indices]
        need_pv = c.risk_metrics.requires_discounted_cashflows()
        set_legs = [[linear_leg_of(p) for p in ns.products if is_linear(p)] for ns in sets]
        if berm_units is None:
            berm_units = [(p, r) for r, ns in enumerate(sets) for p in ns.products if is_rate_bermudan(p)]
        if len(berm_units) > B.IRC_MAX_BERM:
            raise NotImplementedError(f"at most {B.IRC_MAX_BERM} Bermudan options per launch group")
        unit_legs = [linear_leg_of(p) for p in unit_products]
        float_keys = [[] for _ in range(n_dates)]   # per date: list of (t1,t2)
        for leg in [l for legs in set_legs for l in legs] + unit_legs:
            for d, per in leg.floating.items():
                for key in per:
                    if key not in float_keys[date_idx[d]]:
                        float_keys[date_idx[d]].append(key)
        float_off = [0]
        for di in range(n_dates):
            float_off.append(float_off[-1] + len(float_keys[di]))
        n_float = float_off[-1]
        float_pos = {(di, key): float_off[di] + j for di in range(n_dates) for j, key in enumerate(float_keys[di])}
        float_coef, float_inv_tau = [], np.zeros(n_float)
        for di in range(n_dates):
            for key in float_keys[di]:
                alpha, Bc = self.vas.bond_coefficients(pv, key[0], key[1])
                float_coef += [alpha, Bc]
                float_inv_tau[float_pos[(di, key)]] = 1.0 / (key[1] - key[0])

        def leg_tables(legs_per_row):
            fix = np.zeros((len(legs_per_row), n_dates))
            flo = np.zeros((len(legs_per_row), max(n_float, 1)))[:, :n_float]
            for r, legs in enumerate(legs_per_row):
                for leg in legs:
                    for d, amt in leg.fixed.items():
                        fix[r, date_idx[d]] += amt
                    for d, per in leg.floating.items():
                        for key, wgt in per.items():
                            flo[r, float_pos[(date_idx[d], key)]] += wgt
            return fix, flo

        set_fix, set_float = leg_tables(set_legs)
        unit_fix, unit_float = leg_tables([[l] for l in unit_legs])

        # ---- date tables -------------------------------------------------------------
        expo_times = c.exposure_timeline.tolist() if c.risk_metrics.requires_exposure_profiles() else []
        metric_times = c.metric_exposure_timeline.tolist() if c.risk_metrics.requires_exposure_profiles() else []
        n_expo, n_metric = len(expo_times), len(metric_times)
        flags = np.zeros(n_dates, dtype=np.int32)
        date_expo = np.full(n_dates, -1, dtype=np.int32)
        date_metric = np.full(n_dates, -1, dtype=np.int32)
        date_reg = np.full(n_dates, -1, dtype=np.int32)
        cash_dates = set()
        for leg in [l for legs in set_legs for l in legs] + unit_legs:
            cash_dates.update(leg.fixed.keys())
            cash_dates.update(leg.floating.keys())
        for d in cash_dates:
            flags[date_idx[d]] |= B.DATE_HAS_CASHFLOW
        for e, t in enumerate(expo_times):
            flags[date_idx[t]] |= B.DATE_HAS_EXPOSURE
            date_expo[date_idx[t]] = e
        reg_list = expo_times if reg_times is None else list(reg_times)
        for k, t in enumerate(reg_list):
            flags[date_idx[t]] |= B.DATE_HAS_REGRESSION
            date_reg[date_idx[t]] = k
        for m, t in enumerate(metric_times):
            flags[date_idx[t]] |= B.DATE_HAS_METRIC
            date_metric[date_idx[t]] = m

        basis = np.array([self.basis_at(t) for t in expo_times]).reshape(n_expo, 2)
        reg_basis = np.array([self.basis_at(t) for t in reg_list]).reshape(len(reg_list), 2)

        # ---- Bermudan exercise units: per date, the zero-bond decomposition of the underlying ----
        ex_by_date = [[] for _ in range(n_dates)]
        for b, (prod, _) in enumerate(berm_units):
            ex_dates = prod.product_timeline.tolist()
            for i, t in enumerate(ex_dates):
                ex_by_date[date_idx[t]].append((b, i, t, i == len(ex_dates) - 1))
        ex_off, ex_unit, ex_last, ex_const, ex_term_off, term_coef, term_w, ex_basis = [0], [], [], [], [0], [], [], []
        ex_index = {}
        for di in range(n_dates):
            for b, i, t, last in ex_by_date[di]:
                # (memoised per backend: the pre-simulation plan and the main plan ask for the same decompositions)
                memo = self.__dict__.setdefault("_terms_cache", {})
                key_ut = (id(berm_units[b][0].underlying), t)
                if key_ut not in memo:
                    memo[key_ut] = underlying_terms(berm_units[b][0].underlying, t)
                const, terms = memo[key_ut]
                ex_index[(b, i)] = len(ex_unit)
                ex_unit.append(b)
                ex_last.append(int(last))
                ex_const.append(const)
                for T, wgt in terms:
                    alpha, Bc = self.vas.bond_coefficients(pv, t, T)
                    term_coef += [alpha, Bc]
                    term_w.append(wgt)
                ex_term_off.append(len(term_w))
                ex_basis.append(self.basis_at(t))
                flags[di] |= B.DATE_HAS_EXERCISE
            ex_off.append(len(ex_unit))

        # ---- netting-set terms ---------------------------------------------------------
        metrics = c.risk_metrics.metrics
        acc = 0
        if need_pv:
            acc |= B.ACC_PV
        kinds = {m.metric_type for m in metrics}
        if kinds & {MetricType.CE, MetricType.EPE, MetricType.EEPE}:
            acc |= B.ACC_POS
        if MetricType.ENE in kinds:
            acc |= B.ACC_NEG
        if MetricType.PFE in kinds:
            acc |= B.ACC_SPILL
        cva_metrics = [m for m in metrics if m.metric_type == MetricType.CVA]
        lgd = 0.0
        cva_coef = [zero] * (2 * n_metric)
        cva_metric = None
        if cva_metrics and sets:
            if len({m.counterparty_id for m in cva_metrics}) > 1 or len({m.recovery_rate for m in cva_metrics}) > 1:
                raise NotImplementedError("one CVA counterparty / recovery per run is supported for now")
            cva_metric = cva_metrics[0]
            if not self.has_cir or cva_metric.counterparty_id not in self.cir.asset_ids:
                raise Exception("Not all models set for xVA valuation.")
            acc |= B.ACC_CVA
            lgd = 1.0 - cva_metric.recovery_rate
            cva_coef = []
            if nt == 0 and not self.cir.deterministic and n_metric > 1:
                Cs, Bs = self.cir.conditional_survival_values([x.v for x in pc], metric_times[:-1], metric_times[1:])
                for Ck, Bk in zip(Cs, Bs):
                    cva_coef += [D._val(float(Ck)), D._val(float(Bk))]
                cva_coef += [zero, zero]
            else:
                for m in range(n_metric):
                    if m < n_metric - 1:
                        Ck, Bk = self.cir.conditional_survival_coefficients(pc, metric_times[m], metric_times[m + 1])
                        cva_coef += [Ck, Bk]
                    else:
                        cva_coef += [zero, zero]
        set_thr = np.array([ns.threshold for ns in sets], dtype=np.float64)
        set_flags = np.zeros(len(sets), dtype=np.int32)
        set_lag = np.full((len(sets), max(n_metric, 1)), -1, dtype=np.int32)[:, :n_metric]
        for r, (si, ns) in enumerate(zip(set_indices, sets)):
            if ns.is_collateralized():
                set_flags[r] |= 1
                delayed = c.netting_set_delayed_exposure_indices[si].tolist()
                for m in range(n_metric):
                    if delayed[m] >= 0:
                        lag = int(c.metric_exposure_indices[m]) - delayed[m]
                        # the general template (tangents, exercise units) keeps the last IRC_MAX_LAG exposures of a path in
                        # registers; the value-only kernel a ring in shared memory sized for the plan's largest lag
                        general = nt != 0 or bool(berm_units)
                        if lag >= (B.IRC_MAX_LAG if general else 64):
                            raise NotImplementedError(
                                f"MPoR look-back spans {lag} exposure dates; this kernel keeps {(B.IRC_MAX_LAG if general else 64) - 1}")
                        set_lag[r, m] = lag
            if cva_metric is not None and (ns.counterparty_id is None or ns.counterparty_id == cva_metric.counterparty_id):
                set_flags[r] |= 2

        t = {}  # keep-alive of every host array referenced by the descriptor
        d = B.IrcDesc()
        d.nt, d.scheme = nt, (B.SCHEME_ANALYTICAL if self.scheme == SimulationScheme.ANALYTICAL else B.SCHEME_EULER)
        d.has_cir = int(self.has_cir)
        d.cir_deterministic = int(self.has_cir and self.cir.deterministic)
        d.vas_noise = self.vas_idx if self.has_cir else 0
        d.cir_noise = self.cir_idx if self.has_cir else 0
        if self.hybrid is not None:
            d.ext_numeraire, d.ext_rate = 1, float(self.hybrid["ext_rate"])
            d.ext_slot = int(self.hybrid.get("ext_slot", -1)) if nt else -1

        def fp(name, arr):
            t[name], ptr = B.as_dp(arr)
            return ptr

        def ip(name, arr):
            t[name], ptr = B.as_ip(arr)
            return ptr

        d.vas = fp("vas", _packs(pv))
        d.cir = fp("cir", _packs(pc) if self.has_cir else np.zeros(1))
        d.cir_init = fp("cir_init", cir_init.pack() if self.has_cir else np.zeros(1))
        d.chol = fp("chol", _packs(chol))
        d.n_sub, d.n_dates, d.n_pre_dates = n_sub, n_dates, grid.n_pre_dates
        d.step_dt = fp("step_dt", np.array(grid.dt))
        d.step_date = ip("step_date", np.array(grid.date_after))
        d.step_vas = fp("step_vas", _packs(step_vas))
        d.step_cir = fp("step_cir", _packs(step_cir) if self.has_cir else np.zeros(1))
        d.date_flags, d.date_expo = ip("flags", flags), ip("date_expo", date_expo)
        d.date_metric, d.date_reg = ip("date_metric", date_metric), ip("date_reg", date_reg)
        d.date_float_off = ip("float_off", np.array(float_off))
        d.float_coef = fp("float_coef", _packs(float_coef) if float_coef else np.zeros(1))
        d.float_inv_tau = fp("float_inv_tau", float_inv_tau if n_float else np.zeros(1))
        d.n_sets, d.n_expo, d.n_metric, d.acc_flags = len(sets), n_expo, n_metric, acc
        d.set_fix, d.set_float = fp("set_fix", set_fix), fp("set_float", set_float if n_float else np.zeros(1))
        d.set_threshold, d.set_flags = fp("set_thr", set_thr), ip("set_flags", set_flags)
        d.set_lag = ip("set_lag", set_lag if n_metric else np.zeros(1))
        d.expo_coef = fp("expo_coef", np.zeros(max(n_expo * len(sets) * 3 * w, 1)))
        d.expo_basis = fp("expo_basis", basis if n_expo else np.zeros(1))
        d.cva_coef = fp("cva_coef", _packs(cva_coef) if n_metric else np.zeros(1))
        d.lgd = lgd
        d.n_units, d.n_reg = len(unit_products), len(reg_list)
        d.unit_fix = fp("unit_fix", unit_fix if unit_products else np.zeros(1))
        d.unit_float = fp("unit_float", unit_float if (unit_products and n_float) else np.zeros(1))
        d.unit_last_reg = ip("unit_last_reg", np.zeros(max(len(unit_products), 1)))
        d.reg_basis = fp("reg_basis", reg_basis if len(reg_list) else np.zeros(1))
        d.n_berm = len(berm_units)
        if berm_units:
            d.berm_set = ip("berm_set", np.array([r for _, r in berm_units]))
            d.berm_strike = fp("berm_strike", np.array([float(p.strike) for p, _ in berm_units]))
            d.berm_sign = fp("berm_sign", np.array([1.0 if p.option_type == OptionType.CALL else -1.0
                                                    for p, _ in berm_units]))
            d.date_ex_off, d.ex_unit, d.ex_last = ip("ex_off", ex_off), ip("ex_unit", ex_unit), ip("ex_last", ex_last)
            d.ex_term_off, d.ex_const = ip("ex_term_off", ex_term_off), fp("ex_const", ex_const)
            d.term_coef = fp("term_coef", _packs(term_coef) if term_coef else np.zeros(1))
            d.term_w = fp("term_w", term_w if term_w else np.zeros(1))
            d.ex_basis = fp("ex_basis", np.array(ex_basis))
        info = dict(grid=grid, n_expo=n_expo, n_metric=n_metric, basis=basis, acc=acc, reg_basis=reg_basis,
                    set_flags=set_flags, noise_dim=2 if self.has_cir else 1, n_float=n_float,
                    berm_units=berm_units, ex_index=ex_index, n_ex=len(ex_unit), reg_times=reg_list,
                    expo_times=expo_times)
        return d, t, info

    def param_used(self, kind):
        """Which model parameters a metric's value is connected to.  The reference returns
        None from autograd only for parameters outside the graph (controller.py:618-624);
        a ModelConfig writes every sub-model's step into one state tensor
        (model_config.py:261-276), so all of its parameters are connected and unused ones
        come back as 0.0 - verified against the reference (tests/golden/wwr_cva_greeks.json)."""
        return [True] * len(self.c.model.model_params)

    # ------------------------------------------------------------------ execution
    def _rng(self, seed, inject, n_total):
        r = B.Rng()
        r.seed, r.stream = seed, self.c.rng_stream
        if inject is not None:
            r.mode = B.RNG_INJECT
            r.d_z = inject.data_ptr()
            r.d_u = None
            r.n_paths_total = n_total
        else:
            r.mode = B.RNG_PHILOX
            r.n_paths_total = n_total
        return r

    def presim_coefficients(self, products, dev, while_running=None):
        """Regression coefficients [product][n_expo][3] in the standardised basis.
        `while_running`: host work to do after the first group's kernels are queued and before
        their moments are read back (the main plan's lowering overlaps the pre-simulation)."""
        c = self.c
        L = B.lib()
        n_pre = c.num_paths_presim
        out = {}
        if n_pre <= 0:
            raise ValueError("Exposure metrics need a pre-simulation: num_paths_presim must be positive.")
        inject = c.injected_normals.get("pre") if c.injected_normals else None
        for g0 in range(0, len(products), B.IRC_MAX_UNITS):
            group = products[g0:g0 + B.IRC_MAX_UNITS]
            desc, keep, info = self.lower([], group)
            plan = C.c_void_p()
            B.check(L.mcre_irc_create(C.byref(desc), C.byref(plan)))
            try:
                begin, count = RT.shard_range(n_pre, CHUNK_PATHS)
                slots = L.mcre_irc_presim_slots(plan)
                tslots = L.mcre_irc_presim_tangent_slots(plan) if self.nt else 0
                moments = torch.zeros(slots + tslots, dtype=torch.float64, device=dev)
                # bound the scratch: process the local paths in batches of whole chunks
                per_path = L.mcre_irc_presim_scratch_bytes(plan, CHUNK_PATHS) // CHUNK_PATHS
                batch = c.presim_batch_paths * 20 * max(info["n_expo"], 1) // max(per_path, 1) if self.nt else c.presim_batch_paths
                batch = max(CHUNK_PATHS, (batch // CHUNK_PATHS) * CHUNK_PATHS)
                for b0 in range(0, count, batch):
                    bn = min(batch, count - b0)
                    scratch = torch.empty(L.mcre_irc_presim_scratch_bytes(plan, bn), dtype=torch.uint8, device=dev)
                    partial = torch.empty(L.mcre_irc_partial_bytes(plan, bn, CHUNK_PATHS, 1) // 8 + 1,
                                          dtype=torch.float64, device=dev)
                    part_m = torch.zeros(slots + tslots, dtype=torch.float64, device=dev)
                    rng = self._rng(42, inject, n_pre)
                    sh = B.Shard(begin + b0, bn, CHUNK_PATHS)
                    B.check(L.mcre_irc_presim(plan, C.byref(rng), C.byref(sh), scratch.data_ptr(),
                                              partial.data_ptr(), part_m.data_ptr(),
                                              part_m[slots:].data_ptr() if tslots else None, RT.stream_ptr()))
                    moments += part_m
                    del scratch, partial
                if while_running is not None:
                    while_running()
                    while_running = None
                moments = RT.all_reduce_tree(moments)
                mom_all = RT.to_host(moments)
                mom = mom_all[:slots].reshape(info["n_expo"], -1)
                tmom = mom_all[slots:].reshape(len(group), self.nt, info["n_expo"], 9) if tslots else None
            finally:
                L.mcre_irc_destroy(plan)
            nu = 1 if len(group) <= 1 else (2 if len(group) <= 2 else 4)
            assert mom.shape[1] == 5 + 3 * nu
            G = mom[:, [[0, 1, 2], [1, 2, 3], [2, 3, 4]]]          # [n_expo, 3, 3] Gram matrices of [1, u, u^2]
            for u, prod in enumerate(group):
                rhs = mom[:, 5 + 3 * u: 8 + 3 * u]
                coefs = solve_normal_equations_batch(G, rhs)
                dcoefs = None
                if tmom is not None:
                    # d(coefficients)/d(model parameters): the reference differentiates through lstsq
                    dcoefs = np.stack([regression_tangents(G[k], rhs[k], coefs[k], tmom[u, :, k, :])
                                       for k in range(info["n_expo"])])            # [n_expo, nt, 3]
                out[id(prod)] = (coefs, info["basis"], dcoefs)
        return out

    def _device_regression_ok(self, prods, groups):
        """Pre-simulation -> normal equations -> main simulation as one stream of kernels (no read-back in between):
        value-only runs of linear products whose regression units fit one launch and whose sets fit one group."""
        c = self.c
        return (os.environ.get("MCRE_DEVICE_SOLVE", "1") != "0" and self.nt == 0 and len(groups) == 1
                and 0 < len(prods) <= B.IRC_MAX_UNITS and all(is_linear(p) for p in c.products))

    def presim_on_device(self, prods, dev, while_running=None):
        """Like presim_coefficients, but the moments never leave the device: they are all-reduced and solved there
        (mcre_irc_solve_coefficients).  -> (coef_unit [n_units][n_expo][3], coef_sum [n_expo][n_sets][3]) device
        tensors in the standardised basis, and the basis."""
        c = self.c
        L = B.lib()
        n_pre = c.num_paths_presim
        if n_pre <= 0:
            raise ValueError("Exposure metrics need a pre-simulation: num_paths_presim must be positive.")
        inject = c.injected_normals.get("pre") if c.injected_normals else None
        desc, keep, info = self.lower([], prods)
        n_expo, n_sets = info["n_expo"], len(c.netting_sets)
        plan = C.c_void_p()
        B.check(L.mcre_irc_create(C.byref(desc), C.byref(plan)))
        try:
            begin, count = RT.shard_range(n_pre, CHUNK_PATHS)
            slots = L.mcre_irc_presim_slots(plan)
            moments = torch.zeros(slots, dtype=torch.float64, device=dev)
            batch = max(CHUNK_PATHS, (c.presim_batch_paths // CHUNK_PATHS) * CHUNK_PATHS)
            for b0 in range(0, count, batch):
                bn = min(batch, count - b0)
                scratch = torch.empty(L.mcre_irc_presim_scratch_bytes(plan, bn), dtype=torch.uint8, device=dev)
                partial = torch.empty(L.mcre_irc_partial_bytes(plan, bn, CHUNK_PATHS, 1) // 8 + 1, dtype=torch.float64, device=dev)
                part_m = moments if count <= batch else torch.zeros(slots, dtype=torch.float64, device=dev)
                rng = self._rng(42, inject, n_pre)
                sh = B.Shard(begin + b0, bn, CHUNK_PATHS)
                B.check(L.mcre_irc_presim(plan, C.byref(rng), C.byref(sh), scratch.data_ptr(), partial.data_ptr(),
                                          part_m.data_ptr(), None, RT.stream_ptr()))
                if part_m is not moments:
                    moments += part_m
                del scratch, partial
            if while_running is not None:
                while_running()
            moments = RT.all_reduce_tree(moments)
            set_of = {id(p): si for si, ns in enumerate(c.netting_sets) for p in ns.products}
            unit_set, unit_ptr = B.as_ip(np.array([set_of[id(p)] for p in prods], dtype=np.int32))
            coef_unit = torch.empty((len(prods), n_expo, 3), dtype=torch.float64, device=dev)
            coef_sum = torch.empty((n_expo, n_sets, 3), dtype=torch.float64, device=dev)
            B.check(L.mcre_irc_solve_coefficients(plan, moments.data_ptr(), unit_ptr, n_sets, coef_unit.data_ptr(),
                                                  coef_sum.data_ptr(), RT.stream_ptr()))
        finally:
            L.mcre_irc_destroy(plan)
        return coef_unit, coef_sum, info["basis"]

    @contextlib.contextmanager
    def _value_only(self):
        """Lower without tangents: the Longstaff-Schwartz policy enters the main pass through hard
        exercise indicators only (bermudan_option.py:121), whose derivative is zero, so its
        pre-simulation needs no tangents even when differentiate=True."""
        saved = (self.nt, getattr(self, "_dual_cache", None), getattr(self, "_step_cache", None))
        self.nt, self._dual_cache, self._step_cache = 0, None, None
        try:
            yield
        finally:
            self.nt, self._dual_cache, self._step_cache = saved

    def presim_bermudan(self, prod, dev):
        """Longstaff-Schwartz pre-simulation of one Bermudan option (controller.py:294-383):
        forward pass spills x, numeraire and immediate values; one fused moments(+exercise
        update) launch per regression date, latest first; the 3x3 normal equations are solved
        on the host (multi-GPU: after an all-reduce of the 8 moments).
        -> (coef per regression date [n_reg][3] in the standardised basis, reg_times, reg_basis)"""
        c = self.c
        L = B.lib()
        n_pre = c.num_paths_presim
        expo_times = c.exposure_timeline.tolist() if c.risk_metrics.requires_exposure_profiles() else []
        ptl = prod.product_timeline.tolist()
        reg_times = sorted(set(prod.regression_timeline.tolist()) | set(expo_times))
        if len(ptl) == 1 and not expo_times:
            # one exercise date and no exposure dates (a European option's proxy in a PV run): nothing is regressed -
            # there is no continuation after the only date - and the reference does not pre-simulate either
            return np.zeros((len(reg_times), 3)), reg_times, np.array([self.basis_at(t) for t in reg_times]), None
        if n_pre <= 0:
            raise ValueError("Exercise products need a pre-simulation: num_paths_presim must be positive.")
        # tangents are needed only where the coefficients enter smoothly: the alive-state exposure proxies.  The
        # exercise policy is a hard indicator (zero derivative), so PV-only runs pre-simulate values only.
        with_tan = bool(self.nt) and bool(expo_times)
        if with_tan:
            desc, keep, info = self.lower([], [], berm_units=[(prod, 0)], reg_times=reg_times)
        else:
            with self._value_only():
                desc, keep, info = self.lower([], [], berm_units=[(prod, 0)], reg_times=reg_times)
        n_reg, n_ex = len(reg_times), info["n_ex"]
        basis = info["reg_basis"]
        inject = c.injected_normals.get("pre") if c.injected_normals else None
        plan = C.c_void_p()
        B.check(L.mcre_irc_create(C.byref(desc), C.byref(plan)))
        try:
            begin, count = RT.shard_range(n_pre, CHUNK_PATHS)
            n = max(count, 1)
            scratch = torch.zeros(L.mcre_irc_lsm_scratch_bytes(plan, n) // 8 + 1, dtype=torch.float64, device=dev)
            rng = self._rng(42, inject, n_pre)
            sh = B.Shard(begin, count, CHUNK_PATHS)
            B.check(L.mcre_irc_lsm_forward(plan, C.byref(rng), C.byref(sh), scratch.data_ptr(), RT.stream_ptr()))
        finally:
            L.mcre_irc_destroy(plan)
        xs = scratch[:n_reg * n].view(n_reg, n)
        ns_ = scratch[n_reg * n:2 * n_reg * n].view(n_reg, n)
        imm_all = scratch[2 * n_reg * n:(2 * n_reg + n_ex) * n].view(n_ex, n)
        assert [info["ex_index"][(0, i)] for i in range(len(ptl))] == list(range(len(ptl)))  # one unit: date order
        imm = imm_all
        tangents = None
        if with_tan:
            nt, off = self.nt, (2 * n_reg + n_ex) * n
            dxs = scratch[off:off + n_reg * nt * n].view(n_reg, nt, n)
            dns = scratch[off + n_reg * nt * n:off + 2 * n_reg * nt * n].view(n_reg, nt, n)
            dim_ = scratch[off + 2 * n_reg * nt * n:off + (2 * n_reg + n_ex) * nt * n].view(n_ex, nt, n)
            tangents = dict(nt=nt, dxs=dxs, dnums=dns, dimm=dim_)
        if not with_tan and os.environ.get("MCRE_DEVICE_SOLVE", "1") != "0":
            out = backward_induction_device(xs, ns_, imm, ptl, reg_times, basis, count, CHUNK_PATHS, dev)
        else:
            out = backward_induction(xs, ns_, imm, ptl, reg_times, basis, count, CHUNK_PATHS, dev, tangents=tangents)
        coef, dcoef = out if with_tan else (out, None)
        return coef, reg_times, basis, dcoef

    def _pfe_sensitivities(self, plan, quant, spill, begin, count, n_main, n_metric, n_sets, inject, dev):
        """Gradient of the PFE order statistics: the reference differentiates torch.sort(...)[index]
        (pfe_metric.py:59-71), i.e. the pathwise gradient of the selected path.  The selected paths are located
        in the spilled exposures (smallest global path id among ties, whose tangents agree anyway), replayed with
        tangents on every rank (Philox is counter based) and their unsecured-exposure tangents picked per date."""
        L = B.lib()
        for q, (vals, tans) in ((q, quant[0][q]) for q in list(quant[0])):
            targets = torch.tensor([[quant[r][q][0][m][0] for m in range(n_metric)] for r in range(n_sets)],
                                   dtype=torch.float64, device=dev).reshape(-1)
            index = torch.empty(n_sets * n_metric, dtype=torch.int64, device=dev)
            B.check(L.mcre_select_locate(spill.data_ptr(), spill.stride(1), count, n_sets * n_metric, targets.data_ptr(),
                                         index.data_ptr(), RT.stream_ptr()))
            gidx = torch.where(index < count, index + begin, torch.full_like(index, torch.iinfo(torch.int64).max))
            _, world = RT.dist_info()
            if world > 1:
                import torch.distributed as dist
                dist.all_reduce(gidx, op=dist.ReduceOp.MIN)
            paths, inverse = torch.unique(gidx, return_inverse=True)
            n_list = int(paths.numel())
            tan = torch.zeros((n_list, n_metric, n_sets, self.nt), dtype=torch.float64, device=dev)
            B.check(L.mcre_irc_set_path_replay(plan, paths.data_ptr(), tan.data_ptr()))
            try:
                slots = L.mcre_irc_main_slots(plan)
                acc = torch.zeros(slots, dtype=torch.float64, device=dev)
                shift = torch.zeros(slots, dtype=torch.float64, device=dev)
                partial = torch.empty(L.mcre_irc_partial_bytes(plan, n_list, CHUNK_PATHS, 0) // 8 + 1, dtype=torch.float64, device=dev)
                rng = self._rng(43, inject, n_main)
                sh = B.Shard(0, n_list, CHUNK_PATHS)
                B.check(L.mcre_irc_mainsim(plan, C.byref(rng), C.byref(sh), partial.data_ptr(), acc.data_ptr(),
                                           shift.data_ptr(), None, RT.stream_ptr()))
            finally:
                B.check(L.mcre_irc_set_path_replay(plan, None, None))
            tan_h, inv_h = tan.cpu().numpy(), inverse.cpu().numpy().reshape(n_sets, n_metric)
            for r in range(n_sets):
                quant[r][q] = (quant[r][q][0], [tan_h[inv_h[r, m], m, r, :] for m in range(n_metric)])

    def run(self):
        c = self.c
        dev = RT.compute_device()
        L = B.lib()
        n_main = c.num_paths_mainsim
        n_sets = len(c.netting_sets)
        need_expo = c.risk_metrics.requires_exposure_profiles()
        timings = {}
        import time
        t0 = time.perf_counter()
        coef_by_product = {}
        berm_coef = {}
        group_size = B.IRC_MAX_SETS if self.nt == 0 else 2
        groups = [list(range(g0, min(g0 + group_size, n_sets))) for g0 in range(0, n_sets, group_size)]
        lowered = {}

        def lower_main():
            for gi, idxs in enumerate(groups):
                lowered[gi] = self.lower(idxs, [])

        dev_reg = None          # (coef_unit, coef_sum, basis, products): regression solved on the device
        events = None
        if c.requires_regression:
            prods = [p for p in c.products if c._product_requires_regression(p) and is_linear(p)]
            if prods and need_expo and self._device_regression_ok(prods, groups):
                # no host synchronisation until the results are read: phases are timed with CUDA events
                events = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
                events[0].record()
                dev_reg = self.presim_on_device(prods, dev, while_running=lower_main) + (prods,)
                events[1].record()
            elif prods and need_expo:
                # the main plans are lowered on the host while the pre-simulation kernels run
                coef_by_product = self.presim_coefficients(prods, dev, while_running=lower_main)
            # expose the coefficients in the reference's raw monomial basis (controller.regression_coeffs)
            for p in prods:
                if id(p) in coef_by_product:
                    coefs, basis, _ = coef_by_product[id(p)]
                    degen = [t <= self.vas.t0() for t in c.exposure_timeline.tolist()]
                    c.regression_coeffs[p.product_id][:, 0, :] = torch.tensor(to_raw_basis(coefs, basis, degen))
            expo_times = c.exposure_timeline.tolist() if need_expo else []
            for p in c.products:
                if not is_rate_bermudan(p):
                    continue
                coef, reg_times, basis, dcoef = self.presim_bermudan(p, dev)
                berm_coef[id(p)] = (coef, {t: k for k, t in enumerate(reg_times)}, basis, dcoef)
                raw = to_raw_basis(coef, basis, [t <= self.vas.t0() for t in reg_times])
                ridx = {t: k for k, t in enumerate(reg_times)}
                for j, t in enumerate(p.regression_timeline.tolist()):
                    p.regression_coeffs[j, 1, :] = torch.tensor(raw[ridx[t]])
                for e, t in enumerate(expo_times):
                    c.regression_coeffs[p.product_id][e, 1, :] = torch.tensor(raw[ridx[t]])
        if dev_reg is None:
            torch.cuda.synchronize(dev)
        timings["preprocessing"] = time.perf_counter() - t0
        t1 = time.perf_counter()

        results = [None] * n_sets
        inject = c.injected_normals.get("main") if c.injected_normals else None
        if not lowered:
            lower_main()
        for gi, idxs in enumerate(groups):
            desc, keep, info = lowered[gi]
            w = 1 + self.nt
            n_expo, n_metric = info["n_expo"], info["n_metric"]
            coef = np.zeros((max(n_expo, 1), len(idxs), 3, w))
            for r, si in enumerate(idxs):
                for p in c.netting_sets[si].products:
                    if id(p) in coef_by_product:
                        coef[:n_expo, r, :, 0] += coef_by_product[id(p)][0]
                        if self.nt and coef_by_product[id(p)][2] is not None:
                            coef[:n_expo, r, :, 1:] += np.transpose(coef_by_product[id(p)][2], (0, 2, 1))
            plan = C.c_void_p()
            B.check(L.mcre_irc_create(C.byref(desc), C.byref(plan)))
            try:
                if dev_reg is not None:
                    B.check(L.mcre_irc_set_coefficients_device(plan, dev_reg[1].data_ptr(), RT.stream_ptr()))
                else:
                    coef_flat, coef_ptr = B.as_dp(coef[:n_expo])
                    B.check(L.mcre_irc_set_coefficients(plan, coef_ptr, RT.stream_ptr()))
                if info["berm_units"]:
                    units = info["berm_units"]
                    exc = np.zeros((max(info["n_ex"], 1), 3, w))
                    bex = np.zeros((max(n_expo, 1), len(units), 3, w))
                    for b, (p, _) in enumerate(units):
                        bc, ridx, _, dbc = berm_coef[id(p)]
                        for i, t in enumerate(p.product_timeline.tolist()):
                            exc[info["ex_index"][(b, i)], :, 0] = bc[ridx[t]]
                        for e, t in enumerate(info["expo_times"]):
                            bex[e, b, :, 0] = bc[ridx[t]]
                            if dbc is not None:      # d(alive-state exposure coefficients)/d(parameters)
                                bex[e, b, :, 1:] = dbc[ridx[t]].T
                    exc_k, exc_ptr = B.as_dp(exc)
                    bex_k, bex_ptr = B.as_dp(bex)
                    B.check(L.mcre_irc_set_exercise_coefficients(plan, exc_ptr, bex_ptr, RT.stream_ptr()))
                chunk_paths = CHUNK_PATHS if self.hybrid is None else self.hybrid["chunk"]
                begin, count = RT.shard_range(n_main, chunk_paths)
                slots = L.mcre_irc_main_slots(plan)
                acc = torch.zeros(slots, dtype=torch.float64, device=dev)
                shift = torch.zeros(slots, dtype=torch.float64, device=dev)
                partial = torch.empty(L.mcre_irc_partial_bytes(plan, max(count, 1), chunk_paths, 0) // 8 + 1,
                                      dtype=torch.float64, device=dev)
                spill = None
                if info["acc"] & B.ACC_SPILL:
                    spill = torch.empty((len(idxs), n_metric, max(count, 1)), dtype=torch.float64, device=dev)
                pv_spill = None
                if self.hybrid is not None and (info["acc"] & B.ACC_PV):
                    pv_spill = torch.zeros((len(idxs), max(count, 1)), dtype=torch.float64, device=dev)
                    B.check(L.mcre_irc_set_pv_spill(plan, pv_spill.data_ptr()))
                rng = self._rng(43, inject, n_main)
                sh = B.Shard(begin, count, chunk_paths)
                B.check(L.mcre_irc_mainsim(plan, C.byref(rng), C.byref(sh), partial.data_ptr(), acc.data_ptr(),
                                           shift.data_ptr(), spill.data_ptr() if spill is not None else None,
                                           RT.stream_ptr()))
                acc = RT.all_reduce_tree(acc)
                if events is not None:
                    events[2].record()
                acc_h = RT.to_host(acc)
                shift_h = RT.to_host(shift)
                quant = None
                cap = None
                if self.hybrid is not None:
                    # per-path exposures / cashflows handed to the combining backend; no metric is finished here
                    cap = dict(idxs=list(idxs), spill=spill, pv_spill=pv_spill, tan=None, pv_tan=None)
                    if self.nt and spill is not None:
                        # tangents of every path's exposures: the replay mode of the tangent build (made for the
                        # gradient of PFE order statistics) over the whole shard, [path][exposure date][set][nt]
                        paths = torch.arange(begin, begin + max(count, 1), dtype=torch.int64, device=dev)
                        cap["tan"] = torch.zeros((max(count, 1), n_metric, len(idxs), self.nt), dtype=torch.float64, device=dev)
                        B.check(L.mcre_irc_set_path_replay(plan, paths.data_ptr(), cap["tan"].data_ptr()))
                        try:
                            acc2, shift2 = torch.zeros_like(acc), torch.zeros_like(shift)
                            sh2 = B.Shard(0, count, chunk_paths)
                            B.check(L.mcre_irc_mainsim(plan, C.byref(rng), C.byref(sh2), partial.data_ptr(), acc2.data_ptr(),
                                                       shift2.data_ptr(), None, RT.stream_ptr()))
                        finally:
                            B.check(L.mcre_irc_set_path_replay(plan, None, None))
                    self.hybrid["captured"].append(cap)
                    spill = None
                if spill is not None:
                    from mcre.select import order_statistics
                    quant = order_statistics(c, spill, count, n_main)
                    if self.nt:
                        self._pfe_sensitivities(plan, quant, spill, begin, count, n_main, n_metric, len(idxs), inject, dev)
            finally:
                L.mcre_irc_destroy(plan)
            ns_t = 1 if len(idxs) <= 1 else (2 if len(idxs) <= 2 else 4)
            nv = 4 + 2 * self.nt
            acc_h = acc_h.reshape(n_metric + 1, ns_t, nv)
            shift_h = shift_h.reshape(n_metric + 1, ns_t, nv)
            if cap is not None and self.nt and (info["acc"] & B.ACC_PV):
                cap["pv_tan"] = [acc_h[n_metric, r, 4:4 + self.nt] / n_main for r in range(len(idxs))]
            from mcre.finish import irc_raw_to_neutral
            for r, si in enumerate(idxs):
                res = irc_raw_to_neutral(acc_h[:, r, :], shift_h[:, r, :], n_main, self.nt, n_metric, info["acc"])
                if quant is not None:
                    res["pfe"] = quant[r]
                res["param_used"] = self.param_used
                results[si] = res
        torch.cuda.synchronize(dev)
        for orig, proxy in getattr(c, "_european_proxies", []):
            # the one state of the European product = the alive state of its proxy
            orig_coeffs = c.regression_coeffs[proxy.product_id]
            if orig_coeffs.shape[0]:
                c._proxy_parent.regression_coeffs[orig.product_id][:, 0, :] = orig_coeffs[:, 1, :]
        timings["path_generation"] = time.perf_counter() - t1
        timings["request_resolution"] = 0.0
        if dev_reg is not None:
            # expose the coefficients in the reference's raw monomial basis (controller.regression_coeffs)
            coef_unit, _, basis, prods = dev_reg
            coef_h = RT.to_host(coef_unit)
            degen = [t <= self.vas.t0() for t in c.exposure_timeline.tolist()]
            for u, p in enumerate(prods):
                c.regression_coeffs[p.product_id][:, 0, :] = torch.tensor(to_raw_basis(coef_h[u], basis, degen))
            # phase split on the device clock (the host never waited between the phases)
            pre_s = events[0].elapsed_time(events[1]) * 1e-3
            main_s = events[1].elapsed_time(events[2]) * 1e-3
            total = timings["preprocessing"] + timings["path_generation"]
            timings["preprocessing"] = pre_s
            timings["path_generation"] = max(total - pre_s, main_s)
        return results, timings
