"""SimulationController with the reference's constructor, attributes and result layout
(reference: src/controller/controller.py:21-709), executing on hand-written sm_100a
kernels through the C-ABI library instead of per-sub-step torch ops.

Flow of run_simulation():
  1. pick the backend (model / product family) and lower the run into flat plan tables
  2. pre-simulation pass (Philox key 42): Gram / right-hand-side moments per regression
     date -> host solve of the (degree+1)^2 normal equations
  3. main pass (Philox key 43): fused path stepping + cashflows + exposure + netting /
     collateral + metric integrands, block-reduced into per-date accumulators
  4. finish metrics on the host from O(T) sums; assemble SimulationResults
"""
from __future__ import annotations

import logging
import os
import time
from collections import defaultdict
from typing import Sequence

import numpy as np

from common.packages import *
from common.enums import SimulationScheme
from controller.simulation_results import SimulationResults
from maths.regression import PolyomialRegression, RegressionFunction
from metrics.metric import Metric, MetricType
from metrics.risk_metrics import RiskMetrics
from models.model import Model
from models.model_config import ModelConfig
from products.netting_set import NettingSet

logger = logging.getLogger(__name__)

_SINGLE_EVAL = {MetricType.PV, MetricType.CVA, MetricType.EEPE, MetricType.CE}


def _merge_grads(a, b):
    """Sum two per-parameter gradient lists where None means "not connected"."""
    if a is None:
        return list(b)
    return [y if x is None else (x if y is None else x + y) for x, y in zip(a, b)]


class SimulationController:
    def __init__(self, netting_sets: Sequence[NettingSet], model: Model, risk_metrics: RiskMetrics,
                 num_paths_mainsim: int, num_paths_presim: int, num_steps: int,
                 simulation_scheme: SimulationScheme, differentiate: bool = False,
                 regression_function: RegressionFunction = PolyomialRegression(degree=2)):
        self.risk_metrics = risk_metrics
        netting_sets = list(netting_sets)
        if len(netting_sets) == 0:
            raise ValueError("Provide at least one netting set.")
        seen = set()
        for ns in netting_sets:
            for p in ns.products:
                if id(p) in seen:
                    raise ValueError("A product instance cannot belong to more than one netting set.")
                seen.add(id(p))
        self.netting_sets = netting_sets
        self.products = [p for ns in netting_sets for p in ns.products]
        self.product_to_netting_set_idx = [i for i, ns in enumerate(netting_sets) for _ in ns.products]

        # exposure grids: metric dates, plus (t - MPoR) look-back dates of collateralised sets
        self.metric_exposure_timeline = risk_metrics.exposure_timeline.clone()
        self.exposure_timeline = self._build_internal_exposure_timeline()
        # (.tolist(): iterating a tensor element by element costs microseconds per element)
        self._exposure_time_to_idx = {t: i for i, t in enumerate(self.exposure_timeline.tolist())}
        self.metric_exposure_indices = torch.tensor(
            [self._exposure_time_to_idx[t] for t in self.metric_exposure_timeline.tolist()],
            dtype=torch.long, device=device)
        self.netting_set_delayed_exposure_indices = self._build_netting_set_delayed_exposure_indices()

        if risk_metrics.any_xva:
            if not isinstance(model, ModelConfig):
                raise Exception("ModelConfig needs to be provided for xVA valuation.")
            if not all(cp in model.id_to_model for cp in risk_metrics.counterparty_ids):
                raise Exception("Not all models set for xVA valuation.")

        self.model = model
        self.num_paths_presim = num_paths_presim
        self.num_paths_mainsim = num_paths_mainsim
        self.num_steps = num_steps
        self.simulation_scheme = simulation_scheme
        self.differentiate = differentiate
        self.regression_function = regression_function
        self.requires_higher_order_derivatives = False
        for pid, prod in enumerate(self.products):
            prod.product_id = pid
        if differentiate:
            self.model.requires_grad()

        # regression coefficients per product: [exposure date, product state, basis]
        # (tensors without elements - products that are never regressed, runs without exposure dates - are shared per
        # shape: books of tens of thousands of products spent 0.3 s here in 100k torch.zeros calls)
        self.regression_coeffs = []
        n_expo, n_basis = len(self.exposure_timeline), regression_function.get_degree()
        empty = {}
        for prod in self.products:
            states = prod.get_num_states()
            if len(prod.regression_timeline) == 0:
                key = (0, states, n_basis)
                if key not in empty:
                    empty[key] = torch.zeros(key, dtype=FLOAT, device=device)
                prod.regression_coeffs = empty[key]
            else:
                prod._allocate_regression_coeffs(regression_function)
            if n_expo == 0:
                key = (0, states, n_basis)
                if key not in empty:
                    empty[key] = torch.zeros(key, dtype=FLOAT, device=device)
                self.regression_coeffs.append(empty[key])
            else:
                self.regression_coeffs.append(torch.zeros((n_expo, states, n_basis), dtype=FLOAT, device=device))

        # simulation grid = product modelling dates U exposure dates, merged on float equality
        times = {t for p in self.products for t in p.modeling_timeline.tolist()}
        times |= set(self.exposure_timeline.tolist())
        self.simulation_timeline = torch.tensor(sorted(times), dtype=FLOAT, device=device)
        self.requires_regression = any(self._product_requires_regression(p) for p in self.products)

        # ---- execution knobs of this implementation (not in the reference) ----
        #: Philox key high word: scenario / sweep index (disjoint streams per scenario)
        self.rng_stream = 0
        #: parity mode: {"pre": Tensor[n_sub, N_pre, d], "main": Tensor[n_sub, N_main, d]} on the GPU
        self.injected_normals = None
        #: "philox" (default: counter-based draws made in registers) or "torch": the reference's own
        #: torch.randn stream (seeds 42 / 43) regenerated on the host and injected (mcre/compat.py)
        self.rng_compat = os.environ.get("MCRE_RNG", "philox").lower()
        #: pre-simulation scratch is bounded by processing this many local paths at a time
        self.presim_batch_paths = 1 << 22
        #: least-squares solver of gas-storage regressions: "auto", "lapack" (the reference's own routine on the host,
        #: reproduces its rank decisions) or "moments" (device moments + small normal equations); mcre/storage.py
        self.storage_regression = "auto"
        self.last_timings = {}

    # ------------------------------------------------------------------ timelines
    def _build_internal_exposure_timeline(self):
        if not self.risk_metrics.requires_exposure_profiles():
            return self.risk_metrics.exposure_timeline.clone()
        times = set(self.risk_metrics.exposure_timeline.tolist())
        for ns in self.netting_sets:
            if ns.is_collateralized():
                times.update(ns.get_collateral_query_times(self.risk_metrics.exposure_timeline).tolist())
        return torch.tensor(sorted(times), dtype=FLOAT, device=device)

    def _build_netting_set_delayed_exposure_indices(self):
        out = []
        n = len(self.metric_exposure_timeline)
        for ns in self.netting_sets:
            idx = torch.full((n,), -1, dtype=torch.long, device=device)
            if ns.is_collateralized():
                delayed = (self.metric_exposure_timeline - ns.margin_period_of_risk).tolist()
                vals = [self._exposure_time_to_idx[d] if d >= 0.0 else -1 for d in delayed]
                idx = torch.tensor(vals, dtype=torch.long, device=device)
            out.append(idx)
        return out

    @staticmethod
    def _make_unique_names(base_names):
        counts, out = defaultdict(int), []
        for name in base_names:
            counts[name] += 1
            out.append(name if counts[name] == 1 else f"{name}#{counts[name]}")
        return out

    # ------------------------------------------------------------------ dispatch rules
    def _product_requires_regression(self, product):
        if len(product.regression_timeline) > 0:
            return True
        if not self.risk_metrics.requires_exposure_profiles():
            return False
        return not self._can_use_analytic_exposure_for_product(product)

    def _can_use_analytic_exposure_for_product(self, product):
        ok = {MetricType.PV, MetricType.EPE, MetricType.PFE}
        return all(m.metric_type in ok for m in self.risk_metrics.metrics) and product.supports_analytic_exposure(self.model)

    def _can_evaluate_metric_analytically_for_product(self, product, metric):
        return (metric.metric_type == MetricType.PV
                and metric.evaluation_type == Metric.EvaluationType.ANALYTICAL
                and product.supports_analytic_pv(self.model))

    def _can_skip_monte_carlo_for_product(self, product):
        if self.risk_metrics.requires_exposure_profiles():
            return False
        if not any(m.evaluation_type == Metric.EvaluationType.ANALYTICAL for m in self.risk_metrics.metrics):
            return False        # (asked several times per product of a run: the common case leaves here)
        return all(self._can_evaluate_metric_analytically_for_product(product, m) for m in self.risk_metrics.metrics)

    def compute_higher_derivatives(self):
        self.requires_higher_order_derivatives = True

    def inject_normals(self, pre=None, main=None):
        """Parity mode: use the given standard-normal draws
        [n_sub, N, noise_dim] (the reference's torch.randn stream) instead of Philox."""
        from mcre.runtime import compute_device
        dev = compute_device()
        self.injected_normals = {}
        if pre is not None:
            self.injected_normals["pre"] = torch.as_tensor(pre, dtype=torch.float64).to(dev).contiguous()
        if main is not None:
            self.injected_normals["main"] = torch.as_tensor(main, dtype=torch.float64).to(dev).contiguous()

    # ------------------------------------------------------------------ run
    def _select_backend(self):
        from mcre.storage import StorageBackend
        if StorageBackend.supports(self):
            return StorageBackend(self)
        from mcre.irc import IrcBackend, with_single_exercise_proxies
        view = with_single_exercise_proxies(self)
        if view is not self:
            view._proxy_parent = self
        if IrcBackend.supports(view):
            return IrcBackend(view)
        from mcre.equity import EquityBackend
        from mcre.hybrid import EquityCreditGreeks
        if EquityCreditGreeks.supports(self):
            return EquityCreditGreeks(self)
        if EquityBackend.supports(self):
            return EquityBackend(self)
        from mcre.hybrid import HybridBackend
        if HybridBackend.supports(self):
            return HybridBackend(self)
        raise NotImplementedError(
            f"No CUDA backend for model {type(self.model).__name__} with products "
            f"{sorted({type(p).__name__ for p in self.products})}")

    def _analytic_pv(self, product):
        """Closed-form PV of a product the Monte Carlo can skip (controller.py:204-229) and, with
        differentiate=True, its first (and on request second) derivatives with respect to the model
        parameters: a scalar closed form evaluated on the host with torch autograd, like the reference."""
        params = self.model.get_model_params()
        if not self.differentiate:
            return float(product.compute_pv_analytically(self.model).squeeze()), None, None
        for q in params:
            q.requires_grad_(True)
        try:
            pv = product.compute_pv_analytically(self.model).squeeze()
            second = self.requires_higher_order_derivatives
            g = torch.autograd.grad(pv, params, retain_graph=second, create_graph=second, allow_unused=True)
            hess = None
            if second:
                hess = []
                for gi in g:
                    if gi is None or not gi.requires_grad:
                        hess.append(tuple(None for _ in params))
                        continue
                    row = torch.autograd.grad(gi, params, retain_graph=True, allow_unused=True)
                    hess.append(tuple(None if h is None else h.detach().cpu().numpy() for h in row))
            grads = [None if gi is None else gi.detach().cpu().numpy() for gi in g]
            return float(pv.detach()), grads, hess
        finally:
            for q in params:
                q.requires_grad_(False)

    def run_simulation(self) -> SimulationResults:
        t0 = time.perf_counter()
        mc_products = [p for p in self.products if not self._can_skip_monte_carlo_for_product(p)]
        from products.storage import Storage
        if any(self._product_requires_regression(p) and not isinstance(p, Storage) for p in mc_products):
            # (gas storages regress on any polynomial degree: mcre/storage.py)
            # the pre-simulation kernels accumulate the moments of the quadratic basis [1, x, x^2] (8 sums, 3x3 normal
            # equations) and the main kernels evaluate 3 coefficients: any other basis must not be evaluated as this one
            from maths.regression import PolyomialRegression
            rf = self.regression_function
            if type(rf) is not PolyomialRegression or rf.degree != 2:
                raise NotImplementedError(
                    "regression_function: only PolyomialRegression(degree=2) is implemented by the CUDA regression "
                    f"kernels (got {type(rf).__name__}(degree={getattr(rf, 'degree', None)})); the reference builds "
                    "regression_function.get_regression_matrix(x) for any basis (controller.py:361-374)")
        n_params = len(self.model.get_model_params())
        analytic = [[0.0 for _ in self.risk_metrics.metrics] for _ in self.netting_sets]
        analytic_grads = [[None for _ in self.risk_metrics.metrics] for _ in self.netting_sets]
        analytic_hess = [[None for _ in self.risk_metrics.metrics] for _ in self.netting_sets]
        has_pathwise = [False] * len(self.netting_sets)
        for pi, p in enumerate(self.products):
            si = self.product_to_netting_set_idx[pi]
            if self._can_skip_monte_carlo_for_product(p):
                pv, g, h = self._analytic_pv(p)
                for mi, _ in enumerate(self.risk_metrics.metrics):
                    analytic[si][mi] += pv
                    if g is not None:
                        analytic_grads[si][mi] = _merge_grads(analytic_grads[si][mi], g)
                    if h is not None:
                        cur = analytic_hess[si][mi]
                        analytic_hess[si][mi] = h if cur is None else [tuple(_merge_grads(list(a), list(b))) for a, b in zip(cur, h)]
            else:
                has_pathwise[si] = True
        self._analytic_grads = analytic_grads
        if mc_products and self.injected_normals is None and self.rng_compat == "torch":
            from mcre.compat import inject_reference_stream
            inject_reference_stream(self)
        raw, timings = None, {"preprocessing": 0.0, "path_generation": 0.0, "request_resolution": 0.0}
        if mc_products:
            backend = self._select_backend()
            if self.requires_higher_order_derivatives and not getattr(backend, "second", False):
                raise NotImplementedError("second-order sensitivities of Monte Carlo values: present values of equity "
                                          "books on one Black-Scholes model (mcre/equity.py); analytic PVs otherwise")
            raw, timings = backend.run()
        t3 = time.perf_counter()
        from mcre.finish import finish_results
        results, grads = finish_results(self, raw, analytic, has_pathwise)
        t4 = time.perf_counter()
        timings["valuation"] = t4 - t3
        timings["total"] = t4 - t0
        self.last_timings = timings
        logger.info(
            "Simulation completed for %d netting set(s) and %d product(s): "
            "preprocessing=%.6fs path_generation=%.6fs request_resolution=%.6fs valuation=%.6fs total=%.6fs",
            len(self.netting_sets), len(self.products), timings["preprocessing"], timings["path_generation"],
            timings["request_resolution"], timings["valuation"], timings["total"])
        higher = []
        if self.requires_higher_order_derivatives:
            # [set][metric][evaluation][parameter i] -> tuple over parameters (simulation_results.py:5-338)
            def hessian_of(si, mi):
                h = analytic_hess[si][mi]
                mc = raw[si].get("pv_hess") if (raw is not None and raw[si] is not None) else None
                if mc is not None and self.risk_metrics.metrics[mi].metric_type == MetricType.PV:
                    # pathwise Hessian of the simulated part (mcre/equity.py, csrc/dual2.cuh) + the analytic part
                    rows = [tuple(float(x) for x in row) for row in mc]
                    h = rows if h is None else [tuple(_merge_grads(list(a), list(b))) for a, b in zip(h, rows)]
                return h if h is not None else [tuple(None for _ in range(n_params)) for _ in range(n_params)]
            higher = [[[hessian_of(si, mi)] for mi in range(len(self.risk_metrics.metrics))]
                      for si in range(len(self.netting_sets))]
        return SimulationResults(
            results, grads, higher,
            netting_set_names=self._make_unique_names([ns.get_name() for ns in self.netting_sets]),
            metric_names=self._make_unique_names([m.get_name() for m in self.risk_metrics.metrics]),
            model_param_names=self.model.get_model_param_names())
