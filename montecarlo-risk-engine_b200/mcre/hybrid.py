"""Netting sets that mix equity and interest-rate products under one ModelConfig of Black-Scholes market models
(the numeraire), a Vasicek short rate and optionally the counterparty's CIR++ intensity - the reference's
tests/exposure_tests/cva_large_netting_set_derivatives.py (model_config.py:8-276; the controller sums all products of a
set before the netting terms, controller.py:506-563).

The two product families have their own fused kernels.  Here they run on column slices of ONE joint correlated draw:

  1. mcre_correlated_normals materialises w = L z per (sub-step, path) for the pre-simulation (key 42) and the main
     simulation (key 43): L = Cholesky factor of the joint correlation, z = Philox or the injected reference stream;
  2. the rate products run through the interest-rate backend on the Vasicek column with the numeraire of the
     Black-Scholes model (`ext_numeraire`), spilling per-path netted exposures [exposure date][path] and discounted
     cashflows instead of finishing metrics;
  3. the equity products run through the equity backend's accumulating launches on the equity (+ credit) columns,
     starting from those spills: threshold / MPoR collateral, positive parts, PFE order statistics and the CVA integrand
     then see the whole netting set.

Sensitivities (differentiate=True): a second pass on plans with tangents leaves per-path tangents of both families'
exposures (interest-rate kernels: the path-replay spill; equity kernel: mcre_eq_set_exposure_tangent_accumulator), and
mcre_exposure_tangent_sums differentiates the netting terms and positive parts on the combined duals.
The noise tensor is O(sub-steps x paths x factors): these books run on thousands to a few million paths.
"""
from __future__ import annotations

import copy
import ctypes as C
import time

import numpy as np
import torch

from common.enums import SimulationScheme
from mcre import binding as B
from mcre import runtime as RT
from mcre.timegrid import build_time_grid

#: largest joint noise tensor (bytes) this backend materialises per simulation pass
MAX_NOISE_BYTES = 24 << 30


def _families(model):
    """(black-scholes indices, vasicek index, credit index or None) of a hybrid ModelConfig, or None."""
    from models.black_scholes import BlackScholesModel
    from models.cirpp import CIRPPModel
    from models.model_config import ModelConfig
    from models.vasicek import VasicekModel
    if not isinstance(model, ModelConfig):
        return None
    bs = [i for i, m in enumerate(model.models) if type(m) is BlackScholesModel]
    vas = [i for i, m in enumerate(model.models) if isinstance(m, VasicekModel)]
    cir = [i for i, m in enumerate(model.models) if isinstance(m, CIRPPModel)]
    if not bs or len(vas) != 1 or len(cir) > 1 or len(bs) + len(vas) + len(cir) != len(model.models):
        return None
    if model.id_to_model["numeraire"] not in bs:
        return None
    return bs, vas[0], (cir[0] if cir else None)


def metric_gradients(c, values, tan, pv_grad, chunk, dev, credit=None):
    """Netting-set terms and metric parts on per-path duals (mcre_exposure_tangent_sums): `values[si]` = the value run's
    dict of set si (exposures before netting terms "_expo", default weights "_cva_w"), tan[si] [parameters][exposure
    date][path] the tangents of those exposures, pv_grad[si] the PV gradient.
    `credit` (stochastic intensity): {"w_tan": [n_metric][4][n] tangents of the per-path default weights w.r.t. the
    credit model's parameters, "offset": index of its first parameter}; the weights themselves are the value run's.
    -> per set {"pv": grad, "pos": [grad per metric date], "neg": [...], "cva": grad} over the flattened parameters."""
    from metrics.metric import MetricType
    L = B.lib()
    n_par = len(c.model.model_params)
    n_main = c.num_paths_mainsim
    n_expo, n_metric = len(c.exposure_timeline), len(c.metric_exposure_timeline)
    need_expo = c.risk_metrics.requires_exposure_profiles()
    begin, count = RT.shard_range(n_main, chunk)
    n = max(count, 1)
    out = []
    cva_metric = next((m for m in c.risk_metrics.metrics if m.metric_type == MetricType.CVA), None)
    metric_expo = np.asarray(c.metric_exposure_indices.tolist(), dtype=np.int32)
    for si, ns in enumerate(c.netting_sets):
        res = {"pv": pv_grad[si]}
        if need_expo:
            lag = np.full(n_metric, -1, dtype=np.int32)
            if ns.is_collateralized():
                delayed = c.netting_set_delayed_exposure_indices[si].tolist()
                for m in range(n_metric):
                    if delayed[m] >= 0:
                        lag[m] = int(metric_expo[m]) - delayed[m]
            w = np.zeros(n_metric)
            cva_w = values[si].get("_cva_w")
            per_path = credit is not None and cva_metric is not None and cva_w is not None
            if cva_metric is not None and cva_w is not None and not per_path:
                # deterministic credit: the same weights on every path; every rank reads its first local path
                w_local = cva_w[:, 0].clone() if count > 0 else torch.zeros(n_metric, dtype=torch.float64, device=dev)
                w = RT.to_host(w_local) * (1.0 - cva_metric.recovery_rate)
            n_wtan = 4 if per_path else 0
            n_chunks = (n + chunk - 1) // chunk
            slots = n_metric * (n_par + n_wtan) * 3
            partial = torch.empty(n_chunks * slots + 1, dtype=torch.float64, device=dev)
            sums = torch.zeros(slots, dtype=torch.float64, device=dev)
            me_k, me_p = B.as_ip(metric_expo)
            lg_k, lg_p = B.as_ip(lag)
            w_k, w_p = B.as_dp(w)
            if per_path:
                # stochastic intensity: per-path weights; the credit model's own parameters enter through the weights
                # (rows n_par .. n_par + 3 of the result)
                B.check(L.mcre_exposure_tangent_sums_paths(
                    values[si]["_expo"].data_ptr(), tan[si].data_ptr(), count, n_expo, n_par, n_metric, me_p, lg_p,
                    int(ns.is_collateralized()), float(ns.threshold), cva_w.data_ptr(), 1.0 - cva_metric.recovery_rate,
                    credit["w_tan"].data_ptr(), n_wtan, chunk, partial.data_ptr(), sums.data_ptr(), RT.stream_ptr()))
            else:
                B.check(L.mcre_exposure_tangent_sums(values[si]["_expo"].data_ptr(), tan[si].data_ptr(), count, n_expo, n_par,
                                                     n_metric, me_p, lg_p, int(ns.is_collateralized()), float(ns.threshold),
                                                     w_p, chunk, partial.data_ptr(), sums.data_ptr(), RT.stream_ptr()))
            sm = RT.to_host(RT.all_reduce_tree(sums)).reshape(n_metric, n_par + n_wtan, 3) / n_main
            res["pos"] = [sm[m, :n_par, 0] for m in range(n_metric)]
            res["neg"] = [sm[m, :n_par, 1] for m in range(n_metric)]
            res["cva"] = sm[:n_metric - 1, :n_par, 2].sum(axis=0) if n_metric > 1 else np.zeros(n_par)
            if per_path and n_metric > 1:
                off = credit["offset"]
                res["cva"][off:off + 4] += sm[:n_metric - 1, n_par:, 2].sum(axis=0)
            # PFE: the reference differentiates torch.sort(...)[index] (pfe_metric.py:59-71) = the pathwise gradient of the
            # selected path.  The path is located in the value run's unsecured exposures (smallest global id among ties,
            # like the interest-rate backend), its dual read from the per-path tensors on the rank that owns it.
            res["pfe"] = {}
            for q, (vals, _) in (values[si].get("pfe") or {}).items():
                unsec = values[si]["_unsec"]                                       # [1][n_metric][n]
                targets = torch.tensor([v for v, _ in vals], dtype=torch.float64, device=dev)
                index = torch.empty(n_metric, dtype=torch.int64, device=dev)
                B.check(L.mcre_select_locate(unsec.data_ptr(), unsec.stride(1), count, n_metric, targets.data_ptr(),
                                             index.data_ptr(), RT.stream_ptr()))
                gidx = torch.where(index < count, index + begin, torch.full_like(index, torch.iinfo(torch.int64).max))
                _, world = RT.dist_info()
                if world > 1:
                    import torch.distributed as dist
                    dist.all_reduce(gidx, op=dist.ReduceOp.MIN)
                expo_h = values[si]["_expo"]
                grad = torch.zeros((n_metric, n_par), dtype=torch.float64, device=dev)
                g_host = gidx.tolist()
                h = float(ns.threshold)
                for m in range(n_metric):
                    lp = g_host[m] - begin
                    if not (0 <= lp < count):
                        continue                                                   # another rank owns the path
                    xe = int(metric_expo[m])

                    def dthr(x):
                        return 1.0 if (x > h or x < -h) else 0.0
                    if ns.is_collateralized():
                        g = tan[si][:, xe, lp].clone()
                        if lag[m] >= 0:
                            g -= dthr(float(expo_h[xe - lag[m], lp])) * tan[si][:, xe - lag[m], lp]
                    else:
                        g = dthr(float(expo_h[xe, lp])) * tan[si][:, xe, lp]
                    grad[m] = g
                grad_h = RT.to_host(RT.all_reduce_tree(grad)) + 0.0     # (-0.0 of a masked tangent -> 0.0, as the sum over ranks gives)
                res["pfe"][q] = [grad_h[m] for m in range(n_metric)]
        out.append(res)
    return out


def attach_gradients(results, grads, used, used_cva=None):
    """Puts the gradients of metric_gradients next to the values of the value run's per-set dicts.  used / used_cva:
    which parameters the metrics / the CVA are connected to in the reference's autograd graph."""
    from metrics.metric import MetricType
    for res, g in zip(results, grads):
        res["pv"] = (res["pv"][0], g["pv"])
        if "pos" in res:
            res["pos"] = (res["pos"][0], g["pos"])
        if "neg" in res:
            res["neg"] = (res["neg"][0], g["neg"])
        if "cva" in res and res["cva"][0] != (0.0, 0.0):
            res["cva"] = (res["cva"][0], g["cva"])
        for q, rows in (g.get("pfe") or {}).items():
            res["pfe"][q] = (res["pfe"][q][0], rows)
        res["param_used"] = lambda kind, used=used, uc=used_cva: (uc if (uc is not None and kind == MetricType.CVA) else used)


def _credit_checks(c, cir, what):
    from metrics.metric import MetricType
    if cir is not None and not cir.deterministic and any(m.metric_type == MetricType.CVA for m in c.risk_metrics.metrics):
        raise NotImplementedError(f"sensitivities of the CVA of {what}: deterministic credit (the default weights of a "
                                  "stochastic intensity carry tangents the equity launch does not have)")


class EquityCreditGreeks:
    """differentiate=True on an equity book whose ModelConfig carries the counterparty's CIR++ model (CVA of equity
    books, tests/exposure_tests/cva_perfprmance_large_netting_set.py with sensitivities): the value run is the equity
    backend's accumulating path; a second pass on a plan with tangents - the credit factor rides along value-only, so the
    joint draw and the paths are the same - leaves per-path exposure tangents, and metric_gradients differentiates netting
    terms, positive parts and the CVA integrand.  The market parameters' sensitivities are complete; the credit model's
    come back as None in its deterministic mode (outside the reference's autograd graph)."""

    @staticmethod
    def supports(ctrl):
        """Equity books whose sensitivities need per-path duals: a credit model in the ModelConfig, a PFE metric (the
        gradient of an order statistic is one path's tangent; the fused kernel only has tangent sums), or a netting set
        split over several launches."""
        from metrics.metric import MetricType
        from mcre.equity import EQ_BS, EquityBackend, _is_path_dependent, credit_of, eq_ntrk, family_of
        if not ctrl.differentiate or not EquityBackend.supports(ctrl):
            return False
        if credit_of(ctrl.model)[0] is not None:
            return True
        fam = family_of(ctrl.model)
        if fam is None or fam[0] != EQ_BS or not ctrl.risk_metrics.requires_exposure_profiles():
            return False
        if any(m.metric_type == MetricType.PFE for m in ctrl.risk_metrics.metrics):
            return True
        # exposure sensitivities of a netting set with more path-dependent products than one launch with tangents tracks:
        # the book is split over launches, whose per-path exposures and duals are netted here
        return any(sum(_is_path_dependent(p) for p in ns.products) > eq_ntrk(3) for ns in ctrl.netting_sets)

    def __init__(self, ctrl):
        from mcre.equity import credit_of
        self.c = ctrl
        self.credit, self.credit_idx = credit_of(ctrl.model)
        if self.credit is not None and ctrl.simulation_scheme != SimulationScheme.EULER:
            # same restriction as the reference: only Black-Scholes pairs have a joint exact covariance (model_config.py:216-221)
            raise NotImplementedError("Analytical covariance is currently only supported for Black-Scholes-type model pairs.")

    def _view(self, differentiate):
        c = self.c
        v = copy.copy(c)
        v.differentiate = differentiate
        v._credit_passenger = differentiate
        v.regression_coeffs = list(c.regression_coeffs) if not differentiate else [rc.clone() for rc in c.regression_coeffs]
        return v

    def run(self):
        from mcre.equity import EquityBackend, is_equity_exercise, main_chunk
        c = self.c
        dev = RT.compute_device()
        t_start = time.perf_counter()
        n_main, n_sets = c.num_paths_mainsim, len(c.netting_sets)
        n_par = len(c.model.model_params)
        n_expo = len(c.exposure_timeline)
        need_expo = c.risk_metrics.requires_exposure_profiles()
        chunk = main_chunk(n_main)
        begin, count = RT.shard_range(n_main, chunk)
        n = max(count, 1)

        def presim(eb, view):
            eb.presim_exercise_all([p for p in view.products if is_equity_exercise(p)], dev)
            if need_expo:
                reg = [p for p in view.products if not view._can_use_analytic_exposure_for_product(p) and not is_equity_exercise(p)]
                if reg:
                    eb.presim_regression(reg, dev)
        view = self._view(False)
        eb = EquityBackend(view)
        presim(eb, view)
        results = [eb._run_split_book(si, dev, n_main, n_par, chunk=chunk) for si in range(n_sets)]
        tview = self._view(True)
        et = EquityBackend(tview)
        presim(et, tview)
        tan = [torch.zeros((n_par, n_expo, n), dtype=torch.float64, device=dev) if need_expo else None for _ in range(n_sets)]
        pv_grad = []
        from metrics.metric import MetricType
        stochastic_cva = (self.credit is not None and not self.credit.deterministic and need_expo
                          and any(m.metric_type == MetricType.CVA for m in c.risk_metrics.metrics))
        credit_out = {} if stochastic_cva else None
        for si in range(n_sets):
            accum_t, g_pv = et.exposure_tangent_pass(si, dev, n_main, chunk, credit_out=credit_out)
            pv_grad.append(g_pv)
            if need_expo:
                for a, asset in enumerate(et.assets):
                    for k, g in enumerate(asset.gmap):
                        tan[si][g] += accum_t[:, a, k, :]
        credit = None
        if stochastic_cva and "w_tan" in credit_out:
            credit = {"w_tan": credit_out["w_tan"], "offset": c.model.param_offsets()[self.credit_idx]}
        grads = metric_gradients(c, results, tan, pv_grad, chunk, dev, credit=credit)
        used = [True] * n_par
        used_cva = list(used)
        if self.credit is not None and self.credit.deterministic:
            # the deterministic intensity is outside the reference's autograd graph (None); a stochastic one rides in the
            # joint state tensor: its parameters come back as 0.0 for the metrics they do not move
            offs = c.model.param_offsets()
            for k in range(len(self.credit.model_params)):
                used[offs[self.credit_idx] + k] = False
                used_cva[offs[self.credit_idx] + k] = False
        attach_gradients(results, grads, used, used_cva)
        torch.cuda.synchronize(dev)
        total = time.perf_counter() - t_start
        return results, {"preprocessing": 0.0, "path_generation": total, "request_resolution": 0.0}


class HybridBackend:
    @staticmethod
    def supports(ctrl):
        return _families(ctrl.model) is not None

    def __init__(self, ctrl):
        from mcre.irc import is_linear
        self.c = c = ctrl
        self.differentiate = bool(c.differentiate)
        self.bs_idx, self.vas_idx, self.cir_idx = _families(c.model)
        if c.differentiate:
            from metrics.metric import MetricType
            cir = c.model.models[self.cir_idx] if self.cir_idx is not None else None
            if len(self.bs_idx) != 1:
                raise NotImplementedError("sensitivities of hybrid books: one Black-Scholes market model")
        if c.simulation_scheme != SimulationScheme.EULER:
            # the reference defines inter-model covariances for Black-Scholes pairs only (model_config.py:201-221)
            raise NotImplementedError("Inter covariance not implemented for the requested pair of models.")
        models = c.model.models
        vas_assets = set(models[self.vas_idx].asset_ids)
        eq_assets = {a for i in self.bs_idx for a in models[i].asset_ids}
        self.rate_products, self.equity_products = [], []
        for p in c.products:
            a = p.get_asset_id()
            if a in vas_assets:
                if not is_linear(p):
                    raise NotImplementedError("hybrid books: bonds and swaps on the short-rate model")
                self.rate_products.append(p)
            elif a in eq_assets:
                self.equity_products.append(p)
            else:
                raise NotImplementedError(f"hybrid books: no market model for asset {a!r}")

    # ------------------------------------------------------------------ joint noise
    def _joint_noise(self, which, seed, n_total, n_sub, dev):
        """-> device tensor [n_sub][n_total][dim] of correlated draws."""
        from mcre.dual import D, cholesky_dual
        from mcre.paths import joint_matrix
        c = self.c
        dim = int(c.model.simulation_dim)
        if n_sub * n_total * dim * 8 > MAX_NOISE_BYTES:
            raise NotImplementedError(f"hybrid books materialise the joint noise: {n_sub} sub-steps x {n_total} paths x "
                                      f"{dim} factors exceeds {MAX_NOISE_BYTES >> 30} GB")
        corr = joint_matrix(c.model, c.simulation_scheme, None)
        Lm = cholesky_dual([[D(x, None, 0) for x in row] for row in corr])
        chol, chol_p = B.as_dp(np.array([[x.v for x in row] for row in Lm]))
        rng = B.Rng()
        rng.seed, rng.stream, rng.n_paths_total = seed, c.rng_stream, n_total
        z = c.injected_normals.get(which) if c.injected_normals else None
        if z is not None:
            if tuple(z.shape) != (n_sub, n_total, dim):
                raise ValueError(f"injected {which} normals must be [{n_sub}, {n_total}, {dim}]")
            rng.mode, rng.d_z = B.RNG_INJECT, z.data_ptr()
        else:
            rng.mode = B.RNG_PHILOX
        out = torch.empty((max(n_sub, 1), n_total, dim), dtype=torch.float64, device=dev)
        B.check(B.lib().mcre_correlated_normals(C.byref(rng), n_sub, dim, chol_p, n_total, out.data_ptr(), RT.stream_ptr()))
        return out

    def _columns(self):
        """Noise columns of (equity models, vasicek, credit) in the joint draw."""
        off, cols = 0, []
        for m in self.c.model.models:
            cols.append(off)
            off += m.simulation_dim
        return [cols[i] for i in self.bs_idx], cols[self.vas_idx], (cols[self.cir_idx] if self.cir_idx is not None else None)

    # ------------------------------------------------------------------ sub-controllers
    def _sub(self, model, netting_sets, risk_metrics=None, differentiate=False):
        c = self.c
        sub = copy.copy(c)
        sub.model = model
        sub.netting_sets = netting_sets
        sub.products = [p for ns in netting_sets for p in ns.products]
        sub.product_to_netting_set_idx = [i for i, ns in enumerate(netting_sets) for _ in ns.products]
        if risk_metrics is not None:
            n = len(c.exposure_timeline)
            sub.risk_metrics = risk_metrics
            sub.metric_exposure_timeline = c.exposure_timeline.clone()
            sub.metric_exposure_indices = torch.arange(n, dtype=torch.long)
            sub.netting_set_delayed_exposure_indices = [torch.full((n,), -1, dtype=torch.long) for _ in netting_sets]
        sub.requires_regression = any(sub._product_requires_regression(p) for p in sub.products)
        sub.differentiate = differentiate
        sub._credit_passenger = differentiate     # per-path duals: the metrics are differentiated in metric_gradients
        sub.injected_normals = {}
        sub.last_timings = {}
        return sub

    # ------------------------------------------------------------------ sensitivities
    def _tangent_pass(self, noise_eq, noise_ir, values, chunk, dev):
        """First-order sensitivities of every metric (the reference: torch.autograd through the whole run,
        controller.py:609-627).  Both families run once more on plans with tangents and leave per-path tangents of their
        netted exposures; mcre_exposure_tangent_sums applies the netting-set terms and the positive / negative parts to
        the combined duals.  `values`: the value run's per-set dicts (exposures before netting terms, default weights).
        -> per set {"pv": grad, "pos": [grad per metric date], "neg": [...], "cva": grad} over the flattened parameters."""
        from metrics.metric import MetricType
        from metrics.pfe_metric import PFEMetric
        from metrics.pv_metric import PVMetric
        from metrics.risk_metrics import RiskMetrics
        from mcre.equity import EquityBackend, is_equity_exercise
        from mcre.irc import IrcBackend
        from products.netting_set import NettingSet
        c = self.c
        L = B.lib()
        models = c.model.models
        offs = c.model.param_offsets()
        n_par = len(c.model.model_params)
        n_main = c.num_paths_mainsim
        n_sets = len(c.netting_sets)
        n_expo, n_metric = len(c.exposure_timeline), len(c.metric_exposure_timeline)
        need_expo = c.risk_metrics.requires_exposure_profiles()
        need_pv = c.risk_metrics.requires_discounted_cashflows()
        num_rate = offs[c.model.id_to_model["numeraire"]] + 2
        begin, count = RT.shard_range(n_main, chunk)
        n = max(count, 1)
        # per-path tangents of the netted exposure of every set with respect to the flattened parameters
        tan = [torch.zeros((n_par, n_expo, n), dtype=torch.float64, device=dev) if need_expo else None for _ in range(n_sets)]
        pv_grad = [np.zeros(n_par) for _ in range(n_sets)]

        # ---- rate products: tangent build with 8 slots = 4 Vasicek parameters, the numeraire model's rate, 3 unused
        rate_of_set = [[p for p in ns.products if p in self.rate_products] for ns in c.netting_sets]
        if self.rate_products:
            rows = [si for si in range(n_sets) if rate_of_set[si]]
            ir_sets = [NettingSet(name=f"rates_{si}", products=rate_of_set[si]) for si in rows]
            ir_metrics = ([PFEMetric(0.5)] if need_expo else []) + ([PVMetric()] if need_pv or not need_expo else [])
            rm = RiskMetrics(ir_metrics, exposure_timeline=c.exposure_timeline.tolist() if need_expo else None)
            sub = self._sub(models[self.vas_idx], ir_sets, rm, differentiate=True)
            sub.injected_normals = noise_ir
            sub.regression_coeffs = [rc.clone() for rc in c.regression_coeffs]     # (the value run's stay as they are)
            ib = IrcBackend(sub)
            ib.nt = 8
            num_model = models[c.model.id_to_model["numeraire"]]
            ib.hybrid = {"ext_rate": float(num_model.param_values()[2]), "ext_slot": 4, "chunk": chunk, "captured": []}
            ib.run()
            slot_to_global = [offs[self.vas_idx] + k for k in range(4)] + [num_rate]
            for cap in ib.hybrid["captured"]:
                for r, k in enumerate(cap["idxs"]):
                    si = rows[k]
                    if cap["tan"] is not None:
                        t = cap["tan"][:, :, r, :].permute(2, 1, 0)          # [path][date][slot] -> [slot][date][path]
                        for slot, g in enumerate(slot_to_global):
                            tan[si][g] += t[slot]
                    if cap["pv_tan"] is not None:
                        for slot, g in enumerate(slot_to_global):
                            pv_grad[si][g] += cap["pv_tan"][r][slot]

        # ---- equity products: Black-Scholes plan with lane-local tangents (spot, volatility, rate) ----------------
        b = self.bs_idx[0]
        eq_sets = []
        for ns in c.netting_sets:
            view = copy.copy(ns)
            view.products = [p for p in ns.products if p in self.equity_products]
            eq_sets.append(view)
        cir = models[self.cir_idx] if self.cir_idx is not None else None
        stochastic_cva = (cir is not None and not cir.deterministic and need_expo
                          and any(m.metric_type == MetricType.CVA for m in c.risk_metrics.metrics))
        if stochastic_cva:
            # the credit factor as a passenger of the equity plan (the value run's layout: market model, then credit):
            # its column of the joint draw feeds the tangents of the per-path default weights
            from models.model_config import ModelConfig
            sub = self._sub(ModelConfig(models=[models[b], cir]), eq_sets, differentiate=True)
            sub.injected_normals = dict(noise_eq)
        else:
            sub = self._sub(models[b], eq_sets, differentiate=True)
            sub.injected_normals = {k: v[:, :, :1].contiguous() for k, v in noise_eq.items()}
        credit_out = {} if stochastic_cva else None
        sub.regression_coeffs = [rc.clone() for rc in c.regression_coeffs]
        eb = EquityBackend(sub)
        eb.presim_exercise_all([p for p in sub.products if is_equity_exercise(p)], dev)
        if need_expo:
            reg = [p for p in sub.products if not sub._can_use_analytic_exposure_for_product(p) and not is_equity_exercise(p)]
            if reg:
                eb.presim_regression(reg, dev)
        for si in range(n_sets):
            if not eq_sets[si].products:
                continue
            accum_t, g_pv = eb.exposure_tangent_pass(si, dev, n_main, chunk, credit_out=credit_out)
            for k in range(3):
                if need_expo:
                    tan[si][offs[b] + k] += accum_t[:, 0, k, :]
                pv_grad[si][offs[b] + k] += g_pv[k]

        credit = None
        if stochastic_cva:
            if "w_tan" not in credit_out:
                raise NotImplementedError("sensitivities of the CVA of hybrid books under a stochastic intensity: a netting "
                                          "set with equity products (the weights' tangents ride with an equity launch)")
            credit = {"w_tan": credit_out["w_tan"], "offset": offs[self.cir_idx]}
        return metric_gradients(c, values, tan, pv_grad, chunk, dev, credit=credit)

    def run(self):
        from metrics.metric import MetricType
        from metrics.pfe_metric import PFEMetric
        from metrics.pv_metric import PVMetric
        from metrics.risk_metrics import RiskMetrics
        from models.model_config import ModelConfig
        from mcre.equity import EquityBackend
        from mcre.irc import IrcBackend
        from products.netting_set import NettingSet
        c = self.c
        dev = RT.compute_device()
        t_start = time.perf_counter()
        models = c.model.models
        n_main, n_pre = c.num_paths_mainsim, c.num_paths_presim
        t0 = float(models[0].calibration_date[0])
        n_sub = build_time_grid(t0, c.simulation_timeline.tolist(), c.num_steps).n_sub
        need_expo = c.risk_metrics.requires_exposure_profiles()
        need_pv = c.risk_metrics.requires_discounted_cashflows()
        kinds = {m.metric_type for m in c.risk_metrics.metrics}
        eq_cols, vas_col, cir_col = self._columns()
        with_credit = MetricType.CVA in kinds and self.cir_idx is not None
        chunk = 256 if n_main < (1 << 18) else 4096

        # ---- one joint draw per pass, sliced per family (contiguous copies: the kernels index [sub-step][path][column])
        w_main = self._joint_noise("main", 43, n_main, n_sub, dev)
        eq_take = eq_cols + ([cir_col] if with_credit else [])
        noise_eq = {"main": w_main[:, :, eq_take].contiguous()}
        noise_ir = {"main": w_main[:, :, [vas_col]].contiguous()}
        del w_main
        if c.requires_regression and n_pre > 0:
            w_pre = self._joint_noise("pre", 42, n_pre, n_sub, dev)
            noise_eq["pre"] = w_pre[:, :, eq_take].contiguous()
            noise_ir["pre"] = w_pre[:, :, [vas_col]].contiguous()
            del w_pre

        # ---- rate products: per-path exposures and cashflows under the Black-Scholes numeraire --------------------------
        n_expo = len(c.exposure_timeline)
        n_sets = len(c.netting_sets)
        rate_of_set = [[p for p in ns.products if p in self.rate_products] for ns in c.netting_sets]
        extra_expo, extra_pv = [None] * n_sets, [None] * n_sets
        t_pre = 0.0
        if self.rate_products:
            rows = [si for si in range(n_sets) if rate_of_set[si]]
            ir_sets = [NettingSet(name=f"rates_{si}", products=rate_of_set[si]) for si in rows]
            ir_metrics = ([PFEMetric(0.5)] if need_expo else []) + ([PVMetric()] if need_pv or not need_expo else [])
            rm = RiskMetrics(ir_metrics, exposure_timeline=c.exposure_timeline.tolist() if need_expo else None)
            sub = self._sub(models[self.vas_idx], ir_sets, rm)
            sub.injected_normals = noise_ir
            ib = IrcBackend(sub)
            num_model = models[c.model.id_to_model["numeraire"]]
            ib.hybrid = {"ext_rate": float(num_model.param_values()[2]), "chunk": chunk, "captured": []}
            _, t_ir = ib.run()
            t_pre += t_ir.get("preprocessing", 0.0)
            for cap in ib.hybrid["captured"]:
                for r, k in enumerate(cap["idxs"]):
                    si = rows[k]
                    if cap["spill"] is not None:
                        extra_expo[si] = cap["spill"][r]
                    if cap["pv_spill"] is not None:
                        extra_pv[si] = cap["pv_spill"][r]

        # ---- equity products + netting-set terms + metrics on the combined accumulators ------------------------------------
        eq_models = [models[i] for i in self.bs_idx] + ([models[self.cir_idx]] if with_credit else [])
        eq_model = eq_models[0] if len(eq_models) == 1 else ModelConfig(models=eq_models)
        eq_sets = []
        for ns in c.netting_sets:
            view = copy.copy(ns)
            view.products = [p for p in ns.products if p in self.equity_products]
            eq_sets.append(view)
        sub = self._sub(eq_model, eq_sets)
        sub.injected_normals = noise_eq
        eb = EquityBackend(sub)
        t1 = time.perf_counter()
        from mcre.equity import is_equity_exercise
        eb.presim_exercise_all([p for p in sub.products if is_equity_exercise(p)], dev)
        if need_expo:
            reg = [p for p in sub.products if not sub._can_use_analytic_exposure_for_product(p) and not is_equity_exercise(p)]
            if reg:
                eb.presim_regression(reg, dev)
        torch.cuda.synchronize(dev)
        t_pre += time.perf_counter() - t1
        n_params = len(sub.model.model_params)
        results = []
        for si in range(n_sets):
            res = eb._run_split_book(si, dev, n_main, n_params, chunk=chunk, extra_expo=extra_expo[si], extra_pv=extra_pv[si])
            res["param_used"] = lambda kind: [False] * len(c.model.model_params)
            results.append(res)
        if self.differentiate:
            grads = self._tangent_pass(noise_eq, noise_ir, results, chunk, dev)
            # parameters of the credit model: its deterministic mode reads the market hazard curve only, the reference's
            # autograd graph does not reach them (None); everything else is connected
            offs = c.model.param_offsets()
            used = [True] * len(c.model.model_params)
            if self.cir_idx is not None and models[self.cir_idx].deterministic:
                # (a stochastic intensity rides in the joint state tensor: 0.0 where it moves nothing, numbers for the CVA)
                for k in range(len(models[self.cir_idx].model_params)):
                    used[offs[self.cir_idx] + k] = False
            attach_gradients(results, grads, used)
        torch.cuda.synchronize(dev)
        total = time.perf_counter() - t_start
        return results, {"preprocessing": t_pre, "path_generation": total - t_pre, "request_resolution": 0.0}
