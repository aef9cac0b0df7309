"""Host-side finishing of metrics from the O(T) accumulators the kernels return.

Restates Metric._compute_mc_mean_and_error (reference: src/metrics/metric.py:26-35) and
the per-metric result shapes of src/metrics/*.py on sums instead of per-path vectors:
  value = c + S1/N,   var = (S2 - S1^2/N)/(N-1),   mc_error = sqrt(var/N)
with S1 = sum(x - c), S2 = sum((x - c)^2) and c the pilot-path shift.
"""
from __future__ import annotations

import math

import numpy as np

from metrics.metric import Metric, MetricType


def mean_and_error(s1, s2, shift, n):
    mean = shift + s1 / n
    if n < 2:
        return mean, float("nan")
    var = (s2 - s1 * s1 / n) / (n - 1)
    return mean, math.sqrt(max(var, 0.0)) / math.sqrt(n)


def _zero_result(metric, n_metric):
    n_eval = 1 if metric.metric_type in {MetricType.PV, MetricType.CVA, MetricType.EEPE} else n_metric
    return [(0.0, 0.0) for _ in range(n_eval)]


def finish_results(ctrl, raw, analytic, has_pathwise):
    """-> (results[set][metric] = [(value, err), ...], grads[set][metric][eval] = tuple)"""
    metrics = ctrl.risk_metrics.metrics
    n_metric = len(ctrl.metric_exposure_timeline)
    n_params = len(ctrl.model.model_params)
    results, grads = [], []
    for si, ns in enumerate(ctrl.netting_sets):
        r = raw[si] if raw is not None else None
        set_results, set_grads = [], []
        for mi, metric in enumerate(metrics):
            kind = metric.metric_type
            vals, tans = [], []
            if (kind == MetricType.CVA and ns.counterparty_id is not None
                    and getattr(metric, "counterparty_id", None) != ns.counterparty_id):
                vals = _zero_result(metric, n_metric)
                tans = [None] * len(vals)
            elif kind == MetricType.PV:
                if r is not None and has_pathwise[si]:
                    v, e = r["pv"][0]
                    t = r["pv"][1]
                else:
                    v, e, t = 0.0, 0.0, None
                if metric.evaluation_type == Metric.EvaluationType.ANALYTICAL:
                    v += analytic[si][mi]
                vals, tans = [(v, e)], [t]
            elif kind == MetricType.CE:
                vals, tans = [r["pos"][0][0]], [r["pos"][1][0]]
            elif kind == MetricType.EPE:
                vals, tans = list(r["pos"][0]), list(r["pos"][1])
            elif kind == MetricType.ENE:
                vals, tans = list(r["neg"][0]), list(r["neg"][1])
            elif kind == MetricType.EEPE:
                epe = np.array([v for v, _ in r["pos"][0]])
                # time average of EPE; "error" = unbiased std over time / sqrt(T) (eepe_metric.py:11-15)
                with np.errstate(invalid="ignore", divide="ignore"):
                    err = float(np.std(epe, ddof=1) / math.sqrt(len(epe))) if len(epe) > 1 else float("nan")
                t = None
                if r["pos"][1][0] is not None:
                    t = np.mean(np.stack(r["pos"][1]), axis=0)
                vals, tans = [(float(epe.mean()), err)], [t]
            elif kind == MetricType.PFE:
                q = r["pfe"][metric.quantile]
                vals, tans = list(q[0]), list(q[1])
            elif kind == MetricType.CVA:
                vals, tans = [r["cva"][0]], [r["cva"][1]]
            else:
                raise NotImplementedError(kind)
            set_results.append(vals)
            if ctrl.differentiate:
                used = r["param_used"](kind) if r is not None else [False] * n_params
                # closed-form part of an analytic PV (controller._analytic_pv): host autograd gradients
                ag = None
                if kind == MetricType.PV and metric.evaluation_type == Metric.EvaluationType.ANALYTICAL:
                    ag = getattr(ctrl, "_analytic_grads", None)
                    ag = ag[si][mi] if ag is not None else None
                per_eval = []
                for t in tans:
                    row = [None] * n_params
                    if t is not None:
                        row = [np.asarray(t[i]) if used[i] else None for i in range(n_params)]
                    if ag is not None:
                        row = [a if b is None else (b if a is None else np.asarray(a + b)) for a, b in zip(ag, row)]
                    per_eval.append(tuple(row))
                set_grads.append(per_eval)
        results.append(set_results)
        if ctrl.differentiate:
            grads.append(set_grads)
    return results, grads


def irc_raw_to_neutral(acc, shift, n, nt, n_metric, flags):
    """Slot block [n_metric+1][4+2nt] of one netting set -> neutral per-metric dict pieces."""
    from mcre import binding as B
    out = {}

    def series(col, tcol):
        vals, tans = [], []
        for m in range(n_metric):
            vals.append(mean_and_error(acc[m, col], acc[m, col + 1], shift[m, col], n))
            tans.append(acc[m, tcol:tcol + nt] / n if nt else None)
        return vals, tans

    if flags & B.ACC_POS:
        out["pos"] = series(0, 4)
    if flags & B.ACC_NEG:
        out["neg"] = series(2, 4 + nt)
    tail = n_metric
    if flags & B.ACC_PV:
        out["pv"] = (mean_and_error(acc[tail, 0], acc[tail, 1], shift[tail, 0], n),
                     acc[tail, 4:4 + nt] / n if nt else None)
    if flags & B.ACC_CVA:
        out["cva"] = (mean_and_error(acc[tail, 2], acc[tail, 3], shift[tail, 2], n),
                      acc[tail, 4 + nt:4 + 2 * nt] / n if nt else None)
    return out
