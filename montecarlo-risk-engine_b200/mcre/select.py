"""PFE order statistics via the radix-select kernels (csrc/select.cu)."""
from __future__ import annotations

import ctypes as C
import math

import numpy as np
import torch

from mcre import binding as B
from mcre import runtime as RT
from metrics.metric import MetricType


def select_rows(values, n_local, ranks):
    """values: device tensor [rows, stride]; ranks: int64 [rows, R] global 0-based ranks.
    Returns numpy [rows, R] of exact order statistics over all ranks' paths."""
    L = B.lib()
    L.mcre_select_create.argtypes = [C.c_int32, C.c_int32, C.POINTER(C.c_void_p)]
    L.mcre_select_destroy.argtypes = [C.c_void_p]
    L.mcre_select_destroy.restype = None
    L.mcre_select_begin.argtypes = [C.c_void_p, C.POINTER(C.c_int64), C.c_void_p]
    L.mcre_select_count.argtypes = [C.c_void_p, C.c_void_p, C.c_int64, C.c_int64, C.c_int32, C.c_void_p, C.c_void_p]
    L.mcre_select_scan.argtypes = [C.c_void_p, C.c_int32, C.c_void_p, C.c_void_p]
    L.mcre_select_finish.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
    L.mcre_select_compact.argtypes = [C.c_void_p, C.c_void_p, C.c_int64, C.c_int64, C.c_int32, C.c_void_p]
    rows, R = ranks.shape
    dev = values.device
    plan = C.c_void_p()
    B.check(L.mcre_select_create(rows, R, C.byref(plan)))
    try:
        rk = np.ascontiguousarray(ranks, dtype=np.int64)
        st = RT.stream_ptr()
        B.check(L.mcre_select_begin(plan, rk.ctypes.data_as(C.POINTER(C.c_int64)), st))
        hist = torch.zeros(rows * R * 256, dtype=torch.int64, device=dev)
        stride = values.stride(0) if values.dim() == 2 else n_local
        for p in range(8):
            B.check(L.mcre_select_count(plan, values.data_ptr(), stride, n_local, p, hist.data_ptr(), st))
            _, world = RT.dist_info()
            if world > 1:
                import torch.distributed as dist
                dist.all_reduce(hist)  # integer sum: exact and order independent
            B.check(L.mcre_select_scan(plan, p, hist.data_ptr(), st))
            if p == 1:
                # 16 bits fixed: keep only the elements that can still be a wanted order statistic
                B.check(L.mcre_select_compact(plan, values.data_ptr(), stride, n_local, 2, st))
        out = torch.empty(rows * R, dtype=torch.float64, device=dev)
        B.check(L.mcre_select_finish(plan, out.data_ptr(), st))
        return out.cpu().numpy().reshape(rows, R)
    finally:
        L.mcre_select_destroy(plan)


def pfe_from_order_statistics(q, n, lo, mid, hi, q_index):
    """(value, standard error) of the quantile estimate (reference: pfe_metric.py:13-44)."""
    if q_index == 0 or q_index == n - 1:
        return mid, 0.0
    if lo == mid and hi == mid:
        return mid, 0.0
    f = max((hi - lo) / 2.0, 1e-6)
    return mid, math.sqrt(q * (1.0 - q) / (n * f * f))


def order_statistics(ctrl, spill, n_local, n_total):
    """spill: [n_sets_in_group, n_metric, n_local] unsecured exposures.
    -> per set: {quantile: ([(pfe, se)] per metric date, [None...])}"""
    n_sets, n_metric = spill.shape[0], spill.shape[1]
    pfes = [m for m in ctrl.risk_metrics.metrics if m.metric_type == MetricType.PFE]
    out = [dict() for _ in range(n_sets)]
    flat = spill.reshape(n_sets * n_metric, spill.shape[2])
    for metric in pfes:
        qi = metric.quantile_index(n_total)
        ranks = np.array([max(qi - 1, 0), min(max(qi, 0), n_total - 1), min(qi + 1, n_total - 1)], dtype=np.int64)
        sel = select_rows(flat, n_local, np.tile(ranks, (n_sets * n_metric, 1)))
        sel = sel.reshape(n_sets, n_metric, 3)
        for s in range(n_sets):
            vals = [pfe_from_order_statistics(metric.quantile, n_total, *sel[s, m], q_index=qi)
                    for m in range(n_metric)]
            out[s][metric.quantile] = (vals, [None] * n_metric)
    return out
