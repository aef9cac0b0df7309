"""The plan cache of the interest-rate / credit plan compiler (mcre/irc.py): a sweep over the inter-model correlation
reuses every lowered table except the Cholesky factor.  A cached plan must be indistinguishable from a fresh lowering
of the same run - checked table by table, on the host (no device needed), for the headline config and a
two-netting-set collateralised book."""
import ctypes as C

import numpy as np
import pytest

import cases


def _tables(desc, keep):
    out = {k: np.array(v, copy=True) for k, v in keep.items()}
    scal = {}
    for name, ctype in desc._fields_:
        v = getattr(desc, name)
        if isinstance(v, (int, float)):
            scal[name] = v
    return out, scal


def _lower_all(ns, builder, rho, **kw):
    from mcre.irc import IrcBackend
    model, sets, metrics, tl = builder(ns, rho=rho, **kw)
    rm = ns.RiskMetrics(metrics, exposure_timeline=tl)
    sc = ns.SimulationController(sets, model, rm, 4096, 4096, 1, ns.SimulationScheme.EULER)
    be = IrcBackend(sc)
    pre = be.lower([], sc.products)
    main = be.lower(list(range(len(sets))), [])
    return pre, main


@pytest.mark.parametrize("builder,kw", [(cases.cfg3_wwr, {}), (cases.wwr_cva, dict(n_expo=21, maturity=5.0))])
def test_cached_plan_equals_fresh_lowering(builder, kw, monkeypatch):
    from mcre import irc
    ns = cases.Namespace()
    irc._PLAN_CACHE.clear()
    monkeypatch.setenv("MCRE_PLAN_CACHE", "1")
    _lower_all(ns, builder, 0.3, **kw)                      # fills the cache
    n_entries = len(irc._PLAN_CACHE)
    assert n_entries == 2
    cached = _lower_all(ns, builder, -0.7, **kw)            # served from the cache, new Cholesky factor
    assert len(irc._PLAN_CACHE) == n_entries
    monkeypatch.setenv("MCRE_PLAN_CACHE", "0")
    fresh = _lower_all(ns, builder, -0.7, **kw)
    for (dc, kc, ic), (df, kf, if_) in zip(cached, fresh):
        tc, sc_ = _tables(dc, kc)
        tf, sf = _tables(df, kf)
        assert sc_ == sf
        assert tc.keys() == tf.keys()
        for k in tc:
            assert np.array_equal(tc[k], tf[k]), k
        assert ic["n_expo"] == if_["n_expo"] and ic["acc"] == if_["acc"]
    # the Cholesky factor really is this run's
    chol = cached[1][1]["chol"]
    assert abs(chol[2] - (-0.7)) < 1e-15 and abs(chol[3] - np.sqrt(1 - 0.49)) < 1e-15


def test_cache_key_separates_different_runs(monkeypatch):
    """Anything lower() reads other than the correlation changes the key: parameters, schedules, grids, metrics."""
    from mcre import irc
    ns = cases.Namespace()
    monkeypatch.setenv("MCRE_PLAN_CACHE", "1")
    irc._PLAN_CACHE.clear()
    _lower_all(ns, cases.wwr_cva, 0.3, n_expo=21, maturity=5.0)
    _lower_all(ns, cases.wwr_cva, 0.3, n_expo=21, maturity=5.0, vol=0.1)        # model parameter
    _lower_all(ns, cases.wwr_cva, 0.3, n_expo=11, maturity=5.0)                 # exposure grid
    _lower_all(ns, cases.wwr_cva, 0.3, n_expo=21, maturity=4.0)                 # product schedule
    _lower_all(ns, cases.wwr_cva, 0.3, n_expo=21, maturity=5.0, extra_metrics=False)   # metric list
    assert len(irc._PLAN_CACHE) == 10
    # tangent plans and exercise products are never cached
    model, sets, metrics, tl = cases.wwr_cva(ns, n_expo=11, maturity=2.5)
    sc = ns.SimulationController(sets, model, ns.RiskMetrics(metrics, exposure_timeline=tl), 1024, 1024, 1,
                                 ns.SimulationScheme.EULER, True)
    be = irc.IrcBackend(sc)
    assert be._lower_key([0], [], None, None) is None
