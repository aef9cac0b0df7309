"""BASELINE.json's configurations at FULL size on the GPU, checked through size-independent
properties (the oracle cannot run these sizes): shard-and-combine identity, determinism,
linearity of netted exposures, exact order statistics against a full sort, martingale /
dominance relations, consistency of exposure at t = 0 with PV."""
import ctypes as C

import numpy as np
import pytest
import torch

import cases
import parity_helpers as helpers

pytestmark = pytest.mark.gpu


def _controller(ns, model, sets, metrics, tl, n_main, n_pre, steps, scheme, diff=False):
    rm = ns.RiskMetrics(metrics, exposure_timeline=tl) if tl is not None else ns.RiskMetrics(metrics)
    return ns.SimulationController(sets, model, rm, n_main, n_pre, steps, scheme, diff)


def test_config3_wwr_cva_full_size_shards_combine_bit_exactly():
    """2^24 paths x 240 steps: the accumulators of two half-range shards, tree-combined like the
    multi-GPU all-reduce does, equal the single full-range launch bit for bit; the CVA is positive,
    reproducible, and agrees with a 2^22-path run within 4 combined standard errors."""
    from mcre import binding as B
    from mcre import runtime as RT
    from mcre.irc import CHUNK_PATHS, IrcBackend
    ns = cases.Namespace()
    n, n_pre = 1 << 24, 1 << 20
    tl = np.arange(241) / 24.0
    model, sets, metrics, _ = cases.wwr_cva(ns, rho=0.5, maturity=10.0, extra_metrics=False)
    sc = _controller(ns, model, sets, metrics, tl, n, n_pre, 1, ns.SimulationScheme.EULER)
    res = sc.run_simulation()
    cva, err = float(res.get_results("irs", "cva[GM]")[0]), float(res.get_mc_error("irs", "cva[GM]")[0])
    assert cva > 0.0 and 0.0 < err < 0.01 * cva
    res2 = _controller(ns, *cases.wwr_cva(ns, rho=0.5, maturity=10.0, extra_metrics=False)[:3], tl, n, n_pre, 1,
                       ns.SimulationScheme.EULER).run_simulation()
    assert float(res2.get_results("irs", "cva[GM]")[0]) == cva            # deterministic
    small = _controller(ns, *cases.wwr_cva(ns, rho=0.5, maturity=10.0, extra_metrics=False)[:3], tl, 1 << 22, n_pre, 1,
                        ns.SimulationScheme.EULER).run_simulation()
    c4, e4 = float(small.get_results("irs", "cva[GM]")[0]), float(small.get_mc_error("irs", "cva[GM]")[0])
    assert abs(c4 - cva) <= 4.0 * np.hypot(err, e4)
    assert 1.6 < e4 / err < 2.4                                               # error ~ 1 / sqrt(N)

    # shard-and-combine through the C ABI (same plan, same coefficients)
    L, dev = B.lib(), RT.compute_device()
    be = IrcBackend(sc)
    coefs = be.presim_coefficients(sc.products, dev)
    desc, keep, info = be.lower([0], [])
    plan = C.c_void_p()
    B.check(L.mcre_irc_create(C.byref(desc), C.byref(plan)))
    try:
        coef = np.zeros((info["n_expo"], 1, 3, 1))
        coef[:, 0, :, 0] = coefs[id(sc.products[0])][0]
        arr, ptr = B.as_dp(coef)
        B.check(L.mcre_irc_set_coefficients(plan, ptr, RT.stream_ptr()))
        slots = L.mcre_irc_main_slots(plan)

        def run(begin, count):
            acc = torch.zeros(slots, dtype=torch.float64, device=dev)
            shift = torch.zeros(slots, dtype=torch.float64, device=dev)
            partial = torch.empty(L.mcre_irc_partial_bytes(plan, count, CHUNK_PATHS, 0) // 8 + 1, dtype=torch.float64, device=dev)
            rng = B.Rng()
            rng.mode, rng.seed, rng.stream, rng.n_paths_total = B.RNG_PHILOX, 43, 0, n
            sh = B.Shard(begin, count, CHUNK_PATHS)
            B.check(L.mcre_irc_mainsim(plan, C.byref(rng), C.byref(sh), partial.data_ptr(), acc.data_ptr(), shift.data_ptr(),
                                       None, RT.stream_ptr()))
            return acc.cpu(), shift.cpu()

        full, shift_full = run(0, n)
        lo, shift_lo = run(0, n // 2)
        hi, shift_hi = run(n // 2, n // 2)
        assert torch.equal(shift_lo, shift_full) and torch.equal(shift_hi, shift_full)   # pilot path is global path 0
        assert torch.equal(RT.tree_sum([lo, hi]), full)
    finally:
        L.mcre_irc_destroy(plan)


def test_config2_irs_profiles_full_size_linearity_and_signs():
    """2^22 paths x 120 quarterly steps, EE / EPE / ENE / EEPE / PFE with an MPoR-collateralised set:
    sign constraints, PFE >= EPE-consistent ordering, and linearity - a netting set holding the same
    swap twice has exactly twice the exposure profile (regression proxies are linear in the cashflows)."""
    ns = cases.Namespace()
    n = 1 << 22
    tl = np.arange(121) * 0.25
    model = ns.VasicekModel(0., 0.03, 0.05, 0.02, 0.02)
    swap = lambda: ns.InterestRateSwap(0.0, 30.0, 1.0, 0.03, 0.25, 0.25, ns.IRSType.PAYER)
    sets = [ns.NettingSet(name="one", products=[swap()]),
            ns.NettingSet(name="two", products=[swap(), swap()]),
            ns.NettingSet(name="collateralised", products=[swap()], margin_period_of_risk=0.25)]
    metrics = [ns.PVMetric(), ns.EPEMetric(), ns.ENEMetric(), ns.EEPEMetric(), ns.PFEMetric(0.95)]
    res = _controller(ns, model, sets, metrics, tl, n, n, 1, ns.SimulationScheme.EULER).run_simulation()
    epe1, epe2 = np.array(res.get_results("one", "epe")), np.array(res.get_results("two", "epe"))
    ene1, pfe1, pfe2 = np.array(res.get_results("one", "ene")), np.array(res.get_results("one", "pfe[0.95]")), np.array(res.get_results("two", "pfe[0.95]"))
    assert np.all(epe1 >= 0.0) and np.all(ene1 <= 0.0)
    assert np.all(pfe1[1:-1] >= epe1[1:-1] + ene1[1:-1])          # 95 % quantile above the mean exposure
    helpers.assert_close(epe2, 2.0 * epe1, 1e-10, 1e-13, "EPE linearity")
    helpers.assert_close(pfe2, 2.0 * pfe1, 1e-10, 1e-13, "PFE linearity")
    helpers.assert_close(res.get_results("two", "pv"), 2.0 * np.array(res.get_results("one", "pv")), 1e-10, 1e-13, "PV linearity")
    helpers.assert_close(res.get_results("one", "eepe"), [epe1.mean()], 1e-12, 0.0, "EEPE = time average of EPE")
    # MPoR = one grid step, no threshold: unsecured exposure = E(t_m) - E(t_m-1) path by path, so its mean
    # telescopes: EPE_c + ENE_c at date m = EE(t_m) - EE(t_m-1) of the uncollateralised set
    epe_c, ene_c = np.array(res.get_results("collateralised", "epe")), np.array(res.get_results("collateralised", "ene"))
    ee = epe1 + ene1
    helpers.assert_close((epe_c + ene_c)[1:], ee[1:] - ee[:-1], 1e-9, 1e-12, "collateral telescoping identity")
    assert epe1[-1] == 0.0 and res.get_mc_error("one", "epe")[0] == 0.0    # matured swap, deterministic t = 0


def test_pfe_order_statistics_equal_full_sort_at_full_size():
    """Radix select (csrc/select.cu) vs torch.sort on 8 rows x 2^22 doubles with ties, negatives, zeros."""
    from mcre.select import select_rows
    g = torch.Generator(device="cuda").manual_seed(5)
    n, rows = 1 << 22, 8
    x = torch.randn(rows, n, dtype=torch.float64, device="cuda", generator=g)
    x[1] = torch.round(x[1] * 4.0) / 4.0                  # heavy ties
    x[2] = torch.clamp(x[2], min=0.0)                      # half the row equal to zero
    x[3] = -torch.abs(x[3]) * 1e-300                       # tiny negatives
    x[4, ::2] = 0.0
    ranks = np.array([[0, 1, 2], [n // 2 - 1, n // 2, n // 2 + 1], [int(0.95 * n) - 1, int(0.95 * n), int(0.95 * n) + 1],
                      [n - 3, n - 2, n - 1]] * 2, dtype=np.int64)
    got = select_rows(x, n, ranks)
    srt = torch.sort(x, dim=1).values.cpu().numpy()
    want = np.stack([srt[r, ranks[r]] for r in range(rows)])
    assert np.array_equal(got, want)


def test_config4_bermudan_swaption_full_size_consistency():
    """2^22 paths x 40 exercise dates: the exposure at t = 0 (regressed continuation value of the
    pre-simulation) agrees with the main simulation's PV within 4 standard errors, the exposure profile
    decays to zero at the last exercise date, and more exercise rights are worth more."""
    ns = cases.Namespace()
    n = 1 << 22
    model, sets, metrics, tl = cases.bermudan_swaption(ns, n_ex=40)
    res = _controller(ns, model, sets, metrics, tl, n, n, 1, ns.SimulationScheme.EULER).run_simulation()
    pv, pv_err = float(res.get_results("bermudan", "pv")[0]), float(res.get_mc_error("bermudan", "pv")[0])
    epe = np.array(res.get_results("bermudan", "epe"))
    assert pv > 0 and abs(epe[0] - pv) <= 4.0 * np.hypot(pv_err, pv_err)   # both are N-path estimates of the same value
    assert epe[-1] == 0.0 and np.all(epe >= 0.0)
    model2, sets2, metrics2, tl2 = cases.bermudan_swaption(ns, n_ex=8)
    res8 = _controller(ns, model2, sets2, [ns.PVMetric()], None, n, n, 1, ns.SimulationScheme.EULER).run_simulation()
    assert float(res8.get_results("bermudan", "pv")[0]) < pv


def test_config5_heston_basket_full_size_martingale_and_dominance():
    """2^24 paths x 252 QE sub-steps x 5 assets: the discounted basket is a martingale (strike-0 call
    = basket spot within 4 standard errors), up-and-out <= vanilla, arithmetic Asian <= vanilla call on
    the same basket (Jensen), Greeks finite and delta-like sensitivities positive."""
    ns = cases.Namespace()
    n = 1 << 24
    model, sets, metrics, _ = cases.heston_basket5(ns)
    ids = [f"h{i}" for i in range(5)]
    w = [0.2] * 5
    sets.append(ns.NettingSet(name="forward", products=[ns.BasketOption(1.0, ids, w, 0.0, ns.OptionType.CALL)]))
    sets.append(ns.NettingSet(name="vanilla", products=[ns.BasketOption(1.0, ids, w, 100.0, ns.OptionType.CALL)]))
    res = _controller(ns, model, sets, metrics, None, n, 0, 21, ns.SimulationScheme.QE).run_simulation()
    fwd, fwd_err = float(res.get_results("forward", "pv")[0]), float(res.get_mc_error("forward", "pv")[0])
    assert abs(fwd - 100.0) <= 4.0 * fwd_err + 2e-2           # allowance for the QE discretisation bias at 252 steps
    vanilla = float(res.get_results("vanilla", "pv")[0])
    # the reference does not discount path-dependent payoffs (numeraire of the FIRST monitoring date,
    # asian_option.py:90, barrier_option.py:312): compare undiscounted-to-undiscounted
    assert 0.0 < float(res.get_results("barrier", "pv")[0]) < vanilla * np.exp(0.03)
    assert 0.0 < float(res.get_results("asian", "pv")[0]) < vanilla * np.exp(0.03)
    # pathwise Greeks at 2^22 paths: all 35 finite, spot sensitivities of the Asian call positive
    model, sets, metrics, _ = cases.heston_basket5(ns)
    resg = _controller(ns, model, sets, metrics, None, 1 << 22, 0, 21, ns.SimulationScheme.QE, True).run_simulation()
    g = np.array([float(x) for x in resg.get_derivatives("asian", "pv")[0]])
    assert g.shape == (35,) and np.all(np.isfinite(g)) and np.all(g[0::7] > 0.0)
