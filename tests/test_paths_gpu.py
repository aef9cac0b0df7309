"""Model-level parity: the CUDA path generator (csrc/paths.cu, shared step functions)
against the oracle's restatement of each reference model, with the reference's own
torch.randn stream injected and with native Philox."""
import numpy as np
import pytest
import torch

import cases
from cases import HAZARDS

pytestmark = pytest.mark.gpu


def _models(ns):
    S = ns.SimulationScheme
    bsm = ns.BlackScholesMulti(0.0, 0.03, ["a", "b", "c"], [100.0, 90.0, 110.0], [0.2, 0.3, 0.25],
                               np.array([[1.0, 0.5, 0.2], [0.5, 1.0, 0.3], [0.2, 0.3, 1.0]]))
    vas = lambda: ns.VasicekModel(0.0, 0.03, 0.05, 0.5, 0.02, asset_id="r")
    cir = lambda det=False: ns.CIRPPModel(0.0, "cp", HAZARDS, 0.1, 0.01, 0.02, 0.0001, deterministic=det)
    hyb = ns.ModelConfig([vas(), cir()], inter_asset_correlation_matrix=np.array([0.4]))
    hyb_det = ns.ModelConfig([vas(), cir(True)], inter_asset_correlation_matrix=np.array([0.0]))
    bs4 = ns.ModelConfig([ns.BlackScholesModel(0.0, 100.0 + i, 0.01, 0.3 + 0.02 * i, asset_id=f"e{i}") for i in range(4)],
                         inter_asset_correlation_matrix=np.array([[0.5] for _ in range(6)]))
    mixed = ns.ModelConfig([ns.BlackScholesModel(0.0, 100.0, 0.02, 0.3, asset_id="eq"), vas(), cir(True)],
                           numeraire_model_idx=1, inter_asset_correlation_matrix=np.array([0.2, 0.0, 0.0]))
    sch = ns.SchwartzTwoFactorModel(0.0, [0.0, 0.5, 1.0, 2.0], [50.0, 52.0, 51.0, 55.0], 0.03, 1.2, 0.4, 0.02, 0.15, 0.3)
    hes = ns.HestonModel(0.0, 100.0, 0.03, 0.4, -0.7, 2.0, 0.04, 0.04)
    return [
        ("bs_analytical", ns.BlackScholesModel(0.0, 100.0, 0.05, 0.2), S.ANALYTICAL),
        ("bs_euler", ns.BlackScholesModel(0.0, 100.0, 0.05, 0.2), S.EULER),
        ("bsm_analytical", bsm, S.ANALYTICAL), ("bsm_euler", bsm, S.EULER),
        ("vasicek_analytical", vas(), S.ANALYTICAL), ("vasicek_euler", vas(), S.EULER),
        ("cirpp_euler", cir(), S.EULER),
        ("hybrid_euler", hyb, S.EULER), ("hybrid_det_euler", hyb_det, S.EULER),
        ("bs4_analytical", bs4, S.ANALYTICAL), ("bs4_euler", bs4, S.EULER), ("mixed_euler", mixed, S.EULER),
        ("schwartz_analytical", sch, S.ANALYTICAL), ("schwartz_euler", sch, S.EULER),
        ("heston_euler", hes, S.EULER), ("heston_qe", hes, S.QE),
    ]


TIMELINE = [0.0, 0.1, 0.25, 0.5, 0.75, 1.0, 1.7]


@pytest.mark.parametrize("idx", range(16))
@pytest.mark.parametrize("draws", ["torch", "philox"])
def test_paths_match_oracle(idx, draws):
    from oracle import engine, models
    from engine.engine import MonteCarloEngine
    ns = cases.Namespace()
    name, model, scheme = _models(ns)[idx]
    n, steps = 512, 3
    n_sub = engine.count_substeps(0.0, TIMELINE, steps)
    dim = models.noise_dim(model)
    qe = scheme.name == "QE"
    eng = MonteCarloEngine(torch.tensor(TIMELINE, dtype=torch.float64), scheme, model, n, steps)
    if draws == "torch":
        d = engine.torch_reference_draws(43, n, n_sub, dim, qe)
        eng.injected_normals, eng.injected_uniforms = d.z, d.u
        tol = 1e-12
    else:
        d = engine.PhiloxDraws(43, n, n_sub, dim, with_uniforms=qe)
        tol = 1e-9   # device vs host libm in Box-Muller
    got = eng.generate_paths().cpu().numpy()
    p = [np.float64(v) for v in models.param_values(model)]
    want = engine.generate_paths(model, p, TIMELINE, n, steps, scheme, d, smoothing=False)
    want = np.stack([np.stack([np.broadcast_to(c, (n,)) for c in st], axis=1) for st in want], axis=1)
    assert got.shape == want.shape, name
    err = np.max(np.abs(got - want) / np.maximum(1.0, np.abs(want)))
    assert err < tol, f"{name} [{draws}]: max rel err {err:.3e}"


def test_philox_uniform_bits_match_oracle():
    """Same Philox4x32-10 stream on both sides: the first normal of each path agrees to libm accuracy."""
    from oracle import philox
    from engine.engine import MonteCarloEngine
    ns = cases.Namespace()
    model = ns.BlackScholesModel(0.0, 1.0, 0.0, 1.0)
    n = 4096
    eng = MonteCarloEngine(torch.tensor([1.0], dtype=torch.float64), ns.SimulationScheme.ANALYTICAL, model, n, 1)
    s = eng.generate_paths().cpu().numpy()[:, 0, 0]
    z = np.log(s) + 0.5          # S = exp(z - 1/2)
    want = philox.normals(np.arange(n, dtype=np.uint64), 1, 1, 43)[0, :, 0]
    assert np.max(np.abs(z - want)) < 1e-12
    assert abs(z.mean()) < 5 / np.sqrt(n) and abs(z.std() - 1) < 0.05


def test_deterministic_cirpp_step_tracks_hazard_curve():
    """Reference test tests/pytests/test_cirpp.py:8-45 through the CUDA single-step seam."""
    ns = cases.Namespace()
    model = ns.CIRPPModel(0.0, "cp", {1.0: 0.02, 2.0: 0.03, 5.0: 0.04}, 0.2, 0.03, 0.01, 0.02, deterministic=True)
    state = model.get_state(num_paths=4)
    assert torch.allclose(state[:, 0], torch.full((4,), 0.02, dtype=state.dtype))
    nxt = model.simulate_time_step_euler(torch.tensor(0.0), torch.tensor(1.5), state, torch.zeros((4, 1), dtype=state.dtype))
    assert torch.allclose(nxt[:, 0], torch.full((4,), 0.03, dtype=state.dtype))
    assert torch.allclose(nxt[:, 1], torch.full((4,), 0.03, dtype=state.dtype))
    cs = model.survival_probability(torch.tensor(1.0), torch.tensor(2.0), nxt[:, 0])
    assert torch.allclose(cs, torch.full((4,), float(np.exp(-0.03)), dtype=state.dtype))
