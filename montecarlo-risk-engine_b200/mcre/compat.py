"""Debugging / compatibility helpers behind Model.get_state and
Model.simulate_time_step_* (reference: src/models/*.py).  They run one sub-step of the
CUDA path generator with the caller's noise, so even this seam has no CPU fallback."""
from __future__ import annotations

import torch

from common.packages import FLOAT
from mcre.timegrid import TimeGrid


def initial_state(model, num_paths):
    from mcre.paths import initial_state_values
    row = torch.tensor(initial_state_values(model), dtype=FLOAT)
    return row.unsqueeze(0).expand(num_paths, -1).clone()


def single_step(model, scheme, time1, time2, state, corr_randn):
    """state [N, D], corr_randn [N, d] already-correlated noise -> next state [N, D]."""
    from mcre.paths import generate
    t1 = float(torch.as_tensor(time1).reshape(-1)[0])
    t2 = float(torch.as_tensor(time2).reshape(-1)[0])
    state = torch.as_tensor(state, dtype=FLOAT)
    n = state.shape[0]
    grid = TimeGrid(dates=[t2], n_pre_dates=0, t1=[t1], t2=[t2], dt_nominal=[t2 - t1], dt=[t2 - t1], date_after=[0])
    out = torch.empty_like(state)
    # the generator takes one initial state per launch: group identical rows
    uniq, inv = torch.unique(state, dim=0, return_inverse=True)
    z = torch.as_tensor(corr_randn, dtype=FLOAT).reshape(n, -1)
    for k in range(uniq.shape[0]):
        sel = (inv == k).nonzero().reshape(-1)
        paths = generate(model, [t2], len(sel), 1, scheme, seed=0, inject_z=z[sel].unsqueeze(0), grid=grid,
                         init_state=uniq[k].tolist(), identity_chol=True)
        out[sel] = paths[:, 0, :].cpu()
    return out


def torch_reference_draws(seed, n_paths, n_sub, dim, qe=False):
    """The reference's own random stream for one engine run, regenerated on the host: it seeds the
    (global) torch generator in MonteCarloEngine.__init__ (engine.py:25: 42 for the pre-simulation, 43 for
    the main simulation), then draws per sub-step one torch.randn(N, d) (model.py:47) and, under the Heston QE
    scheme, one torch.rand_like(m) after it (heston.py:191-192).  A private generator with the same seed
    yields the same numbers without touching the caller's global RNG state.
    -> (z [n_sub, N, d], u [n_sub, N] or None), float64 host tensors."""
    g = torch.Generator(device="cpu")
    g.manual_seed(int(seed))
    z = torch.empty((n_sub, n_paths, dim), dtype=torch.float64)
    u = torch.empty((n_sub, n_paths), dtype=torch.float64) if qe else None
    for s in range(n_sub):
        torch.randn(n_paths, dim, generator=g, dtype=torch.float64, out=z[s])
        if qe:
            u[s] = torch.rand(n_paths, 1, generator=g, dtype=torch.float64)[:, 0]
    return z, u


def inject_reference_stream(ctrl):
    """RNG compatibility mode (`SimulationController.rng_compat = "torch"` or MCRE_RNG=torch): feed the
    kernels the reference's torch.randn stream instead of Philox, so seeded known answers of the reference
    (tests/pytests/test_american_option.py:61, test_cva.py:188-189) are reproduced digit for digit.  Only
    the normals are made on the host; stepping, payoffs, regressions and metrics stay on the GPU."""
    from common.enums import SimulationScheme
    from mcre.timegrid import build_time_grid
    model = ctrl.model
    t0 = float(model.calibration_date[0]) if not hasattr(model, "models") else float(model.models[0].calibration_date[0])
    n_sub = build_time_grid(t0, ctrl.simulation_timeline.tolist(), ctrl.num_steps).n_sub
    dim = int(model.simulation_dim)
    qe = ctrl.simulation_scheme == SimulationScheme.QE
    pre = None
    if ctrl.requires_regression and ctrl.num_paths_presim > 0:
        pre = torch_reference_draws(42, ctrl.num_paths_presim, n_sub, dim, qe)
    main = torch_reference_draws(43, ctrl.num_paths_mainsim, n_sub, dim, qe)
    ctrl.inject_normals(pre=None if pre is None else pre[0], main=main[0])
    if qe:
        ctrl.injected_uniforms = {"pre": None if pre is None else pre[1], "main": main[1]}
    # Brownian-bridge barrier options draw from their own numpy default_rng(12345) at every payoff call
    # (barrier_option.py:49-50, 174, 200): once in the pre-simulation if the product is regressed, then in the
    # main simulation; first barrier before second
    import numpy as np
    bridge = {"pre": {}, "main": {}}
    for prod in ctrl.products:
        if not getattr(prod, "use_brownian_bridge", False):
            continue
        gen = np.random.default_rng(12345)
        n_int = len(prod.modeling_timeline) - 1
        double = prod.barrier2 is not None and prod.barrier_option_type2 is not None
        passes = []
        if pre is not None and ctrl._product_requires_regression(prod):
            passes.append(("pre", ctrl.num_paths_presim))
        passes.append(("main", ctrl.num_paths_mainsim))
        for which, n in passes:
            u1 = gen.uniform(0, 1, size=(n, n_int))
            u2 = gen.uniform(0, 1, size=(n, n_int)) if double else None
            bridge[which][prod.product_id] = (u1, u2)
    if bridge["main"]:
        ctrl.injected_bridge = bridge
