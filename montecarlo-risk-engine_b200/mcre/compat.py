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
