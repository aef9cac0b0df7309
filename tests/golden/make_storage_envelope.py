"""Generates tests/golden/storage_envelope.json with the UNMODIFIED reference (/root/reference/src): the
reachable-inventory envelope, action dates, rate-curve samples and cost look-ups of the two storage contracts of
tests/cases.py:storage_s2f (storage_helpers.py:56-437, storage.py:47-66).  Build container only:
    python tests/golden/make_storage_envelope.py"""
import json
import os
import sys
import types

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
sys.path.insert(0, "/root/reference/src")
_helpers = types.ModuleType("helpers")
_helpers.__path__ = ["/root/reference/src/helpers"]
sys.modules["helpers"] = _helpers
mpl = types.ModuleType("matplotlib")
mpl.pyplot = types.ModuleType("matplotlib.pyplot")
sys.modules.setdefault("matplotlib", mpl)
sys.modules.setdefault("matplotlib.pyplot", mpl.pyplot)

import numpy as np  # noqa: E402
import torch  # noqa: E402

import cases  # noqa: E402


def describe(ns, which):
    _, sets, _, _ = cases.storage_s2f(ns, which=which)
    st = sets[0].products[0]
    cfg = st.storage_config
    levels = np.linspace(-10.0, 300000.0 if which == "storage2" else 100.0, 37)
    dates = [0.0, 10.5, 200.0, 272.9999, 273.0, 400.0, 1000.0]
    return dict(
        envelope=[[w.start_date, w.end_date, w.vmin, w.vmax] for w in cfg.volume_constraints],
        action_dates=st.product_timeline.tolist(), next_dates=st.next_action_dates.tolist(),
        levels=levels.tolist(), dates=dates,
        injection=[cfg.interpolate_rate_tensor(torch.tensor(levels), cfg.get_injection_flexibility_slice(t)).tolist() for t in dates],
        withdrawal=[[cfg.get_withdrawal_flexibility_rate(t, float(v)) for v in levels] for t in dates],
        costs=[[cfg.get_variable_injection_cost(t), cfg.get_variable_withdrawal_cost(t)] for t in dates],
        grid=[[cfg.grid_step(0.0, 90.0, 10), cfg.state_scale(0.0, 90.0, 10)], [cfg.grid_step(5.0, 5.0, 10), cfg.state_scale(5.0, 5.0, 10)]],
        volume_of_state=st.state_to_volume(200.0, torch.tensor([0.0, 2.5, 9.0])).tolist())


if __name__ == "__main__":
    ns = cases.Namespace()
    out = {w: describe(ns, w) for w in ("storage1", "storage2")}
    with open(os.path.join(HERE, "storage_envelope.json"), "w") as f:
        json.dump(out, f)
    print("wrote storage_envelope.json")
