"""Results container with the reference's nested layout and accessors
(reference: src/controller/simulation_results.py:5-338).

Layout (kept identical so downstream scripts keep working):
  results[set][metric]                 -> list of (value, mc_error) per evaluation
  derivatives[set][metric][eval]       -> tuple over model parameters (None = unused)
  second_derivatives[set][metric][eval][param_i] -> tuple over model parameters
All leaves are numpy 0-d arrays (the reference converts tensors with .numpy()).
"""
import numpy as np

_LEGACY = {
    "netting_set": ("prod_idx", "product", "product_idx"),
    "metric": ("metric_idx", "metric_set_idx"),
    "evaluation_idx": ("evaluation_index",),
}


def _leafify(obj):
    """Tensors / python floats -> numpy; containers keep their type."""
    if obj is None:
        return None
    if isinstance(obj, (list, tuple)):
        return type(obj)(_leafify(x) for x in obj)
    if hasattr(obj, "detach"):
        return obj.detach().cpu().numpy()
    if isinstance(obj, (float, int, np.floating)):
        return np.asarray(obj, dtype=np.float64)
    return obj


class SimulationResults:
    def __init__(self, results, derivatives, second_derivatives, netting_set_names=None,
                 metric_names=None, model_param_names=None, product_names=None):
        self.results = _leafify(results)
        self.derivatives = _leafify(derivatives)
        self.second_derivatives = _leafify(second_derivatives)
        n_sets = len(self.results)
        n_metrics = len(self.results[0]) if n_sets else 0
        if netting_set_names is not None and product_names is not None and netting_set_names != product_names:
            raise ValueError("Provide either 'netting_set_names' or legacy alias 'product_names', not conflicting values.")
        names = netting_set_names if netting_set_names is not None else product_names
        self.netting_set_names = names if names is not None else [f"netting_set_{i}" for i in range(n_sets)]
        self.product_names = self.netting_set_names
        self.metric_names = metric_names if metric_names is not None else [f"metric_{i}" for i in range(n_metrics)]
        self.model_param_names = model_param_names if model_param_names is not None else []
        self._idx = {
            "netting set": {n.lower(): i for i, n in enumerate(self.netting_set_names)},
            "metric": {n.lower(): i for i, n in enumerate(self.metric_names)},
            "model parameter": {n.lower(): i for i, n in enumerate(self.model_param_names)},
        }

    # -- name resolution -----------------------------------------------------
    def _resolve(self, kind, key, available):
        if not isinstance(key, str):
            return key
        table = self._idx[kind]
        if key.lower() not in table:
            raise KeyError(f"Unknown {kind} name '{key}'. Available: {available}")
        return table[key.lower()]

    def _resolve_netting_set_idx(self, netting_set):
        return self._resolve("netting set", netting_set, self.netting_set_names)

    def _resolve_metric_idx(self, metric):
        return self._resolve("metric", metric, self.metric_names)

    def _resolve_param_idx(self, param):
        return self._resolve("model parameter", param, self.model_param_names)

    @staticmethod
    def _merge_legacy(current, kwargs):
        """Fold legacy keyword aliases into (netting_set, metric, evaluation_idx)."""
        out = {}
        for new_name, aliases in _LEGACY.items():
            value = None
            for alias in aliases:
                if alias in kwargs:
                    v = kwargs.pop(alias)
                    if value is None:
                        value = v
                    elif v != value:
                        raise ValueError(
                            f"Conflicting values provided for '{new_name}' and legacy alias '{alias}'.")
            out[new_name] = current[new_name] if current[new_name] is not None else value
        if kwargs:
            raise TypeError("Unexpected keyword argument(s): " + ", ".join(sorted(kwargs)))
        return out["netting_set"], out["metric"], out["evaluation_idx"]

    def _cell(self, table, netting_set, metric, evaluation_idx, legacy):
        ns, m, ev = self._merge_legacy(
            {"netting_set": netting_set, "metric": metric, "evaluation_idx": evaluation_idx}, legacy)
        return table[self._resolve_netting_set_idx(ns)][self._resolve_metric_idx(m)], ev

    # -- names ---------------------------------------------------------------
    def get_product_names(self):
        return list(self.netting_set_names)

    def get_netting_set_names(self):
        return list(self.netting_set_names)

    def get_metric_names(self):
        return list(self.metric_names)

    def get_model_param_names(self):
        return list(self.model_param_names)

    # -- values --------------------------------------------------------------
    def get_results(self, netting_set=None, metric=None, evaluation_idx=None, **legacy_kwargs):
        cell, ev = self._cell(self.results, netting_set, metric, evaluation_idx, legacy_kwargs)
        values = np.array([pair[0] for pair in cell])
        return values if ev is None else values[ev]

    def get_mc_error(self, netting_set=None, metric=None, evaluation_idx=None, **legacy_kwargs):
        cell, ev = self._cell(self.results, netting_set, metric, evaluation_idx, legacy_kwargs)
        errors = np.array([pair[1] for pair in cell])
        return errors if ev is None else errors[ev]

    def get_derivatives(self, netting_set=None, metric=None, param=None, evaluation_idx=None,
                        **legacy_kwargs):
        cell, ev = self._cell(self.derivatives, netting_set, metric, evaluation_idx, legacy_kwargs)
        if param is None and ev is None:
            return cell
        if ev is not None:
            row = cell[ev]
            if param is None:
                return {name: row[i] for i, name in enumerate(self.model_param_names)}
            return row[self._resolve_param_idx(param)]
        p = self._resolve_param_idx(param)
        return np.array([row[p] for row in cell])

    def get_second_derivatives(self, netting_set=None, metric=None, param1=None, param2=None,
                               evaluation_idx=None, **legacy_kwargs):
        cell, ev = self._cell(self.second_derivatives, netting_set, metric, evaluation_idx, legacy_kwargs)
        if param1 is None and param2 is None and ev is None:
            return cell
        names = self.model_param_names

        def named(row):
            return {n: row[i] for i, n in enumerate(names)}

        if ev is not None:
            hess = cell[ev]
            if param1 is None and param2 is None:
                return {n: named(hess[i]) for i, n in enumerate(names)}
            if param2 is None:
                return named(hess[self._resolve_param_idx(param1)])
            if param1 is None:
                c = self._resolve_param_idx(param2)
                return {n: hess[i][c] for i, n in enumerate(names)}
            return hess[self._resolve_param_idx(param1)][self._resolve_param_idx(param2)]
        if param1 is not None and param2 is not None:
            r, c = self._resolve_param_idx(param1), self._resolve_param_idx(param2)
            return np.array([h[r][c] for h in cell])
        raise ValueError("When evaluation_idx is omitted, provide both param1 and param2 or neither.")
