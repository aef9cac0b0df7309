"""Joint simulation of several correlated models
(reference: src/models/model_config.py:8-307)."""
import numpy as np
from models.model import *
from models.black_scholes import BlackScholesModel
from mcre.dual import D


class ModelConfig(Model):
    KIND = 100  # composite

    def __init__(self, models, numeraire_model_idx=0, discount_model_idx=0,
                 inter_asset_correlation_matrix=None):
        assert len(models) > 0, "Provide at least one model."
        assert all(float(models[i].calibration_date) == float(models[i + 1].calibration_date)
                   for i in range(len(models) - 1)), "All models must share the same calibration_date."
        asset_ids = [a for m in models for a in m.asset_ids]
        assert len(asset_ids) == len(set(asset_ids)), \
            "Duplicate asset_ids detected across models. A particular asset can only be simulated by one distinct model."
        super().__init__(calibration_date=float(models[0].calibration_date), asset_ids=asset_ids,
                         simulation_dim=sum(m.simulation_dim for m in models))
        self.state_dim = sum(m.state_dim for m in models)
        self.models = list(models)
        self.id_to_model = {"numeraire": numeraire_model_idx, "discount": discount_model_idx}
        for idx, m in enumerate(models):
            for a in m.asset_ids:
                self.id_to_model[a] = idx
        self.model_state_offset, off = {}, 0
        for idx, m in enumerate(models):
            self.model_state_offset[idx] = off
            off += m.state_dim
        self.model_params = [p for m in models for p in m.get_model_params()]
        self.inter_asset_correlation_matrix = []
        if inter_asset_correlation_matrix is None:
            for i, m1 in enumerate(models):
                for m2 in models[i + 1:]:
                    self.inter_asset_correlation_matrix.append(
                        torch.zeros(m1.num_assets, m2.num_assets, dtype=FLOAT, device=device))
        else:
            for c in inter_asset_correlation_matrix:
                self.inter_asset_correlation_matrix.append(torch.tensor(c, dtype=FLOAT, device=device))

    def get_model_param_names(self):
        names = []
        for m in self.models:
            label = m.asset_ids[0] if len(m.asset_ids) == 1 and m.asset_ids[0] else m.__class__.__name__
            names.extend(f"{label}.{n}" for n in m.get_model_param_names())
        return names

    def requires_grad(self):
        super().requires_grad()
        for m in self.models:
            m.requires_grad()

    def param_offsets(self):
        offs, o = [], 0
        for m in self.models:
            offs.append(o)
            o += len(m.model_params)
        return offs

    def joint_correlation(self, scheme, sub_params):
        """Block correlation over assets, pair blocks in (0,1),(0,2),..,(1,2).. order,
        symmetrised (reference: model_config.py:101-142).  ``sub_params[i]`` are the
        dual parameters of sub-model i."""
        n = self.num_assets
        nt = sub_params[0][0].t.shape[0]
        corr = [[D(0.0, None, nt) for _ in range(n)] for _ in range(n)]
        row, idx = 0, 0
        for i, m1 in enumerate(self.models):
            n1 = m1.num_assets
            blk = m1.intra_correlation(scheme, sub_params[i])
            if len(blk) != n1:
                raise RuntimeError(
                    f"{type(m1).__name__} has {len(blk)} noise dimensions for {n1} asset(s); "
                    "it cannot be placed in a ModelConfig (same restriction as the reference).")
            for a in range(n1):
                for b in range(n1):
                    corr[row + a][row + b] = blk[a][b]
            col = row + n1
            for m2 in self.models[i + 1:]:
                n2 = m2.num_assets
                ic = self.inter_asset_correlation_matrix[idx].numpy()
                up = np.broadcast_to(ic, (n1, n2))
                lo = np.broadcast_to(ic.T if ic.ndim >= 2 else ic, (n2, n1))
                for a in range(n1):
                    for b in range(n2):
                        corr[row + a][col + b] = D(float(up[a, b]), None, nt)
                        corr[col + b][row + a] = D(float(lo[b, a]), None, nt)
                col += n2
                idx += 1
            row += n1
        return [[0.5 * (corr[i][j] + corr[j][i]) for j in range(n)] for i in range(n)]
