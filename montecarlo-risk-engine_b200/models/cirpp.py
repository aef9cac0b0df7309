"""CIR++ default-intensity model: lambda(t) = y(t) + psi(t)
(reference: src/models/cirpp.py:6-317, src/helpers/cs_helper.py:80-107).

All deterministic ingredients (market hazard, psi shift, CIR A/B functions, the
conditional-survival prefactor) are scalars per date; they are evaluated here on
the host as dual numbers and shipped to the kernels as per-step / per-date
tables."""
import math
import numpy as np
from models.model import *
from mcre.dual import D, dexp, dsqrt, dlog, dval


class CIRPPModel(Model):
    KIND = 4  # MCRE_MODEL_CIRPP

    def __init__(self, calibration_date, asset_id, hazard_rates, kappa, theta, volatility, y0,
                 deterministic=False):
        super().__init__(calibration_date=calibration_date, state_dim=2, asset_ids=[asset_id])
        assert 2 * kappa * theta - volatility ** 2 > 0 and y0 > 0, "Feller condition not met."
        # parameter order: [kappa, theta, sigma, y0]
        self.model_params = [torch.tensor(float(v), dtype=FLOAT, device=device)
                             for v in (kappa, theta, volatility, y0)]
        self.tenors = torch.tensor(list(hazard_rates.keys()), dtype=FLOAT, device=device)
        self.hazard_rates = torch.tensor(list(hazard_rates.values()), dtype=FLOAT, device=device)
        self.deterministic = deterministic

    def get_kappa(self):
        return torch.stack([self.model_params[0]])

    def get_theta(self):
        return torch.stack([self.model_params[1]])

    def get_sigma(self):
        return torch.stack([self.model_params[2]])

    def get_y0(self):
        return torch.stack([self.model_params[3]])

    def get_model_param_names(self):
        return ["kappa", "theta", "sigma", "y0"]

    # -- market curve (not model parameters: no tangents) ---------------------
    def market_hazard(self, t):
        """Piecewise-constant hazard, right-closed buckets, flat after the last tenor."""
        ten, haz = self.tenors.tolist(), self.hazard_rates.tolist()
        for tt, h in zip(ten, haz):
            if t <= tt:
                return h
        return haz[-1]

    def market_survival(self, t):
        """S_m(0,t) from the piecewise-constant hazards (cs_helper.py:80-107)."""
        import math
        ten, haz = self.tenors.tolist(), self.hazard_rates.tolist()
        surv, prev, idx = 1.0, 0.0, len(ten) - 1
        for i, mat in enumerate(ten):
            if mat <= t:
                surv *= math.exp(-haz[i] * (mat - prev))
                prev = mat
            else:
                idx = i
                break
        dt = t - prev
        if dt > 0:
            surv *= math.exp(-haz[idx] * dt)
        return surv

    def _market_survival_probability(self, t):
        """Tensor form of market_survival under the reference's (private) name, which its scenario script calls
        (tests/exposure_tests/cirpp_scenarios_vs_deterministic_hazard.py:93)."""
        return torch.tensor(self.market_survival(float(torch.as_tensor(t).reshape(-1)[0])), dtype=FLOAT, device=device)

    # -- CIR building blocks (reference: cirpp.py:85-142) ---------------------
    @staticmethod
    def _h(p):
        return dsqrt(p[0] * p[0] + 2.0 * p[2] * p[2])

    def _memo(self, p):
        """Per-parameter-set cache of A(tau) / B(tau): the plan compiler asks for the same
        maturities again and again (adjacent exposure dates share an end point)."""
        key = tuple(id(x) for x in p)
        cache = getattr(self, "_ab_cache", None)
        if cache is None or cache[0] != key:
            cache = (key, {}, {}, p)          # keeps p alive so the ids stay valid
            self._ab_cache = cache
        return cache

    def cir_A(self, p, tau):
        memo = self._memo(p)[1]
        if tau in memo:
            return memo[tau]
        kappa, theta, sigma = p[0], p[1], p[2]
        h = self._h(p)
        num = 2.0 * h * dexp(0.5 * (kappa + h) * tau)
        den = 2.0 * h + (kappa + h) * (dexp(h * tau) - 1.0)
        memo[tau] = (num / den) ** ((2.0 * kappa * theta) / (sigma * sigma))
        return memo[tau]

    def cir_B(self, p, tau):
        memo = self._memo(p)[2]
        if tau in memo:
            return memo[tau]
        kappa = p[0]
        h = self._h(p)
        e = dexp(h * tau) - 1.0
        memo[tau] = (2.0 * e) / (2.0 * h + (kappa + h) * e)
        return memo[tau]

    def lambda_t(self, t, y_t):
        """Model intensity lambda(t) = y(t) + psi(t); in deterministic mode the state already is the market hazard
        (reference: cirpp.py:240-244).  Host helper."""
        y_t = torch.as_tensor(y_t, dtype=FLOAT)
        if self.deterministic:
            return y_t
        return y_t + float(self.psi(self.dual_params(), float(torch.as_tensor(t).reshape(-1)[0])).v)

    def psi(self, p, t):
        """Deterministic shift psi(t) = lambda_mkt(t) + D(t) - y0 E(t)."""
        kappa, theta, sigma, y0 = p
        h = self._h(p)
        et = dexp(h * t)
        den = 2.0 * h + (kappa + h) * (et - 1.0)
        Dt = (2.0 * kappa * theta / (sigma * sigma)) * (0.5 * (kappa + h) - (h * (kappa + h) * et) / den)
        Et = (4.0 * h * h * et) / (den * den)
        return self.market_hazard(t) + Dt - y0 * Et

    # -- value-only, vectorised over dates (plans without sensitivities are lowered on every run) --
    def _market_vec(self, ts, what):
        f = self.market_hazard if what == "hazard" else self.market_survival
        return np.array([f(float(t)) for t in ts])

    def psi_values(self, pv, ts):
        """psi(t) for an array of times, same formulas as psi()."""
        kappa, theta, sigma, y0 = pv
        ts = np.asarray(ts, dtype=np.float64)
        h = math.sqrt(kappa * kappa + 2.0 * sigma * sigma)
        et = np.exp(h * ts)
        den = 2.0 * h + (kappa + h) * (et - 1.0)
        Dt = (2.0 * kappa * theta / (sigma * sigma)) * (0.5 * (kappa + h) - (h * (kappa + h) * et) / den)
        Et = (4.0 * h * h * et) / (den * den)
        return self._market_vec(ts, "hazard") + Dt - y0 * Et

    def conditional_survival_values(self, pv, t, T):
        """(C, B) arrays of S(t,T | y) = C exp(-B y) for arrays of (t, T), same formulas as
        conditional_survival_coefficients()."""
        kappa, theta, sigma, y0 = pv
        t, T = np.asarray(t, dtype=np.float64), np.asarray(T, dtype=np.float64)
        h = math.sqrt(kappa * kappa + 2.0 * sigma * sigma)

        def A(tau):
            num = 2.0 * h * np.exp(0.5 * (kappa + h) * tau)
            den = 2.0 * h + (kappa + h) * (np.exp(h * tau) - 1.0)
            return (num / den) ** ((2.0 * kappa * theta) / (sigma * sigma))

        def Bf(tau):
            e = np.exp(h * tau) - 1.0
            return (2.0 * e) / (2.0 * h + (kappa + h) * e)

        pref = (self._market_vec(T, "survival") / self._market_vec(t, "survival")) * (A(t) / A(T)) \
            * np.exp(Bf(T) * y0 - Bf(t) * y0)
        return pref * A(T - t), Bf(T - t)

    def conditional_survival_coefficients(self, p, t, T):
        """(C, B) with S(t,T | y_t) = C * exp(-B * y_t) (reference: cirpp.py:246-285).
        Deterministic mode: S = S_m(0,T)/S_m(0,t), i.e. (ratio, 0)."""
        nt = p[0].t.shape[0]
        if self.deterministic:
            return D(self.market_survival(T) / self.market_survival(t), None, nt), D(0.0, None, nt)
        y0 = p[3]
        A0t, A0T = self.cir_A(p, t), self.cir_A(p, T)
        B0t, B0T = self.cir_B(p, t), self.cir_B(p, T)
        pref = (self.market_survival(T) / self.market_survival(t)) * (A0t / A0T) * dexp(B0T * y0 - B0t * y0)
        return pref * self.cir_A(p, T - t), self.cir_B(p, T - t)

    # -- host helpers used by tests (closed form, tensors in / out) -----------
    def survival_probability(self, t, T, y_t):
        t, T = float(torch.as_tensor(t)), float(torch.as_tensor(T))
        C, B = self.conditional_survival_coefficients(self.dual_params(), t, T)
        y = torch.as_tensor(y_t, dtype=FLOAT)
        return C.v * torch.exp(-B.v * y)
