"""Product base class (reference: src/products/product.py:12-217).

Products are contract descriptions.  Unlike the reference they carry no
per-path tensors or request handles; `mcre.lowering` turns them into per-date
cashflow / exercise tables for the kernels."""
from __future__ import annotations
from enum import Enum
from common.packages import *


class OptionType(Enum):
    CALL = 1
    PUT = 2


class SettlementType(Enum):
    PHYSICAL = 0
    CASH = 1


class ProductFamily(Enum):
    GENERIC = "generic"
    VANILLA_TERMINAL_OPTION = "vanilla_terminal_option"
    BINARY_TERMINAL_PAYOFF = "binary_terminal_payoff"
    BASKET_TERMINAL_PAYOFF = "basket_terminal_payoff"
    ASIAN_PATH_TERMINAL = "asian_path_terminal"
    BARRIER_PATH_TERMINAL = "barrier_path_terminal"
    BERMUDAN_EXERCISE = "bermudan_exercise"
    FLEXICALL_EXERCISE = "flexicall_exercise"


def _ft(values):
    return torch.tensor([float(v) for v in values], dtype=FLOAT, device=device)


class Product:
    def __init__(self, asset_ids=None, product_id=0, product_family=ProductFamily.GENERIC):
        self.asset_ids = asset_ids if asset_ids else [""]
        self.product_id = product_id
        self.name = None
        self.product_family = product_family
        self.product_timeline = None
        self.modeling_timeline = None
        self.regression_timeline = _ft([])
        self.regression_coeffs = None

    def get_num_states(self):
        return 1

    def get_initial_state(self):
        return 0

    def get_state_dtype(self):
        return torch.long

    def get_asset_id(self, id=None):
        return self.asset_ids[id] if id else self.asset_ids[0]

    def get_name(self):
        return self.name if self.name else self.__class__.__name__

    def get_product_family(self):
        return self.product_family

    def _allocate_regression_coeffs(self, regression_function):
        self.regression_coeffs = torch.zeros(
            (len(self.regression_timeline), self.get_num_states(), regression_function.get_degree()),
            dtype=FLOAT, device=device)

    # analytic hooks (reference: product.py:199-217)
    def compute_pv_analytically(self, model):
        raise NotImplementedError

    def supports_analytic_pv(self, model):
        return False

    def supports_analytic_exposure(self, model):
        return False
