"""Regression basis description (reference: src/maths/regression.py:3-14).

Only the *description* lives on the host: the design matrix is never built; the
pre-simulation kernel accumulates the Gram moments of this basis directly."""
import torch


class RegressionFunction:
    def __init__(self, degree):
        self.degree = degree

    def get_degree(self):
        # number of basis functions (the reference calls this "degree")
        return self.degree + 1


class PolyomialRegression(RegressionFunction):
    def __init__(self, degree):
        super().__init__(degree)

    def get_regression_matrix(self, explanatory_variables):
        x = torch.as_tensor(explanatory_variables)
        return torch.stack([x ** k for k in range(self.degree + 1)], dim=1)
