"""Equity family backend (Black-Scholes single / multi asset, ModelConfig of BS models,
Heston) - placeholder until csrc/equity.cu lands in this round."""


class EquityBackend:
    @staticmethod
    def supports(ctrl):
        return False
