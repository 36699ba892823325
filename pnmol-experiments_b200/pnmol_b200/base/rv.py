"""Random variables (src/pnmol/base/rv.py:9-14)."""
from collections import namedtuple


class MultivariateNormal(namedtuple("_MultivariateNormal", "mean cov_sqrtm")):
    @property
    def cov(self):
        return self.cov_sqrtm @ self.cov_sqrtm.transpose(-1, -2)
