"""Differential operators, reduced to the two the shipped problems use.

The reference builds operators as higher-order functions over JAX autodiff
(src/pnmol/diffops.py:76-247); only ``laplace()`` (src/pnmol/pde/examples.py:53,167,237,323)
and ``gradient()`` (src/pnmol/discretize.py:128) ever reach the discretisation, both in
one spatial dimension.  Here they are tags that select closed-form kernel derivatives
(``pnmol_b200.kernels``).
"""
from collections import namedtuple


class DifferentialOperator(namedtuple("DifferentialOperator", "name")):
    def __repr__(self):
        return f"DifferentialOperator({self.name})"


def laplace():
    return DifferentialOperator("laplace")


def gradient():
    return DifferentialOperator("gradient")
