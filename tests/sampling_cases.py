"""Cases for the design-space sampling / constraint helpers (util.py:57-214 of the reference), shared by
oracle/gen_sampling.py (verbatim reference run) and tests/test_oracle_golden.py (the in-tree mirror)."""
import numpy as np


def sum_above(x):
    return (np.sum(x, axis=1) > 1.2).astype(float)


def first_small(x):
    return (x[:, 0] < 0.1).astype(float)


def broken(x):
    raise RuntimeError("a constraint function that fails is skipped by the reference (util.py:164-169)")


EVALUATED = [np.array([[0.5, 0.5], [0.25, 0.75]]), np.array([[0.5, 0.5], [0.1, 0.1]])]

# (name, seed, positional (samples, ndim), keywords)
CASES = [
    ("lhs_plain", 0, (12, 3), dict(resolution=5)),
    ("lhs_bounds", 1, (15, 2), dict(resolution=5, lb=[0.1, 0.2], ub=[0.9, 0.7])),
    ("lhs_linear", 2, (10, 2), dict(resolution=4, A=[1.0, 2.0], b=[0.3, 0.5])),
    ("lhs_funcs", 3, (10, 3), dict(resolution=5, func=[sum_above, first_small])),
    ("lhs_single_func", 4, (8, 2), dict(func=sum_above)),
    ("lhs_broken_func", 5, (6, 2), dict(resolution=5, func=[broken, first_small])),
    ("grid", 6, (20, 2), dict(sampleScheme="Grid", resolution=5)),
    ("grid_bounds_no_opt", 7, (30, 2), dict(sampleScheme="Grid", lb=[0.2, 0.0], ub=[1.0, 0.8], opt_sample_size=False)),
    ("compfunc", 8, (9, 3), dict(sampleScheme="CompFunc", resolution=5)),
    ("tight_bounds_retry", 9, (20, 2), dict(resolution=5, ub=[0.3, 0.3])),
    ("evaluated_points", 10, (5, 2), dict(sampleScheme="Grid", resolution=2, evaluatedPoints=EVALUATED, opt_sample_size=False)),
    ("wrong_length_bounds", 11, (7, 3), dict(resolution=5, lb=[0.5, 0.5])),
    ("equality_rows", 12, (6, 2), dict(resolution=5, Aeq=[1.0, 1.0], beq=[0.2, 0.2])),
]
