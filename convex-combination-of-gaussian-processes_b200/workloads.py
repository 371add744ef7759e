"""Seeded synthetic workloads of SURVEY 8(d), shared by bench.py and tests.
The designs are the reference's shipped data inputs, read from the committed
package data file data/reference_designs.npz (made by tests/golden/make_golden.py from the reference's shipped input files)."""
from __future__ import annotations

import os
import numpy as np

_ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
DESIGNS_PATH = os.path.join(os.path.dirname(os.path.abspath(__file__)), "data", "reference_designs.npz")
_designs = None


def designs():
    global _designs
    if _designs is None:
        _designs = dict(np.load(DESIGNS_PATH))
    return _designs


def test_function_4(X):
    """[A]:338 simulator 4: (sin 2x + cos 4x)(sin 8y + cos 4y)."""
    x, y = X[:, 0], X[:, 1]
    return (np.sin(2 * x) + np.cos(4 * x)) * (np.sin(8 * y) + np.cos(4 * y))


def m1_design():
    """M1 primary: maximin 100 pts ([-1,1]^2), y = simulator 4, sigma2 = 1."""
    X = designs()["maximin100"]
    return X, test_function_4(X), 1.0


def m1_candidates(B, seed=20131):
    """B real-line rows (psi1, psi2, phi, zeta): psi ~ N(log 20, .5^2), phi ~ N(0,1), zeta ~ N(0,.5^2)."""
    rng = np.random.default_rng(seed)
    return np.column_stack([rng.normal(np.log(20), 0.5, B), rng.normal(np.log(20), 0.5, B),
                            rng.normal(0, 1, B), rng.normal(0, 0.5, B)])


def me_pool():
    """ME-A: D.old = Initial ME Design (14x2), pool = the 1000 7-point blocks of All_Subdesigns."""
    d = designs()
    return d["me_initial14"], d["me_all_subdesigns"]


def me_params(P, seed=7):
    """Row 0 = the prior medians (0.5, 1, 4) of [M]:981-983, then P-1 draws p~U(0,1),
    theta1~IG(3,2), theta2~IG(5,16)."""
    rng = np.random.default_rng(seed)
    rows = [[0.5, 1.0, 4.0]]
    if P > 1:
        q = P - 1
        rows = np.vstack([rows, np.column_stack([rng.uniform(0, 1, q), 1.0 / rng.gamma(3, 1 / 2.0, q),
                                                 1.0 / rng.gamma(5, 1 / 16.0, q)])])
    return np.asarray(rows, dtype=np.float64)


def synthetic_pool(N=2048, seed=2048):
    """ME-B: N-point random LHS on [-1,1]^2 (same construction as make_golden.py)."""
    rng = np.random.default_rng(seed)
    return (np.column_stack([rng.permutation(N), rng.permutation(N)]) + rng.uniform(size=(N, 2))) / N * 2 - 1
