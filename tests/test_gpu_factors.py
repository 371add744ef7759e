"""GPU (-m gpu): Cholesky factors of the posterior rows kept on the device (ccgp_factors_*, include/ccgp.h) -- the
device-side `factors.frame` ([A]:572-592) + `prediction` ([A]:637-654).  The stored-factor path runs the same site phase
on the same factor as ccgp_predict, so the tables must be bit-identical to a direct call; golden tables pin both."""
import os

import numpy as np
import pytest

from ccgp_b200 import reference_api as api
from ccgp_b200 import CcgpError, GAUSS_ANISO_LAMBDA, GAUSS_ISO, GAUSS_ISO_RAW2, MATERN1D, workloads
from conftest import rel_err

pytestmark = pytest.mark.gpu
TOL = 1e-10


def _pars(rng, family, S, n, d):
    scale = 3.0 * n ** (1.0 / d)
    if family == GAUSS_ANISO_LAMBDA:
        return np.column_stack([rng.uniform(0.1, 0.9, S)] + [scale * rng.uniform(1, 3, S) for _ in range(d)] + [rng.uniform(0.3, 2, S)])
    return np.column_stack([rng.uniform(0.1, 0.9, S), scale * rng.uniform(1, 3, S), scale * rng.uniform(2, 6, S)])


@pytest.mark.parametrize("n,d,family,S,T", [(14, 2, GAUSS_ANISO_LAMBDA, 7, 33), (64, 4, GAUSS_ISO, 1000, 14), (50, 9, GAUSS_ISO, 300, 150),
                                           (100, 2, GAUSS_ANISO_LAMBDA, 64, 625), (100, 2, GAUSS_ANISO_LAMBDA, 256, 625), (100, 2, GAUSS_ANISO_LAMBDA, 3, 1000),
                                           (90, 9, GAUSS_ISO_RAW2, 40, 110), (120, 3, GAUSS_ISO, 5, 70)])
def test_stored_factors_reproduce_direct_prediction(engine, n, d, family, S, T):
    rng = np.random.default_rng(31 * n + S)
    X = rng.uniform(-1, 1, (n, d))
    y = np.sin(2 * X[:, 0]) + 0.1 * rng.normal(size=n)
    pars = _pars(rng, family, S, n, d)
    if S > 3:
        pars[3, 1:] *= 1e-19                                   # singular row: NaN column, status 1 on both paths
    pv, vf = None, -1
    if family == GAUSS_ISO_RAW2:                               # quirk Q2: the vector uses theta1 * (1 + lambda)
        pv = np.column_stack([pars[:, 0], pars[:, 1], pars[:, 1] * (1.0 + pars[:, 2])])
        vf = GAUSS_ISO
    engine.set_design(X, y)
    fac = engine.factors(pars, family, pars_vec=pv, vec_family=vf)
    info = fac.info()
    assert info["rows"] == S and info["stored"] and info["device_bytes"] >= S * n * (n + 1) // 2 * 8
    for rep, (Tn, s2) in enumerate([(T, 2.5), (max(T // 3, 1), 0.7), (T, 2.5)]):           # any number of site sets per factorisation
        Xn = np.vstack([rng.uniform(-1.2, 1.2, (Tn - 1, d)), X[:1]]) if Tn > 1 else X[:1]
        m0, v0, st0 = engine.predict(pars, family, Xn, s2, pars_vec=pv, vec_family=vf)
        m1, v1, st1 = fac.predict(Xn, s2)
        assert np.array_equal(st0, st1)
        assert np.array_equal(m0, m1, equal_nan=True) and np.array_equal(v0, v1, equal_nan=True)      # bit-identical
        if S > 3:
            assert st1[3] == 1 and np.all(np.isnan(m1[:, 3]))
        if family != GAUSS_ISO_RAW2:                           # (Q2 rows use another correlation vector: no interpolation)
            ok = st1 == 0
            assert np.abs(m1[-1, ok] - y[0]).max() < 1e-5 and np.abs(v1[-1, ok]).max() < 1e-5 * s2   # the last site is a design point
    fac.close()
    fac.close()                                                # idempotent


def test_stored_factors_golden_tables(engine, golden, designs):
    engine.set_design(designs["maximin14"], golden["pred14_y"])
    fac = engine.factors(golden["pred14_pars"], GAUSS_ANISO_LAMBDA)
    m, v, st = fac.predict(golden["pred14_Xnew"], 0.9)
    assert np.all(st == 0)
    assert rel_err(m, golden["pred14_mean"]).max() < TOL
    assert np.abs(v - golden["pred14_var"]).max() < TOL * 0.9
    engine.set_design(designs["maximin14"], golden["pred14_y"])            # the identical design: still valid
    assert np.array_equal(fac.predict(golden["pred14_Xnew"], 0.9)[0], m)
    engine.set_design(0.5 * designs["maximin14"], golden["pred14_y"])
    with pytest.raises(CcgpError):                             # the design changed under the stored factors
        fac.predict(golden["pred14_Xnew"], 0.9)
    fac.close()
    he, het = designs["he_train"], designs["he_test"]
    engine.set_design(he[:, :4], he[:, 4])
    fac = api.factors_device(he[:, :4], he[:, 4], golden["predHE_pars"], script="H", engine=engine)
    m, v = api.predict_post_factors(fac, het[:, :4], 30.0)
    assert rel_err(m, golden["predHE_mean"]).max() < TOL
    assert np.abs(v - golden["predHE_var"]).max() / 30.0 < TOL
    fac.close()
    # quirk Q2 ([V]:672): the matrix rows use `lambda`, the vector theta1 (1 + lambda)
    fac = api.factors_device(designs["maximin14"], golden["pred14_y"], golden["predV_pars"], script="V", engine=engine)
    m, v = api.predict_post_factors(fac, golden["pred14_Xnew"], 0.9)
    assert rel_err(m, golden["predV_mean"]).max() < TOL
    assert np.abs(v - golden["predV_var"]).max() < TOL * 0.9
    fac.close()


def test_few_rows_many_sites_split_over_ctas(engine):
    """ccgp_predict with few posterior rows and many sites factors into a scratch buffer and splits the sites of a row over
    the idle CTAs (two launches): same values as the one-launch path (CCGP_PREDICT_NOSPLIT=1), bit for bit."""
    rng = np.random.default_rng(77)
    X, y, _ = workloads.m1_design()
    engine.set_design(X, y)
    for S, T in ((1, 625), (5, 2000), (40, 130)):
        pars = _pars(rng, GAUSS_ANISO_LAMBDA, S, 100, 2)
        g = rng.uniform(-1, 1, (T, 2))
        l0 = engine.launch_count
        a = engine.predict(pars, GAUSS_ANISO_LAMBDA, g, 1.7)
        assert engine.launch_count - l0 == 2
        os.environ["CCGP_PREDICT_NOSPLIT"] = "1"
        try:
            l0 = engine.launch_count
            b = engine.predict(pars, GAUSS_ANISO_LAMBDA, g, 1.7)
            assert engine.launch_count - l0 == 1
        finally:
            os.environ.pop("CCGP_PREDICT_NOSPLIT", None)
        assert np.array_equal(a[0], b[0]) and np.array_equal(a[1], b[1]) and np.array_equal(a[2], b[2])


def test_factors_fall_back_to_refactoring(engine):
    """1-D Matern rows (no tensor-path kernel) and CCGP_FACTORS_OFF keep only the parameters: same tables, nothing stored."""
    rng = np.random.default_rng(5)
    X = np.sort(rng.uniform(0, 1, (8, 1)), axis=0)
    y = np.sin(6 * X[:, 0])
    pars = np.column_stack([rng.uniform(0.2, 0.8, 6), rng.uniform(0.3, 1.0, 6), rng.uniform(0.3, 1.0, 6)])
    Xn = np.linspace(0, 1, 50)[:, None]
    engine.set_design(X, y)
    fac = engine.factors(pars, MATERN1D)
    assert not fac.info()["stored"]
    m0, v0, st0 = engine.predict(pars, MATERN1D, Xn, 1.3)
    m1, v1, st1 = fac.predict(Xn, 1.3)
    assert np.array_equal(m0, m1) and np.array_equal(v0, v1) and np.array_equal(st0, st1)
    fac.close()
    Xd, yd, _ = workloads.m1_design()
    engine.set_design(Xd, yd)
    p2 = _pars(rng, GAUSS_ANISO_LAMBDA, 9, 100, 2)
    os.environ["CCGP_FACTORS_OFF"] = "1"
    try:
        fac = engine.factors(p2, GAUSS_ANISO_LAMBDA)
    finally:
        os.environ.pop("CCGP_FACTORS_OFF", None)
    assert not fac.info()["stored"]
    g = rng.uniform(-1, 1, (40, 2))
    a, b = engine.predict(p2, GAUSS_ANISO_LAMBDA, g, 1.0), fac.predict(g, 1.0)
    assert np.array_equal(a[0], b[0]) and np.array_equal(a[1], b[1])
    fac.close()
