"""GPU (-m gpu): the in-library multi-GPU context (ccgp_create_multi, csrc/multi.cu) against the single-GPU one.
Runs with however many GPUs the box shows: with one GPU the front/child fan-out is exercised without NCCL,
with two or more the (min, index) all-reduce goes through NCCL (gpurun --gpus N).  Per-candidate values must be
BIT-identical to the single-GPU context and every argmin index identical, for every GPU count."""
import os

import numpy as np
import pytest

import ccgp_b200
from ccgp_b200 import GAUSS_ANISO_LAMBDA, GAUSS_ISO, LOGSCALE, MEAN_ZERO_PLUS_TAU2, workloads
from ccgp_b200 import reference_api as api

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def multi():
    eng = ccgp_b200.Engine(n_gpus=int(os.environ.get("CCGP_TEST_GPUS", "0")))
    yield eng
    eng.close()


def test_multi_context_reports_its_gpus(multi):
    import torch
    want = int(os.environ.get("CCGP_TEST_GPUS", "0")) or torch.cuda.device_count()
    assert multi.n_gpus == want


def test_multi_nll_batch_and_argmin_match_single(engine, multi):
    X, y, s2 = workloads.m1_design()
    th = workloads.m1_candidates(20011)                       # not a multiple of anything
    engine.set_design(X, y)
    multi.set_design(X, y)
    a = engine.nll_batch(th, GAUSS_ANISO_LAMBDA, s2, scale=LOGSCALE)
    b = multi.nll_batch(th, GAUSS_ANISO_LAMBDA, s2, scale=LOGSCALE)
    for u, v in zip(a, b):
        assert np.array_equal(u, v)
    c0 = multi.collective_count
    v1, i1 = engine.nll_argmin(th, GAUSS_ANISO_LAMBDA, s2, scale=LOGSCALE)
    v2, i2 = multi.nll_argmin(th, GAUSS_ANISO_LAMBDA, s2, scale=LOGSCALE)
    assert (v1, i1) == (v2, i2) == (a[0].min(), int(a[0].argmin()))
    assert multi.collective_count - c0 == (2 if multi.n_gpus > 1 else 0)
    # ties: duplicate the best row at a higher index -> the lower index still wins, on every GPU count
    th2 = np.vstack([th, th[i1:i1 + 1]])
    assert multi.nll_argmin(th2, GAUSS_ANISO_LAMBDA, s2, scale=LOGSCALE)[1] == i1
    # NaN rows never win
    th3 = th.copy()
    th3[i1, 0] = np.nan
    v3, i3 = multi.nll_argmin(th3, GAUSS_ANISO_LAMBDA, s2, scale=LOGSCALE)
    assert i3 == int(np.nanargmin(np.where(np.arange(len(th)) == i1, np.nan, a[0])))


def test_multi_sweep_and_predict_match_single(engine, multi, designs):
    he, hp = designs["he_train"], designs["he_hyperpars"]
    X, y = he[:, :4], he[:, 4]
    cand = np.vstack([api.sweep_candidates(hp[r, 0:2], hp[r, 2:4], 1000) for r in (0, 311, 623)])
    engine.set_design(X, y)
    multi.set_design(X, y)
    a = engine.nll_batch(cand, GAUSS_ISO, 30.0, mean_mode=MEAN_ZERO_PLUS_TAU2, tau=50.0)
    b = multi.nll_batch(cand, GAUSS_ISO, 30.0, mean_mode=MEAN_ZERO_PLUS_TAU2, tau=50.0)
    assert np.array_equal(a[0], b[0])
    pars = cand[::100]
    ma, va, _ = engine.predict(pars, GAUSS_ISO, designs["he_test"][:, :4], 30.0)
    mb, vb, _ = multi.predict(pars, GAUSS_ISO, designs["he_test"][:, :4], 30.0)
    assert np.array_equal(ma, mb) and np.array_equal(va, vb)
    # factors kept on the device (ccgp_factors_*): rows sliced over the GPUs, same tables
    fac = multi.factors(pars, GAUSS_ISO)
    assert fac.info()["stored"] and fac.info()["rows"] == len(pars)
    mc, vc, _ = fac.predict(designs["he_test"][:, :4], 30.0)
    assert np.array_equal(ma, mc) and np.array_equal(va, vc)
    fac.close()


def test_multi_me_argmin_both_splits(engine, multi):
    D_old, pool = workloads.me_pool()
    params = workloads.me_params(37)
    bv, bi = engine.me_argmin(D_old, pool, params)
    mv, mi = multi.me_argmin(D_old, pool, params)               # parameter rows split over the GPUs: no collective
    assert np.array_equal(bi, mi) and np.array_equal(bv, mv)
    nd, _, _ = multi.me_schur_batch(D_old, pool[:301], params[:5])
    nd1, _, _ = engine.me_schur_batch(D_old, pool[:301], params[:5])
    assert np.array_equal(nd, nd1)
    os.environ["CCGP_MULTI_ME_SPLIT_DESIGNS"] = "1"             # designs split: the per-row (min, index) goes through NCCL
    try:
        c0 = multi.collective_count
        sv, si = multi.me_argmin(D_old, pool, params)
        assert multi.collective_count - c0 == (2 if multi.n_gpus > 1 else 0)
    finally:
        os.environ.pop("CCGP_MULTI_ME_SPLIT_DESIGNS", None)
    assert np.array_equal(bi, si) and np.array_equal(bv, sv)
    one, _, _ = multi.me_schur_batch(D_old, pool, params[:1])   # fewer rows than GPUs: designs are scattered
    assert np.array_equal(one[:, 0], engine.me_schur_batch(D_old, pool, params[:1])[0][:, 0])
