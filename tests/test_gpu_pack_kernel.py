"""GPU (-m gpu): the packed-residency NLL kernel (csrc/factor_pack.cuh; default for 2-D designs of 79..102 points, CCGP_KERNEL=5 elsewhere) against the oracle's golden
values and against the team / warp kernels on large seeded batches (the soak that caught wrong values in experimental
builds of this kernel: a few percent of the candidates off by 1e-2, different from run to run)."""
import os

import numpy as np
import pytest

from ccgp_b200 import GAUSS_ANISO_LAMBDA, LOGSCALE, workloads
from conftest import rel_err

pytestmark = pytest.mark.gpu
TOL = 1e-10


@pytest.fixture()
def packed_kernel():
    old = os.environ.get("CCGP_KERNEL")
    os.environ["CCGP_KERNEL"] = "5"
    yield
    if old is None:
        os.environ.pop("CCGP_KERNEL", None)
    else:
        os.environ["CCGP_KERNEL"] = old


def test_packed_kernel_golden_n100(engine, golden, designs, packed_kernel):
    engine.set_design(designs["maximin100"], golden["c1n100_y"])
    nll, beta, st = engine.nll_batch(golden["c1n100_nat"], GAUSS_ANISO_LAMBDA, 1.0)
    assert engine.last_nll_config()["variant"] >= 500          # the packed kernel really ran
    assert np.all(st == 0)
    assert rel_err(-nll, golden["c1n100_ref"]).max() < TOL
    assert rel_err(beta, golden["c1n100_beta"]).max() < TOL


@pytest.mark.parametrize("n,d,logs", [(100, 2, 1), (104, 2, 0), (96, 2, 0), (97, 3, 0), (64, 4, 0), (50, 9, 0), (21, 2, 0)])
def test_packed_kernel_soak_vs_shipped_choice(engine, n, d, logs):
    rng = np.random.default_rng(100 * n + d)
    B = 1 << 16
    if (n, d) == (100, 2):
        X, y, s2 = workloads.m1_design()
    else:
        X = rng.uniform(-1, 1, (n, d)); y = rng.normal(size=n); s2 = 1.0
    engine.set_design(X, y)
    if logs:
        th, scale = workloads.m1_candidates(B), LOGSCALE
    else:
        t0 = 8.0 / d * n ** (1.0 / d)
        th = np.column_stack([rng.uniform(0.2, 0.8, B)] + [rng.uniform(0.5 * t0, 1.5 * t0, B) for _ in range(d)] + [rng.uniform(0.5, 3.0, B)])
        scale = 0
    os.environ.pop("CCGP_KERNEL", None)
    os.environ["CCGP_NO_PACK"] = "1"                                 # the team / warp kernels (the packed one is the default for d = 2, n = 79..102)
    try:
        ref, rbeta, rst = engine.nll_batch(th, GAUSS_ANISO_LAMBDA, s2, scale=scale)
    finally:
        os.environ.pop("CCGP_NO_PACK", None)
    assert engine.last_nll_config()["variant"] < 500
    os.environ["CCGP_KERNEL"] = "5"
    try:
        for rep in range(2):
            nll, beta, st = engine.nll_batch(th, GAUSS_ANISO_LAMBDA, s2, scale=scale)
            assert engine.last_nll_config()["variant"] >= 500
            assert np.array_equal(st, rst)                                   # same candidates refused
            ok = rst == 0
            # two correct FP64 factorisations with different accumulation orders: kappa * eps apart
            assert rel_err(nll[ok], ref[ok]).max() < 1e-9 and rel_err(beta[ok], rbeta[ok]).max() < 1e-9
            if rep:
                assert np.array_equal(nll, first, equal_nan=True)              # run-to-run bit-identical
            first = nll
    finally:
        os.environ.pop("CCGP_KERNEL", None)


@pytest.mark.parametrize("n,d,logs,B", [(100, 2, 1, (1 << 15) + 37), (97, 3, 0, 4099), (96, 2, 0, 5), (50, 9, 0, 1 << 14), (21, 2, 0, 2000)])
def test_producer_consumer_kernel_bit_identical_to_packed(engine, n, d, logs, B):
    """csrc/factor_pc.cuh (CCGP_KERNEL=6): the packed kernel with the column assembly moved to producer warps (setmaxnreg,
    named barriers, a slot plan with separate raw and solved slots).  Same arithmetic tile by tile, so the values must
    be bit-identical to the packed kernel's, for batch sizes that leave consumers without work in the last CTA too."""
    rng = np.random.default_rng(7 * n + d)
    if (n, d) == (100, 2):
        X, y, s2 = workloads.m1_design()
    else:
        X = rng.uniform(-1, 1, (n, d)); y = rng.normal(size=n); s2 = 1.0
    engine.set_design(X, y)
    if logs:
        th, scale = workloads.m1_candidates(B), LOGSCALE
    else:
        t0 = 8.0 / d * n ** (1.0 / d)
        th = np.column_stack([rng.uniform(0.2, 0.8, B)] + [rng.uniform(0.5 * t0, 1.5 * t0, B) for _ in range(d)] + [rng.uniform(0.5, 3.0, B)])
        scale = 0
    old = os.environ.get("CCGP_KERNEL")
    try:
        os.environ["CCGP_KERNEL"] = "5"
        ref, rbeta, rst = engine.nll_batch(th, GAUSS_ANISO_LAMBDA, s2, scale=scale)
        assert 500 <= engine.last_nll_config()["variant"] < 600
        for split in ("0", "1", "2"):
            os.environ["CCGP_KERNEL"] = "6"
            os.environ["CCGP_PC_SPLIT"] = split
            nll, beta, st = engine.nll_batch(th, GAUSS_ANISO_LAMBDA, s2, scale=scale)
            assert engine.last_nll_config()["variant"] >= 600
            assert np.array_equal(st, rst)
            assert np.array_equal(nll, ref, equal_nan=True) and np.array_equal(beta, rbeta, equal_nan=True)
    finally:
        os.environ.pop("CCGP_PC_SPLIT", None)
        if old is None:
            os.environ.pop("CCGP_KERNEL", None)
        else:
            os.environ["CCGP_KERNEL"] = old
