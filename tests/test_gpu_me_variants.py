"""The ME kernel's schedule and S-block variants (CCGP_ME_BALANCED, CCGP_ME_SYM: read at every launch) give the same
bits as the first version over every shape the kernel is unrolled for, and the values match the oracle's
`Augmented.Mixed.Entropy` ([M]:869-877) / `Entropy` ([M]:856-861)."""
import os

import numpy as np
import pytest

from oracle import ccgp_oracle as orc

pytestmark = pytest.mark.gpu


def _with_env(env, fn):
    old = {k: os.environ.get(k) for k in env}
    os.environ.update(env)
    try:
        return fn()
    finally:
        for k, v in old.items():
            if v is None:
                os.environ.pop(k, None)
            else:
                os.environ[k] = v


FIRST = {"CCGP_ME_BALANCED": "0", "CCGP_ME_SYM": "0"}
VARIANTS = [{"CCGP_ME_BALANCED": "1", "CCGP_ME_SYM": "0"}, {"CCGP_ME_BALANCED": "0", "CCGP_ME_SYM": "1"},
            {"CCGP_ME_BALANCED": "1", "CCGP_ME_SYM": "1"}]


@pytest.mark.parametrize("n_old,n_new,d", [(14, 7, 2), (14, 8, 2), (14, 1, 2), (14, 2, 2), (14, 3, 1), (0, 7, 2), (5, 4, 3),
                                           (16, 6, 2), (20, 5, 4), (24, 8, 2), (30, 7, 2)])
def test_me_variants_bit_identical_and_match_oracle(engine, n_old, n_new, d):
    rng = np.random.default_rng(100 * n_old + 10 * n_new + d)
    D_old = rng.uniform(-1, 1, (n_old, d)) if n_old else None
    C, P = 203, 9                                           # ragged: not a multiple of the 16 designs of a pass
    D_new = rng.uniform(-1, 1, (C, n_new, d))
    params = np.column_stack([rng.uniform(0.1, 0.9, P), rng.uniform(0.2, 3.0, P), rng.uniform(1.0, 12.0, P)])
    ref = _with_env(FIRST, lambda: engine.me_schur_batch(D_old, D_new, params))
    for env in VARIANTS:
        got = _with_env(env, lambda: engine.me_schur_batch(D_old, D_new, params))
        for a, b in zip(ref, got):
            assert np.array_equal(a.view(np.int64) if a.dtype == np.float64 else a, b.view(np.int64) if b.dtype == np.float64 else b), env
    want = orc.me_schur_negdet_batch(D_old, D_new[:40], params) if 0 < n_old <= 16 and d >= 2 else None   # d = 1: 17 collinear points, kappa ~ 1e15
    if want is not None:
        ok = ref[2][:40] == 0
        assert ok.mean() > 0.5
        err = np.abs(ref[0][:40][ok] - want[ok]) / np.abs(want[ok])
        assert err.max() < 1e-5, err.max()                  # sanity only, kappa-limited (random designs: 3e-7 seen); the golden-case gates live in test_gpu_parity.py


def test_me_variants_paired_and_stencil_bit_identical(engine, designs):
    rng = np.random.default_rng(12)
    D_old = designs["me_initial14"]
    P, group, n_new, d = 5, 37, 7, 2
    params = np.column_stack([rng.uniform(0.2, 0.8, P), rng.uniform(0.3, 2.0, P), rng.uniform(2.0, 8.0, P)])
    D_new = rng.uniform(-1, 1, (P * group, n_new, d))
    X = rng.uniform(-1, 1, (P * group, n_new * d))
    ref_p = _with_env(FIRST, lambda: engine.me_schur_paired(D_old, D_new, params, group))
    ref_s = _with_env(FIRST, lambda: engine.me_schur_stencil(D_old, X, n_new, d, params, group))
    for env in VARIANTS:
        got_p = _with_env(env, lambda: engine.me_schur_paired(D_old, D_new, params, group))
        got_s = _with_env(env, lambda: engine.me_schur_stencil(D_old, X, n_new, d, params, group))
        assert np.array_equal(ref_p[0].view(np.int64), got_p[0].view(np.int64)) and np.array_equal(ref_p[1], got_p[1]), env
        assert np.array_equal(ref_s[0].view(np.int64), got_s[0].view(np.int64)) and np.array_equal(ref_s[1], got_s[1]), env
