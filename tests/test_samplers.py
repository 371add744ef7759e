"""Lock-step Metro / Laplace drivers: host logic on CPU with the oracle as the
likelihood, and on the GPU with the CUDA `logpost_batch`."""
import math

import numpy as np
import pytest

from oracle import ccgp_oracle as orc
from ccgp_b200 import samplers


class RowRng:
    """Replays pre-drawn rows: iteration t hands chain c the values u[t, c], z[t, c, :]."""
    def __init__(self, u, z, chains=None):
        self.u, self.z, self.t = u, z, 0
        self.chains = list(range(u.shape[1])) if chains is None else chains

    def random(self, C):
        assert C == len(self.chains)
        return self.u[self.t, self.chains]

    def standard_normal(self, shape):
        out = self.z[self.t, self.chains]
        self.t += 1
        return out


def oracle_logpost_fn(D, y, sigma2, family, script):
    def fn(theta):
        rows = [orc.logpost(D, th, y, sigma2, family, script) for th in np.atleast_2d(theta)]
        return dict(val=np.array([r["val"] for r in rows]), beta=np.array([r["beta"] for r in rows]))
    return fn


def test_levinson_matches_direct_yule_walker():
    rng = np.random.default_rng(3)
    x = np.zeros(400)
    for t in range(2, 400):
        x[t] = 0.6 * x[t - 1] - 0.25 * x[t - 2] + rng.normal()
    ar, var_pred = samplers._ar_yule_walker_aic(x)
    p = ar.size
    assert 1 <= p <= 26
    n = x.size
    xc = x - x.mean()
    r = np.array([np.dot(xc[: n - k], xc[k:]) / n for k in range(p + 1)])
    T = np.array([[r[abs(i - j)] for j in range(p)] for i in range(p)])
    direct = np.linalg.solve(T, r[1:])
    np.testing.assert_allclose(ar, direct, rtol=1e-9, atol=1e-12)
    v = r[0] - direct @ r[1:]
    np.testing.assert_allclose(var_pred, v * n / (n - (p + 1)), rtol=1e-9)


def test_spectrum0_and_geweke_behaviour():
    rng = np.random.default_rng(11)
    phi, n = 0.5, 20000
    e = rng.normal(size=n)
    x = np.zeros(n)
    for t in range(1, n):
        x[t] = phi * x[t - 1] + e[t]
    s0 = samplers.spectrum0_ar(x)
    assert abs(s0 - 1.0 / (1 - phi) ** 2) / 4.0 < 0.1          # sigma^2 / (1 - phi)^2 = 4
    zs = [samplers.geweke_z(x[i * 1000:(i + 1) * 1000]) for i in range(20)]
    assert np.std(zs) < 2.0 and abs(np.mean(zs)) < 1.0         # ~ N(0, 1) on stationary pieces
    trend = x[:1000] + np.linspace(0, 10, 1000)
    assert samplers.geweke_pvalue(trend) < 1e-3
    assert samplers.geweke_pvalue(np.ones(50)) == 0.0           # try-error -> 0 ([A]:531)
    # windows: first ceil(1 + 0.1 (n-1)) values, last from floor(n - 0.5 (n-1)) (1-based)
    z = samplers.geweke_z(np.arange(101.0) + rng.normal(size=101))
    assert z < 0


def test_laplace_batch_on_a_quadratic():
    A = np.array([[3.0, 0.4, 0.1], [0.4, 2.0, -0.3], [0.1, -0.3, 1.5]])
    m = np.array([0.3, -1.2, 2.0])

    def fn(theta):
        dlt = np.atleast_2d(theta) - m
        return dict(val=-0.5 * np.einsum("bi,ij,bj->b", dlt, A, dlt) + 1.25, beta=np.zeros(len(dlt)))

    starts = np.array([[0.0, 0.0, 0.0], [1.0, -2.0, 3.0], [-1.0, 1.0, 1.0]])
    out = samplers.laplace_batch(fn, starts)
    assert out["converge"].all()
    np.testing.assert_allclose(out["mode"], np.repeat(m[None], 3, 0), atol=2e-4)
    np.testing.assert_allclose(out["var"], np.repeat(np.linalg.inv(A)[None], 3, 0), rtol=1e-5, atol=1e-7)
    want = 1.5 * math.log(2 * math.pi) - 0.5 * np.linalg.slogdet(A)[1] + 1.25
    np.testing.assert_allclose(out["int"], want, atol=1e-6)
    # lock-step == one simplex at a time
    solo = samplers.laplace_batch(fn, starts[1:2])
    np.testing.assert_array_equal(solo["mode"][0], out["mode"][1])
    np.testing.assert_array_equal(solo["var"][0], out["var"][1])


def test_metro_multichain_equals_sequential_chains(designs):
    D = designs["maximin14"]
    y = orc.test_function(4, D[:, 0], D[:, 1])
    fn = oracle_logpost_fn(D, y, 1.0, orc.FAMILY_ISO, "I")
    rng = np.random.default_rng(5)
    C, k, T = 3, 3, 400
    u, z = rng.random((T, C)), rng.standard_normal((T, C, k))
    mu = np.array([[1.0, 2.0, 0.0], [1.5, 2.5, 0.3], [0.5, 1.5, -0.4]])
    v = 0.05 * np.eye(3)
    kw = dict(N=60, samp_size=20, batch_size=10, alpha=0.3, logpost_fn=fn)
    multi = samplers.Metro_multichain(mu, v, rng=RowRng(u, z), **kw)
    for c in range(C):
        solo = samplers.Metro_multichain(mu[c:c + 1], v, rng=RowRng(u, z, [c]), **kw)[0]
        assert solo["n_accept"] == multi[c]["n_accept"] and solo["n_proposals"] == multi[c]["n_proposals"]
        np.testing.assert_array_equal(solo["sample"], multi[c]["sample"])
        np.testing.assert_array_equal(solo["beta"], multi[c]["beta"])
    for ch in multi:
        assert ch["sample"].shape == (20, 3) and ch["n_accept"] >= 20
        assert ch["n_accept"] % 10 == 0 or ch["n_accept"] == 60
        assert ch["pv"] >= 0.3 or ch["n_accept"] == 60
    # a by-hand re-run of chain 0 with the reference's loop ([A]:509-534)
    L = np.linalg.cholesky(math.sqrt(2.0) * v)
    th, lold, acc = mu[0].copy(), fn(mu[0:1])["val"][0], []
    for t in range(multi[0]["n_proposals"]):
        cand = th + L @ z[t, 0]
        lc = fn(cand[None])["val"][0]
        if lc - lold > math.log(u[t, 0]):
            th, lold = cand, lc
            acc.append(cand)
    np.testing.assert_allclose(np.array(acc)[-20:], multi[0]["sample"], rtol=0, atol=0)


@pytest.mark.gpu
def test_gpu_laplace_and_metro_match_the_oracle_driven_run(engine, designs):
    from ccgp_b200 import reference_api as api
    D = designs["maximin14"]
    y = orc.test_function(4, D[:, 0], D[:, 1])
    gpu_fn = lambda th: api.logpost_batch(D, th, y, 1.0, script="I", engine=engine)   # noqa: E731
    cpu_fn = oracle_logpost_fn(D, y, 1.0, orc.FAMILY_ISO, "I")
    starts = np.array([[1.0, 2.0, 0.0], [2.0, 3.0, 0.5]])
    lg, lc = samplers.laplace_batch(gpu_fn, starts), samplers.laplace_batch(cpu_fn, starts)
    np.testing.assert_allclose(lg["mode"], lc["mode"], atol=1e-6)
    np.testing.assert_allclose(lg["var"], lc["var"], rtol=1e-4, atol=1e-8)
    rng = np.random.default_rng(9)
    C, T = 8, 600
    u, z = rng.random((T, C)), rng.standard_normal((T, C, 3))
    mu = np.repeat(lc["mode"][:1], C, 0)
    kw = dict(N=80, samp_size=30, batch_size=10, alpha=0.2)
    g = samplers.Metro_multichain(mu, lc["var"][0], logpost_fn=gpu_fn, rng=RowRng(u, z), **kw)
    c = samplers.Metro_multichain(mu, lc["var"][0], logpost_fn=cpu_fn, rng=RowRng(u, z), **kw)
    for a, b in zip(g, c):
        assert a["n_accept"] == b["n_accept"] and a["n_proposals"] == b["n_proposals"]
        np.testing.assert_array_equal(a["sample"], b["sample"])           # same accept decisions
        np.testing.assert_allclose(a["beta"], b["beta"], rtol=1e-10, atol=1e-12)
