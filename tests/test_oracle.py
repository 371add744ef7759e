"""CPU: the oracle against its committed golden vectors and against itself
(reference-faithful path vs minimal path vs 50-digit truth)."""
import numpy as np
import pytest

from oracle import ccgp_oracle as orc
from conftest import rel_err


def test_designs_fixture_shapes(designs):
    assert designs["maximin14"].shape == (14, 2)
    assert designs["maximin100"].shape == (100, 2)
    assert designs["me_all_subdesigns"].shape == (1000, 7, 2)
    assert designs["hyperpars_2d"].shape == (60, 4)
    assert designs["he_train"].shape == (64, 5) and designs["he_hyperpars"].shape == (624, 4)
    assert designs["gv50_train1"].shape == (50, 10) and designs["gv90_train1"].shape == (90, 10)
    assert designs["design1d"].shape == (201, 8)
    # KATs that follow from the shipped data alone
    assert np.allclose(designs["me_plugin21"][:14], designs["me_initial14"], atol=1e-7)
    pool = designs["me_all_subdesigns"].reshape(7000, 2)
    rows = np.array([5374, 6776, 813, 5495, 5487, 6274, 1852]) - 1
    assert np.abs(designs["me_kmedoids21"][14:] - pool[rows]).max() < 1e-7


def test_loglik_reference_matches_golden(golden, designs):
    X = designs["maximin100"]
    y = golden["c1n100_y"]
    for b in range(6):
        r = orc.loglik_reference(X, y, 1.0, orc.FAMILY_ANISO_LAMBDA, golden["c1n100_nat"][b])
        assert rel_err(r["loglik"], golden["c1n100_ref"][b]) < 1e-13
        assert rel_err(r["beta"], golden["c1n100_beta"][b]) < 1e-12


def test_logpost_adds_jacobian_and_prior(golden, designs):
    X = designs["maximin100"]
    y = golden["c1n100_y"]
    th = golden["c1n100_theta"][0]
    r = orc.logpost(X, th, y, 1.0, orc.FAMILY_ANISO_LAMBDA, "A")
    assert rel_err(r["val"], golden["c1n100_logpost_val"][0]) < 1e-13
    psi1, psi2, phi, zeta = th
    jac = -phi - 2 * np.log(1 + np.exp(-phi)) + psi1 + psi2 + zeta
    pri = -psi1 - psi1 ** 2 / 2 - psi2 - psi2 ** 2 / 2 - 4 * zeta - 4 / np.exp(zeta)
    assert abs(r["val"] - (r["loglik"] + jac + pri)) < 1e-12


@pytest.mark.parametrize("case", ["c1n100", "c1n14gls", "c2gls", "gv50", "gv90"])
def test_reference_vs_minimal_gls(golden, case):
    """LU-inverse + chol2inv (what R runs) vs Cholesky + forward solves agree to
    1e-10 relative wherever kappa_1(R) <= 1e6 (the stated parity range)."""
    ok = golden[case + "_kappa"] <= 1e6
    assert ok.sum() >= 8
    assert rel_err(golden[case + "_minimal"][ok], golden[case + "_ref"][ok]).max() < 1e-10


@pytest.mark.parametrize("case", ["c1n100", "c1n14", "c1n14gls", "c2gls", "c2tau"])
def test_minimal_vs_truth(golden, case):
    tr = golden[case + "_truth"]
    ok = ~np.isnan(tr)
    assert ok.sum() >= 2
    assert rel_err(golden[case + "_minimal"][ok], tr[ok]).max() < 1e-10


def test_tau_variant_reference_is_the_inaccurate_one(golden):
    """SURVEY 'hard parts': factorising cR + tau^2 11' directly (the reference) loses
    digits; the Sherman-Morrison form is ~1e4x closer to the 50-digit truth."""
    tr = golden["c1n14_truth"]
    ok = ~np.isnan(tr)
    e_ref = rel_err(golden["c1n14_ref"][ok], tr[ok]).max()
    e_min = rel_err(golden["c1n14_minimal"][ok], tr[ok]).max()
    assert e_min < 1e-11 and e_min < e_ref


def test_halton_and_sweep(golden):
    assert np.array_equal(orc.halton_base2(16), golden["halton_first16"])
    assert np.allclose(orc.halton_base2(4), [0.5, 0.25, 0.75, 0.125])
    c = orc.sweep_candidates([3, 2], [5, 16], 8)
    assert c.shape == (8, 3) and np.all(c[:, 1:] > 0)
    # qigamma(.5, 3, 2) is the prior median quoted at [M]:982 as ~1
    assert abs(orc.qigamma(0.5, 3, 2) - 0.748) < 0.01


def test_likeli_hyperpars_golden(golden, designs):
    X = designs["maximin14"]
    hp = designs["hyperpars_2d"]
    i = int(golden["c1n14_likeli_rows"][0])
    v = orc.likeli_hyperpars(X, golden["c1n14_y"], hp[i, 0:2], hp[i, 2:4], 0.7, N=int(golden["c1n14_likeli_N"]), tau=100.0)
    assert rel_err(np.log(v), np.log(golden["c1n14_likeli"][0])) < 1e-9


def test_predict_oracle_golden(golden, designs):
    X = designs["maximin14"]
    m, v = orc.predict_table(X, golden["pred14_y"], 0.9, orc.FAMILY_ANISO_LAMBDA, golden["pred14_pars"][:2], golden["pred14_Xnew"][:5])
    assert rel_err(m, golden["pred14_mean"][:5, :2]).max() < 1e-12
    assert rel_err(v, golden["pred14_var"][:5, :2]).max() < 1e-10


def test_me_oracle_golden_and_schur_identity(golden, designs):
    D_old = designs["me_initial14"]
    pool = designs["me_all_subdesigns"]
    p, t1, t2 = golden["me_params"][0]
    Rinv = orc.mixed_R_old_inv(D_old, p, t1, t2)
    for c in (0, 7, 199):
        v = orc.Augmented_Mixed_Entropy(D_old, pool[c], p, t1, t2, Rinv)
        assert rel_err(v, golden["me_negdet_200"][c, 0]) < 1e-12
        # det(R_all) = det(R_old) det(Schur)
        det_all = -orc.Entropy(np.vstack([D_old, pool[c]]), p, t1, t2)
        det_old = -orc.Entropy(D_old, p, t1, t2)
        assert rel_err(det_all, det_old * (-v)) < 1e-9
    # the selection is unambiguous in FP64: top-2 gap of the full pool under the prior medians
    nd = np.sort(golden["me_negdet_full_prior"])
    assert (nd[1] - nd[0]) / abs(nd[0]) > 1e-8
    assert int(golden["me_argmin_full_prior"]) == int(np.argmin(golden["me_negdet_full_prior"]))


def test_r_solve_rejects_singular():
    A = np.ones((4, 4))
    assert orc.r_solve(A) is None
    R = orc.Mixed_corr_matrix(np.array([[0.0, 0.0], [0.0, 0.0], [1.0, 1.0]]), orc.FAMILY_ISO, [0.5, 1.0, 2.0])
    assert orc.loglik_reference(np.array([[0.0, 0.0], [0.0, 0.0], [1.0, 1.0]]), [1.0, 2.0, 3.0], 1.0, orc.FAMILY_ISO, [0.5, 1.0, 2.0])["status"] == 2
    assert R.shape == (3, 3)


def test_r_det_matches_numpy():
    rng = np.random.default_rng(0)
    A = rng.normal(size=(7, 7))
    assert rel_err(orc.r_det(A), np.linalg.det(A)) < 1e-12


def test_expanded_square_form_is_what_r_computes():
    """corr.matrix's U + t(U) + V equals the direct form to ~1e-14 but is not exactly symmetric/unit-diagonal."""
    rng = np.random.default_rng(3)
    X = rng.uniform(-1, 1, (30, 2))
    R = orc.corr_matrix(X, [3.0, 5.0])
    Rd = orc.Mixed_corr_matrix_direct(X, orc.FAMILY_ANISO_LAMBDA, [1.0, 3.0, 5.0, 0.0])
    assert np.abs(R - Rd).max() < 1e-13


def test_cgp_objective_restatement_against_a_cholesky_formulation(designs):
    """CGP comparator ([A]:104-135): the literal restatement (LU inverse, det) against the same objective written with
    a Cholesky factor -- two independent codings of one formula."""
    from ccgp_b200 import workloads
    X = designs["maximin14"]
    y = workloads.test_function_4(X)
    Xs, _ = orc.cgp_standardise(X)
    lower, upper = orc.cgp_bounds(Xs)
    assert lower[0] == 0.001 and upper[0] == 1.0 and lower[-1] == 0.0 and upper[-1] == 1.0
    assert np.isclose(upper[1], lower[3]) and np.isclose(upper[3] / lower[3], 3.0)        # alpha_l, kappa_u = 3 alpha_l
    rng = np.random.default_rng(3)
    for w in lower + (upper - lower) * rng.random((12, 5)):
        lam, th, kappa, bw = w[0], w[1:3], w[3], w[4]
        G, L, Gbw = orc.cgp_psi(Xs, th), orc.cgp_psi(Xs, kappa + th), orc.cgp_psi(Xs, th * bw)
        n, one, sig = 14, np.ones(14), np.ones(14)
        for rep in range(5):
            Q = G + lam * np.sqrt(np.outer(sig, sig)) * L
            C = np.linalg.cholesky(Q)
            uy, u1 = np.linalg.solve(C.T, np.linalg.solve(C, y)), np.linalg.solve(C.T, np.linalg.solve(C, one))
            beta = uy.sum() / u1.sum()
            temp = uy - beta * u1
            if rep == 4:
                break
            e = y - beta - G @ temp
            sig = (Gbw @ e ** 2) / (Gbw @ one)
            sig = sig / sig.mean()
        val = 2.0 * np.log(np.diag(C)).sum() + n * np.log((y - beta) @ temp / n)
        assert abs(orc.cgp_var_mle_dk(Xs, y, w) - val) < 1e-7 * max(1.0, abs(val))
