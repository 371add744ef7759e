"""k-medoids (PAM) step of the batch-sequential ME design: the shipped files give an
exact known answer -- 7-medoids of the 7000 points of All_Subdesigns.txt are rows 15-21 of
`k-medoids ME Design.txt` (rows 5374, 6776, 813, 5495, 5487, 6274, 1852 of the point cloud, 1-based)."""
import numpy as np
import pytest

from oracle import ccgp_oracle as orc

KAT_ROWS = np.array([5374, 6776, 813, 5495, 5487, 6274, 1852]) - 1


def test_shipped_medoids_are_rows_of_the_point_cloud(designs):
    P = designs["me_all_subdesigns"].reshape(-1, 2)
    km = designs["me_kmedoids21"]
    np.testing.assert_array_equal(km[:14], designs["me_initial14"])
    np.testing.assert_array_equal(P[KAT_ROWS], km[14:])


def test_oracle_pam_reproduces_the_shipped_design(designs):
    P = designs["me_all_subdesigns"].reshape(-1, 2)
    trace = []
    med, cost, swaps = orc.pam_kmedoids(P, 7, trace=trace)
    assert sorted(med.tolist()) == sorted(KAT_ROWS.tolist())
    assert swaps == 4 and trace[0][0] == "build"
    assert abs(cost - 948.4281295516507) < 1e-9


def test_oracle_pam_small_cases():
    rng = np.random.default_rng(2)
    centers = np.array([[0.0, 0.0], [5.0, 5.0], [0.0, 6.0]])
    P = np.vstack([c + 0.1 * rng.normal(size=(30, 2)) for c in centers])
    med, cost, _ = orc.pam_kmedoids(P, 3)
    assert sorted(m // 30 for m in med) == [0, 1, 2]                      # one medoid per blob
    brute = min(np.linalg.norm(P[:, None] - P[[a, b, c]][None], axis=2).min(axis=1).sum()
                for a in range(0, 30, 3) for b in range(30, 60, 3) for c in range(60, 90, 3))
    assert cost <= brute + 1e-12
    med1, cost1, _ = orc.pam_kmedoids(P[:5], 5)                           # k = n: every point its own medoid
    assert sorted(med1.tolist()) == [0, 1, 2, 3, 4] and cost1 == 0.0
    med2, _, _ = orc.pam_kmedoids(P[:7], 1)
    assert med2[0] == int(np.argmin(np.linalg.norm(P[:7, None] - P[None, :7], axis=2).sum(axis=0)))


@pytest.mark.gpu
def test_gpu_pam_exact_kat_and_oracle_trace(engine, designs):
    from ccgp_b200 import reference_api as api
    P = designs["me_all_subdesigns"].reshape(-1, 2)
    med, cost, swaps = engine.kmedoids_pam(P, 7)
    assert sorted(med.tolist()) == sorted(KAT_ROWS.tolist())              # bit-exact selection
    omed, ocost, oswaps = orc.pam_kmedoids(P, 7)
    assert med.tolist() == omed.tolist() and swaps == oswaps              # same BUILD order, same exchanges
    assert abs(cost - ocost) < 1e-9
    out = api.kmedoids_design(designs["me_initial14"], designs["me_all_subdesigns"].reshape(1000, 7, 2), 7, engine=engine)
    assert sorted(map(tuple, out["Design"][14:])) == sorted(map(tuple, designs["me_kmedoids21"][14:]))
    np.testing.assert_array_equal(out["Design"][:14], designs["me_initial14"])


@pytest.mark.gpu
@pytest.mark.parametrize("n,d,k", [(1, 1, 1), (40, 3, 5), (257, 2, 16), (1000, 4, 7)])
def test_gpu_pam_matches_oracle_on_random_clouds(engine, n, d, k):
    k = min(k, n)
    P = np.random.default_rng(100 + n).uniform(-1, 1, (n, d))
    med, cost, swaps = engine.kmedoids_pam(P, k)
    omed, ocost, oswaps = orc.pam_kmedoids(P, k)
    assert med.tolist() == omed.tolist() and swaps == oswaps
    assert abs(cost - ocost) < 1e-9 * max(1.0, ocost)
