"""CPU: the C-ABI library loads and exports every symbol of include/ccgp.h, the
host logic (sharding, candidate grids, priors) is right, and there is no silent
CPU fallback."""
import os
import re
import subprocess
import sys

import numpy as np
import pytest

import ccgp_b200
from ccgp_b200 import _capi, reference_api as api, sharding, workloads
from oracle import ccgp_oracle as orc

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _have_gpu():
    try:
        import torch
        return torch.cuda.is_available()
    except Exception:
        return False


def test_library_exports_every_declared_symbol():
    hdr = open(os.path.join(ROOT, "include", "ccgp.h")).read()
    declared = sorted(set(re.findall(r"\b(ccgp_[a-z0-9_]+)\s*\(", hdr)))
    assert declared == sorted(_capi.SYMBOLS)
    assert os.path.exists(_capi.LIB_PATH), "libccgp.so not built: run __graft_entry__.build()"
    lib = _capi.load()
    for s in declared:
        assert hasattr(lib, s), s
    assert lib.ccgp_num_params(0, 2) == 3 and lib.ccgp_num_params(1, 2) == 4 and lib.ccgp_num_params(1, 9) == 11
    assert lib.ccgp_num_params(7, 2) == -1


def test_library_contains_sm100a_code_only():
    out = subprocess.run(["cuobjdump", "--list-elf", _capi.LIB_PATH], capture_output=True, text=True)
    if out.returncode != 0:
        pytest.skip("cuobjdump unavailable")
    archs = set(re.findall(r"sm_\d+a?", out.stdout))
    assert archs == {"sm_100a"}, archs


@pytest.mark.skipif(_have_gpu(), reason="checks the no-GPU failure mode")
def test_no_cpu_fallback_engine_fails_loudly():
    with pytest.raises(ccgp_b200.CcgpError) as ei:
        ccgp_b200.Engine(0)
    assert "no CPU fallback" in str(ei.value) or "CUDA" in str(ei.value)


def test_product_code_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "convex-combination-of-gaussian-processes_b200")
    for dp, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                txt = open(os.path.join(dp, f)).read()
                for needle in ("import oracle", "from oracle", "ccgp_oracle", "oracle/"):
                    assert needle not in txt, (f, "product code must not reach into oracle/: " + needle)


def test_shard_range_partitions_exactly():
    for total in (0, 1, 7, 1000, 2 ** 20 + 3):
        for world in (1, 2, 3, 4, 8):
            seen = 0
            prev_hi = 0
            for r in range(world):
                lo, hi = sharding.shard_range(total, r, world)
                assert lo == prev_hi and hi >= lo
                prev_hi = hi
                seen += hi - lo
            assert seen == total and prev_hi == total


_WORKER = r'''
import os, sys
sys.path.insert(0, sys.argv[1])
import torch.distributed as dist
import ccgp_b200
from ccgp_b200 import sharding
rank = int(os.environ["RANK"]); world = int(os.environ["WORLD_SIZE"])
dist.init_process_group("gloo", rank=rank, world_size=world)
import numpy as np
vals = np.array([3.0, 1.5, 2.0, 1.5, 9.0, float("nan"), 1.5, 4.0])   # global which.min -> index 1
lo, hi = sharding.shard_range(len(vals), rank, world)
loc = vals[lo:hi]
if len(loc) and not np.all(np.isnan(loc)):
    i = int(np.nanargmin(loc)); v = float(loc[i]); gi = lo + i
else:
    v, gi = float("nan"), -1
gv, gidx = sharding.allreduce_argmin(v, gi)
assert gidx == 1 and gv == 1.5, (rank, gv, gidx)
# a rank whose shard is all-NaN must not win; all-NaN everywhere -> (-1)
gv2, gidx2 = sharding.allreduce_argmin(float("nan"), -1)
assert gidx2 == -1
dist.barrier(); dist.destroy_process_group()
print("ok", rank)
'''


@pytest.mark.parametrize("world", [2, 3])
def test_allreduce_argmin_gloo(tmp_path, world):
    script = tmp_path / "w.py"
    script.write_text(_WORKER)
    procs = []
    port = 29500 + os.getpid() % 2000 + world
    for r in range(world):
        env = dict(os.environ, RANK=str(r), WORLD_SIZE=str(world), MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
        procs.append(subprocess.Popen([sys.executable, str(script), ROOT], env=env, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True))
    for p in procs:
        out, _ = p.communicate(timeout=240)
        assert p.returncode == 0, out


def test_single_process_argmin_passthrough():
    assert sharding.allreduce_argmin(2.5, 7) == (2.5, 7)
    v, i = sharding.allreduce_argmin(float("nan"), -1)
    assert i == -1 and v != v


def test_host_candidate_grid_matches_oracle():
    assert np.array_equal(api.halton_base2(64), orc.halton_base2(64))
    a = api.sweep_candidates([3, 1], [5, 40], 100)
    b = orc.sweep_candidates([3, 1], [5, 40], 100)
    assert np.array_equal(a, b)


def test_host_prior_and_jacobian_match_oracle():
    rng = np.random.default_rng(2)
    th4 = rng.normal(size=(5, 4))
    th3 = rng.normal(size=(5, 3))
    for b in range(5):
        assert abs(api.log_jacobian(th4, ccgp_b200.GAUSS_ANISO_LAMBDA, 2)[b] - orc.log_jacobian(orc.FAMILY_ANISO_LAMBDA, th4[b], 2)) < 1e-13
        assert abs(api.log_jacobian(th3, ccgp_b200.GAUSS_ISO, 4)[b] - orc.log_jacobian(orc.FAMILY_ISO, th3[b], 4)) < 1e-13
        assert abs(api.log_prior(th4, "A")[b] - orc.log_prior("A", th4[b])) < 1e-13
        for s in ("I", "G"):
            assert abs(api.log_prior(th3, s)[b] - orc.log_prior(s, th3[b])) < 1e-13
        assert abs(api.log_prior(th3, "V", (3, 2, 5, 16))[b] - orc.log_prior("V", th3[b], (3, 2, 5, 16))) < 1e-13
        assert np.allclose(api.transform(th4, ccgp_b200.GAUSS_ANISO_LAMBDA, 2)[b], orc.transform_theta(orc.FAMILY_ANISO_LAMBDA, th4[b], 2), rtol=1e-15)


def test_workloads_are_seeded_and_shaped():
    X, y, s2 = workloads.m1_design()
    assert X.shape == (100, 2) and y.shape == (100,) and s2 == 1.0
    a, b = workloads.m1_candidates(1000), workloads.m1_candidates(1000)
    assert np.array_equal(a, b) and a.shape == (1000, 4)
    # prefix property: the first rows do not depend on B only for the same B; different B reshuffles -> documented
    D_old, pool = workloads.me_pool()
    assert D_old.shape == (14, 2) and pool.shape == (1000, 7, 2)
    q = workloads.me_params(5)
    assert q.shape == (5, 3) and list(q[0]) == [0.5, 1.0, 4.0]
    assert workloads.synthetic_pool(256).shape == (256, 2)


def test_dexp_host_accuracy(tmp_path):
    """ccgp_math.h's exp(-s) (the one the kernels use) vs libm on the host: <= 1 ulp-ish."""
    src = tmp_path / "t.c"
    src.write_text(r'''
#include <stdio.h>
#include <stdlib.h>
#include "%s/convex-combination-of-gaussian-processes_b200/csrc/ccgp_math.h"
int main(){ double worst=0; srand(1);
 for(long i=0;i<3000000;i++){ double s=(i%%3==0)?(rand()/(double)RAND_MAX)*680.0:(i%%3==1?(rand()/(double)RAND_MAX)*5.0:(rand()/(double)RAND_MAX)*0.01);
  double a=dexp_neg(s), b=exp(-s); double e=fabs(a-b)/b; if(e>worst) worst=e; }
 printf("%%.3e\n", worst); return (dexp_neg(0.0)==1.0 && dexp_neg(800.0)==0.0) ? 0 : 1; }
''' % ROOT)
    exe = tmp_path / "t"
    subprocess.check_call(["gcc", "-O2", "-mfma", "-o", str(exe), str(src), "-lm"])
    out = subprocess.run([str(exe)], capture_output=True, text=True)
    assert out.returncode == 0
    assert float(out.stdout) < 3.4e-16


def test_bessel_matern_spline_host_accuracy(tmp_path):
    """ccgp_math.h's K0/K1-based Matern and the cubic spline (the code the kernels run) vs SciPy's kv."""
    from scipy.special import kv, gamma
    src = tmp_path / "b.c"
    src.write_text(r'''
#include <stdio.h>
#include "%s/convex-combination-of-gaussian-processes_b200/csrc/ccgp_math.h"
int main(){ for (int i = 0; i < 400; ++i) { double t = 1e-4 * pow(1.04, i); double k0, k1; ccgp_bessel_k01(t, &k0, &k1);
  printf("%%.17g %%.17g %%.17g %%.17g %%.17g %%.17g\n", t, k0, k1, ccgp_matern(t, 10, 1.0/384.0),
         ccgp_matern(t, 5, 1.0/(1.3293403881791370*2.8284271247461903)), ccgp_spline(t)); } return 0; }
''' % ROOT)
    exe = tmp_path / "b"
    subprocess.check_call(["gcc", "-O2", "-mfma", "-o", str(exe), str(src), "-lm"])
    rows = np.array([[float(v) for v in line.split()] for line in subprocess.run([str(exe)], capture_output=True, text=True).stdout.splitlines()])
    t = rows[:, 0]
    ok = t < 600
    assert np.max(np.abs(rows[ok, 1] / kv(0, t[ok]) - 1)) < 5e-15
    assert np.max(np.abs(rows[ok, 2] / kv(1, t[ok]) - 1)) < 5e-15
    for col, nu in ((3, 5.0), (4, 2.5)):
        want = t ** nu * kv(nu, t) / (gamma(nu) * 2 ** (nu - 1))
        assert np.max(np.abs(rows[ok, col] - want[ok]) / np.maximum(want[ok], 1e-300)) < 2e-14
    assert np.allclose(rows[:, 5], orc.spline_corr_func(1.0, t), rtol=0, atol=5e-16)


def test_oracle_matern_half_integer_closed_forms():
    h = np.linspace(0.0, 2.0, 41)
    for theta in (0.3, 1.7):
        t = 2 * np.sqrt(0.5) * h / theta
        assert np.allclose(orc.Matern_corr_func(0.5, h, theta), np.exp(-t), rtol=1e-13)
        t = 2 * np.sqrt(1.5) * h / theta
        assert np.allclose(orc.Matern_corr_func(1.5, h, theta), (1 + t) * np.exp(-t), rtol=1e-13)
    R = orc.Mixed_corr_matrix(np.linspace(0, 1, 8).reshape(-1, 1), orc.FAMILY_MATERN_SPLINE1D, [0.4, 0.3, 0.5])
    assert np.allclose(R, R.T) and np.allclose(np.diag(R), 1.0)
