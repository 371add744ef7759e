"""Host-side handle over the C ABI (include/ccgp.h): one `Engine` per GPU.

Host-pointer methods take/return NumPy arrays (column-major, as R hands them
over); `*_dev` methods take torch CUDA tensors and only enqueue work on the
engine's stream.  Nothing here computes: every number comes from libccgp.so.
"""
from __future__ import annotations

import ctypes as C
import os
import numpy as np

from . import _capi
from ._capi import CcgpError

GAUSS_ISO, GAUSS_ANISO_LAMBDA, GAUSS_ISO_RAW2, MATERN1D, MATERN_SPLINE1D = 0, 1, 2, 3, 4
NATURAL, LOGSCALE = 0, 1
MEAN_GLS_BETA, MEAN_ZERO_PLUS_TAU2 = 0, 1


def _f(a):
    """float64 column-major view/copy (what the ABI expects)."""
    return np.asfortranarray(np.asarray(a, dtype=np.float64))


def _ptr(a):
    return None if a is None else a.ctypes.data


def _prefer_bundled_nccl():
    """libccgp.so dlopen()s libnccl.so.2 when a multi-GPU context is created.  In a Python process that later imports
    torch, the NCCL loaded first wins (same SONAME) and torch's libtorch_cuda.so needs the newer one it ships with
    (seen on the GPU box: system 2.27.3 loaded first -> `undefined symbol: ncclDevCommCreate` at `import torch`).
    Point CCGP_NCCL_LIB at the pip-bundled library when there is one; an R process has no such conflict."""
    if os.environ.get("CCGP_NCCL_LIB"):
        return
    try:
        import importlib.util
        spec = importlib.util.find_spec("nvidia.nccl")
        for root in (spec.submodule_search_locations if spec else []):
            cand = os.path.join(root, "lib", "libnccl.so.2")
            if os.path.exists(cand):
                os.environ["CCGP_NCCL_LIB"] = cand
                return
    except Exception:  # noqa: BLE001
        pass


class Engine:
    def __init__(self, device: int = 0, n_gpus: int = None):
        """device: one GPU (ccgp_create).  n_gpus: all (<= 0) or the first n GPUs of the box behind one context
        (ccgp_create_multi): batches are sliced over the GPUs inside the library, which.min goes through NCCL."""
        self._lib = _capi.load()
        h = C.c_void_p()
        if n_gpus is None:
            rc = self._lib.ccgp_create(C.byref(h), int(device))
            what = "ccgp_create(device=%d)" % device
        else:
            _prefer_bundled_nccl()
            rc = self._lib.ccgp_create_multi(C.byref(h), int(n_gpus))
            what = "ccgp_create_multi(n_gpus=%d)" % n_gpus
        if rc != 0:
            raise CcgpError("%s failed (%d): %s" % (what, rc, self._lib.ccgp_last_error(None).decode()))
        self._h = h
        self.device = device
        self.n_gpus = self._lib.ccgp_num_gpus(h)
        self.n = 0
        self.d = 0

    # -- plumbing -----------------------------------------------------------------
    def _ck(self, rc):
        if rc != 0:
            raise CcgpError("libccgp error %d: %s" % (rc, self._lib.ccgp_last_error(self._h).decode()))

    def close(self):
        if getattr(self, "_h", None):
            self._lib.ccgp_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    @property
    def collective_count(self) -> int:
        return int(self._lib.ccgp_collective_count(self._h))

    def sync(self):
        self._ck(self._lib.ccgp_sync(self._h))

    def set_stream(self, cuda_stream_ptr):
        """Enqueue on the caller's stream (e.g. torch.cuda.current_stream().cuda_stream, 0 = legacy
        default stream); None = back to the engine's own stream."""
        if cuda_stream_ptr is None:
            self._ck(self._lib.ccgp_use_own_stream(self._h))
        else:
            self._ck(self._lib.ccgp_set_stream(self._h, C.c_void_p(int(cuda_stream_ptr))))

    @property
    def launch_count(self) -> int:
        return int(self._lib.ccgp_launch_count(self._h))

    def last_nll_config(self):
        t, s, c, v = C.c_int(), C.c_int(), C.c_int(), C.c_int()
        self._lib.ccgp_last_nll_config(self._h, C.byref(t), C.byref(s), C.byref(c), C.byref(v))
        return dict(team=t.value, smem_bytes=s.value, ctas_per_sm=c.value, variant=v.value)

    def measure_fp64_peak(self) -> float:
        f = C.c_double()
        self._ck(self._lib.ccgp_measure_fp64_peak(self._h, C.byref(f)))
        return f.value

    def measure_fp64_peak_dmma(self) -> float:
        f = C.c_double()
        self._ck(self._lib.ccgp_measure_fp64_peak_dmma(self._h, C.byref(f)))
        return f.value

    def set_matern_nu(self, nu: float):
        """Smoothness of the 1-D Matern families (integer or half-integer)."""
        self._ck(self._lib.ccgp_set_matern_nu(self._h, float(nu)))

    def num_params(self, family):
        return int(self._lib.ccgp_num_params(family, self.d))

    # -- design -------------------------------------------------------------------
    def set_design(self, X, y):
        X = _f(np.atleast_2d(X))
        y = np.ascontiguousarray(np.asarray(y, dtype=np.float64).reshape(-1))
        if y.shape[0] != X.shape[0]:
            raise ValueError("y has %d entries, X has %d rows" % (y.shape[0], X.shape[0]))
        self._ck(self._lib.ccgp_set_design(self._h, _ptr(X), X.shape[0], X.shape[1], _ptr(y)))
        self.n, self.d = X.shape

    # -- likelihood ---------------------------------------------------------------
    def nll_batch(self, cand, family, sigma2, scale=NATURAL, mean_mode=MEAN_GLS_BETA, tau=0.0):
        """-> (nll[B], beta[B], status[B]) for candidate rows `cand` (B x k)."""
        cand = _f(np.atleast_2d(cand))
        B, k = cand.shape
        if k != self.num_params(family):
            raise ValueError("family %d with d=%d needs %d columns, got %d" % (family, self.d, self.num_params(family), k))
        nll = np.empty(B)
        beta = np.empty(B)
        status = np.empty(B, dtype=np.int32)
        self._ck(self._lib.ccgp_nll_batch(self._h, family, scale, _ptr(cand), B, B, float(sigma2), mean_mode,
                                          float(tau), _ptr(nll), _ptr(beta), _ptr(status)))
        return nll, beta, status

    def nll_batch_dev(self, cand_t, family, sigma2, scale=NATURAL, mean_mode=MEAN_GLS_BETA, tau=0.0,
                      out_nll=None, out_beta=None, out_status=None):
        """torch CUDA tensors; cand_t is k x B row-major (= B x k column-major). Async."""
        import torch
        k, B = cand_t.shape
        assert cand_t.is_cuda and cand_t.dtype == torch.float64 and cand_t.is_contiguous()
        if out_nll is None:
            out_nll = torch.empty(B, dtype=torch.float64, device=cand_t.device)
        if out_beta is None:
            out_beta = torch.empty(B, dtype=torch.float64, device=cand_t.device)
        if out_status is None:
            out_status = torch.empty(B, dtype=torch.int32, device=cand_t.device)
        self._ck(self._lib.ccgp_nll_batch_dev(self._h, family, scale, cand_t.data_ptr(), B, B, float(sigma2), mean_mode,
                                              float(tau), out_nll.data_ptr(), out_beta.data_ptr(), out_status.data_ptr()))
        return out_nll, out_beta, out_status

    def nll_argmin(self, cand, family, sigma2, scale=NATURAL, mean_mode=MEAN_GLS_BETA, tau=0.0):
        cand = _f(np.atleast_2d(cand))
        B, k = cand.shape
        bv, bi = C.c_double(), C.c_int64()
        self._ck(self._lib.ccgp_nll_argmin(self._h, family, scale, _ptr(cand), B, B, float(sigma2), mean_mode, float(tau),
                                           C.byref(bv), C.byref(bi)))
        return bv.value, bi.value

    def argmin_dev(self, vals_t):
        bv, bi = C.c_double(), C.c_int64()
        self._ck(self._lib.ccgp_argmin_dev(self._h, vals_t.data_ptr(), vals_t.numel(), C.byref(bv), C.byref(bi)))
        return bv.value, bi.value

    def rinv_batch(self, cand, family, scale=NATURAL):
        """-> (Rinv[B, n, n], beta[B], status[B]) : logpost's R.Inv and beta."""
        cand = _f(np.atleast_2d(cand))
        B, k = cand.shape
        n = self.n
        rinv = np.empty((B, n, n))
        beta = np.empty(B)
        status = np.empty(B, dtype=np.int32)
        self._ck(self._lib.ccgp_rinv_batch(self._h, family, scale, _ptr(cand), B, B, _ptr(rinv), _ptr(beta), _ptr(status)))
        # each n x n block is column-major; R^-1 is symmetric so the transpose view is the same matrix
        return rinv.transpose(0, 2, 1), beta, status

    def rcond_batch(self, cand, family, scale=NATURAL):
        """-> (rcond[B], beta[B], status[B]): 1 / (||R||_1 ||R^-1||_1), the number base R's solve() tests
        against .Machine$double.eps ([A]:448-449)."""
        cand = _f(np.atleast_2d(cand))
        B, k = cand.shape
        rc = np.empty(B)
        beta = np.empty(B)
        status = np.empty(B, dtype=np.int32)
        self._ck(self._lib.ccgp_rcond_batch(self._h, family, scale, _ptr(cand), B, B, _ptr(rc), _ptr(beta), _ptr(status)))
        return rc, beta, status

    # -- prediction ---------------------------------------------------------------
    def predict(self, pars, family, X_new, sigma2, pars_vec=None, vec_family=-1):
        """-> (mean[T,S], var[T,S], status[S])."""
        pars = _f(np.atleast_2d(pars))
        X_new = _f(np.atleast_2d(X_new))
        S, T = pars.shape[0], X_new.shape[0]
        if X_new.shape[1] != self.d:
            raise ValueError("X_new has %d columns, design has %d" % (X_new.shape[1], self.d))
        pv = None if pars_vec is None else _f(np.atleast_2d(pars_vec))
        mean = np.empty((T, S), order="F")
        var = np.empty((T, S), order="F")
        status = np.empty(S, dtype=np.int32)
        self._ck(self._lib.ccgp_predict(self._h, family, _ptr(pars), S, S, vec_family, _ptr(pv), S,
                                        _ptr(X_new), T, float(sigma2), _ptr(mean), _ptr(var), _ptr(status)))
        return mean, var, status

    def factors(self, pars, family, pars_vec=None, vec_family=-1):
        """Factor the S posterior rows once and keep the factors in HBM (ccgp_factors_create): the device-side
        `factors.frame` ([A]:572-592).  -> `Factors`, whose predict(X_new, sigma2) returns what `predict` returns."""
        pars = _f(np.atleast_2d(pars))
        S = pars.shape[0]
        pv = None if pars_vec is None else _f(np.atleast_2d(pars_vec))
        h = C.c_void_p()
        self._ck(self._lib.ccgp_factors_create(self._h, family, _ptr(pars), S, S, vec_family, _ptr(pv), S, C.byref(h)))
        return Factors(self, h, S)

    # -- ME criteria --------------------------------------------------------------
    @staticmethod
    def _pack_designs(D_new):
        """(C, n_new, d) -> C blocks, each the column-major n_new x d matrix."""
        D_new = np.asarray(D_new, dtype=np.float64)
        if D_new.ndim == 2:
            D_new = D_new[None]
        return np.ascontiguousarray(D_new.transpose(0, 2, 1)), D_new.shape

    def me_schur_batch(self, D_old, D_new, params):
        """-> (negdet[C,P], logdet[C,P], status[C,P]); D_old may be None / empty."""
        blocks, (Cn, n_new, d) = self._pack_designs(D_new)
        params = _f(np.atleast_2d(params))
        P = params.shape[0]
        if D_old is None or len(D_old) == 0:
            Do, n_old = None, 0
        else:
            Do = _f(np.atleast_2d(D_old))
            n_old = Do.shape[0]
        negdet = np.empty((Cn, P), order="F")
        logdet = np.empty((Cn, P), order="F")
        status = np.empty((Cn, P), dtype=np.int32, order="F")
        self._ck(self._lib.ccgp_me_schur_batch(self._h, _ptr(Do), n_old, d, _ptr(blocks), n_new, Cn, _ptr(params), P, P,
                                               _ptr(negdet), _ptr(logdet), _ptr(status)))
        return negdet, logdet, status

    def predict_dev(self, pars_t, family, Xnew_t, sigma2, out_mean, out_var, out_status=None):
        """torch CUDA tensors: pars_t k x S row-major (= S x k column-major), Xnew_t d x T row-major; out_* T*S. Async."""
        k, S = pars_t.shape
        d, T = Xnew_t.shape
        self._ck(self._lib.ccgp_predict_dev(self._h, family, pars_t.data_ptr(), S, S, -1, None, S, Xnew_t.data_ptr(), T,
                                            float(sigma2), out_mean.data_ptr(), out_var.data_ptr(),
                                            out_status.data_ptr() if out_status is not None else None))

    def me_schur_batch_dev(self, D_old_t, n_old, d, D_new_t, n_new, Cn, params_t, P, out_negdet, out_logdet=None, out_status=None):
        self._ck(self._lib.ccgp_me_schur_batch_dev(
            self._h, D_old_t.data_ptr() if D_old_t is not None else None, n_old, d, D_new_t.data_ptr(), n_new, Cn,
            params_t.data_ptr(), P, P, out_negdet.data_ptr(),
            out_logdet.data_ptr() if out_logdet is not None else None,
            out_status.data_ptr() if out_status is not None else None))

    def me_argmin(self, D_old, D_new, params):
        """which.min over the candidate designs per parameter row -> (best_val[P], best_idx[P])."""
        blocks, (Cn, n_new, d) = self._pack_designs(D_new)
        params = _f(np.atleast_2d(params))
        P = params.shape[0]
        if D_old is None or len(D_old) == 0:
            Do, n_old = None, 0
        else:
            Do = _f(np.atleast_2d(D_old))
            n_old = Do.shape[0]
        bv = np.empty(P)
        bi = np.empty(P, dtype=np.int64)
        self._ck(self._lib.ccgp_me_argmin(self._h, _ptr(Do), n_old, d, _ptr(blocks), n_new, Cn, _ptr(params), P, P,
                                          _ptr(bv), _ptr(bi)))
        return bv, bi

    def subset_logdet_batch(self, pool, idx, family, params):
        pool = _f(np.atleast_2d(pool))
        idx = np.asfortranarray(np.asarray(idx, dtype=np.int32))
        Cn, m = idx.shape
        params = np.ascontiguousarray(np.asarray(params, dtype=np.float64).reshape(-1))
        out = np.empty(Cn)
        status = np.empty(Cn, dtype=np.int32)
        self._ck(self._lib.ccgp_subset_logdet_batch(self._h, _ptr(pool), pool.shape[0], pool.shape[1], _ptr(idx), m, Cn, Cn,
                                                    family, _ptr(params), _ptr(out), _ptr(status)))
        return out, status

    def me_schur_paired(self, D_old, D_new, params, group):
        """Design c against parameter row c // group only: D_new (P*group, n_new, d) -> negdet[P*group], status."""
        blocks, (Cn, n_new, d) = self._pack_designs(D_new)
        params = _f(np.atleast_2d(params))
        P = params.shape[0]
        if Cn != P * group:
            raise ValueError("need P * group designs")
        if D_old is None or len(D_old) == 0:
            Do, n_old = None, 0
        else:
            Do = _f(np.atleast_2d(D_old))
            n_old = Do.shape[0]
        negdet = np.empty(Cn)
        status = np.empty(Cn, dtype=np.int32)
        self._ck(self._lib.ccgp_me_schur_paired(self._h, _ptr(Do), n_old, d, _ptr(blocks), n_new, int(group), _ptr(params), P, P,
                                                _ptr(negdet), _ptr(status)))
        return negdet, status

    def me_schur_stencil(self, D_old, X, n_new, d, params, group, h=1e-3, lower=-1.0, upper=1.0):
        """X: (P*group, n_new*d) base designs as c(D) vectors -> vals[P*group, 2m+1] at the central-difference
        stencil (column 0: base, 1+2i: coordinate i + h, 2+2i: coordinate i - h, clipped to the box), status."""
        X = np.ascontiguousarray(np.asarray(X, dtype=np.float64))
        params = _f(np.atleast_2d(params))
        P = params.shape[0]
        K, m = X.shape
        if K != P * group or m != n_new * d:
            raise ValueError("need P * group rows of n_new * d coordinates")
        if D_old is None or len(D_old) == 0:
            Do, n_old = None, 0
        else:
            Do = _f(np.atleast_2d(D_old))
            n_old = Do.shape[0]
        S = 2 * m + 1
        vals = np.empty((K, S))
        status = np.empty((K, S), dtype=np.int32)
        self._ck(self._lib.ccgp_me_schur_stencil(self._h, _ptr(Do), n_old, d, _ptr(X), n_new, int(group), _ptr(params), P, P,
                                                 float(h), float(lower), float(upper), _ptr(vals), _ptr(status)))
        return vals, status

    def kmedoids_pam(self, P, k, max_swaps=1000):
        """PAM k-medoids of the rows of P (Euclidean) -> (medoid row indices[k], total cost, swaps)."""
        P = _f(np.atleast_2d(P))
        n, d = P.shape
        med = np.empty(k, dtype=np.int32)
        cost = np.zeros(1)
        swaps = np.zeros(1, dtype=np.int32)
        self._ck(self._lib.ccgp_kmedoids_pam(self._h, _ptr(P), n, d, int(k), int(max_swaps), _ptr(med), _ptr(cost), _ptr(swaps)))
        return med, float(cost[0]), int(swaps[0])

    def cgp_objective_batch(self, Xs, y, W):
        """CGP comparator: var.MLE.DK ([A]:104-135) for the rows (lambda, theta_1..p, kappa, bw) of W on the standardised
        design Xs -> (values[B], status[B])."""
        Xs = _f(np.atleast_2d(Xs))
        y = np.ascontiguousarray(y, dtype=np.float64)
        W = _f(np.atleast_2d(W))
        n, p = Xs.shape
        B = W.shape[0]
        assert W.shape[1] == p + 3 and y.shape[0] == n
        out = np.empty(B)
        st = np.zeros(B, dtype=np.int32)
        self._ck(self._lib.ccgp_cgp_objective_batch(self._h, _ptr(Xs), _ptr(y), n, p, _ptr(W), B, B, _ptr(out), _ptr(st)))
        return out, st

    def cgp_jackknife(self, Xs, y, w):
        """CGP comparator: Yp_jackknife ([A]:166-199) for ONE parameter row w -> (predictions[n], status[n])."""
        Xs = _f(np.atleast_2d(Xs))
        y = np.ascontiguousarray(y, dtype=np.float64)
        w = np.ascontiguousarray(np.asarray(w, dtype=np.float64).reshape(-1))
        n, p = Xs.shape
        assert w.shape[0] == p + 3 and y.shape[0] == n
        out = np.empty(n)
        st = np.zeros(n, dtype=np.int32)
        self._ck(self._lib.ccgp_cgp_jackknife(self._h, _ptr(Xs), _ptr(y), n, p, _ptr(w), _ptr(out), _ptr(st)))
        return out, st

    def mixed_corr(self, params, family, A, B=None):
        """Mixed correlation block between the rows of A and of B (B=None: A with itself)."""
        A = _f(np.atleast_2d(A))
        params = np.ascontiguousarray(np.asarray(params, dtype=np.float64).reshape(-1))
        na, d = A.shape
        if B is None:
            Bm, nb = None, na
        else:
            Bm = _f(np.atleast_2d(B))
            nb = Bm.shape[0]
        out = np.empty((na, nb), order="F")
        self._ck(self._lib.ccgp_mixed_corr(self._h, family, _ptr(params), _ptr(A), na, _ptr(Bm), nb, d, _ptr(out)))
        return out


class Factors:
    """Cholesky factors of S posterior rows resident on the device (ccgp_factors_*, include/ccgp.h), keyed by row index."""

    def __init__(self, engine, handle, S):
        self._eng, self._h, self.S = engine, handle, S

    def info(self):
        rows, stored, nbytes = C.c_int64(), C.c_int(), C.c_int64()
        self._eng._ck(self._eng._lib.ccgp_factors_info(self._eng._h, self._h, C.byref(rows), C.byref(stored), C.byref(nbytes)))
        return dict(rows=rows.value, stored=bool(stored.value), device_bytes=nbytes.value)

    def predict(self, X_new, sigma2):
        """-> (mean[T,S], var[T,S], status[S]) at the sites X_new, from the stored factors."""
        X_new = _f(np.atleast_2d(X_new))
        T, S = X_new.shape[0], self.S
        if X_new.shape[1] != self._eng.d:
            raise ValueError("X_new has %d columns, design has %d" % (X_new.shape[1], self._eng.d))
        mean = np.empty((T, S), order="F")
        var = np.empty((T, S), order="F")
        status = np.empty(S, dtype=np.int32)
        self._eng._ck(self._eng._lib.ccgp_factors_predict(self._eng._h, self._h, _ptr(X_new), T, float(sigma2),
                                                          _ptr(mean), _ptr(var), _ptr(status)))
        return mean, var, status

    def predict_dev(self, Xnew_t, sigma2, out_mean, out_var, out_status=None):
        """torch CUDA tensors (Xnew_t d x T row-major, out_* T*S); single-GPU contexts.  Async."""
        d, T = Xnew_t.shape
        self._eng._ck(self._eng._lib.ccgp_factors_predict_dev(self._eng._h, self._h, Xnew_t.data_ptr(), T, float(sigma2),
                                                              out_mean.data_ptr(), out_var.data_ptr(),
                                                              out_status.data_ptr() if out_status is not None else None))

    def close(self):
        if getattr(self, "_h", None) and getattr(self._eng, "_h", None):
            self._eng._lib.ccgp_factors_destroy(self._eng._h, self._h)
        self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
