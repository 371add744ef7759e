#!/usr/bin/env python
"""bench.py -- headline benchmark of the hot path (BASELINE.json metric M1):
combined-GP NLL evaluations/sec, batched, n=100, 2-D anisotropic, GLS-beta mean.

  python bench.py --gpus N --steps K --warmup W            (N>1: launched by torchrun)
  python bench.py --impl reference --gpus N --steps K --warmup W

A step = one pass of the hot path over one batch of B seeded candidates per GPU
(X = maximin 100 pts, y = simulator 4, sigma2 = 1, candidates
(psi1, psi2, phi, zeta) from default_rng(20131 + rank)).  `value` is timed with CUDA
events on the launching stream, inputs resident in HBM, max over ranks; `e2e` is the
same metric through the host-pointer C-ABI call (H2D of the candidates + D2H of
nll/beta/status inside the timed region).  One JSON line on stdout (rank 0).
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

N_PTS, DIM = 100, 2
# F_nll(n,d) = P*c_pair + (n^3/3 + n^2/2 + n/6) + 2n^2 + 6n, c_pair = 4d+3 (aniso), + 2P exp
P_PAIRS = N_PTS * (N_PTS - 1) // 2
FLOP_PER_EVAL = P_PAIRS * (4 * DIM + 3) + (N_PTS ** 3 / 3 + N_PTS ** 2 / 2 + N_PTS / 6) + 2 * N_PTS ** 2 + 6 * N_PTS
EXP_PER_EVAL = 2 * P_PAIRS
BYTES_PER_EVAL = 52  # 32 B candidate row in, nll + beta + status out
# SURVEY 8(d) ME Schur determinant (n_old = 14, n_new = 7, d = 2 iso): 4 860 FLOP + 238 exp; 112 B design in + 12 B out.
# One table-driven exp is 10 FP64 pipe instructions (ccgp_math.h), i.e. 20 FLOP-equivalents of the same pipe.
ME_FLOP, ME_EXP, ME_EXP_FLOP_EQ, ME_BYTES = 4860.0, 238.0, 20.0, 124.0


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--batch", type=int, default=1 << 20, help="candidates per GPU per step")
    ap.add_argument("--cpu-sample", type=int, default=0, help="candidates in the CPU-baseline sample (0 = auto)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-me", action="store_true")
    ap.add_argument("--scaling", default="weak", choices=["weak", "strong"],
                    help="weak: --batch candidates per GPU (default). strong: ONE --batch-candidate batch split over the ranks, "
                         "the which.min all-reduce inside the timed region; the headline value then is the strong-scaling one")
    ap.add_argument("--no-strong", action="store_true", help="skip the strong-scaling block of the default run")
    ap.add_argument("--cpu-helper", default=None, help=argparse.SUPPRESS)   # internal: "m1:<in.npy>:<out.npy>" / "me:<in.npz>:<out.npy>"
    return ap.parse_args()


# ------------------------------------------------------------------ CPU arm (oracle = the reference's algorithm)
def _cpu_worker(args):
    os.environ["OPENBLAS_NUM_THREADS"] = "1"
    from oracle import ccgp_oracle as orc
    X, y, s2, th = args
    out = np.empty(len(th))
    for i, t in enumerate(th):
        nat = orc.transform_theta(orc.FAMILY_ANISO_LAMBDA, t, 2)
        out[i] = orc.loglik_reference(X, y, s2, orc.FAMILY_ANISO_LAMBDA, nat)["loglik"]
    return out


def _cpu_me_worker(args):
    os.environ["OPENBLAS_NUM_THREADS"] = "1"
    from oracle import ccgp_oracle as orc
    D_old, pool, prm = args
    return orc.me_schur_negdet_batch(D_old, pool, prm)


class CpuArm:
    """The reference's per-candidate algorithm (LU inverse + beta.MLE + dmnorm: chol, chol2inv) via the
    oracle port on every host core, one single-threaded-BLAS worker process per core."""

    def __init__(self):
        import multiprocessing as mp
        from ccgp_b200 import workloads
        self.cores = len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else os.cpu_count()
        self.X, self.y, self.s2 = workloads.m1_design()
        os.environ["OPENBLAS_NUM_THREADS"] = "1"
        self.pool = mp.get_context("fork").Pool(self.cores)
        self.workloads = workloads

    def run(self, th):
        chunks = np.array_split(th, self.cores * 4)
        t0 = time.perf_counter()
        res = self.pool.map(_cpu_worker, [(self.X, self.y, self.s2, c) for c in chunks if len(c)])
        dt = time.perf_counter() - t0
        return dt, np.concatenate(res)

    def close(self):
        self.pool.close()
        self.pool.join()


def cpu_helper(spec):
    """Internal entry (`--cpu-helper`): the CPU-baseline leg in a FRESH process.  Forking worker pools from the GPU
    process (CUDA context, NCCL and BLAS threads alive) deadlocked on the GPU box; a new interpreter that never touches
    CUDA forks its single-threaded-BLAS workers safely.  Prints one JSON line {"dt": seconds, "cores": C}."""
    kind, fin, fout = spec.split(":")
    if kind == "m1":
        th = np.load(fin)
        arm = CpuArm()
        arm.run(th[:min(len(th), 256)])
        dt, ll = arm.run(th)
        arm.close()
        np.save(fout, ll)
        print(json.dumps({"dt": dt, "cores": arm.cores}))
    else:
        import multiprocessing as mp
        z = np.load(fin)
        D_old, pool, params = z["D_old"], z["pool"], z["params"]
        cores = len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else os.cpu_count()
        os.environ["OPENBLAS_NUM_THREADS"] = "1"
        with mp.get_context("fork").Pool(cores) as pl:
            pl.map(_cpu_me_worker, [(D_old, pool[:8], params[:1])] * cores)
            t0 = time.perf_counter()
            parts = pl.map(_cpu_me_worker, [(D_old, pool, params[i:i + 4]) for i in range(0, len(params), 4)])
            dt = time.perf_counter() - t0
        np.save(fout, np.hstack(parts))
        print(json.dumps({"dt": dt, "cores": cores}))


def run_cpu_helper(kind, **arrays):
    """Run cpu_helper(kind) in a fresh interpreter; returns (seconds, cores, result array) or None on failure."""
    import tempfile
    with tempfile.TemporaryDirectory() as td:
        fin = os.path.join(td, "in.npy" if kind == "m1" else "in.npz")
        fout = os.path.join(td, "out.npy")
        if kind == "m1":
            np.save(fin, arrays["th"])
        else:
            np.savez(fin, **arrays)
        env = dict(os.environ, OPENBLAS_NUM_THREADS="1", CUDA_VISIBLE_DEVICES="")
        for k in ("RANK", "WORLD_SIZE", "LOCAL_RANK"):
            env.pop(k, None)
        try:
            r = subprocess.run([sys.executable, os.path.abspath(__file__), "--cpu-helper", "%s:%s:%s" % (kind, fin, fout)],
                               env=env, capture_output=True, text=True, timeout=240)
            info = json.loads(r.stdout.strip().splitlines()[-1])
            return info["dt"], info["cores"], np.load(fout)
        except Exception as ex:  # noqa: BLE001
            _log("cpu helper %s failed: %r" % (kind, ex))
            return None


def blas_name():
    try:
        import scipy
        cfg = scipy.show_config(mode="dicts")
        b = cfg["Build Dependencies"]["blas"]
        return "%s %s" % (b.get("name"), b.get("version"))
    except Exception:
        return "scipy-bundled BLAS/LAPACK"


def run_reference(args, rank, world):
    if rank != 0:
        return
    arm = CpuArm()
    per_step = max(256, 256 * arm.cores)
    th_all = arm.workloads.m1_candidates(per_step * (args.steps + args.warmup))
    k = 0
    for _ in range(args.warmup):
        arm.run(th_all[k:k + per_step]); k += per_step
    total = 0.0
    for _ in range(args.steps):
        dt, _ = arm.run(th_all[k:k + per_step]); k += per_step
        total += dt
    arm.close()
    val = per_step * args.steps / total
    line = {
        "impl": "reference", "metric": "combined-GP NLL evals/sec (batched, n=100)", "value": val, "unit": "evals/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * total / args.steps,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": "M1: n=100 d=2 anisotropic GLS-beta NLL (maximin 100 pts, simulator 4, sigma2=1)",
                   "candidates_per_step": per_step},
        "cpu_baseline": {"value": val, "unit": "evals/s", "cores": arm.cores, "kind": "port",
                         "sample": "%d candidates/step of the M1 stream, reference-faithful path (dgesv+dgecon inverse, "
                                   "dpotrf+dpotri dmnorm), 1 process/core, %s" % (per_step, blas_name())},
        "e2e": {"value": val, "unit": "evals/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    _emit(line)


# ------------------------------------------------------------------ clocks
class ClockSampler:
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.rows = []
        self.proc = None
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(gpu_index), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "200"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.th = threading.Thread(target=self._read, daemon=True)
            self.th.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append((time.perf_counter(), line.strip()))

    def stop(self, t0, t1):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.25)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, mx, reasons, pw = [], [], set(), []
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ts, row in self.rows:
            if ts < t0 or ts > t1 + 0.2:
                continue
            f = [x.strip() for x in row.split(",")]
            try:
                sm.append(float(f[1])); mx.append(float(f[2])); pw.append(float(f[3]))
            except Exception:
                continue
            for nm, val in zip(names, f[5:9]):
                if val.lower().startswith("active"):
                    reasons.add(nm)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "power_w_max": max(pw) if pw else None, "samples": len(sm), "reasons": sorted(reasons)}


# ------------------------------------------------------------------ GPU arm
def run_ours(args, rank, world, local_rank):
    import torch
    import torch.distributed as dist
    import ccgp_b200
    from ccgp_b200 import workloads, sharding, GAUSS_ANISO_LAMBDA, LOGSCALE

    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    eng = ccgp_b200.Engine(local_rank)
    stream = torch.cuda.current_stream(dev)
    eng.set_stream(stream.cuda_stream)
    X, y, s2 = workloads.m1_design()
    eng.set_design(X, y)
    B = args.batch
    th_host = workloads.m1_candidates(B, seed=20131 + rank)          # B x 4, real-line scale
    cand_pinned = torch.from_numpy(np.asfortranarray(th_host).T.copy()).pin_memory()   # 4 x B row-major == B x 4 col-major
    cand_dev = cand_pinned.to(dev)
    nll = torch.empty(B, dtype=torch.float64, device=dev)
    beta = torch.empty(B, dtype=torch.float64, device=dev)
    status = torch.empty(B, dtype=torch.int32, device=dev)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)      # > 126 MB L2

    def step():
        eng.nll_batch_dev(cand_dev, GAUSS_ANISO_LAMBDA, s2, scale=LOGSCALE, out_nll=nll, out_beta=beta, out_status=status)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    _log('engine + inputs ready')
    for _ in range(max(args.warmup, 3)):
        step()
    barrier()
    _log('warm-up done')
    # roofline denominators, measured BEFORE the clock sampler starts (their kernels are not part of the timed region)
    peak = eng.measure_fp64_peak()
    peak_dmma = eng.measure_fp64_peak_dmma()
    barrier()
    _log('FP64 peaks measured')
    sampler = ClockSampler(local_rank) if rank == 0 else None
    launches0 = eng.launch_count
    barrier()
    t_wall0 = time.perf_counter()
    evs = []
    for _ in range(args.steps):
        flush.zero_()                                # L2 flush between timed iterations (outside the event pair)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        step()
        e1.record(stream)
        evs.append((e0, e1))
    barrier()
    t_wall1 = time.perf_counter()
    launches = eng.launch_count - launches0
    step_ms = [a.elapsed_time(b) for a, b in evs]
    total_ms = float(sum(step_ms))
    # the path's only collective: which.min of the likelihood over all ranks (value, global index)
    bv, bi = eng.argmin_dev(nll)
    gmin = sharding.allreduce_argmin(bv, bi + rank * B if bi >= 0 else -1, device=dev)
    if world > 1:
        t = torch.tensor([total_ms], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        total_ms = float(t.item())
    clocks = sampler.stop(t_wall0, t_wall1) if sampler else None
    n_bad = int((status != 0).sum().item())

    _log('timed region done')
    # ---- e2e: host buffers through the C-ABI host entry point, copies inside the timed region
    th_f = np.asfortranarray(th_host)
    eng.set_stream(None)
    eng.nll_batch(th_f, GAUSS_ANISO_LAMBDA, s2, scale=LOGSCALE)
    barrier()
    t0 = time.perf_counter()
    ksteps = max(3, min(args.steps, 10))
    for _ in range(ksteps):
        h_nll, h_beta, h_st = eng.nll_batch(th_f, GAUSS_ANISO_LAMBDA, s2, scale=LOGSCALE)
    barrier()
    e2e_s = time.perf_counter() - t0
    if world > 1:
        t = torch.tensor([e2e_s], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        e2e_s = float(t.item())
    assert np.array_equal(h_nll, nll.cpu().numpy()), "host-pointer path and device path disagree"
    cfg = eng.last_nll_config()

    _log('e2e done')
    # ---- M2: ME subset (Schur) determinants/sec, pool(1000) x params(1000) per step, resident inputs
    me = None
    pred = None
    large_n = None
    if not args.no_me:
        eng.set_stream(stream.cuda_stream)
        D_old, pool = workloads.me_pool()
        P = 1000
        params = workloads.me_params(P, seed=7 + rank)
        d_old = torch.from_numpy(np.asfortranarray(D_old).T.copy()).to(dev)
        d_new = torch.from_numpy(np.ascontiguousarray(pool.transpose(0, 2, 1))).to(dev)
        d_par = torch.from_numpy(np.asfortranarray(params).T.copy()).to(dev)
        negdet = torch.empty(1000 * P, dtype=torch.float64, device=dev)
        for _ in range(3):
            eng.me_schur_batch_dev(d_old, 14, 2, d_new, 7, 1000, d_par, P, negdet)
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        me_steps = 5
        for _ in range(me_steps):
            eng.me_schur_batch_dev(d_old, 14, 2, d_new, 7, 1000, d_par, P, negdet)
        e1.record(stream)
        barrier()
        me_ms = e0.elapsed_time(e1)
        if world > 1:
            t = torch.tensor([me_ms], dtype=torch.float64, device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            me_ms = float(t.item())
        me_kern_s = me_ms * 1e-3 / me_steps
        me_tf = ME_FLOP * 1000 * P / me_kern_s / 1e12
        me_tf_exp = (ME_FLOP + ME_EXP * ME_EXP_FLOP_EQ) * 1000 * P / me_kern_s / 1e12
        # e2e: host buffers through ccgp_me_argmin (H2D of the designs + parameter rows, D2H of the P (value, index) pairs)
        eng.set_stream(None)
        eng.me_argmin(D_old, pool, params)
        barrier()
        t0 = time.perf_counter()
        for _ in range(me_steps):
            bv_me, bi_me = eng.me_argmin(D_old, pool, params)
        barrier()
        me_e2e_s = time.perf_counter() - t0
        if world > 1:
            t = torch.tensor([me_e2e_s], dtype=torch.float64, device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            me_e2e_s = float(t.item())
        eng.set_stream(stream.cuda_stream)
        _log('ME kernel + e2e done')
        me_cpu = None
        if rank == 0 and world == 1 and not args.no_cpu_baseline:
            ncores = len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else os.cpu_count()
            rows = 4 * ncores                                   # 1000 designs x rows parameter rows: a few seconds
            got = run_cpu_helper("me", D_old=D_old, pool=pool, params=params[:rows])
            if got is not None:
                dt, cores, nd_cpu = got
                same = bool(np.array_equal(nd_cpu.argmin(axis=0), bi_me[:rows]))
                me_cpu = {"value": 1000 * rows / dt, "unit": "dets/s", "cores": cores, "kind": "port",
                          "sample": "1000 designs x first %d parameter rows, oracle Augmented.Mixed.Entropy (cross Grams, solve(R.old) once per row, "
                                    "dgetrf det), 1 process/core, fresh interpreter" % rows, "argmin_identical_to_gpu": same}
        me = {"metric": "ME subset (Schur) log-dets/sec", "value": world * 1000 * P * me_steps / (me_ms * 1e-3),
              "unit": "dets/s", "workload": "ME-A: Initial ME Design (14x2) + 1000 All_Subdesigns blocks x 1000 parameter rows per GPU per step",
              "ms_per_step": me_ms / me_steps,
              "roofline": {"bound": "tensor", "bound_detail": "FP64 pipe (DFMA); 7x7 Schur blocks are too small for DMMA tiles",
                           "achieved": me_tf, "peak": peak / 1e12, "unit": "TFLOP/s", "frac": me_tf / (peak / 1e12),
                           "achieved_with_exp": me_tf_exp, "frac_with_exp": me_tf_exp / (peak / 1e12),
                           "flop_per_det": ME_FLOP, "exp_per_det": ME_EXP, "exp_flop_equivalent": ME_EXP_FLOP_EQ,
                           "hbm": {"achieved_gbs": (112.0 * 1000 + 8.0 * 1000 * P) / me_kern_s / 1e9, "peak_gbs": _hbm_peak_gbs()},
                           "traffic": _me_traffic(1000 * P)},
              "cpu_baseline": me_cpu,
              "e2e": {"value": world * 1000 * P * me_steps / me_e2e_s, "unit": "dets/s",
                      "h2d_bytes_per_step": int(1000 * 14 * 8 + P * 24 + 14 * 2 * 8), "d2h_bytes_per_step": int(P * 16),
                      "timed": "wall clock around %d synchronous ccgp_me_argmin host calls (1000 designs x %d rows each)" % (me_steps, P)}}
        _log('ME cpu baseline done')
        # ---- predictive table (the prediction() stage of the same fit): S posterior rows x T = 625 grid sites, n = 100
        S_p, T_p = 1000, 625
        rng_p = np.random.default_rng(4242 + rank)
        pars_p = workloads.m1_candidates(S_p, seed=99 + rank)
        nat_p = np.column_stack([1.0 / (1.0 + np.exp(-pars_p[:, 2])), np.exp(pars_p[:, 0]), np.exp(pars_p[:, 1]), np.exp(pars_p[:, 3])])
        gx = np.linspace(-1.0, 1.0, 25)
        Xn = np.array([(a, b) for a in gx for b in gx])                       # the 25 x 25 grid of [A]:851-853
        d_pp = torch.from_numpy(np.asfortranarray(nat_p).T.copy()).to(dev)
        d_xn = torch.from_numpy(np.asfortranarray(Xn).T.copy()).to(dev)
        pm = torch.empty(S_p * T_p, dtype=torch.float64, device=dev)
        pvv = torch.empty(S_p * T_p, dtype=torch.float64, device=dev)
        for _ in range(3):
            eng.predict_dev(d_pp, GAUSS_ANISO_LAMBDA, d_xn, s2, pm, pvv)
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        for _ in range(me_steps):
            eng.predict_dev(d_pp, GAUSS_ANISO_LAMBDA, d_xn, s2, pm, pvv)
        e1.record(stream)
        barrier()
        p_ms = e0.elapsed_time(e1)
        if world > 1:
            t = torch.tensor([p_ms], dtype=torch.float64, device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            p_ms = float(t.item())
        # SURVEY 8(d) counts the reference's algorithm (explicit R^-1: r'R^-1 r = 2 n^2 per site); the kernel solves
        # L v = r instead (n^2 per site) -- both fractions are reported
        pflop = S_p * T_p * (100 * (3 * 2 + 4) + 2.0 * 100 * 100 + 6 * 100) + S_p * 100 ** 3 / 3.0
        pflop_exec = S_p * T_p * (100 * (3 * 2 + 4) + 1.0 * 100 * 100 + 6 * 100) + S_p * 100 ** 3 / 3.0
        p_s = p_ms * 1e-3 / me_steps
        pred = {"metric": "predictive (posterior row, site) pairs/sec", "value": world * S_p * T_p * me_steps / (p_ms * 1e-3),
                "unit": "pairs/s", "workload": "n=100 d=2 anisotropic, S=1000 posterior rows x T=625 (25x25 grid) per GPU per step",
                "ms_per_step": p_ms / me_steps, "tflops_algorithmic": pflop / p_s / 1e12,
                "roofline": {"bound": "tensor", "peak": peak / 1e12, "unit": "TFLOP/s",
                             "achieved": pflop / p_s / 1e12, "frac": pflop / p_s / peak,
                             "achieved_executed": pflop_exec / p_s / 1e12, "frac_executed": pflop_exec / p_s / peak,
                             "note": "frac: SURVEY 8(d) count of the reference's explicit-inverse algorithm (2 n^2 per site); "
                                     "frac_executed: the forward-solve algorithm the kernel runs (n^2 per site); 2 n exp per site in neither"},
                "finite": bool(torch.isfinite(pm).all().item())}
        # the same table from factors kept on the device (ccgp_factors_*: the device-side factors.frame): site phase only
        fac = eng.factors(nat_p, GAUSS_ANISO_LAMBDA)
        pm2 = torch.empty_like(pm)
        pv2 = torch.empty_like(pvv)
        for _ in range(3):
            fac.predict_dev(d_xn, s2, pm2, pv2)
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        for _ in range(me_steps):
            fac.predict_dev(d_xn, s2, pm2, pv2)
        e1.record(stream)
        barrier()
        f_ms = e0.elapsed_time(e1)
        if world > 1:
            t = torch.tensor([f_ms], dtype=torch.float64, device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            f_ms = float(t.item())
        finfo = fac.info()
        pred["from_stored_factors"] = {"value": world * S_p * T_p * me_steps / (f_ms * 1e-3), "unit": "pairs/s", "ms_per_step": f_ms / me_steps,
                                       "device_bytes": finfo["device_bytes"], "stored": finfo["stored"],
                                       "bit_identical_to_direct": bool(torch.equal(pm, pm2) and torch.equal(pvv, pv2))}
        fac.close()
        # ---- large-n blocked FP64-tensor path (north_star: synthetic n = 2048, 2-D anisotropic; SURVEY 8d ME-B(1)), rank 0
        if rank == 0:
            n_big, B_big = 2048, 64
            Xb = workloads.synthetic_pool(n_big, seed=2048)
            yb = np.sin(3 * Xb[:, 0]) * np.cos(2 * Xb[:, 1])
            rng_b = np.random.default_rng(1)
            h2 = 4.0 / n_big
            nat_b = np.column_stack([rng_b.uniform(0.2, 0.8, B_big), rng_b.uniform(1.2, 2.5, B_big) / h2,
                                     rng_b.uniform(1.2, 2.5, B_big) / h2, rng_b.uniform(0.5, 2, B_big)])
            eng.set_design(Xb, yb)
            d_cb = torch.from_numpy(np.asfortranarray(nat_b).T.copy()).to(dev)
            ob = eng.nll_batch_dev(d_cb, GAUSS_ANISO_LAMBDA, 1.0)
            eng.nll_batch_dev(d_cb, GAUSS_ANISO_LAMBDA, 1.0, out_nll=ob[0], out_beta=ob[1], out_status=ob[2])
            torch.cuda.synchronize(dev)
            l0 = eng.launch_count
            evs = []
            for _ in range(9):                                   # queued back to back; clocks sampled while they run
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record(stream)
                eng.nll_batch_dev(d_cb, GAUSS_ANISO_LAMBDA, 1.0, out_nll=ob[0], out_beta=ob[1], out_status=ob[2])
                e1.record(stream)
                evs.append((e0, e1))
            big_clk, big_pw = [], []
            try:
                import pynvml
                pynvml.nvmlInit()
                hnd = pynvml.nvmlDeviceGetHandleByIndex(local_rank)
                while not evs[-1][1].query():
                    big_clk.append(pynvml.nvmlDeviceGetClockInfo(hnd, pynvml.NVML_CLOCK_SM))
                    big_pw.append(pynvml.nvmlDeviceGetPowerUsage(hnd) / 1000.0)
                    time.sleep(0.01)
            except Exception:
                pass
            torch.cuda.synchronize(dev)
            tb = [a.elapsed_time(b) for a, b in evs]
            b_ms = sorted(tb)[len(tb) // 2]
            bflop = B_big * (n_big ** 3 / 3.0 + 2.5 * n_big * n_big)
            large_n = {"metric": "large-n NLL evals/sec (blocked FP64 tensor path)", "value": B_big / (b_ms * 1e-3), "unit": "evals/s",
                       "workload": "synthetic n=2048 d=2 anisotropic, 64 candidates per batch, matrices in HBM (2.2 GB)",
                       "ms_per_batch": b_ms, "ms_per_batch_all": [round(t, 3) for t in tb], "launches_per_batch": int((eng.launch_count - l0) // 9),
                       "sm_mhz_during": (float(np.median(big_clk)) if big_clk else None), "power_w_max_during": (max(big_pw) if big_pw else None),
                       "roofline": {"bound": "tensor", "achieved": bflop / (b_ms * 1e-3) / 1e12, "peak": peak / 1e12, "unit": "TFLOP/s",
                                    "frac": bflop / (b_ms * 1e-3) / peak, "flop_per_eval": bflop / B_big},
                       "not_pd_candidates": int((ob[2] != 0).sum().item())}
            eng.set_design(X, y)
        eng.set_stream(None)

    _log('ME/predict blocks done')
    # ---- strong scaling: ONE batch split over the ranks through the host-pointer call (H2D + kernel + D2H of the rank's
    # slice), then the path's collective -- which.min as two NCCL all-reduces -- INSIDE the timed region.  Three batches:
    # the headline 2^20 stream and the reference's own sweeps ([V]:552-599: 60 x 1728 at n = 14; [H]:549-595: 624 x 1000 at n = 64).
    strong = None
    if not args.no_strong:
        from ccgp_b200 import reference_api as api
        from ccgp_b200 import GAUSS_ISO, MEAN_ZERO_PLUS_TAU2
        eng.set_stream(None)
        strong = {}

        def timed_split(tag, Xd, yd, cand, family, sigma2, scale, mean_mode, tau, reps):
            eng.set_design(Xd, yd)
            total = cand.shape[0]
            lo, hi = sharding.shard_range(total, rank, world)
            mine = np.asfortranarray(cand[lo:hi])
            best = None
            for it in range(reps + 1):
                barrier()
                t0 = time.perf_counter()
                h_nll, _, _ = eng.nll_batch(mine, family, sigma2, scale=scale, mean_mode=mean_mode, tau=tau)
                i = int(np.nanargmin(h_nll)) if np.isfinite(h_nll).any() else -1
                g = sharding.allreduce_argmin(float(h_nll[i]) if i >= 0 else float("nan"), lo + i if i >= 0 else -1, device=dev)
                torch.cuda.synchronize(dev)
                dt = time.perf_counter() - t0
                if world > 1:
                    t = torch.tensor([dt], dtype=torch.float64, device=dev)
                    dist.all_reduce(t, op=dist.ReduceOp.MAX)
                    dt = float(t.item())
                if it > 0:
                    best = dt if best is None else min(best, dt)
            strong[tag] = {"candidates": int(total), "ms": best * 1e3, "value": total / best, "unit": "evals/s",
                           "argmin": {"nll": g[0], "index": g[1]}}

        timed_split("m1_2^20_n100", X, y, workloads.m1_candidates(1 << 20, seed=20131), GAUSS_ANISO_LAMBDA, s2, LOGSCALE, 0, 0.0, 3)
        dsg = workloads.designs()
        X14 = dsg["maximin14"]
        hp = dsg["hyperpars_2d"]
        c14 = np.vstack([api.sweep_candidates(hp[i, 0:2], hp[i, 2:4], 1728) for i in range(hp.shape[0])])
        timed_split("choose_hyperpars_60x1728_n14", X14, workloads.test_function_4(X14), c14, GAUSS_ISO, 0.7, 0, MEAN_ZERO_PLUS_TAU2, 100.0, 5)
        he, hph = dsg["he_train"], dsg["he_hyperpars"]
        c64 = np.vstack([api.sweep_candidates(hph[i, 0:2], hph[i, 2:4], 1000) for i in range(hph.shape[0])])
        timed_split("choose_hyperpars_624x1000_n64", he[:, :4], he[:, 4], c64, GAUSS_ISO, 30.0, 0, MEAN_ZERO_PLUS_TAU2, 50.0, 5)
        strong["timed"] = ("per batch: barrier, then wall clock around [ccgp_nll_batch on the rank's contiguous slice (host buffers: H2D, kernel, D2H) + "
                           "which.min of the slice + 2 NCCL all-reduces (MIN value, MIN index of the winners)], max over ranks, best of the repeats")
        eng.set_design(X, y)

    _log('strong block done')
    # ---- CPU baseline (rank 0, N=1 only): bounded sample of the same workload
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        ncores = len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else os.cpu_count()
        ns = args.cpu_sample or max(2048, 2048 * ncores // 8 * 1)
        ns = min(ns, B)
        got_cpu = run_cpu_helper("m1", th=th_host[:ns])
        if got_cpu is not None:
            dt, cores, ll = got_cpu
            got = -nll[:ns].cpu().numpy()
            relerr = float(np.max(np.abs(got - ll) / np.maximum(np.abs(ll), 1.0)))
            cpu = {"value": ns / dt, "unit": "evals/s", "cores": cores, "kind": "port",
                   "sample": "first %d candidates of the step's batch, reference-faithful oracle path (dgesv+dgecon inverse, "
                             "dpotrf+dpotri dmnorm), 1 process/core in a fresh interpreter, %s" % (ns, blas_name()),
                   "max_rel_err_gpu_vs_cpu": relerr}

    _log('cpu baseline done')
    if rank == 0:
        secs = total_ms * 1e-3
        evals = world * B * args.steps
        value = evals / secs
        kern_s = (total_ms / args.steps) * 1e-3                     # one kernel launch per step per GPU
        achieved_tf = FLOP_PER_EVAL * B / kern_s / 1e12
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except Exception:
            pass
        hbm_peak = peaks.get("hbm_gbs", 6650.0)
        traffic = None
        tpath = os.path.join(ROOT, "profiles", "traffic.json")
        if os.path.exists(tpath):
            try:
                tj = json.load(open(tpath))
                # ncu dram__bytes_read+write of one captured launch, scaled to this launch's candidate count
                traffic = tj["nll_kernel_dram_bytes_per_launch"] / tj["candidates_per_launch"] * B
            except Exception:
                traffic = None
        line = {
            "metric": "combined-GP NLL evals/sec (batched, n=100)", "value": value, "unit": "evals/s",
            "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3), "ms_per_step": total_ms / args.steps,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": "M1: n=100 d=2 anisotropic GLS-beta NLL (maximin 100 pts, simulator 4, sigma2=1)",
                       "candidates_per_gpu_per_step": B, "parallelism": "candidates sharded, dp%d" % world,
                       "l2": "flushed between timed steps (256 MiB memset outside the event pairs); working set/step = %d MiB"
                             % ((B * BYTES_PER_EVAL) >> 20),
                       "kernel": cfg, "not_pd_candidates": n_bad,
                       "argmin": {"nll": gmin[0], "index": gmin[1]}},
            "roofline": {"bound": "tensor", "bound_detail": "FP64 pipe: DFMA and the FP64 tensor form (mma.sync.m8n8k4.f64, DMMA) share it at 64 FMA/clk/SM; tcgen05 has no FP64 kind", "achieved": achieved_tf, "peak": peak / 1e12, "unit": "TFLOP/s",
                         "frac": achieved_tf / (peak / 1e12),
                         "peak_source": "measured live before the timed region: dependent-free DFMA loop on all SMs (MEASURED_PEAKS.json has no FP64 entry)",
                         "peak_detail": {"dfma_loop_tflops": peak / 1e12, "dmma_m8n8k4_loop_tflops": peak_dmma / 1e12,
                                         "theoretical_tflops": 148 * 128 * 1.965e9 / 1e12,
                                         "note": "no library DGEMM is linked (libccgp.so has no cuBLAS dependency); the DMMA loop is the GEMM-shaped cross-check"},
                         "executed_floor": "FP64 pipe time per candidate ~30 k clk of a sub-partition (15 k DMMA + 10.5 k build DFMA + 4.2 k diagonal chain) = 47 % of peak on the FLOP-only count at 100 % pipe occupancy: the 9 900 exponentials are not in flop_per_eval",
                         "flop_per_eval": FLOP_PER_EVAL, "exp_per_eval": EXP_PER_EVAL,
                         "exp_per_s": EXP_PER_EVAL * B / kern_s,
                         "hbm": {"achieved_gbs": BYTES_PER_EVAL * B / kern_s / 1e9, "peak_gbs": hbm_peak,
                                 "frac": BYTES_PER_EVAL * B / kern_s / 1e9 / hbm_peak},
                         "traffic": traffic},
            "cpu_baseline": cpu,
            "e2e": {"value": world * B * ksteps / e2e_s, "unit": "evals/s", "h2d_bytes_per_step": int(B * 32),
                    "d2h_bytes_per_step": int(B * 20), "timed": "wall clock around %d synchronous ccgp_nll_batch host calls" % ksteps},
            "gpu_launches": int(launches),
            "clocks": clocks,
            "me": me, "predict": pred, "large_n": large_n, "strong": strong,
        }
        if args.scaling == "strong" and strong:
            st = strong["m1_2^20_n100"]
            line.update({"scaling": "strong", "value": st["value"], "ms_per_step": st["ms"],
                         "weak": {"value": value, "ms_per_step": total_ms / args.steps}})
            line["config"]["parallelism"] = "ONE batch of 2^20 candidates split contiguously over %d ranks, which.min all-reduce in the timed region" % world
        _emit(line)
    eng.close()
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


_RESULT_FD = None
_T0 = time.perf_counter()


def _log(msg):
    """progress line on stderr (elapsed wall time since start), rank 0 only"""
    if int(os.environ.get("RANK", "0")) == 0:
        sys.stderr.write("[bench %7.1fs] %s\n" % (time.perf_counter() - _T0, msg))
        sys.stderr.flush()


def _emit(line):
    """The one JSON line goes to the real stdout; everything else libraries print (NCCL's version banner,
    warnings) was redirected to stderr for the duration of the run."""
    data = (json.dumps(line) + "\n").encode()
    if _RESULT_FD is None:
        sys.stdout.write(data.decode())
        sys.stdout.flush()
    else:
        os.write(_RESULT_FD, data)


def _hbm_peak_gbs():
    """measured copy bandwidth of this pool's B200s (driver-written MEASURED_PEAKS.json), else the profiling recipe's fallback"""
    try:
        return float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json"))).get("hbm_gbs", 6650.0))
    except Exception:  # noqa: BLE001
        return 6650.0


def _me_traffic(dets):
    """ncu DRAM bytes of one captured ME launch (profiles/traffic.json), scaled to this launch's determinant count."""
    try:
        tj = json.load(open(os.path.join(ROOT, "profiles", "traffic.json")))
        return tj["me_kernel_dram_bytes_per_launch"] / tj["me_dets_per_launch"] * dets
    except Exception:
        return None


def main():
    global _RESULT_FD
    args = parse()
    sys.stdout.flush()
    _RESULT_FD = os.dup(1)          # keep the real stdout for the result line
    os.dup2(2, 1)                   # anything else written to fd 1 (by C libraries too) lands on stderr
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.cpu_helper:
        os.dup2(_RESULT_FD, 1)
        cpu_helper(args.cpu_helper)
        return
    if args.impl == "reference":
        run_reference(args, rank, world)
        return
    if world != args.gpus and world == 1 and args.gpus > 1:
        # plain `python bench.py --gpus N`: re-launch under torchrun, one rank per GPU
        cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(args.gpus),
               "--master-addr", "127.0.0.1", "--master-port", str(29400 + os.getpid() % 500)] + sys.argv
        sys.exit(subprocess.call(cmd, stdout=_RESULT_FD))
    run_ours(args, rank, world, local_rank)


if __name__ == "__main__":
    main()
