"""A/B timing of the NLL kernels on the M1 workload (n=100, d=2 aniso) and the reference's other design sizes:
the shipped choice vs the packed-residency kernel vs its producer/consumer split (factor_pc.cuh, CCGP_KERNEL=6), which must
be bit-identical to the packed kernel.  usage: python tools/time_pc.py [n ...]"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import ccgp_b200  # noqa: E402
from ccgp_b200 import workloads, GAUSS_ISO, GAUSS_ANISO_LAMBDA, LOGSCALE  # noqa: E402

dev = torch.device("cuda", 0)
eng = ccgp_b200.Engine(0)
stream = torch.cuda.current_stream(dev)
eng.set_stream(stream.cuda_stream)
KEYS = ("CCGP_KERNEL", "CCGP_PACK_WARPS", "CCGP_PACK_EVEN", "CCGP_TEAM_NW", "CCGP_DUO_CANDS", "CCGP_TEAM_MAP")
sizes = [int(a.split(":")[0]) for a in sys.argv[1:]] or [100]
dims = {int(a.split(":")[0]): int(a.split(":")[1]) for a in sys.argv[1:] if ":" in a}       # "n:d" overrides the dimension
rng = np.random.default_rng(5)
for n in sizes:
    if n == 100:
        X, y, s2 = workloads.m1_design()
        fam, d = GAUSS_ANISO_LAMBDA, 2
        B = (1 << 18) + 37
        cand = workloads.m1_candidates(B)
        scale = LOGSCALE
    else:
        d = dims.get(n, {14: 2, 50: 9, 64: 4, 90: 9}.get(n, 2))
        fam = GAUSS_ISO if d > 2 else GAUSS_ANISO_LAMBDA
        X = rng.uniform(-1, 1, (n, d)); y = rng.normal(size=n); s2 = 1.0
        B = int(min(1 << 20, max(1 << 16, (1 << 31) // (n * n * n // 3 + 1))))
        k = eng.num_params(fam)
        th = 8.0 / d * n ** (1.0 / d)
        cand = np.column_stack([rng.uniform(0.2, 0.8, B)] + [rng.uniform(0.5 * th, 1.5 * th, B) for _ in range(k - 1)])
        scale = 0
    eng.set_design(X, y)
    cd = torch.from_numpy(np.asfortranarray(cand).T.copy()).to(dev)
    nll = torch.empty(B, dtype=torch.float64, device=dev)
    beta = torch.empty(B, dtype=torch.float64, device=dev)
    status = torch.empty(B, dtype=torch.int32, device=dev)
    configs = [("pack w8", {"CCGP_KERNEL": "5"}), ("auto", {})] + ([("prod/cons", {"CCGP_KERNEL": "6"})] if os.environ.get("TIME_PC", "1") == "1" else [])
    ref = None
    for name, env in configs:
        for kk in KEYS:
            os.environ.pop(kk, None)
        os.environ.update(env)
        ts = []
        for it in range(4):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(stream)
            eng.nll_batch_dev(cd, fam, s2, scale=scale, out_nll=nll, out_beta=beta, out_status=status)
            e1.record(stream)
            torch.cuda.synchronize(dev)
            ts.append(e0.elapsed_time(e1))
        v = nll.cpu().numpy(); bv = beta.cpu().numpy()
        if ref is None:
            ref = (v.copy(), bv.copy())
        ok = np.isfinite(v) & np.isfinite(ref[0])
        err = float(np.max(np.abs(v[ok] - ref[0][ok]) / np.maximum(1.0, np.abs(ref[0][ok]))))
        errb = float(np.max(np.abs(bv[ok] - ref[1][ok]) / np.maximum(1.0, np.abs(ref[1][ok]))))
        same_nan = bool(np.array_equal(np.isfinite(v), np.isfinite(ref[0])))
        ms = min(ts[1:])
        cfg = eng.last_nll_config()
        print("n=%4d d=%d B=%8d %-10s %8.3f ms %8.2f M evals/s  variant %s warps/ctas %s smem %s  maxrel nll %.1e beta %.1e  same NaN set %s" % (
            n, d, B, name, ms, B / ms / 1e3, cfg["variant"], cfg["ctas_per_sm"], cfg["smem_bytes"], err, errb, same_nan))
        sys.stdout.flush()
