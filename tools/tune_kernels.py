"""Kernel choice by design size: times every NLL kernel family on synthetic designs of the reference's sizes.
usage: python tools/tune_kernels.py            (prints one line per (n, d, kernel))"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import ccgp_b200  # noqa: E402
from ccgp_b200 import GAUSS_ISO, GAUSS_ANISO_LAMBDA  # noqa: E402

dev = torch.device("cuda", 0)
eng = ccgp_b200.Engine(0)
stream = torch.cuda.current_stream(dev)
eng.set_stream(stream.cuda_stream)
KEYS = ("CCGP_MMA_NW", "CCGP_NO_MMA", "CCGP_TEAM_NW", "CCGP_TEAM_FUSED", "CCGP_KERNEL", "CCGP_VARIANT")
configs = [("auto", {}), ("team nw4", {"CCGP_KERNEL": "3", "CCGP_TEAM_NW": "4"}), ("team nw3", {"CCGP_KERNEL": "3", "CCGP_TEAM_NW": "3"}), ("team nw3 fused", {"CCGP_KERNEL": "3", "CCGP_TEAM_NW": "3", "CCGP_TEAM_FUSED": "1"}),
           ("team nw2", {"CCGP_KERNEL": "3", "CCGP_TEAM_NW": "2"}), ("warp", {"CCGP_KERNEL": "1"}),
           ("cta nw4", {"CCGP_KERNEL": "4"}), ("cta nw2", {"CCGP_KERNEL": "4", "CCGP_MMA_NW": "2"}),
           ("dfma default", {"CCGP_NO_MMA": "1"}), ("dfma v0 (1 warp)", {"CCGP_NO_MMA": "1", "CCGP_VARIANT": "0"})]
for k in KEYS:
    os.environ.pop(k, None)
os.environ["CCGP_MMA_MIN_NPAD"] = "0"
rng = np.random.default_rng(5)
cases = [(14, 2, GAUSS_ISO), (21, 2, GAUSS_ISO), (30, 2, GAUSS_ISO), (50, 9, GAUSS_ISO), (64, 4, GAUSS_ISO), (90, 9, GAUSS_ISO),
         (100, 2, GAUSS_ANISO_LAMBDA), (110, 2, GAUSS_ANISO_LAMBDA), (128, 2, GAUSS_ANISO_LAMBDA), (200, 2, GAUSS_ANISO_LAMBDA)]
if len(sys.argv) > 1:
    want = [int(a) for a in sys.argv[1:]]
    cases = [c for c in cases if c[0] in want]
for n, d, fam in cases:
    X = rng.uniform(-1, 1, (n, d))
    y = rng.normal(size=n)
    eng.set_design(X, y)
    B = int(min(1 << 20, max(1 << 15, (1 << 30) // (n * n * n // 3 + 1))))
    k = eng.num_params(fam)
    th = 8.0 / d * n ** (1.0 / d)
    cand = np.column_stack([rng.uniform(0.2, 0.8, B)] + [rng.uniform(0.5 * th, 1.5 * th, B) for _ in range(k - 1)])
    cd = torch.from_numpy(np.asfortranarray(cand).T.copy()).to(dev)
    nll = torch.empty(B, dtype=torch.float64, device=dev)
    beta = torch.empty(B, dtype=torch.float64, device=dev)
    status = torch.empty(B, dtype=torch.int32, device=dev)
    ref = None
    for name, env in configs:
        for kk in KEYS:
            os.environ.pop(kk, None)
        os.environ.update(env)
        ts = []
        try:
            for it in range(4):
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record(stream)
                eng.nll_batch_dev(cd, fam, 1.0, out_nll=nll, out_beta=beta, out_status=status)
                e1.record(stream)
                torch.cuda.synchronize(dev)
                ts.append(e0.elapsed_time(e1))
        except Exception as ex:  # noqa: BLE001
            print("n=%4d d=%d %-18s failed: %s" % (n, d, name, ex))
            continue
        v = nll.cpu().numpy()
        if ref is None:
            ref = v.copy()
        ok = np.isfinite(v)
        err = float(np.max(np.abs(v[ok] - ref[ok]) / np.maximum(1.0, np.abs(ref[ok])))) if ok.any() else float("nan")
        ms = min(ts[1:])
        print("n=%4d d=%d B=%8d %-18s %8.3f ms %8.2f M evals/s  variant %s  finite %.3f  maxrel vs first %.1e" % (
            n, d, B, name, ms, B / ms / 1e3, eng.last_nll_config()["variant"], ok.mean(), err))
    sys.stdout.flush()
