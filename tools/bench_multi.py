"""Strong scaling INSIDE libccgp.so (ccgp_create_multi): one caller thread, one batch, G GPUs -- what an R script
gets through .Call.  For G in {1, 2, 4, 8} (as many as the box has): wall clock around ccgp_nll_argmin (host
buffers in, (min, index) out, NCCL all-reduce inside), ccgp_nll_batch, ccgp_me_argmin and ccgp_predict on the
headline batch and on the reference's own sweep sizes.  One JSON line per (G, workload).
usage: python tools/bench_multi.py [out.jsonl]        (single process; run under gpurun --gpus N)"""
import json
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import ccgp_b200  # noqa: E402
from ccgp_b200 import workloads, GAUSS_ANISO_LAMBDA, GAUSS_ISO, LOGSCALE, MEAN_ZERO_PLUS_TAU2  # noqa: E402
from ccgp_b200 import reference_api as api  # noqa: E402

out_path = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "gpurun_out", "bench_multi.jsonl")
ngpu = torch.cuda.device_count()
X, y, s2 = workloads.m1_design()
th = np.asfortranarray(workloads.m1_candidates(1 << 20))
dsg = workloads.designs()
X14 = dsg["maximin14"]
y14 = workloads.test_function_4(X14)
hp = dsg["hyperpars_2d"]
c14 = np.asfortranarray(np.vstack([api.sweep_candidates(hp[i, 0:2], hp[i, 2:4], 1728) for i in range(60)]))
he, hph = dsg["he_train"], dsg["he_hyperpars"]
c64 = np.asfortranarray(np.vstack([api.sweep_candidates(hph[i, 0:2], hph[i, 2:4], 1000) for i in range(624)]))
D_old, pool = workloads.me_pool()
me_par = workloads.me_params(1000)
pars_p = workloads.m1_candidates(1000, seed=99)
nat_p = np.column_stack([1.0 / (1.0 + np.exp(-pars_p[:, 2])), np.exp(pars_p[:, 0]), np.exp(pars_p[:, 1]), np.exp(pars_p[:, 3])])
gx = np.linspace(-1.0, 1.0, 25)
Xn = np.array([(a, b) for a in gx for b in gx])


def best_of(fn, reps=5):
    fn()
    ts = []
    for _ in range(reps):
        t0 = time.perf_counter()
        r = fn()
        ts.append(time.perf_counter() - t0)
    return min(ts), r


lines = []
base = {}
for G in [g for g in (1, 2, 4, 8) if g <= ngpu]:
    eng = ccgp_b200.Engine(n_gpus=G)
    runs = []
    eng.set_design(X, y)
    runs.append(("nll_argmin 2^20 candidates, n=100 aniso (M1 stream)", 1 << 20, "evals/s",
                 lambda: eng.nll_argmin(th, GAUSS_ANISO_LAMBDA, s2, scale=LOGSCALE)))
    runs.append(("nll_batch 2^20 candidates, n=100 aniso (nll, beta, status back to the host)", 1 << 20, "evals/s",
                 lambda: eng.nll_batch(th, GAUSS_ANISO_LAMBDA, s2, scale=LOGSCALE)[0][:1]))
    runs.append(("predict S=1000 x T=625, n=100", 625000, "pairs/s", lambda: eng.predict(nat_p, GAUSS_ANISO_LAMBDA, Xn, s2)[0][:1, :1]))
    for name, units, unit, fn in runs:
        dt, r = best_of(fn)
        lines.append({"gpus": G, "workload": name, "ms": dt * 1e3, "value": units / dt, "unit": unit, "result": [float(v) for v in np.ravel(r)[:2]]})
    eng.set_design(X14, y14)
    dt, r = best_of(lambda: eng.nll_argmin(c14, GAUSS_ISO, 0.7, mean_mode=MEAN_ZERO_PLUS_TAU2, tau=100.0))
    lines.append({"gpus": G, "workload": "choose.hyperpars sweep 60 x 1728, n=14 ([V]:552-599), nll_argmin", "ms": dt * 1e3, "value": len(c14) / dt,
                  "unit": "evals/s", "result": [float(r[0]), float(r[1])]})
    eng.set_design(he[:, :4], he[:, 4])
    dt, r = best_of(lambda: eng.nll_argmin(c64, GAUSS_ISO, 30.0, mean_mode=MEAN_ZERO_PLUS_TAU2, tau=50.0))
    lines.append({"gpus": G, "workload": "choose.hyperpars sweep 624 x 1000, n=64 ([H]:549-595), nll_argmin", "ms": dt * 1e3, "value": len(c64) / dt,
                  "unit": "evals/s", "result": [float(r[0]), float(r[1])]})
    dt, r = best_of(lambda: eng.me_argmin(D_old, pool, me_par))
    lines.append({"gpus": G, "workload": "me_argmin 1000 designs x 1000 parameter rows (rows split, no collective)", "ms": dt * 1e3, "value": 1e6 / dt,
                  "unit": "dets/s", "result": [float(r[1][0]), float(r[1][-1])]})
    os.environ["CCGP_MULTI_ME_SPLIT_DESIGNS"] = "1"
    dt, r = best_of(lambda: eng.me_argmin(D_old, pool, me_par))
    os.environ.pop("CCGP_MULTI_ME_SPLIT_DESIGNS")
    lines.append({"gpus": G, "workload": "me_argmin 1000 designs x 1000 parameter rows (designs split, NCCL (min,index) per row)", "ms": dt * 1e3,
                  "value": 1e6 / dt, "unit": "dets/s", "result": [float(r[1][0]), float(r[1][-1])]})
    for ln in lines:
        if ln["gpus"] == G:
            if G == 1:
                base[ln["workload"]] = ln
            b = base.get(ln["workload"])
            if b:
                ln["speedup_vs_1gpu"] = b["ms"] / ln["ms"]
                ln["efficiency"] = b["ms"] / ln["ms"] / G
                ln["same_result_as_1gpu"] = ln["result"] == b["result"]
    eng.close()
with open(out_path, "w") as f:
    for ln in lines:
        f.write(json.dumps(ln) + "\n")
        print(json.dumps(ln))
