"""GPU tuner: time every factor-kernel variant on the SURVEY 8(d) workload shapes.
Run on the B200 box:  python tools/tune.py [--quick]  -> gpurun_out/tune.json"""
import json
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import ccgp_b200  # noqa: E402
from ccgp_b200 import workloads, GAUSS_ISO, GAUSS_ANISO_LAMBDA, LOGSCALE, MEAN_ZERO_PLUS_TAU2  # noqa: E402

VARIANTS = [(32, 4, 4), (32, 8, 4), (64, 4, 4), (64, 8, 4), (64, 4, 8), (128, 4, 4), (128, 8, 4), (128, 4, 8),
            (256, 4, 4), (256, 4, 4), (32, 8, 4), (32, 4, 4), (32, 8, 8), (96, 4, 4), (64, 8, 4), (128, 4, 4),
            (32, 4, 4), (32, 8, 4), (32, 4, 4), (32, 4, 4), (32, 4, 8)]


def main():
    quick = "--quick" in sys.argv
    only = [a for a in sys.argv[1:] if not a.startswith("--")]
    dev = torch.device("cuda", 0)
    eng = ccgp_b200.Engine(0)
    stream = torch.cuda.current_stream(dev)
    eng.set_stream(stream.cuda_stream)
    D = workloads.designs()
    rng = np.random.default_rng(0)
    cases = {}
    X, y, s2 = workloads.m1_design()
    cases["n100_d2_aniso"] = (X, y, s2, GAUSS_ANISO_LAMBDA, LOGSCALE, 0, 0.0, lambda B: workloads.m1_candidates(B), 1 << 18)
    X14 = D["maximin14"]
    cases["n14_d2_iso_tau"] = (X14, workloads.test_function_4(X14), 0.7, GAUSS_ISO, 0, MEAN_ZERO_PLUS_TAU2, 100.0,
                               lambda B: np.column_stack([rng.uniform(0.05, 0.95, B), 1 / rng.gamma(3, 0.5, B), 1 / rng.gamma(5, 1 / 16.0, B)]), 1 << 20)
    he = D["he_train"]
    cases["n64_d4_iso"] = (he[:, :4], he[:, 4], 30.0, GAUSS_ISO, 0, 0, 0.0,
                           lambda B: np.column_stack([rng.uniform(0.05, 0.95, B), 1 / rng.gamma(3, 1.0, B), 1 / rng.gamma(5, 1 / 40.0, B)]), 1 << 19)
    for tag in ("gv50", "gv90"):
        tr = D[tag + "_train1"]
        cases[tag + "_d9_iso"] = (tr[:, :9], tr[:, 9], 13.0, GAUSS_ISO, 0, 0, 0.0,
                                  lambda B: np.column_stack([rng.uniform(0.05, 0.95, B), 0.06 / rng.gamma(3, 1.0, B), 1 / rng.gamma(5, 0.5, B)]), 1 << 18)
    results = {}
    for name, (Xc, yc, s2c, fam, scale, mm, tau, gen, B) in cases.items():
        if only and name not in only:
            continue
        if quick:
            B //= 4
        eng.set_design(Xc, yc)
        cand = torch.from_numpy(np.asfortranarray(gen(B)).T.copy()).to(dev)
        res = {}
        for v, (team, tr, ks) in enumerate(VARIANTS):
            os.environ["CCGP_VARIANT"] = str(v)
            try:
                out = eng.nll_batch_dev(cand, fam, s2c, scale=scale, mean_mode=mm, tau=tau)
                torch.cuda.synchronize()
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record(stream)
                eng.nll_batch_dev(cand, fam, s2c, scale=scale, mean_mode=mm, tau=tau, out_nll=out[0], out_beta=out[1], out_status=out[2])
                e1.record(stream)
                torch.cuda.synchronize()
                ms = e0.elapsed_time(e1)
                cfg = eng.last_nll_config()
                nbad = int((out[2] != 0).sum().item())
                res[v] = dict(team=team, tr=tr, ks=ks, ms=ms, evals_per_s=B / (ms * 1e-3), ctas_per_sm=cfg["ctas_per_sm"], smem=cfg["smem_bytes"], bad=nbad)
                print("%-16s v%-2d team=%-3d TR=%d KS=%d ctas/SM=%-2d  %9.3f ms  %12.0f evals/s  bad=%d" % (
                    name, v, team, tr, ks, cfg["ctas_per_sm"], ms, B / (ms * 1e-3), nbad), flush=True)
            except Exception as ex:  # noqa: BLE001
                print("%-16s v%-2d failed: %s" % (name, v, ex), flush=True)
        os.environ.pop("CCGP_VARIANT", None)
        results[name] = dict(B=B, variants=res)
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    json.dump(results, open(os.path.join(ROOT, "gpurun_out", "tune.json"), "w"), indent=1)
    print("fp64 peak TFLOP/s:", eng.measure_fp64_peak() / 1e12)


if __name__ == "__main__":
    main()
