"""Fixture from the only stored OUTPUT the reference ships: `Ground Vibrations Emulator/Results/Size 50 Results 1.txt`
(written by [G]:760-761 after an unseeded MCMC fit on `Training Set Size 50 Sample 1`).  Run in the build
container (needs /root/reference): python tests/golden/make_gv_results.py -> tests/golden/gv50_results1.npz"""
import os

import numpy as np

REF = "/root/reference/Ground Vibrations Emulator/Results/Size 50 Results 1.txt"
rows = [l.split() for l in open(REF).read().strip().splitlines()]
hdr = [h.strip('"') for h in rows[0]]
data = np.array([[float(v) for v in r[1:]] for r in rows[1:]])
col = {h: i for i, h in enumerate(hdr)}
out = dict(X_test=data[:, :9], y_true=data[:, col["y.true"]], y_hat_combined=data[:, col["y.hat.Combined"]],
           ll_combined=data[:, col["LL.Combined"]], ul_combined=data[:, col["UL.Combined"]],
           y_hat_single=data[:, col["y.hat.single"]], y_hat_cgp=data[:, col["y.hat.CGP"]])
np.savez_compressed(os.path.join(os.path.dirname(os.path.abspath(__file__)), "gv50_results1.npz"), **out)
rm = lambda a: float(np.sqrt(np.mean((a - out["y_true"]) ** 2)))
print("RMSPE combined %.4f single %.4f cgp %.4f; coverage combined %.3f" % (
    rm(out["y_hat_combined"]), rm(out["y_hat_single"]), rm(out["y_hat_cgp"]),
    float(np.mean((out["y_true"] >= out["ll_combined"]) & (out["y_true"] <= out["ul_combined"])))))
