#!/bin/bash
# sweep of the balanced ME schedule's grid size (CTAs per SM), P = 1000 and the reference's 60-row sweep
mkdir -p gpurun_out
python - > gpurun_out/me_sweep.txt 2>&1 <<'PY'
import os, sys, subprocess
sys.argv = ["x"]
import numpy as np, torch
sys.path.insert(0, os.getcwd())
import ccgp_b200
from ccgp_b200 import workloads
dev = torch.device("cuda", 0)
eng = ccgp_b200.Engine(0)
stream = torch.cuda.current_stream(dev); eng.set_stream(stream.cuda_stream)
D_old, pool = workloads.me_pool()
C = pool.shape[0]
d_old = torch.from_numpy(np.asfortranarray(D_old).T.copy()).to(dev)
d_new = torch.from_numpy(np.stack([p.flatten(order="F") for p in pool])).to(dev)
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
for P in (1000, 60, 250, 4000):
    params = workloads.me_params(P)
    d_par = torch.from_numpy(np.asfortranarray(params).T.copy()).to(dev)
    out = torch.empty(C * P, dtype=torch.float64, device=dev)
    def run():
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        flush.zero_(); e0.record(stream)
        eng.me_schur_batch_dev(d_old, 14, 2, d_new, 7, C, d_par, P, out)
        e1.record(stream); torch.cuda.synchronize(dev)
        return e0.elapsed_time(e1)
    os.environ["CCGP_ME_BALANCED"] = "0"
    t = sorted(run() for _ in range(9)); print("P=%d chunked: min %.4f med %.4f ms  %.3f G dets/s" % (P, t[0], t[4], C * P / t[0] / 1e6)); ref = out.clone()
    os.environ["CCGP_ME_BALANCED"] = "1"
    for ctas in (5, 8, 10, 12, 13, 15, 16, 20, 24, 25, 30, 32, 40, 48, 64):
        os.environ["CCGP_ME_CTAS"] = str(ctas)
        out.zero_()
        t = sorted(run() for _ in range(9))
        print("P=%d balanced ctas/SM %2d: min %.4f med %.4f ms  %.3f G dets/s  same bits %s" % (P, ctas, t[0], t[4], C * P / t[0] / 1e6, bool(torch.equal(out.view(torch.int64), ref.view(torch.int64)))))
PY
cat gpurun_out/me_sweep.txt
