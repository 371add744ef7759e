"""Kernel-only timing of the ME criterion (M2): 1000 All_Subdesigns designs x P parameter rows, device-resident
inputs, CUDA events on the context's stream.  usage: python tools/time_me.py [P] [reps]
(also the target of the ncu capture of me_schur_kernel: ncu --set full -k regex:me_schur -c 1 python tools/time_me.py 1000 2)"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import ccgp_b200  # noqa: E402
from ccgp_b200 import workloads  # noqa: E402

P = int(sys.argv[1]) if len(sys.argv) > 1 else 1000
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 10
dev = torch.device("cuda", 0)
eng = ccgp_b200.Engine(0)
stream = torch.cuda.current_stream(dev)
eng.set_stream(stream.cuda_stream)
D_old, pool = workloads.me_pool()
params = workloads.me_params(P)
C = pool.shape[0]
d_old = torch.from_numpy(np.asfortranarray(D_old).T.copy()).to(dev)          # column-major n_old x d
d_new = torch.from_numpy(np.stack([p.flatten(order="F") for p in pool])).to(dev)
d_par = torch.from_numpy(np.asfortranarray(params).T.copy()).to(dev)
out = torch.empty(C * P, dtype=torch.float64, device=dev)
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)


def run_once():
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    flush.zero_()
    e0.record(stream)
    eng.me_schur_batch_dev(d_old, 14, 2, d_new, 7, C, d_par, P, out)
    e1.record(stream)
    torch.cuda.synchronize(dev)
    return e0.elapsed_time(e1)


if len(sys.argv) > 3 and sys.argv[3] == "ab":
    # A/B of the chunked schedule against the balanced one (CCGP_ME_BALANCED, read at every launch): same bits, time
    os.environ["CCGP_ME_BALANCED"] = "0"
    os.environ["CCGP_ME_SYM"] = "0"
    t_old = min(run_once() for _ in range(reps + 1))
    ref = out.clone()
    print("chunked schedule (8 CTAs/SM grid): %.4f ms  %.3f G dets/s" % (t_old, C * P / t_old / 1e6))
    os.environ["CCGP_ME_BALANCED"] = "1"
    for ctas in (0, 3, 4, 5, 6, 8):
        os.environ["CCGP_ME_CTAS"] = str(ctas)
        out.zero_()
        t_new = min(run_once() for _ in range(reps + 1))
        same = bool(torch.equal(out.view(torch.int64), ref.view(torch.int64)))
        print("balanced schedule, CTAs/SM %s: %.4f ms  %.3f G dets/s  bit-identical to chunked: %s"
              % (ctas or "occupancy query", t_new, C * P / t_new / 1e6, same))
    os.environ.pop("CCGP_ME_CTAS")
    for bal in ("0", "1"):
        os.environ["CCGP_ME_BALANCED"] = bal
        os.environ["CCGP_ME_SYM"] = "1"
        out.zero_()
        t_new = min(run_once() for _ in range(reps + 1))
        same = bool(torch.equal(out.view(torch.int64), ref.view(torch.int64)))
        print("pairwise S block (CCGP_ME_SYM=1), balanced %s: %.4f ms  %.3f G dets/s  bit-identical to the first version: %s"
              % (bal, t_new, C * P / t_new / 1e6, same))
    sys.exit(0)

ts = []
for it in range(reps):
    flush.zero_()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    eng.me_schur_batch_dev(d_old, 14, 2, d_new, 7, C, d_par, P, out)
    e1.record(stream)
    torch.cuda.synchronize(dev)
    ts.append(e0.elapsed_time(e1))
ms = min(ts[1:]) if reps > 1 else ts[0]
flop = 4860.0 * C * P
print("ME %d x %d: %.4f ms  %.3f G dets/s  %.2f TFLOP/s algorithmic (4860 FLOP + 238 exp per det)" % (C, P, ms, C * P / ms / 1e6, flop / ms / 1e9))
v = out.cpu().numpy().reshape(P, C)
print("argmin row0", int(v[0].argmin()), "value", v[0].min())
