"""Kernel-only timing of the M1 NLL batch for several batch sizes and kernel switches (CUDA events).
usage: python tools/time_nll.py [B ...]     env switches are toggled in-process (read at every launch)"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import ccgp_b200  # noqa: E402
from ccgp_b200 import workloads, GAUSS_ANISO_LAMBDA, LOGSCALE  # noqa: E402

Bs = [int(a) for a in sys.argv[1:]] or [1 << 16, 1 << 18, 1 << 20]
dev = torch.device("cuda", 0)
eng = ccgp_b200.Engine(0)
stream = torch.cuda.current_stream(dev)
eng.set_stream(stream.cuda_stream)
X, y, s2 = workloads.m1_design()
eng.set_design(X, y)
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
configs = [("auto", {}), ("team nw4", {"CCGP_KERNEL": "3", "CCGP_TEAM_NW": "4"}), ("team nw3", {"CCGP_KERNEL": "3", "CCGP_TEAM_NW": "3"}), ("team nw3 fused", {"CCGP_KERNEL": "3", "CCGP_TEAM_NW": "3", "CCGP_TEAM_FUSED": "1"}),
           ("team nw2", {"CCGP_KERNEL": "3", "CCGP_TEAM_NW": "2"}), ("warp", {"CCGP_KERNEL": "1"}),
           ("cta nw4", {"CCGP_KERNEL": "4"}), ("dfma", {"CCGP_NO_MMA": "1"}),
           ("team nw3 map1", {"CCGP_KERNEL": "3", "CCGP_TEAM_NW": "3", "CCGP_TEAM_MAP": "1"}),
           ("team nw2 map1", {"CCGP_KERNEL": "3", "CCGP_TEAM_NW": "2", "CCGP_TEAM_MAP": "1"}),
           ("team nw4 map1", {"CCGP_KERNEL": "3", "CCGP_TEAM_NW": "4", "CCGP_TEAM_MAP": "1"}),
           ("team nw4 map2", {"CCGP_KERNEL": "3", "CCGP_TEAM_NW": "4", "CCGP_TEAM_MAP": "2"})]
extra = os.environ.get("TIME_NLL_CONFIGS")
if extra:
    configs = [c for c in configs if c[0] in extra.split(",")]
ref = None
for B in Bs:
    th = workloads.m1_candidates(B)
    cand = torch.from_numpy(np.asfortranarray(th).T.copy()).to(dev)
    nll = torch.empty(B, dtype=torch.float64, device=dev)
    beta = torch.empty(B, dtype=torch.float64, device=dev)
    status = torch.empty(B, dtype=torch.int32, device=dev)
    for name, env in configs:
        for k in ("CCGP_MMA_NW", "CCGP_NO_MMA", "CCGP_TEAM_NW", "CCGP_TEAM_FUSED", "CCGP_KERNEL", "CCGP_TEAM_MAP"):
            os.environ.pop(k, None)
        os.environ.update(env)
        ts = []
        for it in range(5):
            flush.zero_()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(stream)
            eng.nll_batch_dev(cand, GAUSS_ANISO_LAMBDA, s2, scale=LOGSCALE, out_nll=nll, out_beta=beta, out_status=status)
            e1.record(stream)
            torch.cuda.synchronize(dev)
            ts.append(e0.elapsed_time(e1))
        ms = min(ts[1:])
        v = nll.cpu().numpy()
        if name == "dfma":
            ref = v.copy()
        print("B=%8d %-18s %8.3f ms  %6.2f M evals/s  cfg %s  nll[0]=%.15g" % (B, name, ms, B / ms / 1e3, eng.last_nll_config(), v[0]))
    sys.stdout.flush()
