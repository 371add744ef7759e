"""Throughput of the blocked large-n path (SURVEY 8d ME-B(1)): n = 2048, d = 2 anisotropic,
batch of 64 candidates.  Run on the B200 box."""
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import ccgp_b200  # noqa: E402
from ccgp_b200 import workloads, GAUSS_ANISO_LAMBDA  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 2048
B = int(sys.argv[2]) if len(sys.argv) > 2 else 64
X = workloads.synthetic_pool(n, seed=2048)
y = np.sin(3 * X[:, 0]) * np.cos(2 * X[:, 1])
h2 = 4.0 / n
rng = np.random.default_rng(1)
nat = np.column_stack([rng.uniform(0.2, 0.8, B), rng.uniform(1.2, 2.5, B) / h2, rng.uniform(1.2, 2.5, B) / h2, rng.uniform(0.5, 2, B)])
eng = ccgp_b200.Engine(0)
dev = torch.device("cuda", 0)
stream = torch.cuda.current_stream(dev)
eng.set_stream(stream.cuda_stream)
eng.set_design(X, y)
cand = torch.from_numpy(np.asfortranarray(nat).T.copy()).to(dev)
out = eng.nll_batch_dev(cand, GAUSS_ANISO_LAMBDA, 1.0)
torch.cuda.synchronize()
l0 = eng.launch_count
reps = 5
tms = []
for _ in range(reps):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    eng.nll_batch_dev(cand, GAUSS_ANISO_LAMBDA, 1.0, out_nll=out[0], out_beta=out[1], out_status=out[2])
    e1.record(stream)
    torch.cuda.synchronize()
    tms.append(e0.elapsed_time(e1))
print("per-rep ms (one call per event pair; includes the host's enqueue time):", " ".join("%.2f" % t for t in tms))
# nine batches queued back to back, as bench.py's large_n block does: the host runs ahead of the GPU
tq = []
for _ in range(3):
    evs = [torch.cuda.Event(enable_timing=True) for _ in range(10)]
    evs[0].record(stream)
    for i in range(9):
        eng.nll_batch_dev(cand, GAUSS_ANISO_LAMBDA, 1.0, out_nll=out[0], out_beta=out[1], out_status=out[2])
        evs[i + 1].record(stream)
    torch.cuda.synchronize()
    tq.append(evs[0].elapsed_time(evs[9]) / 9)
    print("   queued batches, ms each:", " ".join("%.2f" % evs[i].elapsed_time(evs[i + 1]) for i in range(9)))
print("queued x9, ms per batch:", " ".join("%.2f" % t for t in tq))
reps += 27
ms = min(tq)
flop = n ** 3 / 3 + n ** 2 / 2 + 2 * n * n + (n * (n - 1) / 2) * 11
print("n=%d B=%d: %.2f ms per batch, %.1f evals/s, %.2f TFLOP/s FP64 (algorithmic), %d launches/batch, bad=%d" % (
    n, B, ms, B / (ms * 1e-3), flop * B / (ms * 1e-3) / 1e12, (eng.launch_count - l0) // reps, int((out[2] != 0).sum())))
