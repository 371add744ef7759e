"""BASELINE configs[3]: the ground-vibrations study ([G]:689-762) over ALL 17 training/test pairs the reference ships
(9 of size 50, 8 of size 90), sets sharded over the GPUs of the box: one process per GPU, set i goes to rank i mod N,
no data-path collective (the fits are independent; the per-set RMSPE lines are gathered by the parent).  Per set: Laplace
start -> C lock-step Metropolis chains (one batched logpost per step) -> predictive table of the pooled sample at the
150 / 110 test sites -> RMSPE of the posterior-mean prediction.
usage: python tools/fit_gv_sets.py [--gpus N] [--chains C] [--out file.jsonl]     (N = 0: every N in 1, 2, 4, 8 the box has)"""
import argparse
import json
import os
import subprocess
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tools"))


def sets():
    z = np.load(os.path.join(ROOT, "tests", "golden", "gv_sets.npz"))
    out = []
    for size, count in ((50, 9), (90, 8)):
        for i in range(1, count + 1):
            out.append(("size %d sample %d" % (size, i), z["train%d_%d" % (size, i)], z["test%d_%d" % (size, i)]))
    return out


def worker(rank, world, chains):
    import ccgp_b200
    import fit_gv
    eng = ccgp_b200.Engine(rank)
    S = sets()
    fit_gv.fit(eng, 2, N=200, samp_size=100, train=S[0][1], test=S[0][2])          # warm-up (library load, first launches)
    print(json.dumps({"rank": rank, "ready": True}), flush=True)
    sys.stdin.readline()                                                           # the parent starts all ranks together
    t0 = time.perf_counter()
    for i, (name, tr, te) in enumerate(S):
        if i % world != rank:
            continue
        t1 = time.perf_counter()
        r = fit_gv.fit(eng, chains, seed=100 + i, train=tr, test=te)
        rmspe = float(np.sqrt(np.mean((r["yhat"] - r["y_true"]) ** 2)))
        print(json.dumps({"rank": rank, "set": name, "n": int(tr.shape[0]), "rmspe": rmspe, "sd_y_true": float(np.std(r["y_true"])),
                          "seconds": time.perf_counter() - t1, "proposals": int(r["proposals"]), "samples": int(r["samples"]),
                          "launches": int(eng.launch_count)}), flush=True)
    print(json.dumps({"rank": rank, "done": True, "seconds": time.perf_counter() - t0}), flush=True)
    eng.close()


def run(world, chains):
    procs = [subprocess.Popen([sys.executable, os.path.abspath(__file__), "--worker", str(r), "--world", str(world), "--chains", str(chains)],
                              stdin=subprocess.PIPE, stdout=subprocess.PIPE, text=True) for r in range(world)]
    for p in procs:
        assert json.loads(p.stdout.readline())["ready"]
    t0 = time.perf_counter()
    for p in procs:
        p.stdin.write("go\n"); p.stdin.flush()
    lines, rank_s = [], []
    for p in procs:
        for ln in p.stdout:
            d = json.loads(ln)
            if d.get("done"):
                rank_s.append(d["seconds"])
            else:
                lines.append(d)
        p.wait()
    wall = time.perf_counter() - t0
    return {"gpus": world, "chains": chains, "sets": len(lines), "wall_s": wall, "max_rank_s": max(rank_s),
            "proposals_per_s": sum(d["proposals"] for d in lines) / max(rank_s), "mean_rmspe_50": float(np.mean([d["rmspe"] for d in lines if d["n"] == 50])),
            "mean_rmspe_90": float(np.mean([d["rmspe"] for d in lines if d["n"] == 90])), "per_set": sorted(lines, key=lambda d: d["set"])}


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=0)
    ap.add_argument("--chains", type=int, default=32)
    ap.add_argument("--out", default=os.path.join(ROOT, "gpurun_out", "fit_gv_sets.jsonl"))
    ap.add_argument("--worker", type=int, default=-1)
    ap.add_argument("--world", type=int, default=1)
    a = ap.parse_args()
    if a.worker >= 0:
        worker(a.worker, a.world, a.chains)
        sys.exit(0)
    import torch
    have = torch.cuda.device_count()
    res, base = [], None
    for G in ([a.gpus] if a.gpus > 0 else [g for g in (1, 2, 4, 8) if g <= have]):
        r = run(G, a.chains)
        base = base or r
        r["speedup_vs_first"] = base["max_rank_s"] / r["max_rank_s"]
        r["efficiency"] = r["speedup_vs_first"] * base["gpus"] / G
        res.append(r)
        print("GPUs %d: 17 sets in %.2f s (max over ranks; wall %.2f), %.0f logpost proposals/s, mean RMSPE n=50 %.3f, n=90 %.3f, speed-up %.2f (efficiency %.2f)" % (
            G, r["max_rank_s"], r["wall_s"], r["proposals_per_s"], r["mean_rmspe_50"], r["mean_rmspe_90"], r["speedup_vs_first"], r["efficiency"]), flush=True)
    os.makedirs(os.path.dirname(a.out), exist_ok=True)
    with open(a.out, "w") as f:
        for r in res:
            f.write(json.dumps(r) + "\n")
