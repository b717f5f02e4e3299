#!/usr/bin/env python
"""A/B runs of the library's experiment knobs on one resident data set: for every configuration (a set of
NGSLD_* environment variables, read by the library at each scan) scan the same slab of first sites with results left in
HBM and print pairs/s and the kernel times.

  python scripts/sweep.py --n-sites 50000 --n-ind 500 --pairs 20000000 \
      --cfg default: --cfg dense:NGSLD_EM_PATH=warp --cfg unfused:NGSLD_CELL_FUSE=0,NGSLD_PEARSON_CTAS=2"""
import argparse
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests", "golden"))
import gen_synth  # noqa: E402
import ngsld_b200 as N  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--n-sites", type=int, default=50000)
ap.add_argument("--n-ind", type=int, default=500)
ap.add_argument("--seed", type=int, default=11)
ap.add_argument("--pairs", type=int, default=20_000_000, help="pairs per measured scan (a slab of first sites)")
ap.add_argument("--max-kb-dist", type=int, default=0)
ap.add_argument("--rnd-sample", type=float, default=1.0)
ap.add_argument("--call-geno", action="store_true")
ap.add_argument("--reps", type=int, default=2)
ap.add_argument("--cfg", action="append", default=[], help="label:VAR=value,VAR=value")
a = ap.parse_args()

t0 = time.time()
GL, pos = gen_synth.synth_fast(a.n_sites, a.n_ind, a.seed)
gl, expg, maf = N.prepare_sites(GL, call_geno=a.call_geno)
del GL
eng = N.Engine(0)
eng.set_sites(gl, expg, maf)
eng.set_positions(np.diff(np.concatenate([[0], pos])).astype(np.float64), None)
P = N.ScanParams.make(max_kb_dist=a.max_kb_dist, rnd_sample=a.rnd_sample, seed=1)
# slab of first sites [0, hi) holding about --pairs rows
lo, hi = 1, a.n_sites
while lo < hi:
    mid = (lo + hi) // 2
    if eng.count(P, 0, mid) < a.pairs:
        lo = mid + 1
    else:
        hi = mid
hi = lo
print(json.dumps({"setup_s": round(time.time() - t0, 1), "n_sites": a.n_sites, "n_ind": a.n_ind, "s1_hi": hi,
                  "rows": eng.count(P, 0, hi)}), flush=True)
for spec in a.cfg or ["default:"]:
    label, _, kv = spec.partition(":")
    env = dict(x.split("=", 1) for x in kv.split(",") if x)
    for k, v in env.items():
        os.environ[k] = v
    best = None
    try:
        for _ in range(a.reps):
            st = eng.scan_device(P, 0, hi)
            if best is None or st["ms_device_total"] < best["ms_device_total"]:
                best = st
    except N.NgsldError as e:
        print(json.dumps({"label": label, "error": str(e)}), flush=True)
        best = None
    for k in env:
        del os.environ[k]
    if best:
        st = best
        print(json.dumps({"label": label, "env": env, "kernel": st["em_kernel"],
                          "Mpairs_per_s": round(st["n_pairs"] / st["ms_device_total"] / 1e3, 2),
                          "ms_total": round(st["ms_device_total"], 1), "ms_em": round(st["ms_em"], 1),
                          "ms_pearson": round(st["ms_pearson"], 1),
                          "passes_per_pair": round(st["sum_em_passes"] / st["n_pairs"], 2),
                          "cells_per_pair": round(st["sum_cells"] / max(1, st["n_cell_pairs"]), 1),
                          "resid_pairs": st["n_resid_pairs"]}), flush=True)
eng.close()
