#!/usr/bin/env python
"""north_star numerics contract at scale: the default fast kernel vs the bit-faithful kernel (md5-pinned to the
reference; the oracle itself would need hours) on the same pairs of a BASELINE-sized input.  Reports the largest
deviations per column and the number of nIter / r2_ExpG mismatches.

  python scripts/contract_at_scale.py                                       # config 3: 50 000 x 500, first 800 first-sites, all pairs
  python scripts/contract_at_scale.py --geno /dev/shm/c5.glf --n-sites 1000000 --n-ind 2000 --head 60000 \
         --rnd-sample 0.01 --seed 1 --s1-hi 6000                            # config 5: sampled pairs of a 60 000-site prefix
  python scripts/contract_at_scale.py --geno /dev/shm/c4.glf --n-sites 200000 --n-ind 1000 --head 50000 --max-kb-dist 500 --s1-hi 300"""
import argparse
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests", "golden"))
import gen_synth  # noqa: E402
import ngsld_b200 as N  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--n-sites", type=int, default=50000)
ap.add_argument("--n-ind", type=int, default=500)
ap.add_argument("--s1-hi", type=int, default=800)
ap.add_argument("--seed", type=int, default=1, help="--seed of the scan (random sampling)")
ap.add_argument("--data-seed", type=int, default=11)
ap.add_argument("--geno", default="", help="binary GL file (default: generate the config-3 input)")
ap.add_argument("--head", type=int, default=0, help="use only the first HEAD sites of the file")
ap.add_argument("--max-kb-dist", type=int, default=0)
ap.add_argument("--rnd-sample", type=float, default=1.0)
a = ap.parse_args()
if a.geno:
    n = a.head or a.n_sites
    GL = np.fromfile(a.geno, "<f8", count=n * a.n_ind * 3).reshape(n, a.n_ind, 3)
    pos = np.array([int(l.split("\t")[1]) for l in open(a.geno + ".pos").read().splitlines()[:n]])
else:
    GL, pos = gen_synth.synth_fast(a.n_sites, a.n_ind, a.data_seed)
gl, expg, maf = N.prepare_sites(GL)
del GL
kw = dict(max_kb_dist=a.max_kb_dist, rnd_sample=a.rnd_sample, seed=a.seed)
with N.Engine(0) as eng:
    eng.set_sites(gl, expg, maf)
    eng.set_positions(np.diff(np.concatenate([[0], pos])).astype(np.float64), None)
    fast = eng.scan(N.ScanParams.make(**kw), 0, a.s1_hi)
    st = eng.stats()
    strict = eng.scan(N.ScanParams.make(strict=1, **kw), 0, a.s1_hi)
out = {"pairs": int(len(fast)), "n_ind": a.n_ind, "fast_kernel": st["em_kernel"],
       "cells_per_pair": st["sum_cells"] / max(1, st["n_cell_pairs"]), "pairs_left_to_dense_kernel": st["n_resid_pairs"],
       "same_pairs": bool(np.array_equal(fast["s1"], strict["s1"]) and np.array_equal(fast["s2"], strict["s2"])),
       "n_iter_mismatches": int((fast["n_iter"] != strict["n_iter"]).sum()),
       "r2_expg_bit_mismatches": int((fast["r2_expg"].view(np.uint64) != strict["r2_expg"].view(np.uint64)).sum()),
       "not_converged_pairs": int((strict["n_iter"] == 100).sum()), "mean_n_iter": float(strict["n_iter"].mean())}
same = fast["n_iter"] == strict["n_iter"]
for f in ("D", "Dp", "r2", "hap", "hap_maf"):
    x, y = fast[f][same], strict[f][same]
    fin = np.isfinite(x) & np.isfinite(y)
    d = np.abs(x[fin] - y[fin])
    out["max_abs_diff_" + f] = float(d.max()) if d.size else 0.0
    out["nan_pattern_mismatches_" + f] = int((np.isnan(x) != np.isnan(y)).sum())
    if f in ("Dp", "r2"):
        out["pairs_over_1e-9_" + f] = int((d > 1e-9).sum())
print(json.dumps(out))
