#!/usr/bin/env python
"""north_star numerics contract at bench scale: fast kernel vs the bit-faithful kernel on the same tens of millions of
pairs of the 50 000 x 500 workload (the strict kernel is md5-pinned to the reference; the oracle itself would need
hours for this many pairs).  Reports the largest deviations per column and the number of nIter / r2_ExpG mismatches."""
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
ap.add_argument("--seed", type=int, default=11)
a = ap.parse_args()
GL, pos = gen_synth.synth_fast(a.n_sites, a.n_ind, a.seed)
gl, expg, maf = N.prepare_sites(GL)
del GL
with N.Engine(0) as eng:
    eng.set_sites(gl, expg, maf)
    eng.set_positions(np.diff(np.concatenate([[0], pos])).astype(np.float64), None)
    fast = eng.scan(N.ScanParams.make(max_kb_dist=0), 0, a.s1_hi)
    kern = eng.stats()["em_kernel"]
    strict = eng.scan(N.ScanParams.make(max_kb_dist=0, strict=1), 0, a.s1_hi)
out = {"pairs": int(len(fast)), "fast_kernel": kern, "n_iter_mismatches": int((fast["n_iter"] != strict["n_iter"]).sum()),
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
