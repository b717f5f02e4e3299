#!/usr/bin/env python
"""Run one BASELINE.json-style configuration (possibly scaled down) on GPU 0 and print what happened.

  python scripts/run_config.py --n-sites 200000 --n-ind 1000 --max-kb-dist 500           # config 4 (banded)
  python scripts/run_config.py --n-sites 100000 --n-ind 2000 --rnd-sample 0.01 --seed 1  # config 5, 1/10 of the sites
  --mode device|rows|tsv : leave results in HBM / binary rows to a host sink / TSV bytes to /dev/null"""
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
ap.add_argument("--n-sites", type=int, required=True)
ap.add_argument("--n-ind", type=int, required=True)
ap.add_argument("--max-kb-dist", type=int, default=0)
ap.add_argument("--max-snp-dist", type=int, default=0)
ap.add_argument("--rnd-sample", type=float, default=1.0)
ap.add_argument("--seed", type=int, default=1)
ap.add_argument("--data-seed", type=int, default=12)
ap.add_argument("--mode", default="device", choices=["device", "rows", "tsv"])
ap.add_argument("--s1-hi", type=int, default=0, help="scan only first sites [0, s1_hi)")
ap.add_argument("--strict", action="store_true")
a = ap.parse_args()

t0 = time.time()
GL, pos = gen_synth.synth_fast(a.n_sites, a.n_ind, a.data_seed)
t1 = time.time()
gl, expg, maf = N.prepare_sites(GL)
del GL
t2 = time.time()
dist = np.diff(np.concatenate([[0], pos])).astype(np.float64)
labels = [f"chr1:{p}" for p in pos] if a.mode == "tsv" else None
eng = N.Engine(0)
eng.set_sites(gl, expg, maf)
eng.set_positions(dist, labels)
t3 = time.time()
P = N.ScanParams.make(max_kb_dist=a.max_kb_dist, max_snp_dist=a.max_snp_dist, rnd_sample=a.rnd_sample, seed=a.seed,
                      strict=int(a.strict))
hi = a.s1_hi or a.n_sites
n_rows = eng.count(P, 0, hi)
t4 = time.time()
extra = {}
if a.mode == "device":
    st = eng.scan_device(P, 0, hi)
elif a.mode == "rows":
    acc = [0, 0]

    def sink(rows):
        acc[0] += len(rows)
        acc[1] += int(rows["n_iter"].sum())
    eng.scan_sink(P, sink, 0, hi)
    st = eng.stats()
    extra = {"rows_seen": acc[0], "sum_n_iter": acc[1]}
else:
    with open(os.devnull, "wb") as fh:
        class W:
            n = 0

            def write(self, b):
                W.n += len(b)
                fh.write(b)
        eng.scan_tsv(P, 0, hi, out=W())
    st = eng.stats()
    extra = {"tsv_bytes": W.n}
t5 = time.time()
print(json.dumps({"config": vars(a), "planned_rows": n_rows, "pairs": st["n_pairs"], "kernel": st["em_kernel"],
                  "mean_passes": st["sum_em_passes"] / max(1, st["n_pairs"]), "launches": st["n_launches"],
                  "s_synth": round(t1 - t0, 2), "s_prepare_host": round(t2 - t1, 2), "s_upload": round(t3 - t2, 2),
                  "s_count": round(t4 - t3, 2), "s_scan_wall": round(t5 - t4, 3), "ms_device": st["ms_device_total"],
                  "ms_em": st["ms_em"], "ms_pearson": st["ms_pearson"], "ms_format": st["ms_format"], "ms_plan": st["ms_plan"],
                  "pairs_per_s_wall": st["n_pairs"] / max(t5 - t4, 1e-9), **extra}))
eng.close()
