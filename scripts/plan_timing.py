#!/usr/bin/env python
"""Host-side cost of starting a scan (plan, sampling counts, kernel choice, buffers) for sampled and windowed scans:
prints the library's own split (NGSLD_DEBUG_PLAN=1) for a few slabs."""
import os
import sys
import time

import numpy as np

os.environ["NGSLD_DEBUG_PLAN"] = "1"
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests", "golden"))
import gen_synth  # noqa: E402
import ngsld_b200 as N  # noqa: E402

n_sites, n_ind = int(sys.argv[1]) if len(sys.argv) > 1 else 120000, int(sys.argv[2]) if len(sys.argv) > 2 else 64
GL, pos = gen_synth.synth_fast(n_sites, n_ind, 5)
gl, expg, maf = N.prepare_sites(GL)
with N.Engine(0) as eng:
    eng.set_sites(gl, expg, maf)
    eng.set_positions(np.diff(np.concatenate([[0], pos])).astype(np.float64), None)
    for kw in (dict(max_kb_dist=0, rnd_sample=0.01, seed=1), dict(max_kb_dist=0, rnd_sample=0.5, seed=1), dict(max_kb_dist=100)):
        P = N.ScanParams.make(**kw)
        b = eng.partition(P, 16)
        for k in (0, 7, 15):
            t0 = time.perf_counter()
            st = eng.scan_device(P, int(b[k]), int(b[k + 1]))
            print(kw, "slab", k, "rows", st["n_pairs"], "wall %.1f ms" % (1e3 * (time.perf_counter() - t0)),
                  "plan %.1f ms device %.1f ms" % (st["ms_plan"], st["ms_device_total"]), flush=True)
