#!/usr/bin/env python
"""Synthetic binary genotype-likelihood file of BASELINE-config size, written in parallel.

  python scripts/gen_big.py N_SITES N_IND SEED OUT [PROCS]

Same model as tests/golden/gen_synth.py (haplotype copying with LD, Poisson(2) depth, 1 % error), in
independent blocks of 4096 sites (one numpy stream per block) so that the blocks can be produced by a pool of
processes straight into the memory-mapped output file; `OUT.pos` gets "chr1\\t<pos>" with gaps of 1..999 bp."""
import multiprocessing as mp
import sys

import numpy as np

BLOCK = 4096


def block(args):
    out, n_sites, n_ind, seed, b = args
    s0 = b * BLOCK
    n = min(BLOCK, n_sites - s0)
    rng = np.random.default_rng([seed, b])
    nh = 2 * n_ind
    f = rng.uniform(0.05, 0.5, n)
    fresh = rng.random((n, nh)) < f[:, None]
    keep = rng.random((n, nh)) < 0.9
    keep[0] = False
    idx = np.where(~keep, np.arange(n)[:, None], 0)
    idx = np.maximum.accumulate(idx, axis=0)
    H = np.take_along_axis(fresh, idx, axis=0)
    G = H[:, 0::2].astype(np.int8) + H[:, 1::2]
    depth = rng.poisson(2.0, G.shape)
    alt = rng.binomial(depth, np.where(G == 0, 0.01, np.where(G == 1, 0.5, 0.99)))
    ref = depth - alt
    pa = np.array([0.01, 0.5, 0.99])
    L = np.exp(alt[..., None] * np.log(pa) + ref[..., None] * np.log(1 - pa))
    mm = np.memmap(out, dtype="<f8", mode="r+", shape=(n_sites, n_ind, 3))
    mm[s0:s0 + n] = L / L.sum(-1, keepdims=True)
    mm.flush()
    return b


def main():
    n_sites, n_ind, seed, out = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3]), sys.argv[4]
    procs = int(sys.argv[5]) if len(sys.argv) > 5 else max(1, mp.cpu_count() - 2)
    mm = np.memmap(out, dtype="<f8", mode="w+", shape=(n_sites, n_ind, 3))
    del mm
    jobs = [(out, n_sites, n_ind, seed, b) for b in range((n_sites + BLOCK - 1) // BLOCK)]
    with mp.Pool(procs) as pool:
        for _ in pool.imap_unordered(block, jobs):
            pass
    pos = np.cumsum(np.random.default_rng([seed, 1 << 30]).integers(1, 1000, n_sites))
    with open(out + ".pos", "w") as fh:
        fh.write("".join(f"chr1\t{p}\n" for p in pos))


if __name__ == "__main__":
    main()
