#!/usr/bin/env python
"""Synthetic binary genotype-likelihood fixtures (test + bench input generator).

Writes `<out>` = raw little-endian doubles, site-major [n_sites][n_ind][3], normal scale, each triple
normalised to sum 1 -- exactly the layout the reference's binary branch reads
(reference shared/read_data.cpp:28-31, size check ngsLD.cpp:55) -- and `<out>.pos` ("chr1\\t<pos>").

Model: haplotype-copying LD (each haplotype keeps its previous allele with
prob. 0.9), Poisson(2) read depth, 1 % base error.  numpy PCG64 streams, so digests are tied to numpy.
"""
import sys
import numpy as np


def synth(n_sites, n_ind, seed):
    rng = np.random.default_rng(seed)
    nh = 2 * n_ind
    f = rng.uniform(0.05, 0.5, n_sites)
    H = np.empty((n_sites, nh), np.int8)
    H[0] = rng.random(nh) < f[0]
    for s in range(1, n_sites):
        fresh = rng.random(nh) < f[s]
        keep = rng.random(nh) < 0.9
        H[s] = np.where(keep, H[s - 1], fresh)
    G = H[:, 0::2] + H[:, 1::2]
    depth = rng.poisson(2.0, G.shape)
    err = 0.01
    alt = rng.binomial(depth, np.where(G == 0, err, np.where(G == 1, 0.5, 1 - err)))
    ref = depth - alt
    pa = np.array([err, 0.5, 1 - err])
    GL = (pa[None, None, :] ** alt[..., None]) * ((1 - pa)[None, None, :] ** ref[..., None])
    GL = GL / GL.sum(-1, keepdims=True)
    pos = np.cumsum(rng.integers(1, 1000, n_sites))
    return GL.astype('<f8'), pos


def synth_fast(n_sites, n_ind, seed, block=4096):
    """Same statistical model, vectorised over blocks of sites for bench-sized inputs
    (LD is broken every `block` sites; different random stream from synth())."""
    rng = np.random.default_rng(seed)
    nh = 2 * n_ind
    out = np.empty((n_sites, n_ind, 3), '<f8')
    pa = np.array([0.01, 0.5, 0.99])
    lpa, l1pa = np.log(pa), np.log(1 - pa)
    prev = None
    for s0 in range(0, n_sites, block):
        b = min(block, n_sites - s0)
        f = rng.uniform(0.05, 0.5, b)
        fresh = rng.random((b, nh)) < f[:, None]
        keep = rng.random((b, nh)) < 0.9
        keep[0] = False if prev is None else keep[0]
        # index of the last "fresh" site at or before s for each haplotype
        idx = np.where(~keep, np.arange(b)[:, None], 0)
        if prev is not None:
            idx = np.where(keep, -1, np.arange(b)[:, None])
        idx = np.maximum.accumulate(idx, axis=0)
        H = np.where(idx >= 0, np.take_along_axis(fresh, np.maximum(idx, 0), axis=0),
                     prev[None, :] if prev is not None else False)
        prev = H[-1].copy()
        G = H[:, 0::2].astype(np.int8) + H[:, 1::2]
        depth = rng.poisson(2.0, G.shape)
        alt = rng.binomial(depth, np.where(G == 0, 0.01, np.where(G == 1, 0.5, 0.99)))
        ref = depth - alt
        L = np.exp(alt[..., None] * lpa + ref[..., None] * l1pa)
        out[s0:s0 + b] = L / L.sum(-1, keepdims=True)
    pos = np.cumsum(rng.integers(1, 1000, n_sites))
    return out, pos


def write(out, GL, pos, chrom='chr1'):
    GL.astype('<f8').tofile(out)
    with open(out + '.pos', 'w') as fh:
        fh.write(''.join(f"{chrom}\t{p}\n" for p in pos))


if __name__ == '__main__':
    n_sites, n_ind, seed = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3])
    GL, pos = synth(n_sites, n_ind, seed)
    write(sys.argv[4], GL, pos)
