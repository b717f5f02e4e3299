"""Size-independent properties at bench-like sizes (-m gpu): what can be checked when the oracle would take hours.

  * fast vs strict kernel on a random sample of a large scan (both on the GPU; strict is md5-pinned to the reference)
  * partition shards concatenate to the whole scan, byte for byte (the multi-GPU sharding unit)
  * checksum of checksums: the nIter column sums to the kernel's own pass counter
  * invariants of the domain: haplotype frequencies sum to 1, |D'| <= 1, 0 <= r2 <= 1, swapping the two sites of a
    pair swaps hap01/hap10 and leaves D, r2, r2_ExpG unchanged
  * every EM kernel family gives the same answer on the same pairs (warp-per-pair with 1/2/4 warps, group kernels)"""
import os

import numpy as np
import pytest

import helpers as H
import ngsld_b200 as N

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def G():
    import gpu_helpers
    return gpu_helpers


@pytest.fixture(scope="module")
def big(G):
    GL, pos = H.gen_synth.synth_fast(3000, 500, 77)
    gl, expg, maf = N.prepare_sites(GL)
    eng = N.Engine(0)
    eng.set_sites(gl, expg, maf)
    eng.set_positions(np.diff(np.concatenate([[0], pos])).astype(np.float64), None)
    yield eng, (gl, expg, maf)
    eng.close()


@pytest.mark.parametrize("path", ["cell", "warp"])
def test_large_scan_fast_vs_strict_sample_and_pass_checksum(G, big, path, monkeypatch):
    eng, arrays = big
    monkeypatch.setenv("NGSLD_EM_PATH", path)
    P = N.ScanParams.make(max_kb_dist=0)
    rows = eng.scan(P)                                   # 4.5 M pairs
    st = eng.stats()
    assert len(rows) == 3000 * 2999 // 2 and st["em_kernel"].startswith("em" + path + "::")
    it = rows["n_iter"].astype(np.int64)
    assert int(np.where(it < 100, it + 1, 100).sum()) == st["sum_em_passes"]
    if path == "cell":   # 500 individuals at 2x depth: a pair has ~160 distinct (p, q) combinations, none is left over
        assert st["n_cell_pairs"] + st["n_resid_pairs"] == len(rows) and st["n_resid_pairs"] <= len(rows) // 100
        assert 100 < st["sum_cells"] / st["n_cell_pairs"] < 250
    sel = np.sort(np.random.default_rng(5).choice(len(rows), 20000, replace=False))
    strict = eng.pairs(rows["s1"][sel], rows["s2"][sel], strict=True)
    G.assert_fast_close(rows[sel], strict)
    # the oracle itself on a handful
    few = sel[:40]
    G.assert_strict_equal(strict[:40], G.oracle_rows(arrays, rows["s1"][few], rows["s2"][few]))
    # domain invariants over the whole scan
    hap = rows["hap"]
    ok = np.isfinite(rows["r2"])
    assert np.all(np.abs(hap.sum(1) - 1) < 1e-12)
    assert np.all(np.abs(rows["Dp"][ok]) <= 1 + 1e-9) and np.all((rows["r2"][ok] >= 0) & (rows["r2"][ok] <= 1 + 1e-9))
    assert np.all((rows["r2_expg"] >= 0) & (rows["r2_expg"] <= 1 + 1e-12))
    assert np.all(np.diff(rows["s1"].astype(np.int64)) >= 0)       # (s1, s2) order


def test_large_scan_shards_concatenate(big):
    eng, _ = big
    P = N.ScanParams.make(max_kb_dist=0, max_snp_dist=700)
    whole = eng.scan(P)
    b = eng.partition(P, 4)
    parts = [eng.scan(P, int(b[k]), int(b[k + 1])) for k in range(4)]
    assert sum(len(p) for p in parts) == len(whole)
    assert max(len(p) for p in parts) - min(len(p) for p in parts) <= 2 * 700
    assert np.concatenate(parts).tobytes() == whole.tobytes()


def test_site_swap_symmetry(big):
    eng, _ = big
    rng = np.random.default_rng(9)
    a = rng.integers(0, 3000, 5000).astype(np.uint32)
    b = rng.integers(0, 3000, 5000).astype(np.uint32)
    keep = a != b
    a, b = a[keep], b[keep]
    for strict in (True, False):
        x, y = eng.pairs(a, b, strict=strict), eng.pairs(b, a, strict=strict)
        assert np.array_equal(x["n_iter"], y["n_iter"])
        tol = 0 if strict else 1e-12
        # swapping the sites swaps the roles of hap01 and hap10; in strict mode the arithmetic is NOT symmetric in
        # the two sites (association order), so only the fast-kernel tolerance applies to both
        for f, g in (("D", "D"), ("r2", "r2")):
            fin = np.isfinite(x[f]) & np.isfinite(y[g])
            assert np.all(np.abs(x[f][fin] - y[g][fin]) <= 1e-9)
        assert np.all(np.abs(x["hap"][:, 1] - y["hap"][:, 2]) <= 1e-9) and np.all(np.abs(x["hap"][:, 0] - y["hap"][:, 0]) <= 1e-9)
        assert G_same(x["r2_expg"], y["r2_expg"], 1e-12)


def G_same(a, b, tol):
    fin = np.isfinite(a) & np.isfinite(b)
    return bool(np.all(np.abs(a[fin] - b[fin]) <= tol)) and bool(np.all(np.isnan(a) == np.isnan(b)))


W = {"NGSLD_EM_PATH": "warp"}
CELL = {"NGSLD_EM_PATH": "cell"}


@pytest.mark.parametrize("n_ind,env", [(500, dict(W, NGSLD_WARP_R="4")), (500, dict(W, NGSLD_WARP_G="2")), (500, dict(W, NGSLD_WARP_G="4")),
                                       (500, {"NGSLD_EM_PATH": "tile"}), (500, {"NGSLD_EM_PATH": "list"}),
                                       (1000, W), (1000, {"NGSLD_EM_PATH": "list"}), (250, dict(W, NGSLD_WARP_R="3")),
                                       (90, W),
                                       # class-compressed kernel: default shape, r2_ExpG in a kernel of its own, cells mostly
                                       # in the shared-memory tail, hardly any room (most pairs go to the dense kernel)
                                       (500, CELL), (500, dict(CELL, NGSLD_CELL_FUSE="0")),
                                       (500, dict(CELL, NGSLD_CELL_R="2", NGSLD_CELL_TCAP="160")),
                                       (500, dict(CELL, NGSLD_CELL_R="4", NGSLD_CELL_TCAP="32")),
                                       (1000, CELL), (2000, CELL), (2000, dict(CELL, NGSLD_CELL_TCAP="64")), (90, CELL), (250, {})])
def test_every_kernel_family_agrees_with_strict(G, n_ind, env, monkeypatch):
    GL, _ = H.gen_synth.synth_fast(160, n_ind, 1234 + n_ind)
    gl, expg, maf = N.prepare_sites(GL)
    with N.Engine(0) as eng:
        eng.set_sites(gl, expg, maf)
        strict = eng.scan(N.ScanParams.make(max_kb_dist=0, strict=1))
        for k, v in env.items():
            monkeypatch.setenv(k, v)
        for ign in (0, 1):
            fast = eng.scan(N.ScanParams.make(max_kb_dist=0, ignore_miss_data=ign))
            ref = strict if not ign else eng.scan(N.ScanParams.make(max_kb_dist=0, ignore_miss_data=1, strict=1))
            G.assert_fast_close(fast, ref)


def test_device_site_terms_equal_host_x87(monkeypatch):
    """The per-site x87 Pearson terms computed by the emulation on the device and by the host FPU's native long
    double give byte-identical rows."""
    GL, _ = H.gen_synth.synth(300, 77, 31)
    GL[5, :] = [1.0, 0.0, 0.0]          # monomorphic site: all deviations exactly zero
    GL[9, ::2] = [0.0, 0.0, 1.0]
    gl, expg, maf = N.prepare_sites(GL)
    P = N.ScanParams.make(max_kb_dist=0, strict=1)
    with N.Engine(0) as eng:
        eng.set_sites(gl, expg, maf)
        dev = eng.scan(P)
        monkeypatch.setenv("NGSLD_HOST_TERMS", "1")
        eng.set_sites(gl, expg, maf)
        host = eng.scan(P)
    assert dev.tobytes() == host.tobytes()


@pytest.mark.parametrize("bin_size,n_bins,kw", [(250.0, 400, dict(max_kb_dist=100)), (1000.0, 50, dict(max_kb_dist=0)),
                                                (333.5, 64, dict(max_kb_dist=0, rnd_sample=0.3, seed=5))])
def test_decay_bins_equal_numpy_binning_of_the_rows(bin_size, n_bins, kw):
    """ngsld_scan_decay (the binning of scripts/fit_LDdecay.R done on the device) against the same binning of the
    scanned rows in numpy: counts exact, sums to 1e-9 relative (floating-point atomics reorder the additions)."""
    GL, pos = H.gen_synth.synth(500, 60, 8)
    GL[7, :] = [1.0, 0.0, 0.0]                      # monomorphic site: NaN statistics must be left out per column
    gl, expg, maf = N.prepare_sites(GL)
    dist = np.diff(np.concatenate([[0], pos])).astype(np.float64)
    dist[300] = np.inf                              # a chromosome change: infinite distances are dropped
    P = N.ScanParams.make(**kw)
    with N.Engine(0) as eng:
        eng.set_sites(gl, expg, maf)
        eng.set_positions(dist, None)
        rows = eng.scan(P)
        bins, outside = eng.scan_decay(P, bin_size, n_bins)
    k = np.ceil(rows["dist"] / bin_size) - 1
    inside = np.isfinite(rows["dist"]) & (k < n_bins)
    assert outside == int((~inside).sum()) and inside.sum() > 1000
    k = k[inside].astype(np.int64)
    for j, f in enumerate(("r2_expg", "D", "Dp", "r2")):
        v = rows[f][inside]
        fin = np.isfinite(v)
        cnt = np.bincount(k[fin], minlength=n_bins)
        tot = np.bincount(k[fin], weights=v[fin], minlength=n_bins)
        assert np.array_equal(bins["n"][:, j], cnt), f
        assert np.allclose(bins["sum"][:, j], tot, rtol=1e-9, atol=1e-12), f


@pytest.mark.parametrize("flags,max_cells", [(["--call_geno"], 16), (["--call_geno", "--N_thresh", "0.4", "--call_thresh", "0.4"], 16)])
def test_called_genotypes_run_on_a_count_table(G, flags, max_cells):
    """--call_geno (reference gen_func.cpp:886-914) makes every likelihood triple one-hot, or flat for an individual
    without data (flat triples stay flat: `best = -1`) or below N_thresh: a pair then has at most 4 x 4 distinct
    combinations and the class-compressed EM iterates on that count table, whatever the sample size.  Same nIter,
    1e-9, as for any other input."""
    GL, _ = H.gen_synth.synth(120, 500, 4242)
    opt = H.parse_flags(["--max_kb_dist", "0"] + flags)
    eng, arrays = G.engine_for(GL, opt)
    with eng:
        fast = eng.scan(G.scan_params(opt, False))
        st = eng.stats()
        strict = eng.scan(G.scan_params(opt, True))
        ign = eng.scan(N.ScanParams.make(max_kb_dist=0, ignore_miss_data=1))
        ign_strict = eng.scan(N.ScanParams.make(max_kb_dist=0, ignore_miss_data=1, strict=1))
    assert st["em_kernel"].startswith("emcell::") and st["n_resid_pairs"] == 0
    assert st["sum_cells"] <= max_cells * st["n_cell_pairs"]
    G.assert_fast_close(fast, strict)
    G.assert_fast_close(ign, ign_strict)
    few = np.random.default_rng(2).choice(len(fast), 60, replace=False)
    G.assert_strict_equal(strict[few], G.oracle_rows(arrays, strict["s1"][few], strict["s2"][few]))


def test_uncompressible_data_keeps_the_dense_kernel():
    """Continuous likelihoods (every triple distinct): the palettes overflow, ngsld_set_sites sees that the class-
    compressed kernel cannot pay, and the dense warp-per-pair kernel runs (also when NGSLD_EM_PATH=cell asks for the
    other one: there is no palette to run it on)."""
    rng = np.random.default_rng(12)
    GL = rng.dirichlet([0.8, 0.8, 0.8], (40, 300))
    gl, expg, maf = N.prepare_sites(GL)
    with N.Engine(0) as eng:
        eng.set_sites(gl, expg, maf)
        P = N.ScanParams.make(max_kb_dist=0)
        a = eng.scan(P)
        assert eng.stats()["em_kernel"].startswith("emwarp::")
        strict = eng.scan(N.ScanParams.make(max_kb_dist=0, strict=1))
        # a few coded sites among uncoded ones: pairs touching an uncoded site are handed on to the dense kernel
        GL2 = GL.copy()
        GL2[::3] = H.gen_synth.synth(14, 300, 5)[0]
        gl2, expg2, maf2 = N.prepare_sites(GL2)
        eng.set_sites(gl2, expg2, maf2)
        os.environ["NGSLD_EM_PATH"] = "cell"
        try:
            b = eng.scan(P)
            st = eng.stats()
        finally:
            del os.environ["NGSLD_EM_PATH"]
        strict2 = eng.scan(N.ScanParams.make(max_kb_dist=0, strict=1))
    assert st["em_kernel"].startswith("emcell::") and st["n_resid_pairs"] > 0 and st["n_cell_pairs"] == 14 * 13 // 2
    import gpu_helpers
    gpu_helpers.assert_fast_close(b, strict2)
    gpu_helpers.assert_fast_close(a, strict)


@pytest.mark.parametrize("kw", [dict(), dict(ignore_miss_data=True), dict(call_geno=True, N_thresh=0.3, call_thresh=0.9),
                                dict(log_scale=True)])
def test_device_side_preparation_is_close_to_the_host_path(G, kw):
    """ngsld_set_sites_raw (K0, opt-in): likelihoods / expected genotypes / allele frequencies prepared by a kernel agree
    with the glibc host path to a few ulp, and a scan on them stays inside the 1e-9 contract (it is NOT bit-identical:
    CUDA's log/exp are not glibc's -- that is why the host path is the default)."""
    GL, _ = H.gen_synth.synth(200, 330, 77)
    GL[::9, ::4] = [1 / 3, 1 / 3, 1 / 3]
    GL[5, 7] = [0.0, 0.0, 0.0]                     # all-zero triple: becomes flat
    raw = np.log(GL) if kw.get("log_scale") else GL
    if kw.get("log_scale"):
        raw[5, 7] = [-3.0, -3.0, -3.0]
    gl, expg, maf = N.prepare_sites(raw, **kw)
    P = N.ScanParams.make(max_kb_dist=0, ignore_miss_data=int(kw.get("ignore_miss_data", False)))
    with N.Engine(0) as host, N.Engine(0) as dev:
        host.set_sites(gl, expg, maf)
        maf_d = dev.set_sites_raw(raw, **kw)
        a, b = host.scan(P), dev.scan(P)
    ok = np.isfinite(maf)
    assert np.array_equal(np.isnan(maf), np.isnan(maf_d))
    assert np.max(np.abs(maf_d[ok] - maf[ok]) / np.maximum(maf[ok], 1e-300)) < 1e-13
    assert np.array_equal(a["s1"], b["s1"]) and np.array_equal(a["n_used"], b["n_used"])
    assert np.mean(a["n_iter"] != b["n_iter"]) < 1e-3          # a last-bit input change may move a pair across the 1e-5 test
    same = a["n_iter"] == b["n_iter"]
    for f in ("D", "hap", "r2_expg"):
        x, y = a[f][same], b[f][same]
        fin = np.isfinite(x) & np.isfinite(y)
        assert np.all(np.isnan(x) == np.isnan(y)) and np.max(np.abs(x[fin] - y[fin])) < 1e-9, f
    with pytest.raises(N.NgsldError):
        bad = raw.copy()
        bad[3, 3, 1] = np.nan
        with N.Engine(0) as e:
            e.set_sites_raw(bad, **kw)
