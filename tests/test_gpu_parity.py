"""GPU parity tests proper (-m gpu): the CUDA path, called through the C ABI, against the oracle and the
golden reference outputs.  Bars: strict kernel bit-exact in every column and byte-identical TSV; fast
kernel r2_ExpG bit-exact, D/D'/r2 within 1e-9 at equal nIter (BASELINE.json north_star)."""
import os

import numpy as np
import pytest

import helpers as H
import ngsld_b200 as N
from helpers import O

pytestmark = pytest.mark.gpu

SMALL = [(fx, v) for fx in ("edge", "tiny") for v in H.MANIFEST["fixtures"][fx]["variants"]]
MID = [("s", "ext"), ("s", "rnd01"), ("s", "kb20")]
# the headline sample sizes: p = 600 x 100 (group kernels), q = 300 x 500 (warp-per-pair kernel, one warp per pair),
# u = 300 x 1000 banded (two warps per pair), v = 300 x 2000 randomly sampled (four warps per pair)
BIG = [("p", "ext"), ("q", "ext"), ("u", "kb50"), ("v", "rnd30")]


@pytest.fixture(scope="module")
def G():
    import gpu_helpers
    return gpu_helpers


def _run_tsv(G, fx, variant, tmp, strict):
    v = H.MANIFEST["fixtures"][fx]["variants"][variant]
    raw, labels, dist, opt = H.load_fixture(fx, tmp, v["flags"], v["pos"], v.get("geno"))
    eng, _ = G.engine_for(raw, opt, labels, dist)
    with eng:
        return eng.scan_tsv(G.scan_params(opt, strict)), v


@pytest.mark.parametrize("fx,variant", SMALL + MID + BIG)
def test_strict_tsv_is_byte_identical_to_reference(G, fx, variant, tmp_path_factory):
    got, v = _run_tsv(G, fx, variant, tmp_path_factory.getbasetemp(), strict=True)
    gold = H.golden_bytes(fx, variant)
    if gold is not None:
        assert got == gold
    assert H.md5(got) == v["md5"]


def _parse(tsv):
    lines = tsv.decode().splitlines()
    head = lines[0].split("\t")
    rows = [l.split("\t") for l in lines[1:]]
    return head, rows


@pytest.mark.parametrize("fx,variant", SMALL + MID + BIG)
def test_fast_tsv_matches_reference_within_contract(G, fx, variant, tmp_path_factory):
    tmp = tmp_path_factory.getbasetemp()
    got, v = _run_tsv(G, fx, variant, tmp, strict=False)
    gold = H.golden_bytes(fx, variant)
    if gold is None:
        gold = H.oracle_tsv(fx, tmp, v["flags"], v["pos"], v.get("geno"))
        assert H.md5(gold) == v["md5"]
    hg, rg = _parse(got)
    hr, rr = _parse(gold)
    assert hg == hr and len(rg) == len(rr)
    col = {n: i for i, n in enumerate(hr)}
    exact = ["site1", "site2", "dist", "r2_ExpG"] + [c for c in ("sample_size", "maf1", "maf2", "loglike", "nIter") if c in col]
    for a, b in zip(rg, rr):
        for c in exact:
            assert a[col[c]] == b[col[c]], (c, a, b)
        for c in hr:
            if c in exact:
                continue
            x, y = a[col[c]], b[col[c]]
            if x != y:  # text may differ in the last printed digit only
                assert abs(float(x) - float(y)) <= 1.000001e-6 * max(1.0, abs(float(y))), (c, x, y)


@pytest.mark.parametrize("fx,variant", [("tiny", "ext"), ("tiny", "nomiss"), ("edge", "ext"), ("s", "ext")])
def test_rows_strict_bit_exact_and_fast_within_tolerance(G, fx, variant, tmp_path_factory):
    tmp = tmp_path_factory.getbasetemp()
    v = H.MANIFEST["fixtures"][fx]["variants"][variant]
    raw, labels, dist, opt = H.load_fixture(fx, tmp, v["flags"], v["pos"], v.get("geno"))
    eng, arrays = G.engine_for(raw, opt, labels, dist)
    with eng:
        strict = eng.scan(G.scan_params(opt, True))
        fast = eng.scan(G.scan_params(opt, False))
    assert len(strict) == v["rows"] == len(fast)
    sel = np.arange(len(strict)) if len(strict) <= 3000 else np.random.default_rng(0).choice(len(strict), 3000, False)
    ref = G.oracle_rows(arrays, strict["s1"][sel], strict["s2"][sel], opt["ignore_miss"])
    G.assert_strict_equal(strict[sel], ref)
    G.assert_fast_close(fast[sel], ref)
    # fast vs strict over ALL rows (both on the GPU)
    G.assert_fast_close(fast, strict)
    assert np.array_equal(fast["s1"], strict["s1"]) and np.array_equal(fast["s2"], strict["s2"])
    assert np.array_equal(fast["dist"], strict["dist"])


# every (individuals-per-lane, lanes-per-group) kernel instantiation family, incl. ragged tails; for each sample size
# the default kernel family, the class-compressed kernel forced (NGSLD_EM_PATH=cell) and, from 160 individuals, the
# dense warp-per-pair kernel forced (NGSLD_EM_PATH=warp)
@pytest.mark.parametrize("n_ind", [1, 2, 3, 7, 24, 32, 33, 64, 65, 100, 128, 129, 159, 160, 250, 256, 257, 500, 513, 1000, 1025, 2047, 2500, 5001])
@pytest.mark.parametrize("ignore_miss", [False, True])
def test_every_group_shape_against_oracle(G, n_ind, ignore_miss, monkeypatch):
    n_sites = 14 if n_ind <= 256 else 8
    GL, _ = H.gen_synth.synth(n_sites, n_ind, 1000 + n_ind)
    opt = H.parse_flags(["--max_kb_dist", "0"] + (["--ignore_miss_data"] if ignore_miss else []))
    eng, arrays = G.engine_for(GL, opt)
    s1, s2 = np.triu_indices(n_sites, 1)
    ref = G.oracle_rows(arrays, s1, s2, ignore_miss)
    with eng:
        G.assert_strict_equal(eng.pairs(s1, s2, ignore_miss, strict=True), ref)
        for path in [None, "cell"] + (["warp"] if n_ind >= 160 else []):
            if path:
                monkeypatch.setenv("NGSLD_EM_PATH", path)
            G.assert_fast_close(eng.pairs(s1, s2, ignore_miss, strict=False), ref)
            p = G.scan_params(opt, False)
            rows = eng.scan(p)            # window path
            kernel = eng.stats()["em_kernel"]
            assert np.array_equal(rows["s1"], s1) and np.array_equal(rows["s2"], s2)
            G.assert_fast_close(rows, ref)
            if path:
                assert kernel.startswith("em" + path + "::"), kernel
            monkeypatch.delenv("NGSLD_EM_PATH", raising=False)


@pytest.mark.parametrize("path", ["list", "tile"])
def test_list_and_tile_paths_agree_bitwise(G, path, tmp_path_factory, monkeypatch):
    tmp = tmp_path_factory.getbasetemp()
    v = H.MANIFEST["fixtures"]["s"]["variants"]["kb20"]
    raw, labels, dist, opt = H.load_fixture("s", tmp, v["flags"], True)
    eng, _ = G.engine_for(raw, opt, labels, dist)
    with eng:
        monkeypatch.setenv("NGSLD_EM_PATH", path)
        a = eng.scan(G.scan_params(opt, False))
        monkeypatch.setenv("NGSLD_EM_PATH", "list")
        b = eng.scan(G.scan_params(opt, False))
    assert a.tobytes() == b.tobytes()


def test_small_chunks_and_ranges_concatenate_to_the_whole(G, tmp_path_factory):
    """Chunked streaming and first-site range partitioning (the multi-GPU sharding unit) must not change a byte."""
    tmp = tmp_path_factory.getbasetemp()
    v = H.MANIFEST["fixtures"]["s"]["variants"]["ext"]
    raw, labels, dist, opt = H.load_fixture("s", tmp, v["flags"], True)
    eng, _ = G.engine_for(raw, opt, labels, dist)
    p = G.scan_params(opt, False)
    with eng:
        whole = eng.scan_tsv(p, header=False)
        eng.set_chunk_rows(777)
        assert eng.scan_tsv(p, header=False) == whole
        eng.set_chunk_rows(0)
        bounds = eng.partition(p, 3)
        assert bounds[0] == 0 and bounds[-1] == eng.n_sites and np.all(np.diff(bounds.astype(np.int64)) >= 0)
        parts = [eng.scan_tsv(p, int(bounds[k]), int(bounds[k + 1]), header=False) for k in range(3)]
        counts = [eng.count(p, int(bounds[k]), int(bounds[k + 1])) for k in range(3)]
    assert b"".join(parts) == whole
    assert sum(counts) == v["rows"] and max(counts) - min(counts) < 2 * eng.n_sites


def test_sampling_reproduces_reference_pair_set(G, tmp_path_factory):
    tmp = tmp_path_factory.getbasetemp()
    v = H.MANIFEST["fixtures"]["s"]["variants"]["rnd01"]
    raw, labels, dist, opt = H.load_fixture("s", tmp, v["flags"], True)
    eng, arrays = G.engine_for(raw, opt, labels, dist)
    with eng:
        rows = eng.scan(G.scan_params(opt, False))
    assert len(rows) == v["rows"] == 789
    gold = H.oracle_tsv("s", tmp, v["flags"], True).decode().splitlines()[1:]
    want = [tuple(l.split("\t")[:2]) for l in gold]
    got = [(labels[a], labels[b]) for a, b in zip(rows["s1"], rows["s2"])]
    assert got == want


def test_errors_are_reported_not_thrown(G):
    GL, _ = H.gen_synth.synth(6, 5, 1)
    gl, expg, maf = N.prepare_sites(GL)
    with N.Engine(0) as eng:
        with pytest.raises(N.NgsldError):
            eng.scan(N.ScanParams.make(max_kb_dist=0))           # no sites yet
        bad = maf.copy()
        bad[2] = 1.5
        with pytest.raises(N.NgsldError) as ei:
            eng.set_sites(gl, expg, bad)                          # haplo_freq: "invalid allele frequencies"
        assert ei.value.code == -4
        eng.set_sites(gl, expg, maf)
        with pytest.raises(N.NgsldError):
            eng.scan(N.ScanParams.make(max_kb_dist=0, rnd_sample=0.0))
        with pytest.raises(N.NgsldError):
            eng.pairs([0], [99])
        assert len(eng.scan(N.ScanParams.make(max_kb_dist=5))) == 0   # no positions: every distance is inf


def test_empty_and_degenerate_scans(G):
    """No pairs at all: one site, a window that excludes every pair, a maf filter that drops every site."""
    GL, pos = H.gen_synth.synth(12, 9, 3)
    gl, expg, maf = N.prepare_sites(GL)
    dist = np.diff(np.concatenate([[0], pos])).astype(np.float64) + 5000.0   # every gap > 5 kb
    with N.Engine(0) as eng:
        eng.set_sites(gl[:1], expg[:1], maf[:1])
        assert len(eng.scan(N.ScanParams.make(max_kb_dist=0))) == 0
        assert eng.scan_tsv(N.ScanParams.make(max_kb_dist=0)) == N.tsv_header(False)
        eng.set_sites(gl, expg, maf)
        eng.set_positions(dist, None)
        assert len(eng.scan(N.ScanParams.make(max_kb_dist=1))) == 0
        assert len(eng.scan(N.ScanParams.make(max_kb_dist=0, min_maf=1.0))) == 0
        assert eng.count(N.ScanParams.make(max_kb_dist=0, max_snp_dist=1)) == 11
        two = eng.scan(N.ScanParams.make(max_kb_dist=0), 10, 12)                # the last two first sites: one pair
        assert len(two) == 1 and (two["s1"][0], two["s2"][0]) == (10, 11)
        assert len(eng.pairs([], [])) == 0


@pytest.mark.parametrize("fx,variant", [("edge", "ext"), ("edge", "nopos"), ("tiny", "ext"), ("tiny", "plain")])
def test_host_format_fallback_is_byte_identical(G, fx, variant, tmp_path_factory):
    """The chunk-level fallback the device formatter takes for values >= 1e9 (host snprintf), forced on in a
    subprocess: same bytes, including the -nan / inf / -0.000000 spellings of the edge fixture."""
    import subprocess
    import sys
    code = (
        "import sys; sys.path.insert(0, %r); sys.path.insert(0, %r)\n"
        "import helpers as H, gpu_helpers as G\n"
        "v = H.MANIFEST['fixtures'][%r]['variants'][%r]\n"
        "raw, labels, dist, opt = H.load_fixture(%r, %r, v['flags'], v['pos'], v.get('geno'))\n"
        "eng, _ = G.engine_for(raw, opt, labels, dist)\n"
        "sys.stdout.buffer.write(eng.scan_tsv(G.scan_params(opt, True)))\n"
    ) % (H.ROOT, H.HERE, fx, variant, fx, str(tmp_path_factory.getbasetemp()))
    out = subprocess.run([sys.executable, "-c", code], capture_output=True, env=dict(os.environ, NGSLD_FORCE_HOST_FORMAT="1"))
    assert out.returncode == 0, out.stderr.decode()
    assert out.stdout == H.golden_bytes(fx, variant)


@pytest.mark.parametrize("n_ind", [40, 200, 700])
def test_all_missing_site_under_ignore_miss_stops_at_once_like_the_reference(G, n_ind):
    """A site whose every individual is flat gets maf = 0/0 under --ignore_miss_data; its pairs have NaN frequencies,
    and the reference's `if (d > eps)` chain leaves eps at 0, so the EM stops in the pass that produced the NaN
    (nIter 0).  Every kernel family must report that, not 100."""
    GL, _ = H.gen_synth.synth(10, n_ind, 500 + n_ind)
    GL[3, :] = [1 / 3, 1 / 3, 1 / 3]
    GL[6, : n_ind // 2] = [1 / 3, 1 / 3, 1 / 3]          # half missing: still a valid site
    opt = H.parse_flags(["--max_kb_dist", "0", "--ignore_miss_data"])
    eng, arrays = G.engine_for(GL, opt)
    assert np.isnan(arrays[2][3])
    s1, s2 = np.triu_indices(10, 1)
    ref = G.oracle_rows(arrays, s1, s2, True)
    touched = (s1 == 3) | (s2 == 3)
    assert np.all(ref["n_iter"][touched] == 0) and np.all(np.isnan(ref["hap"][touched]))
    with eng:
        G.assert_strict_equal(eng.pairs(s1, s2, True, strict=True), ref)
        fast = eng.scan(G.scan_params(opt, False))
        G.assert_fast_close(fast, ref)
        assert eng.stats()["sum_em_passes"] == int(np.where(ref["n_iter"] < 100, ref["n_iter"] + 1, 100).sum())


def test_text_buffers_follow_the_row_slot_size(G, tmp_path_factory):
    """One context, text scans with growing row slots: plain -> --extend_out -> longer labels.  The text buffers must
    be re-sized with the slot (they used to be kept whenever the row count fitted)."""
    tmp = tmp_path_factory.getbasetemp()
    v = H.MANIFEST["fixtures"]["s"]["variants"]["ext"]
    raw, labels, dist, opt = H.load_fixture("s", tmp, v["flags"], True)
    eng, _ = G.engine_for(raw, opt, None, dist)            # "(null)" labels first
    with eng:
        opt_plain = dict(opt, extend_out=False)
        a = eng.scan_tsv(G.scan_params(opt_plain, True))
        b = eng.scan_tsv(G.scan_params(opt, True))
        eng.set_positions(dist, labels)
        c = eng.scan_tsv(G.scan_params(opt, True))
        long_labels = [l + ":" + "x" * 40 for l in labels]
        eng.set_positions(dist, long_labels)
        d = eng.scan_tsv(G.scan_params(opt, True))
    assert H.md5(c) == v["md5"]
    assert a.count(b"\n") == b.count(b"\n") == c.count(b"\n") == d.count(b"\n") == v["rows"] + 1
    assert b.replace(b"(null)", b"") != b and len(b) > len(a)
    short = c.decode().splitlines()[1:]
    for la, lb in zip(short[:200], d.decode().splitlines()[1:201]):
        fa, fb = la.split("\t"), lb.split("\t")
        assert fb[0] == fa[0] + ":" + "x" * 40 and fb[1] == fa[1] + ":" + "x" * 40 and fa[2:] == fb[2:]


def test_text_into_one_buffer_and_shared_site_table(G, tmp_path_factory):
    """ngsld_scan_tsv_into (device -> one caller buffer, page-locked or pageable) gives the bytes of ngsld_scan_tsv, and a
    context that received its site table from another context by device-to-device copy (ngsld_share_sites) scans to the
    same bytes as the one that was fed from the host."""
    tmp = tmp_path_factory.getbasetemp()
    v = H.MANIFEST["fixtures"]["s"]["variants"]["ext"]
    raw, labels, dist, opt = H.load_fixture("s", tmp, v["flags"], True)
    eng, _ = G.engine_for(raw, opt, labels, dist)
    with eng:
        for strict in (True, False):
            p = G.scan_params(opt, strict)
            whole = eng.scan_tsv(p, header=False)
            for pinned in (True, False):
                text, n_rows = eng.scan_tsv_into(p, pinned=pinned)
                assert n_rows == v["rows"] and text == whole
            eng.set_chunk_rows(1000)
            assert eng.scan_tsv_into(p)[0] == whole
            eng.set_chunk_rows(0)
        assert H.md5(N.tsv_header(True) + eng.scan_tsv_into(G.scan_params(opt, True))[0]) == v["md5"]
        n_dev = N.load_library().ngsld_device_count()
        for dev in sorted({0, n_dev - 1}):
            with N.Engine(dev) as other:
                other.share_sites_from(eng)
                for strict in (True, False):
                    p = G.scan_params(opt, strict)
                    assert other.scan_tsv(p) == eng.scan_tsv(p)
                q = G.scan_params(dict(opt, max_kb_dist=20, rnd_sample=0.5, seed=3), False)
                assert other.scan(q).tobytes() == eng.scan(q).tobytes()
