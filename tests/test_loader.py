"""Host side of the drop-in (CPU tests): the input readers and the device-free planner behind the C ABI.

The readers are checked end to end against the reference: file -> ngsld_load_geno / ngsld_load_positions ->
ngsld_prepare_sites -> (oracle scan, the checker) must reproduce, byte for byte, what the unmodified reference
binary printed for the same text inputs (tests/golden/make_golden_text.py)."""
import gzip
import os

import numpy as np
import pytest

import helpers as H
import ngsld_b200 as N
from helpers import O

TEXT = H.MANIFEST["text_cases"]


def load_case(name):
    c = TEXT[name]
    opt = H.parse_flags(c["flags"])
    probs = "--probs" in c["flags"] or "--log_scale" in c["flags"]
    cells, log_cells = N.load_geno(os.path.join(H.GOLD, c["geno"]), c["n_ind"], c["n_sites"], probs=probs,
                                   log_scale=opt["log_scale"])
    gl, expg, maf = N.prepare_sites(cells, log_scale=opt["log_scale"], from_log_cells=log_cells,
                                    ignore_miss_data=opt["ignore_miss"], call_geno=opt["call_geno"],
                                    N_thresh=opt["n_thresh"], call_thresh=opt["call_thresh"])
    labels, dist = N.read_positions(os.path.join(H.GOLD, c["pos"]), c["n_sites"], header=c["posH"])
    return c, opt, (gl, expg, maf), labels, dist


@pytest.mark.parametrize("name", sorted(TEXT))
def test_text_inputs_reproduce_reference_output(name, tmp_path):
    c, opt, (gl, expg, maf), labels, dist = load_case(name)
    out = str(tmp_path / "o.ld")
    O.run(gl, expg, maf, dist, labels, opt["max_kb_dist"], opt["max_snp_dist"], opt["min_maf"], opt["rnd_sample"],
          opt["seed"], opt["ignore_miss"], opt["extend_out"], n_threads=2, out_path=out)
    got = open(out, "rb").read()
    gold = gzip.open(os.path.join(H.GOLD, f"tiny.{name}.ld.gz"), "rb").read()
    assert H.md5(gold) == c["md5"]
    assert got == gold


def test_text_and_binary_inputs_load_identically():
    raw = np.fromfile(os.path.join(H.GOLD, "tiny.glf"), "<f8").reshape(40, 24, 3)
    cells, log_cells = N.load_geno(os.path.join(H.GOLD, "tiny.beagle.gz"), 24, 40)
    assert log_cells and np.array_equal(cells, np.log(raw))
    b, lc = N.load_geno(os.path.join(H.GOLD, "tiny.glf"), 24, 40)
    assert not lc and b.tobytes() == raw.tobytes()
    g, lc = N.load_geno(os.path.join(H.GOLD, "tiny.geno.gz"), 24, 40, probs=False)
    assert lc and set(np.unique(g)) <= {0.0, -1e15, np.log(1 / 3)}


def test_positions_rules(tmp_path):
    p = tmp_path / "a.pos"
    p.write_text("#c\nchrA\t10\textra\tcol\n\nchrA\t25\nchrB\t7\nchrB\t9\n")
    labels, dist = N.read_positions(str(p), 4)
    assert labels == ["chrA:10\textra\tcol", "chrA:25", "chrB:7", "chrB:9"]       # only the first tab becomes ':'
    assert dist.tolist() == [10.0, 15.0, np.inf, 2.0]
    gz = tmp_path / "a.pos.gz"
    with gzip.open(gz, "wt") as fh:
        fh.write("chr\tpos\nchrA\t10\nchrA\t25\n")
    labels, dist = N.read_positions(str(gz), 2, header=True)
    assert labels == ["chrA:10", "chrA:25"] and dist.tolist() == [10.0, 15.0]
    for body, msg in (("chrA\t10\nchrA\t10\n", "invalid distance between adjacent sites!"),
                      ("chrA\t10\n", "wrong number of lines in POS file!"),
                      ("chrA 10\nchrA 12\n", "wrong POS file format!")):
        q = tmp_path / "bad.pos"
        q.write_text(body)
        with pytest.raises(N.NgsldError) as ei:
            N.read_positions(str(q), 2)
        assert msg in str(ei.value) and "[read_dist]" in str(ei.value)


def test_geno_reader_errors(tmp_path):
    raw = np.full((3, 2, 3), 1 / 3)
    f = tmp_path / "x.glf"
    raw.tofile(f)
    with pytest.raises(N.NgsldError) as ei:
        N.load_geno(str(f), 2, 4)
    assert "premature EOF" in str(ei.value)
    with pytest.raises(N.NgsldError) as ei:
        N.load_geno(str(f), 2, 2)
    assert "not at EOF" in str(ei.value)
    g = tmp_path / "g.gz"
    with gzip.open(g, "wt") as fh:
        fh.write("0 1 3\n")
    with pytest.raises(N.NgsldError) as ei:
        N.load_geno(str(g), 3, 1, probs=False)
    assert "{-1,0,1,2}" in str(ei.value)
    with gzip.open(g, "wt") as fh:
        fh.write("hdr a b\n0 1 2\n0 1\n")          # a short line after the first site is an error, not a header
    with pytest.raises(N.NgsldError) as ei:
        N.load_geno(str(g), 3, 2, probs=False)
    assert "Less fields than expected" in str(ei.value)
    with pytest.raises(N.NgsldError) as ei:
        N.load_geno(str(tmp_path / "missing.gz"), 3, 1)
    assert ei.value.code == -6


# ---- device-free planner --------------------------------------------------------------------------
PLANS = [dict(max_kb_dist=0), dict(max_kb_dist=3), dict(max_kb_dist=0, max_snp_dist=5), dict(max_kb_dist=0, min_maf=0.3),
         dict(max_kb_dist=2, max_snp_dist=7, min_maf=0.25), dict(max_kb_dist=0, rnd_sample=0.5, seed=12345),
         dict(max_kb_dist=4, rnd_sample=0.1, seed=7, min_maf=0.25)]


@pytest.mark.parametrize("kw", PLANS)
def test_plan_count_and_partition_match_oracle_scan(kw, tmp_path):
    GL, pos = H.gen_synth.synth(90, 12, 5)
    gl, expg, maf = N.prepare_sites(GL)
    dist = np.diff(np.concatenate([[0], pos])).astype(np.float64)
    dist[40] = np.inf                                                  # a chromosome change
    P = N.ScanParams.make(**kw)

    def oracle_count(lo, hi):
        n, _ = O.run(gl, expg, maf, dist, None, kw.get("max_kb_dist", 100), kw.get("max_snp_dist", 0),
                     kw.get("min_maf", 0.0), kw.get("rnd_sample", 1.0), kw.get("seed", 1), False, False, lo, hi,
                     out_path="/dev/null")
        return n
    total = oracle_count(0, 90)
    assert N.plan_count(maf, dist, P) == total
    assert N.plan_count(maf, dist, P, 13, 57) == oracle_count(13, 57)
    for parts in (1, 2, 3, 8):
        b = N.plan_partition(maf, dist, P, parts).astype(np.int64)
        assert b[0] == 0 and b[-1] == 90 and np.all(np.diff(b) >= 0)
        counts = [N.plan_count(maf, dist, P, int(b[k]), int(b[k + 1])) for k in range(parts)]
        assert sum(counts) == total
        assert max(counts) - min(counts) <= 2 * 90                      # balanced to within one first site's rows


def test_plan_without_positions_has_no_finite_distance():
    maf = np.full(10, 0.3)
    assert N.plan_count(maf, None, N.ScanParams.make(max_kb_dist=5)) == 0
    assert N.plan_count(maf, None, N.ScanParams.make(max_kb_dist=0)) == 45
    with pytest.raises(N.NgsldError):
        N.plan_count(maf, None, N.ScanParams.make(max_kb_dist=0, rnd_sample=0.0))
