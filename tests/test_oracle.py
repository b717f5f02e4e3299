"""Pins the CPU oracle (oracle/ngsld_oracle.c) to the reference: md5 equality with outputs of the
unmodified reference binary on every golden fixture, GSL's own known answer for the taus generator,
and -- when oracle/_ref/ngsLD is present -- a live run of the reference itself."""
import ctypes as C
import os

import numpy as np
import pytest

import helpers as H
from helpers import O

CASES = [(fx, v) for fx, d in H.MANIFEST["fixtures"].items() for v in d["variants"]]


def test_taus_known_answer():
    # GSL rng/test.c: rng_test(gsl_rng_taus, 1, 10000, 2733957125UL)
    t = O.Taus()
    O.lib().orc_taus_set(C.byref(t), 1)
    for _ in range(9999):
        O.lib().orc_taus_get(C.byref(t))
    assert O.lib().orc_taus_get(C.byref(t)) == 2733957125


def test_taus_seed_zero_is_one():
    a, b = O.Taus(), O.Taus()
    O.lib().orc_taus_set(C.byref(a), 0)
    O.lib().orc_taus_set(C.byref(b), 1)
    assert [O.lib().orc_taus_get(C.byref(a)) for _ in range(5)] == [O.lib().orc_taus_get(C.byref(b)) for _ in range(5)]


@pytest.mark.parametrize("fx,variant", CASES)
def test_oracle_matches_reference_golden(fx, variant, tmp_path_factory):
    tmp = tmp_path_factory.getbasetemp()
    v = H.MANIFEST["fixtures"][fx]["variants"][variant]
    got = H.oracle_tsv(fx, tmp, v["flags"], use_pos=v["pos"], geno=v.get("geno"), n_threads=os.cpu_count() or 4)
    gold = H.golden_bytes(fx, variant)
    if gold is not None:
        assert H.md5(gold) == v["md5"]
        assert got == gold
    assert H.md5(got) == v["md5"]
    assert got.count(b"\n") - 1 == v["rows"]


@pytest.mark.skipif(not O.have_ref(), reason="oracle/_ref/ngsLD not built (needs /root/reference)")
@pytest.mark.parametrize("threads", [1, 3])
def test_reference_binary_live(threads, tmp_path):
    """The reference itself, run here: thread count must not change the sorted output, and the
    oracle restatement must equal it byte for byte on a fresh random fixture (not a stored one)."""
    GL, pos = H.gen_synth.synth(50, 17, 99)
    geno = str(tmp_path / "live.glf")
    H.gen_synth.write(geno, GL, pos)
    out = str(tmp_path / "ref.ld")
    O.run_ref(["--geno", geno, "--probs", "--n_ind", "17", "--n_sites", "50", "--pos", geno + ".pos",
               "--max_kb_dist", "7", "--extend_out"], out, n_threads=threads)
    ref = open(out, "rb").read()
    gl, expg, maf = O.preprocess(GL)
    labels, dist = O.read_pos(geno + ".pos")
    mine = str(tmp_path / "orc.ld")
    O.run(gl, expg, maf, dist, labels, max_kb_dist=7, out_path=mine, n_threads=2)
    assert sorted(ref.splitlines()) == sorted(open(mine, "rb").read().splitlines())
    if threads == 1:
        assert ref == open(mine, "rb").read()


def test_survey_digests():
    """Digests recorded independently in SURVEY.md App. D for the same generator + reference."""
    fx = H.MANIFEST["fixtures"]
    assert fx["p"]["variants"]["ext"]["md5"] == "f9c52f5e8dd422d542e9b55c0a5e2701"
    assert fx["q"]["variants"]["ext"]["md5"] == "e141c2ae9780434e2bbe2ec6ef5ce977"
    assert fx["edge"]["variants"]["ext"]["md5"] == "d6c15d75d16040b9c6e3cc0545e8f6d3"


def test_preprocess_rejects_nan():
    raw = np.full((2, 2, 3), 0.25)
    raw[1, 1, 0] = np.nan
    with pytest.raises(ValueError):
        O.preprocess(raw)


def test_pearson_matches_independent_long_double_restatement():
    """gsl_stats_correlation's recurrence (GSL statistics/covariance_source.c) restated a second time, in numpy's
    x87 long double, must agree bit for bit with the oracle's C version (and therefore with the GSL stand-in the
    reference binary was built against).  This does not prove what libgsl itself does -- it is not in this image --
    but it removes transcription errors from the one piece of third-party arithmetic on the path."""
    ld = np.longdouble
    assert np.finfo(ld).nmant == 63, "needs x87 80-bit long double"
    rng = np.random.default_rng(42)
    for n in (2, 3, 7, 24, 100, 501):
        for rep in range(6):
            x = rng.uniform(0, 2, n)
            y = np.clip(0.6 * x + rng.normal(0, 0.4, n), 0, 2)
            if rep == 0:
                x[:] = 0.5                       # zero variance -> 0/0
            if rep == 1:
                y = x.copy()                     # r = 1
            mean_x, mean_y = ld(x[0]), ld(y[0])
            sxx = syy = sxy = ld(0)
            for i in range(1, n):
                ratio = ld(i / (i + 1.0))
                dx, dy = ld(x[i]) - mean_x, ld(y[i]) - mean_y
                sxx += dx * dx * ratio
                syy += dy * dy * ratio
                sxy += dx * dy * ratio
                mean_x += dx / ld(i + 1.0)
                mean_y += dy / ld(i + 1.0)
            with np.errstate(all="ignore"):
                r = np.float64(sxy / ld(np.sqrt(np.float64(sxx)) * np.sqrt(np.float64(syy))))
                want = r * r
            got = O.lib().orc_pearson_r2(np.ascontiguousarray(x), np.ascontiguousarray(y), n)
            assert (np.isnan(want) and np.isnan(got)) or np.float64(got).tobytes() == np.float64(want).tobytes(), (n, rep)


def _random_flags(rng):
    flags = ["--probs"]
    flags += ["--max_kb_dist", str(int(rng.choice([0, 0, 3, 8, 20])))]
    if rng.random() < 0.4:
        flags += ["--max_snp_dist", str(int(rng.integers(1, 12)))]
    if rng.random() < 0.4:
        flags += ["--min_maf", f"{rng.uniform(0.05, 0.35):.3f}"]
    if rng.random() < 0.4:
        flags += ["--rnd_sample", f"{rng.uniform(0.05, 0.9):.3f}", "--seed", str(int(rng.integers(1, 100000)))]
    if rng.random() < 0.3:
        flags += ["--ignore_miss_data"]
    if rng.random() < 0.3:
        t = rng.uniform(0.5, 0.99)
        flags += ["--call_geno", "--N_thresh", f"{rng.uniform(0.0, t):.3f}", "--call_thresh", f"{t:.3f}"]
    if rng.random() < 0.7:
        flags += ["--extend_out"]
    return flags


@pytest.mark.skipif(not O.have_ref(), reason="oracle/_ref/ngsLD not built (needs /root/reference)")
@pytest.mark.parametrize("seed", range(14))
def test_oracle_fuzz_against_live_reference(seed, tmp_path):
    """Random fixtures x random flag combinations: the oracle restatement must reproduce the unmodified reference
    binary byte for byte (--n_threads 1 order)."""
    rng = np.random.default_rng(1000 + seed)
    n_sites, n_ind = int(rng.integers(20, 70)), int(rng.integers(2, 30))
    GL, pos = H.gen_synth.synth(n_sites, n_ind, 2000 + seed)
    if seed % 3 == 0:
        GL[rng.integers(0, n_sites), :] = [1 / 3] * 3          # a site without information
        GL[rng.integers(0, n_sites), rng.integers(0, n_ind)] = [0.0, 0.0, 0.0]
    if seed % 4 == 1:
        GL[rng.integers(0, n_sites), :] = [1.0, 0.0, 0.0]      # monomorphic
    geno = str(tmp_path / "f.glf")
    H.gen_synth.write(geno, GL, pos)
    if seed % 5 == 2:                                         # two chromosomes
        lines = open(geno + ".pos").read().splitlines()
        half = len(lines) // 2
        lines = lines[:half] + [l.replace("chr1", "chr2") for l in lines[half:]]
        open(geno + ".pos", "w").write("\n".join(lines) + "\n")
    flags = _random_flags(rng)
    ref_out = str(tmp_path / "ref.ld")
    O.run_ref(["--geno", geno, "--n_ind", str(n_ind), "--n_sites", str(n_sites), "--pos", geno + ".pos"] + flags, ref_out)
    opt = H.parse_flags(flags)
    raw = np.fromfile(geno, "<f8").reshape(n_sites, n_ind, 3)
    gl, expg, maf = O.preprocess(raw, opt["log_scale"], opt["ignore_miss"], opt["call_geno"], opt["n_thresh"], opt["call_thresh"])
    labels, dist = O.read_pos(geno + ".pos")
    mine = str(tmp_path / "orc.ld")
    O.run(gl, expg, maf, dist, labels, opt["max_kb_dist"], opt["max_snp_dist"], opt["min_maf"], opt["rnd_sample"],
          opt["seed"], opt["ignore_miss"], opt["extend_out"], out_path=mine, n_threads=2)
    assert open(mine, "rb").read() == open(ref_out, "rb").read(), flags
    # and the product's host preparation agrees with the oracle's on the same input
    import ngsld_b200 as N
    a = N.prepare_sites(raw, log_scale=opt["log_scale"], ignore_miss_data=opt["ignore_miss"], call_geno=opt["call_geno"],
                        N_thresh=opt["n_thresh"], call_thresh=opt["call_thresh"])
    for x, y in zip(a, (gl, expg, maf)):
        assert x.tobytes() == y.tobytes()
    P = N.ScanParams.make(max_kb_dist=opt["max_kb_dist"], max_snp_dist=opt["max_snp_dist"], min_maf=opt["min_maf"],
                          rnd_sample=opt["rnd_sample"], seed=opt["seed"])
    assert N.plan_count(maf, dist, P) == open(ref_out, "rb").read().count(b"\n") - 1


@pytest.mark.skipif(not os.environ.get("NGSLD_SLOW"), reason="~90 s on 8 cores: set NGSLD_SLOW=1")
def test_oracle_reproduces_survey_t2k_digest(tmp_path):
    """SURVEY.md App. D fixture t2k (2000 x 100, seed 1, 1 999 000 rows, --extend_out): the digest of the unmodified
    reference's --n_threads 1 output was recorded there independently of this repository."""
    GL, pos = H.gen_synth.synth(2000, 100, 1)
    geno = str(tmp_path / "t2k.glf")
    H.gen_synth.write(geno, GL, pos)
    assert H.md5(open(geno, "rb").read()) == "c0a38c1d2f02f7d201d24f1dd957d079"
    assert H.md5(open(geno + ".pos", "rb").read()) == "2fc7fcad38b1917a54ec61cd2bbc5cf6"
    gl, expg, maf = O.preprocess(GL)
    labels, dist = O.read_pos(geno + ".pos")
    out = str(tmp_path / "t2k.ld")
    n, _ = O.run(gl, expg, maf, dist, labels, 0, 0, 0.0, 1.0, 1, False, True, n_threads=os.cpu_count() or 4, out_path=out)
    assert n == 1999000
    assert H.md5(open(out, "rb").read()) == "740e0c65315d7e4d9d2cffbfba956dc4"
