"""Loader quirks checked against the UNMODIFIED reference binary run here (oracle/_ref/ngsLD; skipped when it did not
travel): oddly formatted text inputs go through the reference on one side and through ngsld_load_geno /
ngsld_load_positions / ngsld_prepare_sites + the oracle scan (the checker) on the other; outputs must be identical."""
import gzip
import os
import subprocess

import numpy as np
import pytest

import helpers as H
import ngsld_b200 as N
from helpers import O

pytestmark = pytest.mark.skipif(not O.have_ref(), reason="oracle/_ref/ngsLD not built (needs /root/reference)")

N_SITES, N_IND = 14, 5


def _gl(seed=3):
    rng = np.random.default_rng(seed)
    GL = rng.dirichlet([0.7] * 3, (N_SITES, N_IND))
    GL[2, 1] = [1 / 3] * 3
    GL[4, :] = [0.0, 0.0, 1.0]
    return GL


def _pos_lines():
    return [("chrA" if s < 9 else "chrB", 100 + 37 * s) for s in range(N_SITES)]


def _fmt_variants(v, k):
    """the same value written in different but equivalent spellings"""
    r = repr(float(v))
    return [r, "+" + r, f"{v:.17e}", r][k % 4] if v != 0 else ["0", "0.0", "0e0", "+0"][k % 4]


def write_case(tmp, name):
    GL = _gl()
    pos = _pos_lines()
    geno, posf = str(tmp / f"{name}.gz"), str(tmp / f"{name}.pos")
    flags, use_header = ["--probs"], False
    if name == "mixed_separators_and_spellings":
        rows = []
        for s in range(N_SITES):
            vals = [_fmt_variants(v, s + k) for k, v in enumerate(GL[s].ravel())]
            rows.append(f"{pos[s][0]}_{pos[s][1]} \t A\tC  " + " \t".join(vals) + " ")
        text = "marker a1 a2 " + " ".join(f"i{k}" for k in range(N_IND * 3)) + "\n" + "\n".join(rows) + "\n"
    elif name == "no_header_extra_numeric_columns":
        rows = ["7\t8\t9\t" + "\t".join(repr(float(v)) for v in GL[s].ravel()) for s in range(N_SITES)]
        text = "\n".join(rows) + "\n"
    elif name == "log_scale_with_neg_inf":
        with np.errstate(divide="ignore"):
            LG = np.log(GL)
        rows = ["\t".join("-inf" if np.isneginf(v) else repr(float(v)) for v in LG[s].ravel()) for s in range(N_SITES)]
        text = "\n".join(rows) + "\n"
        flags = ["--log_scale"]
    elif name == "called_genotypes_with_labels":
        G = GL.argmax(-1)
        G[2, 1] = -1
        rows = [f"{pos[s][0]}\t{pos[s][1]}\t" + "\t".join(str(int(g)) for g in G[s]) for s in range(N_SITES)]
        text = "\n".join(rows) + "\n"
        flags = []
    else:
        raise KeyError(name)
    with gzip.open(geno, "wt") as fh:
        fh.write(text)
    if name == "mixed_separators_and_spellings":
        use_header = True
        body = "chrom\tposition\tnote\n# comment\n\n" + "".join(f"{c}\t{p}\tx{k}\ty\n" for k, (c, p) in enumerate(pos))
        with gzip.open(posf + ".gz", "wt") as fh:
            fh.write(body)
        posf += ".gz"
    else:
        with open(posf, "w") as fh:
            fh.write("".join(f"{c}\t{p}\n" for c, p in pos))
    return geno, posf, flags, use_header


@pytest.mark.parametrize("name", ["mixed_separators_and_spellings", "no_header_extra_numeric_columns",
                                  "log_scale_with_neg_inf", "called_genotypes_with_labels"])
@pytest.mark.parametrize("extra", [["--max_kb_dist", "0", "--extend_out"], ["--max_kb_dist", "1", "--min_maf", "0.1"]])
def test_text_quirks_match_the_reference_binary(name, extra, tmp_path):
    geno, posf, flags, use_header = write_case(tmp_path, name)
    ref_out = str(tmp_path / "ref.ld")
    cmd = [O.REF_BIN, "--geno", geno, "--n_ind", str(N_IND), "--n_sites", str(N_SITES), "--posH" if use_header else "--pos",
           posf] + flags + extra + ["--n_threads", "1", "--verbose", "0", "--out", ref_out]
    subprocess.check_call(cmd, stderr=subprocess.DEVNULL)
    opt = H.parse_flags(flags + extra)
    probs = bool(flags)
    cells, log_cells = N.load_geno(geno, N_IND, N_SITES, probs=probs, log_scale=opt["log_scale"])
    gl, expg, maf = N.prepare_sites(cells, log_scale=opt["log_scale"], from_log_cells=log_cells)
    labels, dist = N.read_positions(posf, N_SITES, header=use_header)
    mine = str(tmp_path / "mine.ld")
    O.run(gl, expg, maf, dist, labels, opt["max_kb_dist"], opt["max_snp_dist"], opt["min_maf"], opt["rnd_sample"],
          opt["seed"], opt["ignore_miss"], opt["extend_out"], out_path=mine)
    assert open(mine, "rb").read() == open(ref_out, "rb").read()


def test_crlf_line_endings_fail_in_both(tmp_path):
    """chomp() removes one trailing character only, so the last field of a CRLF line keeps its \\r, is not numeric and
    is dropped: the reference aborts with 'Less fields than expected' on the second site, and so does the loader."""
    GL = _gl()
    geno = str(tmp_path / "crlf.gz")
    with gzip.open(geno, "wb") as fh:
        fh.write(b"".join(("\t".join(repr(float(v)) for v in GL[s].ravel()) + "\r\n").encode() for s in range(N_SITES)))
    r = subprocess.run([O.REF_BIN, "--geno", geno, "--probs", "--n_ind", str(N_IND), "--n_sites", str(N_SITES),
                        "--max_kb_dist", "0", "--verbose", "0", "--out", str(tmp_path / "o.ld")], capture_output=True)
    assert r.returncode != 0
    with pytest.raises(N.NgsldError):
        N.load_geno(geno, N_IND, N_SITES)
