"""The ngsLD-compatible command line (ngsld_b200/bin/ngsLD).  CPU: flag handling, validation messages and exit
status follow the reference (parse_args.cpp:63-183, shared/gen_func.cpp:12-18).  GPU (-m gpu): the output file is
byte-identical to the unmodified reference's for every golden case, binary and text inputs alike."""
import gzip
import os
import subprocess

import pytest

import helpers as H

CLI = os.path.join(H.ROOT, "ngsld_b200", "bin", "ngsLD")
TINY = os.path.join(H.GOLD, "tiny.glf")


def run_cli(args, **kw):
    return subprocess.run([CLI] + args, capture_output=True, **kw)


@pytest.fixture(scope="module", autouse=True)
def _built():
    if not os.path.exists(CLI):
        subprocess.check_call(["make", "-s", "-C", os.path.join(H.ROOT, "ngsld_b200", "csrc"), "all"])


@pytest.mark.parametrize("args,msg", [
    ([], "genotype input file (--geno) missing!"),
    (["--geno", TINY], "number of individuals (--n_ind) missing!"),
    (["--geno", TINY, "--n_ind", "24"], "number of sites (--n_sites) missing!"),
    (["--geno", TINY, "--n_ind", "24", "--n_sites", "40"], "position file necessary in order to filter by maximum distance!"),
    (["--geno", TINY, "--n_ind", "24", "--n_sites", "40", "--max_kb_dist", "0", "--min_maf", "1.5"], "minimum allele frequency must be in [0,1]!"),
    (["--geno", TINY, "--n_ind", "24", "--n_sites", "40", "--max_kb_dist", "0", "--call_geno"], "can only call genotypes from likelihoods/probabilities!"),
    (["--geno", TINY, "--n_ind", "24", "--n_sites", "40", "--max_kb_dist", "0", "--rnd_sample", "0"], "proportion of comparisons to sample must be in ]0,1]!"),
    (["--geno", TINY, "--n_ind", "24", "--n_sites", "40", "--max_kb_dist", "0", "--n_threads", "0"], "number of threads cannot be less than 1!"),
])
def test_argument_validation_messages(args, msg):
    r = run_cli(args + ["--verbose", "0"])
    assert r.returncode == 255                                   # exit(-1)
    assert f"ERROR: [parse_cmd_args] {msg}".encode() in r.stderr
    assert r.stdout == b""


def test_unknown_option_and_outH_exit_like_the_reference():
    assert run_cli(["--bogus"]).returncode == 255
    assert run_cli(["--outH", "x"]).returncode == 255            # in the reference's table but without a case


def test_file_checks():
    r = run_cli(["--geno", "/nonexistent.glf", "--n_ind", "2", "--n_sites", "2", "--max_kb_dist", "0", "--verbose", "0"])
    assert r.returncode == 255 and b"ERROR: [main] cannot check GENO file size!" in r.stderr
    r = run_cli(["--geno", TINY, "--n_ind", "24", "--n_sites", "41", "--max_kb_dist", "0", "--verbose", "0"])
    assert r.returncode == 255 and b"ERROR: [main] invalid/corrupt genotype input file!" in r.stderr


def test_single_dash_long_options_and_argument_echo():
    r = run_cli(["-geno", TINY, "-n_ind", "24", "-n_sites", "41", "-max_kb_dist", "0", "-probs", "-N_thresh", "0.2"])
    assert b"==> Input Arguments:" in r.stderr and b"\tcall_geno: true\n" in r.stderr and b"\tn_sites: 41\n" in r.stderr
    assert b"BINARY input file (always lkl)" in r.stderr


@pytest.mark.skipif(os.path.exists("/dev/nvidia0"), reason="box has a GPU")
def test_no_gpu_is_fatal_not_a_cpu_fallback():
    r = run_cli(["--geno", TINY, "--n_ind", "24", "--n_sites", "40", "--max_kb_dist", "0", "--verbose", "0"])
    assert r.returncode == 255 and b"no CUDA device available" in r.stderr
    assert r.stdout.count(b"\n") == 1                            # only the header was written before the failure


# ---- GPU: byte identity with the reference ------------------------------------------------------------
BIN_CASES = [(fx, v) for fx in ("edge", "tiny") for v in H.MANIFEST["fixtures"][fx]["variants"]]


@pytest.mark.gpu
@pytest.mark.parametrize("fx,variant", BIN_CASES)
def test_cli_strict_output_is_byte_identical_binary_inputs(fx, variant, tmp_path):
    v = H.MANIFEST["fixtures"][fx]["variants"][variant]
    d = H.MANIFEST["fixtures"][fx]
    geno = os.path.join(H.GOLD, v.get("geno") or fx + ".glf")
    args = ["--geno", geno, "--n_ind", str(d["n_ind"]), "--n_sites", str(d["n_sites"])] + v["flags"]
    if v["pos"]:
        args += ["--pos", os.path.join(H.GOLD, fx + ".glf.pos")]
    out = tmp_path / "o.ld"
    r = run_cli(args + ["--gpu_strict", "--gpu_n", "1", "--verbose", "0", "--out", str(out)])
    assert r.returncode == 0, r.stderr.decode()
    assert out.read_bytes() == H.golden_bytes(fx, variant)


@pytest.mark.gpu
@pytest.mark.parametrize("fx,variant", [("p", "ext"), ("q", "ext"), ("u", "kb50"), ("v", "rnd30")])
def test_cli_strict_output_md5_at_headline_sample_sizes(fx, variant, tmp_path):
    """n_ind = 100 / 500 / 1000 (banded) / 2000 (sampled): the CLI's file has the md5 of the unmodified reference's."""
    d = H.MANIFEST["fixtures"][fx]
    v = d["variants"][variant]
    geno, pos = H.fixture_paths(fx, tmp_path)
    out = tmp_path / "o.ld"
    r = run_cli(["--geno", geno, "--n_ind", str(d["n_ind"]), "--n_sites", str(d["n_sites"]), "--pos", pos] + v["flags"] +
                ["--gpu_strict", "--verbose", "0", "--out", str(out)])
    assert r.returncode == 0, r.stderr.decode()
    got = out.read_bytes()
    assert got.count(b"\n") - 1 == v["rows"]
    assert H.md5(got) == v["md5"]


@pytest.mark.gpu
@pytest.mark.parametrize("name", sorted(H.MANIFEST["text_cases"]))
def test_cli_strict_output_is_byte_identical_text_inputs(name, tmp_path):
    c = H.MANIFEST["text_cases"][name]
    args = ["--geno", os.path.join(H.GOLD, c["geno"]), "--n_ind", str(c["n_ind"]), "--n_sites", str(c["n_sites"]),
            "--posH" if c["posH"] else "--pos", os.path.join(H.GOLD, c["pos"])] + c["flags"]
    r = run_cli(args + ["--gpu_strict", "--gpu_n", "1", "--verbose", "0"])          # to stdout
    assert r.returncode == 0, r.stderr.decode()
    assert r.stdout == gzip.open(os.path.join(H.GOLD, f"tiny.{name}.ld.gz"), "rb").read()


@pytest.mark.gpu
def test_cli_fast_kernel_and_all_gpus_keep_reference_row_order(tmp_path):
    """Default (fast) kernel on every visible GPU: same rows in the same order, text equal up to the last digit."""
    fx = H.MANIFEST["fixtures"]["s"]
    geno, pos = H.fixture_paths("s", tmp_path)
    out = tmp_path / "s.ld"
    r = run_cli(["--geno", geno, "--probs", "--n_ind", str(fx["n_ind"]), "--n_sites", str(fx["n_sites"]), "--pos", pos,
                 "--max_kb_dist", "20", "--extend_out", "--gpu_stats", "--verbose", "0", "--out", str(out)])
    assert r.returncode == 0, r.stderr.decode()
    assert b"[gpu 0]" in r.stderr
    got = out.read_bytes().splitlines()
    want = H.oracle_tsv("s", tmp_path, fx["variants"]["kb20"]["flags"]).splitlines()
    assert len(got) == len(want) == fx["variants"]["kb20"]["rows"] + 1
    for a, b in zip(got, want):
        fa, fb = a.split(b"\t"), b.split(b"\t")
        assert fa[:4] == fb[:4] and fa[7:10] == fb[7:10] and fa[-2:] == fb[-2:]
        if a != b:
            for x, y in zip(fa[4:], fb[4:]):
                assert x == y or abs(float(x) - float(y)) <= 1.000001e-6 * max(1.0, abs(float(y)))


@pytest.mark.gpu
@pytest.mark.parametrize("slab_rows", ["97", "1000"])
def test_cli_many_slabs_keep_bytes_and_order(slab_rows, tmp_path):
    """The slab queue + writer thread with far more slabs than GPUs: output must not change by a byte."""
    v = H.MANIFEST["fixtures"]["tiny"]["variants"]["ext"]
    args = ["--geno", TINY, "--n_ind", "24", "--n_sites", "40", "--pos", TINY + ".pos"] + v["flags"]
    out = tmp_path / "o.ld"
    r = run_cli(args + ["--gpu_strict", "--verbose", "0", "--gpu_stats", "--out", str(out)],
                env=dict(os.environ, NGSLD_CLI_SLAB_ROWS=slab_rows))
    assert r.returncode == 0, r.stderr.decode()
    assert out.read_bytes() == H.golden_bytes("tiny", "ext")
    assert b"slabs of" in r.stderr


REF = os.path.join(H.ROOT, "oracle", "_ref", "ngsLD")


@pytest.mark.skipif(not os.path.exists(REF), reason="oracle/_ref/ngsLD not built (needs /root/reference)")
@pytest.mark.parametrize("args", [
    [],
    ["--geno", TINY],
    ["--geno", TINY, "--n_ind", "24"],
    ["--geno", TINY, "--n_ind", "24", "--n_sites", "40"],
    ["--geno", TINY, "--n_ind", "24", "--n_sites", "40", "--max_kb_dist", "0", "--min_maf", "-0.1"],
    ["--geno", TINY, "--n_ind", "24", "--n_sites", "40", "--max_kb_dist", "0", "--rnd_sample", "1.5"],
    ["--geno", TINY, "--n_ind", "24", "--n_sites", "39", "--max_kb_dist", "0"],
    ["--geno", "/nonexistent/x.glf", "--n_ind", "24", "--n_sites", "40", "--max_kb_dist", "0"],
    ["--geno", TINY, "--n_ind", "24", "--n_sites", "40", "--max_kb_dist", "0", "--nonsense"],
    ["--geno", os.path.join(H.GOLD, "tiny.geno.gz"), "--n_ind", "24", "--n_sites", "40", "--max_kb_dist", "0", "--call_geno"],
])
def test_fatal_errors_match_the_reference_binary(args, tmp_path):
    """Same exit status and the same 'ERROR: [func] message' line as the unmodified reference for bad invocations
    (all of them fail before any pair is computed, so no GPU is needed)."""
    extra = ["--verbose", "0", "--out", str(tmp_path / "o.ld")]
    mine = run_cli(args + extra)
    ref = subprocess.run([REF] + args + extra, capture_output=True)

    def err_line(b):
        lines = [l for l in b.decode(errors="replace").splitlines() if l.startswith("ERROR:")]
        return lines[0] if lines else None
    assert mine.returncode == ref.returncode
    assert err_line(mine.stderr) == err_line(ref.stderr)


@pytest.mark.gpu
@pytest.mark.parametrize("extra,kw", [([], {}), (["--gpu_prune_keep_heavy"], dict(keep_heavy=True))])
def test_cli_prune_mode_gives_the_pruning_scripts_site_list(extra, kw, tmp_path):
    """--gpu_prune: the site list scripts/prune_graph.pl would produce from the TSV (restated in oracle/prune_oracle.py),
    without the TSV ever being written; all visible GPUs."""
    from oracle import prune_oracle as PO
    fx = H.MANIFEST["fixtures"]["s"]
    geno, pos = H.fixture_paths("s", tmp_path)
    out, excl = tmp_path / "kept.txt", tmp_path / "excl.txt"
    base = ["--geno", geno, "--probs", "--n_ind", str(fx["n_ind"]), "--n_sites", str(fx["n_sites"]), "--pos", pos,
            "--max_kb_dist", "20", "--gpu_strict", "--verbose", "0"]
    r = run_cli(base + ["--gpu_prune", str(out), "--gpu_prune_max_kb_dist", "10", "--gpu_prune_min_weight", "0.2",
                        "--gpu_prune_excl", str(excl), "--gpu_stats"] + extra)
    assert r.returncode == 0, r.stderr.decode()
    assert r.stdout == b"" and b"[prune]" in r.stderr
    tsv = H.oracle_tsv("s", tmp_path, fx["variants"]["kb20"]["flags"])
    nodes, edges = PO.read_edges(tsv, max_kb_dist=10, min_weight=0.2)
    want_kept, want_excl = PO.prune(nodes, edges, **kw)
    got_kept = out.read_text().split()
    assert set(got_kept) == want_kept and len(got_kept) == len(want_kept)
    assert sorted(excl.read_text().split()) == sorted(want_excl) and len(want_excl) > 10
    if not kw:
        assert excl.read_text().split() == want_excl


@pytest.mark.gpu
def test_cli_binary_side_format(tmp_path):
    """--gpu_out_bin: --out holds the rows as 112-byte records in the TSV's row order; formatting them gives the TSV."""
    import numpy as np
    import ngsld_b200 as N
    v = H.MANIFEST["fixtures"]["tiny"]["variants"]["ext"]
    args = ["--geno", TINY, "--n_ind", "24", "--n_sites", "40", "--pos", TINY + ".pos"] + v["flags"]
    out = tmp_path / "o.bin"
    r = run_cli(args + ["--gpu_strict", "--gpu_out_bin", "--verbose", "0", "--out", str(out)],
                env=dict(os.environ, NGSLD_CLI_SLAB_ROWS="100"))
    assert r.returncode == 0, r.stderr.decode()
    rows = np.fromfile(out, N.ROW_DTYPE)
    gold = H.golden_bytes("tiny", "ext").decode().splitlines()[1:]
    assert len(rows) == len(gold) == v["rows"]
    labels = [l.replace("\t", ":") for l in open(TINY + ".pos").read().splitlines()]
    for row, line in list(zip(rows, gold))[::7]:
        f = line.split("\t")
        assert (labels[row["s1"]], labels[row["s2"]]) == (f[0], f[1])
        assert "%.0f" % row["dist"] == f[2] and int(f[-1]) == row["n_iter"] and int(f[7]) == row["n_used"]
        for val, txt in zip([row["r2_expg"], row["D"], row["Dp"], row["r2"]], f[3:7]):
            assert ("%f" % val).replace("nan", "-nan").replace("--", "-") == txt


@pytest.mark.gpu
def test_cli_slab_files_concatenate_to_the_output(tmp_path):
    """--gpu_out_shards: one file per slab, written in parallel; cat in name order is the reference's file."""
    v = H.MANIFEST["fixtures"]["tiny"]["variants"]["ext"]
    args = ["--geno", TINY, "--n_ind", "24", "--n_sites", "40", "--pos", TINY + ".pos"] + v["flags"]
    out = tmp_path / "o.ld"
    r = run_cli(args + ["--gpu_strict", "--gpu_out_shards", "--verbose", "0", "--gpu_stats", "--out", str(out)],
                env=dict(os.environ, NGSLD_CLI_SLAB_ROWS="97"))
    assert r.returncode == 0, r.stderr.decode()
    parts = sorted(tmp_path.glob("o.ld.part-*"))
    assert len(parts) >= 5 and not out.exists()
    assert b"".join(p.read_bytes() for p in parts) == H.golden_bytes("tiny", "ext")
