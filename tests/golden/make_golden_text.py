#!/usr/bin/env python
"""Adds the TEXT-input golden cases (gz genotype-likelihood tables, called genotypes, --posH) to
tests/golden/manifest.json under "text_cases", again from the UNMODIFIED reference binary
(oracle/_ref/ngsLD).  Run in the build container only, after make_golden.py:

    python tests/golden/make_golden_text.py

Inputs are derived from the committed tiny.glf (40 sites x 24 individuals): the reference reads them
through read_geno()'s text branch (shared/read_data.cpp:48-103) and read_dist()/--posH
(shared/read_data.cpp:165-218), i.e. the same cases examples/test.sh exercises with ANGSD output."""
import gzip
import hashlib
import json
import os
import subprocess
import tempfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
REF = os.path.join(ROOT, "oracle", "_ref", "ngsLD")
N_SITES, N_IND = 40, 24


def md5(b):
    return hashlib.md5(b).hexdigest()


def gz_write(path, text):
    with gzip.GzipFile(path, "wb", mtime=0) as fh:
        fh.write(text.encode())


def make_inputs():
    GL = np.fromfile(os.path.join(HERE, "tiny.glf"), "<f8").reshape(N_SITES, N_IND, 3)
    pos = [l.rstrip("\n").split("\t") for l in open(os.path.join(HERE, "tiny.glf.pos"))]
    head = "marker\tallele1\tallele2\t" + "\t".join(f"Ind{i}" for i in range(N_IND) for _ in range(3)) + "\n"
    # beagle-like: 3 label columns (the marker is dropped as non-numeric, the allele codes are numeric but only the
    # LAST 3*n_ind numeric fields are used), shortest round-trip decimals so the text equals the binary doubles
    rows = ["\t".join([f"{c}_{p}", "0", "1"] + [repr(float(v)) for v in GL[s].ravel()]) for s, (c, p) in enumerate(pos)]
    gz_write(os.path.join(HERE, "tiny.beagle.gz"), head + "\n".join(rows) + "\n")
    rows = ["\t".join([f"{c}_{p}", "0", "1"] + [repr(float(v)) for v in np.log(GL[s]).ravel()]) for s, (c, p) in enumerate(pos)]
    gz_write(os.path.join(HERE, "tiny.beagle_log.gz"), head + "\n".join(rows) + "\n")
    # called genotypes -1/0/1/2, blank-separated, no header; chromosome label + position in front
    G = GL.argmax(-1)
    G[np.ptp(GL, axis=-1) < 1e-9] = -1
    rows = [" ".join([c, p] + [str(int(g)) for g in G[s]]) for s, (c, p) in enumerate(pos)]
    gz_write(os.path.join(HERE, "tiny.geno.gz"), "\n".join(rows) + "\n")
    # position file with a header line, a comment and an empty line (read_file skips the latter two)
    with open(os.path.join(HERE, "tiny.posH"), "w") as fh:
        fh.write("chr\tpos\n# a comment\n\n" + "".join(f"{c}\t{p}\n" for c, p in pos))


CASES = {
    "textgl": dict(geno="tiny.beagle.gz", pos="tiny.glf.pos", posH=False, flags=["--probs", "--max_kb_dist", "0", "--extend_out"]),
    "textgl_log": dict(geno="tiny.beagle_log.gz", pos="tiny.glf.pos", posH=False, flags=["--log_scale", "--max_kb_dist", "0", "--extend_out"]),
    "textgl_posH": dict(geno="tiny.beagle.gz", pos="tiny.posH", posH=True, flags=["--probs", "--max_kb_dist", "4", "--extend_out"]),
    "textgl_filters": dict(geno="tiny.beagle.gz", pos="tiny.glf.pos", posH=False,
                           flags=["--probs", "--max_kb_dist", "5", "--min_maf", "0.3", "--ignore_miss_data", "--extend_out"]),
    "textgl_call": dict(geno="tiny.beagle.gz", pos="tiny.glf.pos", posH=False,
                        flags=["--probs", "--max_kb_dist", "0", "--call_geno", "--N_thresh", "0.3", "--call_thresh", "0.9", "--extend_out"]),
    "textgeno": dict(geno="tiny.geno.gz", pos="tiny.glf.pos", posH=False, flags=["--max_kb_dist", "0", "--extend_out"]),
    "textgeno_rnd": dict(geno="tiny.geno.gz", pos="tiny.glf.pos", posH=False,
                         flags=["--max_kb_dist", "0", "--rnd_sample", "0.5", "--seed", "12345"]),
}


def main():
    make_inputs()
    man_path = os.path.join(HERE, "manifest.json")
    man = json.load(open(man_path))
    man["text_cases"] = {}
    for name, c in CASES.items():
        with tempfile.NamedTemporaryFile(suffix=".ld") as out:
            cmd = [REF, "--geno", os.path.join(HERE, c["geno"]), "--n_ind", str(N_IND), "--n_sites", str(N_SITES),
                   "--posH" if c["posH"] else "--pos", os.path.join(HERE, c["pos"])] + c["flags"] + \
                  ["--n_threads", "1", "--verbose", "0", "--out", out.name]
            subprocess.check_call(cmd, stderr=subprocess.DEVNULL)
            res = open(out.name, "rb").read()
        with gzip.GzipFile(os.path.join(HERE, f"tiny.{name}.ld.gz"), "wb", mtime=0) as fh:
            fh.write(res)
        man["text_cases"][name] = dict(c, md5=md5(res), rows=res.count(b"\n") - 1, n_sites=N_SITES, n_ind=N_IND)
        print(name, md5(res), res.count(b"\n") - 1)
    json.dump(man, open(man_path, "w"), indent=1, sort_keys=True)


if __name__ == "__main__":
    main()
