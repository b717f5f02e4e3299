#!/usr/bin/env python
"""Regenerates tests/golden/ from the UNMODIFIED reference binary (oracle/_ref/ngsLD, built by
`make -C oracle ref` from /root/reference against the GSL stand-in).  Run in the build container only:

    python tests/golden/make_golden.py

Small fixtures (edge, tiny) are committed whole (inputs + gzip'd reference TSV); the larger synthetic
ones (s, p, q, u, v) are regenerated from gen_synth.py at test time and pinned by md5 of input and of the
reference's output (manifest.json).
"""
import gzip
import hashlib
import json
import os
import subprocess
import sys
import tempfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, HERE)
import gen_synth  # noqa: E402

REF = os.path.join(ROOT, "oracle", "_ref", "ngsLD")


def md5(b):
    return hashlib.md5(b).hexdigest()


def edge_fixture():
    rng = np.random.default_rng(3)
    GL = rng.dirichlet([0.3] * 3, (8, 6))
    GL[1, :] = [1, 0, 0]
    GL[2, :] = [1 / 3] * 3
    GL[3, 0] = [0, 0, 0]
    GL[4, :] = [0, 0, 1]
    GL[5, :3] = [1, 0, 0]
    GL[5, 3:] = [0, 0, 1]
    pos = "".join(f"c1\t{100 * (k + 1)}\n" for k in range(6)) + "c2\t700\nc2\t800\n"
    return GL.astype("<f8"), pos


def run_ref(geno, n_ind, n_sites, pos, flags):
    with tempfile.NamedTemporaryFile(suffix=".ld") as out:
        cmd = [REF, "--geno", geno, "--n_ind", str(n_ind), "--n_sites", str(n_sites)]
        if pos:
            cmd += ["--pos", pos]
        cmd += flags + ["--n_threads", "1", "--verbose", "0", "--out", out.name]
        subprocess.check_call(cmd, stderr=subprocess.DEVNULL)
        return open(out.name, "rb").read()


# name -> (flags, needs_pos)
VARIANTS = {
    "ext": (["--probs", "--max_kb_dist", "0", "--extend_out"], True),
    "plain": (["--probs", "--max_kb_dist", "0"], True),
    "kb3": (["--probs", "--max_kb_dist", "3", "--extend_out"], True),
    "snp5": (["--probs", "--max_kb_dist", "0", "--max_snp_dist", "5", "--extend_out"], True),
    "maf20": (["--probs", "--max_kb_dist", "0", "--min_maf", "0.2", "--extend_out"], True),
    "rnd": (["--probs", "--max_kb_dist", "0", "--rnd_sample", "0.5", "--seed", "12345", "--extend_out"], True),
    "nomiss": (["--probs", "--max_kb_dist", "0", "--ignore_miss_data", "--extend_out"], True),
    "callgeno": (["--probs", "--max_kb_dist", "0", "--call_geno", "--extend_out"], True),
    "callgeno_thr": (["--probs", "--max_kb_dist", "0", "--call_geno", "--N_thresh", "0.3", "--call_thresh", "0.9",
                      "--extend_out"], True),
    "nopos": (["--probs", "--max_kb_dist", "0", "--extend_out"], False),
}


def main():
    man = {"reference": "fgvieira/ngsLD 1.2.1 (596bec1f), shim-built oracle/_ref/ngsLD, --n_threads 1", "fixtures": {}}
    with tempfile.TemporaryDirectory() as tmp:
        # ---- committed small fixtures -------------------------------------------------------------
        GL, pos = edge_fixture()
        GL.tofile(os.path.join(HERE, "edge.glf"))
        open(os.path.join(HERE, "edge.glf.pos"), "w").write(pos)
        GLt, post = gen_synth.synth(40, 24, 4)
        # sprinkle exact-missing and exact-called cells so --ignore_miss_data / --call_geno paths bite
        GLt[::7, ::5] = [1 / 3, 1 / 3, 1 / 3]
        gen_synth.write(os.path.join(HERE, "tiny.glf"), GLt, post)
        np.log(GLt).astype("<f8").tofile(os.path.join(HERE, "tiny_log.glf"))
        for name, n_sites, n_ind in (("edge", 8, 6), ("tiny", 40, 24)):
            geno = os.path.join(HERE, name + ".glf")
            fx = {"n_sites": n_sites, "n_ind": n_ind, "input_md5": md5(open(geno, "rb").read()), "variants": {}}
            for v, (flags, needs_pos) in VARIANTS.items():
                out = run_ref(geno, n_ind, n_sites, geno + ".pos" if needs_pos else None, flags)
                with gzip.GzipFile(os.path.join(HERE, f"{name}.{v}.ld.gz"), "wb", mtime=0) as fh:
                    fh.write(out)
                fx["variants"][v] = {"flags": flags, "pos": needs_pos, "md5": md5(out), "rows": out.count(b"\n") - 1}
            if name == "tiny":
                flags = ["--log_scale", "--max_kb_dist", "0", "--extend_out"]
                out = run_ref(os.path.join(HERE, "tiny_log.glf"), n_ind, n_sites, geno + ".pos", flags)
                with gzip.GzipFile(os.path.join(HERE, "tiny.logscale.ld.gz"), "wb", mtime=0) as fh:
                    fh.write(out)
                fx["variants"]["logscale"] = {"flags": flags, "pos": True, "md5": md5(out),
                                              "rows": out.count(b"\n") - 1, "geno": "tiny_log.glf"}
            man["fixtures"][name] = fx
        # ---- regenerated fixtures, md5-pinned -----------------------------------------------------
        for name, n_sites, n_ind, seed, variants in (
                ("s", 400, 40, 2, {"ext": VARIANTS["ext"][0],
                                   "rnd01": ["--probs", "--max_kb_dist", "0", "--rnd_sample", "0.01", "--seed", "1"],
                                   "kb20": ["--probs", "--max_kb_dist", "20", "--extend_out"]}),
                ("p", 600, 100, 5, {"ext": VARIANTS["ext"][0]}),
                ("q", 300, 500, 7, {"ext": VARIANTS["ext"][0]}),
                # the multi-warp-per-pair shapes of BASELINE configs 4 and 5: banded window / random sampling
                ("u", 300, 1000, 21, {"kb50": ["--probs", "--max_kb_dist", "50", "--extend_out"]}),
                ("v", 300, 2000, 22, {"rnd30": ["--probs", "--max_kb_dist", "0", "--rnd_sample", "0.3", "--seed", "7",
                                                "--extend_out"]})):
            geno = os.path.join(tmp, name + ".glf")
            GLs, poss = gen_synth.synth(n_sites, n_ind, seed)
            gen_synth.write(geno, GLs, poss)
            fx = {"n_sites": n_sites, "n_ind": n_ind, "seed": seed, "input_md5": md5(open(geno, "rb").read()),
                  "pos_md5": md5(open(geno + ".pos", "rb").read()), "variants": {}}
            for v, flags in variants.items():
                out = run_ref(geno, n_ind, n_sites, geno + ".pos", flags)
                fx["variants"][v] = {"flags": flags, "pos": True, "md5": md5(out), "rows": out.count(b"\n") - 1}
                print(name, v, fx["variants"][v]["md5"], fx["variants"][v]["rows"], flush=True)
            man["fixtures"][name] = fx
    json.dump(man, open(os.path.join(HERE, "manifest.json"), "w"), indent=1, sort_keys=True)


if __name__ == "__main__":
    main()
