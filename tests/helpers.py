"""Shared test helpers: fixture materialisation, flag parsing, oracle-side TSV production."""
import gzip
import hashlib
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
GOLD = os.path.join(HERE, "golden")
sys.path.insert(0, GOLD)
import gen_synth  # noqa: E402
from oracle import oracle as O  # noqa: E402  (test infrastructure only)

MANIFEST = json.load(open(os.path.join(GOLD, "manifest.json")))


def md5(b):
    return hashlib.md5(b).hexdigest()


def fixture_paths(name, tmpdir, geno=None):
    """(geno_path, pos_path) for a manifest fixture; small ones are committed, larger are regenerated."""
    fx = MANIFEST["fixtures"][name]
    if name in ("edge", "tiny"):
        return os.path.join(GOLD, geno or (name + ".glf")), os.path.join(GOLD, name + ".glf.pos")
    path = os.path.join(str(tmpdir), name + ".glf")
    if not os.path.exists(path):
        GL, pos = gen_synth.synth(fx["n_sites"], fx["n_ind"], fx["seed"])
        gen_synth.write(path, GL, pos)
    assert md5(open(path, "rb").read()) == fx["input_md5"], "numpy stream drifted: regenerate goldens"
    return path, path + ".pos"


def golden_bytes(name, variant):
    p = os.path.join(GOLD, f"{name}.{variant}.ld.gz")
    return gzip.open(p, "rb").read() if os.path.exists(p) else None


def parse_flags(flags):
    """Reference CLI flags (parse_args.cpp:35-132) -> dict of the knobs that shape the pair scan."""
    o = dict(log_scale=False, max_kb_dist=100, max_snp_dist=0, min_maf=0.0, ignore_miss=False, call_geno=False,
             n_thresh=0.0, call_thresh=0.0, rnd_sample=1.0, seed=1, extend_out=False)
    it = iter(flags)
    for f in it:
        if f == "--probs":
            pass
        elif f == "--log_scale":
            o["log_scale"] = True
        elif f == "--max_kb_dist":
            o["max_kb_dist"] = int(next(it))
        elif f == "--max_snp_dist":
            o["max_snp_dist"] = int(next(it))
        elif f == "--min_maf":
            o["min_maf"] = float(next(it))
        elif f == "--ignore_miss_data":
            o["ignore_miss"] = True
        elif f == "--call_geno":
            o["call_geno"] = True
        elif f == "--N_thresh":
            o["n_thresh"] = float(next(it)); o["call_geno"] = True
        elif f == "--call_thresh":
            o["call_thresh"] = float(next(it)); o["call_geno"] = True
        elif f == "--rnd_sample":
            o["rnd_sample"] = float(next(it))
        elif f == "--seed":
            o["seed"] = int(next(it))
        elif f == "--extend_out":
            o["extend_out"] = True
        else:
            raise ValueError(f)
    return o


def load_fixture(name, tmpdir, flags, use_pos=True, geno=None):
    fx = MANIFEST["fixtures"][name]
    gpath, ppath = fixture_paths(name, tmpdir, geno)
    opt = parse_flags(flags)
    raw = np.fromfile(gpath, "<f8").reshape(fx["n_sites"], fx["n_ind"], 3)
    labels, dist = (O.read_pos(ppath) if use_pos else (None, None))
    return raw, labels, dist, opt


def oracle_tsv(name, tmpdir, flags, use_pos=True, geno=None, n_threads=4):
    raw, labels, dist, opt = load_fixture(name, tmpdir, flags, use_pos, geno)
    gl, expg, maf = O.preprocess(raw, opt["log_scale"], opt["ignore_miss"], opt["call_geno"], opt["n_thresh"],
                                 opt["call_thresh"])
    out = os.path.join(str(tmpdir), f"{name}.oracle.ld")
    O.run(gl, expg, maf, dist, labels, opt["max_kb_dist"], opt["max_snp_dist"], opt["min_maf"], opt["rnd_sample"],
          opt["seed"], opt["ignore_miss"], opt["extend_out"], n_threads=n_threads, out_path=out)
    return open(out, "rb").read()
