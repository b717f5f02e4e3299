"""TEST INFRASTRUCTURE — ctypes loader for oracle/liboracle.so (the CPU restatement) and helpers to run
the shim-built reference binary oracle/_ref/ngsLD.  Imported only by tests/, __graft_entry__.smoke()
and bench.py's cpu_baseline / --impl reference legs; never by ngsld_b200/."""
import ctypes as C
import os
import subprocess
import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "liboracle.so")
REF_BIN = os.path.join(HERE, "_ref", "ngsLD")

_u64, _dbl, _int = C.c_uint64, C.c_double, C.c_int
_pd = np.ctypeslib.ndpointer(np.float64, flags="C_CONTIGUOUS")
_pu = np.ctypeslib.ndpointer(np.uint64, flags="C_CONTIGUOUS")


class PairOut(C.Structure):
    _fields_ = [("r2pear", _dbl), ("D", _dbl), ("Dp", _dbl), ("r2", _dbl), ("hap", _dbl * 4),
                ("hmaf", _dbl * 2), ("chi2", C.c_float), ("n_used", _u64), ("n_iter", _u64)]


class Taus(C.Structure):
    _fields_ = [("a", C.c_uint32), ("b", C.c_uint32), ("c", C.c_uint32)]


def build():
    subprocess.check_call(["make", "-s", "-C", HERE, "liboracle.so"])
    if os.path.exists("/root/reference/ngsLD.cpp"):
        subprocess.check_call(["make", "-s", "-C", HERE, "ref"])


_lib = None


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            build()
        L = C.CDLL(LIB_PATH)
        L.orc_taus_set.argtypes = [C.POINTER(Taus), _u64]
        L.orc_taus_get.argtypes = [C.POINTER(Taus)]
        L.orc_taus_get.restype = C.c_uint32
        L.orc_site_seeds.argtypes = [_u64, _u64, _pu]
        L.orc_preprocess.argtypes = [_pd, _u64, _u64, _int, _int, _int, _dbl, _dbl, _pd, _pd, _pd]
        L.orc_preprocess.restype = _int
        L.orc_pearson_r2.argtypes = [_pd, _pd, _u64]
        L.orc_pearson_r2.restype = _dbl
        L.orc_pair.argtypes = [_pd, _pd, _pd, _u64, _u64, _u64, _int, C.POINTER(PairOut)]
        L.orc_run.argtypes = [_pd, _pd, _pd, _pd, C.POINTER(C.c_char_p), _u64, _u64, _u64, _u64, _dbl, _dbl,
                              _u64, _int, _int, _u64, _u64, _int, C.c_char_p, _int, C.POINTER(_u64)]
        L.orc_run.restype = _u64
        L.orc_pos_dist.argtypes = [C.POINTER(C.c_char_p), C.POINTER(C.c_char_p), _u64, _pd]
        L.orc_pos_dist.restype = _int
        _lib = L
    return _lib


def preprocess(raw, log_scale=False, ignore_miss=False, call_geno=False, n_thresh=0.0, call_thresh=0.0):
    """raw [n_sites, n_ind, 3] float64 -> (gl, expg, maf) exactly as the reference prepares them."""
    raw = np.ascontiguousarray(raw, np.float64)
    n_sites, n_ind, _ = raw.shape
    gl = np.empty_like(raw)
    expg = np.empty((n_sites, n_ind))
    maf = np.empty(n_sites)
    rc = lib().orc_preprocess(raw, n_sites, n_ind, int(log_scale), int(ignore_miss), int(call_geno),
                              n_thresh, call_thresh, gl, expg, maf)
    if rc != 0:
        raise ValueError("NaN found! Is the file format correct?")
    return gl, expg, maf


def read_pos(path, header=False):
    """labels ('chr:pos', only the first tab replaced) and inter-site distances like the reference
    (ngsLD.cpp:119-132, shared/read_data.cpp:165-218, shared/gen_func.cpp:238-282)."""
    import gzip
    op = gzip.open if open(path, "rb").read(2) == b"\x1f\x8b" else open
    lines = []
    skip = 1 if header else 0
    with op(path, "rt") as fh:
        for ln in fh:
            ln = ln.rstrip("\n").rstrip("\r") if ln.endswith("\n") else ln
            if not ln or ln.startswith("#"):
                continue
            if skip:
                skip -= 1
                continue
            lines.append(ln)
    n = len(lines)
    chrs = (C.c_char_p * n)(*[l.split("\t")[0].encode() for l in lines])
    poss = (C.c_char_p * n)(*[l.split("\t")[1].encode() for l in lines])
    dist = np.empty(n)
    if lib().orc_pos_dist(chrs, poss, n, dist) != 0:
        raise ValueError("invalid distance between adjacent sites!")
    labels = [l.replace("\t", ":", 1) for l in lines]
    return labels, dist


def pair(gl, expg, maf, s1, s2, ignore_miss=False):
    o = PairOut()
    lib().orc_pair(gl, expg, maf, gl.shape[1], s1, s2, int(ignore_miss), C.byref(o))
    return o


def run(gl, expg, maf, pos_dist=None, labels=None, max_kb_dist=0, max_snp_dist=0, min_maf=0.0,
        rnd_sample=1.0, seed=1, ignore_miss=False, extend_out=True, s1_lo=0, s1_hi=None, n_threads=1,
        out_path=None, header=True):
    """Whole scan; returns (n_pairs, sum_em_passes).  Rows go to out_path in (s1, s2) order."""
    n_sites, n_ind = expg.shape
    if pos_dist is None:
        pos_dist = np.full(n_sites, np.inf)
    lab = None
    if labels is not None:
        lab = (C.c_char_p * n_sites)(*[l.encode() for l in labels])
    it = _u64(0)
    n = lib().orc_run(gl, expg, maf, np.ascontiguousarray(pos_dist, np.float64), lab, n_sites, n_ind,
                      int(max_kb_dist), int(max_snp_dist), float(min_maf), float(rnd_sample), int(seed),
                      int(ignore_miss), int(extend_out), s1_lo, n_sites if s1_hi is None else s1_hi,
                      n_threads, out_path.encode() if out_path else None, int(header), C.byref(it))
    return n, it.value


def have_ref():
    return os.path.exists(REF_BIN) and os.access(REF_BIN, os.X_OK)


def run_ref(args, out_path, n_threads=1):
    """Run the unmodified reference CLI (shim-built) with --verbose 0."""
    cmd = [REF_BIN] + list(args) + ["--n_threads", str(n_threads), "--verbose", "0", "--out", out_path]
    subprocess.check_call(cmd, stderr=subprocess.DEVNULL)
