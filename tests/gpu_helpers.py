"""Helpers for the -m gpu parity tests: run the CUDA path through the C ABI (ngsld_b200.Engine) on a
manifest fixture and compare with the oracle (test infrastructure)."""
import numpy as np

import helpers as H
import ngsld_b200 as N
from helpers import O


def engine_for(raw, opt, labels=None, dist=None, device=0):
    gl, expg, maf = N.prepare_sites(raw, log_scale=opt["log_scale"], ignore_miss_data=opt["ignore_miss"],
                                    call_geno=opt["call_geno"], N_thresh=opt["n_thresh"],
                                    call_thresh=opt["call_thresh"])
    eng = N.Engine(device)
    eng.set_sites(gl, expg, maf)
    eng.set_positions(dist, labels)
    return eng, (gl, expg, maf)


def scan_params(opt, strict):
    return N.ScanParams.make(max_kb_dist=opt["max_kb_dist"], max_snp_dist=opt["max_snp_dist"], min_maf=opt["min_maf"],
                             rnd_sample=opt["rnd_sample"], seed=opt["seed"], ignore_miss_data=int(opt["ignore_miss"]),
                             extend_out=int(opt["extend_out"]), strict=int(strict))


def oracle_rows(site_arrays, s1, s2, ignore_miss=False):
    """Oracle results for explicit pairs as a ROW_DTYPE array (dist/s1/s2 not filled)."""
    gl, expg, maf = site_arrays
    out = np.zeros(len(s1), N.ROW_DTYPE)
    for k, (a, b) in enumerate(zip(s1, s2)):
        o = O.pair(gl, expg, maf, int(a), int(b), ignore_miss)
        out[k]["r2_expg"], out[k]["D"], out[k]["Dp"], out[k]["r2"] = o.r2pear, o.D, o.Dp, o.r2
        out[k]["hap"] = list(o.hap)
        out[k]["hap_maf"] = list(o.hmaf)
        out[k]["chi2"] = o.chi2
        out[k]["n_iter"], out[k]["n_used"] = o.n_iter, o.n_used
        out[k]["s1"], out[k]["s2"] = a, b
    return out


def bits(a):
    return np.ascontiguousarray(a).view(np.uint64 if a.dtype == np.float64 else np.uint32)


def same_bits_or_nan(a, b):
    """bit-equal, treating any NaN as equal to any NaN (the GPU's canonical NaN has another payload)."""
    a, b = np.asarray(a), np.asarray(b)
    return bool(np.all((bits(a) == bits(b)) | (np.isnan(a) & np.isnan(b))))


def assert_strict_equal(got, ref):
    for f in ("r2_expg", "D", "Dp", "r2", "hap", "hap_maf", "chi2"):
        assert same_bits_or_nan(got[f], ref[f]), f
    assert np.array_equal(got["n_iter"], ref["n_iter"])
    assert np.array_equal(got["n_used"], ref["n_used"])


def _derived(f):
    """D', r2 from haplotype frequencies (reference ngsLD.cpp:296-306), vectorised float64."""
    f0, f1, f2, f3 = f[..., 0], f[..., 1], f[..., 2], f[..., 3]
    with np.errstate(all="ignore"):
        m0, m1 = 1 - (f0 + f1), 1 - (f0 + f2)
        D = f0 * f3 - f1 * f2
        den = np.where(D < 0, -np.minimum(m0 * m1, (1 - m0) * (1 - m1)), np.minimum(m0 * (1 - m1), (1 - m0) * m1))
        q = D / np.sqrt(m0 * m1 * (1 - m0) * (1 - m1))
        return D / den, q * q


def _condition(f):
    """kappa = sum_k |d g / d f_k| for g in (D', r2), by central differences at the reference frequencies."""
    kd, kr = np.zeros(len(f)), np.zeros(len(f))
    for k in range(4):
        h = 1e-6 * np.abs(f[:, k]) + 1e-300
        up, dn = f.copy(), f.copy()
        up[:, k] += h
        dn[:, k] -= h
        (d1, r1), (d0, r0) = _derived(up), _derived(dn)
        with np.errstate(all="ignore"):
            kd += np.nan_to_num(np.abs(d1 - d0) / (2 * h), nan=np.inf, posinf=np.inf)
            kr += np.nan_to_num(np.abs(r1 - r0) / (2 * h), nan=np.inf, posinf=np.inf)
    return kd, kr


# Tally of the fast-kernel contract over a test session (printed by tests/conftest.py at the end of the run):
# how many pair values were compared and how many of them needed more than the flat 1e-9 of north_star.
SLACK = {"values": 0, "needed_slack": 0, "max_kappa_of_slack": 0.0, "max_err": 0.0, "nan_pattern_forgiven": 0}
WELL_CONDITIONED = 1e5   # kappa below this: the quotient is as stable as its inputs, the flat 1e-9 must hold


def assert_fast_close(got, ref, tol=1e-9):
    """north_star contract: r2_ExpG bit-exact; D, D', r2 within 1e-9 at the same iteration count.

    hap / hap_maf / D: absolute 1e-9 (observed ~1e-16).  D' and r2 are quotients whose denominators vanish
    for near-monomorphic pairs, so their bound is tol + kappa * 1e-15 with kappa the condition number of the
    quotient at the reference frequencies.  Every value that needs more than the flat `tol` is counted in SLACK,
    and none of them may be well conditioned (kappa <= 1e5, where kappa * 1e-15 is 1e-10 at most): only where a
    1e-15 change of a frequency already moves the reference's own value does the bound widen."""
    assert same_bits_or_nan(got["r2_expg"], ref["r2_expg"]), "r2_expg not bit-exact"
    assert np.array_equal(got["n_iter"], ref["n_iter"]), "nIter differs"
    assert np.array_equal(got["n_used"], ref["n_used"])
    fref = np.asarray(ref["hap"], np.float64)
    kd, kr = _condition(np.nan_to_num(fref))
    bound = {"D": 0.0, "hap": 0.0, "hap_maf": 0.0, "Dp": kd, "r2": kr}
    for f in ("D", "Dp", "r2", "hap", "hap_maf"):
        a, b = np.asarray(got[f], np.float64), np.asarray(ref[f], np.float64)
        fin = np.isfinite(a) & np.isfinite(b)
        kappa = bound[f] if np.ndim(bound[f]) == 0 or a.ndim == 1 else bound[f][:, None]
        kappa = np.broadcast_to(np.asarray(kappa, np.float64), a.shape)
        allowed = tol + 1e-15 * kappa
        # NaN / inf patterns must agree wherever the value is well conditioned
        odd = (np.isnan(a) != np.isnan(b)) | (np.isinf(a) != np.isinf(b))
        assert not np.any(odd & (kappa <= WELL_CONDITIONED)), f + " NaN/inf pattern"
        SLACK["nan_pattern_forgiven"] += int(odd.sum())
        err = np.abs(a[fin] - b[fin])
        SLACK["values"] += int(err.size)
        if err.size == 0:
            continue
        over = err > tol
        assert not np.any(over & (kappa[fin] <= WELL_CONDITIONED)), (f, "well-conditioned value beyond 1e-9", float(err[over].max()))
        assert np.all(err <= allowed[fin]), (f, float(err.max()))
        SLACK["needed_slack"] += int(over.sum())
        SLACK["max_err"] = max(SLACK["max_err"], float(err.max()))
        if over.any():
            SLACK["max_kappa_of_slack"] = max(SLACK["max_kappa_of_slack"], float(kappa[fin][over].max()))
