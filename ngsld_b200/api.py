"""ctypes mirror of include/ngsld_b200.h (one method per C entry point, same argument meaning).

The binding is deliberately thin: every call goes straight into libngsld_b200.so.  Nothing here computes
LD on the CPU; if the shared library is absent the import of the library raises, loudly."""
import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))

ROW_DTYPE = np.dtype([("dist", "<f8"), ("r2_expg", "<f8"), ("D", "<f8"), ("Dp", "<f8"), ("r2", "<f8"),
                      ("hap", "<f8", (4,)), ("hap_maf", "<f8", (2,)), ("chi2", "<f4"), ("n_iter", "<u4"),
                      ("n_used", "<u4"), ("s1", "<u4"), ("s2", "<u4"), ("reserved", "<u4")], align=True)
assert ROW_DTYPE.itemsize == 112
DECAY_DTYPE = np.dtype([("n", "<u8", (4,)), ("sum", "<f8", (4,))])

E_CODES = {-1: "NGSLD_E_INVALID", -2: "NGSLD_E_CUDA", -3: "NGSLD_E_NOMEM", -4: "NGSLD_E_DATA", -5: "NGSLD_E_SINK",
           -6: "NGSLD_E_IO"}


class NgsldError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__(f"{E_CODES.get(code, code)}: {msg}")
        self.code = code


class ScanParams(C.Structure):
    """ngsld_scan_params: the reference's `params` fields that shape the scan (ngsLD.hpp:11-44)."""
    _fields_ = [("max_kb_dist", C.c_uint64), ("max_snp_dist", C.c_uint64), ("min_maf", C.c_double),
                ("rnd_sample", C.c_double), ("seed", C.c_uint64), ("ignore_miss_data", C.c_int),
                ("extend_out", C.c_int), ("strict", C.c_int), ("reserved", C.c_int)]

    @classmethod
    def make(cls, **kw):
        p = cls()
        load_library().ngsld_scan_defaults(C.byref(p))
        for k, v in kw.items():
            if not hasattr(p, k):
                raise TypeError(k)
            setattr(p, k, v)
        return p


class ScanStats(C.Structure):
    _fields_ = [("n_pairs", C.c_uint64), ("sum_em_passes", C.c_uint64), ("n_launches", C.c_uint64),
                ("ms_em", C.c_double), ("ms_pearson", C.c_double), ("ms_format", C.c_double),
                ("ms_device_total", C.c_double), ("ms_plan", C.c_double), ("h2d_bytes", C.c_uint64),
                ("d2h_bytes", C.c_uint64), ("em_kernel", C.c_char * 64), ("n_cell_pairs", C.c_uint64),
                ("sum_cells", C.c_uint64), ("sum_cell_passes", C.c_uint64), ("n_resid_pairs", C.c_uint64)]

    def as_dict(self):
        d = {k: getattr(self, k) for k, _ in self._fields_}
        d["em_kernel"] = d["em_kernel"].decode()
        return d


class PruneParams(C.Structure):
    """ngsld_prune_params: the edge filter of scripts/prune_graph.pl (--max_kb_dist * 1000, --min_weight, --field_weight,
    --weight_type, weight precision)."""
    _fields_ = [("max_dist", C.c_double), ("min_weight", C.c_double), ("field", C.c_int), ("weight_type", C.c_int),
                ("weight_precision", C.c_int), ("reserved", C.c_int)]

    @classmethod
    def make(cls, max_dist=float("inf"), min_weight=0.0, field=7, weight_type="a", weight_precision=4):
        return cls(max_dist, min_weight, field, ord(weight_type), weight_precision, 0)


EDGE_DTYPE = np.dtype([("s1", "<u4"), ("s2", "<u4"), ("label", "<i4")])
EDGE_SINK = C.CFUNCTYPE(C.c_int, C.c_void_p, C.c_void_p, C.c_uint64)
ROW_SINK = C.CFUNCTYPE(C.c_int, C.c_void_p, C.c_void_p, C.c_uint64)
TEXT_SINK = C.CFUNCTYPE(C.c_int, C.c_void_p, C.c_void_p, C.c_uint64, C.c_uint64)


def lib_path():
    return os.environ.get("NGSLD_B200_LIB", os.path.join(_HERE, "libngsld_b200.so"))


_LIB = None


def load_library():
    """dlopen libngsld_b200.so and declare every symbol of include/ngsld_b200.h."""
    global _LIB
    if _LIB is not None:
        return _LIB
    path = lib_path()
    if not os.path.exists(path):
        raise ImportError(f"{path} not found: build it with `make -C ngsld_b200/csrc` (or __graft_entry__.build()); "
                          "ngsld_b200 has no CPU fallback")
    L = C.CDLL(path)
    vp, u64, dbl, i32 = C.c_void_p, C.c_uint64, C.c_double, C.c_int
    pd = np.ctypeslib.ndpointer(np.float64, flags="C_CONTIGUOUS")
    pu32 = np.ctypeslib.ndpointer(np.uint32, flags="C_CONTIGUOUS")
    pu64 = np.ctypeslib.ndpointer(np.uint64, flags="C_CONTIGUOUS")
    sig = {
        "ngsld_abi_version": (i32, []),
        "ngsld_device_count": (i32, []),
        "ngsld_create": (i32, [C.POINTER(vp), i32]),
        "ngsld_destroy": (None, [vp]),
        "ngsld_last_error": (C.c_char_p, [vp]),
        "ngsld_set_stream": (i32, [vp, vp]),
        "ngsld_set_chunk_rows": (i32, [vp, u64]),
        "ngsld_prepare_sites": (i32, [pd, u64, u64, i32, i32, i32, i32, dbl, dbl, i32, pd, pd, pd]),
        "ngsld_set_sites": (i32, [vp, pd, pd, pd, u64, u64]),
        "ngsld_set_sites_raw": (i32, [vp, pd, u64, u64, i32, i32, i32, i32, dbl, dbl, vp]),
        "ngsld_set_positions": (i32, [vp, vp, vp]),
        "ngsld_scan_defaults": (None, [C.POINTER(ScanParams)]),
        "ngsld_scan_count": (i32, [vp, u64, u64, C.POINTER(ScanParams), C.POINTER(u64)]),
        "ngsld_partition": (i32, [vp, C.POINTER(ScanParams), i32, pu64]),
        "ngsld_scan": (i32, [vp, u64, u64, C.POINTER(ScanParams), ROW_SINK, vp]),
        "ngsld_scan_into": (i32, [vp, u64, u64, C.POINTER(ScanParams), vp, u64, C.POINTER(u64)]),
        "ngsld_scan_tsv": (i32, [vp, u64, u64, C.POINTER(ScanParams), TEXT_SINK, vp]),
        "ngsld_scan_device": (i32, [vp, u64, u64, C.POINTER(ScanParams)]),
        "ngsld_scan_tsv_into": (i32, [vp, u64, u64, C.POINTER(ScanParams), vp, u64, C.POINTER(u64), C.POINTER(u64)]),
        "ngsld_tsv_row_bound": (u64, [vp, i32]),
        "ngsld_tsv_row_bound_for": (u64, [C.c_uint32, i32]),
        "ngsld_alloc_host": (i32, [C.POINTER(vp), C.c_size_t]),
        "ngsld_free_host": (None, [vp]),
        "ngsld_share_sites": (i32, [vp, vp]),
        "ngsld_get_stats": (i32, [vp, C.POINTER(ScanStats)]),
        "ngsld_scan_decay": (i32, [vp, u64, u64, C.POINTER(ScanParams), dbl, u64, vp, C.POINTER(u64)]),
        "ngsld_pairs": (i32, [vp, pu32, pu32, u64, i32, i32, vp]),
        "ngsld_scan_edges": (i32, [vp, u64, u64, C.POINTER(ScanParams), C.POINTER(PruneParams), EDGE_SINK, vp, vp]),
        "ngsld_prune_graph": (i32, [u64, vp, vp, vp, u64, i32, vp, vp, C.POINTER(u64)]),
        "ngsld_site_seeds": (i32, [u64, u64, pu64]),
        "ngsld_plan_count": (i32, [pd, vp, u64, C.POINTER(ScanParams), u64, u64, C.POINTER(u64)]),
        "ngsld_plan_partition": (i32, [pd, vp, u64, C.POINTER(ScanParams), i32, pu64]),
        "ngsld_load_geno": (i32, [C.c_char_p, i32, i32, i32, u64, u64, pd, C.POINTER(i32)]),
        "ngsld_load_positions": (i32, [C.c_char_p, i32, u64, pd, C.POINTER(vp), C.POINTER(u64)]),
        "ngsld_free": (None, [vp]),
        "ngsld_tsv_header": (i32, [i32, C.c_char_p, C.c_size_t]),
        "ngsld_probe_fp64": (i32, [vp, C.POINTER(dbl)]),
    }
    for name, (res, args) in sig.items():
        fn = getattr(L, name)  # AttributeError here = header and library out of sync
        fn.restype = res
        fn.argtypes = args
    _LIB = L
    return L


EXPORTED = ["ngsld_abi_version", "ngsld_device_count", "ngsld_create", "ngsld_destroy", "ngsld_last_error", "ngsld_set_stream",
            "ngsld_set_chunk_rows", "ngsld_prepare_sites", "ngsld_set_sites", "ngsld_set_positions",
            "ngsld_scan_defaults", "ngsld_scan_count", "ngsld_partition", "ngsld_scan", "ngsld_scan_into",
            "ngsld_scan_tsv", "ngsld_scan_device", "ngsld_get_stats", "ngsld_pairs", "ngsld_site_seeds",
            "ngsld_tsv_header", "ngsld_probe_fp64", "ngsld_plan_count", "ngsld_plan_partition", "ngsld_load_geno",
            "ngsld_load_positions", "ngsld_free", "ngsld_scan_decay", "ngsld_scan_tsv_into", "ngsld_tsv_row_bound",
            "ngsld_alloc_host", "ngsld_free_host", "ngsld_share_sites", "ngsld_set_sites_raw", "ngsld_scan_edges", "ngsld_prune_graph", "ngsld_tsv_row_bound_for"]


def prepare_sites(raw, log_scale=False, from_log_cells=False, ignore_miss_data=False, call_geno=False,
                  N_thresh=0.0, call_thresh=0.0, n_threads=0):
    """ngsld_prepare_sites: raw [n_sites, n_ind, 3] -> (gl, expg, maf) (host, glibc, bit-identical to the
    reference's read_geno / est_maf / conv_space path)."""
    raw = np.ascontiguousarray(raw, np.float64)
    n_sites, n_ind, three = raw.shape
    assert three == 3
    gl = np.empty_like(raw)
    expg = np.empty((n_sites, n_ind))
    maf = np.empty(n_sites)
    rc = load_library().ngsld_prepare_sites(raw, n_sites, n_ind, int(log_scale), int(from_log_cells),
                                            int(ignore_miss_data), int(call_geno), N_thresh, call_thresh,
                                            n_threads, gl, expg, maf)
    if rc != 0:
        raise NgsldError(rc, "NaN found! Is the file format correct?" if rc == -4 else "invalid arguments")
    return gl, expg, maf


def site_seeds(seed, n_sites):
    out = np.empty(n_sites, np.uint64)
    load_library().ngsld_site_seeds(seed, n_sites, out)
    return out


def tsv_header(extend_out):
    buf = C.create_string_buffer(512)
    n = load_library().ngsld_tsv_header(int(extend_out), buf, 512)
    return buf.raw[:n]


def _global_error():
    return (load_library().ngsld_last_error(None) or b"").decode()


def load_geno(path, n_ind, n_sites, is_bin=None, probs=True, log_scale=False):
    """ngsld_load_geno: genotype file -> (cells [n_sites, n_ind, 3], log_cells flag).  is_bin defaults to the
    reference's rule: everything not ending in ".gz" is binary (ngsLD.cpp:45-57)."""
    if is_bin is None:
        is_bin = not path.endswith(".gz")
    cells = np.empty((n_sites, n_ind, 3))
    lc = C.c_int(0)
    rc = load_library().ngsld_load_geno(path.encode(), int(is_bin), int(probs or is_bin), int(log_scale), n_ind,
                                        n_sites, cells, C.byref(lc))
    if rc != 0:
        raise NgsldError(rc, _global_error())
    return cells, bool(lc.value)


def read_positions(path, n_sites, header=False):
    """ngsld_load_positions: --pos / --posH file -> (labels, pos_dist) with the reference's rules (ngsLD.cpp:119-132,
    shared/read_data.cpp:165-218): gz or plain; empty and '#' lines skipped; only the first tab of a line becomes
    ':'; +inf at a chromosome change; adjacent distance must be >= 1."""
    dist = np.empty(n_sites)
    blob, nbytes = C.c_void_p(), C.c_uint64(0)
    L = load_library()
    rc = L.ngsld_load_positions(path.encode(), int(header), n_sites, dist, C.byref(blob), C.byref(nbytes))
    if rc != 0:
        raise NgsldError(rc, _global_error())
    raw = C.string_at(blob, nbytes.value)
    L.ngsld_free(blob)
    labels = [x.decode() for x in raw.split(b"\0")[:n_sites]]
    return labels, dist


def prune_graph(n_sites, labels, seen, edges, keep_heavy=False):
    """ngsld_prune_graph (host only): (kept [n_sites] uint8: 1 kept / 0 excluded / 2 in no row, excluded sites in order)."""
    seen = np.ascontiguousarray(seen, np.uint8)
    edges = np.ascontiguousarray(edges, EDGE_DTYPE)
    kept = np.zeros(n_sites, np.uint8)
    excl = np.zeros(n_sites, np.uint32)
    n_ex = C.c_uint64(0)
    lab_ptr = None
    if labels is not None:
        arr = (C.c_char_p * n_sites)(*[l.encode() if isinstance(l, str) else l for l in labels])
        lab_ptr = C.cast(arr, C.c_void_p)
    rc = load_library().ngsld_prune_graph(n_sites, lab_ptr, seen.ctypes.data_as(C.c_void_p), edges.ctypes.data_as(C.c_void_p),
                                          len(edges), int(keep_heavy), kept.ctypes.data_as(C.c_void_p),
                                          excl.ctypes.data_as(C.c_void_p), C.byref(n_ex))
    if rc != 0:
        raise NgsldError(rc, "invalid pruning input")
    return kept, excl[:n_ex.value].copy()


def plan_count(maf, pos_dist, params, s1_lo=0, s1_hi=None):
    """ngsld_plan_count: rows first sites [s1_lo, s1_hi) will produce, computed without a device."""
    maf = np.ascontiguousarray(maf, np.float64)
    n = C.c_uint64(0)
    pd_ptr = None
    if pos_dist is not None:
        pos_dist = np.ascontiguousarray(pos_dist, np.float64)
        pd_ptr = pos_dist.ctypes.data_as(C.c_void_p)
    rc = load_library().ngsld_plan_count(maf, pd_ptr, len(maf), C.byref(params), s1_lo,
                                         len(maf) if s1_hi is None else s1_hi, C.byref(n))
    if rc != 0:
        raise NgsldError(rc, _global_error())
    return n.value


def plan_partition(maf, pos_dist, params, n_parts):
    """ngsld_plan_partition: equal-row-count first-site ranges for n_parts workers, without a device."""
    maf = np.ascontiguousarray(maf, np.float64)
    b = np.zeros(n_parts + 1, np.uint64)
    pd_ptr = None
    if pos_dist is not None:
        pos_dist = np.ascontiguousarray(pos_dist, np.float64)
        pd_ptr = pos_dist.ctypes.data_as(C.c_void_p)
    rc = load_library().ngsld_plan_partition(maf, pd_ptr, len(maf), C.byref(params), n_parts, b)
    if rc != 0:
        raise NgsldError(rc, _global_error())
    return b


class Engine:
    """One ngsld_ctx bound to one GPU."""

    def __init__(self, device=0):
        self._lib = load_library()
        h = C.c_void_p()
        rc = self._lib.ngsld_create(C.byref(h), device)
        if rc != 0:
            raise NgsldError(rc, self._lib.ngsld_last_error(None).decode())
        self._h = h
        self.n_sites = self.n_ind = 0
        self._keep = []

    def close(self):
        if getattr(self, "_h", None):
            self._lib.ngsld_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    def _check(self, rc):
        if rc != 0:
            raise NgsldError(rc, self._lib.ngsld_last_error(self._h).decode())

    def set_stream(self, cuda_stream_ptr):
        self._check(self._lib.ngsld_set_stream(self._h, C.c_void_p(cuda_stream_ptr)))

    def set_chunk_rows(self, rows):
        self._check(self._lib.ngsld_set_chunk_rows(self._h, rows))

    def set_sites(self, gl, expg, maf):
        gl = np.ascontiguousarray(gl, np.float64)
        expg = np.ascontiguousarray(expg, np.float64)
        maf = np.ascontiguousarray(maf, np.float64)
        n_sites, n_ind = expg.shape
        assert gl.shape == (n_sites, n_ind, 3) and maf.shape == (n_sites,)
        self._check(self._lib.ngsld_set_sites(self._h, gl, expg, maf, n_sites, n_ind))
        self.n_sites, self.n_ind = n_sites, n_ind

    def set_sites_raw(self, raw, log_scale=False, from_log_cells=False, ignore_miss_data=False, call_geno=False,
                      N_thresh=0.0, call_thresh=0.0):
        """ngsld_set_sites_raw: file cells -> device, prepared there (opt-in; not bit-identical to the host path).
        Returns the allele frequencies."""
        raw = np.ascontiguousarray(raw, np.float64)
        n_sites, n_ind, three = raw.shape
        assert three == 3
        maf = np.empty(n_sites)
        self._check(self._lib.ngsld_set_sites_raw(self._h, raw, n_sites, n_ind, int(log_scale), int(from_log_cells),
                                                  int(ignore_miss_data), int(call_geno), N_thresh, call_thresh,
                                                  maf.ctypes.data_as(C.c_void_p)))
        self.n_sites, self.n_ind = n_sites, n_ind
        return maf

    def set_positions(self, pos_dist=None, labels=None):
        pd_ptr = None
        if pos_dist is not None:
            pos_dist = np.ascontiguousarray(pos_dist, np.float64)
            assert pos_dist.shape == (self.n_sites,)
            pd_ptr = pos_dist.ctypes.data_as(C.c_void_p)
        lab_ptr = None
        if labels is not None:
            assert len(labels) == self.n_sites
            arr = (C.c_char_p * self.n_sites)(*[l.encode() if isinstance(l, str) else l for l in labels])
            lab_ptr = C.cast(arr, C.c_void_p)
        self._check(self._lib.ngsld_set_positions(self._h, pd_ptr, lab_ptr))

    def count(self, params, s1_lo=0, s1_hi=None):
        n = C.c_uint64(0)
        hi = self.n_sites if s1_hi is None else s1_hi
        self._check(self._lib.ngsld_scan_count(self._h, s1_lo, hi, C.byref(params), C.byref(n)))
        return n.value

    def partition(self, params, n_parts):
        b = np.zeros(n_parts + 1, np.uint64)
        self._check(self._lib.ngsld_partition(self._h, C.byref(params), n_parts, b))
        return b

    def scan(self, params, s1_lo=0, s1_hi=None):
        """All rows of first sites [s1_lo, s1_hi) as a structured array (ngsld_scan_into)."""
        hi = self.n_sites if s1_hi is None else s1_hi
        n = self.count(params, s1_lo, hi)
        out = np.zeros(n, ROW_DTYPE)
        got = C.c_uint64(0)
        self._check(self._lib.ngsld_scan_into(self._h, s1_lo, hi, C.byref(params), out.ctypes.data_as(C.c_void_p), n,
                                              C.byref(got)))
        assert got.value == n
        return out

    def scan_sink(self, params, fn, s1_lo=0, s1_hi=None):
        """ngsld_scan with a Python callback fn(rows_view) -> None (rows_view is only valid inside the call)."""
        hi = self.n_sites if s1_hi is None else s1_hi

        def _cb(user, ptr, n):
            buf = (C.c_char * (n * ROW_DTYPE.itemsize)).from_address(ptr)
            fn(np.frombuffer(buf, ROW_DTYPE, count=n))
            return 0
        cb = ROW_SINK(_cb)
        self._check(self._lib.ngsld_scan(self._h, s1_lo, hi, C.byref(params), cb, None))

    def scan_tsv(self, params, s1_lo=0, s1_hi=None, header=True, out=None):
        """TSV text (device-formatted).  Returns bytes, or writes to the binary file object `out`."""
        hi = self.n_sites if s1_hi is None else s1_hi
        chunks = []

        def _cb(user, ptr, n_bytes, n_rows):
            b = C.string_at(ptr, n_bytes)
            if out is not None:
                out.write(b)
            else:
                chunks.append(b)
            return 0
        cb = TEXT_SINK(_cb)
        head = tsv_header(params.extend_out) if header else b""
        if out is not None:
            out.write(head)
        self._check(self._lib.ngsld_scan_tsv(self._h, s1_lo, hi, C.byref(params), cb, None))
        return None if out is not None else head + b"".join(chunks)

    def scan_tsv_into(self, params, s1_lo=0, s1_hi=None, pinned=True):
        """ngsld_scan_tsv_into: the scan's TSV text through one (page-locked) buffer sized with ngsld_tsv_row_bound."""
        hi = self.n_sites if s1_hi is None else s1_hi
        cap = max(1, self.count(params, s1_lo, hi)) * self._lib.ngsld_tsv_row_bound(self._h, params.extend_out)
        buf = C.c_void_p()
        if pinned:
            rc = self._lib.ngsld_alloc_host(C.byref(buf), cap)
            if rc != 0:
                raise NgsldError(rc, _global_error())
            ptr = buf
        else:
            keep = C.create_string_buffer(cap)
            ptr = C.cast(keep, C.c_void_p)
        try:
            nb, nr = C.c_uint64(0), C.c_uint64(0)
            self._check(self._lib.ngsld_scan_tsv_into(self._h, s1_lo, hi, C.byref(params), ptr, cap, C.byref(nb), C.byref(nr)))
            return C.string_at(ptr, nb.value), nr.value
        finally:
            if pinned:
                self._lib.ngsld_free_host(buf)

    def share_sites_from(self, other):
        """ngsld_share_sites: take over the site table of another Engine (another GPU) by device-to-device copy."""
        self._check(self._lib.ngsld_share_sites(self._h, other._h))
        self.n_sites, self.n_ind = other.n_sites, other.n_ind

    def scan_edges(self, params, prune, s1_lo=0, s1_hi=None):
        """ngsld_scan_edges: (edges as EDGE_DTYPE array in (s1, s2) order, seen [n_sites] uint8)."""
        hi = self.n_sites if s1_hi is None else s1_hi
        parts = []

        def _cb(user, ptr, n):
            buf = (C.c_char * (n * EDGE_DTYPE.itemsize)).from_address(ptr)
            parts.append(np.frombuffer(buf, EDGE_DTYPE, count=n).copy())
            return 0
        cb = EDGE_SINK(_cb)
        seen = np.zeros(self.n_sites, np.uint8)
        self._check(self._lib.ngsld_scan_edges(self._h, s1_lo, hi, C.byref(params), C.byref(prune), cb, None,
                                               seen.ctypes.data_as(C.c_void_p)))
        return (np.concatenate(parts) if parts else np.zeros(0, EDGE_DTYPE)), seen

    def scan_device(self, params, s1_lo=0, s1_hi=None):
        hi = self.n_sites if s1_hi is None else s1_hi
        self._check(self._lib.ngsld_scan_device(self._h, s1_lo, hi, C.byref(params)))
        return self.stats()

    def scan_decay(self, params, bin_size, n_bins, s1_lo=0, s1_hi=None):
        """ngsld_scan_decay: per-distance-bin counts and sums of r2_ExpG, D, Dp, r2 accumulated on the device.
        Returns (bins structured array with fields n[4], sum[4], rows that fell outside every bin)."""
        hi = self.n_sites if s1_hi is None else s1_hi
        bins = np.zeros(n_bins, DECAY_DTYPE)
        outside = C.c_uint64(0)
        self._check(self._lib.ngsld_scan_decay(self._h, s1_lo, hi, C.byref(params), float(bin_size), n_bins,
                                               bins.ctypes.data_as(C.c_void_p), C.byref(outside)))
        return bins, outside.value

    def stats(self):
        st = ScanStats()
        self._check(self._lib.ngsld_get_stats(self._h, C.byref(st)))
        return st.as_dict()

    def pairs(self, s1, s2, ignore_miss_data=False, strict=False):
        s1 = np.ascontiguousarray(s1, np.uint32)
        s2 = np.ascontiguousarray(s2, np.uint32)
        out = np.zeros(len(s1), ROW_DTYPE)
        self._check(self._lib.ngsld_pairs(self._h, s1, s2, len(s1), int(ignore_miss_data), int(strict),
                                          out.ctypes.data_as(C.c_void_p)))
        return out

    def probe_fp64(self):
        g = C.c_double(0)
        self._check(self._lib.ngsld_probe_fp64(self._h, C.byref(g)))
        return g.value
