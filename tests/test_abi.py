"""CPU-side checks of the drop-in boundary: libngsld_b200.so loads, exports every symbol that
include/ngsld_b200.h declares (and nothing undeclared), refuses to run without a GPU instead of falling
back to the CPU, and its host-only entry points agree with the oracle."""
import ctypes as C
import os
import re
import subprocess

import numpy as np
import pytest

import helpers as H
import ngsld_b200 as N
from helpers import O

HEADER = os.path.join(H.ROOT, "include", "ngsld_b200.h")


def declared_symbols():
    src = open(HEADER).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(ngsld_[a-z0-9_]+)\s*\(", src)) - {"ngsld_row_sink", "ngsld_text_sink"})


def test_library_exports_exactly_the_header():
    lib = N.load_library()
    decl = declared_symbols()
    assert len(decl) >= 25
    for name in decl:
        assert hasattr(lib, name), f"{name} declared in the header but not exported"
    assert sorted(N.EXPORTED) == decl, "python mirror and header disagree"
    out = subprocess.run(["nm", "-D", "--defined-only", N.lib_path()], capture_output=True, text=True).stdout
    exported = sorted(set(re.findall(r"\bT (ngsld_[a-z0-9_]+)$", out, flags=re.M)))
    assert exported == decl, "library exports symbols the header does not declare (or misses some)"


def test_no_torch_or_cpp_types_in_the_abi():
    src = open(HEADER).read()
    assert "torch" not in src and "std::" not in src and "at::" not in src
    assert 'extern "C"' in src


def test_abi_version_and_row_layout():
    assert N.load_library().ngsld_abi_version() == 2
    assert N.ROW_DTYPE.itemsize == 112


@pytest.mark.skipif(os.path.exists("/dev/nvidia0"), reason="box has a GPU")
def test_create_fails_loudly_without_a_gpu():
    """No CPU fallback: without a device the product path refuses to run."""
    with pytest.raises(N.NgsldError) as ei:
        N.Engine(0)
    assert ei.value.code == -2 and "no CPU fallback" in str(ei.value)
    assert N.load_library().ngsld_device_count() == 0


def test_product_package_never_imports_the_oracle():
    for root, _, files in os.walk(os.path.join(H.ROOT, "ngsld_b200")):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".cpp", ".h")):
                txt = open(os.path.join(root, f), errors="ignore").read()
                assert "liboracle" not in txt and "from oracle" not in txt and "import oracle" not in txt, f


def test_prepare_sites_matches_oracle_bitwise():
    GL, _ = H.gen_synth.synth(60, 33, 21)
    GL[3, 4] = [0.0, 0.0, 0.0]
    GL[5, :] = [1 / 3] * 3
    GL[7, 2] = [1.0, 0.0, 0.0]
    for kw in (dict(), dict(ignore_miss_data=True), dict(call_geno=True),
               dict(call_geno=True, N_thresh=0.3, call_thresh=0.9)):
        a = N.prepare_sites(GL, n_threads=3, **kw)
        b = O.preprocess(GL, False, kw.get("ignore_miss_data", False), kw.get("call_geno", False),
                         kw.get("N_thresh", 0.0), kw.get("call_thresh", 0.0))
        for x, y in zip(a, b):
            assert x.tobytes() == y.tobytes(), kw
    a = N.prepare_sites(np.log(GL[8:20]), log_scale=True)
    b = O.preprocess(np.log(GL[8:20]), True)
    for x, y in zip(a, b):
        assert x.tobytes() == y.tobytes()


def test_prepare_sites_rejects_nan():
    raw = np.full((2, 2, 3), 0.25)
    raw[1, 1, 0] = np.nan
    with pytest.raises(N.NgsldError) as ei:
        N.prepare_sites(raw)
    assert ei.value.code == -4


def test_site_seeds_match_oracle():
    for seed in (0, 1, 12345, 2 ** 31 - 1):
        want = np.empty(50, np.uint64)
        O.lib().orc_site_seeds(seed, 50, want)
        assert np.array_equal(N.site_seeds(seed, 50), want)


def test_header_line():
    assert N.tsv_header(False) == b"site1\tsite2\tdist\tr2_ExpG\tD\tDp\tr2\n"
    assert N.tsv_header(True).endswith(b"\tchi2\tloglike\tnIter\n") and N.tsv_header(True).count(b"\t") == 18
