"""Static checks on the shipped machine code (no GPU needed): the occupancy the kernels were tuned for depends on their
register counts, and a local-memory spill in one of the hot loops would cost more than any of the measured optimisations
gained.  Numbers are the ones DESIGN.md §3 / profiles/r2_sass_census.txt state."""
import os
import re
import shutil
import subprocess

import pytest

import helpers as H

LIB = os.path.join(H.ROOT, "ngsld_b200", "libngsld_b200.so")


def _res_usage():
    if shutil.which("cuobjdump") is None or not os.path.exists(LIB):
        pytest.skip("cuobjdump or the built library is not available")
    out = subprocess.run(["cuobjdump", "-res-usage", LIB], capture_output=True, text=True).stdout
    res, name = {}, None
    for ln in out.splitlines():
        m = re.search(r"Function (\S+):", ln)
        if m:
            name = m.group(1)
            continue
        m = re.search(r"REG:(\d+) STACK:(\d+) SHARED:(\d+) LOCAL:(\d+)", ln)
        if m and name:
            res[name] = tuple(int(x) for x in m.groups())
    return res


def _one(res, pattern):
    hits = [v for k, v in res.items() if re.search(pattern, k)]
    assert len(hits) == 1, (pattern, len(hits))
    return hits[0]


def test_library_holds_only_sm_100a_code():
    if shutil.which("cuobjdump") is None or not os.path.exists(LIB):
        pytest.skip("cuobjdump or the built library is not available")
    out = subprocess.run(["cuobjdump", "-lelf", LIB], capture_output=True, text=True).stdout
    archs = set(re.findall(r"\.(sm_\w+)\.cubin", out))
    assert archs == {"sm_100a"}, archs


def test_cell_kernels_keep_their_register_budget_and_do_not_spill():
    res = _res_usage()
    # six register levels: 128 registers -> 4 CTAs of 4 warps per SM; four levels: 96 registers -> 5 CTAs
    reg, stack, shared, local = _one(res, r"em_cell_kernelILi6ELb1ELi4EEE")
    assert reg <= 128 and stack == 0 and local == 0, (reg, stack, local)
    assert shared <= 8192  # static part (per-pair totals of a batch, work statistics); the tables are dynamic
    reg, stack, shared, local = _one(res, r"em_cell_kernelILi4ELb1ELi5EEE")
    assert reg <= 102 and local == 0 and stack <= 16, (reg, stack, local)
    # the dense warp-per-pair kernel behind it: 152 registers leave room for the r2_ExpG CTA beside three EM CTAs
    reg, stack, shared, local = _one(res, r"em_warp_kernelILi6ELb0ELi2ELi1EEE")
    assert reg <= 152 and stack == 0 and local == 0, (reg, stack, local)


def test_x87_loop_of_the_fused_kernel_is_the_limb_version():
    """The r2_ExpG loop of the fused kernel: blocks of four individuals fetched with 128-bit loads, no spills."""
    if shutil.which("cuobjdump") is None or not os.path.exists(LIB):
        pytest.skip("cuobjdump or the built library is not available")
    out = subprocess.run(["cuobjdump", "-sass", "-fun",
                          "_ZN6emcell14em_cell_kernelILi6ELb1ELi4EEEv9SiteTable9PairChunkNS_8CellArgsEP11DevCounters", LIB],
                         capture_output=True, text=True).stdout
    assert out.count("LDG.E.128") >= 8            # significand blocks (and the ratio table)
    assert "LDL" not in out and "STL" not in out  # nothing lives in local memory
    assert out.count("MATCH.ANY") == 4            # joint classes: the four matches of a block of individuals
    n_instr = len(re.findall(r"^\s+/\*[0-9a-f]{4,5}\*/", out, flags=re.M))
    assert 3500 < n_instr < 6500, n_instr
