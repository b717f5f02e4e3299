#!/usr/bin/env python
"""Summarise an .ncu-rep (one kernel launch) into the handful of numbers DESIGN.md / profiles/ quote.
usage: python scripts/ncu_summary.py gpurun_out/x.ncu-rep [--source]  (needs `ncu` on PATH; no GPU needed)"""
import csv
import io
import subprocess
import sys
from collections import Counter

KEYS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "launch__registers_per_thread",
        "launch__grid_size", "launch__block_size", "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed",
        "lts__t_sector_hit_rate.pct", "smsp__inst_executed.sum", "smsp__average_warp_latency_per_inst_issued.ratio",
        "sm__cycles_active.avg", "launch__shared_mem_per_block_dynamic"]


def raw(rep):
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    return rows[0], rows[1], rows[2:]


def main():
    rep = sys.argv[1]
    hdr, units, launches = raw(rep)
    for vals in launches:
        d = dict(zip(hdr, zip(units, vals)))
        print("kernel:", d.get("Kernel Name", ("", "?"))[1])
        for k in KEYS:
            if k in d:
                print(f"  {k:78s} {d[k][1]:>16s} {d[k][0]}")
        st = {h.split("smsp__average_warps_issue_stalled_")[1].split("_per_issue_active")[0]: float(v[1])
              for h, v in d.items() if h.startswith("smsp__average_warps_issue_stalled_") and v[1] not in ("", "n/a")}
        print("  warp stall cycles per issued instruction:",
              ", ".join(f"{k} {v:.2f}" for k, v in sorted(st.items(), key=lambda x: -x[1])[:8]))
    if "--source" in sys.argv:
        out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
        rows = list(csv.reader(io.StringIO(out)))
        hdr = rows[1]
        ix = {h: i for i, h in enumerate(hdr)}
        inst, samp = Counter(), Counter()
        for r in rows[2:]:
            toks = [t for t in r[ix["Source"]].split() if not t.startswith("@")]
            op = toks[0].split(".")[0] if toks else "?"
            inst[op] += int(r[ix["Instructions Executed"]])
            samp[op] += int(r[ix["# Samples"]])
        ti, ts = sum(inst.values()), sum(samp.values())
        print("  executed instruction mix (warp-level) and share of stall samples:")
        for op, n in inst.most_common(14):
            print(f"    {op:10s} {100 * n / ti:6.2f}% of instructions  {100 * samp[op] / max(ts, 1):6.2f}% of samples")


if __name__ == "__main__":
    main()
