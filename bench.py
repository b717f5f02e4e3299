#!/usr/bin/env python
"""bench.py — pairs/s of the pairwise-LD hot path on B200 (driver contract; see DESIGN.md "Measurement").

A *step* is one pass of the hot path over one batch: all pairs of a contiguous slab of first sites
(≈ --batch-pairs pairs) of the 50 000-SNP x 500-individual all-pairs workload (BASELINE.json configs[2],
the configuration the metric and the north-star target are quoted on; it fits one GPU).  Batches of
successive steps are different slabs, and the genotype-likelihood matrix (600 MB) is larger than L2.

  value    — pairs/s with the site table already resident in HBM and results left in HBM
             (ngsld_scan_device), CUDA events on the library's stream, max over ranks.
  e2e      — the same batches through the C ABI with HOST buffers: ngsld_set_sites (H2D of the prepared
             site arrays) + ngsld_scan (binary rows delivered to a host sink), copies inside the timed region.
  roofline — dominant kernel (the fast EM kernel): algorithmic bytes B_alg * pairs / summed kernel time
             against the measured HBM peak, plus the FP64 issue fraction that actually bounds the path.
  cpu_baseline — the unmodified reference binary (oracle/_ref/ngsLD) on all host cores, on a bounded
             prefix sample of the same workload (same n_ind).

Multi-GPU (torchrun, one rank per GPU): the upper triangle is partitioned into equal-pair-count first-site
ranges (ngsld_partition); each rank runs its own batches from its own range; no data-path collective.
`--impl reference` times only the reference CPU implementation (rank 0)."""
import argparse
import json
import os
import statistics
import subprocess
import sys
import tempfile
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests", "golden"))

CHUNK_ROWS = 4 << 20  # rows per device chunk (the library's default): one EM-kernel launch per chunk
METRIC = "snp_pairs_per_sec"
UNIT = "pairs/s"


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--n-sites", type=int, default=50000)
    ap.add_argument("--n-ind", type=int, default=500)
    ap.add_argument("--seed", type=int, default=11)
    ap.add_argument("--batch-pairs", type=int, default=0, help="pairs per step and GPU (0 = auto)")
    ap.add_argument("--cpu-seconds", type=float, default=15.0, help="target CPU time of the reference sample")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-strong", action="store_true", help="skip the whole-workload time-to-solution legs")
    ap.add_argument("--strong-out", default="", help="where the CLI leg writes its TSV (default: /dev/shm if it has room, "
                                                     "else /dev/null)")
    ap.add_argument("--strict", action="store_true", help="bit-faithful EM kernel instead of the fast one")
    ap.add_argument("--em-path", default="", choices=["", "cell", "warp", "list", "tile"],
                    help="force an EM kernel family (default: the library's choice; 'warp' = dense warp-per-pair kernel)")
    return ap.parse_args()


def bytes_per_pair(n_ind):
    """both GL rows + both expected-genotype rows + 7 result doubles."""
    return 2 * n_ind * 24 + 2 * n_ind * 8 + 56


def workload_name(a):
    return f"synthetic binary GL, {a.n_sites} SNPs x {a.n_ind} ind, all pairs (--max_kb_dist 0)"


def make_inputs(a):
    import gen_synth
    GL, pos = gen_synth.synth_fast(a.n_sites, a.n_ind, a.seed)
    return GL, pos


# ------------------------------------------------------------------------------------------------
# reference CPU implementation on a bounded sample
def ref_sample_run(GL, pos, n_sub, threads, tmpdir):
    """Time the reference CLI (or, if it did not travel, the oracle port) on the first n_sub sites."""
    from oracle import oracle as O  # test infrastructure: the CPU baseline leg is allowed to execute it
    n_ind = GL.shape[1]
    n_pairs = n_sub * (n_sub - 1) // 2
    if O.have_ref():
        g = os.path.join(tmpdir, f"sample_{n_sub}.glf")
        if not os.path.exists(g):
            GL[:n_sub].astype("<f8").tofile(g)
            with open(g + ".pos", "w") as fh:
                fh.write("".join(f"chr1\t{p}\n" for p in pos[:n_sub]))
        t0 = time.perf_counter()
        O.run_ref(["--geno", g, "--probs", "--n_ind", str(n_ind), "--n_sites", str(n_sub), "--pos", g + ".pos",
                   "--max_kb_dist", "0"], "/dev/null", n_threads=threads)
        return n_pairs, time.perf_counter() - t0, "reference"
    gl, expg, maf = O.preprocess(GL[:n_sub])
    t0 = time.perf_counter()
    n, _ = O.run(gl, expg, maf, None, None, extend_out=False, n_threads=threads, out_path="/dev/null")
    return n, time.perf_counter() - t0, "port"


def ref_sample_mean_passes(GL, pos, n_sub, threads, tmpdir):
    """Mean EM passes per pair of the reference on the sample (untimed extra run with --extend_out: the nIter column is
    the 0-based index of the converging pass, 100 = all 100 passes ran without converging)."""
    from oracle import oracle as O
    if not O.have_ref():
        return None
    g = os.path.join(tmpdir, f"sample_{n_sub}.glf")
    out = os.path.join(tmpdir, "sample.ld")
    O.run_ref(["--geno", g, "--probs", "--n_ind", str(GL.shape[1]), "--n_sites", str(n_sub), "--pos", g + ".pos",
               "--max_kb_dist", "0", "--extend_out"], out, n_threads=threads)
    tot = n = 0
    with open(out, "rb") as fh:
        next(fh)
        for ln in fh:
            it = int(ln[ln.rfind(b"\t") + 1:])
            tot += it + 1 if it < 100 else 100
            n += 1
    os.unlink(out)
    return tot / max(1, n)


def pick_sample(GL, pos, threads, target_s, tmpdir):
    """Calibrate on a tiny prefix, then size the sample for ~target_s seconds of CPU wall time."""
    n_sub = 160
    for _ in range(4):
        n, dt, _ = ref_sample_run(GL, pos, n_sub, threads, tmpdir)
        if dt >= target_s / 4 or n_sub >= GL.shape[0]:
            break
        rate = n / max(dt - 0.02, 1e-3)  # minus process start-up
        nxt = int(min(GL.shape[0], (2 * rate * target_s) ** 0.5))
        n_sub = max(n_sub + 1, min(nxt, n_sub * 4))
    n, dt, _ = ref_sample_run(GL, pos, n_sub, threads, tmpdir) if dt < target_s / 4 else (n, dt, None)
    return int(min(GL.shape[0], max(160, n_sub * (target_s / max(dt, 1e-3)) ** 0.5)))


def reference_arm(a):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    threads = os.cpu_count() or 1
    GL, pos = make_inputs(argparse.Namespace(n_sites=min(a.n_sites, 4096), n_ind=a.n_ind, seed=a.seed))
    with tempfile.TemporaryDirectory() as td:
        n_sub = pick_sample(GL, pos, threads, max(2.0, a.cpu_seconds / 2), td)
        times, n_pairs, kind = [], 0, "reference"
        for k in range(a.warmup + a.steps):
            n_pairs, dt, kind = ref_sample_run(GL, pos, n_sub, threads, td)
            if k >= a.warmup:
                times.append(dt)
        mean_passes = ref_sample_mean_passes(GL, pos, n_sub, threads, td)
    total = sum(times)
    v = n_pairs * len(times) / total
    sample = f"first {n_sub} sites of the workload, all pairs = {n_pairs} pairs per step, --n_threads {threads}"
    line = {"impl": "reference", "metric": METRIC, "value": v, "unit": UNIT, "n_gpus": a.gpus, "steps": a.steps,
            "warmup": a.warmup, "ms_per_step": 1e3 * total / len(times), "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": workload_name(a), "n_sites": a.n_sites, "n_ind": a.n_ind},
            "cpu_baseline": {"value": v, "unit": UNIT, "cores": threads, "kind": kind, "sample": sample,
                             "mean_em_passes_per_pair": mean_passes,
                             "pair_ind_passes_per_s": v * a.n_ind * mean_passes if mean_passes else None},
            "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------------
class ClockSampler:
    """nvidia-smi clocks/throttle reasons of one GPU during the timed region (recipe in B200_PROFILING.md)."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.rows, self.proc = [], None
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(index), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.th = threading.Thread(target=self._read, daemon=True)
            self.th.start()
        except OSError:
            self.proc = None

    def _read(self):
        for ln in self.proc.stdout:
            self.rows.append((time.perf_counter(), [x.strip() for x in ln.split(",")]))

    def stop(self, t0, t1):
        if not self.proc:
            return None
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except subprocess.TimeoutExpired:
            self.proc.kill()
        rows = [r for t, r in self.rows if t0 <= t <= t1 and len(r) >= 7] or [r for _, r in self.rows if len(r) >= 7]
        if not rows:
            return None
        sm = [float(r[0]) for r in rows]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [n for k, n in enumerate(names) if any(r[3 + k].lower().startswith("active") for r in rows)]
        pw = [float(r[2]) for r in rows if r[2].replace(".", "", 1).isdigit()]
        return {"sm_mhz": statistics.median(sm), "sm_max_mhz": float(rows[0][1]), "reasons": reasons,
                "samples": len(rows), "power_w_median": statistics.median(pw) if pw else None}


def slab_bounds(n_sites, lo, hi, batch_pairs):
    """First-site slabs [a, b) inside [lo, hi) of about batch_pairs pairs each (all-pairs triangle)."""
    per = (n_sites - 1 - np.arange(lo, hi, dtype=np.int64))
    cum = np.concatenate([[0], np.cumsum(per)])
    slabs, a = [], 0
    while a < hi - lo and cum[-1] - cum[a] > 0:
        b = int(np.searchsorted(cum, cum[a] + batch_pairs, side="left"))
        b = max(a + 1, min(b, hi - lo))
        if cum[b] - cum[a] > 0:
            slabs.append((lo + a, lo + b, int(cum[b] - cum[a])))
        a = b
    return slabs


def measured_hbm_peak(path):
    """(GB/s, where it came from): the driver-written MEASURED_PEAKS.json if present (key `hbm_gbs`, looked for at any
    nesting depth), else the fallback B200_PROFILING.md states."""
    def find(obj):
        if isinstance(obj, dict):
            for k, v in obj.items():
                if isinstance(v, (int, float)) and "hbm" in k.lower() and "gb" in k.lower() and v > 0:
                    return float(v)
            for v in obj.values():
                r = find(v)
                if r:
                    return r
        elif isinstance(obj, list):
            for v in obj:
                r = find(v)
                if r:
                    return r
        return None
    try:
        v = find(json.load(open(path)))
        if v:
            return v, "measured (MEASURED_PEAKS.json hbm_gbs)"
    except (OSError, ValueError):
        pass
    return 6650.0, "fallback (B200_PROFILING.md)"


def reduce_over_ranks(values, device, world):
    """(max over ranks, sum over ranks) of a vector of per-rank numbers; the backend is whatever the process
    group was initialised with (nccl on the GPU box, gloo in the CPU tests)."""
    import torch
    import torch.distributed as dist
    vec = torch.tensor(values, device=device, dtype=torch.float64)
    mx, sm = vec.clone(), vec.clone()
    if world > 1:
        dist.all_reduce(mx, op=dist.ReduceOp.MAX)
        dist.all_reduce(sm, op=dist.ReduceOp.SUM)
    return mx.tolist(), sm.tolist()


def strong_legs(a, eng, P, N, GL, pos, bounds, rank, world, local, barrier, reduce_over_ranks):
    """Whole-workload time to solution (all n(n-1)/2 pairs once, split over the ranks = GPUs):
      rows_to_host_sinks  every rank scans its equal-pair-count first-site range through ngsld_scan, rows delivered to a
                          host sink; wall time between two barriers (= the slowest rank).
      tsv_cli             the drop-in CLI (ngsld_b200/bin/ngsLD --gpu_n <ranks>) on the same input file, process start to
                          exit, TSV written to --strong-out (rank 0 runs it; the other ranks wait on the CPU)."""
    import shutil
    lo, hi = int(bounds[rank]), int(bounds[rank + 1])
    seen = [0]

    def sink(rows):
        seen[0] += len(rows)
    flag = f"/dev/shm/ngsld_bench_{os.environ.get('MASTER_PORT', '0')}_{a.seed}"
    if rank == 0 and os.path.exists(flag + ".done"):  # left over from a run that died
        os.unlink(flag + ".done")
    barrier()
    t0 = time.perf_counter()
    eng.scan_sink(P, sink, lo, hi)
    mine = time.perf_counter() - t0
    st = eng.stats()
    barrier()
    wall = time.perf_counter() - t0
    mx, sm = reduce_over_ranks([wall, mine, float(seen[0]), st["ms_em"]], "cuda", world)
    mn = [-x for x in reduce_over_ranks([-mine], "cuda", world)[0]]
    n_total = a.n_sites * (a.n_sites - 1) // 2
    out = {"workload": f"all {n_total} pairs of the workload once, {world} GPU(s)",
           "rows_to_host_sinks": {"seconds": mx[0], "pairs": int(sm[2]), "pairs_per_s": sm[2] / mx[0],
                                  "slowest_rank_s": mx[1], "fastest_rank_s": mn[0],
                                  "imbalance": mx[1] / (sm[1] / world), "d2h_bytes": int(sm[2]) * 112,
                                  "api": "ngsld_scan(row sink) on each rank's ngsld_partition range"}}
    # ---- the CLI on the same data ----
    cli = os.path.join(ROOT, "ngsld_b200", "bin", "ngsLD")
    if rank == 0 and os.path.exists(cli):
        geno = flag + ".glf"
        GL.astype("<f8").tofile(geno)
        with open(geno + ".pos", "w") as fh:
            fh.write("".join(f"chr1\t{p}\n" for p in pos))
        target = a.strong_out
        if not target:
            target = flag + ".ld" if shutil.disk_usage("/dev/shm").free > 130e9 else "/dev/null"
        base = [cli, "--geno", geno, "--probs", "--n_ind", str(a.n_ind), "--n_sites", str(a.n_sites), "--pos", geno + ".pos",
                "--max_kb_dist", "0", "--gpu_n", str(world), "--gpu_stats", "--verbose", "0", "--out", target]

        def cli_leg(extra, what):
            import glob
            t1 = time.perf_counter()
            r = subprocess.run(base + extra, capture_output=True, text=True)
            dt = time.perf_counter() - t1
            files = [target] if target != "/dev/null" and os.path.exists(target) else []
            if extra and target != "/dev/null":
                files = sorted(glob.glob(target + ".part-*"))
            size = sum(os.path.getsize(f) for f in files) if files else None
            for f in files:
                os.unlink(f)
            err = r.stderr.splitlines()
            return {"seconds": dt, "returncode": r.returncode, "pairs_per_s": n_total / dt if r.returncode == 0 else None,
                    "out": target if target == "/dev/null" else "/dev/shm (tmpfs), " + what, "tsv_bytes": size, "files": len(files),
                    "time_line": next((l for l in err if l.startswith("[time]")), None),
                    "writers": [l for l in err if l.startswith("[writer")],
                    "command": "ngsLD --geno <600 MB binary GL> --probs --n_ind 500 --n_sites 50000 --pos <pos> --max_kb_dist 0 "
                               f"--gpu_n {world} --out <out> {' '.join(extra)}: process start to exit"}
        out["tsv_cli"] = cli_leg([], "one file (the reference's --out)")
        out["tsv_cli_shards"] = cli_leg(["--gpu_out_shards"], "one file per slab of first sites, cat in name order = the output")
        for f in (geno, geno + ".pos"):
            os.unlink(f)
        open(flag + ".done", "w").close()
    elif os.path.exists(cli):
        while not os.path.exists(flag + ".done"):  # wait on the CPU: a pending NCCL barrier would occupy SMs
            time.sleep(0.2)
    barrier()
    if rank == 0 and os.path.exists(flag + ".done"):
        os.unlink(flag + ".done")
    return out


def main():
    a = parse()
    if a.impl == "reference":
        return reference_arm(a)

    import torch
    import torch.distributed as dist
    import ngsld_b200 as N
    if a.em_path:
        os.environ["NGSLD_EM_PATH"] = a.em_path

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: ngsld_b200 has no CPU fallback")
    torch.cuda.set_device(local)
    if world > 1:
        # stdout carries exactly one JSON line: NCCL prints its version banner to stdout while the communicator is
        # created, so file descriptor 1 points at stderr until the first collective has completed
        sys.stdout.flush()
        saved_fd = os.dup(1)
        os.dup2(2, 1)
        try:
            dist.init_process_group("nccl", device_id=torch.device("cuda", local))
            dist.barrier()
            torch.cuda.synchronize()
        finally:
            sys.stdout.flush()
            os.dup2(saved_fd, 1)
            os.close(saved_fd)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    GL, pos = make_inputs(a)
    gl, expg, maf = N.prepare_sites(GL)

    def pinned(x):  # host buffers of the end-to-end leg live in pinned memory (driver contract)
        t = torch.empty(x.shape, dtype=torch.float64, pin_memory=True)
        t.numpy()[...] = x
        return t
    pin_keep = [pinned(gl), pinned(expg), pinned(maf)]
    gl, expg, maf = (t.numpy() for t in pin_keep)
    pos_dist = np.diff(np.concatenate([[0], pos])).astype(np.float64)
    eng = N.Engine(local)
    stream = torch.cuda.Stream()
    eng.set_stream(stream.cuda_stream)
    eng.set_sites(gl, expg, maf)
    eng.set_positions(pos_dist, None)
    P = N.ScanParams.make(max_kb_dist=0, strict=int(a.strict))
    bounds = eng.partition(P, world)
    lo, hi = int(bounds[rank]), int(bounds[rank + 1])
    fp64_peak = eng.probe_fp64()  # GFLOP/s, measured DFMA issue rate of this GPU

    # batch size: auto = ~1.5 s of device time per step, found with one calibration slab
    batch = a.batch_pairs
    if batch <= 0:
        cal = slab_bounds(a.n_sites, lo, hi, 2_000_000)[0]
        eng.scan_device(P, cal[0], cal[1])
        st = eng.scan_device(P, cal[0], cal[1])
        rate = st["n_pairs"] / (st["ms_device_total"] * 1e-3)
        batch = int(min(max(rate * (0.25 if a.strict else 1.5), 2_000_000), 96_000_000))
        if world > 1:
            t = torch.tensor([batch], device="cuda", dtype=torch.int64)
            dist.all_reduce(t, op=dist.ReduceOp.MIN)
            batch = int(t.item())
    slabs = slab_bounds(a.n_sites, lo, hi, batch)
    slabs = [s for s in slabs if s[2] >= batch // 2] or slabs
    n_steps = a.warmup + a.steps

    def slab(k):
        return slabs[k % len(slabs)]

    # ---------------- leg 1: HBM-resident ----------------
    for k in range(a.warmup):
        eng.scan_device(P, *slab(k)[:2])
    sampler = ClockSampler(local)
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    t_wall0 = time.perf_counter()
    ev0.record(stream)
    pairs = launches = passes = n_chunks = 0
    cell_pairs = cells = cell_passes = resid_pairs = 0
    ms_em = ms_pearson = 0.0
    em_kernel = ""
    for k in range(a.warmup, n_steps):
        st = eng.scan_device(P, *slab(k)[:2])
        em_kernel = st["em_kernel"]
        pairs += st["n_pairs"]
        launches += st["n_launches"]
        passes += st["sum_em_passes"]
        ms_em += st["ms_em"]
        ms_pearson += st["ms_pearson"]
        n_chunks += -(-st["n_pairs"] // CHUNK_ROWS)
        cell_pairs += st["n_cell_pairs"]
        cells += st["sum_cells"]
        cell_passes += st["sum_cell_passes"]
        resid_pairs += st["n_resid_pairs"]
    ev1.record(stream)
    barrier()
    t_wall1 = time.perf_counter()
    clocks = sampler.stop(t_wall0, t_wall1)
    ms_dev = ev0.elapsed_time(ev1)

    # ---------------- leg 2: end to end through the C ABI with host buffers ----------------
    e2e = None
    if not a.no_e2e:
        seen = [0]

        def sink(rows):
            seen[0] += len(rows)

        def e2e_step(k):
            eng.set_sites(gl, expg, maf)       # H2D of the prepared site arrays (host buffers)
            eng.set_positions(pos_dist, None)
            eng.scan_sink(P, sink, *slab(k)[:2])  # rows D2H into pinned chunks, handed to the sink
            return eng.stats()

        e2e_step(0)
        f0, f1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        barrier()
        seen[0] = 0
        f0.record(stream)
        h2d = d2h = 0
        e_steps = max(1, min(a.steps, 3))
        for k in range(a.warmup, a.warmup + e_steps):
            st = e2e_step(k)
            d2h += st["d2h_bytes"]
            h2d += st["h2d_bytes"] + gl.nbytes + expg.nbytes + maf.nbytes  # plan arrays + the site table
        f1.record(stream)
        barrier()
        ms_e2e = f0.elapsed_time(f1)
        e2e = {"pairs": seen[0], "ms": ms_e2e, "h2d": h2d / e_steps, "d2h": d2h / e_steps, "steps": e_steps}

    # ---------------- leg 3: time to solution for the WHOLE workload (strong scaling over the ranks) ----------------
    strong = None
    if not a.no_strong:
        strong = strong_legs(a, eng, P, N, GL, pos, bounds, rank, world, local, barrier, reduce_over_ranks)

    # ---------------- reduce over ranks ----------------
    mx, sm = reduce_over_ranks([ms_dev, float(pairs), float(launches), float(passes), ms_em, ms_pearson,
                                e2e["ms"] if e2e else 0.0, float(e2e["pairs"]) if e2e else 0.0], "cuda", world)

    if rank == 0:
        ms_max, tot_pairs = mx[0], sm[1]
        value = tot_pairs / (ms_max * 1e-3)
        bpp = bytes_per_pair(a.n_ind)
        # dominant kernel = EM; its per-launch duration from the CUDA events the library records around it
        # on its own launch stream (rank 0's numbers)
        em_s = ms_em * 1e-3
        achieved = pairs * bpp / em_s / 1e9
        hbm_peak, peak_src = measured_hbm_peak(os.path.join(ROOT, "MEASURED_PEAKS.json"))
        # FP64 work that actually bounds the kernel.  Dense kernels: per (individual, EM pass) 9 DMUL + 18 DFMA (+ 1 MUFU
        # seed) = 27 FP64 instructions = 45 flop [em_warp.cuh header; SURVEY.md §8(d) rounds this to 40].  Class-compressed
        # kernel: the same E-step once per CELL (distinct likelihood combination of the pair) plus one DMUL for the
        # cell's weight = 28 instructions = 46 flop per (cell, pass); pairs it left to the dense kernel count as dense.
        n_chunks = max(1, n_chunks)
        cell_mode = cell_pairs > 0
        if cell_mode:
            dense_passes = passes * (resid_pairs / max(1, pairs))  # left-over pairs: assume the mean pass count
            fp64_instr = cell_passes * 28.0 + dense_passes * a.n_ind * 27.0
            flop = cell_passes * 46.0 + dense_passes * a.n_ind * 45.0
        else:
            fp64_instr = passes * a.n_ind * 27.0
            flop = passes * a.n_ind * 45.0
        fp64_achieved = flop / em_s / 1e9
        issue = fp64_instr / em_s / (fp64_peak * 1e9 / 2.0) if fp64_peak else None
        # DRAM traffic of that kernel per launch, from the committed ncu --set full capture (bytes per pair x pairs/launch)
        traffic, traffic_src = None, None
        for name in ("r2_traffic.json", "r1_traffic.json"):
            tj = os.path.join(ROOT, "profiles", name)
            if os.path.exists(tj) and traffic is None:
                t = json.load(open(tj))
                if t.get("kernel") == em_kernel and a.n_ind == t.get("n_ind", 500):
                    traffic = t["dram_bytes_per_pair"] * pairs / n_chunks
                    traffic_src = t["source"]
        roofline = {"bound": "hbm", "achieved": achieved, "peak": hbm_peak, "unit": "GB/s", "frac": achieved / hbm_peak,
                    "traffic": traffic, "traffic_source": traffic_src,
                    "algorithmic_bytes_per_launch": bpp * pairs / n_chunks, "peak_source": peak_src, "kernel": em_kernel,
                    "algorithmic_bytes_per_pair": bpp, "launch_ms_avg": ms_em / n_chunks,
                    "launches_timed": n_chunks,
                    "note": "path is FP64-issue-bound (~%.0f EM passes/pair), not HBM-bound; see fp64" % (passes / max(1, pairs)),
                    "fp64": {"achieved_gflops": fp64_achieved, "peak_gflops": fp64_peak,
                             "frac": fp64_achieved / fp64_peak if fp64_peak else None,
                             "issue_frac": issue,
                             "peak_source": "ngsld_probe_fp64 (DFMA issue micro-benchmark with 2 register operands, this GPU, "
                                            "this run); DFMAs reading 3 distinct registers issue at 0.73 of it "
                                            "(scripts/micro/fp64_ops.cu)",
                             "flop_per_ind_pass": 45, "fp64_instr_per_ind_pass": 27,
                             "flop_per_cell_pass": 46, "fp64_instr_per_cell_pass": 28,
                             "mean_em_passes_per_pair": passes / max(1, pairs)}}
        if cell_mode:
            roofline["cells"] = {"pairs_on_cell_kernel": cell_pairs, "pairs_left_to_dense_kernel": resid_pairs,
                                 "mean_cells_per_pair": cells / max(1, cell_pairs), "individuals": a.n_ind,
                                 "note": "class-compressed EM: one weighted E-step per distinct (p, q) likelihood "
                                         "combination of a pair instead of one per individual; r2_ExpG (x87 emulation, "
                                         "integer pipes) runs inside the same kernel"}
        line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": a.steps, "warmup": a.warmup,
                "ms_per_step": ms_max / a.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
                "dtype": "f64", "data": "synthetic",
                "config": {"workload": workload_name(a), "n_sites": a.n_sites, "n_ind": a.n_ind,
                           "pairs_per_step_per_gpu": int(pairs // a.steps), "kernel": "strict" if a.strict else "fast",
                           "partition": f"equal-pair-count first-site ranges, {world} part(s), no collective",
                           "l2": "inputs larger than L2 (site table %.0f MB, a different slab each step)" % (gl.nbytes / 1e6)},
                "gpu_launches": int(sm[2]), "roofline": roofline, "clocks": clocks}
        if e2e:
            e_value = sm[7] / (mx[6] * 1e-3)
            line["e2e"] = {"value": e_value, "unit": UNIT, "h2d_bytes_per_step": int(e2e["h2d"]),
                           "d2h_bytes_per_step": int(e2e["d2h"]), "steps": e2e["steps"],
                           "api": "ngsld_set_sites + ngsld_set_positions + ngsld_scan(row sink) per step"}
        if strong:
            line["strong"] = strong
        if world == 1 and not a.no_cpu_baseline:
            threads = os.cpu_count() or 1
            with tempfile.TemporaryDirectory() as td:
                n_sub = pick_sample(GL, pos, threads, a.cpu_seconds, td)
                n, dt, kind = ref_sample_run(GL, pos, n_sub, threads, td)
            line["cpu_baseline"] = {"value": n / dt, "unit": UNIT, "cores": threads, "kind": kind,
                                    "sample": f"first {n_sub} sites of the workload, all pairs = {n} pairs, "
                                              f"--n_threads {threads}, {dt:.1f} s"}
        print(json.dumps(line), flush=True)
    eng.close()
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
