"""The N>1 path on CPU (gloo, world_size 2): every rank derives the same equal-pair-count first-site ranges from the
device-free planner (ngsld_plan_partition), owns one range, and the shards concatenated in rank order are the
single-process output -- no data-path collective.  The per-rank shard is produced by the oracle here (there is no
GPU in this container and the product has no CPU path); on the GPU box tests/test_gpu_parity.py checks the same
concatenation property with the CUDA scan.  Also covers bench.py's cross-rank reduction."""
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import helpers as H

CASES = [("kb20", dict(max_kb_dist=20)), ("ext", dict(max_kb_dist=0)), ("rnd01", dict(max_kb_dist=0, rnd_sample=0.01, seed=1))]


def _worker(rank, world, port, tmp, results):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    sys.path.insert(0, H.ROOT)
    sys.path.insert(0, H.HERE)
    import bench
    import ngsld_b200 as N
    from oracle import oracle as O
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        fx = H.MANIFEST["fixtures"]["s"]
        raw, labels, pdist, _ = H.load_fixture("s", tmp, fx["variants"]["ext"]["flags"], True)
        gl, expg, maf = N.prepare_sites(raw)
        for variant, kw in CASES:
            v = fx["variants"][variant]
            opt = H.parse_flags(v["flags"])
            P = N.ScanParams.make(**kw)
            bounds = N.plan_partition(maf, pdist, P, world).astype(np.int64)
            all_bounds = [None] * world
            dist.all_gather_object(all_bounds, bounds.tolist())
            assert all(b == all_bounds[0] for b in all_bounds), "ranks disagree on the partition"
            lo, hi = int(bounds[rank]), int(bounds[rank + 1])
            mine = N.plan_count(maf, pdist, P, lo, hi)
            shard = os.path.join(tmp, f"shard{rank}.{variant}.ld")
            n, _ = O.run(gl, expg, maf, pdist, labels, opt["max_kb_dist"], opt["max_snp_dist"], opt["min_maf"],
                         opt["rnd_sample"], opt["seed"], opt["ignore_miss"], opt["extend_out"], lo, hi,
                         out_path=shard, header=(rank == 0))
            assert n == mine
            mx, sm = bench.reduce_over_ranks([float(mine), float(rank + 1)], "cpu", world)
            assert sm[0] == v["rows"] and mx[1] == world and sm[1] == world * (world + 1) / 2
            shards = [None] * world
            dist.all_gather_object(shards, open(shard, "rb").read())
            if rank == 0:
                whole = b"".join(shards)
                assert H.md5(whole) == v["md5"], variant
                counts = [s.count(b"\n") for s in shards]
                assert abs((counts[0] - 1) - counts[1]) <= 2 * fx["n_sites"]
        results[rank] = "ok"
    finally:
        dist.destroy_process_group()


def test_two_ranks_partition_and_concatenate(tmp_path):
    H.fixture_paths("s", tmp_path)  # materialise once, before the fork
    mgr = mp.Manager()
    results = mgr.dict()
    port = 29500 + os.getpid() % 2000
    mp.spawn(_worker, args=(2, port, str(tmp_path), results), nprocs=2, join=True)
    assert dict(results) == {0: "ok", 1: "ok"}


def test_bench_slabs_tile_the_rank_range():
    sys.path.insert(0, H.ROOT)
    import bench
    n = 5000
    for lo, hi in ((0, n), (1200, 3100)):
        slabs = bench.slab_bounds(n, lo, hi, 400_000)
        assert slabs[0][0] == lo and slabs[-1][1] <= hi
        for (a, b, c), (a2, _, _) in zip(slabs, slabs[1:]):
            assert b == a2
        for a, b, c in slabs:
            assert c == sum(n - 1 - s for s in range(a, b))
        assert all(abs(c - 400_000) < n for _, _, c in slabs[:-1])
