"""ngsld_b200 — B200-native drop-in for the pairwise-LD path of fgvieira/ngsLD.

The product is libngsld_b200.so (hand-written sm_100a kernels behind the C ABI in
include/ngsld_b200.h) plus the ngsLD-compatible CLI bin/ngsLD; this package is the thin Python mirror of
that ABI used by tests and bench.py.  There is no CPU fallback: importing works anywhere, but
`Engine()` raises if the library or a CUDA device is missing."""
from .api import (DECAY_DTYPE, EDGE_DTYPE, EXPORTED, Engine, PruneParams, prune_graph, NgsldError, ROW_DTYPE, ScanParams, lib_path, load_geno, load_library,  # noqa: F401
                  plan_count, plan_partition, prepare_sites, read_positions, site_seeds, tsv_header)

__all__ = ["DECAY_DTYPE", "EDGE_DTYPE", "EXPORTED", "Engine", "PruneParams", "prune_graph", "NgsldError", "ROW_DTYPE", "ScanParams", "lib_path", "load_geno", "load_library",
           "plan_count", "plan_partition", "prepare_sites", "read_positions", "site_seeds", "tsv_header"]
