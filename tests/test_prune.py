"""LD pruning: ngsld_prune_graph (host, no GPU) and ngsld_scan_edges (device filter, -m gpu) against
the restatement of scripts/prune_graph.pl in oracle/prune_oracle.py (test infrastructure; parity unpinned: the script's
Perl modules are not installed here)."""
import numpy as np
import pytest

import helpers as H
import ngsld_b200 as N
from oracle import prune_oracle as PO


def random_graph(rng, n, m, max_label=9000, negative=False):
    pairs = set()
    while len(pairs) < m:
        a, b = rng.integers(0, n, 2)
        if a != b:
            pairs.add((min(a, b), max(a, b)))
    e = np.zeros(len(pairs), N.EDGE_DTYPE)
    for k, (a, b) in enumerate(sorted(pairs)):
        e[k] = (a, b, rng.integers(-max_label if negative else 0, max_label))
    return e


@pytest.mark.parametrize("seed,n,m,keep_heavy,negative", [(1, 30, 60, False, False), (2, 200, 1500, False, False),
                                                          (3, 200, 900, True, False), (4, 80, 300, False, True),
                                                          (5, 500, 499, False, False), (6, 50, 0, False, False)])
def test_prune_graph_equals_the_script_restatement(seed, n, m, keep_heavy, negative):
    rng = np.random.default_rng(seed)
    edges = random_graph(rng, n, m, negative=negative)
    labels = [f"chr{rng.integers(1, 4)}:{rng.integers(1, 10**6)}_{k}" for k in range(n)]
    labels[3] = labels[3].upper()                               # ties are broken case-insensitively
    seen = np.ones(n, np.uint8)
    seen[n - 1] = 0 if not np.any((edges["s1"] == n - 1) | (edges["s2"] == n - 1)) else 1
    kept, excl = N.prune_graph(n, labels, seen, edges, keep_heavy)
    nodes = [labels[s] for s in range(n) if seen[s]]
    want_kept, want_excl = PO.prune(nodes, {(labels[a], labels[b]): int(w) for a, b, w in edges}, keep_heavy)
    assert {labels[s] for s in range(n) if kept[s] == 1} == want_kept
    assert sorted(labels[s] for s in excl) == sorted(want_excl)
    if not keep_heavy:                                           # removal order is fully determined by the tie rule
        assert [labels[s] for s in excl] == want_excl
    assert np.all(kept[seen == 0] == 2)
    # no edge survives between kept nodes (unless every remaining weight is <= 0, where the script stops too)
    alive = kept == 1
    left = edges[alive[edges["s1"]] & alive[edges["s2"]]]
    assert not negative and np.all(left["label"] <= 0) or negative


@pytest.mark.gpu
@pytest.mark.parametrize("variant,q", [("kb20", dict(max_dist=10000.0, min_weight=0.2)),
                                       ("ext", dict(max_dist=float("inf"), min_weight=0.5, field=6)),
                                       ("ext", dict(max_dist=50000.0, min_weight=0.05, field=5, weight_type="e")),
                                       ("kb20", dict(max_dist=20000.0, min_weight=0.3, weight_type="n", field=4))])
def test_device_edge_filter_and_pruning_equal_the_script_on_the_tsv(variant, q, tmp_path_factory):
    """The edges ngsld_scan_edges delivers are exactly the rows the script keeps when it reads the TSV text (weights as
    printed, six decimals), and pruning them gives the script's site list."""
    import gpu_helpers as G
    tmp = tmp_path_factory.getbasetemp()
    v = H.MANIFEST["fixtures"]["s"]["variants"][variant]
    raw, labels, dist, opt = H.load_fixture("s", tmp, v["flags"], True)
    eng, _ = G.engine_for(raw, opt, labels, dist)
    P = G.scan_params(opt, True)
    Q = N.PruneParams.make(**q)
    with eng:
        tsv = eng.scan_tsv(P)
        edges, seen = eng.scan_edges(P, Q)
        fast_edges, _ = eng.scan_edges(G.scan_params(opt, False), Q)
    nodes, want = PO.read_edges(tsv, max_kb_dist=q["max_dist"] / 1000, min_weight=q["min_weight"], field_weight=q.get("field", 7),
                                weight_type=q.get("weight_type", "a"))
    got = {(labels[a], labels[b]): int(w) for a, b, w in edges}
    assert got == want and len(want) > 50
    assert {labels[s] for s in range(len(labels)) if seen[s]} == set(nodes)
    kept, excl = N.prune_graph(len(labels), labels, seen, edges)
    want_kept, want_excl = PO.prune(nodes, want)
    assert {labels[s] for s in range(len(labels)) if kept[s] == 1} == want_kept
    assert [labels[s] for s in excl] == want_excl
    # the fast kernel sees the same graph up to weights that sit on a printing boundary
    assert abs(len(fast_edges) - len(edges)) <= 2
