"""TEST INFRASTRUCTURE -- restatement of the reference's downstream pruning script, used only by tests/ as the checker
of ngsld_scan_edges / ngsld_prune_graph.  Never imported by the product.

Follows /root/reference/scripts/prune_graph.pl line by line (reading loop :103-142, unlinked nodes :163-172,
prune_graph_idx :246-312, remove_node_idx :316-330) on the TSV TEXT the main program writes, with plain dictionaries
instead of Graph::Easy.  PARITY UNPINNED: the script itself cannot run in this image (Graph::Easy, List::AllUtils,
Scalar::Util::Numeric are not installed; prune_ngsLD.py needs graph-tool), so this restatement is the definition the
tests hold the product to.  Two behaviours of the script are deliberately not restated: it dies on a non-integer
distance ("inf" across chromosomes, :113-116) -- here such a row is simply no edge -- and the order in which it prints
the surviving nodes (hash order)."""
import math


def read_edges(tsv_text, max_kb_dist=float("inf"), min_weight=0.0, field_dist=3, field_weight=7, weight_type="a",
               weight_precision=4, header=True):
    """-> (nodes in order of first appearance, {(a, b): label})   [prune_graph.pl:103-142]"""
    max_dist = max_kb_dist * 1000
    prec = 10 ** weight_precision
    nodes, edges = {}, {}
    lines = tsv_text.decode().splitlines() if isinstance(tsv_text, bytes) else tsv_text.splitlines()
    for ln in lines[1 if header else 0:]:
        f = ln.split("\t")
        nodes.setdefault(f[0], len(nodes))
        nodes.setdefault(f[1], len(nodes))
        try:
            weight = float(f[field_weight - 1])
        except ValueError:
            continue
        if math.isnan(weight) or math.isinf(weight):      # Math::BigFloat is_nan / is_inf -> next
            continue
        try:
            dist = int(f[field_dist - 1])
        except ValueError:                                 # "inf": the script dies here; treated as no edge
            continue
        if dist > max_dist:
            continue
        if weight_type == "a":
            weight = abs(weight)
        if weight < min_weight:
            continue
        if weight_type == "n":
            weight = 1
        edges[(f[0], f[1])] = int(weight * prec)           # perl int(): truncation towards zero
    return list(nodes), edges


def prune(nodes, edges, keep_heavy=False):
    """-> (kept node set, excluded nodes in order)   [prune_graph.pl:163-172, 246-330]"""
    adj = {n: {} for n in nodes}
    for (a, b), w in edges.items():
        adj[a][b] = w
        adj[b][a] = w
    kept = {n for n in nodes if not adj[n]}                # unlinked nodes are printed first and removed
    idx = {n: sum(adj[n].values()) for n in nodes if adj[n]}
    excl = []

    def remove(v):
        for u, w in adj[v].items():
            if u in idx:
                idx[u] -= w
                del adj[u][v]
        adj[v] = {}
        del idx[v]
        excl.append(v)
    while idx:
        heavy = sorted(idx, key=lambda n: (-idx[n], n.lower()))[0]
        if idx[heavy] <= 0:
            break
        if keep_heavy:
            for child in list(adj[heavy]):
                remove(child)
        else:
            remove(heavy)
    return kept | set(idx), excl
