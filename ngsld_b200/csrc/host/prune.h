// Greedy graph pruning of the reference's downstream script scripts/prune_graph.pl (sub prune_graph_idx), on an edge
// list the device filtered out of the pair scan (aux::prune_edges_kernel).  Host side: O(E log N).
#pragma once
#include <stdint.h>

#include <vector>

#include "../../../include/ngsld_b200.h"

namespace prune {

// seen[s] != 0: site s occurs in at least one row of the scan (the scripts add both nodes of every input line).
// name_rank[s]: rank of the site's label in case-insensitive string order (ties broken by site index).
// kept[s]: 1 = in the pruned set (unlinked from the start, or left after pruning), 0 = excluded, 2 = in no row.
void run(uint64_t n_sites, const uint8_t *seen, const uint32_t *name_rank, const ngsld_edge *edges, uint64_t n_edges,
         bool keep_heavy, uint8_t *kept, std::vector<uint32_t> &excluded);

void label_ranks(uint64_t n_sites, const char *const *labels, std::vector<uint32_t> &rank);

}  // namespace prune
