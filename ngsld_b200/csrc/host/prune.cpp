// See prune.h.  Follows scripts/prune_graph.pl:246-330 (prune_graph_idx / remove_node_idx): node weight = sum of the
// integer labels of its edges; repeatedly take the node with the largest weight (ties: case-insensitive name order),
// stop when that weight is <= 0, and delete it (default) or its neighbours (--keep_heavy), updating the weights of the
// nodes that lose an edge.
#include "prune.h"

#include <ctype.h>
#include <string.h>

#include <algorithm>
#include <numeric>
#include <set>
#include <string>

namespace prune {

void label_ranks(uint64_t n_sites, const char *const *labels, std::vector<uint32_t> &rank) {
  rank.resize(n_sites);
  std::vector<uint32_t> order(n_sites);
  std::iota(order.begin(), order.end(), 0u);
  if (labels) {
    std::vector<std::string> lc(n_sites);
    for (uint64_t s = 0; s < n_sites; s++) {
      lc[s] = labels[s] ? labels[s] : "";
      for (auto &ch : lc[s]) ch = (char)tolower((unsigned char)ch);
    }
    std::stable_sort(order.begin(), order.end(), [&](uint32_t a, uint32_t b) { return lc[a] < lc[b]; });
  }
  for (uint64_t k = 0; k < n_sites; k++) rank[order[k]] = (uint32_t)k;
}

void run(uint64_t n_sites, const uint8_t *seen, const uint32_t *name_rank, const ngsld_edge *edges, uint64_t n_edges,
         bool keep_heavy, uint8_t *kept, std::vector<uint32_t> &excluded) {
  excluded.clear();
  // adjacency in compressed rows
  std::vector<uint64_t> off(n_sites + 1, 0);
  for (uint64_t e = 0; e < n_edges; e++) {
    off[edges[e].s1 + 1]++;
    off[edges[e].s2 + 1]++;
  }
  for (uint64_t s = 0; s < n_sites; s++) off[s + 1] += off[s];
  std::vector<uint32_t> nbr(2 * n_edges);
  std::vector<int32_t> lab(2 * n_edges);
  {
    std::vector<uint64_t> fill(off.begin(), off.end() - 1);
    for (uint64_t e = 0; e < n_edges; e++) {
      const ngsld_edge &x = edges[e];
      nbr[fill[x.s1]] = x.s2; lab[fill[x.s1]++] = x.label;
      nbr[fill[x.s2]] = x.s1; lab[fill[x.s2]++] = x.label;
    }
  }
  std::vector<long long> weight(n_sites, 0);
  std::vector<uint8_t> alive(n_sites, 0);
  typedef std::pair<long long, uint32_t> Key;  // (-weight, name rank): begin() = heaviest, then first by name
  std::set<Key> heap;
  std::vector<uint32_t> site_of_rank(n_sites);
  for (uint64_t s = 0; s < n_sites; s++) site_of_rank[name_rank[s]] = (uint32_t)s;
  for (uint64_t s = 0; s < n_sites; s++) {
    kept[s] = seen[s] ? 1 : 2;
    if (off[s + 1] == off[s]) continue;  // unlinked: printed first and dropped from the graph (prune_graph.pl:165-172)
    for (uint64_t k = off[s]; k < off[s + 1]; k++) weight[s] += lab[k];
    alive[s] = 1;
    heap.insert(Key(-weight[s], name_rank[s]));
  }
  auto remove_node = [&](uint32_t v) {
    heap.erase(Key(-weight[v], name_rank[v]));
    alive[v] = 0;
    for (uint64_t k = off[v]; k < off[v + 1]; k++) {
      const uint32_t u = nbr[k];
      if (!alive[u]) continue;
      heap.erase(Key(-weight[u], name_rank[u]));
      weight[u] -= lab[k];
      heap.insert(Key(-weight[u], name_rank[u]));
    }
    kept[v] = 0;
    excluded.push_back(v);
  };
  while (!heap.empty()) {
    const Key top = *heap.begin();
    if (-top.first <= 0) break;
    const uint32_t v = site_of_rank[top.second];
    if (keep_heavy) {
      for (uint64_t k = off[v]; k < off[v + 1]; k++)
        if (alive[nbr[k]]) remove_node(nbr[k]);
    } else {
      remove_node(v);
    }
  }
}

}  // namespace prune
