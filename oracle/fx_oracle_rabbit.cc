/*
 * fx_oracle_rabbit.cc -- TEST INFRASTRUCTURE ONLY (see fx_oracle.h).
 * Restatement of DataLoaderDFS (DataLoader.cu:324-385) and DataLoaderRabbit (DataLoader.cu:455-655).
 * C++ rather than C for one reason: the reference orders each Rabbit round with an UNSTABLE sort
 * (ranges::sort by current degree, DataLoader.cu:545-546); how equal degrees fall is a property of
 * libstdc++'s introsort, so the only faithful restatement calls the same std::sort on the same
 * sequence.  Structure follows the reference (per-vertex map of modularity weights, dendrogram of
 * Tree_Node, recursive leaf walk); the product (flex_b200/csrc/fx_order.cu) uses flat arrays.
 */
#include <algorithm>
#include <cstdint>
#include <map>
#include <vector>

#include "fx_oracle.h"

extern "C" void orc_order_dfs(int64_t n, const uint32_t* rowptr, const uint32_t* col, uint64_t* rank) {
  /* vo_to_dfs[0] = n marks vertex 0 as visited until the end (:344, :388) */
  std::vector<uint32_t> vo_to_dfs((size_t)n, 0);
  if (n == 0) return;
  vo_to_dfs[0] = (uint32_t)n;
  uint32_t placed = 0; /* rowPtr.size() - 1 in the reference */
  for (int64_t root = 0; root < n;) {
    std::vector<std::pair<uint32_t, uint32_t>> stack{{rowptr[root], rowptr[root + 1]}};
    if (root) vo_to_dfs[root] = placed;
    placed++;
    while (!stack.empty()) {
      auto& it = stack.back();
      while (it.first < it.second && vo_to_dfs[col[it.first]]) it.first++;
      if (it.first >= it.second) { stack.pop_back(); continue; }
      const uint32_t dst = col[it.first++];
      vo_to_dfs[dst] = placed++;
      stack.push_back({rowptr[dst], rowptr[dst + 1]});
    }
    if (placed >= (uint32_t)n) break;
    while (++root < n && vo_to_dfs[root]) {}
  }
  vo_to_dfs[0] = 0;
  for (int64_t i = 0; i < n; ++i) rank[i] = vo_to_dfs[i];
}

namespace {
struct Tree_Node {
  Tree_Node(Tree_Node* a, Tree_Node* b) : lchild(a), rchild(b), v_idx(-1) {}
  explicit Tree_Node(int v) : lchild(nullptr), rchild(nullptr), v_idx(v) {}
  Tree_Node() : lchild(nullptr), rchild(nullptr), v_idx(-1) {}
  Tree_Node *lchild, *rchild;
  int v_idx;
  void leaves_apply(std::vector<int>& perm) {
    if (lchild) { lchild->leaves_apply(perm); rchild->leaves_apply(perm); }
    else perm.push_back(v_idx);
  }
};
struct Vertex {
  std::map<int, int> dst_wht;
  Tree_Node leaf_node, cluster_node, *tree_node;
  int deg, round;
};
}  // namespace

/* vo_mp[new] = old.  Returns -1 if a round makes no progress (reference assert :585). */
extern "C" int orc_order_rabbit(int64_t n, const uint32_t* rowptr, const uint32_t* col, int is_directed,
                                int32_t* vo_mp) {
  std::vector<uint32_t> v_this_round((size_t)n), v_next_round;
  std::vector<Vertex> mgraph((size_t)n);
  long long n_edges = 0;
  for (int64_t v = 0; v < n; ++v) {
    Vertex& vo = mgraph[v];
    for (uint32_t e = rowptr[v]; e < rowptr[v + 1]; ++e) {
      const int d = (int)col[e];
      if (d != v) {
        vo.dst_wht[d] = 1;
        if (is_directed) mgraph[d].dst_wht[(int)v] = 1;
      }
    }
    vo.deg = (int)vo.dst_wht.size(); /* taken before later vertices add their reverse edges (:527) */
    n_edges += vo.deg;
    vo.leaf_node = Tree_Node((int)v);
    vo.tree_node = &vo.leaf_node;
    v_this_round[v] = (uint32_t)v;
    vo.round = 0;
  }
  const double two_m_inv = 1.0 / double(2 * n_edges);
  for (int round = 1; !v_this_round.empty(); round++) {
    std::sort(v_this_round.begin(), v_this_round.end(),
              [&](uint32_t a, uint32_t b) { return mgraph[a].deg < mgraph[b].deg; });
    for (uint32_t u : v_this_round) {
      Vertex& uo = mgraph[u];
      if (uo.round == round) continue;
      double dQ_max = -1;
      int v = -1;
      const double dv_2m = uo.deg * two_m_inv;
      for (auto [d, w] : uo.dst_wht) {
        const double dq = w - mgraph[d].deg * dv_2m;
        if (!(dq <= dQ_max)) { dQ_max = dq; v = d; } /* set_max common.h:96-102 */
      }
      if (dQ_max <= 0) continue;
      Vertex& vo = mgraph[v];
      vo.deg += uo.deg;
      for (auto [d, w] : uo.dst_wht) {
        if (d == v) continue;
        vo.dst_wht[d] += w;
        auto& dodw = mgraph[d].dst_wht;
        if (!dodw.count((int)u)) continue;
        dodw[v] += dodw[(int)u];
        dodw.erase((int)u);
      }
      vo.dst_wht.erase((int)u);
      uo.cluster_node = Tree_Node(vo.tree_node, uo.tree_node);
      uo.tree_node = nullptr;
      vo.tree_node = &uo.cluster_node;
      if (vo.round == round) continue;
      vo.round = round;
      v_next_round.push_back((uint32_t)v);
    }
    if (!(v_next_round.size() < v_this_round.size())) return -1;
    std::swap(v_this_round, v_next_round);
    v_next_round.clear();
  }
  std::vector<int> perm;
  perm.reserve((size_t)n);
  for (auto& vo : mgraph)
    if (vo.tree_node) vo.tree_node->leaves_apply(perm);
  for (int64_t i = 0; i < n; ++i) vo_mp[i] = perm[(size_t)i];
  return 0;
}
