// fx_order.cu -- DEG / RCM / Gorder rank computation for the reordering hooks.
//
// Behavioural contract (bit-exact permutations): order_deg.cu:19-44, order_rcm.cu:15-33,
// algo_bfs.cu:11-38, adjlist.cu:62-151, order_gorder.cu:13-143, unitheap.cu:16-217 of the
// reference (vendored there from lecfab/rescience-gorder).  Host code, as in the reference:
// Gorder is a serial priority-queue greedy and none of this is inside tPre/tElap.
// Data layout differs from the reference (flat CSR-style int arrays built by counting, no
// Edgelist of pairs, no std::function ranker); the algorithms' decisions do not.
#include <algorithm>
#include <climits>
#include <cmath>
#include <map>
#include <numeric>

#include "fx_common.cuh"

namespace fx {

typedef uint64_t ul;

// in+out degree over the CSR edges (Edgelist::compute_degrees edgelist.cu:89-104)
static void degrees(const fx_matrix* m, std::vector<ul>& out, std::vector<ul>& in) {
  const int64_t n = m->n;
  out.assign(n, 0); in.assign(n, 0);
  for (int64_t r = 0; r < n; ++r) {
    out[r] = m->rowptr[r + 1] - m->rowptr[r];
    for (uint32_t e = m->rowptr[r]; e < m->rowptr[r + 1]; ++e) in[m->col[e]]++;
  }
}

// rank_from_deg (order_deg.cu:19-39): position after sorting by degree (DESC or ASC), id ASC
int order_deg(const fx_matrix* m, bool desc, std::vector<ul>& rank) {
  const int64_t n = m->n;
  std::vector<ul> dout, din;
  degrees(m, dout, din);
  std::vector<uint32_t> ids(n);
  std::iota(ids.begin(), ids.end(), 0u);
  std::vector<ul> deg(n);
  for (int64_t i = 0; i < n; ++i) deg[i] = dout[i] + din[i];
  if (desc) std::stable_sort(ids.begin(), ids.end(), [&](uint32_t a, uint32_t b) { return deg[a] > deg[b]; });
  else std::stable_sort(ids.begin(), ids.end(), [&](uint32_t a, uint32_t b) { return deg[a] < deg[b]; });
  rank.assign(n, 0);
  for (int64_t i = 0; i < n; ++i) rank[ids[i]] = (ul)i;
  return FX_OK;
}

// adjacency of the graph renumbered by `rk`, neighbour lists sorted ascending
// (Dadjlist / Badjlist built with a ranker + sort_neighbours, adjlist.cu:62-73,83-87,127-151,158-187)
struct Adj {
  std::vector<ul> cd;        // n+1 (out) ; for both-sided: 2n+1 with in-lists shifted by n
  std::vector<uint32_t> adj;
};
static void build_adj(const fx_matrix* m, const std::vector<ul>& rk, bool both, Adj& g) {
  const int64_t n = m->n, nnz = m->nnz;
  const int64_t nodes = both ? 2 * n : n;
  g.cd.assign(nodes + 1, 0);
  for (int64_t r = 0; r < n; ++r) {
    g.cd[rk[r] + 1] += m->rowptr[r + 1] - m->rowptr[r];
    if (both)
      for (uint32_t e = m->rowptr[r]; e < m->rowptr[r + 1]; ++e) g.cd[rk[m->col[e]] + n + 1]++;
  }
  for (int64_t i = 0; i < nodes; ++i) g.cd[i + 1] += g.cd[i];
  g.adj.resize(both ? 2 * nnz : nnz);
  std::vector<ul> fill(g.cd.begin(), g.cd.end() - 1);
  for (int64_t r = 0; r < n; ++r) {
    const ul u = rk[r];
    for (uint32_t e = m->rowptr[r]; e < m->rowptr[r + 1]; ++e) {
      const ul v = rk[m->col[e]];
      g.adj[fill[u]++] = (uint32_t)v;
      if (both) g.adj[fill[v + n]++] = (uint32_t)u;
    }
  }
  for (int64_t i = 0; i < nodes; ++i) std::sort(g.adj.begin() + g.cd[i], g.adj.begin() + g.cd[i + 1]);
}

// order_rcm (order_rcm.cu:15-33, directed=true): degree ASC renumbering, BFS from node 0 with
// restarts at the next unplaced id (algo_bfs.cu:11-38), reversed and composed.
int order_rcm(const fx_matrix* m, std::vector<ul>& rank) {
  const int64_t n = m->n;
  std::vector<ul> rdeg;
  order_deg(m, false, rdeg);
  Adj g;
  build_adj(m, rdeg, false, g);
  std::vector<uint8_t> placed(n, 0);
  std::vector<uint32_t> order;
  order.reserve(n);
  size_t i = 0;
  for (int64_t c = 0; c < n; ++c) {
    if (placed[c]) continue;
    order.push_back((uint32_t)c);
    placed[c] = 1;
    while (i < order.size()) {
      const uint32_t w = order[i++];
      for (ul q = g.cd[w]; q < g.cd[w + 1]; ++q) {
        const uint32_t v = g.adj[q];
        if (placed[v]) continue;
        placed[v] = 1;
        order.push_back(v);
      }
    }
  }
  std::vector<ul> rbfs(n);
  for (int64_t p = 0; p < n; ++p) rbfs[order[p]] = (ul)p;  // rank_from_order tools.cu:31-43
  rank.assign(n, 0);
  for (int64_t u = 0; u < n; ++u) rank[u] = (ul)(n - 1) - rbfs[rdeg[u]];
  return FX_OK;
}

// ---- Gorder -------------------------------------------------------------------------------
namespace {
struct Heap {  // UnitHeap (unitheap.cuh:24-62): doubly linked list ordered by key with lazy updates
  std::vector<int> update, key;
  std::vector<ul> prev, next, first, second;
  size_t heapsize = 0;
  ul top = 0, huge = 0, none = 0;
  static constexpr int infty = INT_MAX / 2;
  explicit Heap(ul size) {
    none = size + 2;
    huge = (ul)std::sqrt((double)size);
    key.assign(size, infty); prev.assign(size, none); next.assign(size, none);
    update.assign(size, infty);
  }
  void insert(ul i, int k) { key[i] = k; update[i] = -k; heapsize++; }
  void reconstruct() {  // unitheap.cu:32-60 (indices 0..heapsize-1: valid because nothing is isolated)
    std::vector<ul> g(heapsize);
    std::iota(g.begin(), g.end(), (ul)0);
    std::stable_sort(g.begin(), g.end(), [&](ul a, ul b) { return key[a] > key[b]; });
    top = g[0];
    int cur = key[top];
    first.assign((size_t)10 * cur + 1, none);
    second.assign((size_t)10 * cur + 1, none);
    first[cur] = top;
    for (size_t i = 0; i < g.size(); ++i) {
      const ul v = g[i];
      prev[v] = i > 0 ? g[i - 1] : none;
      next[v] = i + 1 < g.size() ? g[i + 1] : none;
      const int k = key[v];
      if (k != cur) { second[cur] = g[i - 1]; first[k] = v; cur = k; }
    }
    second[cur] = g.back();
  }
  void erase_key(ul i, ul nx, ul pv) {
    const int k = key[i];
    if (first[k] == second[k]) first[k] = second[k] = none;
    else if (i == first[k]) first[k] = nx;
    else if (i == second[k]) second[k] = pv;
  }
  void del(ul i) {
    update[i] = infty;
    const ul pv = prev[i], nx = next[i];
    if (pv != none) next[pv] = nx;
    if (nx != none) prev[nx] = pv;
    erase_key(i, nx, pv);
    if (top == i) top = nx;
    prev[i] = next[i] = none;
    heapsize--;
  }
  void decrease_top() {  // unitheap.cu:87-133
    const ul nx = next[top];
    if (nx == none) return;
    const int k = key[top];
    const int leftover = update[top] / 2;
    const int nk = k + update[top] - leftover;
    if (nk >= key[nx]) return;
    update[top] = leftover;
    ul tail = second[k];
    ul nl = next[tail];
    while (nl != none && key[nl] >= nk) { tail = second[key[nl]]; nl = next[tail]; }
    prev[nx] = none;
    prev[top] = tail;
    next[top] = nl;
    next[tail] = top;
    if (nl != none) prev[nl] = top;
    erase_key(top, nx, none);
    key[top] = nk;
    second[nk] = top;
    if (first[nk] == none) first[nk] = top;
    top = nx;
  }
  ul extract_max() {
    ul t;
    do { t = top; if (update[top] < 0) decrease_top(); } while (top != t);
    del(top);
    return t;
  }
  void increment_key(ul i) {  // unitheap.cu:162-192
    const ul head = first[key[i]];
    const ul pv = prev[i], nx = next[i];
    if (head != i) {
      next[pv] = nx;
      if (nx != none) prev[nx] = pv;
      const ul pl = prev[head];
      prev[i] = pl; next[i] = head; prev[head] = i;
      if (pl != none) next[pl] = i;
    }
    erase_key(i, nx, pv);
    const int k = ++key[i];
    second[k] = i;
    if (first[k] == none) {
      first[k] = i;
      if (k > key[top]) top = i;
    }
    if (k + 4 >= (int)first.size()) {
      const size_t ns = (size_t)(first.size() * 1.5);
      first.resize(ns, none); second.resize(ns, none);
    }
  }
  void lazy(ul i, int up) {
    if (update[i] == infty) return;
    if (update[i] == 0 && up > 0) increment_key(i);
    else update[i] += up;
  }
};
}  // namespace

int order_gorder(const fx_matrix* m, int window, std::vector<ul>& rank) {
  const int64_t n = m->n;
  std::vector<ul> rrcm;
  order_rcm(m, rrcm);
  Adj g;
  build_adj(m, rrcm, true, g);
  auto deg_out = [&](ul u) { return g.cd[u + 1] - g.cd[u]; };
  auto deg_in = [&](ul u) { return g.cd[u + 1 + n] - g.cd[u + n]; };
  for (int64_t u = 0; u < n; ++u)
    if (deg_out(u) + deg_in(u) == 0) {
      set_error("Gorder: isolated vertex %lld (the reference's UnitHeap::ReConstruct, unitheap.cu:32-36, is only "
                "defined when every vertex is inserted)", (long long)u);
      return FX_ERR_FORMAT;
    }
  Heap heap((ul)n);
  for (int64_t u = 0; u < n; ++u) heap.insert((ul)u, (int)deg_in(u));
  heap.reconstruct();
  std::vector<ul> order;
  order.reserve(n);
  std::vector<ul> old_par, new_par;
  auto move_window = [&](ul nn, ul on) {  // order_gorder.cu:88-143
    const uint32_t* oi = g.adj.data() + g.cd[on + n];
    const uint32_t* oe = g.adj.data() + g.cd[on + n + 1];
    const uint32_t* ni = g.adj.data() + g.cd[nn + n];
    const uint32_t* ne = g.adj.data() + g.cd[nn + n + 1];
    if (on == nn) oi = oe;
    else if (deg_out(on) <= heap.huge)
      for (ul q = g.cd[on]; q < g.cd[on + 1]; ++q) heap.lazy(g.adj[q], -1);
    old_par.clear(); new_par.clear();
    while (true) {
      int factor = -1;
      if (oi >= oe) {
        if (ni >= ne) break;
        factor = 1;
      } else if (ni < ne) {
        if (*ni == *oi) { ++oi; ++ni; continue; }
        if (*ni < *oi) factor = 1;
      }
      if (factor == -1) { if (deg_out(*oi) <= heap.huge) old_par.push_back(*oi); ++oi; }
      else { if (deg_out(*ni) <= heap.huge) new_par.push_back(*ni); ++ni; }
    }
    for (ul par : old_par) {
      heap.lazy(par, -1);
      for (ul q = g.cd[par]; q < g.cd[par + 1]; ++q) if (g.adj[q] != on) heap.lazy(g.adj[q], -1);
    }
    if (deg_out(nn) <= heap.huge)
      for (ul q = g.cd[nn]; q < g.cd[nn + 1]; ++q) heap.lazy(g.adj[q], +1);
    for (ul par : new_par) {
      heap.lazy(par, +1);
      for (ul q = g.cd[par]; q < g.cd[par + 1]; ++q) if (g.adj[q] != nn) heap.lazy(g.adj[q], +1);
    }
  };
  const ul hub = heap.top;
  order.push_back(hub);
  heap.del(hub);
  move_window(hub, hub);
  while (heap.heapsize > 0) {
    const ul nn = heap.extract_max();
    order.push_back(nn);
    ul on = nn;
    if (order.size() > (size_t)window) on = order[order.size() - window - 1];
    move_window(nn, on);
  }
  std::vector<ul> rg(n);
  for (int64_t p = 0; p < n; ++p) rg[order[p]] = (ul)p;
  rank.assign(n, 0);
  for (int64_t u = 0; u < n; ++u) rank[u] = rg[rrcm[u]];  // order_gorder.cu:26-29
  return FX_OK;
}

// ---- DFS order (DataLoaderDFS, DataLoader.cu:324-385) ---------------------------------------
// Pre-order numbering of an iterative depth-first search from vertex 0, neighbours in column order,
// restarted at the next unvisited vertex.  rank[old] = new.
int order_dfs(const fx_matrix* m, std::vector<ul>& rank) {
  const int64_t n = m->n;
  rank.assign(n, 0);
  std::vector<uint8_t> seen(n, 0);
  std::vector<std::pair<uint32_t, uint32_t>> stack;  // (cursor, end) into col
  ul next = 0;
  for (int64_t root = 0; root < n; ++root) {
    if (seen[root]) continue;
    seen[root] = 1;
    rank[root] = next++;
    stack.push_back({m->rowptr[root], m->rowptr[root + 1]});
    while (!stack.empty()) {
      auto& it = stack.back();
      while (it.first < it.second && seen[m->col[it.first]]) ++it.first;
      if (it.first >= it.second) { stack.pop_back(); continue; }
      const uint32_t v = m->col[it.first++];
      seen[v] = 1;
      rank[v] = next++;
      stack.push_back({m->rowptr[v], m->rowptr[v + 1]});
    }
  }
  return FX_OK;
}

// ---- Rabbit order (DataLoaderRabbit, DataLoader.cu:455-655; Shiokawa'13 iterative variant) ----
// Incremental modularity merging (opt_iterative, no hub grouping, cluster shyness 1), dendrogram
// leaves in order.  vo_mp[new] = old is produced directly.  The per-round vertex order comes from
// an UNSTABLE sort by current degree in the reference (ranges::sort); the same libstdc++ std::sort
// is used here on the same sequence so that ties fall the same way.
int order_rabbit(const fx_matrix* m, bool is_directed, std::vector<int32_t>& vo_mp) {
  const int64_t n = m->n;
  struct Vtx {
    std::map<int, int> w;       // modularity weights to neighbouring clusters
    int left = -1, right = -1;  // dendrogram: cluster node = (left subtree root, right subtree root)
    int tree = -1;              // current root node id of this vertex's tree, -1 once merged away
    int deg = 0, round = 0;
  };
  std::vector<Vtx> g(n);
  // dendrogram nodes: ids [0,n) are leaves, id n+u is the cluster node created when u was merged
  std::vector<int> lch(2 * n, -1), rch(2 * n, -1);
  long long n_edges = 0;
  std::vector<uint32_t> cur(n), nxt;
  for (int64_t v = 0; v < n; ++v) {
    for (uint32_t e = m->rowptr[v]; e < m->rowptr[v + 1]; ++e) {
      const int d = (int)m->col[e];
      if (d != v) {
        g[v].w[d] = 1;
        if (is_directed) g[d].w[(int)v] = 1;  // force_undirected = dl.is_directed
      }
    }
    // as in the reference (DataLoader.cu:527), the degree is taken HERE: it counts the reverse edges
    // that lower-numbered vertices have already inserted, not those higher-numbered ones add later
    g[v].deg = (int)g[v].w.size();
    n_edges += g[v].deg;
    g[v].tree = (int)v;
    cur[v] = (uint32_t)v;
  }
  const double two_m_inv = 1.0 / double(2 * n_edges);
  for (int round = 1; !cur.empty(); ++round) {
    std::sort(cur.begin(), cur.end(), [&](uint32_t a, uint32_t b) { return g[a].deg < g[b].deg; });
    for (uint32_t u : cur) {
      Vtx& uo = g[u];
      if (uo.round == round) continue;
      double dq_max = -1;
      int v = -1;
      const double dv_2m = uo.deg * two_m_inv;
      for (auto& [d, w] : uo.w) {
        const double dq = w - g[d].deg * dv_2m;
        if (dq > dq_max) { dq_max = dq; v = d; }
      }
      if (dq_max <= 0) continue;
      Vtx& vo = g[v];
      vo.deg += uo.deg;
      for (auto& [d, w] : uo.w) {
        if (d == v) continue;
        vo.w[d] += w;
        auto& dw = g[d].w;
        auto it = dw.find((int)u);
        if (it == dw.end()) continue;
        dw[v] += it->second;
        dw.erase((int)u);
      }
      vo.w.erase((int)u);
      lch[n + u] = vo.tree;  // Tree_Node(vo.tree_node, uo.tree_node)
      rch[n + u] = uo.tree;
      uo.tree = -1;
      vo.tree = (int)(n + u);
      if (vo.round == round) continue;
      vo.round = round;
      nxt.push_back((uint32_t)v);
    }
    if (!(nxt.size() < cur.size())) { set_error("Rabbit: no progress (assert DataLoader.cu:585)"); return FX_ERR_FORMAT; }
    std::swap(cur, nxt);
    nxt.clear();
  }
  vo_mp.clear();
  vo_mp.reserve(n);
  std::vector<int> st;
  for (int64_t v = 0; v < n; ++v) {
    if (g[v].tree < 0) continue;
    st.push_back(g[v].tree);
    while (!st.empty()) {  // leaves left to right
      const int node = st.back();
      st.pop_back();
      if (node < n) { vo_mp.push_back(node); continue; }
      st.push_back(rch[node]);
      st.push_back(lch[node]);
    }
  }
  return FX_OK;
}

}  // namespace fx
