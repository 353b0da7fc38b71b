/*
 * fx_oracle_order.c -- TEST INFRASTRUCTURE ONLY (see fx_oracle.h).
 * C restatement of the reference's vertex orderings: order_deg.cu, order_rcm.cu, algo_bfs.cu,
 * adjlist.cu, order_gorder.cu, unitheap.cu (vendored there from lecfab/rescience-gorder).
 * rank[u] = new position of vertex u.  Written to follow the reference's structure step by step
 * (edge list in CSR order -> ranked adjacency lists -> BFS / unit heap); the writes the reference
 * makes through reserve()d-but-empty vectors (tools.cu:31-43, edgelist.cu:98-101, adjlist.cu:19-20)
 * are done here into properly sized arrays -- same values.
 */
#include <limits.h>
#include <math.h>
#include <stdlib.h>
#include <string.h>

#include "fx_oracle.h"

typedef uint64_t ul;

/* Edgelist::compute_degrees (edgelist.cu:89-104): degOut/degIn/deg over the CSR edges */
static void edge_degrees(int64_t n, const uint32_t *rowptr, const uint32_t *col, ul *dout, ul *din) {
  memset(dout, 0, sizeof(ul) * (size_t)n);
  memset(din, 0, sizeof(ul) * (size_t)n);
  for (int64_t u = 0; u < n; ++u)
    for (uint32_t e = rowptr[u]; e < rowptr[u + 1]; ++e) { dout[u]++; din[col[e]]++; }
}

typedef struct { ul key, val; } kv_t;
static int cmp_desc(const void *a, const void *b) { /* compare_nodedeg_desc order_deg.cu:8 */
  const kv_t *x = a, *y = b;
  if (x->val != y->val) return x->val > y->val ? -1 : 1;
  return x->key < y->key ? -1 : x->key > y->key;
}
static int cmp_asc(const void *a, const void *b) { /* compare_nodedeg_asc order_deg.cu:11 */
  const kv_t *x = a, *y = b;
  if (x->val != y->val) return x->val < y->val ? -1 : 1;
  return x->key < y->key ? -1 : x->key > y->key;
}

/* order_deg(h, desc) -> rank_from_deg (order_deg.cu:19-44) */
void orc_order_deg(int64_t n, const uint32_t *rowptr, const uint32_t *col, int desc, uint64_t *rank) {
  ul *dout = malloc(sizeof(ul) * (size_t)(n ? n : 1)), *din = malloc(sizeof(ul) * (size_t)(n ? n : 1));
  edge_degrees(n, rowptr, col, dout, din);
  kv_t *s = malloc(sizeof(kv_t) * (size_t)(n ? n : 1));
  for (int64_t u = 0; u < n; ++u) { s[u].key = (ul)u; s[u].val = dout[u] + din[u]; }
  qsort(s, (size_t)n, sizeof(kv_t), desc ? cmp_desc : cmp_asc);
  for (int64_t u = 0; u < n; ++u) rank[s[u].key] = (ul)u;
  free(s); free(dout); free(din);
}

static int cmp_ul(const void *a, const void *b) { ul x = *(const ul *)a, y = *(const ul *)b; return x < y ? -1 : x > y; }

/* Dadjlist(h, rank) / Badjlist(h, rank): build_from_edgelist_ranked (adjlist.cu:62-73):
 * degrees by ranked id, cumulated degrees, adjacency filled in edge order, neighbours sorted. */
typedef struct { int64_t n; int both; ul *cd; ul *adj; } adj_t;
static void adj_build(adj_t *g, int64_t n, const uint32_t *rowptr, const uint32_t *col, const ul *rk, int both) {
  int64_t nodes = both ? 2 * n : n, e = rowptr[n];
  g->n = n; g->both = both;
  g->cd = calloc((size_t)nodes + 1, sizeof(ul));
  g->adj = malloc(sizeof(ul) * (size_t)((both ? 2 * e : e) + 1));
  ul *deg = calloc((size_t)nodes + 1, sizeof(ul));
  for (int64_t u = 0; u < n; ++u)
    for (uint32_t q = rowptr[u]; q < rowptr[u + 1]; ++q) {
      deg[rk[u]]++;                      /* degOut (Dadjlist::compute_degrees adjlist.cu:137) */
      if (both) deg[rk[col[q]] + n]++;   /* degIn slid by n (Badjlist adjlist.cu:169) */
    }
  for (int64_t u = 0; u < nodes; ++u) { g->cd[u + 1] = g->cd[u] + deg[u]; deg[u] = 0; }
  for (int64_t u0 = 0; u0 < n; ++u0)
    for (uint32_t q = rowptr[u0]; q < rowptr[u0 + 1]; ++q) {
      ul u = rk[u0], v = rk[col[q]];
      g->adj[g->cd[u] + deg[u]++] = v;
      if (both) g->adj[g->cd[v + n] + deg[v + n]++] = u;
    }
  for (int64_t u = 0; u < nodes; ++u) qsort(g->adj + g->cd[u], (size_t)(g->cd[u + 1] - g->cd[u]), sizeof(ul), cmp_ul);
  free(deg);
}
static void adj_free(adj_t *g) { free(g->cd); free(g->adj); }

/* order_rcm(h, directed=true) (order_rcm.cu:15-33) with algo_bfs(g, 0) (algo_bfs.cu:11-38) */
void orc_order_rcm(int64_t n, const uint32_t *rowptr, const uint32_t *col, uint64_t *rank) {
  ul *rdeg = malloc(sizeof(ul) * (size_t)(n ? n : 1));
  orc_order_deg(n, rowptr, col, 0, rdeg); /* degree ASC */
  adj_t g;
  adj_build(&g, n, rowptr, col, rdeg, 0);
  char *placed = calloc((size_t)(n ? n : 1), 1);
  ul *order = malloc(sizeof(ul) * (size_t)(n ? n : 1));
  int64_t cnt = 0, i = 0;
  for (int64_t c = 0; c < n; ++c) {
    ul u = (ul)c; /* (c + u0) % n with u0 = 0 */
    if (placed[u]) continue;
    order[cnt++] = u; placed[u] = 1;
    while (i < cnt) {
      ul w = order[i++];
      for (ul q = g.cd[w]; q < g.cd[w + 1]; ++q) {
        ul v = g.adj[q];
        if (placed[v]) continue;
        placed[v] = 1; order[cnt++] = v;
      }
    }
  }
  ul *rbfs = malloc(sizeof(ul) * (size_t)(n ? n : 1));
  for (int64_t p = 0; p < n; ++p) rbfs[order[p]] = (ul)p; /* rank_from_order tools.cu:31-43 */
  for (int64_t u = 0; u < n; ++u) rank[u] = (ul)(n - 1) - rbfs[rdeg[u]];
  free(rbfs); free(order); free(placed); adj_free(&g); free(rdeg);
}

/* ---- UnitHeap (unitheap.cuh / unitheap.cu) ---- */
#define INFTY (INT_MAX / 2)
typedef struct { int key; ul prev, next; } lelem;
typedef struct { ul first, second; } hdr;
typedef struct {
  int *update; lelem *LL; hdr *H; size_t hsz; size_t heapsize; ul top, huge, none; int64_t size;
} uheap;
static void h_resize(uheap *h, size_t ns) {
  h->H = realloc(h->H, sizeof(hdr) * ns);
  for (size_t i = h->hsz; i < ns; ++i) { h->H[i].first = h->none; h->H[i].second = h->none; }
  h->hsz = ns;
}
static void h_init(uheap *h, ul size) { /* unitheap.cu:16-22 */
  memset(h, 0, sizeof(*h));
  h->size = (int64_t)size;
  h->none = size + 2;
  h->huge = (ul)sqrt((double)size);
  h->LL = malloc(sizeof(lelem) * (size_t)(size ? size : 1));
  h->update = malloc(sizeof(int) * (size_t)(size ? size : 1));
  for (ul i = 0; i < size; ++i) { h->LL[i].key = INFTY; h->LL[i].prev = h->LL[i].next = h->none; h->update[i] = INFTY; }
}
static void h_insert(uheap *h, ul i, int key) { h->LL[i].key = key; h->update[i] = -key; h->heapsize++; } /* :23-28 */
static uheap *g_sort_heap;
static int cmp_heap(const void *a, const void *b) { /* key DESC, id ASC (:37-39) */
  ul x = *(const ul *)a, y = *(const ul *)b;
  int kx = g_sort_heap->LL[x].key, ky = g_sort_heap->LL[y].key;
  if (kx != ky) return kx > ky ? -1 : 1;
  return x < y ? -1 : x > y;
}
static void h_reconstruct(uheap *h) { /* :32-60: indices 0..heapsize-1 */
  size_t m = h->heapsize;
  ul *g = malloc(sizeof(ul) * (m ? m : 1));
  for (size_t i = 0; i < m; ++i) g[i] = i;
  g_sort_heap = h;
  qsort(g, m, sizeof(ul), cmp_heap);
  h->top = g[0];
  int cur = h->LL[h->top].key;
  h_resize(h, (size_t)10 * cur + 1);
  h->H[cur].first = h->top;
  for (size_t i = 0; i < m; ++i) {
    ul v = g[i];
    h->LL[v].prev = i > 0 ? g[i - 1] : h->none;
    h->LL[v].next = i + 1 < m ? g[i + 1] : h->none;
    int key = h->LL[v].key;
    if (key != cur) { h->H[cur].second = g[i - 1]; h->H[key].first = v; cur = key; }
  }
  h->H[cur].second = g[m - 1];
  free(g);
}
static void h_erase_key(uheap *h, ul i, ul next, ul prev) { /* :63-71 */
  int key = h->LL[i].key;
  if (h->H[key].first == h->H[key].second) h->H[key].first = h->H[key].second = h->none;
  else if (i == h->H[key].first) h->H[key].first = next;
  else if (i == h->H[key].second) h->H[key].second = prev;
}
static void h_delete(uheap *h, ul i) { /* :136-150 */
  h->update[i] = INFTY;
  ul prev = h->LL[i].prev, next = h->LL[i].next;
  if (prev != h->none) h->LL[prev].next = next;
  if (next != h->none) h->LL[next].prev = prev;
  h_erase_key(h, i, next, prev);
  if (h->top == i) h->top = next;
  h->LL[i].prev = h->LL[i].next = h->none;
  h->heapsize--;
}
static void h_decrease_top(uheap *h) { /* :87-133 */
  ul top = h->top, next = h->LL[top].next;
  if (next == h->none) return;
  int key = h->LL[top].key;
  int leftover = h->update[top] / 2;
  int new_key = key + h->update[top] - leftover;
  if (new_key >= h->LL[next].key) return;
  h->update[top] = leftover;
  ul level_tail = h->H[key].second;
  ul next_level = h->LL[level_tail].next;
  while (next_level != h->none && h->LL[next_level].key >= new_key) {
    level_tail = h->H[h->LL[next_level].key].second;
    next_level = h->LL[level_tail].next;
  }
  h->LL[next].prev = h->none;
  h->LL[top].prev = level_tail;
  h->LL[top].next = next_level;
  h->LL[level_tail].next = top;
  if (next_level != h->none) h->LL[next_level].prev = top;
  h_erase_key(h, top, next, h->none);
  h->LL[top].key = new_key;
  h->H[new_key].second = top;
  if (h->H[new_key].first == h->none) h->H[new_key].first = top;
  h->top = next;
}
static ul h_extract_max(uheap *h) { /* :74-84 */
  ul tmptop;
  do { tmptop = h->top; if (h->update[h->top] < 0) h_decrease_top(h); } while (h->top != tmptop);
  h_delete(h, h->top);
  return tmptop;
}
static void h_increment_key(uheap *h, ul i) { /* :162-192 */
  ul level_head = h->H[h->LL[i].key].first;
  ul prev = h->LL[i].prev, next = h->LL[i].next;
  if (level_head != i) {
    h->LL[prev].next = next;
    if (next != h->none) h->LL[next].prev = prev;
    ul prev_level = h->LL[level_head].prev;
    h->LL[i].prev = prev_level;
    h->LL[i].next = level_head;
    h->LL[level_head].prev = i;
    if (prev_level != h->none) h->LL[prev_level].next = i;
  }
  h_erase_key(h, i, next, prev);
  int key = ++h->LL[i].key;
  h->H[key].second = i;
  if (h->H[key].first == h->none) {
    h->H[key].first = i;
    if (key > h->LL[h->top].key) h->top = i;
  }
  if (key + 4 >= (int)h->hsz) h_resize(h, (size_t)(h->hsz * 1.5));
}
static void h_lazy(uheap *h, ul i, int up) { /* :153-160 */
  if (h->update[i] == INFTY) return;
  if (h->update[i] == 0 && up > 0) h_increment_key(h, i);
  else h->update[i] += up;
}

/* complete_gorder(h, window) (order_gorder.cu:13-31) = RCM, Badjlist, order_gorder (:35-84) with
 * move_window (:88-143); all locality weights are 1 (order_gorder.cuh:20-28).
 * Returns -1 if some vertex is isolated (UnitHeap::ReConstruct is then ill-defined). */
int orc_order_gorder(int64_t n, const uint32_t *rowptr, const uint32_t *col, int window, uint64_t *rank) {
  ul *rrcm = malloc(sizeof(ul) * (size_t)(n ? n : 1));
  orc_order_rcm(n, rowptr, col, rrcm);
  adj_t g;
  adj_build(&g, n, rowptr, col, rrcm, 1);
#define DEGOUT(u) (g.cd[(u) + 1] - g.cd[(u)])
#define DEGIN(u) (g.cd[(u) + 1 + n] - g.cd[(u) + n])
  for (int64_t u = 0; u < n; ++u)
    if (DEGOUT(u) + DEGIN(u) == 0) { adj_free(&g); free(rrcm); return -1; }
  uheap hp;
  h_init(&hp, (ul)n);
  for (int64_t u = 0; u < n; ++u) h_insert(&hp, (ul)u, (int)DEGIN(u));
  h_reconstruct(&hp);
  ul *order = malloc(sizeof(ul) * (size_t)(n ? n : 1));
  int64_t cnt = 0;
  ul *tmp_old = malloc(sizeof(ul) * (size_t)(n + 1)), *tmp_new = malloc(sizeof(ul) * (size_t)(n + 1));
  ul new_node = hp.top, old_node = new_node;
  order[cnt++] = new_node;
  h_delete(&hp, new_node);
  for (;;) {
    /* move_window(g, heap, new_node, old_node) */
    const ul *oi = g.adj + g.cd[old_node + n], *oe = g.adj + g.cd[old_node + n + 1];
    const ul *ni = g.adj + g.cd[new_node + n], *ne = g.adj + g.cd[new_node + n + 1];
    if (old_node == new_node) oi = oe;
    else if (DEGOUT(old_node) <= hp.huge)
      for (ul q = g.cd[old_node]; q < g.cd[old_node + 1]; ++q) h_lazy(&hp, g.adj[q], -1);
    int64_t no = 0, nn = 0;
    for (;;) {
      int factor = -1;
      if (oi >= oe) { if (ni >= ne) break; factor = 1; }
      else if (ni < ne) {
        if (*ni == *oi) { ++oi; ++ni; continue; }
        if (*ni < *oi) factor = 1;
      }
      if (factor == -1) { if (DEGOUT(*oi) <= hp.huge) tmp_old[no++] = *oi; ++oi; }
      else { if (DEGOUT(*ni) <= hp.huge) tmp_new[nn++] = *ni; ++ni; }
    }
    for (int64_t t = 0; t < no; ++t) {
      ul par = tmp_old[t];
      h_lazy(&hp, par, -1);
      for (ul q = g.cd[par]; q < g.cd[par + 1]; ++q) if (g.adj[q] != old_node) h_lazy(&hp, g.adj[q], -1);
    }
    if (DEGOUT(new_node) <= hp.huge)
      for (ul q = g.cd[new_node]; q < g.cd[new_node + 1]; ++q) h_lazy(&hp, g.adj[q], +1);
    for (int64_t t = 0; t < nn; ++t) {
      ul par = tmp_new[t];
      h_lazy(&hp, par, +1);
      for (ul q = g.cd[par]; q < g.cd[par + 1]; ++q) if (g.adj[q] != new_node) h_lazy(&hp, g.adj[q], +1);
    }
    if (hp.heapsize == 0) break;
    new_node = h_extract_max(&hp);
    order[cnt++] = new_node;
    old_node = new_node;
    if (cnt > window) old_node = order[cnt - window - 1];
  }
  ul *rg = malloc(sizeof(ul) * (size_t)(n ? n : 1));
  for (int64_t p = 0; p < n; ++p) rg[order[p]] = (ul)p;
  for (int64_t u = 0; u < n; ++u) rank[u] = rg[rrcm[u]];
  free(rg); free(order); free(tmp_old); free(tmp_new);
  free(hp.LL); free(hp.update); free(hp.H);
  adj_free(&g); free(rrcm);
  return 0;
}
