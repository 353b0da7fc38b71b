// fx_reorder.cu -- L1 reordering hooks: DataLoaderDeg / DataLoaderRcm / DataLoaderGorder
// (DataLoader.cu:658-857) and DataLoader::perm_apply (DataLoader.cu:244-321).
//
// The rank computations are host C++ (the reference's are too, and Gorder is inherently serial,
// order_gorder.cu:35-84) and live in fx_order.cc-style functions below; the permutation itself is
// applied with counting (no per-row vector<pair> + sort): because rank is a bijection, visiting the
// NEW columns in ascending order and appending to the destination rows yields rows that are already
// column-sorted -- one O(nnz) pass over the transposed structure.
#include <algorithm>
#include <numeric>

#include "fx_common.cuh"

namespace fx {
int finish_matrix(fx_matrix* m, const std::string& name, int order, bool do_upload);
int order_deg(const fx_matrix* m, bool desc, std::vector<uint64_t>& rank);
int order_rcm(const fx_matrix* m, std::vector<uint64_t>& rank);
int order_gorder(const fx_matrix* m, int window, std::vector<uint64_t>& rank);
int order_dfs(const fx_matrix* m, std::vector<uint64_t>& rank);
int order_rabbit(const fx_matrix* m, bool is_directed, std::vector<int32_t>& vo_mp);
int ensure_census(const fx_matrix* m);

// rank[old] = new.  Produces vo_mp[new]=old and the permuted CSR with ascending columns.
int perm_apply(const fx_matrix* src, const uint64_t* rank, fx_matrix* dst) {
  const int64_t n = src->n, nnz = src->nnz;
  dst->n = n; dst->nnz = nnz; dst->k = src->k;
  dst->vo_mp.assign(n, -1);
  for (int64_t o = 0; o < n; ++o) {
    if (rank[o] >= (uint64_t)n || dst->vo_mp[rank[o]] != -1) {
      set_error("rank is not a permutation (vertex %lld -> %llu)", (long long)o, (unsigned long long)rank[o]);
      return FX_ERR_ARG;  // reference: assert(vold_to_new[v_old]==n) DataLoader.cu:256
    }
    dst->vo_mp[rank[o]] = (int32_t)o;
  }
  dst->rowptr.assign(n + 1, 0);
  for (int64_t vn = 0; vn < n; ++vn) {
    const int64_t vo = dst->vo_mp[vn];
    dst->rowptr[vn + 1] = dst->rowptr[vn] + (src->rowptr[vo + 1] - src->rowptr[vo]);
  }
  dst->col.resize(nnz);
  dst->val.resize(nnz);
  // transpose-style pass: bucket the edges by NEW column, then emit columns in ascending order
  std::vector<uint32_t> cptr(n + 2, 0);
  for (int64_t e = 0; e < nnz; ++e) cptr[rank[src->col[e]] + 1]++;
  for (int64_t i = 0; i < n; ++i) cptr[i + 1] += cptr[i];
  std::vector<uint32_t> t_row(nnz);
  std::vector<float> t_val(nnz);
  {
    std::vector<uint32_t> cur(cptr.begin(), cptr.begin() + n + 1);
    for (int64_t vo = 0; vo < n; ++vo) {
      const uint32_t vn = (uint32_t)rank[vo];
      for (uint32_t e = src->rowptr[vo]; e < src->rowptr[vo + 1]; ++e) {
        const uint32_t cn = (uint32_t)rank[src->col[e]];
        t_row[cur[cn]] = vn;
        t_val[cur[cn]++] = src->val[e];
      }
    }
  }
  std::vector<uint32_t> cur(dst->rowptr.begin(), dst->rowptr.begin() + n);
  for (int64_t cn = 0; cn < n; ++cn)
    for (uint32_t q = cptr[cn]; q < cptr[cn + 1]; ++q) {
      const uint32_t vn = t_row[q];
      dst->col[cur[vn]] = (uint32_t)cn;
      dst->val[cur[vn]++] = t_val[q];
    }
  return FX_OK;
}
}  // namespace fx

extern "C" int fx_reorder_with_rank(const fx_matrix* m, const uint64_t* rank, int order_tag, fx_matrix** out) {
  FX_REQUIRE(m && rank && out, FX_ERR_ARG, "fx_reorder_with_rank: null");
  FX_REQUIRE(!m->col.empty() || m->nnz == 0, FX_ERR_UNSUPPORTED, "matrix has no host CSR (created from device arrays)");
  auto d = new fx_matrix();
  int rc = fx::perm_apply(m, rank, d);
  if (rc == FX_OK) rc = fx::finish_matrix(d, std::string(m->info.graph_name) + ".csv", order_tag, false);
  if (rc != FX_OK) { fx_matrix_free(d); return rc; }
  d->info.c = m->info.c;
  *out = d;
  return FX_OK;
}

extern "C" int fx_reorder(const fx_matrix* m, int order, fx_matrix** out) {
  FX_REQUIRE(m && out, FX_ERR_ARG, "fx_reorder: null");
  FX_REQUIRE(!m->col.empty() || m->nnz == 0, FX_ERR_UNSUPPORTED, "matrix has no host CSR (created from device arrays)");
  std::vector<uint64_t> rank;
  int rc = FX_OK;
  switch (order) {
    case FX_ORDER_OVO: rank.resize(m->n); std::iota(rank.begin(), rank.end(), 0); break;
    case FX_ORDER_DEG: rc = fx::order_deg(m, true, rank); break;   // DataLoader.cu:672 order_deg(h,true)
    case FX_ORDER_RCM: rc = fx::order_rcm(m, rank); break;         // DataLoader.cu:737
    case FX_ORDER_GOR: rc = fx::order_gorder(m, 3, rank); break;   // DataLoader.cu:803 window=3
    case FX_ORDER_DFS: rc = fx::order_dfs(m, rank); break;         // DataLoader.cu:324-385
    case FX_ORDER_RBT: {                                           // DataLoader.cu:455-655
      std::vector<int32_t> vo;
      fx::ensure_census(m);
      rc = fx::order_rabbit(m, m->info.is_directed != 0, vo);
      if (rc == FX_OK) {
        rank.resize(m->n);
        for (int64_t i = 0; i < m->n; ++i) rank[vo[i]] = (uint64_t)i;
      }
      break;
    }
    default: fx::set_error("unknown order %d", order); return FX_ERR_ARG;
  }
  if (rc != FX_OK) return rc;
  return fx_reorder_with_rank(m, rank.data(), order, out);
}
