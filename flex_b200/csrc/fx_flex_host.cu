// fx_flex_host.cu -- host part of the pillar format: Mat::csr2_DiagTiling (mat.cu:680-903).
//
// Round 1 grows each diagonal block row by row from where the previous block ended, so the blocks
// form one serial chain along the diagonal; the reference builds the whole format on the host and so
// does this file -- with flat arrays instead of unordered_map<int,unordered_set<int>> per row
// (claims are one byte per nz; "columns listed by an SM" is a per-column list of SM ids), which is
// what makes it usable on the 10^8-nz shapes.  Outputs are bit-identical to the reference's.
#include <algorithm>
#include <cmath>

#include "fx_common.cuh"
#include "fx_flex.cuh"

namespace fx {

namespace {
struct ColSm {  // alpha_columns_per_sm (mat.cu:712): which SMs listed column c
  std::vector<int> head;
  struct Node { int sm, next; };
  std::vector<Node> nodes;
  explicit ColSm(int n) : head(n, -1) {}
  bool has(unsigned c, int sm) const {
    for (int q = head[c]; q >= 0; q = nodes[q].next) if (nodes[q].sm == sm) return true;
    return false;
  }
  void add(unsigned c, int sm) {
    if (has(c, sm)) return;
    nodes.push_back({sm, head[c]});
    head[c] = (int)nodes.size() - 1;
  }
};

int find_col(const fx_matrix* m, int r, unsigned c) {
  auto b = m->col.begin() + m->rowptr[r], e = m->col.begin() + m->rowptr[r + 1];
  auto it = std::lower_bound(b, e, c);
  return (it != e && *it == c) ? (int)(it - m->col.begin()) : -1;
}

// one row panel of csr2seg_Cmajor (mat.cu:1192-1269) over the nz not yet claimed
int seg_panel(const fx_matrix* m, int tm, int nnz_limit, int ridx, std::vector<uint8_t>& claimed,
              fx_flex_dev::PillarHost& o) {
  const int M = (int)m->n;
  const int rowStart = ridx * tm, rowEnd = std::min(M, rowStart + tm), rows = rowEnd - rowStart;
  const int dif = (int)(0.1 * nnz_limit);
  std::vector<std::vector<unsigned>> ent(rows);
  int remaining = 0;
  for (int i = 0; i < rows; ++i)
    for (unsigned e = m->rowptr[rowStart + i]; e < m->rowptr[rowStart + i + 1]; ++e)
      if (!claimed[e]) { ent[i].push_back(e); ++remaining; }
  std::vector<size_t> cur(rows, 0), prev(rows, 0);
  std::vector<int> atom(rows, 0);
  int nnzInSeg = 0, tiles = 0;
  while (remaining > 0 || nnzInSeg > 0) {
    if (remaining > 0) {
      unsigned j = 0xffffffffu;
      for (int i = 0; i < rows; ++i) if (cur[i] < ent[i].size()) j = std::min(j, m->col[ent[i][cur[i]]]);
      for (int i = 0; i < rows; ++i)
        if (cur[i] < ent[i].size() && m->col[ent[i][cur[i]]] == j) {
          claimed[ent[i][cur[i]]] = 1;
          ++cur[i]; ++atom[i]; ++nnzInSeg; --remaining;
        }
    }
    if ((remaining == 0 && nnzInSeg) || (nnz_limit - nnzInSeg) <= dif || nnzInSeg > nnz_limit) {
      for (int i = 0; i < rows; ++i) {
        o.alpha_rowPtr.push_back(o.alpha_rowPtr.back() + (unsigned)(cur[i] - prev[i]));
        for (size_t q = prev[i]; q < cur[i]; ++q) { o.alpha_colIdx.push_back(m->col[ent[i][q]]); o.alpha_vals.push_back(m->val[ent[i][q]]); }
        nnzInSeg -= (int)(cur[i] - prev[i]);
        prev[i] = cur[i];
      }
      for (int i = 0; i < rows; ++i) {
        const int rl = (int)(m->rowptr[rowStart + i + 1] - m->rowptr[rowStart + i]);
        const unsigned v = (unsigned)m->vo_mp[rowStart + i];
        o.segVoMap.push_back(atom[i] < rl ? (v | 0x80000000u) : v);
        atom[i] = 0;
      }
      o.alpha_pillar_rowPtr.push_back(o.alpha_pillar_rowPtr.back() + (unsigned)rows);
      ++tiles;
    }
  }
  return tiles;
}
}  // namespace

int diag_tiling_host(const fx_matrix* m, int tm, int n_sm, fx_flex_dev::PillarHost& o) {
  FX_REQUIRE(!m->col.empty(), FX_ERR_UNSUPPORTED, "pillar format needs the host CSR");
  const int M = (int)m->n, nnz = (int)m->nnz;
  const int warps_per_sm = 64;   // mat.cu:688
  const float alpha = 0.3f;      // mat.cu:690
  o = fx_flex_dev::PillarHost();
  const int nnz_diagonal_tiles = (int)(alpha * m->rowptr[M]);
  const int partitions_node = warps_per_sm * n_sm;
  const int nnz_p_diagonal_tile = std::max(32, nnz_diagonal_tiles / partitions_node);
  const int thr = (int)(0.85 * nnz_p_diagonal_tile);
  std::vector<int> tile_width(partitions_node, 0);
  std::vector<uint8_t> claimed(std::max(nnz, 1), 0);
  ColSm colsm(M);
  // round 1 (mat.cu:706-759)
  int mat_r_start = 0, warps_with_weights = 0;
  for (int i = 0; i < partitions_node; ++i) {
    mat_r_start += i ? tile_width[i - 1] : 0;
    int cnt = 0, j = mat_r_start;
    const int sm = i / warps_per_sm;
    while (j < M && cnt <= thr) {
      for (unsigned kk = m->rowptr[j];; ++kk) {
        // the reference walks until it meets column j and does not stop at the end of the row
        // (mat.cu:718-727); running off the array is undefined there and refused here
        FX_REQUIRE(kk < (unsigned)nnz, FX_ERR_FORMAT,
                   "row %d has no diagonal entry and the reference's diagonal walk (mat.cu:718-727) leaves the matrix", j);
        if (!(m->col[kk] <= (unsigned)j)) break;
        if ((int)m->col[kk] >= mat_r_start) {
          ++cnt;
          colsm.add(m->col[kk], sm);
          const int e = kk < m->rowptr[j + 1] ? (int)kk : find_col(m, j, m->col[kk]);
          if (e >= 0) claimed[e] = 1;
        }
        if (m->col[kk] == (unsigned)j) break;
      }
      for (int kk = mat_r_start; kk < j; ++kk) {
        const int l = find_col(m, kk, (unsigned)j);
        if (l >= 0) { ++cnt; claimed[l] = 1; colsm.add((unsigned)j, sm); }
      }
      ++j;
    }
    warps_with_weights += cnt > 0;
    FX_REQUIRE(j >= M || cnt, FX_ERR_FORMAT, "empty diagonal block (assert mat.cu:751)");
    tile_width[i] = j - mat_r_start;
  }
  long verify_m = 0;
  for (int w : tile_width) verify_m += w;
  FX_REQUIRE(verify_m == M, FX_ERR_FORMAT, "alpha is too small: diagonal blocks cover %ld of %d rows (assert mat.cu:759)", verify_m, M);
  // round 2 (mat.cu:771-835)
  o.alpha_pillar_rowPtr.push_back(0);
  int nnz_rowPtr = 0, row_end = 0, col_start = 0, col_end = 0;
  for (int i = 0; i < warps_with_weights; ++i) {
    const int row_start = row_end;
    row_end += tile_width[i];
    if (i % warps_per_sm == 0) {
      col_start = col_end;
      for (int idx = 0; idx < warps_per_sm && (i + idx) < warps_with_weights; ++idx) col_end += tile_width[i + idx];
    }
    const int sm = i / warps_per_sm;
    int nnz_warp = 0;
    for (int j = row_start; j < row_end; ++j) {
      int entries = 0;
      o.alpha_rowPtr.push_back((unsigned)nnz_rowPtr);
      for (unsigned kk = m->rowptr[j]; kk < m->rowptr[j + 1]; ++kk) {
        const int l = (int)m->col[kk];
        if (l < col_start) continue;
        if (l >= col_end) break;
        const bool insm = colsm.has((unsigned)l, sm);
        if (claimed[kk] || insm) {
          claimed[kk] = 1;
          FX_REQUIRE(insm, FX_ERR_FORMAT, "claimed nz outside the SM's column set (assert mat.cu:818)");
          o.alpha_colIdx.push_back((unsigned)l);
          o.alpha_vals.push_back(m->val[kk]);
          ++entries; ++nnz_warp; ++nnz_rowPtr;
        }
      }
      const unsigned v = (unsigned)m->vo_mp[j];
      o.segVoMap.push_back(entries < (int)(m->rowptr[j + 1] - m->rowptr[j]) ? (v | 0x80000000u) : v);
    }
    FX_REQUIRE(nnz_warp, FX_ERR_FORMAT, "pillar without nz (assert mat.cu:837)");
    o.alpha_pillar_rowPtr.push_back(o.alpha_pillar_rowPtr.back() + (unsigned)tile_width[i]);
    if (i % warps_per_sm == 0) o.alpha_pillarIdx.push_back((unsigned)i);
  }
  while ((int)o.alpha_pillarIdx.size() <= n_sm) o.alpha_pillarIdx.push_back((unsigned)warps_with_weights);
  FX_REQUIRE((int)o.alpha_rowPtr.size() == M, FX_ERR_FORMAT, "diagonal blocks do not cover every row (assert mat.cu:853)");
  o.empty_wp_p = (1 - (float)warps_with_weights / partitions_node) * 100;
  o.band_nz_p = (float)o.alpha_colIdx.size() / m->rowptr[M] * 100;
  FX_REQUIRE(partitions_node != warps_with_weights, FX_ERR_FORMAT, "no idle warp left for the balance queue (division by zero at mat.cu:862)");
  // round 3 (mat.cu:871-878): leftovers, row-panel segmentation, into the shared balance queue
  o.alpha_rowPtr.push_back((unsigned)nnz_rowPtr);
  int tiles_in_total = warps_with_weights;
  const int tileRows = (M + tm - 1) / tm;
  for (int p = 0; p < tileRows; ++p) tiles_in_total += seg_panel(m, tm, 128, p, claimed, o);
  o.alpha_pillarIdx.push_back((unsigned)tiles_in_total);
  o.n_segs = tiles_in_total;
  o.warps_with_weights = warps_with_weights;
  FX_REQUIRE((int)o.alpha_colIdx.size() == nnz, FX_ERR_FORMAT, "pillar format lost nz (assert mat.cu:895)");
  return FX_OK;
}

}  // namespace fx
